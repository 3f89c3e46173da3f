// Op-level equivariance corrections on their own (fp32): the reference exposes them as free functions
// (src/models/utils.py: get_rotation_2D_matrix :271-298, rotate_encoding :301-321, translate_encodings :325-346,
// translate_encodings2 :349-364) which SimCLR / Hybrid2Model variants and user code call directly.  The
// training step does NOT use these kernels -- it runs the whole chain in one launch (csrc/ntxent.cu); these are
// the same arithmetic behind the reference's function signatures.
//
// Layout: encodings are [n][m][d] fp32 contiguous, d >= 2; only the first two coordinates of every point are
// touched (the reference passes d == 2).  One warp per sample; a sample's m points are walked by the lanes.
#include <math.h>

#include "../../include/peclr_b200.h"
#include "ptx.cuh"

namespace peclr {

constexpr int kWarpsPerBlock = 8;

// mode 0: x += tx * (max_x - min_x), y += ty * (max_y - min_y)   (translate_encodings: range-scaled, detached range)
// mode 1: x += tx, y += ty                                        (translate_encodings2: exact)
__global__ void __launch_bounds__(kWarpsPerBlock * 32) translate_encodings_kernel(float* enc, const float* tx,
                                                                                  const float* ty, int n, int m, int d,
                                                                                  int mode) {
  const int lane = threadIdx.x & 31;
  const int i = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  if (i >= n) return;
  float* e = enc + (size_t)i * m * d;
  float dx = tx[i], dy = ty[i];
  if (mode == 0) {
    float lox = INFINITY, hix = -INFINITY, loy = INFINITY, hiy = -INFINITY;
    for (int k = lane; k < m; k += 32) {
      const float x = e[(size_t)k * d], y = e[(size_t)k * d + 1];
      lox = fminf(lox, x), hix = fmaxf(hix, x), loy = fminf(loy, y), hiy = fmaxf(hiy, y);
    }
    dx *= warp_max(hix) - warp_min(lox);
    dy *= warp_max(hiy) - warp_min(loy);
  }
  for (int k = lane; k < m; k += 32) {
    e[(size_t)k * d] += dx;
    e[(size_t)k * d + 1] += dy;
  }
}

// rot[i] = {alpha, beta, off_x, off_y} of cv2-style getRotationMatrix2D about the (detached) centre of the sample's
// points: trig and offsets in fp64, rounded into fp32 (the reference writes them into a default-dtype matrix).
// out_x = alpha x + beta y + off_x ; out_y = -beta x + alpha y + off_y.
__global__ void __launch_bounds__(kWarpsPerBlock * 32) rotate_encoding_kernel(float* enc, const double* angle,
                                                                              float* rot, int n, int m, int d) {
  const int lane = threadIdx.x & 31;
  const int i = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  if (i >= n) return;
  float* e = enc + (size_t)i * m * d;
  float sx = 0.f, sy = 0.f;
  for (int k = lane; k < m; k += 32) sx += e[(size_t)k * d], sy += e[(size_t)k * d + 1];
  const float cx = warp_sum(sx) / (float)m, cy = warp_sum(sy) / (float)m;
  const double ang = angle[i] * 3.141592653589793 / 180.0;
  const double ca = cos(ang), sa = sin(ang);
  const float al = (float)ca, be = (float)sa;
  const float offx = (float)((1.0 - ca) * (double)cx - sa * (double)cy);
  const float offy = (float)((1.0 - ca) * (double)cy + sa * (double)cx);
  for (int k = lane; k < m; k += 32) {
    const float x = e[(size_t)k * d], y = e[(size_t)k * d + 1];
    e[(size_t)k * d] = fmaf(al, x, fmaf(be, y, offx));
    e[(size_t)k * d + 1] = fmaf(-be, x, fmaf(al, y, offy));
  }
  if (lane == 0 && rot) reinterpret_cast<float4*>(rot)[i] = make_float4(al, be, offx, offy);
}

// gradient through the 2x2 block of the rotation (the centre is detached): in place on g [n][m][d]
__global__ void __launch_bounds__(kWarpsPerBlock * 32) rotate_encoding_bwd_kernel(float* g, const float* rot, int n,
                                                                                  int m, int d) {
  const int lane = threadIdx.x & 31;
  const int i = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  if (i >= n) return;
  const float4 r = reinterpret_cast<const float4*>(rot)[i];
  float* e = g + (size_t)i * m * d;
  for (int k = lane; k < m; k += 32) {
    const float gx = e[(size_t)k * d], gy = e[(size_t)k * d + 1];
    e[(size_t)k * d] = r.x * gx - r.y * gy;
    e[(size_t)k * d + 1] = r.y * gx + r.x * gy;
  }
}

// (n,3,2) fp32 matrix of get_rotation_2D_matrix from fp64 angles (degrees) and fp32 centres; scale as given
__global__ void rotation_matrix_kernel(const double* angle, const float* cx, const float* cy, double scale, float* out,
                                       int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double ang = angle[i] * 3.141592653589793 / 180.0;
  const double al = scale * cos(ang), be = scale * sin(ang);
  float* o = out + (size_t)i * 6;  // [3][2]
  o[0] = (float)al, o[1] = (float)-be;
  o[2] = (float)be, o[3] = (float)al;
  o[4] = (float)((1.0 - al) * (double)cx[i] - be * (double)cy[i]);
  o[5] = (float)((1.0 - al) * (double)cy[i] + be * (double)cx[i]);
}

// get_projection_stats (hybrid2_model.py:92-106): per sample mean / lower median / min / max over the m points of
// coordinates 0 and 1, averaged over the batch.  out[8] = x{mean, median, min, max}, y{...}.  ONE block: its warps
// walk the samples and add their results in a fixed order (reproducible; the op is tiny and logging-only).
__global__ void __launch_bounds__(kWarpsPerBlock * 32) projection_stats_kernel(const float* enc, float* out, int n,
                                                                               int m, int d) {
  __shared__ float part[kWarpsPerBlock][8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float mine = 0.f;  // lane j < 8 accumulates statistic j over this warp's samples
  for (int i = warp; i < n; i += kWarpsPerBlock) {
    const float* e = enc + (size_t)i * m * d;
    const int target = (m - 1) / 2;  // torch.median returns the lower of the two middle values
    float res[8];
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      float sum = 0.f, lo = INFINITY, hi = -INFINITY, med = 0.f;
      for (int k = lane; k < m; k += 32) {
        const float v = e[(size_t)k * d + c];
        sum += v, lo = fminf(lo, v), hi = fmaxf(hi, v);
        int rank = 0;  // position of v in the sorted order, ties broken by index
        for (int j = 0; j < m; ++j) {
          const float o = e[(size_t)j * d + c];
          rank += (o < v) || (o == v && j < k);
        }
        if (rank == target) med = v;
      }
      res[4 * c + 0] = warp_sum(sum) / (float)m;
      res[4 * c + 1] = warp_sum(med);  // exactly one element has the target rank
      res[4 * c + 2] = warp_min(lo);
      res[4 * c + 3] = warp_max(hi);
    }
    float v = res[0];
#pragma unroll
    for (int j = 1; j < 8; ++j) v = lane == j ? res[j] : v;
    mine += v;
  }
  if (lane < 8) part[warp][lane] = mine;
  __syncthreads();
  if (threadIdx.x < 8) {
    float t = 0.f;
    for (int w = 0; w < kWarpsPerBlock; ++w) t += part[w][threadIdx.x];
    out[threadIdx.x] = t / (float)n;
  }
}

}  // namespace peclr

using namespace peclr;

static int check_launch() {
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : -(int)e;
}

extern "C" int peclr_translate_encodings(float* enc, const float* tx, const float* ty, int n, int m, int d, int exact,
                                         void* stream) {
  if (n == 0) return 0;  // empty batch: nothing to do (null pointers are fine then)
  if (!enc || !tx || !ty || n < 0 || m < 1 || d < 2) return -1001;
  translate_encodings_kernel<<<(n + kWarpsPerBlock - 1) / kWarpsPerBlock, kWarpsPerBlock * 32, 0,
                               static_cast<cudaStream_t>(stream)>>>(enc, tx, ty, n, m, d, exact ? 1 : 0);
  return check_launch();
}

extern "C" int peclr_rotate_encoding(float* enc, const double* angle, float* rot, int n, int m, int d, void* stream) {
  if (n == 0) return 0;
  if (!enc || !angle || n < 0 || m < 1 || d < 2) return -1001;
  rotate_encoding_kernel<<<(n + kWarpsPerBlock - 1) / kWarpsPerBlock, kWarpsPerBlock * 32, 0,
                           static_cast<cudaStream_t>(stream)>>>(enc, angle, rot, n, m, d);
  return check_launch();
}

extern "C" int peclr_rotate_encoding_bwd(float* g, const float* rot, int n, int m, int d, void* stream) {
  if (n == 0) return 0;
  if (!g || !rot || n < 0 || m < 1 || d < 2) return -1001;
  rotate_encoding_bwd_kernel<<<(n + kWarpsPerBlock - 1) / kWarpsPerBlock, kWarpsPerBlock * 32, 0,
                               static_cast<cudaStream_t>(stream)>>>(g, rot, n, m, d);
  return check_launch();
}

extern "C" int peclr_rotation_2d_matrix(const double* angle, const float* center_x, const float* center_y,
                                        double scale, float* out, int n, void* stream) {
  if (n == 0) return 0;
  if (!angle || !center_x || !center_y || !out || n < 0) return -1001;
  rotation_matrix_kernel<<<(n + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(angle, center_x, center_y,
                                                                                        scale, out, n);
  return check_launch();
}

extern "C" int peclr_projection_stats(const float* enc, float* out8, int n, int m, int d, void* stream) {
  if (!enc || !out8 || n < 1 || m < 1 || d < 2) return -1001;
  projection_stats_kernel<<<1, kWarpsPerBlock * 32, 0, static_cast<cudaStream_t>(stream)>>>(enc, out8, n, m, d);
  return check_launch();
}
