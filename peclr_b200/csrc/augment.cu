// GPU-side two-view augmentation: the pixel work of the reference's data pipeline in front of the PeCLR step
// (SampleAugmenter.transform_sample, src/data_loader/sample_augmenter.py:47-129, + ToTensor / Normalize,
// data_loader/utils.py:287-293) for the augmentations of the paper's recipe, ONE launch per batch:
//
//   rotate   cv2.warpAffine(INTER_LINEAR, BORDER_CONSTANT 0)       sample_augmenter.py:216-243
//   crop     numpy slice of the rotated image (clipped)             :163-186
//   resize   cv2.resize(INTER_AREA)                                 :188-214
//   jitter   BGR2HSV, h*=, s*=, v = a*v + b, clip, uint8, HSV2BGR    :265-294
//   ToTensor + Normalize -> fp32 NCHW (the reference's `transformed_image{1,2}`)
//
// The arithmetic is OpenCV's 8-bit arithmetic (the reference calls opencv-python; not vendored):
//   * warpAffine: dst -> src coordinates in fixed point (AB_BITS 10, 1/32-pixel positions), bilinear weights in
//     1/32768 units, (sum + 2^14) >> 15; taps outside the image read 0.
//   * INTER_AREA, both scale factors >= 1: area-weighted average (computeResizeAreaTab weights, fp32 accumulation in
//     OpenCV's order: columns inside a row, then rows), or the integer fast path when both factors are integers;
//     otherwise (up-scaling in a dimension) OpenCV's fixed-point bilinear path with the INTER_AREA coefficient rule.
//   * RGB2HSV_b integer conversion (sdiv / hdiv tables, hsv_shift 12) and the fp32 HSV2RGB with truncation to 8 bits.
// A thread owns one output pixel (3 channels) and evaluates the rotated-image pixels its resize cell needs on the
// fly (each is 4 source taps): no intermediate image exists.  Launch-bound (a 128-sample batch of 224 x 224 sources
// is 19 MB in, 25 MB out), so no shared-memory staging: the source taps of neighbouring threads hit L1 / L2.
#include "../../include/peclr_b200.h"
#include <math.h>

#include "ptx.cuh"

namespace peclr {

struct AugView {  // one row of the parameter table (peclr_b200/gpu_augment.py: VIEW_DTYPE)
  double m[6];     // dst -> src affine map (already inverted as cv2.warpAffine does)
  long long src_off;
  int sh, sw;
  int ox, oy, cw, ch;
  double h, s, a, b;
  int rotate, jitter;
};

struct Px {
  int c[3];
};

// Fixed-point source coordinates of cv2.warpAffine: X = (X0(yr) + adelta(xr)) >> 5 with X0 = rn((m1*yr + m2)*1024) + 16,
// adelta = rn(m0*xr*1024) (and Y likewise).  The column terms of a block's crop columns are computed once per block
// into shared memory and the row terms once per thread and row: the inner loops are integer-only (fp64 is slow here).
struct RowTerm {
  int X0, Y0;
};
__device__ __forceinline__ RowTerm row_term(const AugView& v, int yr) {
  RowTerm t;
  t.X0 = __double2int_rn((v.m[1] * yr + v.m[2]) * 1024.0) + 16;
  t.Y0 = __double2int_rn((v.m[4] * yr + v.m[5]) * 1024.0) + 16;
  return t;
}

// value of the ROTATED image at integer (xr, yr): cv2.warpAffine, bit for bit.  cols = {adelta, bdelta} of the crop's
// columns (index xr - ox) in shared memory, or nullptr (computed here).
__device__ __forceinline__ Px rotated_pixel(const uint8_t* __restrict__ img, const AugView& v, int xr, int yr,
                                            const RowTerm& rt, const int2* cols) {
  Px o;
  if (!v.rotate) {
    const uint8_t* p = img + ((size_t)yr * v.sw + xr) * 3;
    o.c[0] = p[0], o.c[1] = p[1], o.c[2] = p[2];
    return o;
  }
  int ad, bd;
  if (cols) {
    const int2 c = cols[xr - v.ox];
    ad = c.x, bd = c.y;
  } else {
    ad = __double2int_rn(v.m[0] * xr * 1024.0), bd = __double2int_rn(v.m[3] * xr * 1024.0);
  }
  const int X = (rt.X0 + ad) >> 5;
  const int Y = (rt.Y0 + bd) >> 5;
  const int sx = X >> 5, sy = Y >> 5, fx = X & 31, fy = Y & 31;
  const int w00 = (32 - fx) * (32 - fy) * 32, w01 = fx * (32 - fy) * 32, w10 = (32 - fx) * fy * 32, w11 = fx * fy * 32;
  const bool x0 = sx >= 0 && sx < v.sw, x1 = sx + 1 >= 0 && sx + 1 < v.sw;
  const bool y0 = sy >= 0 && sy < v.sh, y1 = sy + 1 >= 0 && sy + 1 < v.sh;
  const uint8_t* p = img + ((long long)sy * v.sw + sx) * 3;
  const long long row = (long long)v.sw * 3;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const int p00 = (x0 && y0) ? p[k] : 0, p01 = (x1 && y0) ? p[3 + k] : 0;
    const int p10 = (x0 && y1) ? p[row + k] : 0, p11 = (x1 && y1) ? p[row + 3 + k] : 0;
    o.c[k] = (p00 * w00 + p01 * w01 + p10 * w10 + p11 * w11 + (1 << 14)) >> 15;
  }
  return o;
}

// computeResizeAreaTab for one destination index: cells [first, last] with the weight of each
struct AreaCell {
  int first, count;     // source cells first .. first + count - 1
  float w_first, w_mid, w_last;
  bool has_first, has_last;
};
__device__ __forceinline__ AreaCell area_cell(int d, int ssize, double scale) {
  AreaCell c;
  const double fsx1 = d * scale, fsx2 = fsx1 + scale;
  const double cell = fmin(scale, ssize - fsx1);
  int sx1 = (int)ceil(fsx1), sx2 = (int)floor(fsx2);
  sx2 = min(sx2, ssize - 1);
  sx1 = min(sx1, sx2);
  c.has_first = sx1 - fsx1 > 1e-3;
  c.has_last = fsx2 - sx2 > 1e-3;
  c.w_first = (float)((sx1 - fsx1) / cell);
  c.w_mid = (float)(1.0 / cell);
  c.w_last = (float)(fmin(fmin(fsx2 - sx2, 1.0), cell) / cell);
  c.first = c.has_first ? sx1 - 1 : sx1;
  c.count = (sx2 - sx1) + (c.has_first ? 1 : 0) + (c.has_last ? 1 : 0);
  return c;
}
__device__ __forceinline__ float area_weight(const AreaCell& c, int i) {
  if (i == 0 && c.has_first) return c.w_first;
  if (i == c.count - 1 && c.has_last) return c.w_last;
  return c.w_mid;
}

// INTER_AREA coefficient rule of the bilinear (up-scaling) path: source index + the two 11-bit weights
__device__ __forceinline__ void linear_coeff(int d, int ssize, double scale, double inv_scale, int* s0, int* a0, int* a1) {
  int sx = (int)floor(d * scale);
  float fx = (float)((d + 1) - (sx + 1) * inv_scale);
  fx = fx <= 0.f ? 0.f : fx - floorf(fx);
  if (sx < 0) fx = 0.f, sx = 0;
  if (sx >= ssize - 1) fx = 0.f, sx = ssize - 1;
  *s0 = sx;
  *a0 = max(-32768, min(32767, __float2int_rn(__fmul_rn(__fsub_rn(1.f, fx), 2048.f))));
  *a1 = max(-32768, min(32767, __float2int_rn(__fmul_rn(fx, 2048.f))));
}

// OpenCV's RGB2HSV_b division tables (hsv_shift 12): sdiv[i] = round((255 << 12) / i), hdiv[i] = round((180 << 12) /
// (6 i)); filled once from the host (the same double-precision expression OpenCV evaluates).
__constant__ int c_sdiv[256];
__constant__ int c_hdiv[256];
__device__ __forceinline__ int sdiv_entry(int i) { return c_sdiv[i]; }
__device__ __forceinline__ int hdiv_entry(int i) { return c_hdiv[i]; }

// color_jitter_sample on one 8-bit pixel (channel 0 plays "B", as the reference passes its RGB image to BGR2HSV)
__device__ __forceinline__ Px color_jitter(Px p, const AugView& v) {
  const int b = p.c[0], g = p.c[1], r = p.c[2];
  const int vmax = max(max(b, g), r), vmin = min(min(b, g), r);
  const int diff = vmax - vmin;
  int s = (diff * sdiv_entry(vmax) + (1 << 11)) >> 12;
  int h = vmax == r ? g - b : (vmax == g ? b - r + 2 * diff : r - g + 4 * diff);
  h = (h * hdiv_entry(diff) + (1 << 11)) >> 12;
  h += h < 0 ? 180 : 0;
  h = min(max(h, 0), 255);
  // hue * h, sat * s, val * a + b in float64, clipped to [0, 255], truncated to uint8 (ndarray.astype)
  const int hq = (int)fmin(fmax((double)h * v.h, 0.0), 255.0);
  const int sq = (int)fmin(fmax((double)s * v.s, 0.0), 255.0);
  const int vq = (int)fmin(fmax((double)vmax * v.a + v.b, 0.0), 255.0);
  // HSV2BGR, 8-bit: fp32 HSV2RGB on (h, s / 255, v / 255), result * 255 truncated
  const float fs = __fmul_rn((float)sq, 1.0f / 255.0f), fv = __fmul_rn((float)vq, 1.0f / 255.0f);
  float bb, gg, rr;
  if (sq == 0) {
    bb = gg = rr = fv;
  } else {
    float hh = __fmul_rn((float)hq, 6.f / 180.f);
    if (hh >= 6.f) hh = __fsub_rn(hh, 6.f);
    int sector = (int)floorf(hh);
    hh = __fsub_rn(hh, (float)sector);
    if ((unsigned)sector >= 6u) sector = 0, hh = 0.f;
    float tab[4];
    tab[0] = fv;
    tab[1] = __fmul_rn(fv, __fsub_rn(1.f, fs));
    tab[2] = __fmul_rn(fv, __fsub_rn(1.f, __fmul_rn(fs, hh)));
    tab[3] = __fmul_rn(fv, __fsub_rn(1.f, __fmul_rn(fs, __fsub_rn(1.f, hh))));
    const int sd[6][3] = {{1, 3, 0}, {1, 0, 2}, {3, 0, 1}, {0, 2, 1}, {0, 1, 3}, {2, 1, 0}};
    bb = tab[sd[sector][0]], gg = tab[sd[sector][1]], rr = tab[sd[sector][2]];
  }
  Px o;
  o.c[0] = min(max((int)__fmul_rn(bb, 255.f), 0), 255);
  o.c[1] = min(max((int)__fmul_rn(gg, 255.f), 0), 255);
  o.c[2] = min(max((int)__fmul_rn(rr, 255.f), 0), 255);
  return o;
}

__global__ void __launch_bounds__(128) two_view_augment_kernel(const uint8_t* __restrict__ src,
                                                               const AugView* __restrict__ views, int n, int dh, int dw,
                                                               float m0, float m1, float m2, float s0, float s1,
                                                               float s2, float* __restrict__ out,
                                                               uint8_t* __restrict__ stage, int cols_cap) {
  extern __shared__ int2 s_cols[];
  const int img_i = blockIdx.y, dy = blockIdx.x;
  const AugView v = views[img_i];
  const uint8_t* img = src + v.src_off;
  const int2* cols = nullptr;
  if (v.rotate && v.cw <= cols_cap) {
    for (int i = threadIdx.x; i < v.cw; i += blockDim.x)
      s_cols[i] = make_int2(__double2int_rn(v.m[0] * (v.ox + i) * 1024.0), __double2int_rn(v.m[3] * (v.ox + i) * 1024.0));
    cols = s_cols;
  }
  __syncthreads();
  for (int dx = threadIdx.x; dx < dw; dx += blockDim.x) {
    Px px;
    px.c[0] = px.c[1] = px.c[2] = 0;
    if (v.cw > 0 && v.ch > 0) {
      const double inv_x = (double)dw / v.cw, inv_y = (double)dh / v.ch;
      const double sc_x = 1.0 / inv_x, sc_y = 1.0 / inv_y;
      if (v.cw == dw && v.ch == dh) {  // same size: cv::resize copies
        px = rotated_pixel(img, v, v.ox + dx, v.oy + dy, row_term(v, v.oy + dy), cols);
      } else if (sc_x >= 1.0 && sc_y >= 1.0) {
        const int isx = __double2int_rn(sc_x), isy = __double2int_rn(sc_y);
        if (fabs(sc_x - isx) < 2.220446049250313e-16 && fabs(sc_y - isy) < 2.220446049250313e-16) {
          int sum[3] = {0, 0, 0};  // integer factors: plain box sum
          for (int yy = 0; yy < isy; ++yy) {
            const RowTerm rt = row_term(v, v.oy + dy * isy + yy);
            for (int xx = 0; xx < isx; ++xx) {
              const Px q = rotated_pixel(img, v, v.ox + dx * isx + xx, v.oy + dy * isy + yy, rt, cols);
              sum[0] += q.c[0], sum[1] += q.c[1], sum[2] += q.c[2];
            }
          }
          const float scale = 1.f / (float)(isx * isy);
#pragma unroll
          for (int k = 0; k < 3; ++k)
            px.c[k] = (isx == 2 && isy == 2) ? (sum[k] + 2) >> 2
                                             : min(max(__float2int_rn(__fmul_rn((float)sum[k], scale)), 0), 255);
        } else {
          const AreaCell cx = area_cell(dx, v.cw, sc_x), cy = area_cell(dy, v.ch, sc_y);
          float acc[3] = {0.f, 0.f, 0.f};
          for (int j = 0; j < cy.count; ++j) {
            float buf[3] = {0.f, 0.f, 0.f};
            const RowTerm rt = row_term(v, v.oy + cy.first + j);
            for (int i = 0; i < cx.count; ++i) {
              const Px q = rotated_pixel(img, v, v.ox + cx.first + i, v.oy + cy.first + j, rt, cols);
              const float al = area_weight(cx, i);
#pragma unroll
              for (int k = 0; k < 3; ++k) buf[k] = __fadd_rn(buf[k], __fmul_rn((float)q.c[k], al));
            }
            const float be = area_weight(cy, j);
#pragma unroll
            for (int k = 0; k < 3; ++k)
              acc[k] = j == 0 ? __fmul_rn(be, buf[k]) : __fadd_rn(acc[k], __fmul_rn(be, buf[k]));
          }
#pragma unroll
          for (int k = 0; k < 3; ++k) px.c[k] = min(max(__float2int_rn(acc[k]), 0), 255);
        }
      } else {
        int x0, ax0, ax1, y0, by0, by1;
        linear_coeff(dx, v.cw, sc_x, inv_x, &x0, &ax0, &ax1);
        linear_coeff(dy, v.ch, sc_y, inv_y, &y0, &by0, &by1);
        const int x1 = min(x0 + 1, v.cw - 1), y1 = min(y0 + 1, v.ch - 1);
        const RowTerm r0 = row_term(v, v.oy + y0), r1 = row_term(v, v.oy + y1);
        const Px p00 = rotated_pixel(img, v, v.ox + x0, v.oy + y0, r0, cols);
        const Px p01 = rotated_pixel(img, v, v.ox + x1, v.oy + y0, r0, cols);
        const Px p10 = rotated_pixel(img, v, v.ox + x0, v.oy + y1, r1, cols);
        const Px p11 = rotated_pixel(img, v, v.ox + x1, v.oy + y1, r1, cols);
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          const int h0 = p00.c[k] * ax0 + p01.c[k] * ax1, h1 = p10.c[k] * ax0 + p11.c[k] * ax1;
          px.c[k] = min(max((((by0 * (h0 >> 4)) >> 16) + ((by1 * (h1 >> 4)) >> 16) + 2) >> 2, 0), 255);
        }
      }
    }
    if (v.jitter) px = color_jitter(px, v);
    const size_t pix = (size_t)dy * dw + dx;
    if (stage) {
      uint8_t* st = stage + ((size_t)img_i * dh * dw + pix) * 3;
      st[0] = (uint8_t)px.c[0], st[1] = (uint8_t)px.c[1], st[2] = (uint8_t)px.c[2];
    }
    // ToTensor (x / 255) + Normalize ((t - mean) / std), fp32, CHW
    float* o = out + (size_t)img_i * 3 * dh * dw + pix;
    const size_t plane = (size_t)dh * dw;
    o[0] = __fdiv_rn(__fsub_rn(__fdiv_rn((float)px.c[0], 255.f), m0), s0);
    o[plane] = __fdiv_rn(__fsub_rn(__fdiv_rn((float)px.c[1], 255.f), m1), s1);
    o[2 * plane] = __fdiv_rn(__fsub_rn(__fdiv_rn((float)px.c[2], 255.f), m2), s2);
  }
}

}  // namespace peclr

using namespace peclr;

extern "C" int peclr_two_view_augment(const void* src_u8, long long src_bytes, const void* view_table, int n,
                                      int out_h, int out_w, float mean0, float mean1, float mean2, float std0,
                                      float std1, float std2, float* out, void* stage_u8, void* stream) {
  if (!src_u8 || !view_table || !out || n < 1 || n > 65535 || out_h < 1 || out_w < 1 || src_bytes < 1) return -1001;
  static_assert(sizeof(AugView) == 120, "parameter table row layout (gpu_augment.VIEW_DTYPE)");
  static bool tables_ready = false;
  if (!tables_ready) {
    int sdiv[256], hdiv[256];
    sdiv[0] = hdiv[0] = 0;
    for (int i = 1; i < 256; ++i) {
      sdiv[i] = (int)nearbyint((double)(255 << 12) / (1.0 * i));
      hdiv[i] = (int)nearbyint((double)(180 << 12) / (6.0 * i));
    }
    if (cudaMemcpyToSymbol(c_sdiv, sdiv, sizeof(sdiv)) != cudaSuccess ||
        cudaMemcpyToSymbol(c_hdiv, hdiv, sizeof(hdiv)) != cudaSuccess)
      return -(int)cudaGetLastError();
    tables_ready = true;
  }
  dim3 grid(out_h, n);
  const int cols_cap = 1024;  // crop columns whose fixed-point terms fit the block's shared-memory table (8 KB)
  two_view_augment_kernel<<<grid, 128, cols_cap * sizeof(int2), static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint8_t*>(src_u8), static_cast<const AugView*>(view_table), n, out_h, out_w, mean0, mean1,
      mean2, std0, std1, std2, out, static_cast<uint8_t*>(stage_u8), cols_cap);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : -(int)e;
}
