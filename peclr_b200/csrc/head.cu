// Projection head in fp32 (reference: SimCLR.get_projection_head, src/models/unsupervised/
// simclr_model.py:20-35: Linear(2048,512,bias) -> BatchNorm1d(512) -> ReLU -> Linear(512,128,no bias)).
// The head is < 0.05 % of the step's FLOPs and feeds the fp32 loss chain, so it stays in fp32 on the
// CUDA cores: one strided SIMT GEMM (covers X*W^T, dY*W and dY^T*X) and fused BatchNorm1d(+ReLU)
// forward / backward kernels.
#include "../../include/peclr_b200.h"
#include "ptx.cuh"

namespace peclr {

// C[m, n] (+)= sum_k A[m*sam + k*sak] * B[k*sbk + n*sbn] (+ bias[n]);  32x32 tile, 256 threads, 2x2 each.
// The GEMMs of the head are tiny (M = 2B rows): blockIdx.z splits K so that a launch fills the machine.  With more
// than one split every block stores its partial tile to the workspace slab of its split; the LAST block to arrive
// for a tile (counter in the workspace, self-resetting) adds the slabs in split order and writes C -- a fixed
// summation order, so the result is reproducible (no floating-point atomics).
__global__ void __launch_bounds__(256) sgemm_strided_kernel(const float* __restrict__ A, const float* __restrict__ B,
                                                            float* C, const float* __restrict__ bias, int M, int N,
                                                            int K, long long sam, long long sak, long long sbk,
                                                            long long sbn, long long ldc, int accumulate, int k_per,
                                                            float* partial, unsigned* counters) {
  __shared__ float As[32][33];
  __shared__ float Bs[32][33];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int m0 = blockIdx.y * 32, n0 = blockIdx.x * 32;
  float acc[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
  // loader mapping: pick the fast index along whichever stride is 1
  const bool a_kfast = sak == 1;
  const bool b_nfast = sbn == 1;
  const int k_begin = blockIdx.z * k_per;
  K = min(K, k_begin + k_per);
  for (int k0 = k_begin; k0 < K; k0 += 32) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int idx = threadIdx.x + i * 256;
      const int f = idx & 31, s = idx >> 5;
      {
        const int mm = a_kfast ? s : f, kk = a_kfast ? f : s;
        const int gm = m0 + mm, gk = k0 + kk;
        As[kk][mm] = (gm < M && gk < K) ? A[gm * sam + gk * sak] : 0.f;
      }
      {
        const int nn = b_nfast ? f : s, kk = b_nfast ? s : f;
        const int gn = n0 + nn, gk = k0 + kk;
        Bs[kk][nn] = (gn < N && gk < K) ? B[gk * sbk + gn * sbn] : 0.f;
      }
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 32; ++kk) {
      const float a0 = As[kk][ty * 2], a1 = As[kk][ty * 2 + 1];
      const float b0 = Bs[kk][tx * 2], b1 = Bs[kk][tx * 2 + 1];
      acc[0][0] = fmaf(a0, b0, acc[0][0]);
      acc[0][1] = fmaf(a0, b1, acc[0][1]);
      acc[1][0] = fmaf(a1, b0, acc[1][0]);
      acc[1][1] = fmaf(a1, b1, acc[1][1]);
    }
    __syncthreads();
  }
  if (gridDim.z > 1) {
    __shared__ int is_last;
    const long long slab = (long long)M * N;
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int gm = m0 + ty * 2 + i, gn = n0 + tx * 2 + j;
        if (gm < M && gn < N) __stcg(partial + blockIdx.z * slab + (long long)gm * N + gn, acc[i][j]);
      }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
      unsigned* ctr = counters + blockIdx.y * gridDim.x + blockIdx.x;
      const unsigned prev = atomicAdd(ctr, 1u);
      is_last = prev == gridDim.z - 1;
      if (is_last) *ctr = 0u;  // ready for the next launch
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        acc[i][j] = 0.f;
        const int gm = m0 + ty * 2 + i, gn = n0 + tx * 2 + j;
        if (gm < M && gn < N)
          for (unsigned z = 0; z < gridDim.z; ++z) acc[i][j] += __ldcg(partial + z * slab + (long long)gm * N + gn);
      }
  }
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int gm = m0 + ty * 2 + i, gn = n0 + tx * 2 + j;
      if (gm < M && gn < N) {
        const float v = acc[i][j] + (bias ? bias[gn] : 0.f);
        float* c = C + (long long)gm * ldc + gn;
        *c = accumulate ? *c + v : v;
      }
    }
}

// BatchNorm1d (training) + ReLU on x [M][C]: one block per 32 channels, 8 warps stride over rows.
__global__ void __launch_bounds__(256) bn1d_relu_fwd_kernel(const float* __restrict__ x, const float* gamma,
                                                            const float* beta, float* out, float* mean_out,
                                                            float* invstd_out, float* running_mean,
                                                            float* running_var, int M, int C, float eps,
                                                            float momentum) {
  __shared__ float red[8][32];
  __shared__ float stat[2][32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + lane;
  const bool ok = c < C;
  float s = 0.f;
  for (int r = w; r < M; r += 8) s += ok ? x[(long long)r * C + c] : 0.f;
  red[w][lane] = s;
  __syncthreads();
  if (w == 0) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += red[i][lane];
    stat[0][lane] = t / (float)M;
  }
  __syncthreads();
  const float mean = stat[0][lane];
  float q = 0.f;
  for (int r = w; r < M; r += 8) {
    const float d = ok ? x[(long long)r * C + c] - mean : 0.f;
    q = fmaf(d, d, q);
  }
  __syncthreads();
  red[w][lane] = q;
  __syncthreads();
  if (w == 0) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += red[i][lane];
    const float var = t / (float)M;
    stat[1][lane] = rsqrtf(var + eps);
    if (ok) {
      mean_out[c] = mean;
      invstd_out[c] = stat[1][lane];
      if (running_mean) {
        const float unbiased = M > 1 ? t / (float)(M - 1) : var;
        running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mean;
        running_var[c] = (1.f - momentum) * running_var[c] + momentum * unbiased;
      }
    }
  }
  __syncthreads();
  if (!ok) return;
  const float sc = gamma[c] * stat[1][lane];
  const float sh = beta[c] - mean * sc;
  for (int r = w; r < M; r += 8) out[(long long)r * C + c] = fmaxf(fmaf(x[(long long)r * C + c], sc, sh), 0.f);
}

// backward of the above: g = dout * [out > 0]; dgamma += sum g*xhat; dbeta += sum g;
// dx = gamma*invstd*(g - mean(g) - xhat*mean(g*xhat))
__global__ void __launch_bounds__(256) bn1d_relu_bwd_kernel(const float* __restrict__ dout,
                                                            const float* __restrict__ out,
                                                            const float* __restrict__ x, const float* mean,
                                                            const float* invstd, const float* gamma, float* dx,
                                                            float* dgamma, float* dbeta, int M, int C) {
  __shared__ float red[2][8][32];
  __shared__ float tot[2][32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + lane;
  const bool ok = c < C;
  const float mu = ok ? mean[c] : 0.f, is = ok ? invstd[c] : 0.f;
  float s = 0.f, d = 0.f;
  for (int r = w; r < M; r += 8) {
    if (ok) {
      const long long o = (long long)r * C + c;
      const float g = out[o] > 0.f ? dout[o] : 0.f;
      s += g;
      d = fmaf(g, (x[o] - mu) * is, d);
    }
  }
  red[0][w][lane] = s;
  red[1][w][lane] = d;
  __syncthreads();
  if (w == 0) {
    float a = 0.f, b = 0.f;
    for (int i = 0; i < 8; ++i) a += red[0][i][lane], b += red[1][i][lane];
    tot[0][lane] = a;
    tot[1][lane] = b;
    if (ok) {
      dbeta[c] += a;
      dgamma[c] += b;
    }
  }
  __syncthreads();
  if (!ok) return;
  const float k1 = gamma[c] * is, k2 = tot[0][lane] / (float)M, k3 = tot[1][lane] / (float)M;
  for (int r = w; r < M; r += 8) {
    const long long o = (long long)r * C + c;
    const float g = out[o] > 0.f ? dout[o] : 0.f;
    dx[o] = k1 * (g - k2 - (x[o] - mu) * is * k3);
  }
}

// column sums: out[n] += sum_m x[m][n]   (bias gradient)
__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ x, float* out, int M, int N) {
  __shared__ float red[8][32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + lane;
  float s = 0.f;
  if (c < N)
    for (int r = w; r < M; r += 8) s += x[(long long)r * N + c];
  red[w][lane] = s;
  __syncthreads();
  if (w == 0 && c < N) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += red[i][lane];
    out[c] += t;
  }
}

}  // namespace peclr

using namespace peclr;

static int head_last_error() {
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : -(int)e;
}

// workspace layout: [tile counters, kCounterBytes][splits slabs of M*N floats].  The counters sit at a FIXED place
// (K is only split while there are fewer than 2*148 tiles, so 512 counters always suffice): GEMMs of different
// shapes can share one workspace without one's partials landing on another's (zero-at-rest) counters.
constexpr long long kCounterBytes = 2048;
// split K until there are about two blocks per SM (at least 64 of K per split, whole 32-wide k steps)
static int sgemm_splits(int M, int N, int K, int* k_per) {
  const int tiles = ((N + 31) / 32) * ((M + 31) / 32);
  int splits = (2 * 148 + tiles - 1) / tiles;
  if (splits > K / 64) splits = K / 64;
  if (splits < 1) splits = 1;
  *k_per = ((K + splits - 1) / splits + 31) / 32 * 32;
  return (K + *k_per - 1) / *k_per;
}

extern "C" long long peclr_sgemm_workspace_bytes(int M, int N, int K) {
  if (M < 1 || N < 1 || K < 1) return -1001;
  int k_per = 0;
  const int splits = sgemm_splits(M, N, K, &k_per);
  if (splits == 1) return 0;
  return kCounterBytes + (long long)splits * M * N * 4;
}

extern "C" int peclr_sgemm(const float* A, const float* B, float* C, const float* bias, int M, int N, int K,
                           long long sam, long long sak, long long sbk, long long sbn, long long ldc, int accumulate,
                           void* workspace, long long workspace_bytes, void* stream) {
  if (M < 1 || N < 1 || K < 1) return -1001;
  dim3 grid((N + 31) / 32, (M + 31) / 32);
  int k_per = 0;
  int splits = sgemm_splits(M, N, K, &k_per);
  // no (or too small a) workspace: one split per tile (still exact, just fewer blocks in flight)
  if (splits > 1 && (!workspace || workspace_bytes < peclr_sgemm_workspace_bytes(M, N, K))) {
    splits = 1;
    k_per = (K + 31) / 32 * 32;
  }
  grid.z = splits;
  // the counters must be zero before the first use (the caller zeroes the workspace once; the kernel resets them)
  unsigned* counters = static_cast<unsigned*>(workspace);
  float* partial = splits > 1 ? reinterpret_cast<float*>(static_cast<char*>(workspace) + kCounterBytes) : nullptr;
  sgemm_strided_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(A, B, C, bias, M, N, K, sam, sak, sbk, sbn, ldc,
                                                                accumulate, k_per, partial, counters);
  return head_last_error();
}

extern "C" int peclr_bn1d_relu_fwd(const float* x, const float* gamma, const float* beta, float* out, float* mean_out,
                                   float* invstd_out, float* running_mean, float* running_var, int M, int C, float eps,
                                   float momentum, void* stream) {
  bn1d_relu_fwd_kernel<<<(C + 31) / 32, 256, 0, (cudaStream_t)stream>>>(x, gamma, beta, out, mean_out, invstd_out,
                                                                         running_mean, running_var, M, C, eps,
                                                                         momentum);
  return head_last_error();
}

extern "C" int peclr_bn1d_relu_bwd(const float* dout, const float* out, const float* x, const float* mean,
                                   const float* invstd, const float* gamma, float* dx, float* dgamma, float* dbeta,
                                   int M, int C, void* stream) {
  bn1d_relu_bwd_kernel<<<(C + 31) / 32, 256, 0, (cudaStream_t)stream>>>(dout, out, x, mean, invstd, gamma, dx, dgamma,
                                                                         dbeta, M, C);
  return head_last_error();
}

extern "C" int peclr_colsum_acc(const float* x, float* out, int M, int N, void* stream) {
  colsum_kernel<<<(N + 31) / 32, 256, 0, (cudaStream_t)stream>>>(x, out, M, N);
  return head_last_error();
}
