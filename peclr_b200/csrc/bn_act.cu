// HBM-bound elementwise / reduction kernels of the trunk: training-mode BatchNorm apply (+ residual,
// + ReLU), BatchNorm backward (reduce + apply), stem max-pool (fused with BN + ReLU) forward/backward,
// global average pool, and the NCHW fp32 -> padded space-to-depth bf16 input transform.
//
// All activation tensors are [M = N*H*W][C] bf16 with C contiguous; every thread owns 8 consecutive
// channels (one 16-byte vector), keeps the per-channel BatchNorm coefficients in registers and walks
// rows with kRows independent 128-bit loads in flight, so accesses are fully coalesced and the memory
// system sees enough parallelism.  Grids are sized in multiples of the SM count.
//
// Replaces ATen/cuDNN BatchNorm2d(train) + ReLU + residual add + MaxPool2d + AdaptiveAvgPool2d of the
// reference trunk (src/models/resnet_model.py:16-26; torchvision ResNet.forward).
#include "../../include/peclr_b200.h"
#include "conv_tc.h"
#include "ptx.cuh"

namespace peclr {

constexpr int kRowsDefault = 4;  // rows (16-byte loads) in flight per thread (PECLR_ELT_ROWS overrides: 2 or 4)

struct alignas(16) bf16x8 {
  uint32_t v[4];
};

__device__ __forceinline__ void unpack8(const bf16x8& p, float (&f)[8]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    f[2 * i] = bf16_lo(p.v[i]);
    f[2 * i + 1] = bf16_hi(p.v[i]);
  }
}
__device__ __forceinline__ bf16x8 pack8(const float (&f)[8]) {
  bf16x8 p;
#pragma unroll
  for (int i = 0; i < 4; ++i) p.v[i] = pack_bf16x2(f[2 * i], f[2 * i + 1]);
  return p;
}
// streaming 128-bit load (read once: do not pollute L1) / plain 128-bit store
__device__ __forceinline__ bf16x8 ld8(const __nv_bfloat16* p) {
  bf16x8 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];\n"
               : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3])
               : "l"(p));
  return r;
}
__device__ __forceinline__ bf16x8 ld8_cached(const __nv_bfloat16* p) { return *reinterpret_cast<const bf16x8*>(p); }
__device__ __forceinline__ void st8(__nv_bfloat16* p, const bf16x8& v) { *reinterpret_cast<bf16x8*>(p) = v; }

// mean / variance / inverse std of 8 channels from the fp64 sums the convolution epilogue accumulated.  Mean and
// E[y^2] - mean^2 are formed in fp64 (three double operations per channel): in fp32 the subtraction cancels
// catastrophically for channels with |mean| >> std.  Only the final values are rounded to fp32.
__device__ __forceinline__ void bn_coeffs(const double* sum, const double* sumsq, int c0, double inv_m, float eps,
                                          float (&mean)[8], float (&var)[8], float (&invstd)[8]) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const double m = sum[c0 + i] * inv_m;
    const double v = fma(-m, m, sumsq[c0 + i] * inv_m);
    mean[i] = (float)m;
    var[i] = fmaxf((float)v, 0.f);
    invstd[i] = rsqrtf(var[i] + eps);
  }
}

// Row schedule of a block: grid-strided (rows interleaved across blocks) or one contiguous range per block.
struct RowWalk {
  long long r0, rend, rstep, jstride;
};
template <int kRows>
__device__ __forceinline__ RowWalk row_walk(long long M, int rows_per_block, int roff, int contig) {
  RowWalk w;
  if (contig) {
    const long long unit = (long long)kRows * rows_per_block;
    const long long per = ((M + gridDim.x - 1) / gridDim.x + unit - 1) / unit * unit;
    w.r0 = (long long)blockIdx.x * per + roff;
    w.rend = min(M, ((long long)blockIdx.x + 1) * per);
    w.rstep = unit;
    w.jstride = rows_per_block;
  } else {
    const long long stride = (long long)gridDim.x * rows_per_block;
    w.r0 = (long long)blockIdx.x * rows_per_block + roff;
    w.rend = M;
    w.rstep = kRows * stride;
    w.jstride = stride;
  }
  return w;
}

struct BnApplyArgs {
  const __nv_bfloat16* y;
  const double *sum, *sumsq;
  const float *gamma, *beta;
  const __nv_bfloat16* res;                     // optional second operand of the residual add
  const double *rsum, *rsumsq;  // if non-null, res is a raw conv output with its own BN
  const float *rgamma, *rbeta;
  __nv_bfloat16* out;
  uint8_t* mask_out;  // optional [M][C/8]: bit i of byte (row, channel group) = out[8g + i] > 0
  float *mean_out, *invstd_out, *running_mean, *running_var;
  float *rmean_out, *rinvstd_out, *rrunning_mean, *rrunning_var;
  long long M;
  int C;
  float eps, momentum;
  int relu;
  int contig;  // 1: each block walks one contiguous row range (DRAM locality) instead of a grid-strided one
};

__device__ __forceinline__ void bn_bookkeeping(const float (&mean)[8], const float (&var)[8],
                                               const float (&invstd)[8], int c0, long long M, float momentum,
                                               float* mean_out, float* invstd_out, float* running_mean,
                                               float* running_var) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    mean_out[c0 + i] = mean[i];
    invstd_out[c0 + i] = invstd[i];
    if (running_mean) {
      const float unbiased = M > 1 ? var[i] * (float)((double)M / (double)(M - 1)) : var[i];
      running_mean[c0 + i] = (1.f - momentum) * running_mean[c0 + i] + momentum * mean[i];
      running_var[c0 + i] = (1.f - momentum) * running_var[c0 + i] + momentum * unbiased;
    }
  }
}

template <int kRows>
__global__ void __launch_bounds__(256) bn_apply_kernel(const BnApplyArgs a) {
  pdl_launch_dependents();
  pdl_wait();
  const int tpr = a.C >> 3;  // threads per row
  const int rows_per_block = 256 / tpr;
  const int cg = threadIdx.x % tpr;
  const int roff = threadIdx.x / tpr;
  const int c0 = cg * 8;
  const double inv_m = 1.0 / (double)a.M;
  float mean[8], var[8], invstd[8], scale[8], shift[8];
  bn_coeffs(a.sum, a.sumsq, c0, inv_m, a.eps, mean, var, invstd);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    scale[i] = a.gamma[c0 + i] * invstd[i];
    shift[i] = a.beta[c0 + i] - mean[i] * scale[i];
  }
  float rscale[8], rshift[8];
  const bool has_res = a.res != nullptr;
  const bool res_bn = has_res && a.rsum != nullptr;
  if (res_bn) {
    float rmean[8], rvar[8], rinvstd[8];
    bn_coeffs(a.rsum, a.rsumsq, c0, inv_m, a.eps, rmean, rvar, rinvstd);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      rscale[i] = a.rgamma[c0 + i] * rinvstd[i];
      rshift[i] = a.rbeta[c0 + i] - rmean[i] * rscale[i];
    }
    if (blockIdx.x == 0 && roff == 0)
      bn_bookkeeping(rmean, rvar, rinvstd, c0, a.M, a.momentum, a.rmean_out, a.rinvstd_out, a.rrunning_mean,
                     a.rrunning_var);
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) rscale[i] = 1.f, rshift[i] = 0.f;
  }
  if (blockIdx.x == 0 && roff == 0)
    bn_bookkeeping(mean, var, invstd, c0, a.M, a.momentum, a.mean_out, a.invstd_out, a.running_mean, a.running_var);

  const RowWalk rw = row_walk<kRows>(a.M, rows_per_block, roff, a.contig);
  for (long long r = rw.r0; r < rw.rend; r += rw.rstep) {
    bf16x8 yv[kRows], rv[kRows];
#pragma unroll
    for (int j = 0; j < kRows; ++j) {
      const long long rr = r + j * rw.jstride;
      if (rr < rw.rend) {
        yv[j] = ld8(a.y + rr * a.C + c0);
        if (has_res) rv[j] = ld8(a.res + rr * a.C + c0);
      }
    }
#pragma unroll
    for (int j = 0; j < kRows; ++j) {
      const long long rr = r + j * rw.jstride;
      if (rr >= rw.rend) break;
      float v[8];
      unpack8(yv[j], v);
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = fmaf(v[i], scale[i], shift[i]);
      if (has_res) {
        float q[8];
        unpack8(rv[j], q);
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] += fmaf(q[i], rscale[i], rshift[i]);
      }
      if (a.relu) {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = fmaxf(v[i], 0.f);
      }
      st8(a.out + rr * a.C + c0, pack8(v));
      if (a.mask_out) {
        uint32_t bits = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) bits |= (v[i] > 0.f ? 1u : 0u) << i;
        a.mask_out[rr * tpr + cg] = (uint8_t)bits;
      }
    }
  }
}

// ---- BatchNorm backward ---------------------------------------------------------------------------
// g = dout * relu'(.) where the ReLU mask comes from (mask_mode)
//   0: nothing (dout is already masked)        1: a stored activation tensor (mask > 0)
//   2: recomputed from y: fma(y, gamma*invstd, beta - mean*gamma*invstd) > 0  (no extra HBM read)
//   3: bit mask written by bn_apply (1 byte per 8 channels instead of re-reading the 16-byte activation)
struct BnBwdArgs {
  const __nv_bfloat16 *dout, *mask, *y;
  const float *mean, *invstd, *gamma, *beta;
  double* scratch;         // [2C] fp64 accumulators: sum g, sum g*y (one partial per block, fixed order inside)
  __nv_bfloat16 *dy, *g_out;
  float *dgamma, *dbeta;
  long long M;
  int C, mask_mode;
  int contig;
};

// Pass 1 (kApply = false) accumulates sum g and sum g*y per channel; pass 2 folds everything into three
// per-channel constants: dy = A*g + B*y + C with A = gamma*invstd, B = -A*invstd*k3, C = -A*(k2 - k3*invstd*mean),
// k2 = mean(g), k3 = mean(g*xhat) = invstd*(mean(g*y) - mean*mean(g)) -- that difference is formed in fp64 from the
// fp64 totals (it cancels like a variance).  Few live registers -> 3-4 blocks per SM.
template <bool kApply, int kMask, int kRows>
__global__ void __launch_bounds__(256, kApply ? 2 : 3) bn_bwd_kernel(const BnBwdArgs a) {
  __shared__ float red[kApply ? 1 : 2][kApply ? 1 : 256][kApply ? 1 : 9];
  pdl_launch_dependents();
  pdl_wait();
  const int tpr = a.C >> 3;
  const int rows_per_block = 256 / tpr;
  const int cg = threadIdx.x % tpr;
  const int roff = threadIdx.x / tpr;
  const int c0 = cg * 8;
  float scale[8], shift[8], ka[8], kb[8], kc[8], s[8], d[8];
  const float inv_m = 1.f / (float)a.M;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float mu = a.mean[c0 + i], is = a.invstd[c0 + i], gm = a.gamma[c0 + i];
    if (kMask == 2) {
      scale[i] = gm * is;
      shift[i] = a.beta[c0 + i] - mu * scale[i];
    }
    s[i] = d[i] = 0.f;
    if (kApply) {
      const double sgd = a.scratch[c0 + i], sgyd = a.scratch[a.C + c0 + i];
      const float sg = (float)sgd;
      const float sgx = (float)((double)is * fma(-(double)mu, sgd, sgyd));
      const float k2 = sg * inv_m, k3 = sgx * inv_m;
      ka[i] = gm * is;
      kb[i] = -ka[i] * is * k3;
      kc[i] = -ka[i] * (k2 - k3 * is * mu);
      if (blockIdx.x == 0 && roff == 0) {
        a.dbeta[c0 + i] += sg;
        a.dgamma[c0 + i] += sgx;
      }
    }
  }
  const RowWalk rw = row_walk<kRows>(a.M, rows_per_block, roff, a.contig);
  for (long long r = rw.r0; r < rw.rend; r += rw.rstep) {
    bf16x8 gv[kRows], yv[kRows], mv[kRows];
    uint32_t mb[kRows];
#pragma unroll
    for (int j = 0; j < kRows; ++j) {
      const long long rr = r + j * rw.jstride;
      if (rr < rw.rend) {
        gv[j] = ld8(a.dout + rr * a.C + c0);
        yv[j] = ld8(a.y + rr * a.C + c0);
        if (kMask == 1) mv[j] = ld8(a.mask + rr * a.C + c0);
        if (kMask == 3) mb[j] = reinterpret_cast<const uint8_t*>(a.mask)[rr * tpr + cg];
      }
    }
#pragma unroll
    for (int j = 0; j < kRows; ++j) {
      const long long rr = r + j * rw.jstride;
      if (rr >= rw.rend) break;
      float g[8], yf[8];
      unpack8(gv[j], g);
      unpack8(yv[j], yf);
      if (kMask == 1) {
        float m[8];
        unpack8(mv[j], m);
#pragma unroll
        for (int i = 0; i < 8; ++i) g[i] = m[i] > 0.f ? g[i] : 0.f;
      } else if (kMask == 3) {
#pragma unroll
        for (int i = 0; i < 8; ++i) g[i] = (mb[j] >> i) & 1u ? g[i] : 0.f;
      } else if (kMask == 2) {
#pragma unroll
        for (int i = 0; i < 8; ++i) g[i] = fmaf(yf[i], scale[i], shift[i]) > 0.f ? g[i] : 0.f;
      }
      if (kApply) {
        if (a.g_out) st8(a.g_out + rr * a.C + c0, pack8(g));
#pragma unroll
        for (int i = 0; i < 8; ++i) g[i] = fmaf(ka[i], g[i], fmaf(kb[i], yf[i], kc[i]));
        st8(a.dy + rr * a.C + c0, pack8(g));
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          s[i] += g[i];
          d[i] = fmaf(g[i], yf[i], d[i]);
        }
      }
    }
  }
  if (!kApply) {
#pragma unroll
    for (int i = 0; i < 8; ++i) red[0][threadIdx.x][i] = s[i], red[1][threadIdx.x][i] = d[i];
    __syncthreads();
    if (roff == 0) {
      for (int rr = 1; rr < rows_per_block; ++rr) {
#pragma unroll
        for (int i = 0; i < 8; ++i) s[i] += red[0][rr * tpr + cg][i], d[i] += red[1][rr * tpr + cg][i];
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        red_add_f64(a.scratch + c0 + i, (double)s[i]);
        red_add_f64(a.scratch + a.C + c0 + i, (double)d[i]);
      }
    }
  }
}

static int elt_contig() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("PECLR_ELT_CONTIG");
    v = e ? atoi(e) : 0;
  }
  return v;
}

static int elt_rows() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("PECLR_ELT_ROWS");
    v = e ? atoi(e) : kRowsDefault;
    if (v != 2 && v != 4) v = kRowsDefault;
  }
  return v;
}

template <bool kApply, int kRows>
static void launch_bn_bwd_r(const BnBwdArgs& a, int grid, cudaStream_t st) {
  switch (a.mask_mode) {
    case 0: launch_pdl(bn_bwd_kernel<kApply, 0, kRows>, grid, 256, 0, st, a); break;
    case 1: launch_pdl(bn_bwd_kernel<kApply, 1, kRows>, grid, 256, 0, st, a); break;
    case 2: launch_pdl(bn_bwd_kernel<kApply, 2, kRows>, grid, 256, 0, st, a); break;
    default: launch_pdl(bn_bwd_kernel<kApply, 3, kRows>, grid, 256, 0, st, a); break;
  }
}
template <bool kApply>
static void launch_bn_bwd(const BnBwdArgs& a, int grid, cudaStream_t st) {
  const int r = elt_rows();
  if (r == 2) launch_bn_bwd_r<kApply, 2>(a, grid, st);
  else launch_bn_bwd_r<kApply, 4>(a, grid, st);
}

// ---- stem: a = maxpool3x3/s2/p1(relu(bn(y))), y [N][H][W][64] -> a [N][H/2][W/2][64] ----------------
// also records, per output element, which of the 9 window positions won (first maximum in scan order, as
// ATen does) so the backward pass is a cheap gather.
//
// Both kernels are written for a low instruction count (they were issue-bound, not HBM-bound, as per-channel fp32
// code): relu(bn(.)) is monotone in y, increasing or decreasing with the sign of gamma*invstd, so the window
// maximum and its position are found on the RAW bf16 values (sign bit flipped for negative scales) with packed
// bf16x2 max / compare-mask instructions; BatchNorm + ReLU are applied once, to the winner.  (Where several window
// values map to the same activation -- all clamped to 0 -- the recorded position may differ from ATen's, but no
// gradient flows there: the backward pass masks with relu'.  gamma == 0 exactly is the one degenerate case.)
__device__ __forceinline__ uint32_t bf16x2_max(uint32_t a, uint32_t b) {
  const __nv_bfloat162 r = __hmax2(*reinterpret_cast<const __nv_bfloat162*>(&a), *reinterpret_cast<const __nv_bfloat162*>(&b));
  return *reinterpret_cast<const uint32_t*>(&r);
}
__device__ __forceinline__ uint32_t bf16x2_eq_mask(uint32_t a, uint32_t b) {
  return __heq2_mask(*reinterpret_cast<const __nv_bfloat162*>(&a), *reinterpret_cast<const __nv_bfloat162*>(&b));
}

__global__ void __launch_bounds__(256) stem_bn_relu_pool_kernel(const __nv_bfloat16* __restrict__ y,
                                                                const double* sum, const double* sumsq,
                                                                const float* gamma, const float* beta,
                                                                __nv_bfloat16* __restrict__ out,
                                                                uint8_t* __restrict__ idx_out, float* mean_out,
                                                                float* invstd_out, float* running_mean,
                                                                float* running_var, int N, int H, int W, float eps,
                                                                float momentum, FastDiv fd_wo, FastDiv fd_ho) {
  pdl_launch_dependents();
  pdl_wait();
  const int C = 64, tpr = 8;
  const int cg = threadIdx.x % tpr, roff = threadIdx.x / tpr;
  const int c0 = cg * 8;
  const long long M = (long long)N * H * W;
  float mean[8], var[8], invstd[8], scale[8], shift[8];
  bn_coeffs(sum, sumsq, c0, 1.0 / (double)M, eps, mean, var, invstd);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    scale[i] = gamma[c0 + i] * invstd[i];
    shift[i] = beta[c0 + i] - mean[i] * scale[i];
  }
  if (blockIdx.x == 0 && roff == 0)
    bn_bookkeeping(mean, var, invstd, c0, M, momentum, mean_out, invstd_out, running_mean, running_var);
  uint32_t flip[4];  // sign-bit flips that make "larger raw value" mean "larger activation"
#pragma unroll
  for (int j = 0; j < 4; ++j)
    flip[j] = (scale[2 * j] < 0.f ? 0x00008000u : 0u) | (scale[2 * j + 1] < 0.f ? 0x80000000u : 0u);
  const uint32_t kNegInf = 0xFF80FF80u;
  const int Ho = H / 2, Wo = W / 2;
  // output pixels as one flat range over the grid, 32 pixel lanes x 8 channel groups per block (a whole output row
  // per block iteration left 12 % of the lanes idle at Wo = 56 and whole blocks idle at the end)
  const int items = N * Ho * Wo;
  {
#pragma unroll 2
    for (int q = blockIdx.x * 32 + roff; q < items; q += gridDim.x * 32) {
      const int row = fd_div(fd_wo, q), wo = q - row * Wo;
      const int n = fd_div(fd_ho, row), ho = row - n * Ho;
      const bool top = ho == 0;  // input row 2ho - 1 is padding (2ho + 1 <= H - 1 always: H is even)
      const __nv_bfloat16* img = y + ((size_t)n * H + (top ? 0 : 2 * ho - 1)) * W * C + c0;
      const size_t rstep1 = top ? 0 : (size_t)W * C;  // rows 0 and 1 of the window coincide (clamped) on the top edge
      const bool left = wo == 0;
      const __nv_bfloat16* p0 = img + (size_t)(left ? 0 : 2 * wo - 1) * C;
      const size_t cstep1 = left ? 0 : (size_t)C;
      bf16x8 win[9];
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        const __nv_bfloat16* pr = p0 + (r == 0 ? 0 : rstep1 + (size_t)(r - 1) * W * C);
        win[r * 3 + 0] = ld8_cached(pr);
        win[r * 3 + 1] = ld8_cached(pr + cstep1);
        win[r * 3 + 2] = ld8_cached(pr + cstep1 + C);
      }
      uint32_t best[4], bidx[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint32_t t[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) {
          t[k] = win[k].v[j] ^ flip[j];
          if ((k < 3 && top) || (k % 3 == 0 && left)) t[k] = kNegInf;  // padding never wins (the centre is valid)
        }
        uint32_t m = t[0];
#pragma unroll
        for (int k = 1; k < 9; ++k) m = bf16x2_max(m, t[k]);
        uint32_t id = 0x00080008u;
#pragma unroll
        for (int k = 7; k >= 0; --k) {  // the first position holding the maximum wins
          const uint32_t eq = bf16x2_eq_mask(t[k], m);
          id = (id & ~eq) | ((uint32_t)k * 0x00010001u & eq);
        }
        best[j] = m ^ flip[j];
        bidx[j] = id;
      }
      bf16x8 o;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float a0 = fmaxf(fmaf(bf16_lo(best[j]), scale[2 * j], shift[2 * j]), 0.f);
        const float a1 = fmaxf(fmaf(bf16_hi(best[j]), scale[2 * j + 1], shift[2 * j + 1]), 0.f);
        o.v[j] = pack_bf16x2(a0, a1);
      }
      const size_t oo = ((size_t)row * Wo + wo) * C + c0;
      st8(out + oo, o);
      if (idx_out) {
        uint2 packed;  // bytes 0 and 2 of each 16-bit-lane word
        packed.x = __byte_perm(bidx[0], bidx[1], 0x6420);
        packed.y = __byte_perm(bidx[2], bidx[3], 0x6420);
        *reinterpret_cast<uint2*>(idx_out + oo) = packed;
      }
    }
  }
}

// backward of the above up to (and excluding) the BN-backward apply: g[pixel] = relu'(.) * sum of dpool over the
// (at most 4) windows whose recorded winner is this pixel; also accumulates the BN-backward sums.
// A thread owns 8 channels of the pixel PAIR (2j, 2j + 1) of one input row: the even pixel lies in window column j
// only (as its middle tap), the odd one in columns j (right tap) and j + 1 (left tap); even rows lie in one window
// row, odd rows in two (uniform per block iteration).  Winner tests are byte-parallel (__vcmpeq4 on 4 channels).
__device__ __forceinline__ void pool_bwd_accumulate(const bf16x8& dp, uint2 widx, uint32_t kpos, float (&g)[8]) {
  const uint32_t want = kpos * 0x01010101u;
  const uint32_t m0 = __vcmpeq4(widx.x, want), m1 = __vcmpeq4(widx.y, want);  // 0xFF per matching channel byte
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const uint32_t mb = j < 2 ? m0 : m1;
    const uint32_t lanes = __byte_perm(mb, 0u, (j & 1) ? 0x3322 : 0x1100);  // byte mask -> 16-bit lane mask
    const uint32_t v = dp.v[j] & lanes;
    g[2 * j] += bf16_lo(v);
    g[2 * j + 1] += bf16_hi(v);
  }
}

// One work item = one pixel pair (2j, 2j + 1) of one input row, 8 channels per thread.  Items are dealt to the grid
// as one flat range and every thread keeps TWO of them in flight (all loads of both are issued before either is
// consumed): with one row per block iteration the kernel was bound by the latency of its two dependent memory round
// trips per row (ncu: 42 % of the stall samples on the first use of a load at 23 % occupancy, 13 % at the exit of
// blocks whose rows ran out early).
struct PoolBwdItem {
  bf16x8 y0, y1, dpa, dpc, dqa, dqc;
  uint2 ia, ic, ja, jc;
  size_t o;
  uint32_t krow_a;
  bool live, has_b, has_c;
};

__device__ __forceinline__ void pool_bwd_load(PoolBwdItem& t, int q, int items, const __nv_bfloat16* __restrict__ dpool,
                                              const uint8_t* __restrict__ idx, const __nv_bfloat16* __restrict__ y,
                                              int H, int W, FastDiv fd_wo, FastDiv fd_h, int c0) {
  const int C = 64, Ho = H / 2, Wo = W / 2;
  t.live = q < items;
  if (!t.live) return;
  const int row = fd_div(fd_wo, q), j = q - row * Wo;
  const int n = fd_div(fd_h, row), h = row - n * H;
  // windows (ho, wo) with 2ho-1 <= h <= 2ho+1: ho in {h>>1, (h+1)>>1}; odd rows belong to two window rows (if the
  // second one exists), and h is then the top tap (kernel row 0) of window row ho_b
  const int ho_a = h >> 1, ho_b = (h + 1) >> 1;
  t.has_b = (h & 1) && ho_b < Ho;
  t.has_c = j + 1 < Wo;
  t.krow_a = (uint32_t)(h - (2 * ho_a - 1)) * 3u;
  t.o = ((size_t)row * W + 2 * j) * C + c0;
  t.y0 = ld8(y + t.o), t.y1 = ld8(y + t.o + C);
  const size_t pa0 = (((size_t)n * Ho + ho_a) * Wo + j) * C + c0, pa1 = pa0 + (t.has_c ? C : 0);
  t.dpa = ld8_cached(dpool + pa0), t.dpc = ld8_cached(dpool + pa1);
  t.ia = *reinterpret_cast<const uint2*>(idx + pa0), t.ic = *reinterpret_cast<const uint2*>(idx + pa1);
  if (t.has_b) {
    const size_t pb0 = (((size_t)n * Ho + ho_b) * Wo + j) * C + c0, pb1 = pb0 + (t.has_c ? C : 0);
    t.dqa = ld8_cached(dpool + pb0), t.dqc = ld8_cached(dpool + pb1);
    t.ja = *reinterpret_cast<const uint2*>(idx + pb0), t.jc = *reinterpret_cast<const uint2*>(idx + pb1);
  }
}

__device__ __forceinline__ void pool_bwd_finish(const PoolBwdItem& t, const float (&scale)[8], const float (&shift)[8],
                                                float (&s)[8], float (&d)[8], __nv_bfloat16* __restrict__ g_out) {
  if (!t.live) return;
  float g0[8], g1[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) g0[i] = g1[i] = 0.f;
  pool_bwd_accumulate(t.dpa, t.ia, t.krow_a + 1u, g0);               // even pixel: middle tap of column j
  pool_bwd_accumulate(t.dpa, t.ia, t.krow_a + 2u, g1);               // odd pixel: right tap of column j
  pool_bwd_accumulate(t.dpc, t.ic, t.has_c ? t.krow_a : 0xFFu, g1);  // odd pixel: left tap of column j + 1
  if (t.has_b) {
    pool_bwd_accumulate(t.dqa, t.ja, 1u, g0);
    pool_bwd_accumulate(t.dqa, t.ja, 2u, g1);
    pool_bwd_accumulate(t.dqc, t.jc, t.has_c ? 0u : 0xFFu, g1);
  }
  const int C = 64;
  float yv[8];
  unpack8(t.y0, yv);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    g0[i] = fmaf(yv[i], scale[i], shift[i]) > 0.f ? g0[i] : 0.f;
    s[i] += g0[i];
    d[i] = fmaf(g0[i], yv[i], d[i]);
  }
  st8(g_out + t.o, pack8(g0));
  unpack8(t.y1, yv);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    g1[i] = fmaf(yv[i], scale[i], shift[i]) > 0.f ? g1[i] : 0.f;
    s[i] += g1[i];
    d[i] = fmaf(g1[i], yv[i], d[i]);
  }
  st8(g_out + t.o + C, pack8(g1));
}

__global__ void __launch_bounds__(256, 2) stem_pool_bwd_kernel(const __nv_bfloat16* __restrict__ dpool,
                                                               const uint8_t* __restrict__ idx,
                                                               const __nv_bfloat16* __restrict__ y,
                                                               const float* __restrict__ mean,
                                                               const float* __restrict__ invstd,
                                                               const float* __restrict__ gamma,
                                                               const float* __restrict__ beta,
                                                               __nv_bfloat16* __restrict__ g_out, double* scratch, int N,
                                                               int H, int W, FastDiv fd_wo, FastDiv fd_h) {
  __shared__ float red[2][256][9];
  pdl_launch_dependents();
  pdl_wait();
  const int C = 64, tpr = 8;
  const int cg = threadIdx.x % tpr, roff = threadIdx.x / tpr;
  const int c0 = cg * 8;
  float scale[8], shift[8], s[8], d[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    scale[i] = gamma[c0 + i] * invstd[c0 + i];
    shift[i] = beta[c0 + i] - mean[c0 + i] * scale[i];
    s[i] = d[i] = 0.f;
  }
  const int items = N * H * (W / 2);
  const int stride = gridDim.x * 32;
  for (int q = blockIdx.x * 32 + roff; q < items; q += 2 * stride) {
    PoolBwdItem t0, t1;
    pool_bwd_load(t0, q, items, dpool, idx, y, H, W, fd_wo, fd_h, c0);
    pool_bwd_load(t1, q + stride, items, dpool, idx, y, H, W, fd_wo, fd_h, c0);
    pool_bwd_finish(t0, scale, shift, s, d, g_out);
    pool_bwd_finish(t1, scale, shift, s, d, g_out);
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) red[0][threadIdx.x][i] = s[i], red[1][threadIdx.x][i] = d[i];
  __syncthreads();
  if (roff == 0) {
    for (int rr = 1; rr < 32; ++rr) {
#pragma unroll
      for (int i = 0; i < 8; ++i) s[i] += red[0][rr * tpr + cg][i], d[i] += red[1][rr * tpr + cg][i];
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      red_add_f64(scratch + c0 + i, (double)s[i]);
      red_add_f64(scratch + C + c0 + i, (double)d[i]);
    }
  }
}

// ---- global average pool: x [N][HW][C] bf16 -> out [N][C] fp32 ; and its backward -------------------
__global__ void __launch_bounds__(256) avgpool_fwd_kernel(const __nv_bfloat16* __restrict__ x, float* out, int N,
                                                          int HW, int C) {
  pdl_launch_dependents();
  pdl_wait();
  const int tpr = C >> 3;
  const long long total = (long long)N * tpr;
  const float inv = 1.f / (float)HW;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const int cg = (int)(t % tpr);
    const long long n = t / tpr;
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.f;
    for (int p = 0; p < HW; ++p) {
      float v[8];
      unpack8(ld8(x + (n * HW + p) * C + cg * 8), v);
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] += v[i];
    }
    float4* o = reinterpret_cast<float4*>(out + n * C + cg * 8);
    o[0] = make_float4(acc[0] * inv, acc[1] * inv, acc[2] * inv, acc[3] * inv);
    o[1] = make_float4(acc[4] * inv, acc[5] * inv, acc[6] * inv, acc[7] * inv);
  }
}

__global__ void __launch_bounds__(256) avgpool_bwd_kernel(const float* __restrict__ dout, __nv_bfloat16* dx, int N,
                                                          int HW, int C) {
  pdl_launch_dependents();
  pdl_wait();
  const int tpr = C >> 3;
  const long long total = (long long)N * HW * tpr;
  const float inv = 1.f / (float)HW;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const int cg = (int)(t % tpr);
    const long long row = t / tpr;
    const long long n = row / HW;
    const float4* g = reinterpret_cast<const float4*>(dout + n * C + cg * 8);
    const float4 a = g[0], b = g[1];
    const float v[8] = {a.x * inv, a.y * inv, a.z * inv, a.w * inv, b.x * inv, b.y * inv, b.z * inv, b.w * inv};
    st8(dx + row * C + cg * 8, pack8(v));
  }
}

// ---- input transform: two fp32 NCHW view batches -> one zero-padded, space-to-depth bf16 batch -------------
// out [2B][H/2 + 3][W/2 + 4][16]: block (Y, X) of 2 x 2 image pixels is stored at (Y + 2, X + 2) with its 12 values
// in channel order (dy, dx, c) -> dy * 6 + dx * 3 + c; channels 12..15 and the border are zero.  The 7x7 / stride 2
// stem then is a 4 x 4 / stride 1 convolution over 16-channel pixels (csrc/conv_ops.cu, stem_views).
__global__ void __launch_bounds__(256) stem_input_kernel(const float* __restrict__ x1, const float* __restrict__ x2,
                                                         __nv_bfloat16* __restrict__ out, int B1, int Ntot, int H,
                                                         int W) {
  pdl_launch_dependents();
  pdl_wait();
  const int Hs = H / 2 + 3, Ws = W / 2 + 4;
  const size_t cs = (size_t)H * W;
  // one block iteration = one padded row (n, yp); threads walk the padded blocks (coalesced float2 plane reads)
  for (int row = blockIdx.x; row < Ntot * Hs; row += gridDim.x) {
    const int n = row / Hs, yp = row - n * Hs;
    const int y = yp - 2;
    const bool row_ok = y >= 0 && y < H / 2;
    // (images B1 .. 2 B1 - 1 come from x2; with x2 == NULL the batch is x1 alone)
    const float* src = (n < B1 ? x1 + (size_t)n * 3 * cs : x2 + (size_t)(n - B1) * 3 * cs) +
                       (size_t)(row_ok ? 2 * y : 0) * W;
    uint4* dst = reinterpret_cast<uint4*>(out) + (size_t)row * Ws * 2;
    for (int xp = threadIdx.x; xp < Ws; xp += 256) {
      const int x = xp - 2;
      uint4 lo = make_uint4(0u, 0u, 0u, 0u), hi = make_uint4(0u, 0u, 0u, 0u);
      if (row_ok && x >= 0 && x < W / 2) {
        float2 v[2][3];  // [dy][c] = pixels (2x, 2x + 1) of image row 2y + dy, plane c
#pragma unroll
        for (int dy = 0; dy < 2; ++dy)
#pragma unroll
          for (int c = 0; c < 3; ++c)
            v[dy][c] = *reinterpret_cast<const float2*>(src + c * cs + (size_t)dy * W + 2 * x);
        // channel order: (dy, dx, c)
        lo.x = pack_bf16x2(v[0][0].x, v[0][1].x);
        lo.y = pack_bf16x2(v[0][2].x, v[0][0].y);
        lo.z = pack_bf16x2(v[0][1].y, v[0][2].y);
        lo.w = pack_bf16x2(v[1][0].x, v[1][1].x);
        hi.x = pack_bf16x2(v[1][2].x, v[1][0].y);
        hi.y = pack_bf16x2(v[1][1].y, v[1][2].y);
      }
      dst[2 * xp] = lo;
      dst[2 * xp + 1] = hi;
    }
  }
}

static int num_sms() {
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
  }
  return sms;
}
// grid for a row-walking kernel: enough blocks for `rows` at `rows_per_block * unroll` rows per block
// iteration, capped at `waves` resident blocks per SM.
static int grid_for(long long work_items, int per_block, int waves = 2) {
  static int env_waves = -1;  // tuning knob: PECLR_ELT_WAVES overrides the resident-blocks-per-SM cap
  if (env_waves < 0) {
    const char* e = getenv("PECLR_ELT_WAVES");
    env_waves = e ? atoi(e) : 0;
  }
  if (env_waves > 0) waves = env_waves;
  long long blocks = (work_items + per_block - 1) / per_block;
  const long long cap = (long long)num_sms() * waves;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}
static int last_error() {
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : -(int)e;
}
static bool bad_channels(int C) { return C % 8 || C > 2048 || C < 8 || (256 % (C / 8)) != 0; }

}  // namespace peclr

using namespace peclr;
typedef __nv_bfloat16 bf16;

extern "C" int peclr_bn_apply(const void* y, const double* sum, const double* sumsq, const float* gamma,
                              const float* beta, const void* res, const double* rsum, const double* rsumsq,
                              const float* rgamma, const float* rbeta, void* out, void* mask_out, float* mean_out,
                              float* invstd_out,
                              float* running_mean, float* running_var, float* rmean_out, float* rinvstd_out,
                              float* rrunning_mean, float* rrunning_var, long long M, int C, float eps, float momentum,
                              int relu, void* stream) {
  if (bad_channels(C)) return -1001;
  BnApplyArgs a{(const bf16*)y, sum,      sumsq,      gamma,        beta,         (const bf16*)res, rsum,
                rsumsq,         rgamma,   rbeta,      (bf16*)out,   (uint8_t*)mask_out, mean_out,     invstd_out,       running_mean,
                running_var,    rmean_out, rinvstd_out, rrunning_mean, rrunning_var, M,              C,
                eps,            momentum, relu, elt_contig()};
  const int rows_per_block = 256 / (C / 8);
  const int kr = elt_rows();
  const int grid = grid_for(M, rows_per_block * kr);
  if (kr == 2) launch_pdl(bn_apply_kernel<2>, grid, 256, 0, (cudaStream_t)stream, a);
  else launch_pdl(bn_apply_kernel<4>, grid, 256, 0, (cudaStream_t)stream, a);
  return last_error();
}

extern "C" int peclr_bn_bwd_reduce(const void* dout, const void* mask, const void* y, const float* mean,
                                   const float* invstd, const float* gamma, const float* beta, int mask_mode,
                                   double* scratch, long long M, int C, void* stream) {
  if (bad_channels(C) || mask_mode < 0 || mask_mode > 3 || ((mask_mode & 1) && !mask)) return -1001;
  cudaError_t e = cudaMemsetAsync(scratch, 0, (size_t)2 * C * sizeof(double), (cudaStream_t)stream);
  if (e != cudaSuccess) return -(int)e;
  BnBwdArgs a{(const bf16*)dout, (const bf16*)mask, (const bf16*)y, mean, invstd, gamma, beta, scratch,
              nullptr,           nullptr,           nullptr,        nullptr, M,    C,     mask_mode, elt_contig()};
  const int rows_per_block = 256 / (C / 8);
  launch_bn_bwd<false>(a, grid_for(M, rows_per_block * elt_rows() * 2), (cudaStream_t)stream);
  return last_error();
}

extern "C" int peclr_bn_bwd_apply(const void* dout, const void* mask, const void* y, const float* mean,
                                  const float* invstd, const float* gamma, const float* beta, int mask_mode,
                                  const double* scratch, void* dy, void* g_out, float* dgamma, float* dbeta,
                                  long long M, int C, void* stream) {
  if (bad_channels(C) || mask_mode < 0 || mask_mode > 3 || ((mask_mode & 1) && !mask)) return -1001;
  BnBwdArgs a{(const bf16*)dout, (const bf16*)mask, (const bf16*)y, mean,   invstd, gamma, beta, const_cast<double*>(scratch),
              (bf16*)dy,         (bf16*)g_out,      dgamma,         dbeta,  M,      C,     mask_mode, elt_contig()};
  const int rows_per_block = 256 / (C / 8);
  launch_bn_bwd<true>(a, grid_for(M, rows_per_block * elt_rows()), (cudaStream_t)stream);
  return last_error();
}

extern "C" int peclr_stem_bn_relu_pool(const void* y, const double* sum, const double* sumsq, const float* gamma,
                                       const float* beta, void* out, void* idx_out, float* mean_out,
                                       float* invstd_out, float* running_mean, float* running_var, int N, int H,
                                       int W, float eps, float momentum, void* stream) {
  if ((H & 1) || (W & 1)) return -1001;
  if ((long long)N * (H / 2) * (W / 2) >= (1ll << 30)) return -1001;
  launch_pdl(stem_bn_relu_pool_kernel, grid_for((long long)N * (H / 2) * (W / 2), 32, 3), 256, 0, (cudaStream_t)stream,
             (const bf16*)y, sum, sumsq, gamma, beta, (bf16*)out, (uint8_t*)idx_out, mean_out, invstd_out, running_mean,
             running_var, N, H, W, eps, momentum, make_fastdiv(W / 2), make_fastdiv(H / 2));
  return last_error();
}

extern "C" int peclr_stem_pool_bwd(const void* dpool, const void* idx, const void* y, const float* mean,
                                   const float* invstd, const float* gamma, const float* beta, void* g_out,
                                   double* scratch, int N, int H, int W, void* stream) {
  if ((H & 1) || (W & 1) || !idx) return -1001;
  cudaError_t e = cudaMemsetAsync(scratch, 0, (size_t)2 * 64 * sizeof(double), (cudaStream_t)stream);
  if (e != cudaSuccess) return -(int)e;
  if ((long long)N * H * (W / 2) >= (1ll << 30)) return -1001;
  launch_pdl(stem_pool_bwd_kernel, grid_for((long long)N * H * (W / 2), 64, 2), 256, 0, (cudaStream_t)stream,
             (const bf16*)dpool, (const uint8_t*)idx, (const bf16*)y, mean, invstd, gamma, beta, (bf16*)g_out, scratch,
             N, H, W, make_fastdiv(W / 2), make_fastdiv(H));
  return last_error();
}

extern "C" int peclr_avgpool_fwd(const void* x, float* out, int N, int HW, int C, void* stream) {
  if (C % 8) return -1001;
  launch_pdl(avgpool_fwd_kernel, grid_for((long long)N * (C / 8), 256), 256, 0, (cudaStream_t)stream, (const bf16*)x,
             out, N, HW, C);
  return last_error();
}

extern "C" int peclr_avgpool_bwd(const float* dout, void* dx, int N, int HW, int C, void* stream) {
  if (C % 8) return -1001;
  launch_pdl(avgpool_bwd_kernel, grid_for((long long)N * HW * (C / 8), 256), 256, 0, (cudaStream_t)stream, dout,
             (bf16*)dx, N, HW, C);
  return last_error();
}

extern "C" int peclr_stem_input(const float* x1, const float* x2, void* out, int B, int H, int W, void* stream) {
  if ((H & 1) || (W & 1) || B < 1 || !x1) return -1001;
  const int ntot = x2 ? 2 * B : B;  // x2 == NULL: a plain batch of B images (inference / odd batch sizes)
  launch_pdl(stem_input_kernel, grid_for((long long)ntot * (H / 2 + 3), 1, 8), 256, 0, (cudaStream_t)stream, x1, x2,
             (bf16*)out, B, ntot, H, W);
  return last_error();
}
