// Fused PeCLR loss chain, forward AND backward in ONE launch (fp32):
//
//   p (2B x 128 projection-head output)  ->  L2-normalise -> inverse translate -> inverse rotate ->
//   L2-normalise -> z  ->  NT-Xent over the (global) 2N batch  ->  loss, dloss/dp, 16 projection stats
//
// Reference: Hybrid2Model.get_transformed_projections / contrastive_step
// (src/models/unsupervised/hybrid2_model.py:27-106), translate_encodings / rotate_encoding /
// get_rotation_2D_matrix / vanila_contrastive_loss (src/models/utils.py:154-186, 271-346).  The reference
// runs ~25 ATen launches, materialises five 2N x 2N temporaries and round-trips the rotation matrix
// through the CPU every step; here the 2N x 2N similarity matrix only ever exists as 32 x 64 tiles in
// shared memory / registers.
//
// The op is latency-bound (262 KB of algorithmic traffic at 2N = 256), so it is one cooperative launch
// with software grid barriers between its phases:
//   P1  one warp per local row: statistics, both normalisations, translate, rotate; z rows are written
//       to every rank's z buffer (peer stores over NVLink when world > 1 -> the all-gather is fused).
//   P2  (row-block x column-block) units: S = z z^T / T tile, exp, row sums "neg" (self excluded), S_pos.
//   P3  units again: coefficients c_ik = [e_ik (1/neg_i + 1/neg_k) - 2 [k = pos(i)]] / (T n), g_z += c z.
//   P4  one warp per local row: loss terms and the chain rule back to p.
// Every rank recomputes neg for all global rows, so the backward needs no second collective
// (SURVEY.md section 8(e)); gradients of the GLOBAL mean loss w.r.t. local rows come out complete.
//
// Reproducibility: nothing is accumulated with fp32 atomics.  Row sums, the loss and the statistics are added as
// fp64 atomics (per-tile / per-block / per-row partials computed in a fixed order; their fp64 totals do not depend
// on the arrival order), the gradient partials of the column chunks go to separate slabs that P4 adds in order.
#include <math.h>

#include "../../include/peclr_b200.h"
#include "ptx.cuh"

namespace peclr {

constexpr int D = 128;        // projection width (64 interleaved 2-D points)
constexpr int RB = 32;        // rows per unit
constexpr int CB = 64;        // columns per tile
constexpr int LDS = D + 4;    // padded smem row (floats)

struct NtxentArgs {
  const float* p;          // [2B][128] local projections (view 1 rows, then view 2 rows)
  const double* angle;     // [2B] degrees, may be null when rotate == 0
  const long long* jx;     // [2B] pixels
  const long long* jy;
  float* z;                // [n_glob][128] this rank's copy of the gathered embeddings
  float* const* z_peers;   // [world] every rank's z buffer (device pointers), null when world == 1
  unsigned* const* flag_peers;  // [world] every rank's flag array [world], null when world == 1
  float* rowbuf;           // [2B][4]: 1/|p|, 1/|r|, alpha, beta
  double* neg;             // [n_glob] fp64 accumulators of the row sums
  double* acc;             // [17]: loss, 16 statistics (fp64 accumulators)
  float* spos;             // [n_glob]
  float* gz;               // [chunks][2B][128]: one slab per P3 column chunk, added in order by P4
  float* loss;             // [1]
  float* stats;            // [16] proj1 x{mean,median,min,max}, y{...}, proj2 ...
  float* g_p;              // [2B][128], null for forward only
  unsigned* barrier;       // [2] zeroed by the host before every launch: grid barrier, blocks done with the loss
  int B, world, rank;
  int img_h, img_w, crop, rotate;
  float inv_t;
  unsigned* launch_ctr;    // [1] launches so far (device side, so a captured CUDA graph can be replayed)
  int col_chunk;           // columns per P3 unit (multiple of CB)
  int plain;               // 1: p already holds z (plain NT-Xent, no normalisation / correction / stats)
};

__device__ __forceinline__ void grid_barrier(unsigned* counter, unsigned target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(counter, 1u);
    unsigned spins = 0;
    while (*reinterpret_cast<volatile unsigned*>(counter) < target) {
      if (++spins > (1u << 28)) __trap();
    }
    __threadfence();
  }
  __syncthreads();
}

__device__ __forceinline__ int global_row(int i, int B, int world, int rank) {
  return i < B ? rank * B + i : world * B + rank * B + (i - B);
}

// rank of v among the 64 values held as (a, b) per lane, ties broken by position -> lower median pick
__device__ __forceinline__ float warp_lower_median64(float a, float b, int lane) {
  int ra = 0, rb = 0;
  for (int k = 0; k < 32; ++k) {
    const float oa = __shfl_sync(0xffffffffu, a, k), ob = __shfl_sync(0xffffffffu, b, k);
    // position of a in lane l is 2l, of b is 2l + 1
    ra += (oa < a || (oa == a && 2 * k < 2 * lane)) + (ob < a || (ob == a && 2 * k + 1 < 2 * lane));
    rb += (oa < b || (oa == b && 2 * k < 2 * lane + 1)) + (ob < b || (ob == b && 2 * k + 1 < 2 * lane + 1));
  }
  float m = ra == 31 ? a : (rb == 31 ? b : 0.f);
  return warp_sum(m);  // exactly one lane contributes
}

__global__ void __launch_bounds__(256, 1) ntxent_fused_kernel(const NtxentArgs a) {
  extern __shared__ float sm[];
  float* sI = sm;                 // [RB][LDS]
  float* sK = sm + RB * LDS;      // [CB][LDS]
  float* sC = sK + CB * LDS;      // [RB][CB + 1] coefficients
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n_loc = 2 * a.B;
  const int n_glob = n_loc * a.world;
  const int gwarp = blockIdx.x * 8 + warp, nwarps = gridDim.x * 8;
  unsigned bar_target = 0;
  // launch number (1-based).  With world > 1 it tags the peer flags and selects one of two z buffers, so a fast
  // rank's next step never overwrites embeddings a slower rank is still reading.
  const unsigned epoch = *a.launch_ctr + 1;
  const size_t zsel = a.world > 1 ? (size_t)(epoch & 1) * (size_t)(2 * a.B * a.world) * D : 0;
  float* const zbuf = a.z + zsel;

  // ---------------------------------------------------------------- P1
  if (blockIdx.x == 0 && threadIdx.x < 17) a.acc[threadIdx.x] = 0.0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_glob; i += gridDim.x * blockDim.x) a.neg[i] = 0.0;
  for (int i = gwarp; i < n_loc; i += nwarps) {
    const float4 v = reinterpret_cast<const float4*>(a.p + (size_t)i * D)[lane];  // points 2l, 2l+1
    if (a.plain) {
      reinterpret_cast<float4*>(zbuf + (size_t)global_row(i, a.B, a.world, a.rank) * D)[lane] = v;
      continue;
    }
    float x0 = v.x, y0 = v.y, x1 = v.z, y1 = v.w;
    // projection statistics of the raw head output (hybrid2_model.py:92-106), detached
    {
      const float mx = warp_sum(x0 + x1) * (1.f / 64), my = warp_sum(y0 + y1) * (1.f / 64);
      const float lox = warp_min(fminf(x0, x1)), hix = warp_max(fmaxf(x0, x1));
      const float loy = warp_min(fminf(y0, y1)), hiy = warp_max(fmaxf(y0, y1));
      const float medx = warp_lower_median64(x0, x1, lane), medy = warp_lower_median64(y0, y1, lane);
      if (lane < 8) {  // fp64 sums over the batch (block 0 converts them in P4)
        const float v8 = lane == 0 ? mx : lane == 1 ? medx : lane == 2 ? lox : lane == 3 ? hix
                       : lane == 4 ? my : lane == 5 ? medy : lane == 6 ? loy : hiy;
        atomicAdd(a.acc + 1 + (i < a.B ? 0 : 8) + lane, (double)v8);
      }
    }
    const float pn = sqrtf(warp_sum(x0 * x0 + y0 * y0 + x1 * x1 + y1 * y1));
    const float inv_p = 1.f / fmaxf(pn, 1e-12f);
    x0 *= inv_p, y0 *= inv_p, x1 *= inv_p, y1 *= inv_p;
    if (a.crop) {
      // tx = -(jitter_x / H), ty = -(jitter_y / W) in fp32 (hybrid2_model.py:59-74)
      const float tx = -((float)a.jx[i] / (float)a.img_h), ty = -((float)a.jy[i] / (float)a.img_w);
      const float rx = warp_max(fmaxf(x0, x1)) - warp_min(fminf(x0, x1));
      const float ry = warp_max(fmaxf(y0, y1)) - warp_min(fminf(y0, y1));
      const float dx = tx * rx, dy = ty * ry;
      x0 += dx, x1 += dx, y0 += dy, y1 += dy;
    }
    float al = 1.f, be = 0.f;
    if (a.rotate) {
      // angle' = -angle; trig and offsets in fp64, rounded into an fp32 matrix (utils.py:287-296)
      const double ang = -a.angle[i] * 3.141592653589793 / 180.0;
      const double ca = cos(ang), sa = sin(ang);
      const float cx = warp_sum(x0 + x1) / 64.f, cy = warp_sum(y0 + y1) / 64.f;
      const float offx = (float)((1.0 - ca) * (double)cx - sa * (double)cy);
      const float offy = (float)((1.0 - ca) * (double)cy + sa * (double)cx);
      al = (float)ca, be = (float)sa;
      const float nx0 = fmaf(al, x0, fmaf(be, y0, offx)), ny0 = fmaf(-be, x0, fmaf(al, y0, offy));
      const float nx1 = fmaf(al, x1, fmaf(be, y1, offx)), ny1 = fmaf(-be, x1, fmaf(al, y1, offy));
      x0 = nx0, y0 = ny0, x1 = nx1, y1 = ny1;
    }
    const float rn = sqrtf(warp_sum(x0 * x0 + y0 * y0 + x1 * x1 + y1 * y1));
    const float inv_r = 1.f / fmaxf(rn, 1e-12f);
    const float4 zv = make_float4(x0 * inv_r, y0 * inv_r, x1 * inv_r, y1 * inv_r);
    const int gi = global_row(i, a.B, a.world, a.rank);
    if (a.world == 1) {
      reinterpret_cast<float4*>(zbuf + (size_t)gi * D)[lane] = zv;
    } else {
      for (int r = 0; r < a.world; ++r)
        reinterpret_cast<float4*>(a.z_peers[r] + zsel + (size_t)gi * D)[lane] = zv;
    }
    if (lane == 0) reinterpret_cast<float4*>(a.rowbuf)[i] = make_float4(inv_p, inv_r, al, be);
  }
  bar_target += gridDim.x;
  grid_barrier(a.barrier, bar_target);
  if (a.world > 1) {
    // all local z rows are out (peer stores): publish to every rank, then wait for every rank
    if (blockIdx.x == 0 && threadIdx.x < a.world) {
      __threadfence_system();
      unsigned* remote = a.flag_peers[threadIdx.x] + a.rank;
      asm volatile("st.release.sys.global.u32 [%0], %1;\n" ::"l"(remote), "r"(epoch) : "memory");
      const unsigned* mine = a.flag_peers[a.rank] + threadIdx.x;
      unsigned v = 0, spins = 0;
      do {
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];\n" : "=r"(v) : "l"(mine) : "memory");
        if (++spins > (1u << 28)) __trap();
      } while (v < epoch);
    }
    bar_target += gridDim.x;
    grid_barrier(a.barrier, bar_target);
  }

  // ---------------------------------------------------------------- P2: neg / spos for ALL global rows
  const int half = n_glob / 2;
  {
    const int rbs = (n_glob + RB - 1) / RB, cbs = (n_glob + CB - 1) / CB;
    const int tr = threadIdx.x >> 4, tc = threadIdx.x & 15;  // rows 2*tr.., cols tc + 16*j
    for (int u = blockIdx.x; u < rbs * cbs; u += gridDim.x) {
      const int i0 = (u / cbs) * RB, k0 = (u % cbs) * CB;
      __syncthreads();
      for (int t = threadIdx.x; t < (RB + CB) * (D / 4); t += 256) {
        const int row = t / (D / 4), c4 = t % (D / 4);
        const int g = row < RB ? i0 + row : k0 + (row - RB);
        const float4 v = g < n_glob ? reinterpret_cast<const float4*>(zbuf + (size_t)g * D)[c4]
                                    : make_float4(0.f, 0.f, 0.f, 0.f);
        float* dst = (row < RB ? sI + row * LDS : sK + (row - RB) * LDS) + c4 * 4;
        *reinterpret_cast<float4*>(dst) = v;
      }
      __syncthreads();
      float acc[2][4] = {};
#pragma unroll 4
      for (int d = 0; d < D; d += 4) {
        float4 zi[2], zk[4];
#pragma unroll
        for (int r = 0; r < 2; ++r) zi[r] = *reinterpret_cast<const float4*>(sI + (2 * tr + r) * LDS + d);
#pragma unroll
        for (int c = 0; c < 4; ++c) zk[c] = *reinterpret_cast<const float4*>(sK + (tc + 16 * c) * LDS + d);
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
          for (int c = 0; c < 4; ++c)
            acc[r][c] += zi[r].x * zk[c].x + zi[r].y * zk[c].y + zi[r].z * zk[c].z + zi[r].w * zk[c].w;
      }
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int gi = i0 + 2 * tr + r;
        float rowsum = 0.f;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int gk = k0 + tc + 16 * c;
          const float s = acc[r][c] * a.inv_t;
          if (gi < n_glob && gk < n_glob) {
            if (gk != gi) rowsum += __expf(s);
            if (gk == (gi + half) % n_glob) a.spos[gi] = s;
          }
        }
        // the 16 threads sharing a row are consecutive lanes of a half warp
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) rowsum += __shfl_xor_sync(0xffffffffu, rowsum, o);
        if (tc == 0 && gi < n_glob) atomicAdd(a.neg + gi, (double)rowsum);
      }
    }
  }
  bar_target += gridDim.x;
  grid_barrier(a.barrier, bar_target);

  if (blockIdx.x == 0 && threadIdx.x == 0) *a.launch_ctr = epoch;  // every block read it before the barriers

  // ---------------------------------------------------------------- P3: g_z for LOCAL rows
  if (a.g_p) {
    const int rbs = (n_loc + RB - 1) / RB;
    const int chunks = (n_glob + a.col_chunk - 1) / a.col_chunk;
    const float scale = a.inv_t / (float)n_glob;
    const int tr = threadIdx.x >> 4, tc = threadIdx.x & 15;
    for (int u = blockIdx.x; u < rbs * chunks; u += gridDim.x) {
      const int i0 = (u / chunks) * RB;
      const int kbeg = (u % chunks) * a.col_chunk, kend = min(n_glob, kbeg + a.col_chunk);
      // thread owns output row (threadIdx.x / 8) and 16 features f = (threadIdx.x % 8) * 4 + 32 * j
      const int orow = threadIdx.x >> 3, of = (threadIdx.x & 7) * 4;
      float4 out[4] = {};
      __syncthreads();
      for (int t = threadIdx.x; t < RB * (D / 4); t += 256) {
        const int row = t / (D / 4), c4 = t % (D / 4);
        const int li = i0 + row;
        const float4 v = li < n_loc ? reinterpret_cast<const float4*>(
                                          zbuf + (size_t)global_row(li, a.B, a.world, a.rank) * D)[c4]
                                    : make_float4(0.f, 0.f, 0.f, 0.f);
        *reinterpret_cast<float4*>(sI + row * LDS + c4 * 4) = v;
      }
      for (int k0 = kbeg; k0 < kend; k0 += CB) {
        __syncthreads();
        for (int t = threadIdx.x; t < CB * (D / 4); t += 256) {
          const int row = t / (D / 4), c4 = t % (D / 4);
          const int g = k0 + row;
          const float4 v = g < kend ? reinterpret_cast<const float4*>(zbuf + (size_t)g * D)[c4]
                                    : make_float4(0.f, 0.f, 0.f, 0.f);
          *reinterpret_cast<float4*>(sK + row * LDS + c4 * 4) = v;
        }
        __syncthreads();
        float acc[2][4] = {};
#pragma unroll 4
        for (int d = 0; d < D; d += 4) {
          float4 zi[2], zk[4];
#pragma unroll
          for (int r = 0; r < 2; ++r) zi[r] = *reinterpret_cast<const float4*>(sI + (2 * tr + r) * LDS + d);
#pragma unroll
          for (int c = 0; c < 4; ++c) zk[c] = *reinterpret_cast<const float4*>(sK + (tc + 16 * c) * LDS + d);
#pragma unroll
          for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int c = 0; c < 4; ++c)
              acc[r][c] += zi[r].x * zk[c].x + zi[r].y * zk[c].y + zi[r].z * zk[c].z + zi[r].w * zk[c].w;
        }
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          const int li = i0 + 2 * tr + r;
          const int gi = li < n_loc ? global_row(li, a.B, a.world, a.rank) : -1;
          const float inv_ni = gi >= 0 ? 1.f / (float)a.neg[gi] : 0.f;
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const int gk = k0 + tc + 16 * c;
            float coef = 0.f;
            if (gi >= 0 && gk < kend && gk != gi) {
              coef = __expf(acc[r][c] * a.inv_t) * (inv_ni + 1.f / (float)a.neg[gk]);
              if (gk == (gi + half) % n_glob) coef -= 2.f;
              coef *= scale;
            }
            sC[(2 * tr + r) * (CB + 1) + tc + 16 * c] = coef;
          }
        }
        __syncthreads();
#pragma unroll 4
        for (int k = 0; k < CB; ++k) {
          const float c = sC[orow * (CB + 1) + k];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float4 zk = *reinterpret_cast<const float4*>(sK + k * LDS + of + 32 * j);
            out[j].x = fmaf(c, zk.x, out[j].x);
            out[j].y = fmaf(c, zk.y, out[j].y);
            out[j].z = fmaf(c, zk.z, out[j].z);
            out[j].w = fmaf(c, zk.w, out[j].w);
          }
        }
      }
      const int li = i0 + orow;
      if (li < n_loc) {  // this chunk's slab (plain stores; an empty chunk stores zeros)
        float* dst = a.gz + ((size_t)(u % chunks) * n_loc + li) * D + of;
#pragma unroll
        for (int j = 0; j < 4; ++j) *reinterpret_cast<float4*>(dst + 32 * j) = out[j];
      }
    }
    bar_target += gridDim.x;
    grid_barrier(a.barrier, bar_target);
  }

  // ---------------------------------------------------------------- P4: loss + chain rule to p
  // loss = (1/n) sum_i (log neg_i - S_i,pos(i)/T) over ALL global rows (every rank gets the global loss)
  {
    __shared__ float wpart[8];
    float part = 0.f;
    for (int g = blockIdx.x * blockDim.x + threadIdx.x; g < n_glob; g += gridDim.x * blockDim.x)
      part += logf((float)a.neg[g]) - a.spos[g];
    part = warp_sum(part);
    if (lane == 0) wpart[warp] = part;
    __syncthreads();
    if (threadIdx.x == 0) {
      float bp = 0.f;
      for (int w = 0; w < 8; ++w) bp += wpart[w];
      atomicAdd(a.acc, (double)bp);
      __threadfence();
      if (atomicAdd(a.barrier + 1, 1u) == gridDim.x - 1) {  // last block: every partial is in
        __threadfence();
        *a.loss = (float)(*reinterpret_cast<volatile double*>(a.acc) / (double)n_glob);
      }
    }
    if (blockIdx.x == 0 && threadIdx.x < 16 && a.stats)  // (P1's sums were complete at the first grid barrier)
      a.stats[threadIdx.x] = (float)(__ldcg(a.acc + 1 + threadIdx.x) / (double)a.B);
  }
  if (!a.g_p) return;
  const int gz_chunks = (n_glob + a.col_chunk - 1) / a.col_chunk;
  for (int i = gwarp; i < n_loc; i += nwarps) {
    const int gi = global_row(i, a.B, a.world, a.rank);
    float4 g = reinterpret_cast<const float4*>(a.gz + (size_t)i * D)[lane];
    for (int c = 1; c < gz_chunks; ++c) {  // slab order = fixed summation order
      const float4 v = reinterpret_cast<const float4*>(a.gz + ((size_t)c * n_loc + i) * D)[lane];
      g.x += v.x, g.y += v.y, g.z += v.z, g.w += v.w;
    }
    if (a.plain) {
      reinterpret_cast<float4*>(a.g_p + (size_t)i * D)[lane] = g;
      continue;
    }
    const float4 rb = reinterpret_cast<const float4*>(a.rowbuf)[i];
    const float inv_p = rb.x, inv_r = rb.y, al = rb.z, be = rb.w;
    const float4 zv = reinterpret_cast<const float4*>(zbuf + (size_t)gi * D)[lane];
    // through the second normalisation: g_r = (g_z - z (z . g_z)) / |r|
    const float dot = warp_sum(zv.x * g.x + zv.y * g.y + zv.z * g.z + zv.w * g.w);
    float gx0 = (g.x - zv.x * dot) * inv_r, gy0 = (g.y - zv.y * dot) * inv_r;
    float gx1 = (g.z - zv.z * dot) * inv_r, gy1 = (g.w - zv.w * dot) * inv_r;
    // through the rotation (transpose of the 2x2 block); translation and the detached centre pass nothing
    const float ux0 = al * gx0 - be * gy0, uy0 = be * gx0 + al * gy0;
    const float ux1 = al * gx1 - be * gy1, uy1 = be * gx1 + al * gy1;
    // through the first normalisation: u = p / |p|
    const float4 pv = reinterpret_cast<const float4*>(a.p + (size_t)i * D)[lane];
    const float u0 = pv.x * inv_p, u1 = pv.y * inv_p, u2 = pv.z * inv_p, u3 = pv.w * inv_p;
    const float dot2 = warp_sum(u0 * ux0 + u1 * uy0 + u2 * ux1 + u3 * uy1);
    float4 o;
    if (inv_p < 1e12f) {
      o = make_float4((ux0 - u0 * dot2) * inv_p, (uy0 - u1 * dot2) * inv_p, (ux1 - u2 * dot2) * inv_p,
                      (uy1 - u3 * dot2) * inv_p);
    } else {  // |p| clamped at eps: F.normalize divides by the constant eps
      o = make_float4(ux0 * inv_p, uy0 * inv_p, ux1 * inv_p, uy1 * inv_p);
    }
    reinterpret_cast<float4*>(a.g_p + (size_t)i * D)[lane] = o;
  }
}

}  // namespace peclr

using namespace peclr;

// launch geometry shared by the workspace query and the launch: grid, P3 column chunks
static void ntxent_plan(long long n_loc, long long n_glob, int* grid_out, int* chunks_out, int* col_chunk_out) {
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0)
      sms = 148;
  }
  const int p2_units = (int)(((n_glob + RB - 1) / RB) * ((n_glob + CB - 1) / CB));
  int grid = p2_units < sms ? p2_units : sms;
  if (grid < 1) grid = 1;
  // P3: split the columns so that there are about 2 units per CTA
  const int rbs = (int)((n_loc + RB - 1) / RB);
  int chunks = (2 * grid + rbs - 1) / rbs;
  const int max_chunks = (int)((n_glob + CB - 1) / CB);
  if (chunks > max_chunks) chunks = max_chunks;
  if (chunks < 1) chunks = 1;
  const int col_chunk = (int)(((n_glob + chunks - 1) / chunks + CB - 1) / CB) * CB;
  *grid_out = grid;
  *chunks_out = (int)((n_glob + col_chunk - 1) / col_chunk);
  *col_chunk_out = col_chunk;
}

// workspace layout (in floats): z [1 or 2][n_glob][128] | neg double[n_glob] | acc double[18] | rowbuf [n_loc][4] |
// spos [n_glob, padded to 4] | gz [chunks][n_loc][128] | barrier, done, ..., launch counter
static long long ntxent_layout(long long n_loc, long long n_glob, int world, int chunks, long long* off) {
  const long long zbufs = world > 1 ? 2 : 1;
  off[0] = zbufs * n_glob * D;                    // neg
  off[1] = off[0] + 2 * n_glob;                   // acc
  off[2] = off[1] + 36;                           // rowbuf
  off[3] = off[2] + n_loc * 4;                    // spos
  off[4] = off[3] + (n_glob + 3) / 4 * 4;         // gz
  off[5] = off[4] + (long long)chunks * n_loc * D;  // barrier block (16 words)
  return 4 * (off[5] + 64);
}

extern "C" long long peclr_ntxent_workspace_bytes(int B, int world) {
  const long long n_loc = 2LL * B, n_glob = n_loc * world;
  int grid, chunks, col_chunk;
  ntxent_plan(n_loc, n_glob, &grid, &chunks, &col_chunk);
  long long off[6];
  return ntxent_layout(n_loc, n_glob, world, chunks, off);
}

static int ntxent_launch(const float* p, const double* angle, const long long* jx, const long long* jy, int B,
                         int dim, int img_h, int img_w, int crop, int rotate, float temperature, float* loss,
                         float* stats, float* g_p, void* workspace, long long workspace_bytes, int world, int rank,
                         float* const* z_peers, unsigned* const* flag_peers, int plain, void* stream) {
  if (dim != D || B < 1 || world < 1 || rank < 0 || rank >= world) return -1001;
  if (workspace_bytes < peclr_ntxent_workspace_bytes(B, world)) return -1001;
  if ((crop && (!jx || !jy)) || (rotate && !angle)) return -1001;
  if (world > 1 && (!z_peers || !flag_peers || world > 32)) return -1001;
  const long long n_loc = 2LL * B, n_glob = n_loc * world;
  int grid, chunks, col_chunk;
  ntxent_plan(n_loc, n_glob, &grid, &chunks, &col_chunk);
  long long off[6];
  ntxent_layout(n_loc, n_glob, world, chunks, off);
  NtxentArgs a;
  float* ws = static_cast<float*>(workspace);
  a.p = p, a.angle = angle, a.jx = jx, a.jy = jy;
  a.z = ws;
  a.neg = reinterpret_cast<double*>(ws + off[0]);
  a.acc = reinterpret_cast<double*>(ws + off[1]);
  a.rowbuf = ws + off[2];
  a.spos = ws + off[3];
  a.gz = ws + off[4];
  a.barrier = reinterpret_cast<unsigned*>(ws + off[5]);
  a.launch_ctr = a.barrier + 8;  // lives in the (zero-initialised) workspace, advanced by the kernel
  a.z_peers = z_peers, a.flag_peers = flag_peers;
  a.loss = loss, a.stats = stats, a.g_p = g_p;
  a.B = B, a.world = world, a.rank = rank;
  a.img_h = img_h, a.img_w = img_w, a.crop = crop, a.rotate = rotate;
  a.inv_t = 1.f / temperature;
  a.plain = plain;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  a.col_chunk = col_chunk;
  cudaError_t e = cudaMemsetAsync(a.barrier, 0, 8, st);
  if (e != cudaSuccess) return -(int)e;
  const size_t smem = sizeof(float) * (RB * LDS + CB * LDS + RB * (CB + 1));
  e = cudaFuncSetAttribute(ntxent_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return -(int)e;
  void* args[] = {&a};
  e = cudaLaunchCooperativeKernel((const void*)ntxent_fused_kernel, dim3(grid), dim3(256), args, smem, st);
  return e == cudaSuccess ? 0 : -(int)e;
}

extern "C" int peclr_ntxent_fused(const float* p, const double* angle, const long long* jx, const long long* jy, int B,
                                  int dim, int img_h, int img_w, int crop, int rotate, float temperature, float* loss,
                                  float* stats, float* g_p, void* workspace, long long workspace_bytes, int world,
                                  int rank, float* const* z_peers, unsigned* const* flag_peers, void* stream) {
  if (!stats) return -1001;
  return ntxent_launch(p, angle, jx, jy, B, dim, img_h, img_w, crop, rotate, temperature, loss, stats, g_p,
                       workspace, workspace_bytes, world, rank, z_peers, flag_peers, 0, stream);
}

extern "C" int peclr_ntxent_plain(const float* z, int B, int dim, float temperature, float* loss, float* g_z,
                                  void* workspace, long long workspace_bytes, void* stream) {
  return ntxent_launch(z, nullptr, nullptr, nullptr, B, dim, 1, 1, 0, 0, temperature, loss, nullptr, g_z, workspace,
                       workspace_bytes, 1, 0, nullptr, nullptr, 1, stream);
}
