// Implicit-GEMM convolution on the 5th-gen tensor cores (tcgen05 + TMEM), operands staged by TMA.
//
// Data layout in HBM: activations NHWC bf16, viewed by TMA as rank-4 tensors (C, W, H, N); weights
// bf16 [Cout][taps][Cin] ("KRSC", K-major for fprop) and its transpose [Cin][taps][Cout] for dgrad;
// weight gradients fp32 [Cout][taps][Cin].
//
//   conv_gemm_kernel  (fprop / dgrad): D[pix, n] = sum_taps sum_c A[pix + tap, c] * B[n, tap, c]
//       A tile  = 128 output pixels (a Wb x Hb x Nb box of the pixel lattice) x 64 channels, one TMA box
//                 per filter tap; out-of-image taps are zero-filled by TMA (that is the padding).
//       B tile  = BN output channels x 64, K-major.   D = 128 x BN fp32 accumulator in TMEM.
//       Strided convolutions read/write parity sub-lattices of the image (plain tiled TMA views with
//       doubled strides), so one kernel covers 1x1, 3x3, stride 1 and 2, fprop and dgrad, and the
//       7x7 stem (a 4x4 / stride 1 convolution on the space-to-depth image, whose "channels" are 4-block x
//       16-channel windows of it).
//       Epilogue: TMEM -> registers -> bf16 -> swizzled smem -> TMA store (or TMA reduce-add for
//       gradient accumulation), plus per-channel sum / sum-of-squares of the stored values for the
//       following training-mode BatchNorm.
//
//   conv_wgrad_kernel : dW[n, tap, c] += sum_pix dY[pix, n] * X[pix + tap, c]
//       both operands are loaded exactly as above ([pixels][64 channels] boxes) and consumed as
//       MN-major UMMA operands (contraction over pixels); M = 128 input channels, N = BN output
//       channels, up to 512/BN taps accumulate side by side in TMEM; split over pixel ranges across
//       CTAs: every split writes its fp32 partial to a workspace slab and wgrad_reduce_kernel adds the slabs
//       to dW in a fixed order (no atomics: the weight gradients are bit-reproducible run to run).
//
// Warp roles: warp 0 = TMA producer, warp 1 = MMA issuer (+ TMEM alloc), then the epilogue warps (8 in the
// fprop/dgrad kernel: two per TMEM lane quarter, splitting the columns; 4 in the wgrad kernel).
#include "conv_tc.h"

#include <stdio.h>

#include "ptx.cuh"

namespace peclr {

constexpr int kStageA = 128 * 128;  // 128 pixel rows x 64 bf16

struct GemmParams {
  CUtensorMap a_maps[kMaxViews];
  CUtensorMap b_map;
  CUtensorMap d_map;
  TapTable taps;
  int num_taps, c_chunks;
  int tiles_w, tiles_h, tiles_n, n_tiles;
  FastDiv fd_tiles_w, fd_tiles_h, fd_n_tiles;
  int Wb, Hb, Nb;
  int log_wb, log_wbhb;
  int d_w, d_h, d_n;
  int cout;
  // per-channel statistics of the stored tile values (forward: BatchNorm sums; fused BN-backward reduction: sum g,
  // sum g*y).  Every CTA accumulates its tiles in a fixed order in shared memory and adds ONE fp64 partial per
  // channel to these accumulators (see red_add_f64): the totals do not depend on the arrival order of the CTAs.
  double* stat_sum;
  double* stat_sumsq;
  int reduce_add;
  // fused BatchNorm-backward reduction (dgrad feeding an inner BN + ReLU): with y = that BN's input tile,
  // g = D * [fma(y, gamma*invstd, beta - mean*gamma*invstd) > 0]; stat_sum += sum g, stat_sumsq += sum g*y
  CUtensorMap y_map;
  const float *bn_mean, *bn_invstd, *bn_gamma, *bn_beta;
  int bn_reduce;  // 0: none; 1: as above; 2: "finish" mode, below
  // finish mode (the dgrad that COMPLETES the gradient of a residual block's input = the previous block's output):
  // the tile already in the output tensor (shortcut gradient, or the other branch's dgrad) is loaded through d_map
  // and added in registers, the sum is masked with that block output's ReLU bits (1 byte per 8 channels, written by
  // bn_apply) and stored; stat_sum += sum g, stat_sumsq += sum g*y with y = the previous block's last conv output.
  // One pass instead of TMA reduce-add + a BN-backward reduction pass + a masking pass.
  const uint8_t* mask_bits;
  long long pix_base, pix_w, pix_h, pix_n;  // pixel index of view element (w, h, n) = base + w*pix_w + h*pix_h + n*pix_n
  int mask_row_bytes;
  int acc_stride;      // 2: the tile already in the tensor is valid on the even-row / even-column pixel lattice only
  FastDiv fd_img_w;    // (image width, for the row parity of a flat pixel index)
  // filter-row halo mode (HALO kernels): a tap table entry is a GROUP of halo_taps taps that differ only in dh
  // (dh = taps.dh[g] + t).  One (Hb + halo_taps - 1) x Wb input box per group and channel chunk serves all of them:
  // tap t reads the same shared-memory buffer from row t * Wb on (Wb is a multiple of 8, so that offset is a
  // multiple of the 1024-byte swizzle atom).  Tap t's weights sit halo_kstep columns after tap t - 1's.
  int halo_taps, halo_kstep;
  int halo_a_bytes, halo_stage_a, halo_stages;
  int n_stages;  // pipeline stages in use (<= the kernel's STAGES barriers)
  int d_bufs;    // output staging tiles: 2 = tile i + 1 is drained while tile i's TMA store still reads its buffer
  int stat_copies;  // 512 / BN private copies of the per-CTA statistics (one owner thread per address), or 1 (the
                    // row groups then take turns on the single copy, in a fixed order)
};

// CTAS == 2: the two CTAs of a cluster form one 256 x BN tile (tcgen05 cta_group::2).  Each CTA stages its own
// 128 pixel rows of A and half of the BN weight rows, so the L2 -> shared-memory traffic per FLOP drops by a
// third; the leader CTA issues the MMAs, each CTA drains its own 128 TMEM lanes.
template <int BN, int STAGES, int CTAS, bool HALO = false>
__global__ void __launch_bounds__(320, 1) conv_gemm_kernel(const __grid_constant__ GemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int kStageB = (BN / CTAS) * 128;
  const uint32_t cta_rank = CTAS == 2 ? cluster_ctarank() : 0u;
  // pipeline geometry: fixed for the per-tap kernels; in halo mode the A stage holds the taller box, the B stage
  // one weight tile per tap of the group, and the number of stages is whatever fits (<= STAGES barriers)
  const int n_stages = p.n_stages;
  const uint32_t stage_a = HALO ? (uint32_t)p.halo_stage_a : (uint32_t)kStageA;
  const uint32_t stage_b = HALO ? (uint32_t)(p.halo_taps * kStageB) : (uint32_t)kStageB;
  uint8_t* sA = smem;
  uint8_t* sB = sA + n_stages * stage_a;
  uint8_t* sD = sB + n_stages * stage_b;
  constexpr uint32_t kTileD = (BN / 64) * kStageA;
  uint8_t* sY = sD + p.d_bufs * kTileD;  // y tile of the fused BN-backward reduction (only if p.bn_reduce)
  // finish mode: the tile already in the output tensor, DOUBLE buffered -- it is needed first (by the TMEM drain), and
  // a single buffer exposed its load latency (~3000 clk of a ~6800 clk tile: the load could only be issued after the
  // previous tile's drain)
  uint8_t* sG = sY + (p.bn_reduce ? (BN / 64) * kStageA : 0);
  uint64_t* full = reinterpret_cast<uint64_t*>(sG + (p.bn_reduce == 2 ? 2 * (BN / 64) * kStageA : 0));
  uint64_t* empty = full + STAGES;
  uint64_t* tfull = empty + STAGES;
  uint64_t* tempty = tfull + 2;
  uint64_t* ybar = tempty + 2;  // [0] y tile landed, [1] y tile consumed
  uint64_t* gfull = ybar + 2;   // [b] accumulated-gradient tile landed in buffer b (finish mode)
  uint64_t* gempty = gfull + 2; // [b] buffer b consumed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gempty + 2);
  // per-CTA statistics, one private copy per row group ([512/BN][2][cout]): every address has exactly one owner
  // thread, so the per-tile accumulation is a plain read-modify-write in a fixed order (reproducible; shared fp32
  // atomics are CAS loops and would add in arrival order).  When the copies do not fit next to the pipeline stages
  // (stat_copies == 1) the row groups take turns on one copy, separated by named barriers.
  float* sStat = reinterpret_cast<float*>(tmem_slot + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const bool want_stats = p.stat_sum != nullptr;

  if (warp == 0 && elect_one()) {
    for (int i = 0; i < kMaxViews; ++i) prefetch_tmap(&p.a_maps[i]);
    prefetch_tmap(&p.b_map);
    prefetch_tmap(&p.d_map);
  }
  if (warp == 1) {
    if (elect_one()) {
      for (int i = 0; i < STAGES; ++i) {
        mbar_init(&full[i], 1);
        mbar_init(&empty[i], 1);
      }
      for (int i = 0; i < 2; ++i) {
        mbar_init(&tfull[i], 1);
        mbar_init(&tempty[i], 8 * CTAS);  // one arrival per epilogue warp (of both CTAs: the leader's barrier)
        mbar_init(&ybar[i], 1);
        mbar_init(&gfull[i], 1);
        mbar_init(&gempty[i], 1);
      }
      fence_barrier_init();
    }
    __syncwarp();
    if constexpr (CTAS == 2) tmem_alloc_2sm<2 * BN>(tmem_slot);
    else tmem_alloc<2 * BN>(tmem_slot);
  }
  if (want_stats && warp >= 2) {
    for (int i = threadIdx.x - 64; i < p.stat_copies * 2 * p.cout; i += 256) sStat[i] = 0.f;
  }
  pdl_launch_dependents();
  tc_fence_before();
  __syncthreads();
  if constexpr (CTAS == 2) cluster_sync_all();  // the peer's barriers exist before anything signals them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();  // everything above overlapped the previous kernel's tail; global memory is touched from here on

  // work items: (pair of) m tiles x n tile; CTA `cta_rank` of a pair owns m tile 2 * pair + cta_rank
  const int m_tiles = p.tiles_w * p.tiles_h * p.tiles_n;
  const int total_tiles = ((m_tiles + CTAS - 1) / CTAS) * p.n_tiles;
  const int first_tile = blockIdx.x / CTAS, tile_step = gridDim.x / CTAS;
  const int num_kb = p.num_taps * p.c_chunks;

  if (warp == 0) {
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      int tile_no = 0;
      for (int t = first_tile; t < total_tiles; t += tile_step, ++tile_no) {
        const int tq = fd_div(p.fd_n_tiles, t);
        const int nt = t - tq * p.n_tiles;
        const int mt = tq * CTAS + cta_rank;  // may be one past the end: all-OOB boxes (zero fill)
        const int mq = fd_div(p.fd_tiles_w, mt), nq = fd_div(p.fd_tiles_h, mq);
        const int w0 = (mt - mq * p.tiles_w) * p.Wb;
        const int h0 = (mq - nq * p.tiles_h) * p.Hb;
        const int n0 = nq * p.Nb;
        for (int tap = 0; tap < p.num_taps; ++tap) {
          const CUtensorMap* amap = &p.a_maps[p.taps.view[tap]];
          const int cw = w0 + p.taps.dw[tap];
          const int ch = h0 + p.taps.dh[tap];
          const int kbase = p.taps.koff[tap];
          for (int c = 0; c < p.c_chunks; ++c) {
            mbar_wait(&empty[stage], phase ^ 1);
            if constexpr (HALO) {
              // one tall box for the whole tap group + one weight tile per tap
              const uint32_t bytes = (uint32_t)p.halo_a_bytes + (uint32_t)p.halo_taps * kStageB;
              uint8_t* dstA = sA + stage * stage_a;
              uint8_t* dstB = sB + stage * stage_b;
              if constexpr (CTAS == 2) {
                if (cta_rank == 0) mbar_expect_tx(&full[stage], 2 * bytes);
                tma_load_4d_2sm(amap, &full[stage], dstA, c * 64, cw, ch, n0);
                for (int ht = 0; ht < p.halo_taps; ++ht)
                  tma_load_2d_2sm(&p.b_map, &full[stage], dstB + ht * kStageB, kbase + ht * p.halo_kstep + c * 64,
                                  nt * BN + cta_rank * (BN / 2));
              } else {
                mbar_expect_tx(&full[stage], bytes);
                tma_load_4d(amap, &full[stage], dstA, c * 64, cw, ch, n0);
                for (int ht = 0; ht < p.halo_taps; ++ht)
                  tma_load_2d(&p.b_map, &full[stage], dstB + ht * kStageB, kbase + ht * p.halo_kstep + c * 64, nt * BN);
              }
            } else if constexpr (CTAS == 2) {
              // both CTAs' bytes are counted on the leader's barrier
              if (cta_rank == 0) mbar_expect_tx(&full[stage], 2 * (kStageA + kStageB));
              tma_load_4d_2sm(amap, &full[stage], sA + stage * kStageA, c * 64, cw, ch, n0);
              tma_load_2d_2sm(&p.b_map, &full[stage], sB + stage * kStageB, kbase + c * 64,
                              nt * BN + cta_rank * (BN / 2));
            } else {
              mbar_expect_tx(&full[stage], kStageA + kStageB);
              tma_load_4d(amap, &full[stage], sA + stage * kStageA, c * 64, cw, ch, n0);
              tma_load_2d(&p.b_map, &full[stage], sB + stage * kStageB, kbase + c * 64, nt * BN);
            }
            if (++stage == n_stages) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
        if (p.bn_reduce == 2) {  // the gradient tile to add to (needed first, by the TMEM drain)
          const int gb = tile_no & 1;
          mbar_wait(&gempty[gb], ((tile_no >> 1) & 1) ^ 1);
          mbar_expect_tx(&gfull[gb], (BN / 64) * kStageA);
#pragma unroll
          for (int bx = 0; bx < BN / 64; ++bx)
            tma_load_4d(&p.d_map, &gfull[gb], sG + (gb * (BN / 64) + bx) * kStageA, nt * BN + bx * 64, w0, h0, n0);
        }
        if (p.bn_reduce) {  // this tile's BN input, for the epilogue (own CTA, own barrier)
          mbar_wait(&ybar[1], (tile_no & 1) ^ 1);
          mbar_expect_tx(&ybar[0], (BN / 64) * kStageA);
#pragma unroll
          for (int bx = 0; bx < BN / 64; ++bx)
            tma_load_4d(&p.y_map, &ybar[0], sY + bx * kStageA, nt * BN + bx * 64, w0, h0, n0);
        }
      }
    }
  } else if (warp == 1) {
    if (cta_rank == 0 && elect_one()) {
      constexpr uint32_t idesc = make_idesc_bf16(128 * CTAS, BN, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int t = first_tile; t < total_tiles; t += tile_step, ++it) {
        const int acc = it & 1;
        const uint32_t aphase = (it >> 1) & 1;
        mbar_wait(&tempty[acc], aphase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          if constexpr (HALO) {
            const uint32_t a0 = smem_u32(sA + stage * stage_a), b0 = smem_u32(sB + stage * stage_b);
            const uint32_t row_bytes = (uint32_t)p.Wb * 128u;  // one image row of the box (multiple of 1024 B)
            for (int ht = 0; ht < p.halo_taps; ++ht) {
              const uint64_t a_desc = make_smem_desc(a0 + ht * row_bytes, 0, 1024);
              const uint64_t b_desc = make_smem_desc(b0 + ht * kStageB, 0, 1024);
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                if constexpr (CTAS == 2)
                  umma_bf16_2sm(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | ht | k) != 0);
                else
                  umma_bf16(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | ht | k) != 0);
              }
            }
          } else {
            const uint64_t a_desc = make_smem_desc(smem_u32(sA + stage * kStageA), 0, 1024);
            const uint64_t b_desc = make_smem_desc(smem_u32(sB + stage * kStageB), 0, 1024);
#pragma unroll
            for (int k = 0; k < 4; ++k) {  // 4 x UMMA_K(16) = 64 channels; +32 B per step inside the swizzle atom
              if constexpr (CTAS == 2) umma_bf16_2sm(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0);
              else umma_bf16(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0);
            }
          }
          if constexpr (CTAS == 2) umma_commit_2sm(&empty[stage]);
          else umma_commit(&empty[stage]);
          if (++stage == n_stages) {
            stage = 0;
            phase ^= 1;
          }
        }
        if constexpr (CTAS == 2) umma_commit_2sm(&tfull[acc]);
        else umma_commit(&tfull[acc]);
      }
    }
  } else {
    // 8 epilogue warps: warps w and w + 4 share TMEM lane quarter (w & 3) and split the tile's columns
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    const int et = threadIdx.x - 64;  // 0..255
    // finish mode: this thread's ReLU-mask words (one per column chunk) of tile t.  They are fetched ONE TILE AHEAD:
    // issued at the point of use, this global load was the largest single stall of the kernel (14 % of all samples),
    // and still 9 % when issued at the top of its own tile.
    const bool finish = p.bn_reduce == 2;
    auto fetch_mask = [&](int t, uint32_t* words, bool* acc_valid) {
      const int tq = fd_div(p.fd_n_tiles, t);
      const int nt = t - tq * p.n_tiles;
      const int mt = tq * CTAS + cta_rank;
      const int mq = fd_div(p.fd_tiles_w, mt), nq = fd_div(p.fd_tiles_h, mq);
      const int w = (mt - mq * p.tiles_w) * p.Wb + (row & (p.Wb - 1));
      const int h = (mq - nq * p.tiles_h) * p.Hb + ((row >> p.log_wb) & (p.Hb - 1));
      const int n = nq * p.Nb + (row >> p.log_wbhb);
      // off-image rows: the TMA store clips them, the statistics must not see them
      const bool on_image = t < total_tiles && w < p.d_w && h < p.d_h && n < p.d_n;
      const long long pidx = p.pix_base + (long long)w * p.pix_w + (long long)h * p.pix_h + (long long)n * p.pix_n;
      const long long mask_row = pidx * p.mask_row_bytes;
      // (image width even: the column parity of a flat pixel index is its own parity)
      *acc_valid = p.acc_stride == 1 || (((int)pidx | fd_div(p.fd_img_w, (int)pidx)) & 1) == 0;
#pragma unroll
      for (int ci = 0; ci < BN / 64; ++ci)
        words[ci] = on_image ? *reinterpret_cast<const uint32_t*>(p.mask_bits + mask_row +
                                                                  ((nt * BN + (half * (BN / 64) + ci) * 32) >> 3))
                             : 0u;
    };
    uint32_t mask_next[BN / 64];
    bool acc_valid_next = true;
    if (finish) fetch_mask(first_tile, mask_next, &acc_valid_next);
    int it = 0;
    for (int t = first_tile; t < total_tiles; t += tile_step, ++it) {
      const int acc = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      const int tq = fd_div(p.fd_n_tiles, t);
      const int nt = t - tq * p.n_tiles;
      const int mt = tq * CTAS + cta_rank;
      const int mq = fd_div(p.fd_tiles_w, mt), nq = fd_div(p.fd_tiles_h, mq);
      const int w0 = (mt - mq * p.tiles_w) * p.Wb;
      const int h0 = (mq - nq * p.tiles_h) * p.Hb;
      const int n0 = nq * p.Nb;
      // output staging buffer of this tile; with two buffers the drain below only waits for the store of tile it - 2
      uint8_t* sDt = sD + (p.d_bufs == 2 ? (it & 1) * kTileD : 0u);
      const uint32_t sD32 = smem_u32(sDt);
      uint32_t mask_word[BN / 64];
      bool acc_valid = true;
      if (finish) {
#pragma unroll
        for (int ci = 0; ci < BN / 64; ++ci) mask_word[ci] = mask_next[ci];
        acc_valid = acc_valid_next;
        fetch_mask(t + tile_step, mask_next, &acc_valid_next);
      }
      mbar_wait(&tfull[acc], aphase);
      tc_fence_after();
      if (finish) mbar_wait(&gfull[it & 1], (it >> 1) & 1);
      if (et == 0) {  // the TMA store that last used this buffer has finished reading it
        if (p.d_bufs == 2) tma_wait_group_read1();
        else tma_wait_group_read0();
      }
      named_bar_sync(1, 256);
#pragma unroll
      for (int ci = 0; ci < BN / 64; ++ci) {
        const int chunk = half * (BN / 64) + ci;
        uint32_t r[32];
        tmem_ld32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BN + chunk * 32, r);
        tmem_ld_wait();
        const uint32_t box = sD32 + (chunk >> 1) * kStageA + row * 128;
        if (finish) {
          // r += the tile already in the tensor (same swizzled position in sG as the store position in sD), then the
          // ReLU mask of the block output this is the gradient of: bit k of the word = channel (chunk * 32 + k)
          const uint32_t gdelta = smem_u32(sG) + (it & 1) * (BN / 64) * kStageA - sD32;
          const uint32_t bits = mask_word[ci];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int c16 = (chunk & 1) * 4 + j;
            uint4 g = lds_v4(box + gdelta + ((c16 ^ (row & 7)) << 4));
            if (!acc_valid) g = make_uint4(0u, 0u, 0u, 0u);  // never-written memory: selected away, not multiplied
            const uint32_t gw[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int k = 8 * j + 2 * e;
              const float v0 = __uint_as_float(r[k]) + bf16_lo(gw[e]), v1 = __uint_as_float(r[k + 1]) + bf16_hi(gw[e]);
              r[k] = (bits >> k) & 1u ? __float_as_uint(v0) : 0u;
              r[k + 1] = (bits >> (k + 1)) & 1u ? __float_as_uint(v1) : 0u;
            }
          }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int c16 = (chunk & 1) * 4 + j;  // 16-byte chunk inside the 128-byte row
          uint4 v;
          v.x = pack_bf16x2(__uint_as_float(r[8 * j + 0]), __uint_as_float(r[8 * j + 1]));
          v.y = pack_bf16x2(__uint_as_float(r[8 * j + 2]), __uint_as_float(r[8 * j + 3]));
          v.z = pack_bf16x2(__uint_as_float(r[8 * j + 4]), __uint_as_float(r[8 * j + 5]));
          v.w = pack_bf16x2(__uint_as_float(r[8 * j + 6]), __uint_as_float(r[8 * j + 7]));
          sts_v4(box + ((c16 ^ (row & 7)) << 4), v);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {  // accumulator drained: the (leader's) MMA warp may reuse it
        if constexpr (CTAS == 2) mbar_arrive_cluster(&tempty[acc], 0);
        else mbar_arrive(&tempty[acc]);
      }
      fence_proxy_async_smem();
      named_bar_sync(1, 256);
      if (et == 0) {
        if (finish) mbar_arrive(&gempty[it & 1]);  // every thread has read its part of this sG buffer
#pragma unroll
        for (int b = 0; b < BN / 64; ++b) {
          if (p.reduce_add)
            tma_reduce_add_4d(&p.d_map, sDt + b * kStageA, nt * BN + b * 64, w0, h0, n0);
          else
            tma_store_4d(&p.d_map, sDt + b * kStageA, nt * BN + b * 64, w0, h0, n0);
        }
        tma_commit_group();
      }
      if (want_stats) {
        // column sums of the bf16 values just staged (exactly what the next kernels will read).  A thread owns four
        // adjacent columns (one 8-byte shared load per row) and a contiguous row range; the two half-warps of a warp
        // read two different rows (a full 128-byte line each: conflict free) and meet through one shuffle.  The
        // arithmetic runs on packed fp32 pairs: this loop is half of the epilogue's instructions, and the epilogue is
        // the critical path of the wide 1x1 convolutions.
        constexpr int kQuadWarps = BN / 64;            // warps that together span the BN columns
        constexpr int kRowGroups = 2 * (8 / kQuadWarps);  // = 1024 / BN
        constexpr int kRowsPer = 128 / kRowGroups;        // = BN / 8
        const int ew = et >> 5;                            // epilogue warp 0..7
        const int copy = ew / kQuadWarps;                  // statistics copy (of 512 / BN) this warp adds to
        const int rg = copy * 2 + (lane >> 4);
        const int col = ((ew % kQuadWarps) * 16 + (lane & 15)) * 4;
        const int c16 = (col & 63) >> 3;
        const uint32_t base = sD32 + (col >> 6) * kStageA + ((col & 4) << 1) + rg * kRowsPer * 128;
        const uint32_t ydelta = smem_u32(sY) - sD32;
        float2 s01[2] = {{0.f, 0.f}, {0.f, 0.f}}, s23[2] = {{0.f, 0.f}, {0.f, 0.f}};
        float2 q01[2] = {{0.f, 0.f}, {0.f, 0.f}}, q23[2] = {{0.f, 0.f}, {0.f, 0.f}};
        float2 sc01 = {0.f, 0.f}, sc23 = {0.f, 0.f}, sh01 = {0.f, 0.f}, sh23 = {0.f, 0.f};
        if (p.bn_reduce == 2) mbar_wait(&ybar[0], it & 1);
        if (p.bn_reduce == 1) {
          const int ch = nt * BN + col;
          // (scalar loads: gamma / beta are views into the flat parameter buffer, not necessarily 16-byte aligned)
          sc01 = make_float2(p.bn_gamma[ch] * p.bn_invstd[ch], p.bn_gamma[ch + 1] * p.bn_invstd[ch + 1]);
          sc23 = make_float2(p.bn_gamma[ch + 2] * p.bn_invstd[ch + 2], p.bn_gamma[ch + 3] * p.bn_invstd[ch + 3]);
          sh01 = make_float2(p.bn_beta[ch] - p.bn_mean[ch] * sc01.x, p.bn_beta[ch + 1] - p.bn_mean[ch + 1] * sc01.y);
          sh23 = make_float2(p.bn_beta[ch + 2] - p.bn_mean[ch + 2] * sc23.x,
                             p.bn_beta[ch + 3] - p.bn_mean[ch + 3] * sc23.y);
          mbar_wait(&ybar[0], it & 1);
        }
        // rows of a box that hangs over the image edge are clipped by the TMA store: keep them out of the
        // statistics too (their taps can still reach valid pixels, so they are not zero)
        const bool edge = w0 + p.Wb > p.d_w || h0 + p.Hb > p.d_h || n0 + p.Nb > p.d_n;
#pragma unroll 2
        for (int r8 = 0; r8 < kRowsPer / 8; ++r8) {
          uint2 v[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] = lds_v2(base + r8 * 1024 + j * 128 + ((c16 ^ j) << 4));
          if (edge) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const int r0 = rg * kRowsPer + r8 * 8 + j;
              const int wl = r0 & (p.Wb - 1), hl = (r0 >> p.log_wb) & (p.Hb - 1), nl = r0 >> p.log_wbhb;
              if (w0 + wl >= p.d_w || h0 + hl >= p.d_h || n0 + nl >= p.d_n) v[j] = make_uint2(0u, 0u);
            }
          }
          if (p.bn_reduce) {
            uint2 u[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) u[j] = lds_v2(base + ydelta + r8 * 1024 + j * 128 + ((c16 ^ j) << 4));
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float2 y01 = bf16x2_f2(u[j].x), y23 = bf16x2_f2(u[j].y);
              float2 g01 = bf16x2_f2(v[j].x), g23 = bf16x2_f2(v[j].y);
              if (p.bn_reduce == 1) {  // the gradient passes where the recomputed BN + ReLU output is positive
                const float2 t01 = ffma2(y01, sc01, sh01), t23 = ffma2(y23, sc23, sh23);
                g01.x = t01.x > 0.f ? g01.x : 0.f, g01.y = t01.y > 0.f ? g01.y : 0.f;
                g23.x = t23.x > 0.f ? g23.x : 0.f, g23.y = t23.y > 0.f ? g23.y : 0.f;
              }  // (mode 2: the staged values are the masked gradient already)
              s01[j & 1] = fadd2(s01[j & 1], g01), s23[j & 1] = fadd2(s23[j & 1], g23);
              q01[j & 1] = ffma2(g01, y01, q01[j & 1]), q23[j & 1] = ffma2(g23, y23, q23[j & 1]);
            }
            continue;
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float2 x01 = bf16x2_f2(v[j].x), x23 = bf16x2_f2(v[j].y);
            s01[j & 1] = fadd2(s01[j & 1], x01), s23[j & 1] = fadd2(s23[j & 1], x23);
            q01[j & 1] = ffma2(x01, x01, q01[j & 1]), q23[j & 1] = ffma2(x23, x23, q23[j & 1]);
          }
        }
        // even + odd rows, then lower + upper half-warp (a fixed order), then one owner thread per address
        float4 ts, tq;
        ts.x = s01[0].x + s01[1].x, ts.y = s01[0].y + s01[1].y, ts.z = s23[0].x + s23[1].x, ts.w = s23[0].y + s23[1].y;
        tq.x = q01[0].x + q01[1].x, tq.y = q01[0].y + q01[1].y, tq.z = q23[0].x + q23[1].x, tq.w = q23[0].y + q23[1].y;
        ts.x += __shfl_down_sync(0xffffffffu, ts.x, 16), ts.y += __shfl_down_sync(0xffffffffu, ts.y, 16);
        ts.z += __shfl_down_sync(0xffffffffu, ts.z, 16), ts.w += __shfl_down_sync(0xffffffffu, ts.w, 16);
        tq.x += __shfl_down_sync(0xffffffffu, tq.x, 16), tq.y += __shfl_down_sync(0xffffffffu, tq.y, 16);
        tq.z += __shfl_down_sync(0xffffffffu, tq.z, 16), tq.w += __shfl_down_sync(0xffffffffu, tq.w, 16);
        if (p.stat_copies > 1) {
          if (lane < 16) {
            float4* mine_s = reinterpret_cast<float4*>(sStat + (size_t)copy * 2 * p.cout + nt * BN + col);
            float4* mine_q = reinterpret_cast<float4*>(sStat + (size_t)copy * 2 * p.cout + p.cout + nt * BN + col);
            float4 as = *mine_s, aq = *mine_q;
            as.x += ts.x, as.y += ts.y, as.z += ts.z, as.w += ts.w;
            aq.x += tq.x, aq.y += tq.y, aq.z += tq.z, aq.w += tq.w;
            *mine_s = as, *mine_q = aq;
          }
        } else {
          for (int g = 0; g < 512 / BN; ++g) {  // fixed order: copy 0's rows first
            if (copy == g && lane < 16) {
              float4* mine_s = reinterpret_cast<float4*>(sStat + nt * BN + col);
              float4* mine_q = reinterpret_cast<float4*>(sStat + p.cout + nt * BN + col);
              float4 as = *mine_s, aq = *mine_q;
              as.x += ts.x, as.y += ts.y, as.z += ts.z, as.w += ts.w;
              aq.x += tq.x, aq.y += tq.y, aq.z += tq.z, aq.w += tq.w;
              *mine_s = as, *mine_q = aq;
            }
            named_bar_sync(3, 256);
          }
        }
        if (p.bn_reduce) {  // every thread is done with the y tile: hand the buffer back to the producer
          named_bar_sync(2, 256);
          if (et == 0) mbar_arrive(&ybar[1]);
        }
      }
    }
    if (et == 0) tma_wait_group0();
    if (want_stats) {
      named_bar_sync(1, 256);
      for (int i = et; i < p.cout; i += 256) {
        float s = 0.f, qq = 0.f;
        for (int g = 0; g < p.stat_copies; ++g) s += sStat[g * 2 * p.cout + i], qq += sStat[g * 2 * p.cout + p.cout + i];
        if (s != 0.f || qq != 0.f) {
          red_add_f64(p.stat_sum + i, (double)s);
          red_add_f64(p.stat_sumsq + i, (double)qq);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (CTAS == 2) cluster_sync_all();  // the peer may still be signalling this CTA's barriers / reading smem
  if (warp == 1) {
    tc_fence_after();
    if constexpr (CTAS == 2) tmem_dealloc_2sm<2 * BN>(tmem_base);
    else tmem_dealloc<2 * BN>(tmem_base);
  }
}

// ------------------------------------------------------------------------------------------ wgrad
struct WgradParams {
  CUtensorMap x_maps[kMaxViews];
  CUtensorMap dy_map;
  TapTable taps;
  int num_taps, taps_per_unit, tap_groups;
  int pair_taps;     // Cin == 64: the two 64-channel halves of the 128-row operand hold two different filter taps
  int group_over_m;  // 1x1 filters: the G accumulators of a unit are G consecutive 128-channel m tiles (tap 0)
  int m_tiles, n_tiles, ksplit;
  int tiles_w, tiles_h, tiles_n;
  int Wb, Hb, Nb;
  int cin, a_boxes;
  int64_t ld_co;  // elements between consecutive output channels in dW ( = taps * cin )
  float* dw;
  float* partial;          // ksplit > 1: [ksplit][cout * taps * cin] slabs in dW's layout (plain stores)
  int64_t partial_stride;  // elements per slab
};

template <int BN>
__global__ void __maxnreg__(80) conv_wgrad_kernel(const __grid_constant__ WgradParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int kBoxBytes = 64 * 128;  // 64 pixels x 64 channels bf16
  constexpr int kBBytes = (BN / 64) * kBoxBytes;
  constexpr int kABytes = 2 * kBoxBytes;  // 128 input channels per tap
  const int G = p.taps_per_unit;
  const int stage_bytes = kBBytes + G * kABytes;
  const int stages = (200 * 1024) / stage_bytes;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + stages * stage_bytes);
  uint64_t* empty = full + 32;
  uint64_t* tfull = empty + 32;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && elect_one()) {
    for (int i = 0; i < kMaxViews; ++i) prefetch_tmap(&p.x_maps[i]);
    prefetch_tmap(&p.dy_map);
  }
  if (warp == 1) {
    if (elect_one()) {
      for (int i = 0; i < stages; ++i) {
        mbar_init(&full[i], 1);
        mbar_init(&empty[i], 1);
      }
      mbar_init(tfull, 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc<512>(tmem_slot);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // unit -> (m tile, n tile, tap group, pixel range)
  int u = blockIdx.x;
  const int mt = u % p.m_tiles;
  u /= p.m_tiles;
  const int nt = u % p.n_tiles;
  u /= p.n_tiles;
  const int tg = u % p.tap_groups;
  const int ks = u / p.tap_groups;
  const int tap0 = p.group_over_m ? 0 : tg * G;
  const int mt0 = p.group_over_m ? mt * G : mt;  // first 128-channel tile of this unit
  const int slots = p.pair_taps ? (p.num_taps + 1) / 2 : p.num_taps;  // accumulators needed for all taps
  const int ntap = p.group_over_m ? min(G, (p.cin + 127) / 128 - mt0) : min(G, slots - tap0);
  const int total_chunks = p.tiles_w * p.tiles_h * p.tiles_n;
  const int per = (total_chunks + p.ksplit - 1) / p.ksplit;
  const int c_begin = ks * per;
  const int c_end = min(total_chunks, c_begin + per);

  if (warp == 0) {
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      // pixel-chunk coordinates advance incrementally (one division set per CTA, none per pipeline stage)
      int tw = c_begin % p.tiles_w, th = (c_begin / p.tiles_w) % p.tiles_h, tn = c_begin / (p.tiles_w * p.tiles_h);
      for (int c = c_begin; c < c_end; ++c) {
        const int w0 = tw * p.Wb, h0 = th * p.Hb, n0 = tn * p.Nb;
        if (++tw == p.tiles_w) {
          tw = 0;
          if (++th == p.tiles_h) th = 0, ++tn;
        }
        uint8_t* st = smem + stage * stage_bytes;
        mbar_wait(&empty[stage], phase ^ 1);
        int a_box_count = ntap * p.a_boxes;
        if (p.pair_taps && 2 * (tap0 + ntap) > p.num_taps) --a_box_count;  // odd tap count: last half stays empty
        mbar_expect_tx(&full[stage], kBBytes + a_box_count * kBoxBytes);
#pragma unroll
        for (int b = 0; b < BN / 64; ++b)
          tma_load_4d(&p.dy_map, &full[stage], st + b * kBoxBytes, nt * BN + b * 64, w0, h0, n0);
        for (int g = 0; g < ntap; ++g) {
          if (p.pair_taps) {
            for (int b = 0; b < 2; ++b) {
              const int tap = 2 * (tap0 + g) + b;
              if (tap >= p.num_taps) break;
              tma_load_4d(&p.x_maps[p.taps.view[tap]], &full[stage], st + kBBytes + g * kABytes + b * kBoxBytes, 0,
                          w0 + p.taps.dw[tap], h0 + p.taps.dh[tap], n0);
            }
            continue;
          }
          const int tap = p.group_over_m ? 0 : tap0 + g;
          const int mtile = p.group_over_m ? mt0 + g : mt0;
          const CUtensorMap* xm = &p.x_maps[p.taps.view[tap]];
          for (int b = 0; b < p.a_boxes; ++b)
            tma_load_4d(xm, &full[stage], st + kBBytes + g * kABytes + b * kBoxBytes, mtile * 128 + b * 64,
                        w0 + p.taps.dw[tap], h0 + p.taps.dh[tap], n0);
        }
        if (++stage == stages) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc_bf16(128, BN, 1, 1);
      int stage = 0;
      uint32_t phase = 0;
      for (int c = c_begin; c < c_end; ++c) {
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        const uint32_t st = smem_u32(smem + stage * stage_bytes);
        const uint64_t b_desc = make_smem_desc(st, kBoxBytes, 1024);
        for (int g = 0; g < ntap; ++g) {
          const uint64_t a_desc = make_smem_desc(st + kBBytes + g * kABytes, kBoxBytes, 1024);
#pragma unroll
          for (int k = 0; k < 4; ++k) {  // 16 pixels (k rows of 128 B) per UMMA: +2048 B
            umma_bf16(tmem_base + g * BN, a_desc + 128 * k, b_desc + 128 * k, idesc, (c > c_begin) || (k > 0));
          }
        }
        umma_commit(&empty[stage]);
        if (++stage == stages) {
          stage = 0;
          phase ^= 1;
        }
      }
      umma_commit(tfull);
    }
  } else if (c_end > c_begin) {
    const int q = warp & 3;
    mbar_wait(tfull, 0);
    tc_fence_after();
    for (int g = 0; g < ntap; ++g) {
      int tap = p.group_over_m ? 0 : tap0 + g;
      int ci = (p.group_over_m ? mt0 + g : mt0) * 128 + q * 32 + lane;
      if (p.pair_taps) {  // rows 0..63 belong to tap 2s, rows 64..127 to tap 2s + 1
        tap = 2 * (tap0 + g) + (q >> 1);
        ci = (q & 1) * 32 + lane;
        if (tap >= p.num_taps) ci = p.cin;  // nothing to write
      }
      // one pixel split: read-modify-write of dW by its only writer; several: this split's slab of the workspace
      float* out = (p.ksplit > 1 ? p.partial + ks * p.partial_stride : p.dw) + static_cast<int64_t>(tap) * p.cin + ci;
#pragma unroll 1
      for (int chunk = 0; chunk < BN / 32; ++chunk) {
        uint32_t r[32];
        tmem_ld32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + g * BN + chunk * 32, r);
        tmem_ld_wait();
        if (ci < p.cin) {
          float* o = out + static_cast<int64_t>(nt * BN + chunk * 32) * p.ld_co;
          if (p.ksplit > 1) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              *o = __uint_as_float(r[j]);
              o += p.ld_co;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              *o += __uint_as_float(r[j]);
              o += p.ld_co;
            }
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

// dW += sum over the pixel splits' slabs, in slab order (fixed summation order -> reproducible weight gradients)
__global__ void __launch_bounds__(256) wgrad_reduce_kernel(const float4* __restrict__ partial, float4* dw, int64_t n4,
                                                           int ksplit) {
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n4; i += (int64_t)gridDim.x * 256) {
    float4 acc = dw[i];
#pragma unroll 4
    for (int ks = 0; ks < ksplit; ++ks) {
      const float4 v = __ldcs(partial + (int64_t)ks * n4 + i);
      acc.x += v.x, acc.y += v.y, acc.z += v.z, acc.w += v.w;
    }
    dw[i] = acc;
  }
}

// Few elements, many slabs (the 1x1 convolutions of the first stages: 4096 float4 x ~100 slabs): one thread per
// element would walk the slabs as one long chain of dependent loads (26 us for a 64 KB gradient).  Here S warps-rows of
// a block share each element: row s adds slabs s, s + S, ... (coalesced across the row), the S partial sums meet in
// shared memory and are added in row order -- a fixed order again, so the result stays bit-reproducible.
template <int S>
__global__ void __launch_bounds__(256) wgrad_reduce_split_kernel(const float4* __restrict__ partial, float4* dw,
                                                                 int64_t n4, int ksplit) {
  constexpr int kPos = 256 / S;
  __shared__ float4 part[S][kPos];
  const int s_row = threadIdx.x / kPos, col = threadIdx.x % kPos;
  const int64_t i = (int64_t)blockIdx.x * kPos + col;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (i < n4) {
#pragma unroll 4
    for (int ks = s_row; ks < ksplit; ks += S) {
      const float4 v = __ldcs(partial + (int64_t)ks * n4 + i);
      acc.x += v.x, acc.y += v.y, acc.z += v.z, acc.w += v.w;
    }
  }
  part[s_row][col] = acc;
  __syncthreads();
  if (s_row == 0 && i < n4) {
    float4 out = dw[i];
#pragma unroll
    for (int r = 0; r < S; ++r) {
      const float4 v = part[r][col];
      out.x += v.x, out.y += v.y, out.z += v.z, out.w += v.w;
    }
    dw[i] = out;
  }
}

// The same for MANY convolutions in one launch (a ResNet stage's worth): one table row per convolution, blocks are
// dealt to rows by their prefix sums.  53 per-convolution reductions of ~12 us each (mostly fixed launch cost) become 5.
struct ReduceEntry {
  const float4* partial;
  float4* dw;
  long long n4;
  int ksplit, blk_begin;
};
constexpr int kReduceF4PerBlock = 1024;
__global__ void __launch_bounds__(256) wgrad_reduce_batched_kernel(const ReduceEntry* __restrict__ table, int n) {
  int lo = 0, hi = n - 1;
  while (lo < hi) {  // last row with blk_begin <= blockIdx.x
    const int mid = (lo + hi + 1) >> 1;
    if (table[mid].blk_begin <= (int)blockIdx.x) lo = mid;
    else hi = mid - 1;
  }
  const ReduceEntry e = table[lo];
  const long long base = (long long)(blockIdx.x - e.blk_begin) * kReduceF4PerBlock;
#pragma unroll
  for (int j = 0; j < kReduceF4PerBlock / 256; ++j) {
    const long long i = base + j * 256 + threadIdx.x;
    if (i < e.n4) {
      float4 acc = e.dw[i];
      for (int ks = 0; ks < e.ksplit; ++ks) {
        const float4 v = __ldcs(e.partial + (long long)ks * e.n4 + i);
        acc.x += v.x, acc.y += v.y, acc.z += v.z, acc.w += v.w;
      }
      e.dw[i] = acc;
    }
  }
}

int wgrad_reduce_batched_launch(const void* table, int n, int total_blocks, cudaStream_t stream) {
  if (!table || n < 1 || total_blocks < 1) return PECLR_ERR_ARG;
  wgrad_reduce_batched_kernel<<<total_blocks, 256, 0, stream>>>(static_cast<const ReduceEntry*>(table), n);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : -(int)e;
}
int wgrad_reduce_f4_per_block() { return kReduceF4PerBlock; }

// ------------------------------------------------------------------------------------------ host
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) != cudaSuccess || !sym)
      return nullptr;
    fn = reinterpret_cast<EncodeTiledFn>(sym);
  }
  return fn;
}

// rank-4 bf16 view (C, W, H, N); strides in elements for W, H, N; box (64, bw, bh, bn); 128B swizzle.
static int encode_view(CUtensorMap* m, const View4& v, int bw, int bh, int bn) {
  if (bw > 256 || bh > 256 || bn > 256) return PECLR_ERR_ARG;
  EncodeTiledFn enc = get_encode();
  if (!enc) return PECLR_ERR_DRIVER;
  cuuint64_t dims[4] = {(cuuint64_t)v.c, (cuuint64_t)v.w, (cuuint64_t)v.h, (cuuint64_t)v.n};
  cuuint64_t strides[3] = {(cuuint64_t)v.sw * 2, (cuuint64_t)v.sh * 2, (cuuint64_t)v.sn * 2};
  cuuint32_t box[4] = {64, (cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)bn};
  cuuint32_t es[4] = {1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(v.ptr), dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    fprintf(stderr, "peclr: cuTensorMapEncodeTiled(4d) failed: %d dims=(%llu,%llu,%llu,%llu) strides=(%llu,%llu,%llu) box=(%d,%d,%d)\n",
            (int)r, (unsigned long long)dims[0], (unsigned long long)dims[1], (unsigned long long)dims[2],
            (unsigned long long)dims[3], (unsigned long long)strides[0], (unsigned long long)strides[1],
            (unsigned long long)strides[2], bw, bh, bn);
    return PECLR_ERR_TENSORMAP;
  }
  return 0;
}

static int encode_matrix(CUtensorMap* m, const void* ptr, int64_t cols, int64_t rows, int box_rows) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return PECLR_ERR_DRIVER;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
  cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
  cuuint32_t es[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    fprintf(stderr, "peclr: cuTensorMapEncodeTiled(2d) failed: %d cols=%lld rows=%lld\n", (int)r, (long long)cols,
            (long long)rows);
    return PECLR_ERR_TENSORMAP;
  }
  return 0;
}

// Chooses the pixel box (bw, bh, bn) with bw*bh*bn == target (a power of two) that needs the fewest tiles.
void choose_box(int W, int H, int N, int target, int* bw, int* bh, int* bn) {
  long best = -1;
  for (int w = 1; w <= target; w *= 2)
    for (int h = 1; w * h <= target; h *= 2) {
      const int n = target / (w * h);
      if (w > 256 || h > 256 || n > 256) continue;
      const long tiles = (long)((W + w - 1) / w) * ((H + h - 1) / h) * ((N + n - 1) / n);
      // fewer tiles first; then larger spatial patches (filter-halo reuse in L2), then wider boxes
      if (best < 0 || tiles < best ||
          (tiles == best && (w * h > *bw * *bh || (w * h == *bw * *bh && w > *bw)))) {
        best = tiles;
        *bw = w;
        *bh = h;
        *bn = n;
      }
    }
}

// Halo mode: one image per tile (Nb = 1) and a box width that is a multiple of 8 pixels, so that a shift by whole
// image rows inside the shared-memory box keeps the 1024-byte swizzle-atom alignment.  Fewest tiles first, then the
// tallest box (smallest halo overhead (bh + taps - 1) / bh).
static void choose_box_halo(int W, int H, int* bw, int* bh) {
  long best = -1;
  for (int w = 8; w <= 128; w *= 2) {
    const int h = 128 / w;
    const long tiles = (long)((W + w - 1) / w) * ((H + h - 1) / h);
    if (best < 0 || tiles < best || (tiles == best && h > *bh)) {
      best = tiles;
      *bw = w;
      *bh = h;
    }
  }
}

static int sm_count() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

static bool double_d_enabled() {  // PECLR_CONV_DOUBLE_D=0: single output staging tile (A/B runs)
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("PECLR_CONV_DOUBLE_D");
    v = e ? atoi(e) : 1;
  }
  return v != 0;
}

template <int BN, int STAGES, int CTAS, bool HALO = false>
static int launch_gemm_t(GemmParams& p, bool stats, cudaStream_t stream) {
  constexpr size_t kMaxSmem = 227 * 1024;  // opt-in dynamic shared memory per CTA on sm_100
  // everything but the pipeline stages: alignment slack, output staging (+ y tile), barriers
  const size_t tile_d = (size_t)(BN / 64) * kStageA;
  size_t smem = 1024 + tile_d * (p.bn_reduce == 2 ? 4 : (p.bn_reduce ? 2 : 1)) + 256;
  p.stat_copies = 0;
  if (stats) {
    p.stat_copies = 512 / BN;
    const size_t stages_min = HALO ? 2 * ((size_t)p.halo_stage_a + (size_t)p.halo_taps * (BN / CTAS) * 128)
                                   : (size_t)STAGES * (kStageA + (BN / CTAS) * 128);
    if (smem + stages_min + (size_t)p.stat_copies * 2 * p.cout * 4 > kMaxSmem) p.stat_copies = 1;
    smem += (size_t)p.stat_copies * 2 * p.cout * 4;
  }
  p.d_bufs = 1;
  if constexpr (HALO) {
    const size_t per_stage = (size_t)p.halo_stage_a + (size_t)p.halo_taps * (BN / CTAS) * 128;
    int stages = (int)((kMaxSmem - smem) / per_stage);
    if (stages > STAGES) stages = STAGES;
    if (stages < 2) return PECLR_ERR_ARG;
    p.halo_stages = p.n_stages = stages;
    smem += (size_t)stages * per_stage;
  } else {
    const size_t per_stage = kStageA + (BN / CTAS) * 128;
    p.n_stages = STAGES;
    // a second output staging tile only where it costs no pipeline stage: per-launch A/B (profiles/ab_r02.txt) shows
    // -0.5 us on the launches that keep their stages and +2..10 us on the ones that gave up two of six
    if (double_d_enabled() && smem + tile_d + (size_t)STAGES * per_stage <= kMaxSmem) {
      p.d_bufs = 2;
      smem += tile_d;
    }
    smem += (size_t)p.n_stages * per_stage;
  }
  if (smem > kMaxSmem) return PECLR_ERR_ARG;
  auto kern = conv_gemm_kernel<BN, STAGES, CTAS, HALO>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return -(int)e;
  const int m_tiles = p.tiles_w * p.tiles_h * p.tiles_n;
  const int total = ((m_tiles + CTAS - 1) / CTAS) * p.n_tiles;  // (pairs of) tiles
  const int slots = sm_count() / CTAS;
  const int grid = CTAS * (total < slots ? total : slots);
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(320);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CTAS;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  e = cudaLaunchKernelEx(&cfg, kern, p);
  return e == cudaSuccess ? 0 : -(int)e;
}

bool pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    // measured on B200 (same box, ResNet-50 / ResNet-152 steps): 18.61 vs 18.02 ms and 42.36 vs 40.45 ms WITH vs
    // without -- successors that become resident early take the SM slots the side stream's weight-gradient kernels
    // would have overlapped into.  Off by default; kept as a switch.
    const char* e = getenv("PECLR_PDL");
    v = e ? atoi(e) : 0;
  }
  return v != 0;
}

static bool use_cta_pairs() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("PECLR_CONV_CTA2");
    v = e ? atoi(e) : 1;
  }
  return v != 0;
}

// The 3x3 / stride-1 convolutions with 64 or 128 output channels and the stem use the filter-row halo kernels
// (measured -0.12 ms per ResNet-50 step on the same box; PECLR_CONV_HALO=0 selects the per-tap kernels for A/B runs).
bool conv_halo_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("PECLR_CONV_HALO");
    v = (e ? atoi(e) : 1) != 0 && use_cta_pairs();  // the halo kernels exist as CTA-pair kernels only
  }
  return v != 0;
}

int conv_gemm_launch(const View4* a_views, int num_views, const void* w, int64_t ktot, int64_t nout,
                     const View4& d_view, const TapTable& taps, int num_taps, int c_chunks, double* stat_sum,
                     double* stat_sumsq, int reduce_add, cudaStream_t stream, const BnReduce* bnr, int halo_taps,
                     int halo_kstep) {
  if (num_views < 1 || num_views > kMaxViews || num_taps < 1 || num_taps > kMaxTaps) return PECLR_ERR_ARG;
  if (nout % 64 != 0 || ktot % 64 != 0 || d_view.c != nout) return PECLR_ERR_ARG;
  if (stat_sum && nout > 2048) return PECLR_ERR_ARG;
  GemmParams p;
  memset(&p, 0, sizeof(p));
  int bw = 1, bh = 1, bn = 1;
  const bool halo = halo_taps > 1;
  if (halo) {
    // (callers only group taps for nout in {64, 128}, more than one m tile, single-view stride-1 geometry)
    if (nout % 256 == 0 || d_view.w * d_view.h * d_view.n <= 128 || !use_cta_pairs()) return PECLR_ERR_ARG;
    choose_box_halo((int)d_view.w, (int)d_view.h, &bw, &bh);
    p.halo_taps = halo_taps;
    p.halo_kstep = halo_kstep;
    p.halo_a_bytes = (bh + halo_taps - 1) * bw * 128;
    p.halo_stage_a = (p.halo_a_bytes + 1023) / 1024 * 1024;
  } else {
    choose_box((int)d_view.w, (int)d_view.h, (int)d_view.n, 128, &bw, &bh, &bn);
  }
  int rc;
  for (int i = 0; i < kMaxViews; ++i)
    if ((rc = encode_view(&p.a_maps[i], a_views[i < num_views ? i : 0], bw, halo ? bh + halo_taps - 1 : bh, bn)))
      return rc;
  int BN = nout % 256 == 0 ? 256 : (nout % 128 == 0 ? 128 : 64);
  static int bn_max = -1;  // tuning knob: PECLR_CONV_BN_MAX caps the tile width (64 / 128 / 256)
  if (bn_max < 0) {
    const char* e = getenv("PECLR_CONV_BN_MAX");
    bn_max = e ? atoi(e) : 256;
  }
  if (BN > bn_max && bn_max >= 64) BN = bn_max;
  if (bnr) {
    if (!stat_sum || !stat_sumsq || reduce_add) return PECLR_ERR_ARG;
    if (BN > 128) BN = 128;  // room for the y tile in shared memory
    View4 yv = d_view;
    yv.ptr = bnr->y;
    if ((rc = encode_view(&p.y_map, yv, bw, bh, bn))) return rc;
    p.bn_mean = bnr->mean, p.bn_invstd = bnr->invstd, p.bn_gamma = bnr->gamma, p.bn_beta = bnr->beta;
    p.bn_reduce = 1;
    if (bnr->mask_bits) {
      if (halo || nout % 128 != 0) return PECLR_ERR_ARG;
      p.bn_reduce = 2;
      p.mask_bits = bnr->mask_bits;
      p.pix_base = bnr->pix_base, p.pix_w = bnr->pix_w, p.pix_h = bnr->pix_h, p.pix_n = bnr->pix_n;
      p.mask_row_bytes = (int)(nout / 8);
      p.acc_stride = bnr->acc_stride;
      p.fd_img_w = make_fastdiv(bnr->img_w > 0 ? bnr->img_w : 1);
      if (p.acc_stride != 1 && (p.acc_stride != 2 || bnr->img_w < 2 || (bnr->img_w & 1))) return PECLR_ERR_ARG;
    }
  }
  const bool pairs = use_cta_pairs() && d_view.w * d_view.h * d_view.n > 128;  // at least two m tiles
  if ((rc = encode_matrix(&p.b_map, w, ktot, nout, pairs ? BN / 2 : BN))) return rc;
  if ((rc = encode_view(&p.d_map, d_view, bw, bh, bn))) return rc;
  p.taps = taps;
  p.num_taps = num_taps;
  p.c_chunks = c_chunks;
  p.Wb = bw, p.Hb = bh, p.Nb = bn;
  p.log_wb = __builtin_ctz(bw);
  p.log_wbhb = __builtin_ctz(bw * bh);
  p.d_w = (int)d_view.w, p.d_h = (int)d_view.h, p.d_n = (int)d_view.n;
  p.tiles_w = (int)((d_view.w + bw - 1) / bw);
  p.tiles_h = (int)((d_view.h + bh - 1) / bh);
  p.tiles_n = (int)((d_view.n + bn - 1) / bn);
  p.n_tiles = (int)(nout / BN);
  p.fd_tiles_w = make_fastdiv(p.tiles_w), p.fd_tiles_h = make_fastdiv(p.tiles_h), p.fd_n_tiles = make_fastdiv(p.n_tiles);
  p.cout = (int)nout;
  p.stat_sum = stat_sum;
  p.stat_sumsq = stat_sumsq;
  p.reduce_add = reduce_add;
  const bool stats = stat_sum != nullptr;
  if (halo) {
    if (BN == 128) return launch_gemm_t<128, 8, 2, true>(p, stats, stream);
    if (BN == 64) return launch_gemm_t<64, 8, 2, true>(p, stats, stream);
    return PECLR_ERR_ARG;
  }
  if (p.bn_reduce == 2) {  // four staging tiles (output, y, 2 x gradient-so-far): fewer pipeline stages
    if (pairs) return launch_gemm_t<128, 3, 2>(p, stats, stream);
    return launch_gemm_t<128, 2, 1>(p, stats, stream);
  }
  if (pairs) {
    if (BN == 256) return launch_gemm_t<256, 4, 2>(p, stats, stream);
    if (BN == 128) return launch_gemm_t<128, 6, 2>(p, stats, stream);
    return launch_gemm_t<64, 8, 2>(p, stats, stream);
  }
  if (BN == 256) return launch_gemm_t<256, 3, 1>(p, stats, stream);
  if (BN == 128) return launch_gemm_t<128, 4, 1>(p, stats, stream);
  return launch_gemm_t<64, 6, 1>(p, stats, stream);
}

template <int BN>
static int launch_wgrad_t(const WgradParams& p, cudaStream_t stream) {
  const size_t smem = 1024 + 200 * 1024 + 1024;
  auto kern = conv_wgrad_kernel<BN>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return -(int)e;
  const int grid = p.m_tiles * p.n_tiles * p.tap_groups * p.ksplit;
  kern<<<grid, 192, smem, stream>>>(p);
  e = cudaGetLastError();
  return e == cudaSuccess ? 0 : -(int)e;
}

// Tiling of a weight-gradient launch (everything but the tensor maps): shared by the launch and the workspace query.
static int wgrad_plan(const View4& dy_view, int num_taps, int cin, int cout, WgradParams* pp, int* bn_out) {
  WgradParams& p = *pp;
  int bw = 1, bh = 1, bn = 1;
  choose_box((int)dy_view.w, (int)dy_view.h, (int)dy_view.n, 64, &bw, &bh, &bn);
  // multi-tap filters keep several taps' accumulators side by side in TMEM (dY is loaded once per group)
  const int BN = num_taps > 1 ? (cout % 128 == 0 ? 128 : 64) : (cout % 256 == 0 ? 256 : (cout % 128 == 0 ? 128 : 64));
  p.num_taps = num_taps;
  p.taps_per_unit = num_taps == 9 ? 3 : (num_taps < 4 ? num_taps : 4);  // G * BN <= 512 columns, >= 2 smem stages
  p.tap_groups = (num_taps + p.taps_per_unit - 1) / p.taps_per_unit;
  p.m_tiles = (cin + 127) / 128;
  if (num_taps > 1 && cin == 64) {
    p.pair_taps = 1;
    p.a_boxes = 2;
    const int slots = (num_taps + 1) / 2;
    p.taps_per_unit = slots == 5 ? 3 : 2;
    p.tap_groups = (slots + p.taps_per_unit - 1) / p.taps_per_unit;
  }
  if (num_taps == 1 && cin >= 256) {
    // dY (the wide operand) is loaded once per pixel chunk and multiplied with G input-channel tiles
    p.group_over_m = 1;
    p.taps_per_unit = 512 / BN < p.m_tiles ? 512 / BN : p.m_tiles;
    if (p.taps_per_unit > 4) p.taps_per_unit = 4;  // keep >= 2 pipeline stages in shared memory
    p.tap_groups = 1;
    p.m_tiles = (p.m_tiles + p.taps_per_unit - 1) / p.taps_per_unit;
  }
  p.n_tiles = cout / BN;
  p.a_boxes = (cin >= 128 || p.pair_taps) ? 2 : 1;
  p.Wb = bw, p.Hb = bh, p.Nb = bn;
  p.tiles_w = (int)((dy_view.w + bw - 1) / bw);
  p.tiles_h = (int)((dy_view.h + bh - 1) / bh);
  p.tiles_n = (int)((dy_view.n + bn - 1) / bn);
  const int chunks = p.tiles_w * p.tiles_h * p.tiles_n;
  const int base_units = p.m_tiles * p.n_tiles * p.tap_groups;
  static int split_waves = -1;  // tuning knob: PECLR_WGRAD_WAVES = target CTAs per SM (default 1, measured best)
  if (split_waves < 0) {
    const char* e = getenv("PECLR_WGRAD_WAVES");
    split_waves = e ? atoi(e) : 1;
    if (split_waves < 1) split_waves = 1;
  }
  // as many pixel splits as keep the WHOLE grid resident at once (one CTA per SM: 200 KB of shared memory, all of
  // TMEM): rounding up instead put e.g. 160 or 192 CTAs on 148 SMs, a second wave that ran on a handful of SMs and
  // doubled the kernel's duration (profiles/conv_wgrad_r02_ncu_full.txt: SMs active 50 % of the elapsed time)
  int ksplit = (split_waves * sm_count()) / base_units;
  const int max_split = (chunks + 7) / 8;  // at least 8 pixel chunks (512 pixels) per unit
  if (ksplit > max_split) ksplit = max_split;
  if (ksplit < 1) ksplit = 1;
  // no empty pixel ranges: every split writes its whole slab, the reduction reads all of them
  const int per = (chunks + ksplit - 1) / ksplit;
  ksplit = (chunks + per - 1) / per;
  p.ksplit = ksplit;
  p.cin = cin;
  p.ld_co = (int64_t)num_taps * cin;
  p.partial_stride = (int64_t)cout * num_taps * cin;
  *bn_out = BN;
  return 0;
}

long long conv_wgrad_workspace_bytes(const View4& dy_view, int num_taps, int cin, int cout) {
  if (num_taps < 1 || num_taps > kMaxTaps || cin % 64 != 0 || cout % 64 != 0) return PECLR_ERR_ARG;
  WgradParams p;
  memset(&p, 0, sizeof(p));
  int BN = 0;
  wgrad_plan(dy_view, num_taps, cin, cout, &p, &BN);
  return p.ksplit > 1 ? (long long)p.ksplit * p.partial_stride * 4 : 0;
}

int conv_wgrad_splits(const View4& dy_view, int num_taps, int cin, int cout) {
  if (num_taps < 1 || num_taps > kMaxTaps || cin % 64 != 0 || cout % 64 != 0) return PECLR_ERR_ARG;
  WgradParams p;
  memset(&p, 0, sizeof(p));
  int BN = 0;
  wgrad_plan(dy_view, num_taps, cin, cout, &p, &BN);
  return p.ksplit;
}

int conv_wgrad_launch(const View4* x_views, int num_views, const View4& dy_view, const TapTable& taps, int num_taps,
                      int cin, int cout, float* dw, void* workspace, long long workspace_bytes, cudaStream_t stream,
                      bool defer_reduce) {
  if (num_views < 1 || num_views > kMaxViews || num_taps < 1 || num_taps > kMaxTaps) return PECLR_ERR_ARG;
  if (cin % 64 != 0 || cout % 64 != 0 || dy_view.c != cout) return PECLR_ERR_ARG;
  WgradParams p;
  memset(&p, 0, sizeof(p));
  int BN = 0;
  wgrad_plan(dy_view, num_taps, cin, cout, &p, &BN);
  if (p.ksplit > 1 && (!workspace || workspace_bytes < (long long)p.ksplit * p.partial_stride * 4)) return PECLR_ERR_ARG;
  int rc;
  for (int i = 0; i < kMaxViews; ++i)
    if ((rc = encode_view(&p.x_maps[i], x_views[i < num_views ? i : 0], p.Wb, p.Hb, p.Nb))) return rc;
  if ((rc = encode_view(&p.dy_map, dy_view, p.Wb, p.Hb, p.Nb))) return rc;
  p.taps = taps;
  p.dw = dw;
  p.partial = static_cast<float*>(workspace);
  if (BN == 256) rc = launch_wgrad_t<256>(p, stream);
  else if (BN == 128) rc = launch_wgrad_t<128>(p, stream);
  else rc = launch_wgrad_t<64>(p, stream);
  if (rc || p.ksplit == 1 || defer_reduce) return rc;
  const int64_t n4 = p.partial_stride / 4;
  int64_t blocks = (n4 + 255) / 256;
  const float4* part = reinterpret_cast<const float4*>(p.partial);
  float4* out = reinterpret_cast<float4*>(dw);
  // share each element among S thread rows while that still fills the machine and leaves every row >= 4 slabs
  int S = 1;
  while (S < 16 && blocks * (2 * S) <= 2 * sm_count() && p.ksplit >= 8 * S) S *= 2;
  if (S == 1) {
    if (blocks > 4 * sm_count()) blocks = 4 * sm_count();
    wgrad_reduce_kernel<<<(int)blocks, 256, 0, stream>>>(part, out, n4, p.ksplit);
  } else if (S == 2) {
    wgrad_reduce_split_kernel<2><<<(int)((n4 + 127) / 128), 256, 0, stream>>>(part, out, n4, p.ksplit);
  } else if (S == 4) {
    wgrad_reduce_split_kernel<4><<<(int)((n4 + 63) / 64), 256, 0, stream>>>(part, out, n4, p.ksplit);
  } else if (S == 8) {
    wgrad_reduce_split_kernel<8><<<(int)((n4 + 31) / 32), 256, 0, stream>>>(part, out, n4, p.ksplit);
  } else {
    wgrad_reduce_split_kernel<16><<<(int)((n4 + 15) / 16), 256, 0, stream>>>(part, out, n4, p.ksplit);
  }
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : -(int)e;
}

}  // namespace peclr
