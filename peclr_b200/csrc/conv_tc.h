// Internal interface between the tensor-core GEMM kernels (conv_tc.cu) and the convolution
// front-ends (conv_ops.cu).  Not part of the C ABI (see include/peclr_b200.h for that).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#define PECLR_ERR_ARG (-1001)
#define PECLR_ERR_DRIVER (-1002)
#define PECLR_ERR_TENSORMAP (-1003)

namespace peclr {

// The BN-backward sums (scratch buffers) are kept as kStatReplicas fp32 accumulator sets: block b adds its partial
// into set b % kStatReplicas, the consumer adds the sets in order.  Replication divides the number of atomics
// serialised on one address (the tail of the reduction kernels; measured 4-12 us per launch).  The forward
// statistics of the conv epilogue are one fp64 set: fp64 makes the totals independent of the order in which the
// CTAs' partials arrive (the forward pass, hence the loss, is reproducible run to run); the backward sums stay fp32
// because fp64 arithmetic in the consumers' per-thread prologues is slow on this part (+0.4 ms per step measured).
constexpr int kStatReplicas = 4;
constexpr int kMaxViews = 4;
constexpr int kMaxTaps = 16;

// bf16 tensor seen as (C, W, H, N) with C contiguous; strides in elements.
struct View4 {
  const void* ptr;
  int64_t c, w, h, n;
  int64_t sw, sh, sn;
};

// One entry per filter tap: which view to read, the pixel offset inside that view's lattice and the
// column offset of the tap's weights inside the K dimension of the weight matrix.
struct TapTable {
  int8_t view[kMaxTaps];
  int8_t dw[kMaxTaps];
  int8_t dh[kMaxTaps];
  int32_t koff[kMaxTaps];
};

void choose_box(int W, int H, int N, int target, int* bw, int* bh, int* bn);

// Fused BatchNorm-backward reduction in the epilogue: y = input of the BatchNorm (+ReLU) whose output gradient this
// GEMM produces (same geometry as D); stat_sum / stat_sumsq then receive sum g and sum g*y per channel.
struct BnReduce {
  const void* y;
  const float *mean, *invstd, *gamma, *beta;
};

// D[pix, n] (+)= sum_taps sum_c A_view[pix + tap, c] * Wmat[n, koff(tap) + c]; optional column statistics.
int conv_gemm_launch(const View4* a_views, int num_views, const void* w, int64_t ktot, int64_t nout,
                     const View4& d_view, const TapTable& taps, int num_taps, int c_chunks, void* stat_sum,
                     void* stat_sumsq, int64_t stat_stride, int reduce_add, cudaStream_t stream, const BnReduce* bnr = nullptr,
                     int halo_taps = 0, int halo_kstep = 0);
// halo_taps > 1: every tap-table entry is a group of halo_taps taps with dh = entry.dh + t (t = 0 .. halo_taps - 1),
// whose weights are halo_kstep K-columns apart; the kernel loads one tall input box per group (see conv_tc.cu).
bool conv_halo_enabled();

// dW[n, tap, c] += sum_pix dY[pix, n] * X_view[pix + tap, c]   (fp32, ld between n = num_taps * cin)
int conv_wgrad_launch(const View4* x_views, int num_views, const View4& dy_view, const TapTable& taps, int num_taps,
                      int cin, int cout, float* dw, cudaStream_t stream);

}  // namespace peclr
