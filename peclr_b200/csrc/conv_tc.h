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

// Cross-CTA reductions are reproducible by construction: per-channel statistics (forward BatchNorm sums, BN-backward
// sums) are added as ONE fp64 partial per CTA (each computed in a fixed order) into fp64 accumulators, whose totals
// do not depend on the arrival order; weight-gradient pixel splits go through workspace slabs added in slab order.
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

// Programmatic dependent launch for the kernels of the main stream's dependency chain (ptx.cuh: pdl_wait), enabled
// with PECLR_PDL=1.  Default off: measured slower on the full step (see pdl_enabled() in conv_tc.cu).
bool pdl_enabled();
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                     Args&&... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid, cfg.blockDim = block, cfg.dynamicSmemBytes = smem, cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// Fused BatchNorm-backward reduction in the epilogue: y = input of the BatchNorm (+ReLU) whose output gradient this
// GEMM produces (same geometry as D); stat_sum / stat_sumsq then receive sum g and sum g*y per channel.
struct BnReduce {
  const void* y;
  const float *mean, *invstd, *gamma, *beta;
  // finish mode (mask_bits != nullptr; mean .. beta unused): D = (D_in_tensor + GEMM) masked with the ReLU bits of the
  // block output D is the gradient of ([pixels][C/8] bytes as bn_apply writes them); sums of g and g*y as above.
  // Pixel index of view element (w, h, n) = pix_base + w*pix_w + h*pix_h + n*pix_n.
  const uint8_t* mask_bits = nullptr;
  long long pix_base = 0, pix_w = 0, pix_h = 0, pix_n = 0;
  // finish mode: 2 = only the pixels with even image row AND even image column of D_in_tensor hold a gradient (the
  // scatter of a stride-2 1x1 dgrad into a buffer that was never zeroed); the rest is read as 0.  img_w = image width.
  int acc_stride = 1, img_w = 0;
};

// D[pix, n] (+)= sum_taps sum_c A_view[pix + tap, c] * Wmat[n, koff(tap) + c]; optional column statistics.
int conv_gemm_launch(const View4* a_views, int num_views, const void* w, int64_t ktot, int64_t nout,
                     const View4& d_view, const TapTable& taps, int num_taps, int c_chunks, double* stat_sum,
                     double* stat_sumsq, int reduce_add, cudaStream_t stream, const BnReduce* bnr = nullptr,
                     int halo_taps = 0, int halo_kstep = 0);
// halo_taps > 1: every tap-table entry is a group of halo_taps taps with dh = entry.dh + t (t = 0 .. halo_taps - 1),
// whose weights are halo_kstep K-columns apart; the kernel loads one tall input box per group (see conv_tc.cu).
bool conv_halo_enabled();

// dW[n, tap, c] += sum_pix dY[pix, n] * X_view[pix + tap, c]   (fp32, ld between n = num_taps * cin).  The pixel
// dimension is split over CTAs; with more than one split the partials go through `workspace`
// (conv_wgrad_workspace_bytes) and are added to dW in a fixed order by a second kernel.
long long conv_wgrad_workspace_bytes(const View4& dy_view, int num_taps, int cin, int cout);
int conv_wgrad_launch(const View4* x_views, int num_views, const View4& dy_view, const TapTable& taps, int num_taps,
                      int cin, int cout, float* dw, void* workspace, long long workspace_bytes, cudaStream_t stream,
                      bool defer_reduce = false);
// defer_reduce: only the pixel splits' slabs are written; the caller adds them to dW later, for many convolutions in
// one launch (wgrad_reduce_batched_launch; table rows {partial, dw, n4, ksplit, first block}, see conv_tc.cu).
int conv_wgrad_splits(const View4& dy_view, int num_taps, int cin, int cout);
int wgrad_reduce_batched_launch(const void* table, int n, int total_blocks, cudaStream_t stream);
int wgrad_reduce_f4_per_block();

}  // namespace peclr
