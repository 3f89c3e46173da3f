// Convolution front-ends: turn (N, H, W, Cin, Cout, k, stride) into TMA views + tap tables for the
// tensor-core kernels in conv_tc.cu.  C ABI declared in include/peclr_b200.h.
//
// Replaces, for the ResNet trunk of the reference (src/models/resnet_model.py:16-26, torchvision
// Bottleneck/BasicBlock convolutions), the cuDNN calls behind nn.Conv2d forward / backward.
#include "../../include/peclr_b200.h"
#include "conv_tc.h"

namespace peclr {

static View4 nhwc_view(const void* ptr, int64_t N, int64_t H, int64_t W, int64_t C) {
  return View4{ptr, C, W, H, N, C, W * C, H * W * C};
}
static View4 flat_view(const void* ptr, int64_t M, int64_t C) { return View4{ptr, C, M, 1, 1, C, M * C, M * C}; }
// pixels (2i + ph, 2j + pw) of an NHWC image as their own (C, W/2, H/2, N) tensor
static View4 parity_view(const void* ptr, int64_t N, int64_t H, int64_t W, int64_t C, int ph, int pw) {
  const char* base = static_cast<const char*>(ptr) + ((int64_t)ph * W + pw) * C * 2;
  return View4{base, C, W / 2, H / 2, N, 2 * C, 2 * W * C, H * W * C};
}

// Views + taps of the INPUT side of a k x k / stride s convolution with padding (k-1)/2, expressed on
// the output pixel lattice.  kstride = elements between consecutive taps in the weight matrix.
static int input_side(const void* x, int N, int H, int W, int C, int k, int stride, int kstride, View4* views,
                      TapTable* taps, int* num_taps) {
  memset(taps, 0, sizeof(*taps));
  if (k != 1 && k != 3) return PECLR_ERR_ARG;
  if (stride == 1) {
    views[0] = nhwc_view(x, N, H, W, C);
    int t = 0;
    for (int r = 0; r < k; ++r)
      for (int s = 0; s < k; ++s, ++t) {
        taps->view[t] = 0;
        taps->dh[t] = (int8_t)(r - k / 2);
        taps->dw[t] = (int8_t)(s - k / 2);
        taps->koff[t] = t * kstride;
      }
    *num_taps = t;
    return 1;
  }
  if (stride != 2 || (H & 1) || (W & 1)) return PECLR_ERR_ARG;
  for (int ph = 0; ph < 2; ++ph)
    for (int pw = 0; pw < 2; ++pw) views[ph * 2 + pw] = parity_view(x, N, H, W, C, ph, pw);
  int t = 0;
  for (int r = 0; r < k; ++r)
    for (int s = 0; s < k; ++s, ++t) {
      const int oh = r - k / 2, ow = s - k / 2;  // input pixel = 2 * out + o
      const int ph = oh & 1, pw = ow & 1;
      taps->view[t] = (int8_t)(ph * 2 + pw);
      taps->dh[t] = (int8_t)((oh - ph) / 2);
      taps->dw[t] = (int8_t)((ow - pw) / 2);
      taps->koff[t] = t * kstride;
    }
  *num_taps = t;
  return 4;
}

// Filter-row halo form of a 3 x 3 / stride-1 tap table (see conv_tc.cu): one entry per filter COLUMN s, covering the
// three taps of that column; entry.dh is the topmost input row offset (-1), k0(s) the K-column of the tap that reads
// it and kstep the K-distance to the tap reading the next row down.
static bool halo_ok(int nout, int64_t pixels) { return conv_halo_enabled() && nout % 256 != 0 && pixels > 128; }
static void halo_taps_3x3(TapTable* taps, bool dgrad, int kc) {
  memset(taps, 0, sizeof(*taps));
  for (int s = 0; s < 3; ++s) {
    taps->view[s] = 0;
    taps->dh[s] = -1;
    // fprop: tap (r, s) reads input (h + r - 1, w + s - 1), weights at (r * 3 + s) * Cin
    // dgrad: tap (r, s) reads dy (h + 1 - r, w + 1 - s), weights at (r * 3 + s) * Cout: row offset -1 is r = 2
    taps->dw[s] = (int8_t)(dgrad ? 1 - s : s - 1);
    taps->koff[s] = (dgrad ? 6 + s : s) * kc;
  }
}

}  // namespace peclr

using namespace peclr;

extern "C" int peclr_conv2d_fprop(const void* x, const void* w, void* y, int N, int H, int W, int Cin, int Cout, int k,
                                  int stride, double* stat_sum, double* stat_sumsq, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (Cin % 64 || Cout % 64) return PECLR_ERR_ARG;
  const int Ho = H / stride, Wo = W / stride;
  TapTable taps;
  int num_taps = 0;
  View4 views[kMaxViews];
  if (k == 1 && stride == 1) {
    memset(&taps, 0, sizeof(taps));
    views[0] = flat_view(x, (int64_t)N * H * W, Cin);
    View4 d = flat_view(y, (int64_t)N * H * W, Cout);
    return conv_gemm_launch(views, 1, w, Cin, Cout, d, taps, 1, Cin / 64, stat_sum, stat_sumsq, 0, st);
  }
  const int nv = input_side(x, N, H, W, Cin, k, stride, Cin, views, &taps, &num_taps);
  if (nv < 0) return nv;
  View4 d = nhwc_view(y, N, Ho, Wo, Cout);
  if (k == 3 && stride == 1 && halo_ok(Cout, (int64_t)N * Ho * Wo)) {
    halo_taps_3x3(&taps, false, Cin);
    return conv_gemm_launch(views, 1, w, (int64_t)9 * Cin, Cout, d, taps, 3, Cin / 64, stat_sum, stat_sumsq, 0, st,
                            nullptr, 3, 3 * Cin);
  }
  return conv_gemm_launch(views, nv, w, (int64_t)num_taps * Cin, Cout, d, taps, num_taps, Cin / 64, stat_sum,
                          stat_sumsq, 0, st);
}

// dgrad, optionally with the BatchNorm-backward reduction of the BN (+ReLU) in front of this convolution fused
// into the epilogue (bn_y = that BN's input, same shape as dx; scratch[0:Cin] += sum g, scratch[Cin:2Cin] += sum g*y)
static int dgrad_impl(const void* dy, const void* wt, void* dx, int N, int H, int W, int Cin, int Cout, int k,
                      int stride, int accumulate, const void* bn_y, const float* bn_mean, const float* bn_invstd,
                      const float* bn_gamma, const float* bn_beta, double* scratch, cudaStream_t st) {
  if (Cin % 64 || Cout % 64 || (k != 1 && k != 3)) return PECLR_ERR_ARG;
  if (bn_y && (accumulate || !scratch || Cin > 2048)) return PECLR_ERR_ARG;
  if (accumulate < 0 || accumulate > 2 || (accumulate == 2 && !(k == 1 && stride == 2))) return PECLR_ERR_ARG;
  const int Ho = H / stride, Wo = W / stride;
  TapTable taps;
  memset(&taps, 0, sizeof(taps));
  View4 a;
  const int64_t ktot = (int64_t)k * k * Cout;
  double* s_sum = bn_y ? scratch : nullptr;
  double* s_sq = bn_y ? scratch + Cin : nullptr;
  if (bn_y) {
    cudaError_t e = cudaMemsetAsync(scratch, 0, (size_t)2 * Cin * sizeof(double), st);
    if (e != cudaSuccess) return -(int)e;
  }
  // the y tile of a launch sits at the same offset from bn_y as the output view does from dx
  auto bnr_for = [&](const View4& d, BnReduce* store) -> const BnReduce* {
    if (!bn_y) return nullptr;
    const ptrdiff_t off = static_cast<const char*>(d.ptr) - static_cast<const char*>(dx);
    *store = BnReduce{static_cast<const char*>(bn_y) + off, bn_mean, bn_invstd, bn_gamma, bn_beta};
    return store;
  };
  BnReduce br;
  if (stride == 1) {
    if (k == 1) {
      a = flat_view(dy, (int64_t)N * H * W, Cout);
      View4 d = flat_view(dx, (int64_t)N * H * W, Cin);
      return conv_gemm_launch(&a, 1, wt, ktot, Cin, d, taps, 1, Cout / 64, s_sum, s_sq, accumulate, st,
                              bnr_for(d, &br));
    }
    a = nhwc_view(dy, N, H, W, Cout);
    int t = 0;
    for (int r = 0; r < 3; ++r)
      for (int s = 0; s < 3; ++s, ++t) {
        taps.dh[t] = (int8_t)(1 - r);
        taps.dw[t] = (int8_t)(1 - s);
        taps.koff[t] = t * Cout;
      }
    View4 d = nhwc_view(dx, N, H, W, Cin);
    if (halo_ok(Cin, (int64_t)N * H * W)) {
      halo_taps_3x3(&taps, true, Cout);
      return conv_gemm_launch(&a, 1, wt, ktot, Cin, d, taps, 3, Cout / 64, s_sum, s_sq, accumulate, st,
                              bnr_for(d, &br), 3, -3 * Cout);
    }
    return conv_gemm_launch(&a, 1, wt, ktot, Cin, d, taps, 9, Cout / 64, s_sum, s_sq, accumulate, st,
                            bnr_for(d, &br));
  }
  if (stride != 2 || (H & 1) || (W & 1)) return PECLR_ERR_ARG;
  a = nhwc_view(dy, N, Ho, Wo, Cout);
  if (k == 1) {
    if (bn_y) return PECLR_ERR_ARG;  // (3/4 of dx is plain zero here; no caller needs the fusion)
    // only the even/even pixels of dx receive gradient; accumulate == 2: plain scatter, the caller promises that
    // the other pixels are never read as they are (peclr_conv2d_dgrad_finish_lattice)
    if (accumulate == 2) {
      accumulate = 0;
    } else if (!accumulate) {
      cudaError_t e = cudaMemsetAsync(dx, 0, (size_t)N * H * W * Cin * 2, st);
      if (e != cudaSuccess) return -(int)e;
    }
    View4 d = parity_view(dx, N, H, W, Cin, 0, 0);
    return conv_gemm_launch(&a, 1, wt, ktot, Cin, d, taps, 1, Cout / 64, nullptr, nullptr, accumulate, st);
  }
  // 3x3 stride 2: one launch per parity class of the input pixel; taps r with (ph + 1 - r) even
  for (int ph = 0; ph < 2; ++ph)
    for (int pw = 0; pw < 2; ++pw) {
      int t = 0;
      for (int r = 0; r < 3; ++r) {
        if ((ph + 1 - r) & 1) continue;
        for (int s = 0; s < 3; ++s) {
          if ((pw + 1 - s) & 1) continue;
          taps.view[t] = 0;
          taps.dh[t] = (int8_t)((ph + 1 - r) / 2);
          taps.dw[t] = (int8_t)((pw + 1 - s) / 2);
          taps.koff[t] = (r * 3 + s) * Cout;
          ++t;
        }
      }
      View4 d = parity_view(dx, N, H, W, Cin, ph, pw);
      int rc = conv_gemm_launch(&a, 1, wt, ktot, Cin, d, taps, t, Cout / 64, s_sum, s_sq, accumulate, st,
                                bnr_for(d, &br));
      if (rc) return rc;
    }
  return 0;
}

extern "C" int peclr_conv2d_dgrad(const void* dy, const void* wt, void* dx, int N, int H, int W, int Cin, int Cout,
                                  int k, int stride, int accumulate, void* stream) {
  return dgrad_impl(dy, wt, dx, N, H, W, Cin, Cout, k, stride, accumulate, nullptr, nullptr, nullptr, nullptr,
                    nullptr, nullptr, static_cast<cudaStream_t>(stream));
}

// 1x1 / stride-1 dgrad that COMPLETES the gradient of a residual block's input (see GemmParams "finish mode"):
// dx holds the gradient gathered so far (shortcut branch, or the down-sampling branch's dgrad) and receives
// (dx + dgrad(dy)) * relu'(previous block's output), the mask coming from the bits bn_apply wrote; scratch[0:Cin] =
// sum g, scratch[Cin:2Cin] = sum g*y for the BatchNorm in front of that ReLU (bn_y = its input).
static int dgrad_finish_impl(const void* dy, const void* wt, void* dx, int N, int H, int W, int Cin, int Cout,
                             const void* bn_y, const void* mask_bits, double* scratch, int acc_stride, void* stream) {
  if (!dy || !wt || !dx || !bn_y || !mask_bits || !scratch) return PECLR_ERR_ARG;
  if (acc_stride != 1 && (acc_stride != 2 || (H & 1) || (W & 1) || (int64_t)N * H * W >= (1ll << 31)))
    return PECLR_ERR_ARG;
  if (Cin % 128 || Cout % 64 || Cin > 2048) return PECLR_ERR_ARG;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  cudaError_t e = cudaMemsetAsync(scratch, 0, (size_t)2 * Cin * sizeof(double), st);
  if (e != cudaSuccess) return -(int)e;
  TapTable taps;
  memset(&taps, 0, sizeof(taps));
  const int64_t M = (int64_t)N * H * W;
  View4 a = flat_view(dy, M, Cout);
  View4 d = flat_view(dx, M, Cin);
  BnReduce br;
  br.y = bn_y;
  br.mean = br.invstd = br.gamma = br.beta = nullptr;
  br.mask_bits = static_cast<const uint8_t*>(mask_bits);
  br.pix_base = 0, br.pix_w = 1, br.pix_h = 0, br.pix_n = 0;
  br.acc_stride = acc_stride, br.img_w = W;
  return conv_gemm_launch(&a, 1, wt, Cout, Cin, d, taps, 1, Cout / 64, scratch, scratch + Cin, 0, st, &br);
}

extern "C" int peclr_conv2d_dgrad_finish(const void* dy, const void* wt, void* dx, int N, int H, int W, int Cin,
                                         int Cout, const void* bn_y, const void* mask_bits, double* scratch,
                                         void* stream) {
  return dgrad_finish_impl(dy, wt, dx, N, H, W, Cin, Cout, bn_y, mask_bits, scratch, 1, stream);
}

// The same when the gradient gathered so far is the scatter of a stride-2 1x1 dgrad (the down-sampling shortcut of a
// stage's first block, peclr_conv2d_dgrad with accumulate = 2: no zero fill): only the pixels with even row and even
// column of dx are read, the rest counts as 0 -- the 100-400 MB memset of dx and its re-read disappear.
extern "C" int peclr_conv2d_dgrad_finish_lattice(const void* dy, const void* wt, void* dx, int N, int H, int W,
                                                 int Cin, int Cout, const void* bn_y, const void* mask_bits,
                                                 double* scratch, int acc_stride, void* stream) {
  return dgrad_finish_impl(dy, wt, dx, N, H, W, Cin, Cout, bn_y, mask_bits, scratch, acc_stride, stream);
}

extern "C" int peclr_conv2d_dgrad_bnreduce(const void* dy, const void* wt, void* dx, int N, int H, int W, int Cin,
                                           int Cout, int k, int stride, const void* bn_y, const float* bn_mean,
                                           const float* bn_invstd, const float* bn_gamma, const float* bn_beta,
                                           double* scratch, void* stream) {
  if (!bn_y || !bn_mean || !bn_invstd || !bn_gamma || !bn_beta) return PECLR_ERR_ARG;
  return dgrad_impl(dy, wt, dx, N, H, W, Cin, Cout, k, stride, 0, bn_y, bn_mean, bn_invstd, bn_gamma, bn_beta, scratch,
                    static_cast<cudaStream_t>(stream));
}

// views + tap table of a weight-gradient launch (shared by the workspace query and the launch)
static int wgrad_geometry(const void* x, const void* dy, int N, int H, int W, int Cin, int Cout, int k, int stride,
                          View4* views, TapTable* taps, int* num_taps, View4* d) {
  if (Cin % 64 || Cout % 64) return PECLR_ERR_ARG;
  const int Ho = H / stride, Wo = W / stride;
  if (k == 1 && stride == 1) {
    memset(taps, 0, sizeof(*taps));
    views[0] = flat_view(x, (int64_t)N * H * W, Cin);
    *d = flat_view(dy, (int64_t)N * H * W, Cout);
    *num_taps = 1;
    return 1;
  }
  const int nv = input_side(x, N, H, W, Cin, k, stride, Cin, views, taps, num_taps);
  *d = nhwc_view(dy, N, Ho, Wo, Cout);
  return nv;
}

extern "C" long long peclr_conv2d_wgrad_workspace_bytes(int N, int H, int W, int Cin, int Cout, int k, int stride) {
  TapTable taps;
  int num_taps = 0;
  View4 views[kMaxViews], d;
  const int nv = wgrad_geometry(nullptr, nullptr, N, H, W, Cin, Cout, k, stride, views, &taps, &num_taps, &d);
  if (nv < 0) return nv;
  return conv_wgrad_workspace_bytes(d, num_taps, Cin, Cout);
}

extern "C" int peclr_conv2d_wgrad(const void* x, const void* dy, float* dw, int N, int H, int W, int Cin, int Cout,
                                  int k, int stride, void* workspace, long long workspace_bytes, void* stream) {
  TapTable taps;
  int num_taps = 0;
  View4 views[kMaxViews], d;
  const int nv = wgrad_geometry(x, dy, N, H, W, Cin, Cout, k, stride, views, &taps, &num_taps, &d);
  if (nv < 0) return nv;
  return conv_wgrad_launch(views, nv, d, taps, num_taps, Cin, Cout, dw, workspace, workspace_bytes,
                           static_cast<cudaStream_t>(stream));
}

extern "C" int peclr_conv2d_wgrad_splits(int N, int H, int W, int Cin, int Cout, int k, int stride) {
  TapTable taps;
  int num_taps = 0;
  View4 views[kMaxViews], d;
  const int nv = wgrad_geometry(nullptr, nullptr, N, H, W, Cin, Cout, k, stride, views, &taps, &num_taps, &d);
  if (nv < 0) return nv;
  return conv_wgrad_splits(d, num_taps, Cin, Cout);
}

extern "C" int peclr_conv2d_wgrad_partials(const void* x, const void* dy, float* dw, int N, int H, int W, int Cin,
                                           int Cout, int k, int stride, void* workspace, long long workspace_bytes,
                                           void* stream) {
  TapTable taps;
  int num_taps = 0;
  View4 views[kMaxViews], d;
  const int nv = wgrad_geometry(x, dy, N, H, W, Cin, Cout, k, stride, views, &taps, &num_taps, &d);
  if (nv < 0) return nv;
  return conv_wgrad_launch(views, nv, d, taps, num_taps, Cin, Cout, dw, workspace, workspace_bytes,
                           static_cast<cudaStream_t>(stream), true);
}

extern "C" int peclr_wgrad_reduce_block_f4(void) { return wgrad_reduce_f4_per_block(); }

extern "C" int peclr_wgrad_reduce_batched(const void* table, int num_entries, int total_blocks, void* stream) {
  return wgrad_reduce_batched_launch(table, num_entries, total_blocks, static_cast<cudaStream_t>(stream));
}

// ---- 7x7 / stride 2 / pad 3 stem as a 4x4 / stride 1 convolution on the space-to-depth image ----------------
// xs: [N][H/2 + 3][W/2 + 4][16] bf16 from peclr_stem_input (2 x 2 pixel blocks as 16-channel pixels, block (Y, X)
// at (Y + 2, X + 2)).  out(oy, ox) = sum_{ty, tx < 4} w4[ty][tx] . xs(oy + ty, ox + tx): for output (oy, ox) and
// filter row ty the 4-block window starting at padded block (oy + ty, ox) is 64 contiguous bf16 -- that window is
// the "channel" dimension of an ordinary K-major GEMM operand (a TMA view whose pixel stride, 16 elements, is
// smaller than its 64-element rows).  Weights: [64][4][4 * 16] bf16 from peclr_stem_pack, K = 256 (147 non-zero).
static void stem_views(const void* xs, int N, int H, int W, View4* views, TapTable* taps) {
  const int64_t Hs = H / 2 + 3, Ws = W / 2 + 4;
  memset(taps, 0, sizeof(*taps));
  views[0] = View4{xs, 64, W / 2, Hs, N, 16, Ws * 16, Hs * Ws * 16};
  for (int r = 0; r < 4; ++r) {
    taps->view[r] = 0;
    taps->dh[r] = (int8_t)r;
    taps->dw[r] = 0;
    taps->koff[r] = r * 64;
  }
}

extern "C" int peclr_stem_fprop(const void* xpad, const void* wpack, void* y, int N, int H, int W, double* stat_sum,
                                double* stat_sumsq, void* stream) {
  if ((H & 1) || (W & 1)) return PECLR_ERR_ARG;
  View4 views[kMaxViews];
  TapTable taps;
  stem_views(xpad, N, H, W, views, &taps);
  View4 d = View4{y, 64, W / 2, H / 2, N, 64, (int64_t)(W / 2) * 64, (int64_t)(H / 2) * (W / 2) * 64};
  if (halo_ok(64, (int64_t)N * (H / 2) * (W / 2))) {  // the four filter rows as one tap group
    taps.dh[0] = 0, taps.dw[0] = 0, taps.view[0] = 0, taps.koff[0] = 0;
    return conv_gemm_launch(views, 1, wpack, 4 * 64, 64, d, taps, 1, 1, stat_sum, stat_sumsq, 0,
                            static_cast<cudaStream_t>(stream), nullptr, 4, 64);
  }
  return conv_gemm_launch(views, 1, wpack, 4 * 64, 64, d, taps, 4, 1, stat_sum, stat_sumsq, 0,
                          static_cast<cudaStream_t>(stream));
}

extern "C" long long peclr_stem_wgrad_workspace_bytes(int N, int H, int W) {
  if ((H & 1) || (W & 1)) return PECLR_ERR_ARG;
  View4 d = View4{nullptr, 64, W / 2, H / 2, N, 64, (int64_t)(W / 2) * 64, (int64_t)(H / 2) * (W / 2) * 64};
  return conv_wgrad_workspace_bytes(d, 4, 64, 64);
}

extern "C" int peclr_stem_wgrad(const void* xpad, const void* dy, float* dwpack, int N, int H, int W, void* workspace,
                                long long workspace_bytes, void* stream) {
  if ((H & 1) || (W & 1)) return PECLR_ERR_ARG;
  View4 views[kMaxViews];
  TapTable taps;
  stem_views(xpad, N, H, W, views, &taps);
  View4 d = View4{dy, 64, W / 2, H / 2, N, 64, (int64_t)(W / 2) * 64, (int64_t)(H / 2) * (W / 2) * 64};
  return conv_wgrad_launch(views, 1, d, taps, 4, 64, 64, dwpack, workspace, workspace_bytes,
                           static_cast<cudaStream_t>(stream));
}
