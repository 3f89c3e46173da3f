/* peclr_b200 -- C ABI of the B200-native PeCLR pre-training step (libpeclr_b200.so).
 *
 * The reference (dahiyaaneesh/peclr) is pure Python/PyTorch and has no FFI of its own; its boundary
 * for this path is the Python operator/module API.  These entry points are what
 * that API binds underneath in this build: each one names the reference code it replaces.  The
 * ctypes stub a reference maintainer would add is shown in INTEGRATION.md.
 *
 * Conventions: every pointer is a DEVICE pointer unless stated; bf16 buffers are passed as void*;
 * `stream` is a cudaStream_t passed as void* (use torch.cuda.current_stream().cuda_stream); calls are
 * asynchronous on that stream, never allocate or free, and return 0 on success, a negative code on
 * error (-1001 bad argument, -1002 driver entry point missing, -1003 tensor-map encoding failed,
 * -N = cudaError_t N).  Activations are NHWC bf16; convolution weights are bf16 [Cout][kh*kw][Cin]
 * ("KRSC" = torch.channels_last memory of a (Cout,Cin,kh,kw) tensor); weight gradients are fp32 in
 * the same layout.  Compiled for sm_100a only; there is no CPU or other-arch fallback.
 */
#ifndef PECLR_B200_H_
#define PECLR_B200_H_

#ifdef __cplusplus
extern "C" {
#endif

/* 3 = this header (1: fp32 statistic sums, NHWC4 stem input; 2: fp64 forward sums, fp32 accumulator-set scratch,
 * atomically reduced weight-gradient splits).
 *
 * Reproducibility contract of version 3: no result depends on the order in which thread blocks run.  Per-channel
 * sums (forward BatchNorm statistics, BN-backward sums, optimiser norms, loss-chain row sums) are fp64 accumulators
 * receiving one partial per block, each partial computed in a fixed order; split-K partial products (weight
 * gradients, head GEMMs, the loss chain's gradient) go through workspace slabs that are added in slab order.  Two
 * runs on the same inputs give bit-identical loss, gradients and updated weights. */
int peclr_abi_version(void);

/* ---- ResNet trunk convolutions: tcgen05/TMEM implicit GEMM fed by TMA (csrc/conv_tc.cu) ------------
 * Replace nn.Conv2d forward/backward inside ResNetModel.features
 * (reference src/models/resnet_model.py:16-26,45-52; the 23 shapes of SURVEY.md table A2).
 * k in {1,3}, stride in {1,2}, padding (k-1)/2, Cin and Cout multiples of 64, H and W even if stride 2.
 * H, W are always the INPUT spatial size. */

/* y[N,H/s,W/s,Cout] = conv(x[N,H,W,Cin], w[Cout][k*k][Cin]); if stat_sum != NULL also accumulates
 * per-output-channel sum and sum of squares of the stored bf16 y (training-mode BatchNorm statistics;
 * the two double[Cout] buffers must be zeroed by the caller).  The forward statistics are fp64 accumulators: each
 * CTA adds a partial computed in a fixed order, so the fp32 mean / variance derived from the totals do not depend on
 * the order of arrival and the forward pass (hence the loss) is reproducible from run to run. */
int peclr_conv2d_fprop(const void* x, const void* w, void* y, int N, int H, int W, int Cin, int Cout, int k,
                       int stride, double* stat_sum, double* stat_sumsq, void* stream);
/* dx[N,H,W,Cin] (+)= conv_transpose(dy[N,H/s,W/s,Cout], wt[Cin][k*k][Cout]); wt is the transposed
 * weight copy produced by peclr_weight_transpose.  accumulate = 1 adds into dx (TMA reduce-add); accumulate = 2
 * (k = 1, stride = 2 only): scatter to the sampled pixels without zero-filling the others -- for
 * peclr_conv2d_dgrad_finish_lattice, which never reads them as data. */
int peclr_conv2d_dgrad(const void* dy, const void* wt, void* dx, int N, int H, int W, int Cin, int Cout, int k,
                       int stride, int accumulate, void* stream);
/* The same dgrad with the BatchNorm-backward reduction of the BN + ReLU sitting in front of this convolution fused
 * into the epilogue: bn_y = that BatchNorm's input (shape of dx), bn_* its saved mean / invstd and affine
 * parameters; scratch (double[2*Cin], zeroed by the call) receives scratch[0:Cin] = sum g, scratch[Cin:2Cin] =
 * sum g*y with g = dx * relu'  (what peclr_bn_bwd_reduce with mask_mode 2 computes in a separate pass).  Follow
 * with peclr_bn_bwd_apply(mask_mode 2). */
int peclr_conv2d_dgrad_bnreduce(const void* dy, const void* wt, void* dx, int N, int H, int W, int Cin, int Cout,
                                int k, int stride, const void* bn_y, const float* bn_mean, const float* bn_invstd,
                                const float* bn_gamma, const float* bn_beta, double* scratch, void* stream);
/* The 1x1 / stride-1 dgrad that COMPLETES the gradient of a residual block's input, i.e. of the previous block's
 * output out = relu(bn(y) + shortcut) (torchvision Bottleneck.forward, "out += identity; out = self.relu(out)"):
 * on entry dx holds the part of that gradient gathered so far (shortcut branch, or the down-sampling branch's dgrad),
 * on exit dx = (dx + dgrad(dy)) * relu'(out) with the ReLU mask taken from the bits peclr_bn_apply wrote
 * (mask_bits, uint8 [M][Cin/8]), and scratch (double[2*Cin], zeroed by the call) = {sum g, sum g*y} with bn_y = y.
 * Replaces TMA reduce-add accumulation + peclr_bn_bwd_reduce(mask_mode 3) + the masking inside peclr_bn_bwd_apply
 * for that block output; follow with peclr_bn_bwd_apply(mask_mode 0) on dx.  Cin multiple of 128. */
int peclr_conv2d_dgrad_finish(const void* dy, const void* wt, void* dx, int N, int H, int W, int Cin, int Cout,
                              const void* bn_y, const void* mask_bits, double* scratch, void* stream);
/* The same with acc_stride = 2 for the first block of a stage, where the gradient gathered so far is the down-sampling
 * shortcut's (torchvision Bottleneck.downsample: conv1x1 stride 2): written by peclr_conv2d_dgrad(k = 1, stride = 2,
 * accumulate = 2) = scatter to the pixels with even row and even column WITHOUT zero-filling the rest, which this call
 * then reads as 0 (never as data).  acc_stride = 1 is peclr_conv2d_dgrad_finish.  H, W even. */
int peclr_conv2d_dgrad_finish_lattice(const void* dy, const void* wt, void* dx, int N, int H, int W, int Cin, int Cout,
                                      const void* bn_y, const void* mask_bits, double* scratch, int acc_stride,
                                      void* stream);
/* dw[Cout][k*k][Cin] (fp32) += dy^T * im2col(x).  The pixel dimension is split over thread blocks; the splits'
 * partial products go to `workspace` (peclr_conv2d_wgrad_workspace_bytes for the same geometry; may be NULL when
 * that is 0) and a second kernel adds them to dw in a fixed order. */
long long peclr_conv2d_wgrad_workspace_bytes(int N, int H, int W, int Cin, int Cout, int k, int stride);
int peclr_conv2d_wgrad(const void* x, const void* dy, float* dw, int N, int H, int W, int Cin, int Cout, int k,
                       int stride, void* workspace, long long workspace_bytes, void* stream);
/* The weight gradients of MANY convolutions with ONE ordered reduction (a ResNet stage's worth: the per-convolution
 * reduction launches are mostly fixed cost): peclr_conv2d_wgrad_partials only writes the pixel splits' slabs
 * (ksplit = peclr_conv2d_wgrad_splits(...) of them, each Cout*k*k*Cin floats, into its own workspace region; with
 * ksplit == 1 it accumulates into dw directly and needs no row), then peclr_wgrad_reduce_batched adds, per table
 * row {const float* partial; float* dw; int64 n4 (= elements / 4); int32 ksplit; int32 first_block}, the slabs to
 * dw in slab order.  A row owns ceil(n4 / peclr_wgrad_reduce_block_f4()) consecutive blocks from first_block on;
 * total_blocks is their sum.  `table` is a device array. */
int peclr_conv2d_wgrad_splits(int N, int H, int W, int Cin, int Cout, int k, int stride);
int peclr_conv2d_wgrad_partials(const void* x, const void* dy, float* dw, int N, int H, int W, int Cin, int Cout,
                                int k, int stride, void* workspace, long long workspace_bytes, void* stream);
int peclr_wgrad_reduce_block_f4(void);
int peclr_wgrad_reduce_batched(const void* table, int num_entries, int total_blocks, void* stream);
/* 7x7/stride 2/pad 3 stem (features.0), computed as a 4x4/stride 1 convolution over 2x2 pixel blocks.
 * xpad = [N][H/2+3][W/2+4][16] bf16 from peclr_stem_input (space-to-depth, zero padded),
 * wpack = [64][4][4*16] bf16 from peclr_stem_pack, y = [N][H/2][W/2][64]. */
int peclr_stem_fprop(const void* xpad, const void* wpack, void* y, int N, int H, int W, double* stat_sum,
                     double* stat_sumsq, void* stream);
/* dwpack[64][4][64] fp32 += ... ; fold into the (64,3,7,7) gradient with peclr_stem_unpack_grad */
long long peclr_stem_wgrad_workspace_bytes(int N, int H, int W);
int peclr_stem_wgrad(const void* xpad, const void* dy, float* dwpack, int N, int H, int W, void* workspace,
                     long long workspace_bytes, void* stream);

/* ---- HBM-bound trunk kernels (csrc/bn_act.cu) --------------------------------------------------------
 * Replace nn.BatchNorm2d(train) + ReLU + residual add, MaxPool2d(3,2,1), AdaptiveAvgPool2d(1) of the
 * torchvision blocks the reference wraps (resnet_model.py:16-26). */

/* out = [relu]( bn(y) + residual ), bn from the conv epilogue's sums over M rows.  residual is NULL, a
 * finished activation (rsum == NULL) or a raw conv output with its own BatchNorm (downsample branch).
 * Saves mean / invstd for backward and updates running stats (momentum, unbiased var) as PyTorch does.
 * mask_out (optional, uint8 [M][C/8]) receives the ReLU mask as bits for peclr_bn_bwd_* mask_mode 3. */
int peclr_bn_apply(const void* y, const double* sum, const double* sumsq, const float* gamma, const float* beta,
                   const void* res, const double* rsum, const double* rsumsq, const float* rgamma, const float* rbeta,
                   void* out, void* mask_out, float* mean_out, float* invstd_out, float* running_mean,
                   float* running_var,
                   float* rmean_out, float* rinvstd_out, float* rrunning_mean, float* rrunning_var, long long M, int C,
                   float eps, float momentum, int relu, void* stream);
/* scratch (double[2C], zeroed by the call): scratch[0:C] = sum g, scratch[C:2C] = sum g*y with g = dout * relu';
 * the ReLU mask is (mask_mode)
 * 0: none (dout already masked), 1: the stored activation `mask` > 0, 2: recomputed from y, gamma, beta,
 * 3: `mask` is the bit mask written by peclr_bn_apply. */
int peclr_bn_bwd_reduce(const void* dout, const void* mask, const void* y, const float* mean, const float* invstd,
                        const float* gamma, const float* beta, int mask_mode, double* scratch, long long M, int C,
                        void* stream);
/* dy = gamma*invstd*(g - mean g - xhat*mean(g xhat)); optional g_out = g; dbeta += sum g, dgamma += sum g*xhat */
int peclr_bn_bwd_apply(const void* dout, const void* mask, const void* y, const float* mean, const float* invstd,
                       const float* gamma, const float* beta, int mask_mode, const double* scratch, void* dy,
                       void* g_out, float* dgamma, float* dbeta, long long M, int C, void* stream);
/* out[N,H/2,W/2,64] = maxpool3x3s2p1(relu(bn(y[N,H,W,64])))   (features.1-3); idx_out (uint8, same shape as
 * out, may be NULL) records the winning window position 0..8 for the backward pass */
int peclr_stem_bn_relu_pool(const void* y, const double* sum, const double* sumsq, const float* gamma,
                            const float* beta, void* out, void* idx_out, float* mean_out, float* invstd_out,
                            float* running_mean, float* running_var, int N, int H, int W, float eps, float momentum,
                            void* stream);
/* g_out[N,H,W,64] = relu'(.) * maxpool_backward(dpool) and the BN-backward sums (sum g, sum g*y) into
 * scratch (double[128]) */
int peclr_stem_pool_bwd(const void* dpool, const void* idx, const void* y, const float* mean, const float* invstd,
                        const float* gamma, const float* beta, void* g_out, double* scratch, int N, int H, int W,
                        void* stream);
int peclr_avgpool_fwd(const void* x, float* out, int N, int HW, int C, void* stream);
int peclr_avgpool_bwd(const float* dout, void* dx, int N, int HW, int C, void* stream);
/* cat(transformed_image1, transformed_image2) (hybrid2_model.py:30-32), fp32 NCHW -> zero-padded bf16
 * space-to-depth batch [2B][H/2+3][W/2+4][16] (channel = dy*6 + dx*3 + c of the 2x2 block; 12..15 zero).
 * x2 == NULL: x1 alone, [B][...] (any batch size: inference, a short last validation batch). */
int peclr_stem_input(const float* x1, const float* x2, void* out, int B, int H, int W, void* stream);

/* ---- projection head, fp32 (csrc/head.cu); replaces SimCLR.get_projection_head modules
 * (reference src/models/unsupervised/simclr_model.py:20-35) */
/* C[m,n] (+)= sum_k A[m*sam + k*sak] * B[k*sbk + n*sbn] (+ bias[n]).  Small GEMMs split K over thread blocks: the
 * partial tiles go through `workspace` (peclr_sgemm_workspace_bytes; zeroed ONCE by the caller before first use,
 * it may be shared by GEMMs of different shapes on one stream) and are added in split order by the last block of
 * each tile.  workspace == NULL: no split (same result up to the summation order, fewer blocks in flight). */
long long peclr_sgemm_workspace_bytes(int M, int N, int K);
int peclr_sgemm(const float* A, const float* B, float* C, const float* bias, int M, int N, int K, long long sam,
                long long sak, long long sbk, long long sbn, long long ldc, int accumulate, void* workspace,
                long long workspace_bytes, void* stream);
int peclr_bn1d_relu_fwd(const float* x, const float* gamma, const float* beta, float* out, float* mean_out,
                        float* invstd_out, float* running_mean, float* running_var, int M, int C, float eps,
                        float momentum, void* stream);
int peclr_bn1d_relu_bwd(const float* dout, const float* out, const float* x, const float* mean, const float* invstd,
                        const float* gamma, float* dx, float* dgamma, float* dbeta, int M, int C, void* stream);
int peclr_colsum_acc(const float* x, float* out, int M, int N, void* stream);

/* ---- fused equivariance correction + NT-Xent, forward and backward in one launch (csrc/ntxent.cu) ----
 * Replaces Hybrid2Model.get_transformed_projections after the head + contrastive_step
 * (hybrid2_model.py:40-90), translate_encodings / rotate_encoding / get_rotation_2D_matrix /
 * vanila_contrastive_loss (src/models/utils.py:154-186,271-346) and get_projection_stats (:92-106).
 * p [2B][128] fp32 (view-1 rows then view-2 rows); angle [2B] f64 degrees; jx, jy [2B] int64 pixels;
 * loss [1]; stats [16] = proj1 {x_mean,x_median,x_min,x_max,y_mean,y_median,y_min,y_max}, proj2 {...};
 * g_p [2B][128] = dloss/dp (NULL: forward only).  world > 1: the embedding all-gather is fused into
 * the kernel -- z_peers[r] / flag_peers[r] are rank r's z buffer (the start of its workspace) and flag
 * array (unsigned[world]) mapped into this process (symmetric / peer memory).  The workspace must be zeroed once
 * before first use; the kernel keeps its launch counter there (flags are tagged with it and the gathered
 * embeddings are double buffered by its parity), so the call can be captured in a CUDA graph and replayed.  The loss is the GLOBAL-batch mean on every rank and g_p its gradient w.r.t. the local
 * rows, so parameter gradients are SUMMED across ranks. */
long long peclr_ntxent_workspace_bytes(int B, int world);
int peclr_ntxent_fused(const float* p, const double* angle, const long long* jx, const long long* jy, int B, int dim,
                       int img_h, int img_w, int crop, int rotate, float temperature, float* loss, float* stats,
                       float* g_p, void* workspace, long long workspace_bytes, int world, int rank,
                       float* const* z_peers, unsigned* const* flag_peers, void* stream);

/* NT-Xent alone on already-normalised embeddings z [2B][128] (view-1 rows then view-2 rows): the same kernel
 * without the normalisation / correction phases.  Replaces vanila_contrastive_loss (src/models/utils.py:154-186)
 * as used by SimCLR.contrastive_step (simclr_model.py:37-49).  g_z = dloss/dz (NULL: forward only). */
int peclr_ntxent_plain(const float* z, int B, int dim, float temperature, float* loss, float* g_z, void* workspace,
                       long long workspace_bytes, void* stream);

/* ---- the equivariance corrections as stand-alone operators (csrc/equiv_ops.cu), fp32 ------------------
 * The reference exposes them as free functions that model variants and user code call directly; the training
 * step itself uses the fused kernel above.  enc is [n][m][d] contiguous, d >= 2; only coordinates 0 and 1 of
 * every point are touched; all of them work in place. */
/* translate_encodings (src/models/utils.py:325-346): x += tx*(max_x - min_x), y += ty*(max_y - min_y) with the
 * per-sample range taken over the m points (detached: the gradient is the identity); exact != 0 gives
 * translate_encodings2 (:349-364): x += tx, y += ty. */
int peclr_translate_encodings(float* enc, const float* tx, const float* ty, int n, int m, int d, int exact,
                              void* stream);
/* rotate_encoding (src/models/utils.py:301-321): rotation by angle[i] degrees (f64) about the sample's (detached)
 * mean point with the OpenCV-convention matrix of get_rotation_2D_matrix; rot (optional, float[n][4]) receives
 * {alpha, beta, off_x, off_y} for the backward pass. */
int peclr_rotate_encoding(float* enc, const double* angle, float* rot, int n, int m, int d, void* stream);
/* g[..., :2] <- g[..., :2] @ R[:2,:2]^T, in place (the gradient of rotate_encoding w.r.t. its input) */
int peclr_rotate_encoding_bwd(float* g, const float* rot, int n, int m, int d, void* stream);
/* get_rotation_2D_matrix (src/models/utils.py:271-298): out [n][3][2] fp32 from f64 angles (degrees), fp32
 * centres and a scalar scale; trig and offsets evaluated in f64 and rounded, as the reference does. */
int peclr_rotation_2d_matrix(const double* angle, const float* center_x, const float* center_y, double scale,
                             float* out, int n, void* stream);

/* Hybrid2Model.get_projection_stats (hybrid2_model.py:92-106): out8 = batch means of the per-sample
 * x{mean, lower median, min, max}, y{...} over the m points of enc [n][m][d]. */
int peclr_projection_stats(const float* enc, float* out8, int n, int m, int d, void* stream);

/* ---- fused LARS-Adam over the flat parameter buffer (csrc/lars_adam.cu); replaces
 * LARSWrapper(torch.optim.Adam).step() as configured by BaseModel.configure_optimizers
 * (src/models/base_model.py:57-104).  seg_begin[num_segs+1] are tensor boundaries, seg_wd the weight
 * decay of each tensor's param group (exclude_from_wt_decay, base_model.py:30-51); chunk_* partition
 * the buffer into blocks of peclr_opt_chunk_elems() that never straddle a tensor; norms is double[2*num_segs]
 * scratch.  step is the 1-based Adam step.  p_bf16 (optional) receives the updated weights in bf16. */
int peclr_opt_chunk_elems(void);
int peclr_lars_adam_step(float* p, const float* g, float* m, float* v, void* p_bf16, const long long* seg_begin,
                         const float* seg_wd, int num_segs, const int* chunk_seg, const long long* chunk_begin,
                         int num_chunks, double* norms, float lr, int step, float beta1, float beta2, float adam_eps,
                         int lars, float eta, int clip, float lars_eps, void* stream);
int peclr_cast_bf16(const float* src, void* dst, long long n, void* stream);
/* table: array of {int64 src_off, int64 dst_off, int32 cout, taps, cin, tile_begin}; one launch
 * produces every convolution's [Cin][taps][Cout] bf16 dgrad operand from the fp32 master weights */
int peclr_weight_transpose(const float* src_flat, void* dst_bf16, const void* table, int num_entries, int total_tiles,
                           void* stream);
int peclr_stem_pack(const float* w, void* wpack, void* stream);
int peclr_stem_unpack_grad(const float* gpack, float* g, void* stream);

/* ---- GPU-side two-view augmentation (csrc/augment.cu): the pixel work of SampleAugmenter.transform_sample
 * (reference src/data_loader/sample_augmenter.py:47-129: rotate = cv2.warpAffine, crop, cv2.resize INTER_AREA, HSV
 * colour jitter) + ToTensor / Normalize (data_loader/utils.py:287-293), one launch for all N = 2B view images.
 * src_u8: device buffer of raw 8-bit H x W x 3 images; view_table: N rows of
 *   { double m[6] (dst->src affine map, inverted as cv2.warpAffine does); int64 src_off (bytes); int32 sh, sw (source
 *     size); int32 ox, oy, cw, ch (crop box in the rotated image, clipped); double h, s, a, b (HSV factors);
 *     int32 rotate, jitter (flags) }   -- 120 bytes, peclr_b200/gpu_augment.py VIEW_DTYPE;
 * out: fp32 [N][3][out_h][out_w] normalised views; stage_u8 (optional, [N][out_h][out_w][3]) the 8-bit image before
 * ToTensor.  OpenCV's 8-bit arithmetic is reproduced (fixed-point bilinear warp, area resize, integer HSV). */
int peclr_two_view_augment(const void* src_u8, long long src_bytes, const void* view_table, int n, int out_h,
                           int out_w, float mean0, float mean1, float mean2, float std0, float std1, float std2,
                           float* out, void* stage_u8, void* stream);

/* ---- downstream consumer of the exported encoder (csrc/rn25d_head.cu): the head of RN_25D_wMLPref
 * (reference src/models/rn_25D_wMLPref.py:75-134 forward after the backbone, :6-72 ZrootMLP_ref), inference only
 * (eval-mode BatchNorm1d).  out [B][64] = backbone output (fc of the ResNet trunk: peclr_sgemm on the pooled
 * features); K [nK][3][3] camera matrices, nK = 1 (shared) or B.  `mlp` is a HOST array of 14 device pointers:
 * {W1[128][64], b1, bn1.weight, bn1.bias, bn1.running_mean, bn1.running_var, W2[128][128], b2, bn2.weight, bn2.bias,
 * bn2.running_mean, bn2.running_var, W3[1][128], b3[1]}.  Outputs: kp3d [B][21][3], zrel [B][21][1], kp2d [B][21][2],
 * kp25d [B][21][3] (zrel of the root joint zeroed, as the reference writes it through its views). */
int peclr_rn25d_head(const float* out, const float* K, int nK, int B, const float* const* mlp, float bn_eps,
                     float leaky_slope, float* kp3d, float* zrel, float* kp2d, float* kp25d, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PECLR_B200_H_ */
