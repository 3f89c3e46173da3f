// Head of the downstream 2.5D hand-pose network the exported PeCLR encoders feed (reference:
// src/models/rn_25D_wMLPref.py:75-134 RN_25D_wMLPref.forward after the backbone, and :6-72 ZrootMLP_ref), inference
// (eval-mode BatchNorm1d): from the 64 backbone outputs per image to kp25d / kp2d / zrel / kp3d in ONE launch.
//
//   kp25d = out[:, :63].view(21, 3), zrel = kp25d[..., 2] with zrel[0] = 0 (written through: kp25d shares it)
//   kp3d_unnorm = [x, y, 1] K^-1^T
//   zroot from the scale-normalised bone (3, 8): a z^2 + b z + c = 0 (Iqbal et al. 2018, eq. 6-7), clamped to [4, 50]
//   zroot += MLP(zrel (21), kp3d_unnorm xy (42), zroot (1)):  Linear(64,128) BN LeakyReLU Linear(128,128) BN LeakyReLU
//            Linear(128,1)   (Spurr et al. 2020)
//   kp3d = kp3d_unnorm * (zrel + zroot)
//
// One 128-thread block per image (the op is ~50 kFLOP per image: latency-bound, the MLP weights stay in L2).
#include "../../include/peclr_b200.h"
#include "ptx.cuh"

namespace peclr {

struct Rn25dHeadArgs {
  const float* out;   // [B][64] backbone output (fc)
  const float* K;     // [nK][3][3], nK = 1 (shared) or B
  int nK, B;
  const float *w1, *b1, *g1, *be1, *rm1, *rv1;  // Linear(64,128) + BatchNorm1d(128) (eval: running statistics)
  const float *w2, *b2, *g2, *be2, *rm2, *rv2;  // Linear(128,128) + BatchNorm1d(128)
  const float *w3, *b3;                         // Linear(128,1)
  float bn_eps, slope;
  float *kp3d, *zrel, *kp2d, *kp25d;            // [B][21][3], [B][21][1], [B][21][2], [B][21][3]
};

__global__ void __launch_bounds__(128) rn25d_head_kernel(const Rn25dHeadArgs a) {
  __shared__ float in[64], h1[128], h2[128], un[21][3], zr[21], kinv[9], part[4], zroot_s;
  const int t = threadIdx.x, n = blockIdx.x;
  const float* o = a.out + (size_t)n * 64;
  if (t == 0) {  // K^-1 by the adjugate (K is a camera matrix: well conditioned)
    const float* k = a.K + (size_t)(a.nK == 1 ? 0 : n) * 9;
    const float c00 = k[4] * k[8] - k[5] * k[7], c01 = k[5] * k[6] - k[3] * k[8], c02 = k[3] * k[7] - k[4] * k[6];
    const float inv_det = 1.f / (k[0] * c00 + k[1] * c01 + k[2] * c02);
    kinv[0] = c00 * inv_det, kinv[1] = (k[2] * k[7] - k[1] * k[8]) * inv_det, kinv[2] = (k[1] * k[5] - k[2] * k[4]) * inv_det;
    kinv[3] = c01 * inv_det, kinv[4] = (k[0] * k[8] - k[2] * k[6]) * inv_det, kinv[5] = (k[2] * k[3] - k[0] * k[5]) * inv_det;
    kinv[6] = c02 * inv_det, kinv[7] = (k[1] * k[6] - k[0] * k[7]) * inv_det, kinv[8] = (k[0] * k[4] - k[1] * k[3]) * inv_det;
  }
  __syncthreads();
  if (t < 21) {
    const float x = o[3 * t], y = o[3 * t + 1], z = t == 0 ? 0.f : o[3 * t + 2];  // zrel of the root is 0
    un[t][0] = kinv[0] * x + kinv[1] * y + kinv[2];
    un[t][1] = kinv[3] * x + kinv[4] * y + kinv[5];
    un[t][2] = kinv[6] * x + kinv[7] * y + kinv[8];
    zr[t] = z;
    const size_t q = (size_t)n * 21 + t;
    a.kp2d[q * 2] = x, a.kp2d[q * 2 + 1] = y;
    a.kp25d[q * 3] = x, a.kp25d[q * 3 + 1] = y, a.kp25d[q * 3 + 2] = z;
    a.zrel[q] = z;
  }
  __syncthreads();
  if (t == 0) {
    const float Xm = un[3][0], Ym = un[3][1], Xn = un[8][0], Yn = un[8][1], zm = zr[3], zn = zr[8];
    float qa = (Xn - Xm) * (Xn - Xm) + (Yn - Ym) * (Yn - Ym);
    const float qb = 2.f * (zn * (Xn * Xn + Yn * Yn - Xn * Xm - Yn * Ym) + zm * (Xm * Xm + Ym * Ym - Xn * Xm - Yn * Ym));
    const float qc = (Xn * zn - Xm * zm) * (Xn * zn - Xm * zm) + (Yn * zn - Ym * zm) * (Yn * zn - Ym * zm) +
                     (zn - zm) * (zn - zm) - 1.f;
    float d = qb * qb - 4.f * qa * qc;
    qa = fmaxf(1e-8f, qa);
    d = fmaxf(1e-8f, d);
    zroot_s = fminf(fmaxf((-qb + sqrtf(d)) / (2.f * qa), 4.f), 50.f);
  }
  __syncthreads();
  if (t < 21) in[t] = zr[t];
  else if (t < 63) in[t] = un[(t - 21) >> 1][(t - 21) & 1];
  else if (t == 63) in[63] = zroot_s;
  __syncthreads();
  {
    float acc = a.b1[t];
    const float* w = a.w1 + t * 64;
#pragma unroll 8
    for (int k = 0; k < 64; ++k) acc = fmaf(w[k], in[k], acc);
    acc = (acc - a.rm1[t]) * rsqrtf(a.rv1[t] + a.bn_eps) * a.g1[t] + a.be1[t];
    h1[t] = acc > 0.f ? acc : acc * a.slope;
  }
  __syncthreads();
  {
    float acc = a.b2[t];
    const float* w = a.w2 + t * 128;
#pragma unroll 8
    for (int k = 0; k < 128; ++k) acc = fmaf(w[k], h1[k], acc);
    acc = (acc - a.rm2[t]) * rsqrtf(a.rv2[t] + a.bn_eps) * a.g2[t] + a.be2[t];
    h2[t] = acc > 0.f ? acc : acc * a.slope;
  }
  __syncthreads();
  float v = warp_sum(a.w3[t] * h2[t]);
  if ((t & 31) == 0) part[t >> 5] = v;
  __syncthreads();
  if (t < 21) {
    const float zroot = zroot_s + (part[0] + part[1] + part[2] + part[3] + a.b3[0]);
    const float s = zr[t] + zroot;
    const size_t q = ((size_t)n * 21 + t) * 3;
    a.kp3d[q] = un[t][0] * s, a.kp3d[q + 1] = un[t][1] * s, a.kp3d[q + 2] = un[t][2] * s;
  }
}

}  // namespace peclr

using namespace peclr;

extern "C" int peclr_rn25d_head(const float* out, const float* K, int nK, int B, const float* const* mlp,
                                float bn_eps, float leaky_slope, float* kp3d, float* zrel, float* kp2d, float* kp25d,
                                void* stream) {
  if (!out || !K || !mlp || B < 1 || (nK != 1 && nK != B) || !kp3d || !zrel || !kp2d || !kp25d) return -1001;
  for (int i = 0; i < 14; ++i)
    if (!mlp[i]) return -1001;
  Rn25dHeadArgs a;
  a.out = out, a.K = K, a.nK = nK, a.B = B;
  a.w1 = mlp[0], a.b1 = mlp[1], a.g1 = mlp[2], a.be1 = mlp[3], a.rm1 = mlp[4], a.rv1 = mlp[5];
  a.w2 = mlp[6], a.b2 = mlp[7], a.g2 = mlp[8], a.be2 = mlp[9], a.rm2 = mlp[10], a.rv2 = mlp[11];
  a.w3 = mlp[12], a.b3 = mlp[13];
  a.bn_eps = bn_eps, a.slope = leaky_slope;
  a.kp3d = kp3d, a.zrel = zrel, a.kp2d = kp2d, a.kp25d = kp25d;
  rn25d_head_kernel<<<B, 128, 0, static_cast<cudaStream_t>(stream)>>>(a);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : -(int)e;
}
