// Fused SSIM forward / backward kernels (replace submodules/fused-ssim, ssim.cu:187-366) and
// their LOSS variants, which fold train.py's whole image loss into the same two passes
// (see loss.cu). Included by ssim.cu and loss.cu.
//
// Same math as the reference: separable 11-tap Gaussian (sigma 1.5, taps ssim.cu:9-19), zero
// padding, x pass then y pass, per-pixel SSIM map and the three partial maps dm/dmu1,
// dm/dsigma1^2, dm/dsigma12; backward is three convolutions of (dL * partial). The reference
// runs five (forward) / three (backward) separate conv rounds through one scratch tile with
// ~20 block barriers; here one 32x32 output tile per CTA does ONE horizontal pass producing
// all five (three) row-filtered quantities, then one vertical pass: 3 barriers per tile and
// each input pixel is read from HBM once. Both passes are register-blocked: a thread produces
// 4 consecutive outputs along the filter direction from a 14-value sliding window, which cuts
// the shared-memory loads per output from 11 to 3.5 (the kernels were LSU-bound before).
#pragma once
#include "common.cuh"

namespace cg {
namespace ssimk {

constexpr int TS = 32;         // output tile
constexpr int HALO = 5;
constexpr int IN = TS + 2 * HALO;   // 42
constexpr int NT = 256;
constexpr int RUN = 4;         // outputs per thread along the filter direction
constexpr int WIN = RUN + 10;  // 14-value window

static __constant__ float c_tap[11] = {0.001028380123898387f, 0.0075987582094967365f, 0.036000773310661316f,
                                       0.10936068743467331f,  0.21300552785396576f,   0.26601171493530273f,
                                       0.21300552785396576f,  0.10936068743467331f,   0.036000773310661316f,
                                       0.0075987582094967365f, 0.001028380123898387f};

__device__ __forceinline__ float load_px(const float* __restrict__ img, int y, int x, int H, int W) {
  return (x >= 0 && x < W && y >= 0 && y < H) ? __ldg(img + size_t(y) * W + x) : 0.0f;
}

struct LossParams {
  float threshold, lambda_mse, lambda_dssim, C1, C2;
  int clamp01;   // LOSS variants: img1 is the raw render; clamp it to [0,1] on load (render()'s clamp, fused)
};

// torch.clamp(x, 0, 1) including its NaN behaviour (NaN stays NaN)
__device__ __forceinline__ float clamp01f(float v) { return v < 0.f ? 0.f : (v > 1.f ? 1.f : v); }

// Four fp64 block sums at once; the totals end in threads 0..3 (sum t in thread t). The first two butterfly steps halve
// the number of values a lane still carries (lanes keep sums {0,1} or {2,3}, then one of the pair): 12 SHFLs and 6
// DADDs per warp instead of 40 and 20 for four plain butterflies, and one barrier instead of eight. s_red: NT/32 x 4.
__device__ __forceinline__ double block_sum4_d(double (&v)[4], double* s_red) {
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const bool h4 = (lane & 16u) != 0u, h3 = (lane & 8u) != 0u;
  double k0 = h4 ? v[2] : v[0], k1 = h4 ? v[3] : v[1];
  k0 += __shfl_xor_sync(0xffffffffu, h4 ? v[0] : v[2], 16);
  k1 += __shfl_xor_sync(0xffffffffu, h4 ? v[1] : v[3], 16);
  double k = h3 ? k1 : k0;
  k += __shfl_xor_sync(0xffffffffu, h3 ? k0 : k1, 8);
  k += __shfl_xor_sync(0xffffffffu, k, 4);
  k += __shfl_xor_sync(0xffffffffu, k, 2);
  k += __shfl_xor_sync(0xffffffffu, k, 1);
  if ((lane & 7u) == 0u) s_red[warp * 4 + (lane >> 3)] = k;   // lanes 0, 8, 16, 24 hold the warp's sums 0..3
  __syncthreads();
  double t = 0.0;
  if (threadIdx.x < 4)
    for (int w = 0; w < NT / 32; ++w) t += s_red[w * 4 + threadIdx.x];
  return t;   // valid on threads 0..3
}

// stats block of the LOSS variant (8 doubles, zeroed by the launcher):
//   [0] sum ssim  [1] sum sq over gt>thr  [2] sum sq over gt<=thr  [3] count gt>thr
//   [4] CTA ticket (low 32 bits)  [5] loss  [6] w_pos  [7] w_neg
template <bool LOSS>
__global__ void __launch_bounds__(NT, 3)
ssim_fwd_kernel(int H, int W, LossParams prm, const float* __restrict__ img1, const float* __restrict__ img2,
                float* __restrict__ ssim_map, float* __restrict__ dm_dmu1, float* __restrict__ dm_dsigma1_sq,
                float* __restrict__ dm_dsigma12, double* __restrict__ stats, float* __restrict__ loss_out) {
  pdl_wait();
  __shared__ float s1[IN][IN + 1];
  __shared__ float s2[IN][IN + 1];
  __shared__ float hq[5][IN][TS + 1];   // +1: the RUN-strided stores of the horizontal pass stay conflict-free
  __shared__ double s_red[NT / 32 * 4];
  const size_t plane = size_t(blockIdx.z) * H * W;
  const float* a = img1 + plane;
  const float* b = img2 + plane;
  const int x0 = blockIdx.x * TS, y0 = blockIdx.y * TS;
  for (int i = threadIdx.x; i < IN * IN; i += NT) {
    const int ly = i / IN, lx = i - ly * IN;
    float va = load_px(a, y0 + ly - HALO, x0 + lx - HALO, H, W);
    if (LOSS && prm.clamp01) va = clamp01f(va);
    s1[ly][lx] = va;
    s2[ly][lx] = load_px(b, y0 + ly - HALO, x0 + lx - HALO, H, W);
  }
  __syncthreads();
  // horizontal pass: rows 0..41, RUN consecutive output columns per work item
  for (int i = threadIdx.x; i < IN * (TS / RUN); i += NT) {
    const int ly = i / (TS / RUN), lx = (i - ly * (TS / RUN)) * RUN;
    float p[WIN], q[WIN];
#pragma unroll
    for (int k = 0; k < WIN; ++k) { p[k] = s1[ly][lx + k]; q[k] = s2[ly][lx + k]; }
#pragma unroll
    for (int r = 0; r < RUN; ++r) {
      float m1 = 0.f, m2 = 0.f, e11 = 0.f, e22 = 0.f, e12 = 0.f;
#pragma unroll
      for (int k = 0; k < 11; ++k) {
        const float g = c_tap[k];
        m1 += g * p[r + k];
        m2 += g * q[r + k];
        e11 += g * (p[r + k] * p[r + k]);
        e22 += g * (q[r + k] * q[r + k]);
        e12 += g * (p[r + k] * q[r + k]);
      }
      hq[0][ly][lx + r] = m1; hq[1][ly][lx + r] = m2; hq[2][ly][lx + r] = e11;
      hq[3][ly][lx + r] = e22; hq[4][ly][lx + r] = e12;
    }
  }
  __syncthreads();
  // vertical pass: RUN consecutive output rows per thread
  const int lx = threadIdx.x & 31;
  const int ly0 = (threadIdx.x >> 5) * RUN;
  // one channel at a time: a 14-value column window and 4 accumulators are live, not 5 x 14
  float acc[5][RUN];
#pragma unroll
  for (int c = 0; c < 5; ++c) {
    float col[WIN];
#pragma unroll
    for (int k = 0; k < WIN; ++k) col[k] = hq[c][ly0 + k][lx];
#pragma unroll
    for (int r = 0; r < RUN; ++r) {
      float v = 0.f;
#pragma unroll
      for (int k = 0; k < 11; ++k) v += c_tap[k] * col[r + k];
      acc[c][r] = v;
    }
  }
  double a_ssim = 0.0, a_pos = 0.0, a_neg = 0.0, a_cnt = 0.0;
#pragma unroll
  for (int r = 0; r < RUN; ++r) {
    const float mu1 = acc[0][r], mu2 = acc[1][r], e11 = acc[2][r], e22 = acc[3][r], e12 = acc[4][r];
    const int ly = ly0 + r;
    const int x = x0 + lx, y = y0 + ly;
    if (x < W && y < H) {
      const float sigma1_sq = e11 - mu1 * mu1;
      const float sigma2_sq = e22 - mu2 * mu2;
      const float sigma12 = e12 - mu1 * mu2;
      const float mu1_sq = mu1 * mu1, mu2_sq = mu2 * mu2, mu1_mu2 = mu1 * mu2;
      const float Cc = (2.0f * mu1_mu2 + prm.C1);
      const float D = (2.0f * sigma12 + prm.C2);
      const float A = (mu1_sq + mu2_sq + prm.C1);
      const float Bq = (sigma1_sq + sigma2_sq + prm.C2);
      const size_t o = plane + size_t(y) * W + x;
      const float m = (Cc * D) / (A * Bq);
      if (LOSS) a_ssim += double(m);
      else ssim_map[o] = m;
      if (dm_dmu1) {
        dm_dmu1[o] = ((mu2 * 2.0f * D) / (A * Bq) - (mu2 * 2.0f * Cc) / (A * Bq) - (mu1 * 2.0f * Cc * D) / (A * A * Bq) +
                      (mu1 * 2.0f * Cc * D) / (A * Bq * Bq));
        dm_dsigma1_sq[o] = ((-Cc * D) / (A * Bq * Bq));
        dm_dsigma12[o] = ((2 * Cc) / (A * Bq));
      }
      if (LOSS) {
        const float pp = s1[ly + HALO][lx + HALO], qq = s2[ly + HALO][lx + HALO];
        const float d = pp - qq;
        const float sq = d * d;
        if (qq > prm.threshold) { a_pos += double(sq); a_cnt += 1.0; }
        else a_neg += double(sq);
      }
    }
  }
  if (LOSS) {
    double part[4] = {a_ssim, a_pos, a_neg, a_cnt};
    const double tsum = block_sum4_d(part, s_red);
    if (threadIdx.x < 4) { atomicAdd(stats + threadIdx.x, tsum); __threadfence(); }
    __syncwarp();
    if (threadIdx.x == 0) {
      __threadfence();
      const unsigned total = gridDim.x * gridDim.y;
      const unsigned ticket = atomicAdd(reinterpret_cast<unsigned*>(stats + 4), 1u);
      if (ticket == total - 1) {
        // last CTA: every partial sum is in; finish the scalar on the device (no host sync)
        __threadfence();
        volatile double* vs = stats;
        const double N = double(H) * double(W);
        // weights in fp32 like the torch expression (loss_utils.py:111-112)
        const float np = float(vs[3]), nn = float(N - vs[3]);
        const float w_pos = 5.f * (nn + 1.f) / (np + nn);
        const float w_neg = 1.0f * (np + 1.f) / (np + nn);
        const double Ll1 = (double(w_pos) * vs[1] + double(w_neg) * vs[2]) / N;
        const double ssim = vs[0] / N;
        const double loss = double(prm.lambda_mse) *
                            ((1.0 - double(prm.lambda_dssim)) * Ll1 + double(prm.lambda_dssim) * (1.0 - ssim));
        stats[5] = loss;
        stats[6] = double(w_pos);
        stats[7] = double(w_neg);
        *loss_out = float(loss);
      }
    }
  }
}

// LOSS == false: dL_dimg1 = conv(dL*dm_dmu1) + 2*img1*conv(dL*dm_dsigma1_sq) + img2*conv(dL*dm_dsigma12)
// LOSS == true : dL/dmap is the uniform -g*lambda_mse*lambda_dssim/N (factored out of the convolutions)
//                and the class-balanced MSE term of edge_aware_loss is added in the epilogue.
template <bool LOSS>
__global__ void __launch_bounds__(NT)
ssim_bwd_kernel(int H, int W, LossParams prm, const float* __restrict__ img1, const float* __restrict__ img2,
                const float* __restrict__ dL_dmap, const float* __restrict__ dm_dmu1,
                const float* __restrict__ dm_dsigma1_sq, const float* __restrict__ dm_dsigma12,
                const double* __restrict__ stats, const float* __restrict__ g_loss, float* __restrict__ dL_dimg1) {
  pdl_wait();
  __shared__ float sp[3][IN][IN + 1];
  __shared__ float hq[3][IN][TS + 1];
  const size_t plane = size_t(blockIdx.z) * H * W;
  const int x0 = blockIdx.x * TS, y0 = blockIdx.y * TS;
  for (int i = threadIdx.x; i < IN * IN; i += NT) {
    const int ly = i / IN, lx = i - ly * IN;
    const int y = y0 + ly - HALO, x = x0 + lx - HALO;
    const float g = LOSS ? 1.0f : load_px(dL_dmap + plane, y, x, H, W);
    sp[0][ly][lx] = load_px(dm_dmu1 + plane, y, x, H, W) * g;
    sp[1][ly][lx] = load_px(dm_dsigma1_sq + plane, y, x, H, W) * g;
    sp[2][ly][lx] = load_px(dm_dsigma12 + plane, y, x, H, W) * g;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < IN * (TS / RUN); i += NT) {
    const int ly = i / (TS / RUN), lx = (i - ly * (TS / RUN)) * RUN;
    float w0[WIN], w1[WIN], w2[WIN];
#pragma unroll
    for (int k = 0; k < WIN; ++k) { w0[k] = sp[0][ly][lx + k]; w1[k] = sp[1][ly][lx + k]; w2[k] = sp[2][ly][lx + k]; }
#pragma unroll
    for (int r = 0; r < RUN; ++r) {
      float v0 = 0.f, v1 = 0.f, v2 = 0.f;
#pragma unroll
      for (int k = 0; k < 11; ++k) {
        const float g = c_tap[k];
        v0 += g * w0[r + k];
        v1 += g * w1[r + k];
        v2 += g * w2[r + k];
      }
      hq[0][ly][lx + r] = v0; hq[1][ly][lx + r] = v1; hq[2][ly][lx + r] = v2;
    }
  }
  __syncthreads();
  float k_mse = 0.f, k_ssim = 1.f, w_pos = 0.f, w_neg = 0.f;
  if (LOSS) {
    const float g = g_loss ? __ldg(g_loss) : 1.0f;
    const float invN = 1.0f / (float(H) * float(W));
    w_pos = float(stats[6]);
    w_neg = float(stats[7]);
    k_mse = g * prm.lambda_mse * (1.0f - prm.lambda_dssim) * 2.0f * invN;
    k_ssim = -g * prm.lambda_mse * prm.lambda_dssim * invN;
  }
  const int lx = threadIdx.x & 31;
  const int ly0 = (threadIdx.x >> 5) * RUN;
  float col[3][WIN];
#pragma unroll
  for (int c = 0; c < 3; ++c)
#pragma unroll
    for (int k = 0; k < WIN; ++k) col[c][k] = hq[c][ly0 + k][lx];
#pragma unroll
  for (int r = 0; r < RUN; ++r) {
    float v0 = 0.f, v1 = 0.f, v2 = 0.f;
#pragma unroll
    for (int k = 0; k < 11; ++k) {
      const float gk = c_tap[k];
      v0 += gk * col[0][r + k];
      v1 += gk * col[1][r + k];
      v2 += gk * col[2][r + k];
    }
    const int x = x0 + lx, y = y0 + ly0 + r;
    if (x < W && y < H) {
      const size_t o = plane + size_t(y) * W + x;
      const float p1raw = __ldg(img1 + o), p2 = __ldg(img2 + o);
      const float p1 = (LOSS && prm.clamp01) ? clamp01f(p1raw) : p1raw;
      float d = 0.f;
      d += v0;
      d += p1 * 2.0f * v1;
      d += p2 * v2;
      if (LOSS) {
        const float w = (p2 > prm.threshold) ? w_pos : w_neg;
        const float gval = k_mse * w * (p1 - p2) + k_ssim * d;
        // clamp's adjoint: the gradient passes where 0 <= x <= 1 (and not for NaN), like torch.clamp
        dL_dimg1[o] = (!prm.clamp01 || (p1raw >= 0.f && p1raw <= 1.f)) ? gval : 0.f;
      } else {
        dL_dimg1[o] = d;
      }
    }
  }
}

}  // namespace ssimk
}  // namespace cg
