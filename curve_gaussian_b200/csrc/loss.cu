// Fused image loss of the training step:
//
//   loss = lambda_mse * ((1 - lambda_dssim) * edge_aware_loss(img, gt) + lambda_dssim * (1 - ssim(img, gt)))
//
// i.e. train.py:101-107 with utils/loss_utils.py:94-115 (class-balanced weighted
// MSE; weights come from the counts of gt > threshold / gt <= threshold) and
// fused_ssim (fused-ssim/ssim.cu:187-366; separable 11-tap Gaussian, zero padding).
// The reference spends ~15 elementwise launches, 4 reductions and the SSIM pair on
// this; here ONE forward kernel reads both images once and leaves the scalar loss on
// the device (last-CTA finalisation, no host sync), and ONE backward kernel writes
// dL/dimg. Single-channel images (the pipeline renders 1 channel, config.h:15).
//
// Also here: the per-pixel 3x3 rotation render() applies to the direction channels
// (gaussian_renderer/__init__.py:144), which the reference runs as a (H*W,3)x(3,3) GEMM.
#include "common.cuh"

namespace cg {

namespace {
constexpr int TS = 32;
constexpr int HALO = 5;
constexpr int IN = TS + 2 * HALO;
constexpr int NT = 256;

__constant__ float c_tap_l[11] = {0.001028380123898387f, 0.0075987582094967365f, 0.036000773310661316f,
                                  0.10936068743467331f,  0.21300552785396576f,   0.26601171493530273f,
                                  0.21300552785396576f,  0.10936068743467331f,   0.036000773310661316f,
                                  0.0075987582094967365f, 0.001028380123898387f};

__device__ __forceinline__ float load_px_l(const float* __restrict__ img, int y, int x, int H, int W) {
  return (x >= 0 && x < W && y >= 0 && y < H) ? __ldg(img + size_t(y) * W + x) : 0.0f;
}

// stats block (8 doubles, zeroed by the launcher):
//   [0] sum ssim  [1] sum sq over gt>thr  [2] sum sq over gt<=thr  [3] count gt>thr
//   [4] CTA ticket (low 32 bits)  [5] loss  [6] w_pos  [7] w_neg
struct LossParams {
  float threshold, lambda_mse, lambda_dssim, C1, C2;
};

__device__ __forceinline__ double block_sum(double v, double* s_red) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0.0;
  if (threadIdx.x == 0)
    for (int i = 0; i < NT / 32; ++i) t += s_red[i];
  return t;  // valid on thread 0
}
}  // namespace

__global__ void __launch_bounds__(NT)
edge_ssim_loss_fwd_kernel(int H, int W, LossParams prm, const float* __restrict__ img1, const float* __restrict__ img2,
                          double* __restrict__ stats, float* __restrict__ loss_out, float* __restrict__ dm_dmu1,
                          float* __restrict__ dm_dsigma1_sq, float* __restrict__ dm_dsigma12) {
  __shared__ float s1[IN][IN + 1];
  __shared__ float s2[IN][IN + 1];
  __shared__ float hq[5][IN][TS];
  __shared__ double s_red[NT / 32];
  const int x0 = blockIdx.x * TS, y0 = blockIdx.y * TS;
  for (int i = threadIdx.x; i < IN * IN; i += NT) {
    const int ly = i / IN, lx = i - ly * IN;
    s1[ly][lx] = load_px_l(img1, y0 + ly - HALO, x0 + lx - HALO, H, W);
    s2[ly][lx] = load_px_l(img2, y0 + ly - HALO, x0 + lx - HALO, H, W);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < IN * TS; i += NT) {
    const int ly = i / TS, lx = i - ly * TS;
    float m1 = 0.f, m2 = 0.f, e11 = 0.f, e22 = 0.f, e12 = 0.f;
#pragma unroll
    for (int k = 0; k < 11; ++k) {
      const float p = s1[ly][lx + k], q = s2[ly][lx + k], g = c_tap_l[k];
      m1 += g * p;
      m2 += g * q;
      e11 += g * (p * p);
      e22 += g * (q * q);
      e12 += g * (p * q);
    }
    hq[0][ly][lx] = m1; hq[1][ly][lx] = m2; hq[2][ly][lx] = e11; hq[3][ly][lx] = e22; hq[4][ly][lx] = e12;
  }
  __syncthreads();
  const int lx = threadIdx.x & 31;
  double a_ssim = 0.0, a_pos = 0.0, a_neg = 0.0, a_cnt = 0.0;
  for (int ly = threadIdx.x >> 5; ly < TS; ly += NT / 32) {
    float mu1 = 0.f, mu2 = 0.f, e11 = 0.f, e22 = 0.f, e12 = 0.f;
#pragma unroll
    for (int k = 0; k < 11; ++k) {
      const float g = c_tap_l[k];
      mu1 += g * hq[0][ly + k][lx];
      mu2 += g * hq[1][ly + k][lx];
      e11 += g * hq[2][ly + k][lx];
      e22 += g * hq[3][ly + k][lx];
      e12 += g * hq[4][ly + k][lx];
    }
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
      const size_t o = size_t(y) * W + x;
      a_ssim += double((Cc * D) / (A * Bq));
      if (dm_dmu1) {
        dm_dmu1[o] = ((mu2 * 2.0f * D) / (A * Bq) - (mu2 * 2.0f * Cc) / (A * Bq) - (mu1 * 2.0f * Cc * D) / (A * A * Bq) +
                      (mu1 * 2.0f * Cc * D) / (A * Bq * Bq));
        dm_dsigma1_sq[o] = ((-Cc * D) / (A * Bq * Bq));
        dm_dsigma12[o] = ((2 * Cc) / (A * Bq));
      }
      const float p = s1[ly + HALO][lx + HALO], q = s2[ly + HALO][lx + HALO];
      const float d = p - q;
      const float sq = d * d;
      if (q > prm.threshold) { a_pos += double(sq); a_cnt += 1.0; }
      else a_neg += double(sq);
    }
  }
  const double t0 = block_sum(a_ssim, s_red);
  const double t1 = block_sum(a_pos, s_red);
  const double t2 = block_sum(a_neg, s_red);
  const double t3 = block_sum(a_cnt, s_red);
  if (threadIdx.x == 0) {
    atomicAdd(stats + 0, t0);
    atomicAdd(stats + 1, t1);
    atomicAdd(stats + 2, t2);
    atomicAdd(stats + 3, t3);
    __threadfence();
    const unsigned total = gridDim.x * gridDim.y;
    const unsigned ticket = atomicAdd(reinterpret_cast<unsigned*>(stats + 4), 1u);
    if (ticket == total - 1) {
      __threadfence();
      volatile double* vs = stats;
      const double N = double(H) * double(W);
      // weights in fp32 like the torch expression (loss_utils.py:111-112)
      const float np = float(vs[3]), nn = float(N - vs[3]);
      const float w_pos = 5.f * (nn + 1.f) / (np + nn);
      const float w_neg = 1.0f * (np + 1.f) / (np + nn);
      const double Ll1 = (double(w_pos) * vs[1] + double(w_neg) * vs[2]) / N;
      const double ssim = vs[0] / N;
      const double loss = double(prm.lambda_mse) * ((1.0 - double(prm.lambda_dssim)) * Ll1 + double(prm.lambda_dssim) * (1.0 - ssim));
      stats[5] = loss;
      stats[6] = double(w_pos);
      stats[7] = double(w_neg);
      *loss_out = float(loss);
    }
  }
}

__global__ void __launch_bounds__(NT)
edge_ssim_loss_bwd_kernel(int H, int W, LossParams prm, const float* __restrict__ img1, const float* __restrict__ img2,
                          const double* __restrict__ stats, const float* __restrict__ g_loss,
                          const float* __restrict__ dm_dmu1, const float* __restrict__ dm_dsigma1_sq,
                          const float* __restrict__ dm_dsigma12, float* __restrict__ dL_dimg1) {
  __shared__ float sp[3][IN][IN + 1];
  __shared__ float hq[3][IN][TS];
  const int x0 = blockIdx.x * TS, y0 = blockIdx.y * TS;
  for (int i = threadIdx.x; i < IN * IN; i += NT) {
    const int ly = i / IN, lx = i - ly * IN;
    const int y = y0 + ly - HALO, x = x0 + lx - HALO;
    sp[0][ly][lx] = load_px_l(dm_dmu1, y, x, H, W);
    sp[1][ly][lx] = load_px_l(dm_dsigma1_sq, y, x, H, W);
    sp[2][ly][lx] = load_px_l(dm_dsigma12, y, x, H, W);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < IN * TS; i += NT) {
    const int ly = i / TS, lx = i - ly * TS;
    float v0 = 0.f, v1 = 0.f, v2 = 0.f;
#pragma unroll
    for (int k = 0; k < 11; ++k) {
      const float g = c_tap_l[k];
      v0 += g * sp[0][ly][lx + k];
      v1 += g * sp[1][ly][lx + k];
      v2 += g * sp[2][ly][lx + k];
    }
    hq[0][ly][lx] = v0; hq[1][ly][lx] = v1; hq[2][ly][lx] = v2;
  }
  __syncthreads();
  const float g = g_loss ? __ldg(g_loss) : 1.0f;
  const float invN = 1.0f / (float(H) * float(W));
  const float w_pos = float(stats[6]), w_neg = float(stats[7]);
  const float k_mse = g * prm.lambda_mse * (1.0f - prm.lambda_dssim) * 2.0f * invN;
  const float k_ssim = -g * prm.lambda_mse * prm.lambda_dssim * invN;
  const int lx = threadIdx.x & 31;
  for (int ly = threadIdx.x >> 5; ly < TS; ly += NT / 32) {
    float v0 = 0.f, v1 = 0.f, v2 = 0.f;
#pragma unroll
    for (int k = 0; k < 11; ++k) {
      const float gk = c_tap_l[k];
      v0 += gk * hq[0][ly + k][lx];
      v1 += gk * hq[1][ly + k][lx];
      v2 += gk * hq[2][ly + k][lx];
    }
    const int x = x0 + lx, y = y0 + ly;
    if (x < W && y < H) {
      const size_t o = size_t(y) * W + x;
      const float p1 = __ldg(img1 + o), p2 = __ldg(img2 + o);
      const float dssim = v0 + p1 * 2.0f * v1 + p2 * v2;
      const float w = (p2 > prm.threshold) ? w_pos : w_neg;
      dL_dimg1[o] = k_mse * w * (p1 - p2) + k_ssim * dssim;
    }
  }
}

// out[c][i] = sum_k in[k][i] * M[c][k] with M row-major 3x3 at m[ld*c + k] (ld = row stride in floats).
// transpose != 0 uses M[k][c] instead (the adjoint).
__global__ void __launch_bounds__(256)
rotate_channels_kernel(int64_t n, const float* __restrict__ in, const float* __restrict__ m, int ld, int transpose,
                       float* __restrict__ out) {
  __shared__ float sm[9];
  if (threadIdx.x < 9) {
    const int c = threadIdx.x / 3, k = threadIdx.x % 3;
    sm[threadIdx.x] = transpose ? m[ld * k + c] : m[ld * c + k];
  }
  __syncthreads();
  const int64_t i4 = (int64_t(blockIdx.x) * 256 + threadIdx.x) * 4;
  if (i4 >= n) return;
  if (i4 + 3 < n && (n & 3) == 0) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(in + i4));
    const float4 b = __ldg(reinterpret_cast<const float4*>(in + n + i4));
    const float4 c = __ldg(reinterpret_cast<const float4*>(in + 2 * n + i4));
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      float4 o;
      o.x = a.x * sm[3 * r] + b.x * sm[3 * r + 1] + c.x * sm[3 * r + 2];
      o.y = a.y * sm[3 * r] + b.y * sm[3 * r + 1] + c.y * sm[3 * r + 2];
      o.z = a.z * sm[3 * r] + b.z * sm[3 * r + 1] + c.z * sm[3 * r + 2];
      o.w = a.w * sm[3 * r] + b.w * sm[3 * r + 1] + c.w * sm[3 * r + 2];
      *reinterpret_cast<float4*>(out + r * n + i4) = o;
    }
  } else {
    for (int64_t i = i4; i < n && i < i4 + 4; ++i) {
      const float a = in[i], b = in[n + i], c = in[2 * n + i];
#pragma unroll
      for (int r = 0; r < 3; ++r) out[r * n + i] = a * sm[3 * r] + b * sm[3 * r + 1] + c * sm[3 * r + 2];
    }
  }
}

}  // namespace cg

using namespace cg;

extern "C" {

size_t cg_edge_ssim_loss_stats_bytes(void) { return 8 * sizeof(double); }

int cg_edge_ssim_loss_fwd(int32_t H, int32_t W, const float* img, const float* gt, float threshold, float lambda_mse,
                          float lambda_dssim, float C1, float C2, void* stats, float* loss_out, float* dm_dmu1,
                          float* dm_dsigma1_sq, float* dm_dsigma12, void* stream) {
  CG_ARG(H > 0 && W > 0, "image shape");
  CG_ARG(img && gt && stats && loss_out, "edge_ssim_loss_fwd pointers");
  CG_ARG((dm_dmu1 && dm_dsigma1_sq && dm_dsigma12) || (!dm_dmu1 && !dm_dsigma1_sq && !dm_dsigma12),
         "the three partial maps come together");
  CG_ARG((reinterpret_cast<uintptr_t>(stats) & 7u) == 0, "stats must be 8-byte aligned");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  CG_CUDA(cudaMemsetAsync(stats, 0, 8 * sizeof(double), st));
  dim3 grid((W + TS - 1) / TS, (H + TS - 1) / TS);
  LossParams prm{threshold, lambda_mse, lambda_dssim, C1, C2};
  StageTimer t_(ST_LOSS_FWD, st, 1);
  edge_ssim_loss_fwd_kernel<<<grid, NT, 0, st>>>(H, W, prm, img, gt, reinterpret_cast<double*>(stats), loss_out,
                                                dm_dmu1, dm_dsigma1_sq, dm_dsigma12);
  CG_LAUNCH_CHECK(0, st);
  return CG_OK;
}

int cg_edge_ssim_loss_bwd(int32_t H, int32_t W, const float* img, const float* gt, float threshold, float lambda_mse,
                          float lambda_dssim, const void* stats, const float* g_loss, const float* dm_dmu1,
                          const float* dm_dsigma1_sq, const float* dm_dsigma12, float* dL_dimg, void* stream) {
  CG_ARG(H > 0 && W > 0, "image shape");
  CG_ARG(img && gt && stats && dm_dmu1 && dm_dsigma1_sq && dm_dsigma12 && dL_dimg, "edge_ssim_loss_bwd pointers");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  dim3 grid((W + TS - 1) / TS, (H + TS - 1) / TS);
  LossParams prm{threshold, lambda_mse, lambda_dssim, 0.f, 0.f};
  StageTimer t_(ST_LOSS_BWD, st, 1);
  edge_ssim_loss_bwd_kernel<<<grid, NT, 0, st>>>(H, W, prm, img, gt, reinterpret_cast<const double*>(stats), g_loss,
                                                dm_dmu1, dm_dsigma1_sq, dm_dsigma12, dL_dimg);
  CG_LAUNCH_CHECK(0, st);
  return CG_OK;
}

int cg_rotate_channels(int64_t n, const float* in, const float* m3x3, int32_t ld, int32_t transpose, float* out,
                       void* stream) {
  if (n == 0) return CG_OK;
  CG_ARG(n > 0 && in && m3x3 && out && ld >= 3, "rotate_channels arguments");
  CG_ARG(((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15u) == 0, "16-byte aligned planes");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  count_launches(1);
  const int64_t nthreads = (n + 3) / 4;
  rotate_channels_kernel<<<unsigned((nthreads + 255) / 256), 256, 0, st>>>(n, in, m3x3, ld, transpose, out);
  CG_LAUNCH_CHECK(0, st);
  return CG_OK;
}

}  // extern "C"
