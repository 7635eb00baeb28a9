// Fused image loss of the training step:
//
//   loss = lambda_mse * ((1 - lambda_dssim) * edge_aware_loss(img, gt) + lambda_dssim * (1 - ssim(img, gt)))
//
// i.e. train.py:101-107 with utils/loss_utils.py:94-115 (class-balanced weighted
// MSE; weights come from the counts of gt > threshold / gt <= threshold) and
// fused_ssim (fused-ssim/ssim.cu:187-366; separable 11-tap Gaussian, zero padding).
// The reference spends ~15 elementwise launches, 4 reductions and the SSIM pair on
// this; here ONE forward kernel (ssim_fwd_kernel<true>, ssim_kernels.cuh) reads both images
// once and leaves the scalar loss on the device (last-CTA finalisation, no host sync), and
// ONE backward kernel (ssim_bwd_kernel<true>) writes dL/dimg. Single-channel images (the
// pipeline renders 1 channel, config.h:15).
//
// Also here: the per-pixel 3x3 rotation render() applies to the direction channels
// (gaussian_renderer/__init__.py:144), which the reference runs as a (H*W,3)x(3,3) GEMM.
#include "ssim_kernels.cuh"

namespace cg {

// out[c][i] = sum_k in[k][i] * M[c][k] with M row-major 3x3 at m[ld*c + k] (ld = row stride in floats).
// transpose != 0 uses M[k][c] instead (the adjoint).
__global__ void __launch_bounds__(256)
rotate_channels_kernel(int64_t n, const float* __restrict__ in, const float* __restrict__ m, int ld, int transpose,
                       float* __restrict__ out) {
  pdl_wait();
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
using namespace cg::ssimk;

extern "C" {

size_t cg_edge_ssim_loss_stats_bytes(void) { return 8 * sizeof(double); }

int cg_edge_ssim_loss_fwd(int32_t H, int32_t W, const float* img, const float* gt, float threshold, float lambda_mse,
                          float lambda_dssim, float C1, float C2, int32_t clamp01, void* stats, float* loss_out,
                          float* dm_dmu1, float* dm_dsigma1_sq, float* dm_dsigma12, void* stream) {
  CG_ARG(H > 0 && W > 0, "image shape");
  CG_ARG(img && gt && stats && loss_out, "edge_ssim_loss_fwd pointers");
  CG_ARG((dm_dmu1 && dm_dsigma1_sq && dm_dsigma12) || (!dm_dmu1 && !dm_dsigma1_sq && !dm_dsigma12),
         "the three partial maps come together");
  CG_ARG((reinterpret_cast<uintptr_t>(stats) & 7u) == 0, "stats must be 8-byte aligned");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  CG_CUDA(cudaMemsetAsync(stats, 0, 8 * sizeof(double), st));
  dim3 grid((W + TS - 1) / TS, (H + TS - 1) / TS);
  LossParams prm{threshold, lambda_mse, lambda_dssim, C1, C2, clamp01 ? 1 : 0};
  StageTimer t_(ST_LOSS_FWD, st, 1);
  launch_k(ssim_fwd_kernel<true>, dim3(grid), dim3(NT), 0, st, H, W, prm, img, gt, nullptr, dm_dmu1, dm_dsigma1_sq, dm_dsigma12,
                                            reinterpret_cast<double*>(stats), loss_out);
  CG_LAUNCH_CHECK(0, st);
  return CG_OK;
}

int cg_edge_ssim_loss_bwd(int32_t H, int32_t W, const float* img, const float* gt, float threshold, float lambda_mse,
                          float lambda_dssim, int32_t clamp01, const void* stats, const float* g_loss, const float* dm_dmu1,
                          const float* dm_dsigma1_sq, const float* dm_dsigma12, float* dL_dimg, void* stream) {
  CG_ARG(H > 0 && W > 0, "image shape");
  CG_ARG(img && gt && stats && dm_dmu1 && dm_dsigma1_sq && dm_dsigma12 && dL_dimg, "edge_ssim_loss_bwd pointers");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  dim3 grid((W + TS - 1) / TS, (H + TS - 1) / TS);
  LossParams prm{threshold, lambda_mse, lambda_dssim, 0.f, 0.f, clamp01 ? 1 : 0};
  StageTimer t_(ST_LOSS_BWD, st, 1);
  launch_k(ssim_bwd_kernel<true>, dim3(grid), dim3(NT), 0, st, H, W, prm, img, gt, nullptr, dm_dmu1, dm_dsigma1_sq, dm_dsigma12,
                                            reinterpret_cast<const double*>(stats), g_loss, dL_dimg);
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
  launch_k(rotate_channels_kernel, dim3(unsigned((nthreads + 255) / 256)), dim3(256), 0, st, n, in, m3x3, ld, transpose, out);
  CG_LAUNCH_CHECK(0, st);
  return CG_OK;
}

}  // extern "C"
