// Host entry points of the fused SSIM (replaces submodules/fused-ssim); kernels in ssim_kernels.cuh.
#include "ssim_kernels.cuh"

using namespace cg;
using namespace cg::ssimk;

extern "C" {

int cg_ssim_fwd(int32_t B, int32_t CH, int32_t H, int32_t W, float C1, float C2, const float* img1,
                const float* img2, float* ssim_map, float* dm_dmu1, float* dm_dsigma1_sq, float* dm_dsigma12,
                void* stream) {
  if (B * CH == 0 || H == 0 || W == 0) return CG_OK;
  CG_ARG(B > 0 && CH > 0 && H > 0 && W > 0, "ssim shape");
  CG_ARG(img1 && img2 && ssim_map, "ssim_fwd pointers");
  CG_ARG((dm_dmu1 && dm_dsigma1_sq && dm_dsigma12) || (!dm_dmu1 && !dm_dsigma1_sq && !dm_dsigma12),
         "the three partial maps come together");
  CG_ARG(int64_t(B) * CH <= 65535, "B*CH");
  dim3 grid((W + TS - 1) / TS, (H + TS - 1) / TS, B * CH);
  StageTimer t_(ST_SSIM_FWD, reinterpret_cast<cudaStream_t>(stream), 1);
  LossParams prm{0.f, 0.f, 0.f, C1, C2, 0};
  launch_k(ssim_fwd_kernel<false>, dim3(grid), dim3(NT), 0, reinterpret_cast<cudaStream_t>(stream), H, W, prm, img1, img2, ssim_map, dm_dmu1, dm_dsigma1_sq, dm_dsigma12, nullptr, nullptr);
  CG_LAUNCH_CHECK(0, reinterpret_cast<cudaStream_t>(stream));
  return CG_OK;
}

int cg_ssim_bwd(int32_t B, int32_t CH, int32_t H, int32_t W, float C1, float C2, const float* img1,
                const float* img2, const float* dL_dmap, const float* dm_dmu1, const float* dm_dsigma1_sq,
                const float* dm_dsigma12, float* dL_dimg1, void* stream) {
  (void)C1; (void)C2;
  if (B * CH == 0 || H == 0 || W == 0) return CG_OK;
  CG_ARG(B > 0 && CH > 0 && H > 0 && W > 0, "ssim shape");
  CG_ARG(img1 && img2 && dL_dmap && dm_dmu1 && dm_dsigma1_sq && dm_dsigma12 && dL_dimg1, "ssim_bwd pointers");
  CG_ARG(int64_t(B) * CH <= 65535, "B*CH");
  dim3 grid((W + TS - 1) / TS, (H + TS - 1) / TS, B * CH);
  StageTimer t_(ST_SSIM_BWD, reinterpret_cast<cudaStream_t>(stream), 1);
  LossParams prm{0.f, 0.f, 0.f, 0.f, 0.f, 0};
  launch_k(ssim_bwd_kernel<false>, dim3(grid), dim3(NT), 0, reinterpret_cast<cudaStream_t>(stream), H, W, prm, img1, img2, dL_dmap, dm_dmu1, dm_dsigma1_sq, dm_dsigma12, nullptr, nullptr, dL_dimg1);
  CG_LAUNCH_CHECK(0, reinterpret_cast<cudaStream_t>(stream));
  return CG_OK;
}

}  // extern "C"
