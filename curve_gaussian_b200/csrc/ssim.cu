// Fused SSIM forward / backward (replaces submodules/fused-ssim, ssim.cu:187-366).
//
// Same math as the reference: separable 11-tap Gaussian (sigma 1.5, taps
// ssim.cu:9-19), zero padding, x pass then y pass, per-pixel SSIM map and the
// three partial maps dm/dmu1, dm/dsigma1^2, dm/dsigma12; backward is three
// convolutions of (dL * partial). The reference runs five (forward) / three
// (backward) separate conv rounds through one scratch tile with ~20 block
// barriers; here one 32x32 output tile per CTA does ONE horizontal pass that
// produces all five (three) row-filtered quantities at once, then one vertical
// pass, i.e. 3 barriers per tile and each input pixel is read from HBM once.
#include "common.cuh"

namespace cg {

namespace {
constexpr int TS = 32;         // output tile
constexpr int HALO = 5;
constexpr int IN = TS + 2 * HALO;   // 42
constexpr int NT = 256;

__constant__ float c_tap[11] = {0.001028380123898387f, 0.0075987582094967365f, 0.036000773310661316f,
                                0.10936068743467331f,  0.21300552785396576f,   0.26601171493530273f,
                                0.21300552785396576f,  0.10936068743467331f,   0.036000773310661316f,
                                0.0075987582094967365f, 0.001028380123898387f};

__device__ __forceinline__ float load_px(const float* __restrict__ img, int y, int x, int H, int W) {
  return (x >= 0 && x < W && y >= 0 && y < H) ? __ldg(img + size_t(y) * W + x) : 0.0f;
}
}  // namespace

__global__ void __launch_bounds__(NT)
ssim_fwd_kernel(int H, int W, float C1, float C2, const float* __restrict__ img1, const float* __restrict__ img2,
                float* __restrict__ ssim_map, float* __restrict__ dm_dmu1, float* __restrict__ dm_dsigma1_sq,
                float* __restrict__ dm_dsigma12) {
  __shared__ float s1[IN][IN + 1];
  __shared__ float s2[IN][IN + 1];
  __shared__ float hq[5][IN][TS];
  const size_t plane = size_t(blockIdx.z) * H * W;
  const float* a = img1 + plane;
  const float* b = img2 + plane;
  const int x0 = blockIdx.x * TS, y0 = blockIdx.y * TS;
  for (int i = threadIdx.x; i < IN * IN; i += NT) {
    const int ly = i / IN, lx = i - ly * IN;
    s1[ly][lx] = load_px(a, y0 + ly - HALO, x0 + lx - HALO, H, W);
    s2[ly][lx] = load_px(b, y0 + ly - HALO, x0 + lx - HALO, H, W);
  }
  __syncthreads();
  // horizontal pass: rows 0..41, output columns 0..31
  for (int i = threadIdx.x; i < IN * TS; i += NT) {
    const int ly = i / TS, lx = i - ly * TS;
    float m1 = 0.f, m2 = 0.f, e11 = 0.f, e22 = 0.f, e12 = 0.f;
#pragma unroll
    for (int k = 0; k < 11; ++k) {
      const float p = s1[ly][lx + k], q = s2[ly][lx + k], g = c_tap[k];
      m1 += g * p;
      m2 += g * q;
      e11 += g * (p * p);
      e22 += g * (q * q);
      e12 += g * (p * q);
    }
    hq[0][ly][lx] = m1; hq[1][ly][lx] = m2; hq[2][ly][lx] = e11; hq[3][ly][lx] = e22; hq[4][ly][lx] = e12;
  }
  __syncthreads();
  // vertical pass: 4 output pixels per thread
  const int lx = threadIdx.x & 31;
  for (int ly = threadIdx.x >> 5; ly < TS; ly += NT / 32) {
    float mu1 = 0.f, mu2 = 0.f, e11 = 0.f, e22 = 0.f, e12 = 0.f;
#pragma unroll
    for (int k = 0; k < 11; ++k) {
      const float g = c_tap[k];
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
      const float Cc = (2.0f * mu1_mu2 + C1);
      const float D = (2.0f * sigma12 + C2);
      const float A = (mu1_sq + mu2_sq + C1);
      const float Bq = (sigma1_sq + sigma2_sq + C2);
      const size_t o = plane + size_t(y) * W + x;
      ssim_map[o] = (Cc * D) / (A * Bq);
      if (dm_dmu1) {
        dm_dmu1[o] = ((mu2 * 2.0f * D) / (A * Bq) - (mu2 * 2.0f * Cc) / (A * Bq) - (mu1 * 2.0f * Cc * D) / (A * A * Bq) +
                      (mu1 * 2.0f * Cc * D) / (A * Bq * Bq));
        dm_dsigma1_sq[o] = ((-Cc * D) / (A * Bq * Bq));
        dm_dsigma12[o] = ((2 * Cc) / (A * Bq));
      }
    }
  }
}

__global__ void __launch_bounds__(NT)
ssim_bwd_kernel(int H, int W, const float* __restrict__ img1, const float* __restrict__ img2,
                const float* __restrict__ dL_dmap, const float* __restrict__ dm_dmu1,
                const float* __restrict__ dm_dsigma1_sq, const float* __restrict__ dm_dsigma12,
                float* __restrict__ dL_dimg1) {
  __shared__ float sp[3][IN][IN + 1];
  __shared__ float hq[3][IN][TS];
  const size_t plane = size_t(blockIdx.z) * H * W;
  const int x0 = blockIdx.x * TS, y0 = blockIdx.y * TS;
  for (int i = threadIdx.x; i < IN * IN; i += NT) {
    const int ly = i / IN, lx = i - ly * IN;
    const int y = y0 + ly - HALO, x = x0 + lx - HALO;
    const float g = load_px(dL_dmap + plane, y, x, H, W);
    sp[0][ly][lx] = load_px(dm_dmu1 + plane, y, x, H, W) * g;
    sp[1][ly][lx] = load_px(dm_dsigma1_sq + plane, y, x, H, W) * g;
    sp[2][ly][lx] = load_px(dm_dsigma12 + plane, y, x, H, W) * g;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < IN * TS; i += NT) {
    const int ly = i / TS, lx = i - ly * TS;
    float v0 = 0.f, v1 = 0.f, v2 = 0.f;
#pragma unroll
    for (int k = 0; k < 11; ++k) {
      const float g = c_tap[k];
      v0 += g * sp[0][ly][lx + k];
      v1 += g * sp[1][ly][lx + k];
      v2 += g * sp[2][ly][lx + k];
    }
    hq[0][ly][lx] = v0; hq[1][ly][lx] = v1; hq[2][ly][lx] = v2;
  }
  __syncthreads();
  const int lx = threadIdx.x & 31;
  for (int ly = threadIdx.x >> 5; ly < TS; ly += NT / 32) {
    float v0 = 0.f, v1 = 0.f, v2 = 0.f;
#pragma unroll
    for (int k = 0; k < 11; ++k) {
      const float g = c_tap[k];
      v0 += g * hq[0][ly + k][lx];
      v1 += g * hq[1][ly + k][lx];
      v2 += g * hq[2][ly + k][lx];
    }
    const int x = x0 + lx, y = y0 + ly;
    if (x < W && y < H) {
      const size_t o = plane + size_t(y) * W + x;
      const float p1 = img1[o], p2 = img2[o];
      float d = 0.f;
      d += v0;
      d += p1 * 2.0f * v1;
      d += p2 * v2;
      dL_dimg1[o] = d;
    }
  }
}

}  // namespace cg

using namespace cg;

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
  ssim_fwd_kernel<<<grid, NT, 0, reinterpret_cast<cudaStream_t>(stream)>>>(H, W, C1, C2, img1, img2, ssim_map,
                                                                         dm_dmu1, dm_dsigma1_sq, dm_dsigma12);
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
  ssim_bwd_kernel<<<grid, NT, 0, reinterpret_cast<cudaStream_t>(stream)>>>(H, W, img1, img2, dL_dmap, dm_dmu1,
                                                                         dm_dsigma1_sq, dm_dsigma12, dL_dimg1);
  CG_LAUNCH_CHECK(0, reinterpret_cast<cudaStream_t>(stream));
  return CG_OK;
}

}  // extern "C"
