/*
 * TEST INFRASTRUCTURE ONLY — CPU restatement of fused SSIM (value map, the three
 * partial-derivative maps, and the image gradient).
 * Follows submodules/fused-ssim/ssim.cu:187-366 (== utils/loss_utils.py:46-86 for
 * the value): separable 11-tap Gaussian (sigma 1.5, taps ssim.cu:9-19), zero
 * padding, x pass then y pass, C1/C2 passed by the caller.
 * Pinned by the reference's own known-answer relation (fused-ssim/tests/test.py:57-91:
 * fused == pure-torch ssim, value and gradient) in tests/test_oracle_ssim.py.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

static const float TAP[11] = {0.001028380123898387f, 0.0075987582094967365f, 0.036000773310661316f,
                              0.10936068743467331f,  0.21300552785396576f,   0.26601171493530273f,
                              0.21300552785396576f,  0.10936068743467331f,   0.036000773310661316f,
                              0.0075987582094967365f, 0.001028380123898387f};

/* out = separable conv of `in` (H,W), zero outside the image */
static void conv2(const float* in, float* tmp, float* out, int H, int W) {
#pragma omp parallel for schedule(static)
  for (int y = 0; y < H; ++y)
    for (int x = 0; x < W; ++x) {
      float v = 0.f;
      for (int k = 0; k < 11; ++k) {
        const int xx = x + k - 5;
        v = fmaf(TAP[k], (xx >= 0 && xx < W) ? in[(size_t)y * W + xx] : 0.f, v);
      }
      tmp[(size_t)y * W + x] = v;
    }
#pragma omp parallel for schedule(static)
  for (int y = 0; y < H; ++y)
    for (int x = 0; x < W; ++x) {
      float v = 0.f;
      for (int k = 0; k < 11; ++k) {
        const int yy = y + k - 5;
        v = fmaf(TAP[k], (yy >= 0 && yy < H) ? tmp[(size_t)yy * W + x] : 0.f, v);
      }
      out[(size_t)y * W + x] = v;
    }
}

void or_ssim_fwd(int B, int CH, int H, int W, float C1, float C2, const float* img1, const float* img2,
                 float* ssim_map, float* dm_dmu1, float* dm_dsigma1_sq, float* dm_dsigma12) {
  const size_t n = (size_t)H * W;
  float* buf = (float*)malloc(sizeof(float) * n * 7);
  float *t = buf, *a = buf + n, *mu1 = buf + 2 * n, *mu2 = buf + 3 * n, *e11 = buf + 4 * n, *e22 = buf + 5 * n,
        *e12 = buf + 6 * n;
  for (int p = 0; p < B * CH; ++p) {
    const float* x = img1 + p * n;
    const float* y = img2 + p * n;
    conv2(x, t, mu1, H, W);
    conv2(y, t, mu2, H, W);
    for (size_t i = 0; i < n; ++i) a[i] = x[i] * x[i];
    conv2(a, t, e11, H, W);
    for (size_t i = 0; i < n; ++i) a[i] = y[i] * y[i];
    conv2(a, t, e22, H, W);
    for (size_t i = 0; i < n; ++i) a[i] = x[i] * y[i];
    conv2(a, t, e12, H, W);
    for (size_t i = 0; i < n; ++i) {
      const float m1 = mu1[i], m2 = mu2[i];
      const float s1 = e11[i] - m1 * m1, s2 = e22[i] - m2 * m2, s12 = e12[i] - m1 * m2;
      const float Cc = 2.0f * (m1 * m2) + C1, D = 2.0f * s12 + C2;
      const float A = m1 * m1 + m2 * m2 + C1, Bq = s1 + s2 + C2;
      ssim_map[p * n + i] = (Cc * D) / (A * Bq);
      if (dm_dmu1) {
        dm_dmu1[p * n + i] = (m2 * 2.0f * D) / (A * Bq) - (m2 * 2.0f * Cc) / (A * Bq) -
                             (m1 * 2.0f * Cc * D) / (A * A * Bq) + (m1 * 2.0f * Cc * D) / (A * Bq * Bq);
        dm_dsigma1_sq[p * n + i] = (-Cc * D) / (A * Bq * Bq);
        dm_dsigma12[p * n + i] = (2 * Cc) / (A * Bq);
      }
    }
  }
  free(buf);
}

void or_ssim_bwd(int B, int CH, int H, int W, const float* img1, const float* img2, const float* dL_dmap,
                 const float* dm_dmu1, const float* dm_dsigma1_sq, const float* dm_dsigma12, float* dL_dimg1) {
  const size_t n = (size_t)H * W;
  float* buf = (float*)malloc(sizeof(float) * n * 3);
  float *t = buf, *a = buf + n, *o = buf + 2 * n;
  for (int p = 0; p < B * CH; ++p) {
    const size_t off = p * n;
    for (size_t i = 0; i < n; ++i) a[i] = dm_dmu1[off + i] * dL_dmap[off + i];
    conv2(a, t, o, H, W);
    for (size_t i = 0; i < n; ++i) dL_dimg1[off + i] = o[i];
    for (size_t i = 0; i < n; ++i) a[i] = dm_dsigma1_sq[off + i] * dL_dmap[off + i];
    conv2(a, t, o, H, W);
    for (size_t i = 0; i < n; ++i) dL_dimg1[off + i] += img1[off + i] * 2.0f * o[i];
    for (size_t i = 0; i < n; ++i) a[i] = dm_dsigma12[off + i] * dL_dmap[off + i];
    conv2(a, t, o, H, W);
    for (size_t i = 0; i < n; ++i) dL_dimg1[off + i] += img2[off + i] * o[i];
  }
  free(buf);
}
