/*
 * TEST INFRASTRUCTURE ONLY — CPU restatement of simple-knn's distCUDA2:
 * out[i] = mean of the 3 smallest squared distances from point i to the others.
 * The reference's Morton sort + box culling (simple_knn.cu:46-184) only
 * accelerates an EXACT 3-NN search (its cull is conservative), so the oracle is
 * the brute-force definition with the reference's distance arithmetic
 * (simple_knn.cu:133-146: dx*dx + dy*dy + dz*dz, insertion into 3 best) and
 * final (b0+b1+b2)/3 (simple_knn.cu:183).
 * With fewer than 4 points the reference leaves FLT_MAX in the unused slots.
 */
#include <float.h>
#include <math.h>
#include <stdint.h>

void or_knn_mean_dist2(int64_t P, const float* pts, float* out) {
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < P; ++i) {
    float best[3] = {FLT_MAX, FLT_MAX, FLT_MAX};
    const float x = pts[3 * i], y = pts[3 * i + 1], z = pts[3 * i + 2];
    for (int64_t j = 0; j < P; ++j) {
      if (j == i) continue;
      const float dx = pts[3 * j] - x, dy = pts[3 * j + 1] - y, dz = pts[3 * j + 2] - z;
      float d = fmaf(dz, dz, fmaf(dx, dx, dy * dy));
      for (int k = 0; k < 3; ++k)
        if (best[k] > d) { float t = best[k]; best[k] = d; d = t; }
    }
    out[i] = (best[0] + best[1] + best[2]) / 3.0f;
  }
}
