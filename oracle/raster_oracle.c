/*
 * TEST INFRASTRUCTURE ONLY — CPU restatement (plain C, OpenMP) of the reference
 * curve-Gaussian rasterizer. Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load this; the product
 * (libcurvegs.so) never links or calls it.
 *
 * Pinned by: tests/golden/raster_*.npz — outputs of the UNMODIFIED reference CUDA
 * extension (oracle/_ref/diff_cur_rasterization_C.so built by oracle/build_ref.sh)
 * captured on a B200 by tests/golden/make_raster_golden.py; see
 * tests/test_oracle_raster.py.
 *
 * Follows (reference file:line, submodules/diff-cur-rasterization/cuda_rasterizer):
 *   or_preprocess      forward.cu:78-113 (cov2D), :118-152 (cov3D), :155-274; auxiliary.h:40-55,70-89,151-176
 *   or_bin             rasterizer_impl.cu:70-111 (keys), :35-50 + :306-314 (stable sort), :116-138 (ranges)
 *   or_blend_fwd       forward.cu:279-417
 *   or_blend_bwd       backward.cu:451-675
 *   or_preprocess_bwd  backward.cu:146-325, :329-392, :397-448
 *
 * fp32 rounding: compiled with -ffp-contract=off; the places where the reference
 * build fuses a multiply-add (read off its sm_100 SASS) are written as fmaf()
 * so that radii / tile rects / depth bits come out identical on typical inputs.
 * expf() is glibc's, not CUDA's, so pixel values agree to ~1 ulp of alpha, not
 * bit for bit.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define TILE 16

static inline float dot3p(float a0, float b0, float a1, float b1, float a2, float b2) {
  return fmaf(a2, b2, fmaf(a0, b0, a1 * b1));
}
/* row vector times 4x4, matrix element (i,j) at m[4*i+j] */
static inline void xform(const float* m, float x, float y, float z, float* out, int n) {
  for (int j = 0; j < n; ++j) out[j] = dot3p(m[j], x, m[4 + j], y, m[8 + j], z) + m[12 + j];
}
/* column-major 3x3 like glm: a[c][r]; out = a*b with the k=0,1,2 sum order */
static void mul3(const float a[3][3], const float b[3][3], float out[3][3]) {
  for (int c = 0; c < 3; ++c)
    for (int r = 0; r < 3; ++r) out[c][r] = dot3p(a[0][r], b[c][0], a[1][r], b[c][1], a[2][r], b[c][2]);
}
static void tr3(const float a[3][3], float out[3][3]) {
  for (int c = 0; c < 3; ++c)
    for (int r = 0; r < 3; ++r) out[c][r] = a[r][c];
}
static void quat_matrix(const float* q, float R[3][3]) {
  const float r = q[0], x = q[1], y = q[2], z = q[3];
  const float rz = r * z, rx = r * x, xz = x * z, yy = y * y, zz = z * z;
  const float xy_m = fmaf(x, y, -rz), xy_p = fmaf(x, y, rz);
  const float xz_p = fmaf(r, y, xz), xz_m = fmaf(-r, y, xz);
  const float yz_m = fmaf(y, z, -rx), yz_p = fmaf(y, z, rx);
  const float a = yy + zz, b = fmaf(x, x, zz), c = fmaf(x, x, yy);
  R[0][0] = 1.f - (a + a); R[0][1] = xy_m + xy_m; R[0][2] = xz_p + xz_p;
  R[1][0] = xy_p + xy_p;   R[1][1] = 1.f - (b + b); R[1][2] = yz_m + yz_m;
  R[2][0] = xz_m + xz_m;   R[2][1] = yz_p + yz_p;   R[2][2] = 1.f - (c + c);
}
static void cov3d(const float* scale, float mod, const float* q, float* cov6) {
  float R[3][3], M[3][3], Mt[3][3], S[3][3];
  const float s[3] = {mod * scale[0], mod * scale[1], mod * scale[2]};
  quat_matrix(q, R);
  for (int c = 0; c < 3; ++c)
    for (int r = 0; r < 3; ++r) M[c][r] = s[r] * R[c][r];
  tr3(M, Mt);
  mul3(Mt, M, S);
  cov6[0] = S[0][0]; cov6[1] = S[0][1]; cov6[2] = S[0][2];
  cov6[3] = S[1][1]; cov6[4] = S[1][2]; cov6[5] = S[2][2];
}

typedef struct {
  float t[3], txtz, tytz;
  float T[3][3], V[3][3];
  float c00, c01, c11;
} proj_t;

static void project(const float* p, float fx, float fy, float tanx, float tany, const float* cov6,
                    const float* vm, proj_t* o) {
  float t[3];
  xform(vm, p[0], p[1], p[2], t, 3);
  const float limx = 1.3f * tanx, limy = 1.3f * tany;
  o->txtz = t[0] / t[2];
  o->tytz = t[1] / t[2];
  t[0] = fminf(limx, fmaxf(-limx, o->txtz)) * t[2];
  t[1] = fminf(limy, fmaxf(-limy, o->tytz)) * t[2];
  memcpy(o->t, t, sizeof(t));
  const float tz2 = t[2] * t[2];
  float J[3][3] = {{fx / t[2], 0.f, -(fx * t[0]) / tz2}, {0.f, fy / t[2], -(fy * t[1]) / tz2}, {0.f, 0.f, 0.f}};
  float Wm[3][3] = {{vm[0], vm[4], vm[8]}, {vm[1], vm[5], vm[9]}, {vm[2], vm[6], vm[10]}};
  mul3(Wm, J, o->T);
  float V[3][3] = {{cov6[0], cov6[1], cov6[2]}, {cov6[1], cov6[3], cov6[4]}, {cov6[2], cov6[4], cov6[5]}};
  memcpy(o->V, V, sizeof(V));
  float Tt[3][3], Vt[3][3], X[3][3], C[3][3];
  tr3(o->T, Tt);
  tr3(V, Vt);
  mul3(Tt, Vt, X);
  mul3(X, o->T, C);
  o->c00 = C[0][0]; o->c01 = C[0][1]; o->c11 = C[1][1];
}

static inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* Per-Gaussian projection. rect = (min.x, min.y, max.x, max.y) in tiles. Arrays for culled
 * Gaussians keep radii = 0, tiles = 0 and zeros elsewhere. */
void or_preprocess(int64_t P, const float* means, const float* scales, const float* rots, const float* opac,
                   const float* cov3D_precomp, float mod, const float* vm, const float* pm, int W, int H,
                   float tanx, float tany, int antialiasing, int32_t* radii, float* xy, float* depth,
                   float* conic_o, uint32_t* tiles, int32_t* rect, float* cov3D_out) {
  const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
  const float fy = H / (2.0f * tany), fx = W / (2.0f * tanx);
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < P; ++i) {
    radii[i] = 0; tiles[i] = 0;
    xy[2 * i] = xy[2 * i + 1] = 0.f; depth[i] = 0.f;
    for (int k = 0; k < 4; ++k) { conic_o[4 * i + k] = 0.f; rect[4 * i + k] = 0; }
    const float* p = means + 3 * i;
    float pv[3], ph[4];
    xform(vm, p[0], p[1], p[2], pv, 3);
    if (pv[2] <= 0.2f) continue;
    xform(pm, p[0], p[1], p[2], ph, 4);
    const float pw = 1.0f / (ph[3] + 0.0000001f);
    const float ndcx = ph[0] * pw, ndcy = ph[1] * pw;
    float cov6[6];
    if (cov3D_precomp) memcpy(cov6, cov3D_precomp + 6 * i, sizeof(cov6));
    else cov3d(scales + 3 * i, mod, rots + 4 * i, cov6);
    if (cov3D_out) memcpy(cov3D_out + 6 * i, cov6, sizeof(cov6));
    proj_t pr;
    project(p, fx, fy, tanx, tany, cov6, vm, &pr);
    const float bb = pr.c01 * pr.c01;
    const float det_cov = fmaf(pr.c00, pr.c11, -bb);
    const float a = pr.c00 + 0.3f, c = pr.c11 + 0.3f, b = pr.c01;
    const float det = fmaf(a, c, -bb);
    float hs = 1.0f;
    if (antialiasing) hs = sqrtf(fmaxf(0.000025f, det_cov / det));
    if (det == 0.0f) continue;
    const float di = 1.f / det;
    const float mid = (a + c) * 0.5f;
    const float disc = sqrtf(fmaxf(0.1f, fmaf(mid, mid, -det)));
    const float l1 = mid + disc, l2 = mid - disc;
    const float radf = ceilf(sqrtf(fmaxf(l1, l2)) * 3.f);
    const float px = (float)(fma((double)ndcx + 1.0, (double)W, -1.0) * 0.5);
    const float py = (float)(fma((double)ndcy + 1.0, (double)H, -1.0) * 0.5);
    const int rad = (int)radf;
    const float rf = (float)rad;
    const int mnx = clampi((int)((px - rf) / TILE), 0, gx), mny = clampi((int)((py - rf) / TILE), 0, gy);
    const int mxx = clampi((int)((((px + rf) + TILE) - 1) / TILE), 0, gx);
    const int mxy = clampi((int)((((py + rf) + TILE) - 1) / TILE), 0, gy);
    const uint32_t area = (uint32_t)(mxx - mnx) * (uint32_t)(mxy - mny);
    if (area == 0) continue;
    depth[i] = pv[2];
    radii[i] = rad;
    xy[2 * i] = px; xy[2 * i + 1] = py;
    conic_o[4 * i] = c * di; conic_o[4 * i + 1] = b * -di; conic_o[4 * i + 2] = a * di;
    conic_o[4 * i + 3] = opac[i] * hs;
    tiles[i] = area;
    rect[4 * i] = mnx; rect[4 * i + 1] = mny; rect[4 * i + 2] = mxx; rect[4 * i + 3] = mxy;
  }
}

static uint32_t key_tile_bits(uint32_t n) { /* == getHigherMsb for n >= 1 */
  uint32_t b = 0;
  while (n) { ++b; n >>= 1; }
  return b ? b : 1;
}
/* The halving search of the reference, restated only so tests can check the claim above. */
uint32_t or_higher_msb(uint32_t n) {
  uint32_t pos = 16, step = 16;
  while (step > 1) {
    step >>= 1;
    pos = (n >> pos) ? pos + step : pos - step;
  }
  return (n >> pos) ? pos + 1 : pos;
}
uint32_t or_key_tile_bits(uint32_t n) { return key_tile_bits(n); }

/* Emits keys in Gaussian order / row-major rect order, sorts them with a stable LSD radix
 * sort over the low 32 + tile bits, and derives the per-tile ranges. keys/vals hold R entries
 * (R = sum of tiles); ranges holds 2*gx*gy uint32 (zero for empty tiles). */
int64_t or_bin(int64_t P, const uint32_t* tiles, const int32_t* rect, const float* depth, int W, int H,
               uint64_t* keys, uint32_t* vals, uint32_t* ranges) {
  const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
  int64_t R = 0;
  for (int64_t i = 0; i < P; ++i) {
    if (!tiles[i]) continue;
    uint32_t dbits;
    memcpy(&dbits, depth + i, 4);
    for (int y = rect[4 * i + 1]; y < rect[4 * i + 3]; ++y)
      for (int x = rect[4 * i]; x < rect[4 * i + 2]; ++x) {
        keys[R] = ((uint64_t)((uint32_t)y * (uint32_t)gx + (uint32_t)x) << 32) | dbits;
        vals[R] = (uint32_t)i;
        ++R;
      }
  }
  const int bits = 32 + (int)key_tile_bits((uint32_t)(gx * gy));
  uint64_t* k2 = (uint64_t*)malloc(sizeof(uint64_t) * (size_t)(R ? R : 1));
  uint32_t* v2 = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)(R ? R : 1));
  uint64_t *ka = keys, *kb = k2;
  uint32_t *va = vals, *vb = v2;
  for (int shift = 0; shift < bits; shift += 8) {
    size_t cnt[257] = {0};
    for (int64_t i = 0; i < R; ++i) cnt[((ka[i] >> shift) & 255u) + 1]++;
    for (int d = 0; d < 256; ++d) cnt[d + 1] += cnt[d];
    for (int64_t i = 0; i < R; ++i) {
      size_t pos = cnt[(ka[i] >> shift) & 255u]++;
      kb[pos] = ka[i];
      vb[pos] = va[i];
    }
    uint64_t* tk = ka; ka = kb; kb = tk;
    uint32_t* tv = va; va = vb; vb = tv;
  }
  if (ka != keys) { memcpy(keys, ka, sizeof(uint64_t) * (size_t)R); memcpy(vals, va, sizeof(uint32_t) * (size_t)R); }
  free(k2); free(v2);
  memset(ranges, 0, sizeof(uint32_t) * 2 * (size_t)gx * gy);
  for (int64_t i = 0; i < R; ++i) {
    const uint32_t cur = (uint32_t)(keys[i] >> 32);
    if (i == 0) ranges[2 * cur] = 0;
    else {
      const uint32_t prev = (uint32_t)(keys[i - 1] >> 32);
      if (cur != prev) { ranges[2 * prev + 1] = (uint32_t)i; ranges[2 * cur] = (uint32_t)i; }
    }
    if (i == R - 1) ranges[2 * cur + 1] = (uint32_t)R;
  }
  return R;
}

/* Front-to-back blend of one channel + inverse depth + 4 map channels. */
void or_blend_fwd(int W, int H, const uint32_t* ranges, const uint32_t* point_list, const float* xy,
                  const float* conic_o, const float* colors, const float* depth, const float* all_map,
                  int render_geo, const float* bg, float* out_color, float* out_invd, float* out_map,
                  float* final_T, uint32_t* n_contrib) {
  const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
  const size_t hw = (size_t)W * H;
#pragma omp parallel for schedule(dynamic, 4)
  for (int tile = 0; tile < gx * gy; ++tile) {
    const int tx = tile % gx, ty = tile / gx;
    const uint32_t lo = ranges[2 * tile], hi = ranges[2 * tile + 1];
    for (int ly = 0; ly < TILE; ++ly)
      for (int lx = 0; lx < TILE; ++lx) {
        const int px = tx * TILE + lx, py = ty * TILE + ly;
        if (px >= W || py >= H) continue;
        const float pxf = (float)px, pyf = (float)py;
        float T = 1.f, C = 0.f, D = 0.f, M[4] = {0.f, 0.f, 0.f, 0.f};
        uint32_t contributor = 0, last = 0;
        for (uint32_t k = lo; k < hi; ++k) {
          ++contributor;
          const uint32_t id = point_list[k];
          const float dx = xy[2 * id] - pxf, dy = xy[2 * id + 1] - pyf;
          const float* co = conic_o + 4 * id;
          /* contraction as in the reference build: see blend kernels' SASS */
          const float power = fmaf(fmaf(dx, co[0] * dx, (co[2] * dy) * dy), -0.5f, -((co[1] * dx) * dy));
          if (power > 0.f) continue;
          const float alpha = fminf(0.99f, co[3] * expf(power));
          if (alpha < 1.0f / 255.0f) continue;
          const float test_T = T * (1 - alpha);
          if (test_T < 0.0001f) break;
          C = fmaf(colors[id] * alpha, T, C);
          D = fmaf((1 / depth[id]) * alpha, T, D);
          if (render_geo)
            for (int ch = 0; ch < 4; ++ch) M[ch] = fmaf(all_map[4 * id + ch] * alpha, T, M[ch]);
          T = test_T;
          last = contributor;
        }
        const size_t pid = (size_t)py * W + px;
        final_T[pid] = T;
        n_contrib[pid] = last;
        out_color[pid] = fmaf(T, bg[0], C);
        out_invd[pid] = D;
        if (render_geo)
          for (int ch = 0; ch < 4; ++ch) out_map[ch * hw + pid] = M[ch];
      }
  }
}

static inline void atomic_addf(float* p, float v) {
#pragma omp atomic
  *p += v;
}

/* Back-to-front adjoint of the blend. dL_dinvd / dL_dmap may be NULL (zeros).
 * Accumulators (caller zeroes): d_mean2D (P,3), d_conic (P,4; x,y,w used), d_opacity (P),
 * d_colors (P), d_invdepths (P), d_all_map (P,4). */
void or_blend_bwd(int W, int H, const uint32_t* ranges, const uint32_t* point_list, const float* xy,
                  const float* conic_o, const float* colors, const float* depth, const float* all_map,
                  int render_geo, const float* bg, const float* final_T, const uint32_t* n_contrib,
                  const float* dL_dpix, const float* dL_dinvd, const float* dL_dmap, float* d_mean2D,
                  float* d_conic, float* d_opacity, float* d_colors, float* d_invdepths, float* d_all_map) {
  const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
  const size_t hw = (size_t)W * H;
  const float ddelx_dx = 0.5f * W, ddely_dy = 0.5f * H;
#pragma omp parallel for schedule(dynamic, 4)
  for (int tile = 0; tile < gx * gy; ++tile) {
    const int tx = tile % gx, ty = tile / gx;
    const uint32_t lo = ranges[2 * tile];
    for (int ly = 0; ly < TILE; ++ly)
      for (int lx = 0; lx < TILE; ++lx) {
        const int px = tx * TILE + lx, py = ty * TILE + ly;
        if (px >= W || py >= H) continue;
        const size_t pid = (size_t)py * W + px;
        const float pxf = (float)px, pyf = (float)py;
        const float T_final = final_T[pid];
        float T = T_final;
        const float gC = dL_dpix[pid];
        const float gD = dL_dinvd ? dL_dinvd[pid] : 0.f;
        float gM[4] = {0.f, 0.f, 0.f, 0.f};
        if (render_geo && dL_dmap)
          for (int ch = 0; ch < 4; ++ch) gM[ch] = dL_dmap[ch * hw + pid];
        float last_alpha = 0.f, last_c = 0.f, rec_c = 0.f, last_d = 0.f, rec_d = 0.f;
        float last_m[4] = {0.f, 0.f, 0.f, 0.f}, rec_m[4] = {0.f, 0.f, 0.f, 0.f};
        const float bg_dot = bg[0] * gC;
        for (int64_t pos = (int64_t)n_contrib[pid] - 1; pos >= 0; --pos) {
          const uint32_t id = point_list[lo + pos];
          const float dx = xy[2 * id] - pxf, dy = xy[2 * id + 1] - pyf;
          const float* co = conic_o + 4 * id;
          const float power = fmaf(fmaf(dx, co[0] * dx, (co[2] * dy) * dy), -0.5f, -((co[1] * dx) * dy));
          if (power > 0.f) continue;
          const float G = expf(power);
          const float alpha = fminf(0.99f, co[3] * G);
          if (alpha < 1.0f / 255.0f) continue;
          T = T / (1.f - alpha);
          const float w = alpha * T;
          float dL_dalpha = 0.f;
          const float c = colors[id];
          rec_c = last_alpha * last_c + (1.f - last_alpha) * rec_c;
          last_c = c;
          dL_dalpha += (c - rec_c) * gC;
          atomic_addf(d_colors + id, w * gC);
          if (dL_dinvd) {
            const float invd = 1.f / depth[id];
            rec_d = last_alpha * last_d + (1.f - last_alpha) * rec_d;
            last_d = invd;
            dL_dalpha += (invd - rec_d) * gD;
            atomic_addf(d_invdepths + id, w * gD);
          }
          if (render_geo && dL_dmap)
            for (int ch = 0; ch < 4; ++ch) {
              const float m = all_map[4 * id + ch];
              rec_m[ch] = last_alpha * last_m[ch] + (1.f - last_alpha) * rec_m[ch];
              last_m[ch] = m;
              dL_dalpha += (m - rec_m[ch]) * gM[ch];
              atomic_addf(d_all_map + 4 * id + ch, w * gM[ch]);
            }
          dL_dalpha *= T;
          last_alpha = alpha;
          dL_dalpha += (-T_final / (1.f - alpha)) * bg_dot;
          const float dL_dG = co[3] * dL_dalpha;
          const float gdx = G * dx, gdy = G * dy;
          const float dG_ddelx = -gdx * co[0] - gdy * co[1];
          const float dG_ddely = -gdy * co[2] - gdx * co[1];
          atomic_addf(d_mean2D + 3 * id, dL_dG * dG_ddelx * ddelx_dx);
          atomic_addf(d_mean2D + 3 * id + 1, dL_dG * dG_ddely * ddely_dy);
          atomic_addf(d_conic + 4 * id, -0.5f * gdx * dx * dL_dG);
          atomic_addf(d_conic + 4 * id + 1, -0.5f * gdx * dy * dL_dG);
          atomic_addf(d_conic + 4 * id + 3, -0.5f * gdy * dy * dL_dG);
          atomic_addf(d_opacity + id, G * dL_dalpha);
        }
      }
  }
}

/* conic / mean2D / inverse-depth adjoints -> means3D, cov3D, scales, raw quaternion.
 * d_opacity is updated in place when antialiasing. d_invdepths may be NULL. */
void or_preprocess_bwd(int64_t P, const float* means, const float* scales, const float* rots, const float* opac,
                       const float* cov3D_precomp, float mod, const int32_t* radii, const float* vm,
                       const float* pm, int W, int H, float tanx, float tany, int antialiasing,
                       const float* d_mean2D, const float* d_conic, const float* d_invdepths, float* d_opacity,
                       float* d_means3D, float* d_cov3D, float* d_scales, float* d_rots) {
  const float fy = H / (2.0f * tany), fx = W / (2.0f * tanx);
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < P; ++i) {
    for (int k = 0; k < 3; ++k) d_means3D[3 * i + k] = 0.f;
    for (int k = 0; k < 6; ++k) d_cov3D[6 * i + k] = 0.f;
    if (d_scales) for (int k = 0; k < 3; ++k) d_scales[3 * i + k] = 0.f;
    if (d_rots) for (int k = 0; k < 4; ++k) d_rots[4 * i + k] = 0.f;
    if (!(radii[i] > 0)) continue;
    const float* p = means + 3 * i;
    float cov6[6];
    if (cov3D_precomp) memcpy(cov6, cov3D_precomp + 6 * i, sizeof(cov6));
    else cov3d(scales + 3 * i, mod, rots + 4 * i, cov6);
    proj_t pr;
    project(p, fx, fy, tanx, tany, cov6, vm, &pr);
    const float limx = 1.3f * tanx, limy = 1.3f * tany;
    const float xmul = (pr.txtz < -limx || pr.txtz > limx) ? 0.f : 1.f;
    const float ymul = (pr.tytz < -limy || pr.tytz > limy) ? 0.f : 1.f;
    const float dA = d_conic[4 * i], dB = d_conic[4 * i + 1], dC = d_conic[4 * i + 3];
    float a = pr.c00, b = pr.c01, c = pr.c11;
    float d_root = 0.f;
    if (antialiasing) {
      const float det_cov = a * c - b * b;
      a += 0.3f; c += 0.3f;
      const float det_h = a * c - b * b;
      const float hs = sqrtf(fmaxf(0.000025f, det_cov / det_h));
      const float g = d_opacity[i];
      const float dh = g * opac[i];
      d_opacity[i] = g * hs;
      d_root = (det_cov / det_h) <= 0.000025f ? 0.f : dh / (2 * hs);
    } else { a += 0.3f; c += 0.3f; }
    float g_a = 0.f, g_b = 0.f, g_c = 0.f;
    if (antialiasing) {
      const float w = 0.3f, q = w * w + w * (a + c) + a * c - b * b, df = d_root / (q * q);
      g_a = w * (w * c + c * c + b * b) * df;
      g_c = w * (w * a + a * a + b * b) * df;
      g_b = -2.f * w * b * (w + a + c) * df;
    }
    const float D = a * c - b * b;
    const float k2 = 1.0f / ((D * D) + 0.0000001f);
    const float (*T)[3] = pr.T;
    const float (*V)[3] = pr.V;
    float* oc = d_cov3D + 6 * i;
    if (k2 != 0) {
      g_a += k2 * (-c * c * dA + 2 * b * c * dB + (D - a * c) * dC);
      g_c += k2 * (-a * a * dC + 2 * a * b * dB + (D - a * c) * dA);
      g_b += k2 * 2 * (b * c * dA - (D + 2 * b * b) * dB + a * b * dC);
      oc[0] = T[0][0] * T[0][0] * g_a + T[0][0] * T[1][0] * g_b + T[1][0] * T[1][0] * g_c;
      oc[3] = T[0][1] * T[0][1] * g_a + T[0][1] * T[1][1] * g_b + T[1][1] * T[1][1] * g_c;
      oc[5] = T[0][2] * T[0][2] * g_a + T[0][2] * T[1][2] * g_b + T[1][2] * T[1][2] * g_c;
      oc[1] = 2 * T[0][0] * T[0][1] * g_a + (T[0][0] * T[1][1] + T[0][1] * T[1][0]) * g_b + 2 * T[1][0] * T[1][1] * g_c;
      oc[2] = 2 * T[0][0] * T[0][2] * g_a + (T[0][0] * T[1][2] + T[0][2] * T[1][0]) * g_b + 2 * T[1][0] * T[1][2] * g_c;
      oc[4] = 2 * T[0][2] * T[0][1] * g_a + (T[0][1] * T[1][2] + T[0][2] * T[1][1]) * g_b + 2 * T[1][1] * T[1][2] * g_c;
    }
    float dT0[3], dT1[3];
    for (int cc = 0; cc < 3; ++cc) {
      const float r0 = T[0][0] * V[cc][0] + T[0][1] * V[cc][1] + T[0][2] * V[cc][2];
      const float r1 = T[1][0] * V[cc][0] + T[1][1] * V[cc][1] + T[1][2] * V[cc][2];
      dT0[cc] = 2 * r0 * g_a + r1 * g_b;
      dT1[cc] = 2 * r1 * g_c + r0 * g_b;
    }
    const float dJ00 = vm[0] * dT0[0] + vm[4] * dT0[1] + vm[8] * dT0[2];
    const float dJ02 = vm[2] * dT0[0] + vm[6] * dT0[1] + vm[10] * dT0[2];
    const float dJ11 = vm[1] * dT1[0] + vm[5] * dT1[1] + vm[9] * dT1[2];
    const float dJ12 = vm[2] * dT1[0] + vm[6] * dT1[1] + vm[10] * dT1[2];
    const float* t = pr.t;
    const float tz = 1.f / t[2], tz2 = tz * tz, tz3 = tz2 * tz;
    const float dtx = xmul * -fx * tz2 * dJ02;
    const float dty = ymul * -fy * tz2 * dJ12;
    float dtz = -fx * tz2 * dJ00 - fy * tz2 * dJ11 + (2 * fx * t[0]) * tz3 * dJ02 + (2 * fy * t[1]) * tz3 * dJ12;
    if (d_invdepths) dtz -= d_invdepths[i] / (t[2] * t[2]);
    float dm[3];
    dm[0] = vm[0] * dtx + vm[1] * dty + vm[2] * dtz;
    dm[1] = vm[4] * dtx + vm[5] * dty + vm[6] * dtz;
    dm[2] = vm[8] * dtx + vm[9] * dty + vm[10] * dtz;
    float mh[4];
    xform(pm, p[0], p[1], p[2], mh, 4);
    const float mw = 1.0f / (mh[3] + 0.0000001f);
    const float mul1 = (pm[0] * p[0] + pm[4] * p[1] + pm[8] * p[2] + pm[12]) * mw * mw;
    const float mul2 = (pm[1] * p[0] + pm[5] * p[1] + pm[9] * p[2] + pm[13]) * mw * mw;
    const float gx2 = d_mean2D[3 * i], gy2 = d_mean2D[3 * i + 1];
    for (int k = 0; k < 3; ++k)
      d_means3D[3 * i + k] = dm[k] + ((pm[4 * k] * mw - pm[4 * k + 3] * mul1) * gx2 +
                                      (pm[4 * k + 1] * mw - pm[4 * k + 3] * mul2) * gy2);
    if (!cov3D_precomp && d_scales && d_rots) {
      float R[3][3], M[3][3], dS[3][3], dM[3][3];
      const float* q = rots + 4 * i;
      const float s[3] = {mod * scales[3 * i], mod * scales[3 * i + 1], mod * scales[3 * i + 2]};
      quat_matrix(q, R);
      for (int cc = 0; cc < 3; ++cc)
        for (int r = 0; r < 3; ++r) M[cc][r] = s[r] * R[cc][r];
      dS[0][0] = oc[0]; dS[1][1] = oc[3]; dS[2][2] = oc[5];
      dS[0][1] = dS[1][0] = 0.5f * oc[1];
      dS[0][2] = dS[2][0] = 0.5f * oc[2];
      dS[1][2] = dS[2][1] = 0.5f * oc[4];
      for (int cc = 0; cc < 3; ++cc)
        for (int r = 0; r < 3; ++r)
          dM[cc][r] = 2.0f * (M[0][r] * dS[cc][0] + M[1][r] * dS[cc][1] + M[2][r] * dS[cc][2]);
      /* dMt[c][r] = dM[r][c], Rt[c][r] = R[r][c] */
      float dMt[3][3];
      for (int cc = 0; cc < 3; ++cc)
        for (int r = 0; r < 3; ++r) dMt[cc][r] = dM[r][cc];
      for (int k = 0; k < 3; ++k)
        d_scales[3 * i + k] = R[0][k] * dMt[k][0] + R[1][k] * dMt[k][1] + R[2][k] * dMt[k][2];
      for (int k = 0; k < 3; ++k)
        for (int r = 0; r < 3; ++r) dMt[k][r] *= s[k];
      const float r = q[0], x = q[1], y = q[2], z = q[3];
      float* o = d_rots + 4 * i;
      o[0] = 2 * z * (dMt[0][1] - dMt[1][0]) + 2 * y * (dMt[2][0] - dMt[0][2]) + 2 * x * (dMt[1][2] - dMt[2][1]);
      o[1] = 2 * y * (dMt[1][0] + dMt[0][1]) + 2 * z * (dMt[2][0] + dMt[0][2]) + 2 * r * (dMt[1][2] - dMt[2][1]) - 4 * x * (dMt[2][2] + dMt[1][1]);
      o[2] = 2 * x * (dMt[1][0] + dMt[0][1]) + 2 * r * (dMt[2][0] - dMt[0][2]) + 2 * z * (dMt[1][2] + dMt[2][1]) - 4 * y * (dMt[2][2] + dMt[0][0]);
      o[3] = 2 * r * (dMt[0][1] - dMt[1][0]) + 2 * x * (dMt[2][0] + dMt[0][2]) + 2 * y * (dMt[1][2] + dMt[2][1]) - 4 * z * (dMt[1][1] + dMt[0][0]);
    }
  }
}

int or_num_threads(void) {
  int n = 1;
#ifdef _OPENMP
#pragma omp parallel
  {
#pragma omp master
    n = omp_get_num_threads();
  }
#endif
  return n;
}
