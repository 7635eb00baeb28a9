// Curve control points -> per-Gaussian mean / un-normalised quaternion / scale,
// and the full adjoint back to the control points and widths.
//
// Reference behaviour (Python/ATen, ~50 launches + an autograd graph):
//   scene/gaussian_curve_model.py:58-60 (sample_t), :70-89 (curve point / tangent),
//   :180-198 (prepare_scaling_rot, incl. the two WHOLE-TENSOR norms at :190,:192),
//   utils/general_utils.py:9-86 (rot_to_quat_batch: argmax branch, w >= 0, zero
//   sub-gradient sqrt, 0.1 floor).
// Here: forward = one grid reduction (two sums) + one pass; backward = ONE pass over
// the Gaussians that accumulates per-curve partial sums and the four global sums the
// two whole-tensor norms need, plus a per-curve finalisation (see sample_bwd_partial).
#include "common.cuh"

namespace cg {

struct V3 { float x, y, z; };
__device__ __forceinline__ V3 v3(float x, float y, float z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
__device__ __forceinline__ V3 operator+(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ V3 operator-(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ V3 operator*(float s, V3 a) { return v3(s * a.x, s * a.y, s * a.z); }
__device__ __forceinline__ float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ V3 cross(V3 a, V3 b) { return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }

struct Curve { V3 p0, p1, p2, p3; bool bez; };

__device__ __forceinline__ Curve load_curve(const float* __restrict__ cp, const uint8_t* __restrict__ is_bezier, int64_t b) {
  const float* c = cp + b * 12;
  Curve k;
  k.p0 = v3(c[0], c[1], c[2]); k.p1 = v3(c[3], c[4], c[5]);
  k.p2 = v3(c[6], c[7], c[8]); k.p3 = v3(c[9], c[10], c[11]);
  k.bez = is_bezier ? (is_bezier[b] != 0) : true;
  return k;
}
// Bernstein weights and the weighted sum, op for op in the reference's order (:71-73, :75) and
// WITHOUT fused multiply-adds: the half-step distance |B(t) - B(t - 0.5/n)| cancels ~2 digits,
// so a contraction the ATen kernels do not make would show up at 1e-5 in scaling[:,0].
__device__ __forceinline__ void point_weights(bool bez, float t, float w[4]) {
  const float u = __fsub_rn(1.f, t);
  if (bez) {
    w[0] = __fmul_rn(__fmul_rn(u, u), u);
    w[1] = __fmul_rn(__fmul_rn(3.f, __fmul_rn(u, u)), t);
    w[2] = __fmul_rn(__fmul_rn(3.f, u), __fmul_rn(t, t));
    w[3] = __fmul_rn(__fmul_rn(t, t), t);
  } else { w[0] = u; w[1] = 0.f; w[2] = 0.f; w[3] = t; }
}
__device__ __forceinline__ V3 mul_rn(float s, V3 a) { return v3(__fmul_rn(s, a.x), __fmul_rn(s, a.y), __fmul_rn(s, a.z)); }
__device__ __forceinline__ V3 add_rn(V3 a, V3 b) { return v3(__fadd_rn(a.x, b.x), __fadd_rn(a.y, b.y), __fadd_rn(a.z, b.z)); }
__device__ __forceinline__ V3 curve_point(const Curve& k, float t) {
  float w[4];
  point_weights(k.bez, t, w);
  if (k.bez) return add_rn(add_rn(add_rn(mul_rn(w[0], k.p0), mul_rn(w[1], k.p1)), mul_rn(w[2], k.p2)), mul_rn(w[3], k.p3));
  return add_rn(mul_rn(w[0], k.p0), mul_rn(w[3], k.p3));
}
// tangent = a*(P1-P0) + b*(P2-P1) + c*(P3-P2)  (:81-83); lines: P3-P0 (:86)
__device__ __forceinline__ void tangent_weights(float t, float w[3]) {
  const float u = 1.f - t;
  w[0] = 3.f * (u * u); w[1] = 6.f * u * t; w[2] = 3.f * (t * t);
}
__device__ __forceinline__ V3 curve_tangent(const Curve& k, float t) {
  if (!k.bez) return k.p3 - k.p0;
  float w[3];
  tangent_weights(t, w);
  return w[0] * (k.p1 - k.p0) + w[1] * (k.p2 - k.p1) + w[2] * (k.p3 - k.p2);
}

struct Frame {
  V3 tau, v0, a, c;   // a = tau x (0,0,1), c = tau x a  (both un-normalised)
  float len;
};
__device__ __forceinline__ Frame make_frame(V3 tau) {
  Frame f;
  f.tau = tau;
  f.len = sqrtf(dot(tau, tau));
  f.v0 = (1.f / (f.len + 1e-8f)) * tau;
  f.a = v3(tau.y, -tau.x, 0.f);
  f.c = cross(tau, f.a);
  return f;
}

struct Quat { float q[4]; int k; float sign; float qa; float s; };
// m[r][c]; returns the chosen candidate (general_utils.py:33-86)
__device__ __forceinline__ Quat matrix_to_quat(const float m[3][3]) {
  const float s[4] = {1.0f + m[0][0] + m[1][1] + m[2][2], 1.0f + m[0][0] - m[1][1] - m[2][2],
                      1.0f - m[0][0] + m[1][1] - m[2][2], 1.0f - m[0][0] - m[1][1] + m[2][2]};
  float qa[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) qa[i] = s[i] > 0.f ? sqrtf(s[i]) : 0.f;
  int k = 0;
#pragma unroll
  for (int i = 1; i < 4; ++i) if (qa[i] > qa[k]) k = i;   // first maximum wins, like argmax
  float row[4];
  if (k == 0) { row[0] = qa[0] * qa[0]; row[1] = m[2][1] - m[1][2]; row[2] = m[0][2] - m[2][0]; row[3] = m[1][0] - m[0][1]; }
  else if (k == 1) { row[0] = m[2][1] - m[1][2]; row[1] = qa[1] * qa[1]; row[2] = m[1][0] + m[0][1]; row[3] = m[0][2] + m[2][0]; }
  else if (k == 2) { row[0] = m[0][2] - m[2][0]; row[1] = m[1][0] + m[0][1]; row[2] = qa[2] * qa[2]; row[3] = m[1][2] + m[2][1]; }
  else { row[0] = m[1][0] - m[0][1]; row[1] = m[2][0] + m[0][2]; row[2] = m[2][1] + m[1][2]; row[3] = qa[3] * qa[3]; }
  Quat o;
  o.k = k; o.qa = qa[k]; o.s = s[k];
  const float den = 2.0f * fmaxf(qa[k], 0.1f);
#pragma unroll
  for (int i = 0; i < 4; ++i) o.q[i] = row[i] / den;
  o.sign = o.q[0] < 0.f ? -1.f : 1.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) o.q[i] *= o.sign;
  return o;
}
// adjoint of matrix_to_quat: gq (dL/dq out) -> gm[r][c]
__device__ __forceinline__ void matrix_to_quat_bwd(const float m[3][3], const Quat& f, const float gq_in[4], float gm[3][3]) {
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) gm[r][c] = 0.f;
  const int k = f.k;
  const float den = 2.0f * fmaxf(f.qa, 0.1f);
  float gq[4], num[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) { gq[i] = gq_in[i] * f.sign; num[i] = f.q[i] * f.sign * den; }
  float gnum[4];
  float gden = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) { gnum[i] = gq[i] / den; gden -= gq[i] * num[i] / (den * den); }
  // diagonal term: num[k] = qa^2, den = 2*max(qa, 0.1)
  float gqa = 2.f * f.qa * gnum[k];
  if (f.qa > 0.1f) gqa += 2.f * gden;
  else if (f.qa == 0.1f) gqa += gden;   // torch.maximum splits the gradient on ties
  const float gs = f.s > 0.f ? gqa / (2.f * f.qa) : 0.f;
  const float sg[4][3] = {{1.f, 1.f, 1.f}, {1.f, -1.f, -1.f}, {-1.f, 1.f, -1.f}, {-1.f, -1.f, 1.f}};
  gm[0][0] += sg[k][0] * gs; gm[1][1] += sg[k][1] * gs; gm[2][2] += sg[k][2] * gs;
  // off-diagonal combinations
  auto add = [&](int r1, int c1, float s1, int r2, int c2, float s2, float g) { gm[r1][c1] += s1 * g; gm[r2][c2] += s2 * g; };
  if (k == 0) { add(2, 1, 1, 1, 2, -1, gnum[1]); add(0, 2, 1, 2, 0, -1, gnum[2]); add(1, 0, 1, 0, 1, -1, gnum[3]); }
  else if (k == 1) { add(2, 1, 1, 1, 2, -1, gnum[0]); add(1, 0, 1, 0, 1, 1, gnum[2]); add(0, 2, 1, 2, 0, 1, gnum[3]); }
  else if (k == 2) { add(0, 2, 1, 2, 0, -1, gnum[0]); add(1, 0, 1, 0, 1, 1, gnum[1]); add(1, 2, 1, 2, 1, 1, gnum[3]); }
  else { add(1, 0, 1, 0, 1, -1, gnum[0]); add(2, 0, 1, 0, 2, 1, gnum[1]); add(2, 1, 1, 1, 2, 1, gnum[2]); }
}

__device__ __forceinline__ double block_sum(double v, double* sm) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) sm[w] = v;
  __syncthreads();
  double t = 0.0;
  if (threadIdx.x < 32) {
    t = (threadIdx.x < (blockDim.x >> 5)) ? sm[threadIdx.x] : 0.0;
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  }
  __syncthreads();
  return t;  // valid in thread 0
}

// Four fp64 block sums at once, added to sums[0..3]. A plain butterfly moves every double through five shuffle
// steps (two SHFLs + packing each: a quarter of sample_bwd_point's instructions); here the first two steps halve
// the number of values a lane still carries (lanes keep sums {0,1} or {2,3}, then one of the pair), so 12 SHFLs and 6
// DADDs per warp instead of 40 and 20. sm: 8 x 4 doubles.
__device__ __forceinline__ void block_sum4_add(double (&v)[4], double* sm, double* __restrict__ sums) {
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
  // lanes 0, 8, 16, 24 now hold the warp's sums 0, 1, 2, 3
  if ((lane & 7u) == 0u) sm[warp * 4 + (lane >> 3)] = k;
  __syncthreads();
  if (threadIdx.x < 4) {
    double t = 0.0;
    for (uint32_t w = 0; w < (blockDim.x >> 5); ++w) t += sm[w * 4 + threadIdx.x];
    atomicAdd(&sums[threadIdx.x], t);
  }
}

// sums[0] = sum |a|^2, sums[1] = sum |tau x a|^2
__global__ void __launch_bounds__(256)
sample_reduce_fwd(int64_t B, int n, const float* __restrict__ cp, const uint8_t* __restrict__ is_bezier,
                  const float* __restrict__ tt, double* __restrict__ sums) {
  pdl_wait();
  __shared__ double sm[8];
  const int64_t P = B * n;
  double s1 = 0.0, s2 = 0.0;
  for (int64_t g = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; g < P; g += int64_t(gridDim.x) * blockDim.x) {
    const int64_t b = g / n;
    const int m = int(g - b * n);
    const Curve k = load_curve(cp, is_bezier, b);
    const Frame f = make_frame(curve_tangent(k, tt[m]));
    s1 += double(dot(f.a, f.a));
    s2 += double(dot(f.c, f.c));
  }
  const double t1 = block_sum(s1, sm);
  const double t2 = block_sum(s2, sm);
  if (threadIdx.x == 0) { atomicAdd(&sums[0], t1); atomicAdd(&sums[1], t2); }
}

__global__ void __launch_bounds__(256)
sample_fwd_main(int64_t B, int n, const float* __restrict__ cp, const float* __restrict__ width,
                const uint8_t* __restrict__ is_bezier, const float* __restrict__ tt, float half_step,
                const double* __restrict__ sums, float* __restrict__ xyz, float* __restrict__ rot,
                float* __restrict__ scaling, float* __restrict__ norms) {
  pdl_wait();
  const int64_t P = B * n;
  const int64_t g = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  const float N1 = float(sqrt(sums[0]));
  const float N2 = float(sqrt(sums[1])) / N1;   // ||tau x (a/N1)||_F
  if (g == 0) { norms[0] = N1; norms[1] = N2; }
  if (g >= P) return;
  const int64_t b = g / n;
  const int m = int(g - b * n);
  const Curve k = load_curve(cp, is_bezier, b);
  const float t = tt[m];
  const V3 pos = curve_point(k, t);
  const V3 front = curve_point(k, t - half_step);
  const V3 d = v3(__fsub_rn(pos.x, front.x), __fsub_rn(pos.y, front.y), __fsub_rn(pos.z, front.z));
  const float dist = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(d.x, d.x), __fmul_rn(d.y, d.y)), __fmul_rn(d.z, d.z)));
  const Frame f = make_frame(curve_tangent(k, t));
  const V3 v1 = (1.f / N1) * f.a;
  const V3 v2 = (1.f / (N1 * N2)) * f.c;
  const float mm[3][3] = {{f.v0.x, v1.x, v2.x}, {f.v0.y, v1.y, v2.y}, {f.v0.z, v1.z, v2.z}};
  const Quat q = matrix_to_quat(mm);
  xyz[3 * g] = pos.x; xyz[3 * g + 1] = pos.y; xyz[3 * g + 2] = pos.z;
  reinterpret_cast<float4*>(rot)[g] = make_float4(q.q[0], q.q[1], q.q[2], q.q[3]);
  const float w = expf(width[b]);
  scaling[3 * g] = dist; scaling[3 * g + 1] = w; scaling[3 * g + 2] = w;
}

// Everything the backward needs about one Gaussian, recomputed from the curve.
struct Local {
  Frame f;
  V3 v1, v2;
  float mm[3][3];
  Quat q;
  float gm[3][3];
};
__device__ __forceinline__ Local local_state(const Curve& k, float t, float N1, float N2, const float* gq) {
  Local L;
  L.f = make_frame(curve_tangent(k, t));
  L.v1 = (1.f / N1) * L.f.a;
  L.v2 = (1.f / (N1 * N2)) * L.f.c;
  L.mm[0][0] = L.f.v0.x; L.mm[0][1] = L.v1.x; L.mm[0][2] = L.v2.x;
  L.mm[1][0] = L.f.v0.y; L.mm[1][1] = L.v1.y; L.mm[1][2] = L.v2.y;
  L.mm[2][0] = L.f.v0.z; L.mm[2][1] = L.v1.z; L.mm[2][2] = L.v2.z;
  L.q = matrix_to_quat(L.mm);
  matrix_to_quat_bwd(L.mm, L.q, gq, L.gm);
  return L;
}

// Backward = one pass over the Gaussians + one pass over the curves.
//
// The two whole-tensor norms couple every Gaussian through two scalars,
//   K2 = sum g2.v2      K1 = sum g1.v1 + (sum (g2 x tau).v1 - K2 * sum (v2 x tau).v1) / N2,
// and the tangent gradient of a Gaussian is LINEAR in them: gtau = U - K2*V - K1*W with
//   gc_a = g2/N2, gc_b = v2/N2
//   U = v1 x gc_a + rot((g1 + gc_a x tau)/N1) + (v0 terms)     rot(a) = (-a.y, a.x, 0)
//   V = v1 x gc_b + rot((gc_b x tau)/N1)
//   W = rot(v1/N1)
// sample_bwd_point (thread per Gaussian, the heavy quaternion adjoint) writes gpos, gfront, U,
// V, W to a structure-of-arrays scratch and block-reduces the four global sums;
// sample_bwd_curve (warp per curve) then folds them into the control points with the
// Bernstein / tangent weights, in a fixed order: no atomics on the outputs, bit-reproducible
// gradients, and the heavy math runs once instead of twice (reduce pass + main pass).
constexpr int SB_PT = 14;   // gpos 3, gfront 3, U 3, V 3, W 2 (W.z == 0)

__global__ void __launch_bounds__(256)
sample_bwd_point(int64_t B, int n, const float* __restrict__ cp, const uint8_t* __restrict__ is_bezier,
                 const float* __restrict__ tt, float half_step, const float* __restrict__ norms,
                 const float* __restrict__ dL_dxyz, const float* __restrict__ dL_drot,
                 const float* __restrict__ dL_dscaling, float* __restrict__ pt, double* __restrict__ sums) {
  pdl_wait();
  __shared__ double s_red[32];
  const int64_t P = B * n;
  const int64_t g = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  const float N1 = norms[0], N2 = norms[1];
  double s[4] = {0.0, 0.0, 0.0, 0.0};
  if (g < P) {
    const int64_t b = g / n;
    const int m = int(g - b * n);
    const Curve k = load_curve(cp, is_bezier, b);
    const float t = tt[m];
    V3 gpos = v3(0, 0, 0), gfront = v3(0, 0, 0);
    V3 U = v3(0, 0, 0), V = v3(0, 0, 0), Wv = v3(0, 0, 0);
    if (dL_dxyz) gpos = v3(dL_dxyz[3 * g], dL_dxyz[3 * g + 1], dL_dxyz[3 * g + 2]);
    if (dL_dscaling) {
      const V3 pos = curve_point(k, t), front = curve_point(k, t - half_step);
      const V3 d = pos - front;
      const float dist = sqrtf(dot(d, d));
      if (dist > 0.f) {
        const V3 gd = (dL_dscaling[3 * g] / dist) * d;
        gpos = gpos + gd;
        gfront = v3(-gd.x, -gd.y, -gd.z);
      }
    }
    if (dL_drot) {
      const float4 gq4 = reinterpret_cast<const float4*>(dL_drot)[g];
      const float gq[4] = {gq4.x, gq4.y, gq4.z, gq4.w};
      const Local L = local_state(k, t, N1, N2, gq);
      const V3 g0 = v3(L.gm[0][0], L.gm[1][0], L.gm[2][0]);
      const V3 g1 = v3(L.gm[0][1], L.gm[1][1], L.gm[2][1]);
      const V3 g2 = v3(L.gm[0][2], L.gm[1][2], L.gm[2][2]);
      const V3 tau = L.f.tau;
      s[0] = double(dot(g2, L.v2));
      s[1] = double(dot(g1, L.v1));
      s[2] = double(dot(cross(g2, tau), L.v1));
      s[3] = double(dot(cross(L.v2, tau), L.v1));
      const float iN1 = 1.f / N1, iN2 = 1.f / N2;
      const V3 gc_a = iN2 * g2, gc_b = iN2 * L.v2;
      const V3 ga_a = iN1 * (g1 + cross(gc_a, tau));
      const V3 ga_b = iN1 * cross(gc_b, tau);
      const V3 ga_c = iN1 * L.v1;
      U = cross(L.v1, gc_a);
      U.x -= ga_a.y; U.y += ga_a.x;
      V = cross(L.v1, gc_b);
      V.x -= ga_b.y; V.y += ga_b.x;
      Wv = v3(-ga_c.y, ga_c.x, 0.f);
      // v0 = tau / (|tau| + eps)
      const float len = L.f.len, inv = 1.f / (len + 1e-8f);
      U = U + inv * g0;
      if (len > 0.f) U = U - (dot(g0, tau) * inv * inv / len) * tau;
    }
    const float o[SB_PT] = {gpos.x, gpos.y, gpos.z, gfront.x, gfront.y, gfront.z, U.x, U.y, U.z, V.x, V.y, V.z, Wv.x, Wv.y};
#pragma unroll
    for (int i = 0; i < SB_PT; ++i) pt[int64_t(i) * P + g] = o[i];
  }
  if (dL_drot) block_sum4_add(s, s_red, sums);
}

// One warp per curve; lanes stride over the curve's samples (coalesced reads of the scratch).
__global__ void __launch_bounds__(256, 3)   // (85 registers: the kernel is bound by the latency of its 4 load rounds per warp, so by occupancy)
sample_bwd_curve(int64_t B, int n, const float* __restrict__ width, const uint8_t* __restrict__ is_bezier,
                 const float* __restrict__ tt, float half_step, const float* __restrict__ norms,
                 const double* __restrict__ sums, const float* __restrict__ pt, const float* __restrict__ dL_dscaling,
                 float* __restrict__ dL_dcp, float* __restrict__ dL_dwidth, int accumulate) {
  pdl_wait();
  const int64_t b = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (b >= B) return;
  const int64_t P = B * n;
  const bool bez = is_bezier ? (is_bezier[b] != 0) : true;
  const float K2 = float(sums[0]);
  const float K1 = float(sums[1] + (sums[2] - sums[0] * sums[3]) / double(norms[1]));
  float acc[13];
#pragma unroll
  for (int i = 0; i < 13; ++i) acc[i] = 0.f;
  for (int m = lane; m < n; m += 32) {
    const int64_t g = b * n + m;
    const float t = tt[m];
    float v[SB_PT];
#pragma unroll
    for (int i = 0; i < SB_PT; ++i) v[i] = __ldg(pt + int64_t(i) * P + g);
    // gtau = U - K2 V - K1 W
    const float gtx = v[6] - K2 * v[9] - K1 * v[12];
    const float gty = v[7] - K2 * v[10] - K1 * v[13];
    const float gtz = v[8] - K2 * v[11];
    float w[4], wf[4], cw[4];
    point_weights(bez, t, w);
    point_weights(bez, t - half_step, wf);
    if (bez) {
      float tw[3];
      tangent_weights(t, tw);
      cw[0] = -tw[0]; cw[1] = tw[0] - tw[1]; cw[2] = tw[1] - tw[2]; cw[3] = tw[2];
    } else { cw[0] = -1.f; cw[1] = 0.f; cw[2] = 0.f; cw[3] = 1.f; }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      acc[3 * j] += w[j] * v[0] + wf[j] * v[3] + cw[j] * gtx;
      acc[3 * j + 1] += w[j] * v[1] + wf[j] * v[4] + cw[j] * gty;
      acc[3 * j + 2] += w[j] * v[2] + wf[j] * v[5] + cw[j] * gtz;
    }
    if (dL_dscaling) acc[12] += dL_dscaling[3 * g + 1] + dL_dscaling[3 * g + 2];
  }
#pragma unroll
  for (int i = 0; i < 13; ++i)
    for (int o = 16; o > 0; o >>= 1) acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], o);
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < 12; ++i) dL_dcp[b * 12 + i] = accumulate ? dL_dcp[b * 12 + i] + acc[i] : acc[i];
    const float gw = acc[12] * expf(width[b]);
    dL_dwidth[b] = accumulate ? dL_dwidth[b] + gw : gw;
  }
}

}  // namespace cg

using namespace cg;

extern "C" {

size_t cg_sample_scratch_bytes(int64_t B, int32_t n) {
  // 8 doubles of global sums + the per-Gaussian structure-of-arrays scratch of the backward
  return 128 + size_t(B < 1 ? 1 : B) * size_t(n < 1 ? 1 : n) * SB_PT * sizeof(float);
}

int cg_sample_fwd(int64_t B, int32_t n, const float* curve_points, const float* width, const uint8_t* is_bezier,
                  const float* t, float half_step, float* xyz, float* rotation, float* scaling, float* norms,
                  void* scratch, void* stream) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (B == 0) return CG_OK;
  CG_ARG(B > 0 && n > 0, "B/n");
  CG_ARG(curve_points && width && t && xyz && rotation && scaling && norms && scratch, "sample_fwd pointers");
  CG_ARG((reinterpret_cast<uintptr_t>(rotation) & 15u) == 0, "rotation must be 16-byte aligned");
  CG_ARG((reinterpret_cast<uintptr_t>(scratch) & 7u) == 0, "scratch must be 8-byte aligned");
  double* sums = reinterpret_cast<double*>(scratch);
  const int64_t P = B * n;
  CG_CUDA(cudaMemsetAsync(sums, 0, 8 * sizeof(double), st));
  const int rb = int((P + 255) / 256 < 148 * 8 ? (P + 255) / 256 : 148 * 8);
  StageTimer t_(ST_SAMPLE_FWD, st, 2);
  launch_k(sample_reduce_fwd, dim3(rb), dim3(256), 0, st, B, n, curve_points, is_bezier, t, sums);
  CG_LAUNCH_CHECK(0, st);
  launch_k(sample_fwd_main, dim3(unsigned((P + 255) / 256)), dim3(256), 0, st, B, n, curve_points, width, is_bezier, t, half_step, sums,
                                                             xyz, rotation, scaling, norms);
  CG_LAUNCH_CHECK(0, st);
  return CG_OK;
}

int cg_sample_bwd(int64_t B, int32_t n, const float* curve_points, const float* width, const uint8_t* is_bezier,
                  const float* t, float half_step, const float* norms, const float* dL_dxyz,
                  const float* dL_drotation, const float* dL_dscaling, float* dL_dcurve_points, float* dL_dwidth,
                  void* scratch, int32_t accumulate, void* stream) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (B == 0) return CG_OK;
  CG_ARG(B > 0 && n > 0, "B/n");
  CG_ARG(curve_points && width && t && norms && dL_dcurve_points && dL_dwidth && scratch, "sample_bwd pointers");
  CG_ARG(!dL_drotation || (reinterpret_cast<uintptr_t>(dL_drotation) & 15u) == 0, "dL_drotation must be 16-byte aligned");
  double* sums = reinterpret_cast<double*>(scratch);
  const int64_t P = B * n;
  CG_CUDA(cudaMemsetAsync(sums, 0, 8 * sizeof(double), st));
  float* pt = reinterpret_cast<float*>(reinterpret_cast<char*>(scratch) + 128);
  StageTimer t_(ST_SAMPLE_BWD, st, 2);
  launch_k(sample_bwd_point, dim3(unsigned((P + 255) / 256)), dim3(256), 0, st, B, n, curve_points, is_bezier, t, half_step, norms, dL_dxyz,
                                                             dL_drotation, dL_dscaling, pt, sums);
  CG_LAUNCH_CHECK(0, st);
  launch_k(sample_bwd_curve, dim3(unsigned((B * 32 + 255) / 256)), dim3(256), 0, st, B, n, width, is_bezier, t, half_step, norms, sums, pt,
                                                                  dL_dscaling, dL_dcurve_points, dL_dwidth, int(accumulate));
  CG_LAUNCH_CHECK(0, st);
  return CG_OK;
}

}  // extern "C"
