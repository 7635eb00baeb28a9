// Small fp32 matrix helpers used by the EWA projection kernels.
//
// Bit-exact tile/sort keys need every fp32 operation of the projection to round
// exactly like the reference build does. The reference leaves FMA contraction
// to the compiler (glm column-major mat3 products a0*b0 + a1*b1 + a2*b2, plain
// C expressions elsewhere); which pairs get fused there was read off its
// sm_100 SASS (oracle/_ref, nvcc 12.9) and is PINNED here with explicit
// __fmaf_rn/__fmul_rn/__fadd_rn intrinsics, which no compiler pass may re-fuse
// or split. Patterns:
//   3-term products   fma(a2,b2, fma(a0,b0, rn(a1*b1)))
//   affine transform  fma(z,m8, fma(x,m0, rn(y*m4))) + m12      (plain add last)
//   x*y -/+ r*z       fma(x,y, -/+ rn(r*z));   x*z +/- r*y = fma(+/-r,y, rn(x*z))
//   a*c - b*b         fma(a,c, -rn(b*b))
// M3 stores m[col][row] like glm::mat3 (type_mat3x3.inl).
#pragma once
#include <cuda_runtime.h>

namespace cg {

struct M3 {
  float m[3][3];  // m[col][row]
};

__device__ __forceinline__ M3 m3_cols(float c00, float c01, float c02,
                                      float c10, float c11, float c12,
                                      float c20, float c21, float c22) {
  M3 r;
  r.m[0][0] = c00; r.m[0][1] = c01; r.m[0][2] = c02;
  r.m[1][0] = c10; r.m[1][1] = c11; r.m[1][2] = c12;
  r.m[2][0] = c20; r.m[2][1] = c21; r.m[2][2] = c22;
  return r;
}

__device__ __forceinline__ float dot3p(float a0, float b0, float a1, float b1, float a2, float b2) {
  return __fmaf_rn(a2, b2, __fmaf_rn(a0, b0, __fmul_rn(a1, b1)));
}

__device__ __forceinline__ M3 m3_mul(const M3& a, const M3& b) {
  M3 r;
#pragma unroll
  for (int c = 0; c < 3; ++c)
#pragma unroll
    for (int rw = 0; rw < 3; ++rw)
      r.m[c][rw] = dot3p(a.m[0][rw], b.m[c][0], a.m[1][rw], b.m[c][1], a.m[2][rw], b.m[c][2]);
  return r;
}

__device__ __forceinline__ M3 m3_t(const M3& a) {
  M3 r;
#pragma unroll
  for (int c = 0; c < 3; ++c)
#pragma unroll
    for (int rw = 0; rw < 3; ++rw) r.m[c][rw] = a.m[rw][c];
  return r;
}

// [x y z 1] @ M, M read as m[4*i + j] (auxiliary.h:70-89).
__device__ __forceinline__ float3 xform43(float x, float y, float z, const float* m) {
  float3 r;
  r.x = __fadd_rn(dot3p(m[0], x, m[4], y, m[8], z), m[12]);
  r.y = __fadd_rn(dot3p(m[1], x, m[5], y, m[9], z), m[13]);
  r.z = __fadd_rn(dot3p(m[2], x, m[6], y, m[10], z), m[14]);
  return r;
}
__device__ __forceinline__ float4 xform44(float x, float y, float z, const float* m) {
  float4 r;
  r.x = __fadd_rn(dot3p(m[0], x, m[4], y, m[8], z), m[12]);
  r.y = __fadd_rn(dot3p(m[1], x, m[5], y, m[9], z), m[13]);
  r.z = __fadd_rn(dot3p(m[2], x, m[6], y, m[10], z), m[14]);
  r.w = __fadd_rn(dot3p(m[3], x, m[7], y, m[11], z), m[15]);
  return r;
}

// Rotation matrix of a RAW (not normalised) real-first quaternion, glm column
// layout as in forward.cu:118-152 / backward.cu:329-392.
__device__ __forceinline__ M3 quat_to_m3(float r, float x, float y, float z) {
  const float rz = __fmul_rn(r, z), rx = __fmul_rn(r, x), xz = __fmul_rn(x, z);
  const float yy = __fmul_rn(y, y), zz = __fmul_rn(z, z);
  const float xy_m = __fmaf_rn(x, y, -rz), xy_p = __fmaf_rn(x, y, rz);
  const float xz_p = __fmaf_rn(r, y, xz), xz_m = __fmaf_rn(-r, y, xz);
  const float yz_m = __fmaf_rn(y, z, -rx), yz_p = __fmaf_rn(y, z, rx);
  const float yy_zz = __fadd_rn(yy, zz), xx_zz = __fmaf_rn(x, x, zz), xx_yy = __fmaf_rn(x, x, yy);
  return m3_cols(__fsub_rn(1.f, __fadd_rn(yy_zz, yy_zz)), __fadd_rn(xy_m, xy_m), __fadd_rn(xz_p, xz_p),
                 __fadd_rn(xy_p, xy_p), __fsub_rn(1.f, __fadd_rn(xx_zz, xx_zz)), __fadd_rn(yz_m, yz_m),
                 __fadd_rn(xz_m, xz_m), __fadd_rn(yz_p, yz_p), __fsub_rn(1.f, __fadd_rn(xx_yy, xx_yy)));
}

// Sigma = (S R)^T (S R), upper triangle [xx,xy,xz,yy,yz,zz] (forward.cu:118-152).
__device__ __forceinline__ void cov3d_from_scale_rot(float sx, float sy, float sz, float mod,
                                                     float4 q, float* cov6) {
  const float s[3] = {__fmul_rn(mod, sx), __fmul_rn(mod, sy), __fmul_rn(mod, sz)};
  M3 R = quat_to_m3(q.x, q.y, q.z, q.w);
  M3 M;  // S*R with S diagonal: every entry is the single rounded product s_row * R
#pragma unroll
  for (int c = 0; c < 3; ++c)
#pragma unroll
    for (int rw = 0; rw < 3; ++rw) M.m[c][rw] = __fmul_rn(s[rw], R.m[c][rw]);
  M3 Sg = m3_mul(m3_t(M), M);
  cov6[0] = Sg.m[0][0]; cov6[1] = Sg.m[0][1]; cov6[2] = Sg.m[0][2];
  cov6[3] = Sg.m[1][1]; cov6[4] = Sg.m[1][2]; cov6[5] = Sg.m[2][2];
}

// Exponent of the 2D Gaussian, -0.5*(A dx^2 + C dy^2) - B dx dy, with the
// reference build's rounding: fma(fma(dx, rn(A dx), rn(rn(C dy) dy)), -0.5, -rn(rn(B dx) dy))
// (forward.cu:356, backward.cu:563).
__device__ __forceinline__ float gauss_power(float A, float B, float C, float dx, float dy) {
  const float inner = __fmaf_rn(dx, __fmul_rn(A, dx), __fmul_rn(__fmul_rn(C, dy), dy));
  return __fmaf_rn(inner, -0.5f, -__fmul_rn(__fmul_rn(B, dx), dy));
}

struct Proj2D {
  float3 t;       // clamped view-space point
  float txtz, tytz;
  M3 T;           // W * J
  M3 Vrk;
  float3 cov;     // (c00, c01, c11) before the 0.3 dilation
};

// EWA projection of a 3D covariance (forward.cu:78-113).
__device__ __forceinline__ Proj2D project_cov(float x, float y, float z, float fx, float fy,
                                              float tanx, float tany, const float* cov6,
                                              const float* vm) {
  Proj2D p;
  float3 t = xform43(x, y, z, vm);
  const float limx = __fmul_rn(1.3f, tanx);
  const float limy = __fmul_rn(1.3f, tany);
  p.txtz = __fdiv_rn(t.x, t.z);
  p.tytz = __fdiv_rn(t.y, t.z);
  t.x = __fmul_rn(fminf(limx, fmaxf(-limx, p.txtz)), t.z);
  t.y = __fmul_rn(fminf(limy, fmaxf(-limy, p.tytz)), t.z);
  p.t = t;
  const float tz2 = __fmul_rn(t.z, t.z);
  M3 J = m3_cols(__fdiv_rn(fx, t.z), 0.0f, __fdiv_rn(-__fmul_rn(fx, t.x), tz2),
                 0.0f, __fdiv_rn(fy, t.z), __fdiv_rn(-__fmul_rn(fy, t.y), tz2),
                 0.f, 0.f, 0.f);
  M3 Wm = m3_cols(vm[0], vm[4], vm[8], vm[1], vm[5], vm[9], vm[2], vm[6], vm[10]);
  p.T = m3_mul(Wm, J);
  p.Vrk = m3_cols(cov6[0], cov6[1], cov6[2], cov6[1], cov6[3], cov6[4], cov6[2], cov6[4], cov6[5]);
  M3 c = m3_mul(m3_mul(m3_t(p.T), m3_t(p.Vrk)), p.T);
  p.cov = make_float3(c.m[0][0], c.m[0][1], c.m[1][1]);
  return p;
}

}  // namespace cg
