// Small fp32 matrix helpers used by the EWA projection kernels.
//
// Bit-exact parity with the reference needs the same fp32 expression trees the
// reference gets from glm (column-major mat3, product written as
// a0*b0 + a1*b1 + a2*b2 per element) so that nvcc's FMA contraction lands on
// the same operations. M3 stores m[col][row] like glm::mat3 and mul() sums in
// glm's k = 0,1,2 order (third_party/glm/glm/detail/type_mat3x3.inl, operator*).
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

__device__ __forceinline__ M3 m3_mul(const M3& a, const M3& b) {
  M3 r;
#pragma unroll
  for (int c = 0; c < 3; ++c)
#pragma unroll
    for (int rw = 0; rw < 3; ++rw)
      r.m[c][rw] = a.m[0][rw] * b.m[c][0] + a.m[1][rw] * b.m[c][1] + a.m[2][rw] * b.m[c][2];
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
  r.x = m[0] * x + m[4] * y + m[8] * z + m[12];
  r.y = m[1] * x + m[5] * y + m[9] * z + m[13];
  r.z = m[2] * x + m[6] * y + m[10] * z + m[14];
  return r;
}
__device__ __forceinline__ float4 xform44(float x, float y, float z, const float* m) {
  float4 r;
  r.x = m[0] * x + m[4] * y + m[8] * z + m[12];
  r.y = m[1] * x + m[5] * y + m[9] * z + m[13];
  r.z = m[2] * x + m[6] * y + m[10] * z + m[14];
  r.w = m[3] * x + m[7] * y + m[11] * z + m[15];
  return r;
}

// Rotation matrix of a RAW (not normalised) real-first quaternion, glm column
// layout as in forward.cu:118-152 / backward.cu:329-392.
__device__ __forceinline__ M3 quat_to_m3(float r, float x, float y, float z) {
  return m3_cols(1.f - 2.f * (y * y + z * z), 2.f * (x * y - r * z), 2.f * (x * z + r * y),
                 2.f * (x * y + r * z), 1.f - 2.f * (x * x + z * z), 2.f * (y * z - r * x),
                 2.f * (x * z - r * y), 2.f * (y * z + r * x), 1.f - 2.f * (x * x + y * y));
}

// Sigma = (S R)^T (S R), upper triangle [xx,xy,xz,yy,yz,zz] (forward.cu:118-152).
__device__ __forceinline__ void cov3d_from_scale_rot(float sx, float sy, float sz, float mod,
                                                     float4 q, float* cov6) {
  M3 S = m3_cols(mod * sx, 0.f, 0.f, 0.f, mod * sy, 0.f, 0.f, 0.f, mod * sz);
  M3 R = quat_to_m3(q.x, q.y, q.z, q.w);
  M3 M = m3_mul(S, R);
  M3 Sg = m3_mul(m3_t(M), M);
  cov6[0] = Sg.m[0][0]; cov6[1] = Sg.m[0][1]; cov6[2] = Sg.m[0][2];
  cov6[3] = Sg.m[1][1]; cov6[4] = Sg.m[1][2]; cov6[5] = Sg.m[2][2];
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
  const float limx = 1.3f * tanx;
  const float limy = 1.3f * tany;
  p.txtz = t.x / t.z;
  p.tytz = t.y / t.z;
  t.x = fminf(limx, fmaxf(-limx, p.txtz)) * t.z;
  t.y = fminf(limy, fmaxf(-limy, p.tytz)) * t.z;
  p.t = t;
  M3 J = m3_cols(fx / t.z, 0.0f, -(fx * t.x) / (t.z * t.z),
                 0.0f, fy / t.z, -(fy * t.y) / (t.z * t.z),
                 0.f, 0.f, 0.f);
  M3 Wm = m3_cols(vm[0], vm[4], vm[8], vm[1], vm[5], vm[9], vm[2], vm[6], vm[10]);
  p.T = m3_mul(Wm, J);
  p.Vrk = m3_cols(cov6[0], cov6[1], cov6[2], cov6[1], cov6[3], cov6[4], cov6[2], cov6[4], cov6[5]);
  M3 c = m3_mul(m3_mul(m3_t(p.T), m3_t(p.Vrk)), p.T);
  p.cov = make_float3(c.m[0][0], c.m[0][1], c.m[1][1]);
  return p;
}

}  // namespace cg
