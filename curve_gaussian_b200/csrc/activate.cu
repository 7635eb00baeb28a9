// Per-view activations between the sampled Gaussians and the rasterizer, as ONE
// forward and ONE backward kernel (the reference spends ~60 ATen launches and an
// autograd graph on this per view):
//   rotations = F.normalize(_rotation)                       scene/gaussian_curve_model.py:120-122
//   opacity   = sigmoid(_opacity) broadcast over the samples  :107-110
//   mask      = straight-through 1[sigmoid(_mask) > thr]      gaussian_renderer/__init__.py:72-76
//   all_map   = (first column of R(q^) facing the camera, rotated to view space, 1)
//                                                            :99-105, gaussian_renderer/__init__.py:98-104
#include "common.cuh"

namespace cg {

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

__global__ void __launch_bounds__(256)
activate_fwd(int64_t P, int n, const float* __restrict__ xyz, const float* __restrict__ rot,
             const float* __restrict__ scaling, const float* __restrict__ opacity_logit,
             const float* __restrict__ mask_logit, float mask_thr, const float* __restrict__ campos,
             const float* __restrict__ vm, float* __restrict__ rot_n, float* __restrict__ opacity,
             float* __restrict__ scales, float* __restrict__ all_map) {
  pdl_wait();
  const int64_t g = int64_t(blockIdx.x) * 256 + threadIdx.x;
  if (g >= P) return;
  const float4 q = reinterpret_cast<const float4*>(rot)[g];
  const float nrm = fmaxf(sqrtf(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w), 1e-12f);
  const float r = q.x / nrm, i = q.y / nrm, j = q.z / nrm, k = q.w / nrm;
  reinterpret_cast<float4*>(rot_n)[g] = make_float4(r, i, j, k);
  float m = 1.f;
  if (mask_logit) {
    const float sm = sigmoidf_(mask_logit[g]);
    m = __fadd_rn(__fsub_rn(sm > mask_thr ? 1.f : 0.f, sm), sm);   // ((sm>thr) - sm).detach() + sm
  }
  const float op = sigmoidf_(opacity_logit[g / n]);
  opacity[g] = mask_logit ? op * m : op;
  const float s0 = scaling[3 * g], s1 = scaling[3 * g + 1], s2 = scaling[3 * g + 2];
  if (mask_logit) { scales[3 * g] = s0 * m; scales[3 * g + 1] = s1 * m; scales[3 * g + 2] = s2 * m; }
  else { scales[3 * g] = s0; scales[3 * g + 1] = s1; scales[3 * g + 2] = s2; }
  const float two_s = 2.0f / (r * r + i * i + j * j + k * k);
  float c0 = 1.f - two_s * (j * j + k * k), c1 = two_s * (i * j + k * r), c2 = two_s * (i * k - j * r);
  const float tx = campos[0] - xyz[3 * g], ty = campos[1] - xyz[3 * g + 1], tz = campos[2] - xyz[3 * g + 2];
  if (c0 * tx + c1 * ty + c2 * tz < 0.0f) { c0 = -c0; c1 = -c1; c2 = -c2; }
  reinterpret_cast<float4*>(all_map)[g] =
      make_float4(c0 * vm[0] + c1 * vm[4] + c2 * vm[8], c0 * vm[1] + c1 * vm[5] + c2 * vm[9],
                  c0 * vm[2] + c1 * vm[6] + c2 * vm[10], 1.0f);
}

// Backward: one thread per Gaussian for everything per-Gaussian (coalesced, fully parallel), then one
// warp per curve for the only per-curve quantity, the opacity-logit gradient, as a fixed-order shuffle
// reduction (no atomic scatter, reproducible).
__global__ void __launch_bounds__(256)
activate_bwd_point(int64_t P, int n, const float* __restrict__ xyz, const float* __restrict__ rot,
                   const float* __restrict__ scaling, const float* __restrict__ opacity_logit,
                   const float* __restrict__ mask_logit, float mask_thr, const float* __restrict__ campos,
                   const float* __restrict__ vm, const float* __restrict__ g_rot_n, const float* __restrict__ g_opacity,
                   const float* __restrict__ g_scales, const float* __restrict__ g_all_map,
                   float* __restrict__ g_rot, float* __restrict__ g_scaling, float* __restrict__ g_mask_logit, int accumulate) {
  pdl_wait();
  const int64_t g = int64_t(blockIdx.x) * 256 + threadIdx.x;
  if (g >= P) return;
  // ---- mask / opacity / scales
  float m = 1.f, sm = 0.f;
  if (mask_logit) {
    sm = sigmoidf_(mask_logit[g]);
    m = __fadd_rn(__fsub_rn(sm > mask_thr ? 1.f : 0.f, sm), sm);
  }
  float gs0 = 0.f, gs1 = 0.f, gs2 = 0.f;
  if (g_scales) { gs0 = g_scales[3 * g]; gs1 = g_scales[3 * g + 1]; gs2 = g_scales[3 * g + 2]; }
  g_scaling[3 * g] = gs0 * m; g_scaling[3 * g + 1] = gs1 * m; g_scaling[3 * g + 2] = gs2 * m;
  if (mask_logit && g_mask_logit) {
    const float go = g_opacity ? g_opacity[g] : 0.f;
    const float op = sigmoidf_(opacity_logit[g / n]);
    const float gm = go * op + gs0 * scaling[3 * g] + gs1 * scaling[3 * g + 1] + gs2 * scaling[3 * g + 2];
    const float gml = gm * sm * (1.f - sm);
    g_mask_logit[g] = accumulate ? g_mask_logit[g] + gml : gml;
  }
  // ---- rotation
  const float4 q = reinterpret_cast<const float4*>(rot)[g];
  const float nraw = sqrtf(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w);
  const float nrm = fmaxf(nraw, 1e-12f);
  const float r = q.x / nrm, i = q.y / nrm, j = q.z / nrm, k = q.w / nrm;
  float4 gq = g_rot_n ? reinterpret_cast<const float4*>(g_rot_n)[g] : make_float4(0.f, 0.f, 0.f, 0.f);
  if (g_all_map) {
    const float4 ga = reinterpret_cast<const float4*>(g_all_map)[g];
    // local = dir @ V[:3,:3]  ->  g_dir_r = sum_c ga_c * V[r][c]
    float d0 = ga.x * vm[0] + ga.y * vm[1] + ga.z * vm[2];
    float d1 = ga.x * vm[4] + ga.y * vm[5] + ga.z * vm[6];
    float d2 = ga.x * vm[8] + ga.y * vm[9] + ga.z * vm[10];
    const float S = r * r + i * i + j * j + k * k;
    const float two_s = 2.0f / S;
    const float c0 = 1.f - two_s * (j * j + k * k), c1 = two_s * (i * j + k * r), c2 = two_s * (i * k - j * r);
    const float tx = campos[0] - xyz[3 * g], ty = campos[1] - xyz[3 * g + 1], tz = campos[2] - xyz[3 * g + 2];
    if (c0 * tx + c1 * ty + c2 * tz < 0.0f) { d0 = -d0; d1 = -d1; d2 = -d2; }
    // c0 = 1 - ts*(jj+kk), c1 = ts*(ij+kr), c2 = ts*(ik-jr), ts = 2/S
    const float e0 = j * j + k * k, e1 = i * j + k * r, e2 = i * k - j * r;
    const float g_ts = -d0 * e0 + d1 * e1 + d2 * e2;
    const float g_S = -g_ts * two_s / S;
    gq.x += two_s * (d1 * k - d2 * j) + 2.f * r * g_S;
    gq.y += two_s * (d1 * j + d2 * k) + 2.f * i * g_S;
    gq.z += two_s * (-2.f * j * d0 + d1 * i - d2 * r) + 2.f * j * g_S;
    gq.w += two_s * (-2.f * k * d0 + d1 * r + d2 * i) + 2.f * k * g_S;
  }
  // through x / max(|x|, eps)
  float4 out;
  if (nraw > 1e-12f) {
    const float dt = gq.x * r + gq.y * i + gq.z * j + gq.w * k;
    out = make_float4((gq.x - r * dt) / nrm, (gq.y - i * dt) / nrm, (gq.z - j * dt) / nrm, (gq.w - k * dt) / nrm);
  } else {
    out = make_float4(gq.x / nrm, gq.y / nrm, gq.z / nrm, gq.w / nrm);
  }
  reinterpret_cast<float4*>(g_rot)[g] = out;
}

__global__ void __launch_bounds__(256)
activate_bwd_curve(int64_t B, int n, const float* __restrict__ opacity_logit, const float* __restrict__ mask_logit,
                   float mask_thr, const float* __restrict__ g_opacity, float* __restrict__ g_opacity_logit, int accumulate) {
  pdl_wait();
  const int64_t b = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (b >= B) return;
  float g_op_sum = 0.f;
  if (g_opacity) {
    for (int mi = lane; mi < n; mi += 32) {
      const int64_t g = b * n + mi;
      float m = 1.f;
      if (mask_logit) {
        const float sm = sigmoidf_(mask_logit[g]);
        m = __fadd_rn(__fsub_rn(sm > mask_thr ? 1.f : 0.f, sm), sm);
      }
      g_op_sum += g_opacity[g] * m;
    }
  }
  for (int o = 16; o > 0; o >>= 1) g_op_sum += __shfl_xor_sync(0xffffffffu, g_op_sum, o);
  const float op = sigmoidf_(opacity_logit[b]);
  if (lane == 0) {
    const float gol = g_op_sum * op * (1.f - op);
    g_opacity_logit[b] = accumulate ? g_opacity_logit[b] + gol : gol;
  }
}

}  // namespace cg

using namespace cg;

extern "C" {

int cg_activate_fwd(int64_t B, int32_t n, const float* xyz, const float* rotation, const float* scaling,
                    const float* opacity_logit, const float* mask_logit, float mask_thr, const float* campos,
                    const float* viewmatrix, float* rot_n, float* opacity, float* scales, float* all_map,
                    void* stream) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (B == 0) return CG_OK;
  CG_ARG(B > 0 && n > 0, "B/n");
  CG_ARG(xyz && rotation && scaling && opacity_logit && campos && viewmatrix && rot_n && opacity && scales && all_map,
         "activate_fwd pointers");
  CG_ARG(((reinterpret_cast<uintptr_t>(rotation) | reinterpret_cast<uintptr_t>(rot_n) |
           reinterpret_cast<uintptr_t>(all_map)) & 15u) == 0, "rotation/rot_n/all_map must be 16-byte aligned");
  const int64_t P = B * n;
  StageTimer t_(ST_ACTIVATE_FWD, st, 1);
  launch_k(activate_fwd, dim3(unsigned((P + 255) / 256)), dim3(256), 0, st, P, n, xyz, rotation, scaling, opacity_logit, mask_logit,
                                                         mask_thr, campos, viewmatrix, rot_n, opacity, scales, all_map);
  CG_LAUNCH_CHECK(0, st);
  return CG_OK;
}

int cg_activate_bwd(int64_t B, int32_t n, const float* xyz, const float* rotation, const float* scaling,
                    const float* opacity_logit, const float* mask_logit, float mask_thr, const float* campos,
                    const float* viewmatrix, const float* g_rot_n, const float* g_opacity, const float* g_scales,
                    const float* g_all_map, float* g_rotation, float* g_scaling, float* g_opacity_logit,
                    float* g_mask_logit, int32_t accumulate, void* stream) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (B == 0) return CG_OK;
  CG_ARG(B > 0 && n > 0, "B/n");
  CG_ARG(xyz && rotation && scaling && opacity_logit && campos && viewmatrix && g_rotation && g_scaling &&
             g_opacity_logit, "activate_bwd pointers");
  CG_ARG(!mask_logit || g_mask_logit, "g_mask_logit required with mask_logit");
  const int64_t P = B * n;
  StageTimer t_(ST_ACTIVATE_BWD, st, 2);
  launch_k(activate_bwd_point, dim3(unsigned((P + 255) / 256)), dim3(256), 0, st, P, n, xyz, rotation, scaling, opacity_logit, mask_logit,
                                                               mask_thr, campos, viewmatrix, g_rot_n, g_opacity, g_scales,
                                                               g_all_map, g_rotation, g_scaling, g_mask_logit, int(accumulate));
  CG_LAUNCH_CHECK(0, st);
  launch_k(activate_bwd_curve, dim3(unsigned((B * 32 + 255) / 256)), dim3(256), 0, st, B, n, opacity_logit, mask_logit, mask_thr,
                                                                    g_opacity, g_opacity_logit, int(accumulate));
  CG_LAUNCH_CHECK(0, st);
  return CG_OK;
}

}  // extern "C"
