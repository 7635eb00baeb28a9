// Curve-side regularisers of the training step, the per-iteration terms train.py adds
// to the image loss:
//
//   curve smoothness       train.py:119-124   1 - |cos(dir[b,m], dir[b,m+1])| averaged over adjacent samples,
//                                             dir = first column of quaternion_to_matrix(normalize(_rotation))
//   endpoint connectivity  train.py:133-146   mean distance over all pairs of curve endpoints closer than 0.05
//                                             (excluding a curve's own two endpoints)
//
// The reference spells the first as ~25 ATen ops on (P,3,3) tensors and the second as torch.cdist over the
// 2B endpoints, i.e. a dense (2B)^2 matrix (1.6 GB at B = 10k) plus three masks of the same size. Here each is
// one streaming forward kernel and one backward kernel; the all-pairs term never materialises the matrix.
#include "common.cuh"

namespace cg {

namespace {

__device__ __forceinline__ double block_sum_r(double v, double* s_red) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0.0;
  if (threadIdx.x == 0)
    for (int i = 0; i < int(blockDim.x >> 5); ++i) t += s_red[i];
  return t;  // valid on thread 0
}

// dir = first column of quaternion_to_matrix(q / max(|q|, 1e-12))  (gaussian_curve_model.py:95-97,120-122)
struct Axis { float x, y, z; };
__device__ __forceinline__ Axis axis_of(float4 q) {
  const float nrm = fmaxf(sqrtf(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w), 1e-12f);
  const float r = q.x / nrm, i = q.y / nrm, j = q.z / nrm, k = q.w / nrm;
  const float two_s = 2.0f / (r * r + i * i + j * j + k * k);
  Axis a;
  a.x = 1.f - two_s * (j * j + k * k);
  a.y = two_s * (i * j + k * r);
  a.z = two_s * (i * k - j * r);
  return a;
}
// adjoint of axis_of: d = dL/d(axis) -> dL/dq (raw quaternion)
__device__ __forceinline__ float4 axis_of_bwd(float4 q, float d0, float d1, float d2) {
  const float nraw = sqrtf(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w);
  const float nrm = fmaxf(nraw, 1e-12f);
  const float r = q.x / nrm, i = q.y / nrm, j = q.z / nrm, k = q.w / nrm;
  const float S = r * r + i * i + j * j + k * k;
  const float two_s = 2.0f / S;
  const float e0 = j * j + k * k, e1 = i * j + k * r, e2 = i * k - j * r;
  const float g_ts = -d0 * e0 + d1 * e1 + d2 * e2;
  const float g_S = -g_ts * two_s / S;
  float4 gq;
  gq.x = two_s * (d1 * k - d2 * j) + 2.f * r * g_S;
  gq.y = two_s * (d1 * j + d2 * k) + 2.f * i * g_S;
  gq.z = two_s * (-2.f * j * d0 + d1 * i - d2 * r) + 2.f * j * g_S;
  gq.w = two_s * (-2.f * k * d0 + d1 * r + d2 * i) + 2.f * k * g_S;
  if (nraw > 1e-12f) {
    const float dt = gq.x * r + gq.y * i + gq.z * j + gq.w * k;
    return make_float4((gq.x - r * dt) / nrm, (gq.y - i * dt) / nrm, (gq.z - j * dt) / nrm, (gq.w - k * dt) / nrm);
  }
  return make_float4(gq.x / nrm, gq.y / nrm, gq.z / nrm, gq.w / nrm);
}

constexpr float COS_EPS = 1e-8f;   // F.cosine_similarity default eps
// cos = x.y / (max(|x|, eps) * max(|y|, eps))
__device__ __forceinline__ float cos_sim(Axis x, Axis y) {
  const float nx = fmaxf(sqrtf(x.x * x.x + x.y * x.y + x.z * x.z), COS_EPS);
  const float ny = fmaxf(sqrtf(y.x * y.x + y.y * y.y + y.z * y.z), COS_EPS);
  return (x.x * y.x + x.y * y.y + x.z * y.z) / (nx * ny);
}
// d(1 - |cos(x, y)|)/dx, for norms above eps
__device__ __forceinline__ Axis smooth_pair_grad(Axis x, Axis y) {
  const float nx2 = x.x * x.x + x.y * x.y + x.z * x.z, ny2 = y.x * y.x + y.y * y.y + y.z * y.z;
  const float nx = fmaxf(sqrtf(nx2), COS_EPS), ny = fmaxf(sqrtf(ny2), COS_EPS);
  const float inv = 1.f / (nx * ny);
  const float c = (x.x * y.x + x.y * y.y + x.z * y.z) * inv;
  const float sgn = c > 0.f ? -1.f : (c < 0.f ? 1.f : 0.f);   // d(-|c|)/dc
  const float k = sqrtf(nx2) > COS_EPS ? c / nx2 : 0.f;       // the clamped norm is a constant
  Axis g;
  g.x = sgn * (y.x * inv - k * x.x);
  g.y = sgn * (y.y * inv - k * x.y);
  g.z = sgn * (y.z * inv - k * x.z);
  return g;
}
}  // namespace

// sums[0] += sum over adjacent pairs of (1 - |cos|)
__global__ void __launch_bounds__(256)
curve_smooth_fwd_kernel(int64_t P, int n, const float* __restrict__ rot, double* __restrict__ sums) {
  pdl_wait();
  __shared__ double s_red[8];
  const int64_t g = int64_t(blockIdx.x) * 256 + threadIdx.x;
  double v = 0.0;
  if (g < P && int(g % n) < n - 1) {
    const Axis a = axis_of(reinterpret_cast<const float4*>(rot)[g]);
    const Axis b = axis_of(reinterpret_cast<const float4*>(rot)[g + 1]);
    v = double(1.f - fabsf(cos_sim(a, b)));
  }
  const double t = block_sum_r(v, s_red);
  if (threadIdx.x == 0) atomicAdd(&sums[0], t);
}

__global__ void curve_smooth_finish_kernel(int64_t npairs, const double* __restrict__ sums, float* __restrict__ loss) {
  pdl_wait();
  *loss = npairs > 0 ? float(sums[0] / double(npairs)) : 0.f;
}

__global__ void __launch_bounds__(256)
curve_smooth_bwd_kernel(int64_t P, int n, const float* __restrict__ rot, const float* __restrict__ g_loss,
                        float scale, float* __restrict__ g_rot) {
  pdl_wait();
  const int64_t g = int64_t(blockIdx.x) * 256 + threadIdx.x;
  if (g >= P) return;
  const int m = int(g % n);
  const float4 q = reinterpret_cast<const float4*>(rot)[g];
  const Axis a = axis_of(q);
  float d0 = 0.f, d1 = 0.f, d2 = 0.f;
  if (m < n - 1) {
    const Axis gr = smooth_pair_grad(a, axis_of(reinterpret_cast<const float4*>(rot)[g + 1]));
    d0 += gr.x; d1 += gr.y; d2 += gr.z;
  }
  if (m > 0) {
    const Axis gl = smooth_pair_grad(a, axis_of(reinterpret_cast<const float4*>(rot)[g - 1]));
    d0 += gl.x; d1 += gl.y; d2 += gl.z;
  }
  const float k = (g_loss ? __ldg(g_loss) : 1.f) * scale;
  reinterpret_cast<float4*>(g_rot)[g] = axis_of_bwd(q, k * d0, k * d1, k * d2);
}

// ---------------------------------------------------------------------------
// Endpoint connectivity. Point i < B is curve i's first control point, point B + i its last one.
constexpr int EP_THREADS = 256;

__device__ __forceinline__ float3 endpoint(const float* __restrict__ cp, int64_t B, int64_t i) {
  const float* p = (i < B) ? cp + i * 12 : cp + (i - B) * 12 + 9;
  return make_float3(p[0], p[1], p[2]);
}

// grid (ceil(N/256), chunks): thread = point i, block row y scans the j-chunk y through shared memory.
// part[y][i] = {sum dist, count, sum (p_i - p_j)/dist (3)} over the valid j of that chunk (fixed order).
__global__ void __launch_bounds__(EP_THREADS)
endpoint_conn_pairs_kernel(int64_t B, const float* __restrict__ cp, float thr, int64_t chunk,
                           float* __restrict__ part) {
  pdl_wait();
  __shared__ float4 s_p[EP_THREADS];
  const int64_t N = 2 * B;
  const int64_t i = int64_t(blockIdx.x) * EP_THREADS + threadIdx.x;
  float3 pi = make_float3(0.f, 0.f, 0.f);
  if (i < N) pi = endpoint(cp, B, i);
  const int64_t ci = i < B ? i : i - B;    // curve of point i
  const int64_t j0 = int64_t(blockIdx.y) * chunk, j1 = j0 + chunk < N ? j0 + chunk : N;
  const float thr2 = thr * thr * 1.0001f;  // cheap pre-filter on the squared distance; the exact test follows
  float s = 0.f, cnt = 0.f, vx = 0.f, vy = 0.f, vz = 0.f;
  for (int64_t jb = j0; jb < j1; jb += EP_THREADS) {
    const int64_t j = jb + threadIdx.x;
    __syncthreads();
    if (j < j1) {
      const float3 pj = endpoint(cp, B, j);
      s_p[threadIdx.x] = make_float4(pj.x, pj.y, pj.z, __int_as_float(int(j < B ? j : j - B)));
    }
    __syncthreads();
    const int lim = int(j1 - jb < EP_THREADS ? j1 - jb : EP_THREADS);
    if (i < N) {
      for (int t = 0; t < lim; ++t) {
        const float4 q = s_p[t];
        const float dx = pi.x - q.x, dy = pi.y - q.y, dz = pi.z - q.z;
        const float d2 = dx * dx + dy * dy + dz * dz;
        if (d2 < thr2 && __float_as_int(q.w) != int(ci)) {
          const float d = sqrtf(d2);
          if (d < thr) {
            s += d;
            cnt += 1.f;
            if (d > 0.f) { const float inv = 1.f / d; vx += dx * inv; vy += dy * inv; vz += dz * inv; }
          }
        }
      }
    }
  }
  if (i < N) {
    float* o = part + (int64_t(blockIdx.y) * N + i) * 5;
    o[0] = s; o[1] = cnt; o[2] = vx; o[3] = vy; o[4] = vz;
  }
}

// Folds the chunks: per-point gradient direction v (N,3) and the two totals (sum of distances, pair count).
__global__ void __launch_bounds__(256)
endpoint_conn_fold_kernel(int64_t N, int chunks, const float* __restrict__ part, float* __restrict__ v,
                          double* __restrict__ sums) {
  pdl_wait();
  __shared__ double s_red[8];
  const int64_t i = int64_t(blockIdx.x) * 256 + threadIdx.x;
  double s = 0.0, c = 0.0;
  if (i < N) {
    float vx = 0.f, vy = 0.f, vz = 0.f;
    for (int y = 0; y < chunks; ++y) {
      const float* o = part + (int64_t(y) * N + i) * 5;
      s += double(o[0]); c += double(o[1]);
      vx += o[2]; vy += o[3]; vz += o[4];
    }
    v[3 * i] = vx; v[3 * i + 1] = vy; v[3 * i + 2] = vz;
  }
  const double ts = block_sum_r(s, s_red);
  const double tc = block_sum_r(c, s_red);
  if (threadIdx.x == 0) { atomicAdd(&sums[0], ts); atomicAdd(&sums[1], tc); }
}

__global__ void endpoint_conn_finish_kernel(const double* __restrict__ sums, float* __restrict__ loss) {
  pdl_wait();
  *loss = sums[1] > 0.0 ? float(sums[0] / sums[1]) : 0.f;
}

// dL/dcurve_points: every ordered pair (i,j) and (j,i) is in the mean, so point i gets 2 v_i / count.
__global__ void __launch_bounds__(256)
endpoint_conn_bwd_kernel(int64_t B, const float* __restrict__ v, const double* __restrict__ sums,
                         const float* __restrict__ g_loss, float* __restrict__ g_cp) {
  pdl_wait();
  const int64_t b = int64_t(blockIdx.x) * 256 + threadIdx.x;
  if (b >= B) return;
  const float k = sums[1] > 0.0 ? (g_loss ? __ldg(g_loss) : 1.f) * float(2.0 / sums[1]) : 0.f;
  float* o = g_cp + b * 12;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    o[c] = k * v[3 * b + c];
    o[3 + c] = 0.f;
    o[6 + c] = 0.f;
    o[9 + c] = k * v[3 * (B + b) + c];
  }
}

}  // namespace cg

using namespace cg;

extern "C" {

size_t cg_curve_smooth_scratch_bytes(void) { return 64; }

int cg_curve_smooth_fwd(int64_t B, int32_t n, const float* rotation, void* scratch, float* loss_out, void* stream) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  CG_ARG(B >= 0 && n > 0 && scratch && loss_out, "curve_smooth_fwd arguments");
  CG_ARG(B == 0 || rotation, "rotation");
  CG_ARG((reinterpret_cast<uintptr_t>(rotation) & 15u) == 0 && (reinterpret_cast<uintptr_t>(scratch) & 7u) == 0,
         "rotation must be 16-byte, scratch 8-byte aligned");
  double* sums = reinterpret_cast<double*>(scratch);
  const int64_t P = B * n;
  CG_CUDA(cudaMemsetAsync(sums, 0, 64, st));
  count_launches(P > 0 ? 2 : 1);
  if (P > 0) {
    launch_k(curve_smooth_fwd_kernel, dim3(unsigned((P + 255) / 256)), dim3(256), 0, st, P, n, rotation, sums);
    CG_LAUNCH_CHECK(0, st);
  }
  launch_k(curve_smooth_finish_kernel, dim3(1), dim3(1), 0, st, B * (n - 1), sums, loss_out);
  CG_LAUNCH_CHECK(0, st);
  return CG_OK;
}

int cg_curve_smooth_bwd(int64_t B, int32_t n, const float* rotation, const float* g_loss, float* g_rotation,
                        void* stream) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (B == 0) return CG_OK;
  CG_ARG(B > 0 && n > 0 && rotation && g_rotation, "curve_smooth_bwd arguments");
  CG_ARG(((reinterpret_cast<uintptr_t>(rotation) | reinterpret_cast<uintptr_t>(g_rotation)) & 15u) == 0,
         "rotation/g_rotation must be 16-byte aligned");
  const int64_t P = B * n, npairs = B * (n - 1);
  count_launches(1);
  launch_k(curve_smooth_bwd_kernel, dim3(unsigned((P + 255) / 256)), dim3(256), 0, st, P, n, rotation, g_loss,
                                                                    npairs > 0 ? 1.0f / float(npairs) : 0.f, g_rotation);
  CG_LAUNCH_CHECK(0, st);
  return CG_OK;
}

static int ep_chunks(int64_t N) {
  // enough (point-block, j-chunk) CTAs to fill the GPU a few times over, chunks a multiple of the CTA width
  const int64_t xb = (N + EP_THREADS - 1) / EP_THREADS;
  int64_t want = (148 * 4 + xb - 1) / (xb > 0 ? xb : 1);
  if (want < 1) want = 1;
  if (want > xb) want = xb > 0 ? xb : 1;
  return int(want);
}

size_t cg_endpoint_conn_scratch_bytes(int64_t B) {
  const int64_t N = 2 * (B < 1 ? 1 : B);
  return 64 + size_t(ep_chunks(N)) * size_t(N) * 5 * sizeof(float);
}

int cg_endpoint_conn_fwd(int64_t B, const float* curve_points, float dis_thr, void* scratch, float* v,
                         float* loss_out, void* stream) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  CG_ARG(B >= 0 && scratch && loss_out, "endpoint_conn_fwd arguments");
  CG_ARG((reinterpret_cast<uintptr_t>(scratch) & 7u) == 0, "scratch must be 8-byte aligned");
  double* sums = reinterpret_cast<double*>(scratch);
  CG_CUDA(cudaMemsetAsync(sums, 0, 64, st));
  if (B > 0) {
    CG_ARG(curve_points && v, "curve_points / v");
    const int64_t N = 2 * B;
    const int chunks = ep_chunks(N);
    int64_t chunk = (N + chunks - 1) / chunks;
    chunk = (chunk + EP_THREADS - 1) / EP_THREADS * EP_THREADS;
    float* part = reinterpret_cast<float*>(reinterpret_cast<char*>(scratch) + 64);
    dim3 grid(unsigned((N + EP_THREADS - 1) / EP_THREADS), unsigned((N + chunk - 1) / chunk));
    count_launches(2);
    launch_k(endpoint_conn_pairs_kernel, dim3(grid), dim3(EP_THREADS), 0, st, B, curve_points, dis_thr, chunk, part);
    CG_LAUNCH_CHECK(0, st);
    launch_k(endpoint_conn_fold_kernel, dim3(unsigned((N + 255) / 256)), dim3(256), 0, st, N, int(grid.y), part, v, sums);
    CG_LAUNCH_CHECK(0, st);
  }
  count_launches(1);
  launch_k(endpoint_conn_finish_kernel, dim3(1), dim3(1), 0, st, sums, loss_out);
  CG_LAUNCH_CHECK(0, st);
  return CG_OK;
}

int cg_endpoint_conn_bwd(int64_t B, const float* v, const void* scratch, const float* g_loss, float* g_curve_points,
                         void* stream) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (B == 0) return CG_OK;
  CG_ARG(B > 0 && v && scratch && g_curve_points, "endpoint_conn_bwd arguments");
  count_launches(1);
  launch_k(endpoint_conn_bwd_kernel, dim3(unsigned((B + 255) / 256)), dim3(256), 0, st, B, v, reinterpret_cast<const double*>(scratch),
                                                                     g_loss, g_curve_points);
  CG_LAUNCH_CHECK(0, st);
  return CG_OK;
}

}  // extern "C"
