// Mean squared distance to the 3 nearest neighbours (replaces submodules/simple-knn,
// simple_knn.cu:46-222: bbox -> 30-bit Morton codes -> sort -> boxes of 1024 ->
// box-culled exact search). Init-time only (gaussian_curve_model.py:149).
//
// The cull is conservative, so the result is the exact 3-NN whatever the
// traversal; the Morton order only bounds the work. The sort reuses the
// onesweep radix sort of the rasterizer (Morton code in the low key bits).
#include <float.h>
#include "common.cuh"

namespace cg {

namespace {
constexpr int BOX = 1024;

struct Box { float mnx, mny, mnz, mxx, mxy, mxz; };

__global__ void __launch_bounds__(256)
knn_bbox(int64_t P, const float* __restrict__ pts, float* __restrict__ bb /*6: min xyz, max xyz as ordered ints*/) {
  pdl_wait();
  float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < P; i += int64_t(gridDim.x) * blockDim.x)
    for (int k = 0; k < 3; ++k) { const float v = pts[3 * i + k]; mn[k] = fminf(mn[k], v); mx[k] = fmaxf(mx[k], v); }
  for (int k = 0; k < 3; ++k) {
    for (int o = 16; o > 0; o >>= 1) {
      mn[k] = fminf(mn[k], __shfl_xor_sync(0xffffffffu, mn[k], o));
      mx[k] = fmaxf(mx[k], __shfl_xor_sync(0xffffffffu, mx[k], o));
    }
  }
  if ((threadIdx.x & 31) == 0) {
    // float atomic min/max through the order-preserving int mapping
    for (int k = 0; k < 3; ++k) {
      int a = __float_as_int(mn[k]); a = a >= 0 ? a : a ^ 0x7fffffff;
      int b = __float_as_int(mx[k]); b = b >= 0 ? b : b ^ 0x7fffffff;
      atomicMin(reinterpret_cast<int*>(bb) + k, a);
      atomicMax(reinterpret_cast<int*>(bb) + 3 + k, b);
    }
  }
}
__device__ __forceinline__ float unmap(int a) { return __int_as_float(a >= 0 ? a : a ^ 0x7fffffff); }

__device__ __forceinline__ uint32_t spread10(uint32_t x) {
  x = (x | (x << 16)) & 0x030000FF;
  x = (x | (x << 8)) & 0x0300F00F;
  x = (x | (x << 4)) & 0x030C30C3;
  x = (x | (x << 2)) & 0x09249249;
  return x;
}

__global__ void __launch_bounds__(256)
knn_morton(int64_t P, const float* __restrict__ pts, const float* __restrict__ bb, uint64_t* __restrict__ keys,
           uint32_t* __restrict__ vals) {
  pdl_wait();
  const int64_t i = int64_t(blockIdx.x) * 256 + threadIdx.x;
  if (i >= P) return;
  const int* bi = reinterpret_cast<const int*>(bb);
  uint32_t code = 0;
  for (int k = 0; k < 3; ++k) {
    const float lo = unmap(bi[k]), hi = unmap(bi[3 + k]);
    const float ext = hi - lo;
    float u = ext > 0.f ? (pts[3 * i + k] - lo) / ext : 0.f;
    u = fminf(fmaxf(u, 0.f), 1.f);
    code |= spread10(uint32_t(u * 1023.f)) << k;
  }
  keys[i] = code;
  vals[i] = uint32_t(i);
}

__global__ void __launch_bounds__(BOX)
knn_boxes(int64_t P, const float* __restrict__ pts, const uint32_t* __restrict__ order, Box* __restrict__ boxes) {
  pdl_wait();
  __shared__ float red[6][32];
  const int64_t i = int64_t(blockIdx.x) * BOX + threadIdx.x;
  float v[6] = {FLT_MAX, FLT_MAX, FLT_MAX, -FLT_MAX, -FLT_MAX, -FLT_MAX};
  if (i < P) {
    const uint32_t id = order[i];
    for (int k = 0; k < 3; ++k) v[k] = v[3 + k] = pts[3 * size_t(id) + k];
  }
  for (int k = 0; k < 6; ++k)
    for (int o = 16; o > 0; o >>= 1) {
      const float n = __shfl_xor_sync(0xffffffffu, v[k], o);
      v[k] = k < 3 ? fminf(v[k], n) : fmaxf(v[k], n);
    }
  if ((threadIdx.x & 31) == 0)
    for (int k = 0; k < 6; ++k) red[k][threadIdx.x >> 5] = v[k];
  __syncthreads();
  if (threadIdx.x < 32) {
    for (int k = 0; k < 6; ++k) {
      float t = red[k][threadIdx.x];
      for (int o = 16; o > 0; o >>= 1) {
        const float n = __shfl_xor_sync(0xffffffffu, t, o);
        t = k < 3 ? fminf(t, n) : fmaxf(t, n);
      }
      v[k] = t;
    }
    if (threadIdx.x == 0) boxes[blockIdx.x] = Box{v[0], v[1], v[2], v[3], v[4], v[5]};
  }
}

__device__ __forceinline__ void push3(float d, float best[3]) {
#pragma unroll
  for (int j = 0; j < 3; ++j)
    if (best[j] > d) { const float t = best[j]; best[j] = d; d = t; }
}
__device__ __forceinline__ float dist2(float ax, float ay, float az, float bx, float by, float bz) {
  const float dx = bx - ax, dy = by - ay, dz = bz - az;
  return __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));   // dx*dx + dy*dy + dz*dz as the reference build fuses it
}
__device__ __forceinline__ float box_dist2(const Box& b, float x, float y, float z) {
  float dx = 0.f, dy = 0.f, dz = 0.f;
  if (x < b.mnx || x > b.mxx) dx = fminf(fabsf(x - b.mnx), fabsf(x - b.mxx));
  if (y < b.mny || y > b.mxy) dy = fminf(fabsf(y - b.mny), fabsf(y - b.mxy));
  if (z < b.mnz || z > b.mxz) dz = fminf(fabsf(z - b.mnz), fabsf(z - b.mxz));
  return dx * dx + dy * dy + dz * dz;
}

__global__ void __launch_bounds__(256)
knn_search(int64_t P, const float* __restrict__ pts, const uint32_t* __restrict__ order, const Box* __restrict__ boxes,
           float* __restrict__ out) {
  pdl_wait();
  const int64_t i = int64_t(blockIdx.x) * 256 + threadIdx.x;
  if (i >= P) return;
  const uint32_t me = order[i];
  const float x = pts[3 * size_t(me)], y = pts[3 * size_t(me) + 1], z = pts[3 * size_t(me) + 2];
  float best[3] = {FLT_MAX, FLT_MAX, FLT_MAX};
  // seed the rejection radius with the Morton-order neighbours
  const int64_t lo = i - 3 > 0 ? i - 3 : 0, hi = i + 3 < P - 1 ? i + 3 : P - 1;
  for (int64_t j = lo; j <= hi; ++j) {
    if (j == i) continue;
    const uint32_t o = order[j];
    push3(dist2(x, y, z, pts[3 * size_t(o)], pts[3 * size_t(o) + 1], pts[3 * size_t(o) + 2]), best);
  }
  const float reject = best[2];
  best[0] = best[1] = best[2] = FLT_MAX;
  const int64_t nbox = (P + BOX - 1) / BOX;
  for (int64_t b = 0; b < nbox; ++b) {
    const float bd = box_dist2(boxes[b], x, y, z);
    if (bd > reject || bd > best[2]) continue;
    const int64_t e = (b + 1) * BOX < P ? (b + 1) * BOX : P;
    for (int64_t j = b * BOX; j < e; ++j) {
      if (j == i) continue;
      const uint32_t o = order[j];
      push3(dist2(x, y, z, pts[3 * size_t(o)], pts[3 * size_t(o) + 1], pts[3 * size_t(o) + 2]), best);
    }
  }
  out[me] = (best[0] + best[1] + best[2]) / 3.0f;
}

struct KnnScratch {
  float* bb;
  Box* boxes;
  void* sort;
  static KnnScratch carve(void* base, int64_t P, size_t* bytes) {
    Carver c(base);
    KnnScratch k;
    k.bb = c.take<float>(32);
    k.boxes = c.take<Box>((P + BOX - 1) / BOX + 1);
    Carver sc(nullptr);
    SortBufs<uint64_t>::carve(sc, P);
    k.sort = c.take<char>((sc.used + 127) & ~size_t(127));
    if (bytes) *bytes = (c.used + 127) & ~size_t(127);
    return k;
  }
};
}  // namespace
}  // namespace cg

using namespace cg;

extern "C" {

size_t cg_knn_scratch_bytes(int64_t P) {
  size_t b = 0;
  KnnScratch::carve(nullptr, P < 1 ? 1 : P, &b);
  return b;
}

int cg_knn_mean_dist2(int64_t P, const float* points, float* mean_dist2, void* scratch, void* stream) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (P == 0) return CG_OK;
  CG_ARG(P > 0 && P < (int64_t(1) << 30), "P");
  CG_ARG(points && mean_dist2 && scratch, "knn pointers");
  CG_ARG((reinterpret_cast<uintptr_t>(scratch) & 127u) == 0, "scratch must be 128-byte aligned");
  KnnScratch ks = KnnScratch::carve(scratch, P, nullptr);
  Carver sc(ks.sort);
  SortBufs<uint64_t> bs = SortBufs<uint64_t>::carve(sc, P);
  // bbox init: +max / -max in the ordered-int domain
  const int init[6] = {0x7f7fffff, 0x7f7fffff, 0x7f7fffff, int(0xff7fffffu ^ 0x7fffffffu), int(0xff7fffffu ^ 0x7fffffffu),
                       int(0xff7fffffu ^ 0x7fffffffu)};
  CG_CUDA(cudaMemcpyAsync(ks.bb, init, sizeof(init), cudaMemcpyHostToDevice, st));
  const int rb = int((P + 255) / 256 < 1184 ? (P + 255) / 256 : 1184);
  StageTimer t_(ST_KNN, st, 4);
  launch_k(knn_bbox, dim3(rb), dim3(256), 0, st, P, points, ks.bb);
  CG_LAUNCH_CHECK(0, st);
  launch_k(knn_morton, dim3(unsigned((P + 255) / 256)), dim3(256), 0, st, P, points, ks.bb, bs.keys[0], bs.vals[0]);
  CG_LAUNCH_CHECK(0, st);
  int cur = 0;
  int rc = radix_sort_pairs<uint64_t>(bs, P, 32, &cur, false, st);
  if (rc != CG_OK) return rc;
  const unsigned nbox = unsigned((P + BOX - 1) / BOX);
  launch_k(knn_boxes, dim3(nbox), dim3(BOX), 0, st, P, points, bs.vals[cur], ks.boxes);
  CG_LAUNCH_CHECK(0, st);
  launch_k(knn_search, dim3(unsigned((P + 255) / 256)), dim3(256), 0, st, P, points, bs.vals[cur], ks.boxes, mean_dist2);
  CG_LAUNCH_CHECK(0, st);
  return CG_OK;
}

}  // extern "C"
