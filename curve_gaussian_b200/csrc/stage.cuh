// Device helpers shared by the blend kernels (raster_fwd.cu, raster_bwd.cu): staging of AoS(3) arrays through
// shared memory with 128-bit global accesses, and the mbarrier + 1-D bulk async copy (TMA) primitives that
// stream a tile's sorted records into the shared ring.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace cg {

// Stage `n` consecutive floats (n <= capacity of sm) with 128-bit loads.
__device__ __forceinline__ void stage_floats(const float* __restrict__ src, float* sm, int n) {
  if ((reinterpret_cast<uintptr_t>(src) & 15u) == 0) {
    const int n4 = n >> 2;
    const float4* s4 = reinterpret_cast<const float4*>(src);
    for (int i = threadIdx.x; i < n4; i += blockDim.x) reinterpret_cast<float4*>(sm)[i] = __ldg(s4 + i);
    for (int i = (n4 << 2) + threadIdx.x; i < n; i += blockDim.x) sm[i] = __ldg(src + i);
  } else {
    for (int i = threadIdx.x; i < n; i += blockDim.x) sm[i] = __ldg(src + i);
  }
}
// ... and back: `n` consecutive floats from shared to global memory with 128-bit stores.
__device__ __forceinline__ void unstage_floats(float* __restrict__ dst, const float* sm, int n) {
  if ((reinterpret_cast<uintptr_t>(dst) & 15u) == 0) {
    const int n4 = n >> 2;
    float4* d4 = reinterpret_cast<float4*>(dst);
    for (int i = threadIdx.x; i < n4; i += blockDim.x) d4[i] = reinterpret_cast<const float4*>(sm)[i];
    for (int i = (n4 << 2) + threadIdx.x; i < n; i += blockDim.x) dst[i] = sm[i];
  } else {
    for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = sm[i];
  }
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return uint32_t(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

}  // namespace cg
