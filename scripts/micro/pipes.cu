// Dev microbenchmark (not product): issue cost of SHFL / LDS.128 (lane-divergent, conflict-free) / FFMA / FFMA2 /
// MUFU on sm_100a, full chip, 32 warps per SM. Prints SM-cycles per warp-instruction per SM sub-partition.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>

#define ITERS 2048

template <int NSHFL, int NLDS, int NFMA, int NFMA2, int NMUFU>
__global__ void __launch_bounds__(256) k(float* out, int src_lane_delta, long long* cyc) {
  const long long t0 = clock64();
  __shared__ float4 s[8][64];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int i = lane; i < 64; i += 32) s[w][i] = make_float4(i, 1, 2, 3);
  __syncwarp();
  float a0 = threadIdx.x, a1 = 1.0001f, a2 = 0.5f, a3 = 0.25f;
  float v[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = lane + i;
  uint64_t p0, p1 = 0x3f8000003f800000ull, p2 = 0x3a0000003a000000ull;
  p0 = (uint64_t(__float_as_uint(a0)) << 32) | __float_as_uint(a1);
  const int src = (lane + src_lane_delta) & 31;
  int idx = lane;
#pragma unroll 1
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < NSHFL; ++i) v[i & 7] = __shfl_sync(0xffffffffu, v[i & 7], src);
#pragma unroll
    for (int i = 0; i < NLDS; ++i) {
      float4 t = s[w][(idx + i) & 63];
      v[i & 7] += (t.x + t.y) + (t.z + t.w);
      idx = (idx + 1) & 63;
    }
#pragma unroll
    for (int i = 0; i < NFMA; ++i) v[i & 7] = __fmaf_rn(v[i & 7], a1, a2);
#pragma unroll
    for (int i = 0; i < NFMA2; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p0) : "l"(p1), "l"(p2));
#pragma unroll
    for (int i = 0; i < NMUFU; ++i) v[i & 7] = __expf(v[i & 7] * 1e-3f);
  }
  float r = a3;
#pragma unroll
  for (int i = 0; i < 8; ++i) r += v[i];
  r += float(p0 & 0xffff);
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
  if (threadIdx.x == 0) cyc[blockIdx.x] = clock64() - t0;
}

template <int NSHFL, int NLDS, int NFMA, int NFMA2, int NMUFU>
void run(const char* name, float* out, int sms, double mhz) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int grid = sms * 4;   // 4 CTAs x 8 warps = 32 warps per SM
  static long long* d_cyc = nullptr; if (!d_cyc) cudaMalloc(&d_cyc, grid * 8);
  k<NSHFL, NLDS, NFMA, NFMA2, NMUFU><<<grid, 256>>>(out, 31, d_cyc);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  k<NSHFL, NLDS, NFMA, NFMA2, NMUFU><<<grid, 256>>>(out, 31, d_cyc);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double n_inst = double(NSHFL + NLDS + NFMA + NFMA2 + NMUFU);
  const double warp_inst_per_smsp = 8.0 * ITERS * n_inst;   // 32 warps per SM / 4 SMSP = 8 warps per SMSP
  const double cyc = ms * 1e-3 * mhz * 1e6;
  static long long h[4096]; cudaMemcpy(h, d_cyc, grid * 8, cudaMemcpyDeviceToHost); double avg = 0; for (int i = 0; i < grid; ++i) avg += h[i]; avg /= grid;
  printf("%-28s ms=%.4f  cyc/inst/SMSP: by event %.3f, by clock64 %.3f (eff clock %.0f MHz)\n", name, ms, cyc / warp_inst_per_smsp, avg / warp_inst_per_smsp, avg / (ms * 1e3));
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  const double mhz = khz / 1000.0;
  float* out; cudaMalloc(&out, size_t(p.multiProcessorCount) * 4 * 256 * 4);
  printf("%s SMs=%d clock=%.0f MHz\n", p.name, p.multiProcessorCount, mhz);
  run<8, 0, 0, 0, 0>("shfl x8", out, p.multiProcessorCount, mhz);
  run<0, 8, 0, 0, 0>("lds128 x8 (divergent)", out, p.multiProcessorCount, mhz);
  run<0, 0, 16, 0, 0>("ffma x16", out, p.multiProcessorCount, mhz);
  run<0, 0, 0, 16, 0>("ffma2 x16", out, p.multiProcessorCount, mhz);
  run<0, 0, 0, 0, 8>("mufu.ex2(+fmul) x8", out, p.multiProcessorCount, mhz);
  run<1, 0, 0, 0, 0>("shfl x1", out, p.multiProcessorCount, mhz);
  run<0, 1, 0, 0, 0>("lds128 x1", out, p.multiProcessorCount, mhz);
  run<4, 0, 16, 0, 0>("4shfl 16ffma", out, p.multiProcessorCount, mhz);
  run<0, 2, 16, 0, 0>("2lds 16ffma", out, p.multiProcessorCount, mhz);
  run<0, 0, 8, 8, 0>("8ffma 8ffma2", out, p.multiProcessorCount, mhz);
  run<2, 3, 40, 0, 2>("mix 2shfl 3lds 40ffma 2ex", out, p.multiProcessorCount, mhz);
  run<6, 1, 40, 0, 2>("mix 6shfl 1lds 40ffma 2ex", out, p.multiProcessorCount, mhz);
  run<0, 0, 40, 0, 2>("mix 0shfl 0lds 40ffma 2ex", out, p.multiProcessorCount, mhz);
  return 0;
}
