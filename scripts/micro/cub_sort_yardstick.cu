// Dev yardstick (not product, not linked into libcurvegs): cub::DeviceRadixSort::SortPairs on the same shape of
// data the tile sort sees at C4 - R (u32 tile key < 8160, u32 Gaussian id) pairs, 13 key bits - and on the
// per-Gaussian depth sort (P pairs, 32 key bits). Prints microseconds per call (CUDA events, 20 calls, L2-warm like
// the real step where emit_keys has just written the pairs).
#include <cub/cub.cuh>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

static float time_sort(uint32_t* k0, uint32_t* k1, uint32_t* v0, uint32_t* v1, int n, int end_bit, void* tmp, size_t tmp_bytes,
                       const uint32_t* src_k, const uint32_t* src_v) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  float total = 0.f;
  for (int it = 0; it < 23; ++it) {
    cudaMemcpyAsync(k0, src_k, size_t(n) * 4, cudaMemcpyDeviceToDevice);
    cudaMemcpyAsync(v0, src_v, size_t(n) * 4, cudaMemcpyDeviceToDevice);
    cudaEventRecord(e0);
    cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, k0, k1, v0, v1, n, 0, end_bit);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (it >= 3) total += ms;
  }
  return total / 20.f * 1000.f;
}

int main(int argc, char** argv) {
  const int R = argc > 1 ? atoi(argv[1]) : 8068262, P = 1000000, tiles = 8160;
  std::vector<uint32_t> hk(R), hv(R);
  srand(1);
  for (int i = 0; i < R; ++i) { hk[i] = uint32_t(rand()) % tiles; hv[i] = uint32_t(rand()) % P; }
  uint32_t *sk, *sv, *k0, *k1, *v0, *v1;
  cudaMalloc(&sk, size_t(R) * 4); cudaMalloc(&sv, size_t(R) * 4);
  cudaMalloc(&k0, size_t(R) * 4); cudaMalloc(&k1, size_t(R) * 4); cudaMalloc(&v0, size_t(R) * 4); cudaMalloc(&v1, size_t(R) * 4);
  cudaMemcpy(sk, hk.data(), size_t(R) * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(sv, hv.data(), size_t(R) * 4, cudaMemcpyHostToDevice);
  size_t tmp_bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, k0, k1, v0, v1, R, 0, 32);
  void* tmp; cudaMalloc(&tmp, tmp_bytes);
  printf("cub SortPairs R=%d 13 bits: %.1f us\n", R, time_sort(k0, k1, v0, v1, R, 13, tmp, tmp_bytes, sk, sv));
  printf("cub SortPairs R=%d 16 bits: %.1f us\n", R, time_sort(k0, k1, v0, v1, R, 16, tmp, tmp_bytes, sk, sv));
  for (int i = 0; i < P; ++i) hk[i] = uint32_t(rand()) * 65536u + uint32_t(rand());
  cudaMemcpy(sk, hk.data(), size_t(P) * 4, cudaMemcpyHostToDevice);
  printf("cub SortPairs P=%d 32 bits: %.1f us\n", P, time_sort(k0, k1, v0, v1, P, 32, tmp, tmp_bytes, sk, sv));
  return 0;
}
