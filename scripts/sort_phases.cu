// Per-phase clock breakdown of sort_onesweep_pass on the binning workload (R pairs, 13 key bits).
// nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a scripts/sort_phases.cu -o /tmp/sort_phases && /tmp/sort_phases
#define CG_SORT_TIMING 1
#include "../curve_gaussian_b200/csrc/sort.cu"
#include <vector>
#include <random>
#include <algorithm>
namespace cg {
static thread_local char g_err2[256];
void set_error(const char* fmt, ...) { (void)fmt; }
void count_launches(int) {}
StageTimer::StageTimer(int, cudaStream_t, int) : stage(0), st(nullptr), rec(nullptr) {}
StageTimer::~StageTimer() {}
}
using namespace cg;
int main(int argc, char** argv) {
  const int64_t R = argc > 1 ? atoll(argv[1]) : 7600000;
  const int bits = argc > 2 ? atoi(argv[2]) : 13;
  std::vector<uint32_t> hk(R), hv(R);
  std::mt19937 rng(1);
  for (int64_t i = 0; i < R; ++i) { hk[i] = rng() % 8160; hv[i] = uint32_t(i); }
  Carver c0(nullptr);
  SortBufs<uint32_t>::carve(c0, R);
  void* base; cudaMalloc(&base, c0.used + 256);
  Carver c(base);
  SortBufs<uint32_t> b = SortBufs<uint32_t>::carve(c, R);
  const int64_t ntiles = (R + SORT_TILE - 1) / SORT_TILE;
  long long* ticks; cudaMalloc(&ticks, ntiles * 8 * sizeof(long long));
  std::vector<long long> ht(ntiles * 8);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int rep = 0; rep < 3; ++rep) {
    cudaMemcpy(b.keys[0], hk.data(), R * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(b.vals[0], hv.data(), R * 4, cudaMemcpyHostToDevice);
    long long* tp = (rep == 2) ? ticks : nullptr;
    cudaMemcpyToSymbol(g_sort_ticks, &tp, sizeof(tp));
    int cur = 0;
    cudaEventRecord(e0);
    int rc = radix_sort_pairs<uint32_t>(b, R, bits, &cur, false, 0);
    cudaEventRecord(e1); cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("rep %d rc %d total sort %.1f us (err %s)\n", rep, rc, ms * 1e3, cudaGetErrorString(cudaGetLastError()));
    if (rep == 2) {
      // ticks hold the LAST pass only (each pass overwrites)
      cudaMemcpy(ht.data(), ticks, ntiles * 8 * sizeof(long long), cudaMemcpyDeviceToHost);
      double ph[6] = {0, 0, 0, 0, 0, 0}; long long tmin = ht[0], tmax = 0; double life = 0;
      for (int64_t t = 0; t < ntiles; ++t) {
        for (int p = 0; p < 6; ++p) ph[p] += double(ht[t * 8 + p + 1] - ht[t * 8 + p]);
        life += double(ht[t * 8 + 6] - ht[t * 8]);
        tmin = std::min(tmin, ht[t * 8]); tmax = std::max(tmax, ht[t * 8 + 6]);
      }
      const char* names[6] = {"issue loads", "rank(+load wait)", "digit scan+publish", "look-back", "smem scatter", "write-out"};
      for (int p = 0; p < 6; ++p) printf("  %-20s %9.0f cycles/tile\n", names[p], ph[p] / ntiles);
      printf("  tile lifetime %.0f cycles avg; kernel span %lld cycles; tiles %lld\n", life / ntiles, tmax - tmin, (long long)ntiles);
      // verify
      std::vector<uint32_t> ok(R), ov(R);
      cudaMemcpy(ok.data(), b.keys[cur], R * 4, cudaMemcpyDeviceToHost);
      cudaMemcpy(ov.data(), b.vals[cur], R * 4, cudaMemcpyDeviceToHost);
      bool good = true;
      for (int64_t i = 1; i < R && good; ++i) good = ok[i - 1] < ok[i] || (ok[i - 1] == ok[i] && ov[i - 1] < ov[i]);
      printf("  sorted+stable: %s\n", good ? "yes" : "NO");
    }
  }
  return 0;
}
