// Stable LSD radix sort of (uint32/uint64 key, uint32 value) pairs, 8 bits per pass,
// single-pass-per-digit ("onesweep") with decoupled look-back.
//
// Replaces cub::DeviceRadixSort::SortPairs in the reference binning step
// (rasterizer_impl.cu:309-314) and in simple-knn (simple_knn.cu:207-214). The sort
// must be STABLE: equal keys keep their input order, which is what makes point_list
// bit-exact against the reference (see BinScratch in common.cuh for how the binning
// step splits the reference's one 64-bit sort into a per-Gaussian depth sort and a
// per-instance tile sort with the same result).
//
// HBM traffic: one histogram pass over the keys + per digit pass one read and one
// write of every pair; digit passes whose histogram shows a single occupied bin are
// still run (ping-pong parity stays fixed), they stream at copy speed.
#include "common.cuh"

namespace cg {

namespace {

// Optional per-phase clock instrumentation (scripts/sort_phases.cu defines CG_SORT_TIMING); compiled out otherwise.
#ifdef CG_SORT_TIMING
__device__ long long* g_sort_ticks = nullptr;   // [tile][8]
#define CG_SORT_TICK(n) do { if (threadIdx.x == 0 && g_sort_ticks) g_sort_ticks[size_t(s.tile_id) * 8 + (n)] = clock64(); } while (0)
#else
#define CG_SORT_TICK(n) do { } while (0)
#endif

constexpr uint32_t FLAG_SHIFT = 30;
constexpr uint32_t VALUE_MASK = (1u << FLAG_SHIFT) - 1u;
constexpr uint32_t FLAG_AGG = 1u;   // tile-local count published
constexpr uint32_t FLAG_INC = 2u;   // inclusive prefix published

template <typename K>
__global__ void __launch_bounds__(256)
sort_histogram(const K* __restrict__ keys, int64_t R, const uint32_t* __restrict__ d_n, int passes, int bpp,
               uint32_t* __restrict__ hist) {
  pdl_wait();
  if (d_n) R = min(R, int64_t(*d_n));   // device-side count (capacity mode): R is only the capacity
  __shared__ uint32_t sh[SORT_MAX_PASSES * 256];
  for (int i = threadIdx.x; i < passes * 256; i += blockDim.x) sh[i] = 0;
  __syncthreads();
  int64_t stride = int64_t(gridDim.x) * blockDim.x;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < R; i += stride) {
    K k = keys[i];
    for (int p = 0; p < passes; ++p) atomicAdd(&sh[p * 256 + (uint32_t(k >> (bpp * p)) & ((1u << bpp) - 1u))], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < passes * 256; i += blockDim.x) {
    uint32_t v = sh[i];
    if (v) atomicAdd(&hist[i], v);
  }
}

// One CTA per pass: in-place exclusive scan of that pass's 256 bins.
__global__ void __launch_bounds__(256) sort_scan_bins(uint32_t* __restrict__ hist) {
  pdl_wait();
  __shared__ uint32_t wsum[8];
  uint32_t* h = hist + blockIdx.x * 256;
  uint32_t v = h[threadIdx.x];
  uint32_t lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  uint32_t inc = v;
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t n = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += n;
  }
  if (lane == 31) wsum[w] = inc;
  __syncthreads();
  uint32_t base = 0;
  for (uint32_t i = 0; i < w; ++i) base += wsum[i];
  h[threadIdx.x] = base + inc - v;
}

template <typename K>
struct SortSmem {
  K keys[SORT_TILE];
  uint32_t vals[SORT_TILE];
  uint32_t whist[SORT_THREADS / 32][257];  // per-warp digit counters (+1 bin for tail padding)
  uint32_t local_start[256];
  int32_t gbase[256];                      // global position = gbase[d] + local position
  uint32_t tile_id;
};

#ifndef CG_SORT_CTAS
#define CG_SORT_CTAS 3
#endif
template <typename K>
__global__ void __launch_bounds__(SORT_THREADS, sizeof(K) == 4 ? CG_SORT_CTAS : 2)
sort_onesweep_pass(const K* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
                   K* __restrict__ keys_out, uint32_t* __restrict__ vals_out,
                   int64_t R, const uint32_t* __restrict__ d_n, int shift, uint32_t dmask,
                   const uint32_t* __restrict__ digit_base, uint32_t* __restrict__ ticket,
                   uint32_t* __restrict__ status) {
  pdl_wait();
  if (d_n) R = min(R, int64_t(*d_n));   // device-side count: the grid covers the capacity, surplus CTAs leave below
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SortSmem<K>& s = *reinterpret_cast<SortSmem<K>*>(smem_raw);
  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int WARPS = SORT_THREADS / 32;

  // Dynamic tile id: a tile only ever waits on tiles that already started.
  if (tid == 0) s.tile_id = atomicAdd(ticket, 1u);
  __syncthreads();
  const uint32_t tile = s.tile_id;
  CG_SORT_TICK(0);
  const int64_t tile_base = int64_t(tile) * SORT_TILE;
  // (surplus CTAs of a grid sized for a capacity leave before they touch anything else)
  if (tile_base >= R) return;   // uniform: tickets are handed out in order, so every live tile's predecessors are live
  for (int i = tid; i < WARPS * 257; i += SORT_THREADS) (&s.whist[0][0])[i] = 0;
  __syncthreads();
  const int64_t warp_base = tile_base + int64_t(warp) * (32 * SORT_ITEMS);

  K k[SORT_ITEMS];
  uint32_t v[SORT_ITEMS];
  uint16_t rank[SORT_ITEMS];
#pragma unroll
  for (int i = 0; i < SORT_ITEMS; ++i) {
    int64_t idx = warp_base + i * 32 + lane;
    bool ok = idx < R;
    k[i] = ok ? keys_in[idx] : K(~K(0));
    v[i] = ok ? vals_in[idx] : 0u;
  }

  CG_SORT_TICK(1);   // (the loads are only waited for at first use, inside the ranking)
  // Stable in-warp ranking (items are in memory order). For each item the lanes holding the
  // same digit form a peer group (match.any); the group's first lane bumps the warp's digit
  // counter with ONE shared-memory atomic that returns the group's base. The atomics of
  // successive items to the same counter execute in issue order, so nothing here depends on
  // the previous item through registers: the 16 match/atomic/shuffle chains overlap instead
  // of forming one 16-deep serial chain.
  uint32_t* wh = s.whist[warp];
  const uint32_t lt_mask = (1u << lane) - 1u;
#pragma unroll
  for (int h = 0; h < SORT_ITEMS; h += 8) {
    uint32_t m[8], old[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int64_t idx = warp_base + (h + i) * 32 + lane;
      const uint32_t d = (idx < R) ? (uint32_t(k[h + i] >> shift) & dmask) : 256u;
      m[i] = __match_any_sync(0xffffffffu, d);
      old[i] = 0;
      if (lane == uint32_t(__ffs(int(m[i])) - 1)) old[i] = atomicAdd(&wh[d], uint32_t(__popc(m[i])));
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const uint32_t base = __shfl_sync(0xffffffffu, old[i], __ffs(int(m[i])) - 1);
      rank[h + i] = uint16_t(base + __popc(m[i] & lt_mask));
    }
  }
  __syncthreads();
  CG_SORT_TICK(2);

  // Thread d owns digit d: exclusive scan across warps, tile count, look-back.
  {
    const uint32_t d = tid;
    uint32_t total = 0;
#pragma unroll
    for (int w = 0; w < WARPS; ++w) {
      uint32_t c = s.whist[w][d];
      s.whist[w][d] = total;
      total += c;
    }
    uint32_t* st = status + size_t(tile) * 256 + d;
    const bool live = d <= dmask;   // bins above the digit width stay empty: no publish, no look-back
    if (live) {
      if (tile == 0) {
        atomicExch(st, (FLAG_INC << FLAG_SHIFT) | total);
      } else {
        atomicExch(st, (FLAG_AGG << FLAG_SHIFT) | total);
      }
    }
    // block-wide exclusive scan of `total` over digits -> local_start
    uint32_t inc = total;
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t n = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += n;
    }
    __shared__ uint32_t wsum[WARPS];
    if (lane == 31) wsum[warp] = inc;
    __syncthreads();
    uint32_t wb = 0;
    for (uint32_t i = 0; i < warp; ++i) wb += wsum[i];
    const uint32_t lstart = wb + inc - total;
    s.local_start[d] = lstart;

    CG_SORT_TICK(3);
    // Decoupled look-back, LB predecessors per round trip: all tiles of the first wave publish
    // their aggregates at about the same time, so a one-at-a-time walk ripples through
    // ~ntiles/2 dependent L2 round trips; independent loads of a window cut that by LB.
    uint32_t excl = 0;
    if (tile > 0 && live) {
#ifndef CG_SORT_LB
#define CG_SORT_LB 4
#endif
      constexpr int LB = CG_SORT_LB;
      int64_t t = int64_t(tile) - 1;
      bool done = false;
      while (!done) {
        uint32_t w[LB];
#pragma unroll
        for (int i = 0; i < LB; ++i) {
          const int64_t ti = t - i;
          w[i] = (ti >= 0) ? *reinterpret_cast<volatile const uint32_t*>(status + size_t(ti) * 256 + d)
                           : (FLAG_INC << FLAG_SHIFT);   // tile 0 always publishes FLAG_INC; this is never consumed
        }
        int used = 0;
#pragma unroll
        for (int i = 0; i < LB; ++i) {
          if (!done && used == i) {
            const uint32_t f = w[i] >> FLAG_SHIFT;
            if (f != 0u) {
              excl += w[i] & VALUE_MASK;
              used = i + 1;
              if (f == FLAG_INC) done = true;
            }
          }
        }
        t -= used;
#ifdef CG_SORT_BACKOFF
        if (!done && used == 0) __nanosleep(CG_SORT_BACKOFF);
#endif
      }
      atomicExch(st, (FLAG_INC << FLAG_SHIFT) | ((excl + total) & VALUE_MASK));
    }
    s.gbase[d] = int32_t(digit_base[d] + excl) - int32_t(lstart);
  }
  __syncthreads();
  CG_SORT_TICK(4);

  // Scatter into shared memory at the tile-local sorted position.
#pragma unroll
  for (int i = 0; i < SORT_ITEMS; ++i) {
    int64_t idx = warp_base + i * 32 + lane;
    if (idx < R) {
      uint32_t d = uint32_t(k[i] >> shift) & dmask;
      uint32_t pos = s.local_start[d] + s.whist[warp][d] + rank[i];
      s.keys[pos] = k[i];
      s.vals[pos] = v[i];
    }
  }
  __syncthreads();
  CG_SORT_TICK(5);

  // Coalesced write-out: consecutive threads write consecutive slots of a digit run.
  int64_t rem = R - tile_base;
  const int count = rem < SORT_TILE ? int(rem) : SORT_TILE;
  for (int j = tid; j < count; j += SORT_THREADS) {
    K key = s.keys[j];
    uint32_t d = uint32_t(key >> shift) & dmask;
    int64_t g = int64_t(s.gbase[d]) + j;
    keys_out[g] = key;
    vals_out[g] = s.vals[j];
  }
  CG_SORT_TICK(6);
}

// ---------------------------------------------------------------------------------------------------------------
// Small sorts (n <= SMALL_TILE pairs, e.g. the depth sort of a 5 k-Gaussian scene): ALL passes in one CTA. The keys
// never leave the SM between passes: rank (same stable match.any ranking as the onesweep pass), scatter into
// shared memory, read back in sorted order, next digit. One launch instead of histogram + scan + one launch per
// pass, each of which costs 8-10 us at this size because of launch latency and the look-back machinery alone.
constexpr int SMALL_THREADS = 512;
constexpr int SMALL_ITEMS = 16;
constexpr int SMALL_TILE = SMALL_THREADS * SMALL_ITEMS;   // 8192 pairs
template <typename K>
struct SmallSortSmem {
  K keys[SMALL_TILE];
  uint32_t vals[SMALL_TILE];
  uint32_t whist[SMALL_THREADS / 32][257];
  uint32_t local_start[256];
  uint32_t wsum[SMALL_THREADS / 32];
};

template <typename K>
__global__ void __launch_bounds__(SMALL_THREADS, 1)
sort_small(const K* __restrict__ keys_in, const uint32_t* __restrict__ vals_in, K* __restrict__ keys_out,
           uint32_t* __restrict__ vals_out, int64_t R, const uint32_t* __restrict__ d_n, int passes, int bpp) {
  pdl_wait();
  if (d_n) R = min(R, int64_t(*d_n));
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SmallSortSmem<K>& s = *reinterpret_cast<SmallSortSmem<K>*>(smem_raw);
  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int WARPS = SMALL_THREADS / 32;
  const uint32_t dmask = (1u << bpp) - 1u;
  const int64_t warp_base = int64_t(warp) * (32 * SMALL_ITEMS);
  const uint32_t lt_mask = (1u << lane) - 1u;

  K k[SMALL_ITEMS];
  uint32_t v[SMALL_ITEMS];
#pragma unroll
  for (int i = 0; i < SMALL_ITEMS; ++i) {
    const int64_t idx = warp_base + i * 32 + lane;
    const bool ok = idx < R;
    k[i] = ok ? keys_in[idx] : K(~K(0));
    v[i] = ok ? vals_in[idx] : 0u;
  }
  for (int p = 0; p < passes; ++p) {
    const int shift = bpp * p;
    for (int i = tid; i < WARPS * 257; i += SMALL_THREADS) (&s.whist[0][0])[i] = 0;
    __syncthreads();
    uint32_t* wh = s.whist[warp];
    uint16_t rank[SMALL_ITEMS];
#pragma unroll
    for (int h = 0; h < SMALL_ITEMS; h += 8) {
      uint32_t m[8], old[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int64_t idx = warp_base + (h + i) * 32 + lane;
        const uint32_t d = (idx < R) ? (uint32_t(k[h + i] >> shift) & dmask) : 256u;
        m[i] = __match_any_sync(0xffffffffu, d);
        old[i] = 0;
        if (lane == uint32_t(__ffs(int(m[i])) - 1)) old[i] = atomicAdd(&wh[d], uint32_t(__popc(m[i])));
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const uint32_t base = __shfl_sync(0xffffffffu, old[i], __ffs(int(m[i])) - 1);
        rank[h + i] = uint16_t(base + __popc(m[i] & lt_mask));
      }
    }
    __syncthreads();
    if (tid < 256) {
      // digit tid: exclusive scan of its counts across the warps, then of the digit totals across the digits
      uint32_t total = 0;
#pragma unroll
      for (int w = 0; w < WARPS; ++w) { const uint32_t c = s.whist[w][tid]; s.whist[w][tid] = total; total += c; }
      uint32_t inc = total;
      for (int o = 1; o < 32; o <<= 1) { const uint32_t n = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= uint32_t(o)) inc += n; }
      if (lane == 31) s.wsum[warp] = inc;
      asm volatile("bar.sync 1, 256;" ::: "memory");
      uint32_t wb = 0;
      for (uint32_t i = 0; i < warp; ++i) wb += s.wsum[i];
      s.local_start[tid] = wb + inc - total;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < SMALL_ITEMS; ++i) {
      const int64_t idx = warp_base + i * 32 + lane;
      if (idx < R) {
        const uint32_t d = uint32_t(k[i] >> shift) & dmask;
        const uint32_t pos = s.local_start[d] + s.whist[warp][d] + rank[i];
        s.keys[pos] = k[i];
        s.vals[pos] = v[i];
      }
    }
    __syncthreads();
    if (p + 1 < passes) {
#pragma unroll
      for (int i = 0; i < SMALL_ITEMS; ++i) {
        const int64_t idx = warp_base + i * 32 + lane;
        if (idx < R) { k[i] = s.keys[idx]; v[i] = s.vals[idx]; }
      }
      __syncthreads();
    }
  }
  for (int64_t j = tid; j < R; j += SMALL_THREADS) {
    keys_out[j] = s.keys[j];
    vals_out[j] = s.vals[j];
  }
}

}  // namespace

// Digit geometry of a sort of `end_bit` key bits: number of passes and (equal) bits per pass
// (e.g. 13 tile bits -> 2 x 7 instead of 8 + 5: fewer bins per pass means longer runs per bin in the scatter and
// fewer look-back chains).
void radix_sort_geometry(int end_bit, int key_bits, int* passes, int* bpp) {
  if (end_bit > key_bits) end_bit = key_bits;
  int p = (end_bit + 7) / 8;
  if (p < 1) p = 1;
  *passes = p;
  *bpp = end_bit <= 0 ? 1 : (end_bit + p - 1) / p;
}

// Clears the digit histograms, tickets and look-back words of a sort of n pairs. Called by radix_sort_pairs itself,
// or - ahead of it - by a producer that fills b.hist while it writes the keys (emit_keys) and then passes
// hist_ready = true.
template <typename K>
int radix_sort_begin(const SortBufs<K>& b, int64_t R, int end_bit, cudaStream_t stream) {
  if (R <= 0) return CG_OK;
  int passes, bpp;
  radix_sort_geometry(end_bit, int(8 * sizeof(K)), &passes, &bpp);
  const int64_t ntiles = (R + SORT_TILE - 1) / SORT_TILE;
  // hist, ticket and status are carved back to back: one memset covers all three
  CG_CUDA(cudaMemsetAsync(b.hist, 0,
                          size_t(reinterpret_cast<char*>(b.status + size_t(passes) * ntiles * 256) -
                                 reinterpret_cast<char*>(b.hist)), stream));
  return CG_OK;
}

template <typename K>
int radix_sort_pairs(const SortBufs<K>& b, int64_t R, int end_bit, int* out_buf, bool debug, cudaStream_t stream,
                     const uint32_t* d_n, bool hist_ready, uint32_t* final_vals) {
  *out_buf = 0;
  if (R <= 0) return CG_OK;
  if (R >= (int64_t(1) << FLAG_SHIFT)) {
    set_error("radix sort: %lld pairs exceed the 2^30 look-back word", (long long)R);
    return CG_ERR_CAPACITY;
  }
  int passes, bpp;
  radix_sort_geometry(end_bit, int(8 * sizeof(K)), &passes, &bpp);
  const uint32_t dmask = (1u << bpp) - 1u;
  const int64_t ntiles = (R + SORT_TILE - 1) / SORT_TILE;

  if (R <= SMALL_TILE) {
    // the whole sort in one CTA (the digit histograms a producer may have prepared are not needed)
    static thread_local bool small_attr[64] = {false};
    int sdev = 0;
    CG_CUDA(cudaGetDevice(&sdev));
    CG_ARG(sdev >= 0 && sdev < 64, "device ordinal");
    if (!small_attr[sdev]) {
      CG_CUDA(cudaFuncSetAttribute(sort_small<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(sizeof(SmallSortSmem<K>))));
      small_attr[sdev] = true;
    }
    const int res = passes & 1;   // where the multi-pass sort would leave the result
    if (hist_ready) {
      // the producer's digit counts become exclusive bases, as after the multi-pass sort (the super-tile binning reads
      // its spans from them)
      count_launches(1);
      launch_k(sort_scan_bins, dim3(passes), dim3(256), 0, stream, b.hist);
      CG_LAUNCH_CHECK(debug, stream);
    }
    count_launches(1);
    launch_k(sort_small<K>, dim3(1), dim3(SMALL_THREADS), sizeof(SmallSortSmem<K>), stream, b.keys[0], b.vals[0], b.keys[res],
             final_vals ? final_vals : b.vals[res], R, d_n, passes, bpp);
    CG_LAUNCH_CHECK(debug, stream);
    *out_buf = res;
    return CG_OK;
  }
  // (per device: a process that drives several GPUs sets the attribute on each of them)
  static thread_local bool attr_set[64] = {false};
  int dev = 0;
  CG_CUDA(cudaGetDevice(&dev));
  CG_ARG(dev >= 0 && dev < 64, "device ordinal");
  if (!attr_set[dev]) {
    CG_CUDA(cudaFuncSetAttribute(sort_onesweep_pass<K>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 int(sizeof(SortSmem<K>))));
    attr_set[dev] = true;
  }
  if (!hist_ready) {
    int rc = radix_sort_begin<K>(b, R, end_bit, stream);
    if (rc != CG_OK) return rc;
    int hist_blocks = int(ntiles < 148 * 8 ? ntiles : 148 * 8);
    count_launches(1);
    launch_k(sort_histogram<K>, dim3(hist_blocks), dim3(256), 0, stream, b.keys[0], R, d_n, passes, bpp, b.hist);
    CG_LAUNCH_CHECK(debug, stream);
  }
  count_launches(1 + passes);
  launch_k(sort_scan_bins, dim3(passes), dim3(256), 0, stream, b.hist);
  CG_LAUNCH_CHECK(debug, stream);

  int cur = 0;
  for (int p = 0; p < passes; ++p) {
    launch_k(sort_onesweep_pass<K>, dim3(unsigned(ntiles)), dim3(SORT_THREADS), sizeof(SortSmem<K>), stream, b.keys[cur], b.vals[cur], b.keys[cur ^ 1], (final_vals && p == passes - 1) ? final_vals : b.vals[cur ^ 1], R, d_n, bpp * p, dmask,
        b.hist + p * 256, b.ticket + p, b.status + size_t(p) * ntiles * 256);
    CG_LAUNCH_CHECK(debug, stream);
    cur ^= 1;
  }
  *out_buf = cur;
  return CG_OK;
}

template int radix_sort_pairs<uint32_t>(const SortBufs<uint32_t>&, int64_t, int, int*, bool, cudaStream_t, const uint32_t*, bool, uint32_t*);
template int radix_sort_pairs<uint64_t>(const SortBufs<uint64_t>&, int64_t, int, int*, bool, cudaStream_t, const uint32_t*, bool, uint32_t*);
template int radix_sort_begin<uint32_t>(const SortBufs<uint32_t>&, int64_t, int, cudaStream_t);
template int radix_sort_begin<uint64_t>(const SortBufs<uint64_t>&, int64_t, int, cudaStream_t);

}  // namespace cg
