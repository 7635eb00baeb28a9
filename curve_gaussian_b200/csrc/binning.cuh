// Tile binning without an R-sized sort (rasterizer_impl.cu:70-138, :309-325 semantics: per tile, the Gaussians that
// overlap it, ordered by (depth bits, Gaussian index)).
//
// The reference sorts one 64-bit key per (tile, Gaussian) instance; round 1 of this library sorted the Gaussians by
// depth first and then the R instances by tile alone (two 16-byte passes over R). Here the R instances are never
// sorted at all. A curve Gaussian covers a handful of neighbouring tiles, so the image is cut into SUPER-TILES of
// 8 x 8 tiles and
//   1. every depth-ordered Gaussian is duplicated once per super-tile it overlaps (about 1.5 copies instead of 8
//      instances) and the copies are stably sorted by super-tile index: one small radix pass (emit_keys / sort.cu);
//   2. a super-tile's list is cut into chunks of BIN_CHUNK copies, one warp each. A copy's coverage of its
//      super-tile is a 64-bit mask (a clipped rectangle); lane t of the warp owns tiles t and t + 32 and counts
//      the copies of the chunk that cover them (bin_count);
//   3. per super-tile and tile the chunk counts are scanned (bin_scan_chunks) and the tile totals are scanned in
//      tile order into the tile ranges (bin_scan_tiles);
//   4. the warps walk their chunks again, in list order, and lane t appends the Gaussian index of every copy that
//      covers its tile to that tile's segment of the point list (bin_fill).
// A tile's segment is filled chunk after chunk and, inside a chunk, copy after copy, i.e. in (depth, index) order:
// the same permutation as the reference's sort, with 1/5 of the memory traffic of the two R-sized passes.
#pragma once
#include "common.cuh"

namespace cg {

// number of super-tiles a tile rect overlaps, and the rect in super-tile units
__device__ __forceinline__ uint2 st_rect(uint2 rc, int shift) {
  const uint32_t mnx = (rc.x & 0xffffu) >> shift, mny = (rc.x >> 16) >> shift;
  const uint32_t add = (1u << shift) - 1u;
  const uint32_t mxx = ((rc.y & 0xffffu) + add) >> shift, mxy = ((rc.y >> 16) + add) >> shift;
  return make_uint2(mnx | (mny << 16), mxx | (mxy << 16));
}
__device__ __forceinline__ uint32_t rect_area(uint2 rc) {
  return ((rc.y & 0xffffu) - (rc.x & 0xffffu)) * ((rc.y >> 16) - (rc.x >> 16));
}

// coverage of the super-tile whose first tile is (ox, oy) by a tile rect: bit ty * 8 + tx
__device__ __forceinline__ uint64_t st_mask(uint2 rc, uint32_t ox, uint32_t oy) {
  const int x0 = max(int(rc.x & 0xffffu), int(ox)) - int(ox), x1 = min(int(rc.y & 0xffffu), int(ox) + ST_SIDE) - int(ox);
  const int y0 = max(int(rc.x >> 16), int(oy)) - int(oy), y1 = min(int(rc.y >> 16), int(oy) + ST_SIDE) - int(oy);
  if (x1 <= x0 || y1 <= y0) return 0ull;
  const uint64_t row = uint64_t(((1u << (x1 - x0)) - 1u) << x0);
  const uint64_t all = row * 0x0101010101010101ull;
  const uint64_t upto = y1 >= 8 ? ~0ull : ((1ull << (8 * y1)) - 1ull);
  const uint64_t from = (1ull << (8 * y0)) - 1ull;
  return all & upto & ~from;
}

// chunk_start[s] = first chunk of super-tile s; chunk_start[ns] = number of chunks. One CTA.
// digit_base != NULL (at most 256 super-tiles, i.e. a single sort pass whose digit IS the super-tile index): the
// super-tile spans are read off the sort's exclusive digit bases (n_copies = total, clamped to cap) and written to
// st_ranges here, instead of being recovered from the sorted keys by tile_ranges.
__global__ void __launch_bounds__(1024)
bin_chunk_table(uint32_t ns, uint2* __restrict__ st_ranges, uint32_t* __restrict__ chunk_start,
                const uint32_t* __restrict__ digit_base, const uint32_t* __restrict__ n_copies, uint32_t cap) {
  pdl_wait();
  if (digit_base) {
    const uint32_t n = min(*n_copies, cap);
    for (uint32_t i = threadIdx.x; i < ns; i += 1024) {
      const uint32_t a = digit_base[i], b = (i + 1 < 256u) ? digit_base[i + 1] : n;
      st_ranges[i] = make_uint2(a, b);
    }
    __syncthreads();
  }
  __shared__ uint32_t s_w[32];
  __shared__ uint32_t s_carry;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  const uint32_t lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (uint32_t base = 0; base < ns; base += 1024) {
    const uint32_t i = base + threadIdx.x;
    uint32_t v = 0;
    if (i < ns) { const uint2 r = st_ranges[i]; v = (r.y - r.x + BIN_CHUNK - 1) / BIN_CHUNK; }
    uint32_t inc = v;
    for (int o = 1; o < 32; o <<= 1) { const uint32_t n = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += n; }
    if (lane == 31) s_w[w] = inc;
    __syncthreads();
    if (w == 0) {
      const uint32_t x = s_w[lane];
      uint32_t xi = x;
      for (int o = 1; o < 32; o <<= 1) { const uint32_t n = __shfl_up_sync(0xffffffffu, xi, o); if (lane >= o) xi += n; }
      s_w[lane] = xi - x;
    }
    __syncthreads();
    const uint32_t carry = s_carry;
    if (i < ns) chunk_start[i] = carry + s_w[w] + inc - v;
    __syncthreads();
    if (threadIdx.x == 1023) s_carry = carry + s_w[w] + inc;
    __syncthreads();
  }
  if (threadIdx.x == 0) chunk_start[ns] = s_carry;
}

// Transpose of a 32 x 32 bit matrix held one row per lane (bit c of lane r's word -> bit r of lane c's word): five
// butterfly stages, each swapping the off-diagonal blocks of size j between lanes j apart.
__device__ __forceinline__ uint32_t warp_transpose32(uint32_t a, uint32_t lane) {
  uint32_t m = 0x0000ffffu;
#pragma unroll
  for (int j = 16; j > 0; j >>= 1) {
    const uint32_t p = __shfl_xor_sync(0xffffffffu, a, j);
    const bool hi = (lane & uint32_t(j)) != 0u;
    const uint32_t sh = hi ? (p >> j) : (p << j);
    const uint32_t keep = hi ? ~m : m;
    a = (a & keep) | (sh & ~keep);
    m ^= m << (j >> 1);
  }
  return a;
}

// What a CTA needs to walk its chunk: BIN_CHUNK consecutive copies of one super-tile, BIN_WCHUNK per warp.
struct BinChunk {
  uint32_t s, first, n, ox, oy;
  bool live;
};
__device__ __forceinline__ BinChunk bin_locate(uint32_t chunk, uint32_t ns, uint32_t sgx, const uint2* __restrict__ st_ranges,
                                               const uint32_t* __restrict__ chunk_start) {
  BinChunk c;
  c.live = chunk < __ldg(chunk_start + ns);
  c.s = 0; c.first = 0; c.n = 0; c.ox = 0; c.oy = 0;
  if (!c.live) return c;
  uint32_t lo = 0, hi = ns;   // largest s with chunk_start[s] <= chunk (empty super-tiles repeat the value: take the last)
  while (hi - lo > 1) {
    const uint32_t mid = (lo + hi) >> 1;
    if (__ldg(chunk_start + mid) <= chunk) lo = mid; else hi = mid;
  }
  c.s = lo;
  const uint2 r = st_ranges[lo];
  c.first = r.x + (chunk - __ldg(chunk_start + lo)) * BIN_CHUNK;
  c.n = min(uint32_t(BIN_CHUNK), r.y - c.first);
  c.ox = (lo % sgx) << ST_SHIFT;
  c.oy = (lo / sgx) << ST_SHIFT;
  return c;
}

// The warp's BIN_WCHUNK copies of the chunk as coverage sets per tile: after the transposes lane t holds in m0[e] /
// m1[e] the copies (bit k = copy e * 32 + k of the warp) that cover tiles t / t + 32 of the super-tile. ids (may be
// NULL) receives the Gaussian indices of the warp's copies.
__device__ __forceinline__ void bin_warp_sets(const BinChunk& c, uint32_t warp, uint32_t lane, const uint32_t* __restrict__ dup_list,
                                              const uint2* __restrict__ rect, uint32_t* ids, uint32_t (&m0)[BIN_EPL],
                                              uint32_t (&m1)[BIN_EPL]) {
  const uint32_t w0 = warp * BIN_WCHUNK;
#pragma unroll
  for (int e = 0; e < BIN_EPL; ++e) {
    const uint32_t j = w0 + e * 32 + lane;
    uint64_t m = 0;
    if (j < c.n) {
      const uint32_t id = dup_list[c.first + j];
      if (ids) ids[e * 32 + lane] = id;
      m = st_mask(rect[id], c.ox, c.oy);
    }
    m0[e] = uint32_t(m); m1[e] = uint32_t(m >> 32);
  }
#pragma unroll
  for (int e = 0; e < BIN_EPL; ++e) {
    const bool any = w0 + e * 32 < c.n;   // warp-uniform
    m0[e] = any ? warp_transpose32(m0[e], lane) : 0u;
    m1[e] = any ? warp_transpose32(m1[e], lane) : 0u;
  }
}

// ccnt[chunk][t] = copies of the chunk that cover tile t of its super-tile. One CTA per chunk.
__global__ void __launch_bounds__(256)
bin_count(uint32_t ns, uint32_t sgx, const uint2* __restrict__ st_ranges, const uint32_t* __restrict__ chunk_start,
          const uint32_t* __restrict__ dup_list, const uint2* __restrict__ rect, uint32_t* __restrict__ ccnt) {
  pdl_wait();
  __shared__ uint32_t s_cnt[64];
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // (a fixed grid strides over the chunks: their number is only known on the device)
  for (uint32_t chunk = blockIdx.x;; chunk += gridDim.x) {
    const BinChunk c = bin_locate(chunk, ns, sgx, st_ranges, chunk_start);
    if (!c.live) return;
    if (threadIdx.x < 64) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    uint32_t m0[BIN_EPL], m1[BIN_EPL];
    bin_warp_sets(c, warp, lane, dup_list, rect, nullptr, m0, m1);
    uint32_t c0 = 0, c1 = 0;
#pragma unroll
    for (int e = 0; e < BIN_EPL; ++e) { c0 += __popc(m0[e]); c1 += __popc(m1[e]); }
    if (c0) atomicAdd(&s_cnt[lane], c0);
    if (c1) atomicAdd(&s_cnt[32 + lane], c1);
    __syncthreads();
    if (threadIdx.x < 64) ccnt[size_t(chunk) * 64 + threadIdx.x] = s_cnt[threadIdx.x];
    __syncthreads();
  }
}

// One CTA of 64 threads per super-tile: cbase[chunk][t] = copies covering tile t in the super-tile's earlier
// chunks; tile_cnt[tile] = the tile's total.
__global__ void __launch_bounds__(64)
bin_scan_chunks(uint32_t sgx, int grid_x, int grid_y, const uint32_t* __restrict__ chunk_start,
                const uint32_t* __restrict__ ccnt, uint32_t* __restrict__ cbase, uint32_t* __restrict__ tile_cnt) {
  pdl_wait();
  const uint32_t s = blockIdx.x, t = threadIdx.x;
  const uint32_t c0 = chunk_start[s], c1 = chunk_start[s + 1];
  if (c0 == c1) return;   // (tile_cnt is zero-filled)
  uint32_t run = 0;
  // (16 independent loads per round trip: the kernel is as long as the chain of the largest super-tile)
  for (uint32_t c = c0; c < c1; c += 16) {
    uint32_t v[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) v[k] = (c + k < c1) ? ccnt[size_t(c + k) * 64 + t] : 0u;
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      if (c + k < c1) cbase[size_t(c + k) * 64 + t] = run;
      run += v[k];
    }
  }
  const uint32_t tx = ((s % sgx) << ST_SHIFT) + (t & 7u), ty = ((s / sgx) << ST_SHIFT) + (t >> 3);
  if (tx < uint32_t(grid_x) && ty < uint32_t(grid_y)) tile_cnt[ty * uint32_t(grid_x) + tx] = run;
}

// One CTA (bin_scan_tiles_order in raster_fwd.cu): exclusive scan of the tile totals in tile order -> ranges (empty tiles keep the reference's (0, 0)).
// cap bounds every range (capacity mode: an overflowing frame is flagged and re-run by the caller, but must stay
// inside its buffers).
__device__ __forceinline__ void bin_scan_tiles_body(uint32_t tiles, const uint32_t* __restrict__ tile_cnt, uint32_t cap,
                                                    uint2* __restrict__ ranges) {
  __shared__ uint32_t s_w[32];
  __shared__ uint32_t s_carry;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  const uint32_t lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (uint32_t base = 0; base < tiles; base += 4096) {
    const uint32_t i0 = base + threadIdx.x * 4;
    uint32_t v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) v[k] = (i0 + k < tiles) ? tile_cnt[i0 + k] : 0u;
    const uint32_t mine = v[0] + v[1] + v[2] + v[3];
    uint32_t inc = mine;
    for (int o = 1; o < 32; o <<= 1) { const uint32_t n = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += n; }
    if (lane == 31) s_w[w] = inc;
    __syncthreads();
    if (w == 0) {
      const uint32_t x = s_w[lane];
      uint32_t xi = x;
      for (int o = 1; o < 32; o <<= 1) { const uint32_t n = __shfl_up_sync(0xffffffffu, xi, o); if (lane >= o) xi += n; }
      s_w[lane] = xi - x;
    }
    __syncthreads();
    uint32_t run = s_carry + s_w[w] + inc - mine;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (i0 + k < tiles && v[k]) ranges[i0 + k] = make_uint2(min(run, cap), min(run + v[k], cap));
      run += v[k];
    }
    __syncthreads();
    if (threadIdx.x == 1023) s_carry = run;
    __syncthreads();
  }
}

// point_list[ranges[tile].x + cbase[chunk][t] + rank inside the chunk] = Gaussian index. One CTA per chunk: its
// instances (about six per copy) are first laid out in shared memory tile by tile, then copied out with consecutive
// threads on consecutive addresses: a tile's run of the chunk is ~90 entries long, so the stores fill whole sectors
// (the direct per-lane stores of 4 bytes each cost 32 sectors per instruction and 70 us at 8.8 M instances).
constexpr int BIN_STAGE = 8192;   // staged instances per CTA; a chunk with more writes straight to global memory
__global__ void __launch_bounds__(256)
bin_fill(uint32_t ns, uint32_t sgx, int grid_x, int grid_y, const uint2* __restrict__ st_ranges,
         const uint32_t* __restrict__ chunk_start, const uint32_t* __restrict__ dup_list, const uint2* __restrict__ rect,
         const uint32_t* __restrict__ cbase, const uint2* __restrict__ ranges, uint32_t cap,
         uint32_t* __restrict__ point_list) {
  pdl_wait();
  __shared__ uint16_t s_stage[BIN_STAGE];   // tile (6 bits) << 10 | copy index inside the chunk (10 bits)
  __shared__ uint32_t s_id[8][BIN_WCHUNK];
  static_assert(BIN_CHUNK <= 1024, "a staged entry holds a 10-bit copy index");
  __shared__ uint32_t s_wcnt[8][64];     // per warp and tile: count, then offset inside the tile's run
  __shared__ uint32_t s_loff[65];        // first staged slot of each tile
  __shared__ uint32_t s_goff[64];        // global position of staged slot i of tile t = s_goff[t] + i
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (uint32_t chunk = blockIdx.x;; chunk += gridDim.x) {   // (fixed grid: the number of chunks is only known on the device)
  const BinChunk c = bin_locate(chunk, ns, sgx, st_ranges, chunk_start);
  if (!c.live) return;
  uint32_t m0[BIN_EPL], m1[BIN_EPL];
  bin_warp_sets(c, warp, lane, dup_list, rect, s_id[warp], m0, m1);
  {
    uint32_t c0 = 0, c1 = 0;
#pragma unroll
    for (int e = 0; e < BIN_EPL; ++e) { c0 += __popc(m0[e]); c1 += __popc(m1[e]); }
    s_wcnt[warp][lane] = c0;
    s_wcnt[warp][32 + lane] = c1;
  }
  __syncthreads();
  if (threadIdx.x < 64) {
    // tile t: offsets of the warps inside its run, its run's first staged slot, its segment of the point list
    const uint32_t t = threadIdx.x;
    uint32_t tot = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) { const uint32_t v = s_wcnt[w][t]; s_wcnt[w][t] = tot; tot += v; }
    uint32_t inc = tot;
    for (int o = 1; o < 32; o <<= 1) { const uint32_t n = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= uint32_t(o)) inc += n; }
    if (t == 31) s_loff[64] = inc;   // (borrowed: total of tiles 0..31)
    __syncwarp();
    asm volatile("bar.sync 1, 64;" ::: "memory");
    const uint32_t first = (t >= 32 ? s_loff[64] : 0u) + inc - tot;
    asm volatile("bar.sync 1, 64;" ::: "memory");
    s_loff[t] = first;
    if (t == 63) s_loff[64] = first + tot;
    const uint32_t tx = c.ox + (t & 7u), ty = c.oy + (t >> 3);
    uint32_t g = cap;
    if (tx < uint32_t(grid_x) && ty < uint32_t(grid_y)) g = ranges[ty * uint32_t(grid_x) + tx].x + cbase[size_t(chunk) * 64 + t];
    s_goff[t] = g - first;
  }
  __syncthreads();
  const uint32_t n_inst = s_loff[64];
  const bool staged = n_inst <= uint32_t(BIN_STAGE);
  // lane t appends, in list order, the copies that cover its tiles
  uint32_t p0 = s_loff[lane] + s_wcnt[warp][lane], p1 = s_loff[32 + lane] + s_wcnt[warp][32 + lane];
  const uint32_t g0 = s_goff[lane], g1 = s_goff[32 + lane];
#pragma unroll
  for (int e = 0; e < BIN_EPL; ++e) {
    const uint32_t* ids = s_id[warp] + e * 32;
    uint32_t a = m0[e], b = m1[e];
    while (a) {
      const uint32_t k = uint32_t(__ffs(int(a))) - 1u;
      a &= a - 1u;
      if (staged) s_stage[p0] = uint16_t((lane << 10) | (warp * BIN_WCHUNK + e * 32 + k));
      else if (g0 + p0 < cap) point_list[g0 + p0] = ids[k];
      ++p0;
    }
    while (b) {
      const uint32_t k = uint32_t(__ffs(int(b))) - 1u;
      b &= b - 1u;
      if (staged) s_stage[p1] = uint16_t(((32u + lane) << 10) | (warp * BIN_WCHUNK + e * 32 + k));
      else if (g1 + p1 < cap) point_list[g1 + p1] = ids[k];
      ++p1;
    }
  }
  __syncthreads();
  if (staged) {
    for (uint32_t i = threadIdx.x; i < n_inst; i += 256) {
      const uint32_t en = s_stage[i];
      const uint32_t g = s_goff[en >> 10] + i;
      if (g < cap) point_list[g] = (&s_id[0][0])[en & 1023u];
    }
  }
  __syncthreads();   // (the shared arrays are rewritten by the next chunk)
  }
}

// The reference's 64-bit sort keys, rebuilt for parity checks from the binned state: one CTA per tile.
__global__ void __launch_bounds__(256)
rebuild_keys_ranges(const uint2* __restrict__ ranges, const uint32_t* __restrict__ point_list,
                    const float* __restrict__ depth, uint64_t* __restrict__ keys) {
  pdl_wait();
  const uint2 r = ranges[blockIdx.x];
  for (uint32_t i = r.x + threadIdx.x; i < r.y; i += 256)
    keys[i] = (uint64_t(blockIdx.x) << 32) | __float_as_uint(depth[point_list[i]]);
}

}  // namespace cg
