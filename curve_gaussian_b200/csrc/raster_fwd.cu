// Forward rasterization kernels: EWA preprocess (+tile counts, block sums, depth-sort keys), block-sum scans,
// emission of (super-tile, Gaussian) copies or (tile, Gaussian) instances, tile ranges, and the per-tile
// front-to-back blend (TMA bulk copies of the tile's id list, per-record async gathers). The tile binning itself
// is in binning.cuh, the radix sort in sort.cu.
//
// Reference behaviour followed (semantics, not structure):
//   preprocess   forward.cu:155-274, auxiliary.h:40-55,151-176
//   scan + R     rasterizer_impl.cu:283-287
//   key emission rasterizer_impl.cu:70-111
//   ranges       rasterizer_impl.cu:116-138
//   blend        forward.cu:279-417
#include <cstring>
#include "common.cuh"
#include "math.cuh"
#include "stage.cuh"
#include "binning.cuh"

namespace cg {

__device__ __forceinline__ float ndc_to_pix(float v, int S) {
  // double arithmetic on purpose: the reference's literals are doubles (auxiliary.h:40-43)
  return float(__dmul_rn(__fma_rn(__dadd_rn(double(v), 1.0), double(S), -1.0), 0.5));
}

__global__ void __launch_bounds__(256)
preprocess_fwd(int64_t P, const float* __restrict__ means3D, const float* __restrict__ opacities,
               const float* __restrict__ scales, const float* __restrict__ rotations,
               const float* __restrict__ cov3D_precomp, float scale_modifier,
               const float* __restrict__ viewmatrix, const float* __restrict__ projmatrix,
               int W, int H, float tanx, float tany, float fx, float fy, int grid_x, int grid_y,
               int antialiasing, const float* __restrict__ colors, const float* __restrict__ all_map,
               int32_t* __restrict__ radii, GeomState g) {
  pdl_wait();
  __shared__ __align__(16) float s_mean[768];
  __shared__ __align__(16) float s_scale[768];
  __shared__ float s_vm[16], s_pm[16];
  __shared__ uint32_t s_wsum[8];

  const int64_t blk0 = int64_t(blockIdx.x) * 256;
  const int nhere = int(P - blk0 < 256 ? P - blk0 : 256);
  stage_floats(means3D + blk0 * 3, s_mean, nhere * 3);
  if (scales) stage_floats(scales + blk0 * 3, s_scale, nhere * 3);
  if (threadIdx.x < 16) s_vm[threadIdx.x] = viewmatrix[threadIdx.x];
  else if (threadIdx.x < 32) s_pm[threadIdx.x - 16] = projmatrix[threadIdx.x - 16];
  __syncthreads();

  const int64_t idx = blk0 + threadIdx.x;
  uint32_t touched = 0, depth_key = 0xffffffffu;
  uint2 rect_out = make_uint2(0u, 0u);   // (an empty rect for a culled Gaussian: the binning reads the rect alone)
  int radius_out = 0;
  if (idx < P) {
    const float px = s_mean[3 * threadIdx.x], py = s_mean[3 * threadIdx.x + 1], pz = s_mean[3 * threadIdx.x + 2];
    const float3 p_view = xform43(px, py, pz, s_vm);
    if (!(p_view.z <= 0.2f)) {  // near cull only, no x/y frustum test (auxiliary.h:166)
      const float4 p_hom = xform44(px, py, pz, s_pm);
      const float p_w = __frcp_rn(__fadd_rn(p_hom.w, 0.0000001f));
      const float ndc_x = __fmul_rn(p_hom.x, p_w), ndc_y = __fmul_rn(p_hom.y, p_w);

      float cov6[6];
      if (cov3D_precomp) {
#pragma unroll
        for (int i = 0; i < 6; ++i) cov6[i] = cov3D_precomp[idx * 6 + i];
      } else {
        float4 q;
        if ((reinterpret_cast<uintptr_t>(rotations) & 15u) == 0) q = __ldg(reinterpret_cast<const float4*>(rotations) + idx);
        else q = make_float4(rotations[4 * idx], rotations[4 * idx + 1], rotations[4 * idx + 2], rotations[4 * idx + 3]);
        cov3d_from_scale_rot(s_scale[3 * threadIdx.x], s_scale[3 * threadIdx.x + 1], s_scale[3 * threadIdx.x + 2],
                             scale_modifier, q, cov6);
      }
      Proj2D pr = project_cov(px, py, pz, fx, fy, tanx, tany, cov6, s_vm);
      float3 cov = pr.cov;
      constexpr float h_var = 0.3f;
      // a*c - b*b with one shared rounded b*b (see math.cuh on pinned contraction)
      const float bb = __fmul_rn(cov.y, cov.y);
      const float det_cov = __fmaf_rn(cov.x, cov.z, -bb);
      cov.x = __fadd_rn(cov.x, h_var);
      cov.z = __fadd_rn(cov.z, h_var);
      const float det = __fmaf_rn(cov.x, cov.z, -bb);
      float h_scaling = 1.0f;
      if (antialiasing) h_scaling = sqrtf(fmaxf(0.000025f, __fdiv_rn(det_cov, det)));
      if (det != 0.0f) {
        const float det_inv = __frcp_rn(det);
        const float3 conic = make_float3(__fmul_rn(cov.z, det_inv), __fmul_rn(cov.y, -det_inv), __fmul_rn(cov.x, det_inv));
        const float mid = __fmul_rn(__fadd_rn(cov.x, cov.z), 0.5f);
        const float disc = __fsqrt_rn(fmaxf(0.1f, __fmaf_rn(mid, mid, -det)));
        const float lambda1 = __fadd_rn(mid, disc);
        const float lambda2 = __fsub_rn(mid, disc);
        const float my_radius = ceilf(__fmul_rn(__fsqrt_rn(fmaxf(lambda1, lambda2)), 3.f));
        const float pix_x = ndc_to_pix(ndc_x, W), pix_y = ndc_to_pix(ndc_y, H);
        const int rad = int(my_radius);
        // tile rectangle (auxiliary.h:45-55); int() truncates toward zero
        const int mnx = min(grid_x, max(0, int((pix_x - rad) / TILE_X)));
        const int mny = min(grid_y, max(0, int((pix_y - rad) / TILE_Y)));
        const int mxx = min(grid_x, max(0, int((pix_x + rad + TILE_X - 1) / TILE_X)));
        const int mxy = min(grid_y, max(0, int((pix_y + rad + TILE_Y - 1) / TILE_Y)));
        const uint32_t area = uint32_t(mxx - mnx) * uint32_t(mxy - mny);
        if (area != 0) {
          touched = area;
          radius_out = rad;
          g.depth[idx] = p_view.z;
          depth_key = __float_as_uint(p_view.z);
          // the record the blend kernels gather by Gaussian index: three 16-byte stores
          float4* rec = reinterpret_cast<float4*>(g.grec + idx);
          rec[0] = make_float4(conic.x, conic.y, conic.z, 1.f / p_view.z);
          rec[1] = make_float4(pix_x, pix_y, __ldg(opacities + idx) * h_scaling, colors ? __ldg(colors + idx) : 0.f);
          rec[2] = all_map ? __ldg(reinterpret_cast<const float4*>(all_map) + idx) : make_float4(0.f, 0.f, 0.f, 0.f);
          rect_out = make_uint2(uint32_t(mnx) | (uint32_t(mny) << 16), uint32_t(mxx) | (uint32_t(mxy) << 16));
        }
      }
    }
    radii[idx] = radius_out;
    g.tiles[idx] = touched;
    g.rect[idx] = rect_out;
    // input of the depth sort: key = bits of the view-space depth (culled Gaussians emit nothing, their key only
    // has to be deterministic), value = Gaussian index
    g.gs.keys[0][idx] = depth_key;
    g.gs.vals[0][idx] = uint32_t(idx);
  }
  // block sum of tiles_touched -> blk_sum
  uint32_t v = touched;
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) s_wsum[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t t = 0;
    for (int i = 0; i < 8; ++i) t += s_wsum[i];
    g.blk_sum[blockIdx.x] = t;
  }
}

// Single CTA: exclusive scan of the per-block sums, total -> g.total[0].
__global__ void __launch_bounds__(1024) scan_block_sums(int64_t nblk, GeomState g, int slot, uint32_t* __restrict__ host_out) {
  pdl_wait();
  __shared__ uint32_t s_w[32];
  __shared__ uint32_t s_carry;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  const uint32_t lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  // four consecutive sums per thread: the ~3900 block sums of a 1 M-Gaussian scene are one trip of the loop (this
  // kernel is on the path to the host's copy of R)
  for (int64_t base = 0; base < nblk; base += 4096) {
    const int64_t i0 = base + int64_t(threadIdx.x) * 4;
    uint32_t v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) v[k] = (i0 + k < nblk) ? g.blk_sum[i0 + k] : 0u;
    const uint32_t mine = v[0] + v[1] + v[2] + v[3];
    uint32_t inc = mine;
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t n = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += n;
    }
    if (lane == 31) s_w[w] = inc;
    __syncthreads();
    if (w == 0) {
      uint32_t x = s_w[lane], xi = x;
      for (int o = 1; o < 32; o <<= 1) {
        uint32_t n = __shfl_up_sync(0xffffffffu, xi, o);
        if (lane >= o) xi += n;
      }
      s_w[lane] = xi - x;  // exclusive warp base
    }
    __syncthreads();
    const uint32_t carry = s_carry;
    uint32_t run = carry + s_w[w] + inc - mine;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (i0 + k < nblk) g.blk_prefix[i0 + k] = run;
      run += v[k];
    }
    __syncthreads();
    if (threadIdx.x == 1023) s_carry = run;
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    g.total[slot] = s_carry;
    // the host's copy of R: stored straight into mapped pinned memory (no copy-engine operation between this kernel
    // and the event the host waits on; a D2H copy would also queue behind whatever bulk D2H traffic the application has
    // in flight on other streams)
    if (host_out) { *reinterpret_cast<volatile uint32_t*>(host_out) = s_carry; __threadfence_system(); }
  }
}

// Block sums of tiles_touched (shift = 0) or of the number of super-tiles overlapped (shift = ST_SHIFT), taken in
// depth order (perm = Gaussian indices sorted by depth).
__global__ void __launch_bounds__(256)
perm_block_sums(int64_t P, const uint32_t* __restrict__ perm, GeomState g, int shift) {
  pdl_wait();
  __shared__ uint32_t s_wsum[8];
  const int64_t i = int64_t(blockIdx.x) * 256 + threadIdx.x;
  uint32_t v = 0;
  if (i < P) {
    const uint32_t idx = perm[i];
    // (one gather behind perm[i], not two: a culled Gaussian has an empty rect, and a rect's area is its tile count)
    const uint2 rc = g.rect[idx];
    v = rect_area(shift ? st_rect(rc, shift) : rc);   // tiles, or super-tiles, overlapped
  }
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) s_wsum[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t t = 0;
    for (int i2 = 0; i2 < 8; ++i2) t += s_wsum[i2];
    g.blk_sum[blockIdx.x] = t;
  }
}

// One (tile, Gaussian) pair per overlapped tile, row-major over the rect like
// rasterizer_impl.cu:70-111, but Gaussians are visited in depth order and the key is the
// tile index alone (see BinScratch). Slot i of the scan is Gaussian perm[i].
// The reference walks each rect serially in one thread; here a CTA scans its 256 tile
// counts into shared memory and then every thread produces output slots tid, tid+256, ...
// (binary search of the slot in the scanned counts), so the writes are coalesced and the
// work is balanced no matter how uneven the rects are.
__global__ void __launch_bounds__(256)
emit_keys(int64_t P, const uint32_t* __restrict__ perm, GeomState g, int grid_x, uint32_t* __restrict__ keys,
          uint32_t* __restrict__ vals, uint32_t cap, uint32_t* __restrict__ nr_out, uint32_t* __restrict__ hist, int bpp,
          int shift) {
  pdl_wait();
  // digit histograms of the tile sort that follows (two passes of bpp bits, see radix_sort_begin): counted here,
  // while the keys are in registers anyway, instead of in a pass of their own over the R keys
  __shared__ uint32_t s_hist[2][256];
  s_hist[0][threadIdx.x] = 0;
  s_hist[1][threadIdx.x] = 0;
  const uint32_t dmask = (1u << bpp) - 1u;
  // capacity mode: report {R, R > capacity} to the caller's device counter; slots >= cap are dropped below
  // (the farthest instances, since emission is in depth order) and every later stage works on min(R, cap)
  if (nr_out && blockIdx.x == 0 && threadIdx.x == 0) {
    const uint32_t R = g.total[0];
    nr_out[0] = R;
    nr_out[1] = R > cap ? 1u : 0u;
  }
  __shared__ uint32_t s_w[8];
  __shared__ uint32_t s_off[257];
  __shared__ uint32_t s_idx[256];
  __shared__ uint2 s_rect[256];
  const int64_t i = int64_t(blockIdx.x) * 256 + threadIdx.x;
  const uint32_t lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const uint32_t idx = (i < P) ? perm[i] : 0u;
  uint2 my_rect = make_uint2(0u, 0u);
  if (i < P) {
    my_rect = g.rect[idx];                                   // (empty for a culled Gaussian)
    if (shift) my_rect = st_rect(my_rect, shift);            // (grid_x is then in super-tiles)
  }
  const uint32_t cnt = rect_area(my_rect);
  uint32_t inc = cnt;
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t n = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += n;
  }
  if (lane == 31) s_w[w] = inc;
  s_idx[threadIdx.x] = idx;
  if (cnt) s_rect[threadIdx.x] = my_rect;
  __syncthreads();
  uint32_t wb = 0;
  for (uint32_t k = 0; k < w; ++k) wb += s_w[k];
  s_off[threadIdx.x] = wb + inc - cnt;
  if (threadIdx.x == 255) s_off[256] = wb + inc;
  __syncthreads();
  const uint32_t total = s_off[256];
  const uint32_t base = g.blk_prefix[blockIdx.x];
  for (uint32_t o = threadIdx.x; o < total; o += 256) {
    // last t with s_off[t] <= o; among equal offsets that is the one with a non-zero count
    uint32_t lo = 0, hi = 256;
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const uint32_t mid = (lo + hi) >> 1;
      if (s_off[mid] <= o) lo = mid; else hi = mid;
    }
    const uint32_t local = o - s_off[lo];
    const uint2 rc = s_rect[lo];
    const uint32_t mnx = rc.x & 0xffffu, mny = rc.x >> 16, mxx = rc.y & 0xffffu;
    const uint32_t wdt = mxx - mnx;
    const uint32_t ry = local / wdt, rx = local - ry * wdt;
    if (base + o < cap) {
      const uint32_t key = (mny + ry) * uint32_t(grid_x) + (mnx + rx);
      keys[base + o] = key;
      vals[base + o] = s_idx[lo];
      atomicAdd(&s_hist[0][key & dmask], 1u);
      atomicAdd(&s_hist[1][(key >> bpp) & dmask], 1u);
    }
  }
  __syncthreads();
  if (hist && threadIdx.x <= dmask) {
    if (s_hist[0][threadIdx.x]) atomicAdd(&hist[threadIdx.x], s_hist[0][threadIdx.x]);
    if (s_hist[1][threadIdx.x]) atomicAdd(&hist[256 + threadIdx.x], s_hist[1][threadIdx.x]);
  }
}

// Per-tile [start, end) into the sorted list (rasterizer_impl.cu:116-138); ranges is zero-filled.
// Four keys per thread (one 128-bit load) and a grid-stride loop over a grid of a few CTAs per SM: with one key per
// thread the kernel is bound by the launch rate of its 34 000 tiny CTAs (38 us at C4), not by its 35 MB.
__global__ void __launch_bounds__(256)
tile_ranges(int64_t R, const uint32_t* __restrict__ d_n, const uint32_t* __restrict__ sorted_tiles,
            uint2* __restrict__ ranges) {
  pdl_wait();
  if (d_n) R = min(R, int64_t(*d_n));   // capacity mode: the count lives on the device
  const int64_t nquad = (R + 3) >> 2;
  for (int64_t q = int64_t(blockIdx.x) * 256 + threadIdx.x; q < nquad; q += int64_t(gridDim.x) * 256) {
    const int64_t i0 = q << 2;
    uint32_t k[4];
    if (i0 + 3 < R) {
      const uint4 v = *reinterpret_cast<const uint4*>(sorted_tiles + i0);   // (the key buffers are 128-byte aligned)
      k[0] = v.x; k[1] = v.y; k[2] = v.z; k[3] = v.w;
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) k[j] = (i0 + j < R) ? sorted_tiles[i0 + j] : 0u;
    }
    uint32_t prev = i0 > 0 ? sorted_tiles[i0 - 1] : 0u;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int64_t i = i0 + j;
      if (i < R) {
        const uint32_t cur = k[j];
        if (i == 0) ranges[cur].x = 0;
        else if (cur != prev) { ranges[prev].y = uint32_t(i); ranges[cur].x = uint32_t(i); }
        if (i == R - 1) ranges[cur].y = uint32_t(R);
        prev = cur;
      }
    }
  }
}

// ---------------------------------------------------------------------------
constexpr int BATCH = BLEND_THREADS;

// Launch order of the blend CTAs: tiles by descending list length (counting sort on the length quantised to
// 1024 buckets, one CTA). CTAs are dispatched in index order, so the long tiles start first and the kernel's
// tail is made of short ones instead of whichever tiles happen to sit in the last image rows.
// ntiles counts blend CTAs (tile * BLEND_SUBS + sub). Lengths: the tile's range (end - start) when maxc is
// NULL, else maxc (what the backward walks).
// (lens != NULL: the lengths themselves, one per tile)
__device__ __forceinline__ void order_tiles_body(uint32_t ntiles, const uint2* __restrict__ ranges, const uint32_t* __restrict__ maxc,
                                                 const uint32_t* __restrict__ lens, uint32_t* __restrict__ tile_order) {
  __shared__ uint32_t s_cnt[1024];
  __shared__ uint32_t s_w[32];
  __shared__ uint32_t s_max;
  const uint32_t tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  s_cnt[tid] = 0;
  if (tid == 0) s_max = 1;
  __syncthreads();
  auto length = [&](uint32_t t) { if (lens) return lens[t / BLEND_SUBS]; if (maxc) return maxc[t]; const uint2 r = ranges[t / BLEND_SUBS]; return r.y - r.x; };
  // every pass reads the lengths in batches of 8 independent loads per thread (one CTA: latency, not bandwidth)
  uint32_t mx = 0;
  for (uint32_t base = 0; base < ntiles; base += 8192) {
    uint32_t len[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { const uint32_t t = base + k * 1024 + tid; len[k] = t < ntiles ? length(t) : 0u; }
#pragma unroll
    for (int k = 0; k < 8; ++k) mx = max(mx, len[k]);
  }
  mx = __reduce_max_sync(0xffffffffu, mx);
  if (lane == 0) atomicMax(&s_max, mx);
  __syncthreads();
  const uint32_t top = s_max;
  // bucket 0 = longest lists
  auto bucket = [&](uint32_t len) { return 1023u - uint32_t((uint64_t(len) * 1023u) / top); };
  for (uint32_t base = 0; base < ntiles; base += 8192) {
    uint32_t len[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { const uint32_t t = base + k * 1024 + tid; len[k] = t < ntiles ? length(t) : 0u; }
#pragma unroll
    for (int k = 0; k < 8; ++k) if (base + k * 1024 + tid < ntiles) atomicAdd(&s_cnt[bucket(len[k])], 1u);
  }
  __syncthreads();
  // exclusive scan of the 1024 bucket counts
  const uint32_t v = s_cnt[tid];
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
  s_cnt[tid] = s_w[w] + inc - v;   // first output slot of the bucket
  __syncthreads();
  for (uint32_t base = 0; base < ntiles; base += 8192) {
    uint32_t len[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { const uint32_t t = base + k * 1024 + tid; len[k] = t < ntiles ? length(t) : 0u; }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const uint32_t t = base + k * 1024 + tid;
      if (t < ntiles) tile_order[atomicAdd(&s_cnt[bucket(len[k])], 1u)] = t;   // order inside a bucket is irrelevant
    }
  }
}
__global__ void __launch_bounds__(1024)
order_tiles(uint32_t ntiles, const uint2* __restrict__ ranges, const uint32_t* __restrict__ maxc,
            uint32_t* __restrict__ tile_order) {
  pdl_wait();
  order_tiles_body(ntiles, ranges, maxc, nullptr, tile_order);
}
// Super-tile binning: the single CTA that scans the tile totals into the tile ranges also has the list lengths at
// hand, so it orders the blend CTAs too (one launch and one pass over the ranges less).
__global__ void __launch_bounds__(1024)
bin_scan_tiles_order(uint32_t tiles, const uint32_t* __restrict__ tile_cnt, uint32_t cap, uint2* __restrict__ ranges,
                     uint32_t* __restrict__ tile_order) {
  pdl_wait();
  bin_scan_tiles_body(tiles, tile_cnt, cap, ranges);
  __syncthreads();
  order_tiles_body(tiles * BLEND_SUBS, nullptr, nullptr, tile_cnt, tile_order);
}

// One CTA per 16x16 tile, one thread per pixel, one warp per 8x4 pixel block.
// Staging, three batches deep: thread 0 streams the tile's sorted Gaussian list (the ids: one contiguous span of
// point_list) into a shared ring with cp.async.bulk (TMA 1-D bulk copy + mbarrier) two batches ahead; one batch
// ahead every thread gathers ITS record of the batch from the per-Gaussian table (GeomState::grec, L2-resident)
// with 16-byte cp.async copies straight into shared memory; the current batch is blended. Each warp first
// tests 32 records at a time (one per lane, block_candidate in common.cuh) against its pixel block and only
// walks the instances that can reach it, in list order (forward.cu:331-396 semantics).
#ifndef CG_FWD_CTAS
#define CG_FWD_CTAS 6
#endif
__device__ __forceinline__ void fwd_cp_async16(void* smem, const void* g) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem)), "l"(g) : "memory");
}
template <bool GEO>
__global__ void __launch_bounds__(BLEND_THREADS, CG_FWD_CTAS * (256 / BLEND_THREADS))
blend_fwd(const uint2* __restrict__ ranges, const uint32_t* __restrict__ tile_order, int grid_x,
          const Rec* __restrict__ grec, const uint32_t* __restrict__ point_list, int W, int H,
          const float* __restrict__ bg, float* __restrict__ out_color, float* __restrict__ out_invd,
          float* __restrict__ out_map, float* __restrict__ final_T, uint32_t* __restrict__ n_contrib,
          uint32_t* __restrict__ tile_maxc, uint32_t* __restrict__ cand_lists, uint32_t* __restrict__ cand_ids,
          uint32_t* __restrict__ blk_cnt, uint32_t* __restrict__ cls_count, uint4* __restrict__ cls_list,
          uint32_t nblocks) {
  pdl_wait();
  __shared__ __align__(128) Rec s_rec[2][BATCH];
  __shared__ __align__(128) uint32_t s_ids[3][BATCH + 8];   // (+ the alignment slack of the bulk copy's source)
  __shared__ __align__(8) uint64_t s_full[3];
  __shared__ uint32_t s_maxc;

  const uint32_t cta = tile_order[blockIdx.x];
  const uint32_t tile = cta / BLEND_SUBS, sub = cta % BLEND_SUBS;
  const uint32_t tile_x = tile % uint32_t(grid_x), tile_y = tile / uint32_t(grid_x);
  const uint32_t tid = threadIdx.x;
  const uint32_t lane = tid & 31, warp = tid >> 5;
  const uint32_t blk_x = tile_x * TILE_X + (warp & 1) * 8;
  const uint32_t blk_y = tile_y * TILE_Y + sub * BLEND_ROWS + (warp >> 1) * 4;
  // the pixel is kept as two floats only (the integer coordinates are recovered at the end): at the kernel's
  // 40-register cap the compiler otherwise keeps the integers and converts them again for every walked pair
  float pxf, pyf, T;
  {
    const uint32_t pix_x = blk_x + (lane & 7), pix_y = blk_y + (lane >> 3);
    T = pix_x < uint32_t(W) && pix_y < uint32_t(H) ? 1.0f : -1.0f;   // (see `done` below)
    pxf = float(pix_x); pyf = float(pix_y);
    asm volatile("" : "+f"(pxf), "+f"(pyf));
  }
  // pixel block of this warp, clipped to the image
  const float bx0 = float(blk_x), bx1 = float(min(blk_x + 7u, uint32_t(W) - 1u));
  const float by0 = float(blk_y), by1 = float(min(blk_y + 3u, uint32_t(H) - 1u));
  const uint2 range = ranges[tile];
  const int total = int(range.y - range.x);
  const int rounds = (total + BATCH - 1) / BATCH;

  if (tid == 0) {
    mbar_init(&s_full[0], 1);
    mbar_init(&s_full[1], 1);
    mbar_init(&s_full[2], 1);
    mbar_fence_init();
    s_maxc = 0;
  }
  __syncthreads();
  // ids of batch k: TMA bulk copy of point_list[range.x + k * BATCH, +n) into s_ids[k % 3]. A bulk copy moves
  // 16-byte granules from a 16-byte aligned source, so it starts at the aligned address below the span (ids_off
  // entries early, the same for every batch) and is rounded up at the end; point_list is padded, the surplus
  // entries are never used.
  const uint32_t ids_off = range.x & 3u;
  auto issue_ids = [&](int k) {
    const uint32_t nb = uint32_t(min(BATCH, total - k * BATCH));
    const uint32_t bytes = (((ids_off + nb) * 4u) + 15u) & ~15u;
    const uint32_t* src = point_list + (size_t(range.x) - ids_off) + size_t(k) * BATCH;
    mbar_expect_tx(&s_full[k % 3], bytes);
    bulk_g2s(&s_ids[k % 3][0], src, bytes, &s_full[k % 3]);
  };
  // (ring slot 0 is first used by batch 3: batch 0 takes its ids with plain loads, see below)
  auto ids_parity = [](int k) { return uint32_t(k / 3 + (k % 3 == 0 ? 1 : 0)) & 1u; };
  // records of batch k: every thread gathers the record of its slot
  auto issue_records = [&](int k) {
    mbar_wait(&s_full[k % 3], ids_parity(k));
    const int nb = min(BATCH, total - k * BATCH);
    if (int(tid) < nb) {
      const Rec* r = grec + s_ids[k % 3][ids_off + tid];
      Rec* d = &s_rec[k & 1][tid];
      fwd_cp_async16(&d->ca, &r->ca);
      fwd_cp_async16(&d->x, &r->x);
      if (GEO) fwd_cp_async16(&d->m0, &r->m0);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  if (rounds > 0) {
    if (tid == 0) { if (rounds > 1) issue_ids(1); if (rounds > 2) issue_ids(2); }
    // batch 0 takes its ids with plain loads: one L2 round trip less before the first record can be gathered, which
    // is most of the life of the many short tiles
    const int nb = min(BATCH, total);
    if (int(tid) < nb) {
      const uint32_t id = point_list[size_t(range.x) + tid];
      s_ids[0][ids_off + tid] = id;
      const Rec* r = grec + id;
      Rec* d = &s_rec[0][tid];
      fwd_cp_async16(&d->ca, &r->ca);
      fwd_cp_async16(&d->x, &r->x);
      if (GEO) fwd_cp_async16(&d->m0, &r->m0);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  }

  // "done" is the sign of T (T itself stays > 0: a pixel stops BEFORE T would drop under 1e-4): one register and
  // three instructions per walked pair less than a separate flag in a kernel that runs at its 40-register cap
  float C = 0.f, invd_acc = 0.f;
#define done (T < 0.f)
  float M0 = 0.f, M1 = 0.f, M2 = 0.f, M3v = 0.f;
  uint32_t last_contributor = 0;
  // contributor lists of the two 4x4 halves of this warp's 8x4 block (BinKeep::cand / cand_id): tile-relative
  // positions and Gaussian ids of the instances blended into at least one pixel of the half, in list order
  // (32-bit offsets into the lists: the kernel runs at 40 registers)
  const uint32_t cand_start = 16u * range.x + (sub * BLEND_WARPS + warp) * 2u * uint32_t(total);
  uint32_t cand_l = cand_start, cand_r = cand_start + uint32_t(total);
  const bool left = (lane & 4u) == 0u;

  int todo = total;
  for (int b = 0; b < rounds; ++b, todo -= BATCH) {
    asm volatile("cp.async.wait_group 0;" ::: "memory");     // this thread's record of batch b has landed
    const int num_done = __syncthreads_count(done);          // ... and everybody's is visible
    if (num_done == BLEND_THREADS) {
      // id spans still in flight (batch b+1; in the first iteration also batch 2): drain them before the CTA retires
      if (tid == 0 && b + 1 < rounds) mbar_wait(&s_full[(b + 1) % 3], ids_parity(b + 1));
      if (tid == 0 && b == 0 && rounds > 2) mbar_wait(&s_full[2], ids_parity(2));
      break;
    }
    if (b + 1 < rounds) issue_records(b + 1);
    if (tid == 0 && b > 0 && b + 2 < rounds) issue_ids(b + 2);   // (batches 1 and 2 were requested in the prologue)
    const Rec* batch = s_rec[b & 1];
    const uint32_t* ids = s_ids[b % 3] + ids_off;
    const int n = min(BATCH, todo);
    const uint32_t pos0 = uint32_t(b) * BATCH;
    for (int r = 0; r < n; r += 32) {
      if (__all_sync(0xffffffffu, done)) break;
      const int idx = r + int(lane);
      bool cand = false;
      if (idx < n) {
        const float4 q0 = *reinterpret_cast<const float4*>(&batch[idx].ca);   // ca cb cc invd
        const float4 q1 = *reinterpret_cast<const float4*>(&batch[idx].x);    // x y o col
        cand = block_candidate(q1.x, q1.y, q0.x, q0.y, q0.z, q1.z, bx0, bx1, by0, by1);
      }
      uint32_t mask = __ballot_sync(0xffffffffu, cand);
      uint32_t mine = 0;   // bit k: this pixel blended the chunk's record k
      while (mask) {
        const uint32_t rest = mask & (mask - 1);
        const uint32_t bit = mask ^ rest;          // lowest candidate
        const int j = r + (31 - __clz(int(bit)));
        mask = rest;
        // straight-line form: every lane evaluates the pair and the updates are predicated (five issue slots per pair
        // fewer than skipping with branches, which only pays when all 32 pixels fail the same test)
        const float4 q0 = *reinterpret_cast<const float4*>(&batch[j].ca);   // ca cb cc invd
        const float4 q1 = *reinterpret_cast<const float4*>(&batch[j].x);    // x y o col
        const float dx = __fsub_rn(q1.x, pxf), dy = __fsub_rn(q1.y, pyf);
        const float power = gauss_power(q0.x, q0.y, q0.z, dx, dy);
        const float alpha = fminf(0.99f, __fmul_rn(q1.z, expf(power)));
        const float test_T = __fmul_rn(T, __fsub_rn(1.f, alpha));
        const bool act = !done && !(power > 0.0f) && !(alpha < 1.0f / 255.0f);
        const bool stop = test_T < 0.0001f;
        if (act && !stop) {
          C = __fmaf_rn(__fmul_rn(q1.w, alpha), T, C);
          invd_acc = __fmaf_rn(__fmul_rn(q0.w, alpha), T, invd_acc);
          if (GEO) {
            const float4 mp = *reinterpret_cast<const float4*>(&batch[j].m0);
            M0 = __fmaf_rn(__fmul_rn(mp.x, alpha), T, M0);
            M1 = __fmaf_rn(__fmul_rn(mp.y, alpha), T, M1);
            M2 = __fmaf_rn(__fmul_rn(mp.z, alpha), T, M2);
            M3v = __fmaf_rn(__fmul_rn(mp.w, alpha), T, M3v);
          }
          mine |= bit;
        }
        if (act) T = stop ? -T : test_T;
      }
      if (mine) last_contributor = pos0 + uint32_t(r) + 32u - uint32_t(__clz(int(mine)));   // (1-based position of the last blended)
      // records some pixel of the left / right half blended: lane k files record k of the chunk
      const uint32_t hit_l = __reduce_or_sync(0xffffffffu, left ? mine : 0u);
      const uint32_t hit_r = __reduce_or_sync(0xffffffffu, left ? 0u : mine);
      if ((hit_l | hit_r) >> lane & 1u) {
        const uint32_t lt = (1u << lane) - 1u, pos = pos0 + uint32_t(idx), id = ids[idx];
        if (hit_l >> lane & 1u) { const uint32_t at = cand_l + __popc(hit_l & lt); cand_lists[at] = pos; cand_ids[at] = id; }
        if (hit_r >> lane & 1u) { const uint32_t at = cand_r + __popc(hit_r & lt); cand_lists[at] = pos; cand_ids[at] = id; }
      }
      cand_l += __popc(hit_l);
      cand_r += __popc(hit_r);
    }
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");   // (a gather issued for a batch the CTA did not reach)

#undef done
  T = fabsf(T);
  const uint32_t pix_x = uint32_t(pxf), pix_y = uint32_t(pyf), pix_id = uint32_t(W) * pix_y + pix_x;
  const bool inside = pix_x < uint32_t(W) && pix_y < uint32_t(H);
  if (inside) {
    final_T[pix_id] = T;
    n_contrib[pix_id] = last_contributor;
    out_color[pix_id] = __fmaf_rn(T, bg[0], C);
    out_invd[pix_id] = invd_acc;
    if (GEO) {
      const size_t hw = size_t(H) * W;
      out_map[pix_id] = M0;
      out_map[hw + pix_id] = M1;
      out_map[2 * hw + pix_id] = M2;
      out_map[3 * hw + pix_id] = M3v;
    }
  }
  uint32_t mc = inside ? last_contributor : 0u;
  mc = __reduce_max_sync(0xffffffffu, mc);
  if (lane == 0) atomicMax(&s_maxc, mc);
  if (lane < 2) {
    // each half joins the list of its size class (half octaves of the contributor count): the ring backward takes
    // the classes largest first, which is all the ordering its greedy scheduling needs
    const uint32_t n = lane ? cand_r - cand_start - uint32_t(total) : cand_l - cand_start;
    const uint32_t bid = (tile * 8 + sub * BLEND_WARPS + warp) * 2u + lane;
    blk_cnt[bid] = n;
    if (n) {
      const uint32_t msb = 31u - uint32_t(__clz(int(n)));
      const uint32_t cls = min(uint32_t(RING_CLASSES - 1), 2u * msb + (msb ? ((n >> (msb - 1u)) & 1u) : 0u));
      cls_list[size_t(cls) * nblocks + atomicAdd(&cls_count[cls], 1u)] = make_uint4(bid, n, range.x, uint32_t(total));
    }
  }
  __syncthreads();
  if (tid == 0) tile_maxc[cta] = s_maxc;
}

// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
mark_visible_kernel(int64_t P, const float* __restrict__ means3D, const float* __restrict__ vm, uint8_t* __restrict__ present) {
  pdl_wait();
  const int64_t i = int64_t(blockIdx.x) * 256 + threadIdx.x;
  if (i >= P) return;
  const float3 v = xform43(means3D[3 * i], means3D[3 * i + 1], means3D[3 * i + 2], vm);
  present[i] = (v.z <= 0.2f) ? 0 : 1;
}

// ---------------------------------------------------------------------------
// host side
// Binning mode: super-tile counting (binning.cuh) unless CURVEGS_BINNING=sort or the image has more super-tiles
// than the chunk tables are sized for.
static bool bin_by_supertile(int W, int H) {
  static const bool want = [] { const char* e = getenv("CURVEGS_BINNING"); return !(e && strcmp(e, "sort") == 0); }();
  const size_t nst = size_t((W + TILE_X * ST_SIDE - 1) / (TILE_X * ST_SIDE)) * ((H + TILE_Y * ST_SIDE - 1) / (TILE_Y * ST_SIDE));
  return want && nst <= size_t(BIN_MAX_SUPERTILES);
}
int launch_fwd_geom(const cg_raster_settings* s, int64_t P, const float* means3D, const float* opacities,
                    const float* scales, const float* rotations, const float* cov3D_precomp, const float* colors,
                    const float* all_map, int32_t* radii, void* geom, int64_t* num_rendered, cudaStream_t st) {
  GeomState g = GeomState::carve(geom, P, nullptr);
  const int W = s->image_width, H = s->image_height;
  const int gx = (W + TILE_X - 1) / TILE_X, gy = (H + TILE_Y - 1) / TILE_Y;
  const float fy = H / (2.0f * s->tanfovy), fx = W / (2.0f * s->tanfovx);
  const int64_t nblk = (P + 255) / 256;
  { StageTimer t_(ST_PREPROCESS_FWD, st, 1);
  launch_k(preprocess_fwd, dim3(unsigned(nblk)), dim3(256), 0, st, P, means3D, opacities, scales, rotations, cov3D_precomp,
                                                 s->scale_modifier, s->viewmatrix, s->projmatrix, W, H, s->tanfovx,
                                                 s->tanfovy, fx, fy, gx, gy, s->antialiasing, colors,
                                                 s->render_geo ? all_map : nullptr, radii, g); }
  CG_LAUNCH_CHECK(s->debug, st);
  // R goes to pinned host memory (written by the scan kernel itself); the host then waits on an EVENT recorded right
  // behind that kernel while the stream already carries the next, R-independent stage (depth sort of the Gaussians +
  // offsets in depth order), so the wake-up latency of the host is hidden behind ~0.1 ms of useful GPU work.
  // (num_rendered == NULL: capacity mode, R stays on the device and nothing here touches the host)
  int dev = 0;
  static thread_local uint32_t* h_total[64] = {nullptr};
  static thread_local uint32_t* h_total_dev[64] = {nullptr};   // device-side address of the same mapped allocation
  static thread_local cudaEvent_t h_event[64] = {nullptr};
  if (num_rendered) {
    CG_CUDA(cudaGetDevice(&dev));
    CG_ARG(dev >= 0 && dev < 64, "device ordinal");
  }
  if (num_rendered && !h_total[dev]) {
    CG_CUDA(cudaHostAlloc(reinterpret_cast<void**>(&h_total[dev]), 64, cudaHostAllocMapped));
    CG_CUDA(cudaHostGetDevicePointer(reinterpret_cast<void**>(&h_total_dev[dev]), h_total[dev], 0));
    CG_CUDA(cudaEventCreateWithFlags(&h_event[dev], cudaEventDisableTiming));
  }
  { StageTimer t_(ST_SCAN, st, 1);
  launch_k(scan_block_sums, dim3(1), dim3(1024), 0, st, nblk, g, 0, num_rendered ? h_total_dev[dev] : static_cast<uint32_t*>(nullptr)); }
  CG_LAUNCH_CHECK(s->debug, st);
  if (num_rendered) CG_CUDA(cudaEventRecord(h_event[dev], st));
  {
    int rc, gcur = 0;
    { StageTimer t_(ST_SORT, st, 0);
    rc = radix_sort_pairs<uint32_t>(g.gs, P, 32, &gcur, s->debug != 0, st); }
    if (rc != CG_OK) return rc;
    { StageTimer t_(ST_SCAN, st, 2);
    // offsets, in depth order, of what the binning emits per Gaussian: its tile instances (sort path) or its
    // super-tile copies (total -> g.total[1])
    const bool by_st = bin_by_supertile(W, H);
    launch_k(perm_block_sums, dim3(unsigned(nblk)), dim3(256), 0, st, P, g.gs.vals[gcur], g, by_st ? ST_SHIFT : 0);
    launch_k(scan_block_sums, dim3(1), dim3(1024), 0, st, nblk, g, by_st ? 1 : 0, static_cast<uint32_t*>(nullptr)); }
    CG_LAUNCH_CHECK(s->debug, st);
  }
  if (num_rendered) {
    CG_CUDA(cudaEventSynchronize(h_event[dev]));
    *num_rendered = int64_t(*h_total[dev]);
  }
  return CG_OK;
}

// nr_out == NULL: R is the exact instance count (known on the host). nr_out != NULL (capacity mode): R is the
// capacity of bin_keep / bin_scratch, the true count is read on the device (g.total) by every R-sized kernel and
// {count, overflow} is written to nr_out.
int launch_fwd_blend(const cg_raster_settings* s, int64_t P, int64_t R, void* geom, void* img, void* bin_keep, void* bin_scratch, float* out_color, float* out_invd,
                     float* out_map, uint32_t* nr_out, cudaStream_t st) {
  const uint32_t* d_n = nr_out ? GeomState::carve(geom, P, nullptr).total : nullptr;
  GeomState g = GeomState::carve(geom, P, nullptr);
  const int W = s->image_width, H = s->image_height;
  ImgState im = ImgState::carve(img, W, H, nullptr);
  BinKeep bk = BinKeep::carve(bin_keep, R, nullptr);
  BinScratch bs = BinScratch::carve(bin_scratch, P, R, nullptr);
  const int gx = (W + TILE_X - 1) / TILE_X, gy = (H + TILE_Y - 1) / TILE_Y;
  const int64_t nblk = (P + 255) / 256;
  const size_t tiles = size_t(gx) * gy;

  // tile ranges and, behind them, the size-class counters of the ring backward and the binning's tile counts and
  // super-tile ranges
  CG_CUDA(cudaMemsetAsync(im.ranges, 0, size_t(reinterpret_cast<char*>(im.chunk_start) - reinterpret_cast<char*>(im.ranges)), st));
  const uint32_t* perm = g.gs.vals[radix_sort_result_buf(32)];   // the Gaussians in depth order (cg_raster_fwd_geom)
  const bool by_st = R > 0 && bin_by_supertile(W, H);
  if (by_st) {
    // ---- super-tile binning (binning.cuh) ----
    int rc, cur = 0;
    const int sgx = (gx + ST_SIDE - 1) / ST_SIDE, sgy = (gy + ST_SIDE - 1) / ST_SIDE;
    const uint32_t ns = uint32_t(sgx) * uint32_t(sgy);
    const int end_bit = int(tile_key_bits(ns));
    int passes, bpp;
    radix_sort_geometry(end_bit, 32, &passes, &bpp);
    const bool fused_hist = passes <= 2;
    const uint32_t* d_copies = g.total + 1;   // number of copies: only known on the device
    if (fused_hist) { rc = radix_sort_begin<uint32_t>(bs.is, R, end_bit, st); if (rc != CG_OK) return rc; }
    { StageTimer t_(ST_EMIT_KEYS, st, 1);
    launch_k(emit_keys, dim3(unsigned(nblk)), dim3(256), 0, st, P, perm, g, sgx, bs.is.keys[0], bs.is.vals[0],
             uint32_t(R), nr_out, fused_hist ? bs.is.hist : nullptr, bpp, ST_SHIFT); }
    CG_LAUNCH_CHECK(s->debug, st);
    { StageTimer t_(ST_SORT, st, 0);
    rc = radix_sort_pairs<uint32_t>(bs.is, R, end_bit, &cur, s->debug != 0, st, d_copies, fused_hist, nullptr); }
    if (rc != CG_OK) return rc;
    const unsigned rb = unsigned(min(int64_t(148 * 16), (R + 1023) / 1024));
    const unsigned cb = unsigned(BinScratch::max_chunks(R) - BIN_MAX_SUPERTILES + ns);   // chunks: at most R / BIN_CHUNK + ns + 1
    const bool spans_from_hist = fused_hist && passes == 1;   // (up to 256 super-tiles: 2048 x 2048 pixels)
    { StageTimer t_(ST_TILE_RANGES, st, spans_from_hist ? 4 : 5);
    if (!spans_from_hist) launch_k(tile_ranges, dim3(rb), dim3(256), 0, st, R, d_copies, bs.is.keys[cur], im.st_ranges);
    launch_k(bin_chunk_table, dim3(1), dim3(1024), 0, st, ns, im.st_ranges, im.chunk_start,
             spans_from_hist ? bs.is.hist : nullptr, d_copies, uint32_t(R));
    launch_k(bin_count, dim3(min(cb, 148u * 8u)), dim3(256), 0, st, ns, uint32_t(sgx), im.st_ranges, im.chunk_start, bs.is.vals[cur], g.rect, bs.ccnt);
    launch_k(bin_scan_chunks, dim3(ns), dim3(64), 0, st, uint32_t(sgx), gx, gy, im.chunk_start, bs.ccnt, bs.cbase, im.tile_cnt);
    launch_k(bin_scan_tiles_order, dim3(1), dim3(1024), 0, st, uint32_t(tiles), im.tile_cnt, uint32_t(R), im.ranges, im.tile_order);
    launch_k(bin_fill, dim3(min(cb, 148u * 8u)), dim3(256), 0, st, ns, uint32_t(sgx), gx, gy, im.st_ranges, im.chunk_start, bs.is.vals[cur], g.rect,
             bs.cbase, im.ranges, uint32_t(R), bk.point_list); }
    CG_LAUNCH_CHECK(s->debug, st);
  } else if (R > 0) {
    int rc;
    // ---- sort path: one (tile, Gaussian) pair per overlapped tile, emitted in depth order, ...
    // stable sort by tile only; emit_keys counts the digit histograms of its passes while it writes the keys
    int cur = 0;
    const int end_bit = int(tile_key_bits(uint32_t(tiles)));
    int passes, bpp;
    radix_sort_geometry(end_bit, 32, &passes, &bpp);
    const bool fused_hist = passes <= 2;   // (up to 16 tile bits, i.e. any image below 4096 x 4096 pixels)
    if (fused_hist) { rc = radix_sort_begin<uint32_t>(bs.is, R, end_bit, st); if (rc != CG_OK) return rc; }
    { StageTimer t_(ST_EMIT_KEYS, st, 1);
    launch_k(emit_keys, dim3(unsigned(nblk)), dim3(256), 0, st, P, perm, g, gx, bs.is.keys[0], bs.is.vals[0],
             uint32_t(R), nr_out, fused_hist ? bs.is.hist : nullptr, bpp, 0); }
    CG_LAUNCH_CHECK(s->debug, st);
    { StageTimer t_(ST_SORT, st, 0);
    // (its last pass leaves the sorted Gaussian indices directly in the point list that the backward keeps)
    rc = radix_sort_pairs<uint32_t>(bs.is, R, end_bit, &cur, s->debug != 0, st, d_n, fused_hist, bk.point_list); }
    if (rc != CG_OK) return rc;
    const unsigned rb = unsigned(min(int64_t(148 * 16), (R + 1023) / 1024));
    { StageTimer t_(ST_TILE_RANGES, st, 1);
    launch_k(tile_ranges, dim3(rb), dim3(256), 0, st, R, d_n, bs.is.keys[cur], im.ranges); }
    CG_LAUNCH_CHECK(s->debug, st);
  }
  const dim3 grid{unsigned(tiles) * BLEND_SUBS, 1u, 1u}, block{unsigned(BLEND_THREADS), 1u, 1u};
  StageTimer t_blend(ST_BLEND_FWD, st, by_st ? 1 : 2);
  if (!by_st) launch_k(order_tiles, dim3(1), dim3(1024), 0, st, uint32_t(tiles) * BLEND_SUBS, im.ranges, nullptr, im.tile_order);
  CG_LAUNCH_CHECK(s->debug, st);
  if (s->render_geo)
    launch_k(blend_fwd<true>, dim3(grid), dim3(block), 0, st, im.ranges, im.tile_order, gx, g.grec, bk.point_list, W, H, s->bg,
             out_color, out_invd, out_map, im.final_T, im.n_contrib, im.tile_maxc, bk.cand, bk.cand_id, im.blk_cnt,
             im.cls_count, im.cls_list, uint32_t(tiles) * 16u);
  else
    launch_k(blend_fwd<false>, dim3(grid), dim3(block), 0, st, im.ranges, im.tile_order, gx, g.grec, bk.point_list, W, H, s->bg,
             out_color, out_invd, out_map, im.final_T, im.n_contrib, im.tile_maxc, bk.cand, bk.cand_id, im.blk_cnt,
             im.cls_count, im.cls_list, uint32_t(tiles) * 16u);
  CG_LAUNCH_CHECK(s->debug, st);
  return CG_OK;
}

int launch_rebuild_keys(int64_t P, int64_t R, int W, int H, const void* geom, const void* img, const void* bin_keep,
                        const void* bin_scratch, uint64_t* dst, cudaStream_t st) {
  if (R <= 0) return CG_OK;
  GeomState g = GeomState::carve(const_cast<void*>(geom), P, nullptr);
  ImgState im = ImgState::carve(const_cast<void*>(img), W, H, nullptr);
  BinKeep bk = BinKeep::carve(const_cast<void*>(bin_keep), R, nullptr);
  (void)bin_scratch;   // (the tile of an instance follows from the tile ranges; the sorted tile keys are not needed)
  const size_t tiles = size_t((W + TILE_X - 1) / TILE_X) * ((H + TILE_Y - 1) / TILE_Y);
  count_launches(1);
  launch_k(rebuild_keys_ranges, dim3(unsigned(tiles)), dim3(256), 0, st, im.ranges, bk.point_list, g.depth, dst);
  CG_LAUNCH_CHECK(0, st);
  return CG_OK;
}

int launch_mark_visible(int64_t P, const float* means3D, const float* vm, uint8_t* present, cudaStream_t st) {
  if (P == 0) return CG_OK;
  count_launches(1);
  launch_k(mark_visible_kernel, dim3(unsigned((P + 255) / 256)), dim3(256), 0, st, P, means3D, vm, present);
  CG_LAUNCH_CHECK(0, st);
  return CG_OK;
}

}  // namespace cg
