// Shared helpers for libcurvegs (sm_100a). No torch headers anywhere in csrc/.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
#include <stdio.h>
#include "../../include/curvegs.h"

namespace cg {

constexpr int TILE_X = 16;   // reference config.h:17-18 (BLOCK_X/BLOCK_Y)
constexpr int TILE_Y = 16;
constexpr int TILE_PIX = TILE_X * TILE_Y;
// Blend CTAs: a warp owns an 8x4 pixel block; BLEND_WARPS warps form one CTA, i.e. BLEND_SUBS CTAs share a
// 16x16 tile (and stream the same record list). Measured at C4: 4 warps (half tiles) make the backward 1.5 %
// faster and the forward 2 % slower than 8, so whole tiles stay.
constexpr int BLEND_WARPS = 8;
constexpr int BLEND_THREADS = BLEND_WARPS * 32;
constexpr int BLEND_SUBS = 8 / BLEND_WARPS;
constexpr int BLEND_ROWS = (BLEND_WARPS / 2) * 4;   // pixel rows per CTA
constexpr int RING_CLASSES = 32;   // size classes of the 8x4 blocks' candidate lists (see ImgState::cls_count)
static_assert(BLEND_SUBS == 1, "the block candidate lists (blend_ring.cuh) assume one blend CTA per tile");

void set_error(const char* fmt, ...);

// Optional per-stage device timing (cudaEvents on the launch stream) and a count of
// kernels launched by this library; both are read by bench.py through the C ABI.
enum Stage {
  ST_SAMPLE_FWD = 0, ST_PREPROCESS_FWD, ST_SCAN, ST_EMIT_KEYS, ST_SORT, ST_TILE_RANGES, ST_GATHER, ST_BLEND_FWD,
  ST_BLEND_BWD, ST_PREPROCESS_BWD, ST_SAMPLE_BWD, ST_SSIM_FWD, ST_SSIM_BWD, ST_KNN, ST_ACTIVATE_FWD, ST_ACTIVATE_BWD,
  ST_LOSS_FWD, ST_LOSS_BWD, ST_COUNT
};
void count_launches(int n);
struct StageTimer {
  int stage; cudaStream_t st; void* rec;
  StageTimer(int stage_, cudaStream_t st_, int kernels);
  ~StageTimer();
};

// Programmatic dependent launch (sm_90+): every kernel of the library is launched with the
// programmatic-stream-serialization attribute and starts with pdl_wait(). The next kernel's CTAs can
// then be scheduled while the previous grid drains (its launch latency and ramp-up overlap the tail),
// and block in griddepcontrol.wait until the previous grid has completed and its writes are visible -
// the same ordering as a plain in-stream launch, minus the few-microsecond gap at each of the ~45 kernel
// boundaries of a step. Kernels that precede ours in the stream need no cooperation.
#ifdef __CUDACC__
// wait for the previous grid, then let the NEXT grid's CTAs be staged as soon as every CTA of this grid has
// started (they block in their own wait until this grid is complete)
__device__ __forceinline__ void pdl_wait() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

extern int g_use_pdl;   // cg_set_pdl(); 1 by default

template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                            Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = g_use_pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}
#endif

#define CG_CUDA(expr)                                                        \
  do {                                                                       \
    cudaError_t _e = (expr);                                                 \
    if (_e != cudaSuccess) {                                                 \
      cg::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #expr,             \
                    cudaGetErrorString(_e));                                 \
      return CG_ERR_CUDA;                                                    \
    }                                                                        \
  } while (0)

#define CG_LAUNCH_CHECK(debug, stream)                                       \
  do {                                                                       \
    CG_CUDA(cudaGetLastError());                                             \
    if (debug) CG_CUDA(cudaStreamSynchronize(stream));                       \
  } while (0)

#define CG_ARG(cond, msg)                                                    \
  do {                                                                       \
    if (!(cond)) {                                                           \
      cg::set_error("bad argument: %s (%s)", msg, #cond);                    \
      return CG_ERR_ARG;                                                     \
    }                                                                        \
  } while (0)

// Carve 128-byte aligned typed arrays out of one caller-owned byte blob
// (same idea as rasterizer_impl.h:21-27, but sizes are computed with the very
// same walk so size query and carving cannot disagree).
struct Carver {
  char* p;
  size_t used;
  explicit Carver(void* base) : p(reinterpret_cast<char*>(base)), used(0) {}
  template <typename T>
  T* take(size_t count) {
    size_t off = (used + 127) & ~size_t(127);
    used = off + count * sizeof(T);
    return p ? reinterpret_cast<T*>(p + off) : nullptr;
  }
};

// One projected Gaussian as the blend kernels consume it (GeomState::grec, written once per Gaussian by
// preprocess_fwd). There is no per-tile-instance copy of it: the blend kernels gather the records of a batch of
// their tile's sorted list by Gaussian index (three 16-byte async copies per record, from a table that stays in
// L2: 48 B x P), instead of streaming 48 B x R materialised, sorted records of which the forward's early
// termination leaves three quarters unread. The quads are laid out for the ring backward, whose per-step element
// switch moves whole register quads (blend_ring.cuh).
struct __align__(16) Rec {
  float ca, cb, cc;    // conic xx, xy, yy          (geomState.conic_opacity.xyz)
  float invd;          // 1 / view-space depth
  float x, y;          // pixel-space mean           (geomState.means2D)
  float o, col;        // opacity (conic_opacity.w), colour (1 channel)
  float m0, m1, m2, m3;// all_map channels
};
static_assert(sizeof(Rec) == 48, "record must stay 48 bytes");

#ifdef __CUDACC__
// Can this instance contribute to ANY pixel centre of the block [bx0,bx1] x [by0,by1]?
//
// A pixel contributes only if alpha = min(0.99, o*exp(power)) >= 1/255 (forward.cu:361-363), i.e.
// q(d) = -power = 0.5 d^T Q d <= tau with tau = ln(255 o). q is convex with its minimum (0) at the
// splat centre, so its minimum over the block is 0 if the centre is inside, and otherwise lies on the
// block edge(s) facing the centre, where it is a clamped 1-D quadratic. The instance is skipped only if
// that minimum exceeds tau by more than a bound on the fp32 evaluation error of `power` anywhere in
// the block (E below) plus 2e-3; NaNs and non-positive-definite conics are never skipped. Skipping is
// therefore invisible in the results (bit-identical images, n_contrib and gradients).
__device__ __forceinline__ bool block_candidate(float cx, float cy, float A, float B, float C, float o,
                                                float bx0, float bx1, float by0, float by1) {
  const float chk = cx + cy + A + B + C + o;
  if (chk != chk) return true;                           // NaN anywhere: let the exact path decide
  if (!(A > 0.f && C > 0.f && A * C - B * B > 0.f)) return true;
  if (o <= 0.f) return false;                            // alpha <= 0 < 1/255 for every pixel
  const float tau = fmaxf(__logf(255.0f * o), 0.0f) + 2e-3f;
  const float dxn = fminf(fmaxf(cx, bx0), bx1) - cx;     // offset to the nearest block column (0 if inside)
  const float dyn = fminf(fmaxf(cy, by0), by1) - cy;
  const float Dx = fmaxf(fabsf(bx0 - cx), fabsf(bx1 - cx));
  const float Dy = fmaxf(fabsf(by0 - cy), fabsf(by1 - cy));
  const float E = 1e-6f * (A + C + 2.f * fabsf(B)) * (Dx * Dx + Dy * Dy);
  float qmin = 0.f;
  if (dxn != 0.f || dyn != 0.f) {
    qmin = __int_as_float(0x7f800000);
    if (dxn != 0.f) {
      const float dy = fminf(fmaxf(__fdividef(-B * dxn, C), by0 - cy), by1 - cy);
      qmin = 0.5f * (A * dxn * dxn + C * dy * dy) + B * dxn * dy;
    }
    if (dyn != 0.f) {
      const float dx = fminf(fmaxf(__fdividef(-B * dyn, A), bx0 - cx), bx1 - cx);
      qmin = fminf(qmin, 0.5f * (A * dx * dx + C * dyn * dyn) + B * dx * dyn);
    }
  }
  return !(qmin > tau + E);
}
#endif

// Radix sort geometry (see sort.cu).
constexpr int SORT_THREADS = 256;
#ifndef CG_SORT_ITEMS
#define CG_SORT_ITEMS 16
#endif
constexpr int SORT_ITEMS = CG_SORT_ITEMS;
constexpr int SORT_TILE = SORT_THREADS * SORT_ITEMS;  // 4096 keys per CTA
constexpr int SORT_MAX_PASSES = 8;

// Ping-pong buffers + look-back state of one radix sort of n (key, value) pairs.
template <typename K>
struct SortBufs {
  K* keys[2];
  uint32_t* vals[2];
  uint32_t* hist;     // [SORT_MAX_PASSES][256] global digit histograms -> exclusive bases
  uint32_t* ticket;   // [SORT_MAX_PASSES] dynamic tile id counters
  uint32_t* status;   // [passes][ntiles][256] decoupled look-back words
  size_t status_words;
  static SortBufs carve(Carver& c, int64_t n) {
    SortBufs b;
    int64_t ntiles = (n + SORT_TILE - 1) / SORT_TILE;
    b.keys[0] = c.take<K>(n + 1);
    b.keys[1] = c.take<K>(n + 1);
    b.vals[0] = c.take<uint32_t>(n + 1);
    b.vals[1] = c.take<uint32_t>(n + 1);
    b.hist = c.take<uint32_t>(SORT_MAX_PASSES * 256);
    b.ticket = c.take<uint32_t>(32);
    b.status_words = size_t(sizeof(K)) * size_t(ntiles > 0 ? ntiles : 1) * 256;   // sizeof(K) = max passes
    b.status = c.take<uint32_t>(b.status_words);
    return b;
  }
};

// Super-tile binning geometry (binning.cuh).
constexpr int ST_SHIFT = 3;                    // super-tile = 8 x 8 tiles
constexpr int ST_SIDE = 1 << ST_SHIFT;
constexpr int BIN_EPL = 4;                     // copies per lane
constexpr int BIN_WCHUNK = 32 * BIN_EPL;       // copies per warp
constexpr int BIN_CHUNK = 8 * BIN_WCHUNK;      // copies per chunk (one CTA of 8 warps)
constexpr int BIN_MAX_SUPERTILES = 16384;      // (a 16384 x 16384 pixel image); larger images use the sort path

struct GeomState {
  Rec* grec;         // per Gaussian: conic, 1/depth, mean2D, opacity, colour, all_map (valid where tiles > 0)
  float* depth;
  uint32_t* tiles;
  uint2* rect;       // packed tile rect: x = min.x | min.y<<16, y = max.x | max.y<<16
  uint32_t* blk_sum;
  uint32_t* blk_prefix;
  uint32_t* total;   // [0] = R; [1] = number of (super-tile, Gaussian) copies (super-tile binning)
  SortBufs<uint32_t> gs;   // the P Gaussians sorted by depth bits (value = Gaussian index), see BinScratch
  static GeomState carve(void* base, int64_t P, size_t* bytes) {
    Carver c(base);
    GeomState g;
    int64_t nblk = (P + 255) / 256;
    g.grec = c.take<Rec>(P + 1);
    g.depth = c.take<float>(P);
    g.tiles = c.take<uint32_t>(P);
    g.rect = c.take<uint2>(P);
    g.blk_sum = c.take<uint32_t>(nblk);
    g.blk_prefix = c.take<uint32_t>(nblk);
    g.total = c.take<uint32_t>(32);
    g.gs = SortBufs<uint32_t>::carve(c, P);
    if (bytes) *bytes = (c.used + 127) & ~size_t(127);
    return g;
  }
};

struct ImgState {
  float* final_T;
  uint32_t* n_contrib;
  uint2* ranges;
  uint32_t* tile_maxc;  // per blend CTA (tile * BLEND_SUBS + sub): max n_contrib over its pixels (backward start)
  uint32_t* tile_order; // blend CTA ids, longest list first: the launch order of the blend CTAs (forward, and the
                        // lane-per-pixel backward)
  // 4x4 pixel blocks (16 per tile, block id = tile * 16 + warp * 2 + half: the left / right half of a forward warp's
  // 8x4 block), written by blend_fwd for the ring backward:
  uint32_t* blk_cnt;    // entries of the block's contributor list (BinKeep::cand)
  uint32_t* cls_count;  // [RING_CLASSES] non-empty blocks per size class (class = half-octave of blk_cnt), then
                        // [RING_CLASSES], [RING_CLASSES + 1] = work / exit counters of blend_bwd_ring (left at zero by it);
                        // zero-filled together with `ranges` at the start of every forward
  uint4* cls_list;      // [RING_CLASSES][tiles * 16] block descriptors per class, in arrival order:
                        // {block id, list length, tile range start, tile range length} - everything the ring backward
                        // needs to walk the block, in one 16-byte load
  // super-tile binning (binning.cuh); tile_cnt and st_ranges are zero-filled with `ranges`
  uint32_t* tile_cnt;    // [tiles] instances per tile
  uint2* st_ranges;      // [super-tiles] span of each super-tile in the sorted copy list
  uint32_t* chunk_start; // [super-tiles + 1] first chunk of each super-tile
  static ImgState carve(void* base, int W, int H, size_t* bytes) {
    Carver c(base);
    ImgState s;
    size_t npix = size_t(W) * H;
    size_t tiles = size_t((W + TILE_X - 1) / TILE_X) * ((H + TILE_Y - 1) / TILE_Y);
    s.final_T = c.take<float>(npix);
    s.n_contrib = c.take<uint32_t>(npix);
    s.ranges = c.take<uint2>(tiles);
    s.cls_count = c.take<uint32_t>(RING_CLASSES + 32);   // directly behind ranges: one memset clears all four
    const size_t nst = size_t((W + TILE_X * ST_SIDE - 1) / (TILE_X * ST_SIDE)) * ((H + TILE_Y * ST_SIDE - 1) / (TILE_Y * ST_SIDE));
    s.tile_cnt = c.take<uint32_t>(tiles);
    s.st_ranges = c.take<uint2>(nst);
    s.chunk_start = c.take<uint32_t>(nst + 1);
    s.tile_maxc = c.take<uint32_t>(tiles * BLEND_SUBS);
    s.tile_order = c.take<uint32_t>(tiles * BLEND_SUBS);
    s.blk_cnt = c.take<uint32_t>(tiles * 16);
    s.cls_list = c.take<uint4>(size_t(RING_CLASSES) * tiles * 16);
    if (bytes) *bytes = (c.used + 127) & ~size_t(127);
    return s;
  }
};

struct BinKeep {
  uint32_t* point_list;   // Gaussian index per sorted tile-instance (the reference's point_list)
  // Contributor lists of the 4x4 pixel blocks: block b (0..15) of a tile whose range is [x, y) owns
  // cand[16 * x + b * (y - x) ...]: the tile-relative list positions, ascending, of the instances that the forward
  // blended into at least one pixel of the block, and cand_id[...] their Gaussian indices (written by blend_fwd,
  // consumed back to front by blend_bwd_ring). 128 bytes of address space per instance, of which the forward writes
  // about 13 (1.6 entries per instance-and-8x4-block pair that contributes).
  uint32_t* cand;
  uint32_t* cand_id;
  static BinKeep carve(void* base, int64_t R, size_t* bytes) {
    Carver c(base);
    BinKeep b;
    b.point_list = c.take<uint32_t>(R + 1);
    b.cand = c.take<uint32_t>(16 * size_t(R + 1));
    b.cand_id = c.take<uint32_t>(16 * size_t(R + 1));
    if (bytes) *bytes = (c.used + 127) & ~size_t(127);
    return b;
  }
};

// Binning (binning.cuh) orders the Gaussians once and never sorts the R tile instances:
//   GeomState::gs: the P Gaussians sorted by the bits of their view-space depth (value = Gaussian index);
//                  runs inside cg_raster_fwd_geom, overlapping the host's wait for R
//   BinScratch::is (forward-only scratch): one copy of a Gaussian per 8x8-tile super-tile it overlaps, emitted in
//                  that depth order and stably sorted by super-tile index (one small pass); the per-tile lists are
//                  then filled by counting (ccnt / cbase)
// A tile's list is therefore in (depth, index) order, i.e. exactly the permutation the reference gets from one 64-bit
// sort of tile<<32|depth over the R instances (rasterizer_impl.cu:70-111, :309-314).
// With CURVEGS_BINNING=sort `is` holds the R (tile, id) instances instead, emitted in depth order and stably sorted
// by tile index only (two 16-byte passes over R): the round-1 path, kept as fallback and yardstick.
struct BinScratch {
  SortBufs<uint32_t> is;
  // super-tile binning (binning.cuh): per chunk of BIN_CHUNK copies and tile of its super-tile
  uint32_t* ccnt;    // [chunks][64] copies of the chunk covering the tile
  uint32_t* cbase;   // [chunks][64] the same, summed over the super-tile's earlier chunks
  static size_t max_chunks(int64_t R) { return size_t(R > 0 ? R : 0) / BIN_CHUNK + BIN_MAX_SUPERTILES + 1; }
  static BinScratch carve(void* base, int64_t P, int64_t R, size_t* bytes) {
    (void)P;
    Carver c(base);
    BinScratch b;
    b.is = SortBufs<uint32_t>::carve(c, R);
    b.ccnt = c.take<uint32_t>(max_chunks(R) * 64);
    b.cbase = c.take<uint32_t>(max_chunks(R) * 64);
    if (bytes) *bytes = (c.used + 127) & ~size_t(127);
    return b;
  }
};

// Number of tile-index bits that take part in the sort. The reference finds it
// with a halving search (rasterizer_impl.cu:35-50, getHigherMsb) whose result
// is the bit length of n for n >= 1 and 1 for n == 0; bits above it are zero
// in every key, so the sorted order does not depend on this being tight.
inline uint32_t tile_key_bits(uint32_t n) {
  uint32_t bits = 0;
  while (n) { ++bits; n >>= 1; }
  return bits ? bits : 1;
}

// Stable LSD radix sort of b.keys[0]/b.vals[0] on bits [0,end_bit); *out_buf tells which of the
// two ping-pong buffers holds the result. Instantiated for uint32_t and uint64_t keys (sort.cu).
// With d_n != NULL the number of pairs is min(n, *d_n), read on the device: n is then only the capacity the grids
// and buffers are sized for (sync-free binning, cg_raster_fwd_capacity).
template <typename K>
int radix_sort_pairs(const SortBufs<K>& b, int64_t n, int end_bit, int* out_buf, bool debug, cudaStream_t stream,
                     const uint32_t* d_n = nullptr, bool hist_ready = false, uint32_t* final_vals = nullptr);
// (final_vals != NULL: the last pass writes the sorted VALUES there instead of into its ping-pong buffer; the sorted
//  keys still land in b.keys[*out_buf])
// A producer that writes b.keys[0] can fill the digit histograms itself: radix_sort_begin (clears b.hist and the
// look-back state) -> producer adds, for every key and pass p, one count to b.hist[p * 256 + digit_p(key)]
// (digit geometry from radix_sort_geometry) -> radix_sort_pairs(..., hist_ready = true) skips its histogram pass.
void radix_sort_geometry(int end_bit, int key_bits, int* passes, int* bpp);
template <typename K>
int radix_sort_begin(const SortBufs<K>& b, int64_t n, int end_bit, cudaStream_t stream);
// Which ping-pong buffer radix_sort_pairs leaves the result in (number of digit passes is ceil(end_bit / 8)).
inline int radix_sort_result_buf(int end_bit) { int p = (end_bit + 7) / 8; return (p < 1 ? 1 : p) & 1; }

}  // namespace cg
