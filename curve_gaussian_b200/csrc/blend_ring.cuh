// Ring backward of the alpha blend (backward.cu:451-675 semantics), colour-only variant.
//
// blend_bwd (raster_bwd.cu) maps lanes to pixels, so every (8x4 block, instance) pair ends in a 32-lane
// reduction of 8 terms plus 8 atomics: about a third of its instructions. Here the roles are swapped: a
// lane owns one INSTANCE of a 4x4 pixel block's contributor list (written by blend_fwd: the instances blended
// into at least one pixel of the block) for 16 consecutive steps and accumulates that instance's gradient sums in
// registers, while the 16 PIXELS of the block travel around a ring of 16 lanes, one lane per step, carrying their
// blend state (T, the colour seen behind, dL/dpixel, n_contrib, coordinates) in shuffles. A warp runs two such
// rings (lanes 0-15 and 16-31) over independent block streams. Pixel p meets the instances in list order back to
// front, exactly as in the reference's per-pixel loop, so every per-pixel quantity (T / (1 - alpha) chain,
// accum_rec recursion, alpha, G) has the same bits; only the order of the sum over pixels differs. There is no
// cross-lane reduction at all, and one instance ends in two 128-bit vector reductions (REDG.F32x4).
//
// A ring never drains between blocks: it walks the concatenation of the lists of all the blocks assigned to it.
// Lane i of a ring owns sequence elements i, i+16, ... ("epoch" m = elements 16m..16m+15) and switches to its next
// element at step 16m+i, when pixel 0 arrives; an element flagged as the first of its block re-initialises every
// pixel slot that reaches it from that block's pixel table. Everything that is not per step happens at the epoch
// boundaries, for all lanes at once: list positions are loaded two epochs ahead, the 32-byte records one epoch
// ahead (cp.async from the per-Gaussian table straight into the lane's staging slot), the sums of finished
// instances are flushed one epoch late.
//
// Gradient sums kept per instance (q = o * G * dL/dalpha, d = mean2D - pixel):
//   S0 = sum G dL/dalpha           -> dL/dopacity
//   S1 = sum q dx, S2 = sum q dy   -> dL/dmean2D = -(0.5 W)(A S1 + B S2), -(0.5 H)(C S2 + B S1)
//   S3, S4, S5 = sum q dx dx, q dx dy, q dy dy -> dL/dconic = -0.5 (S3, S4, S5)
//   S7 = sum alpha T dL/dpixel     -> dL/dcolour
#pragma once
#include "common.cuh"
#include "math.cuh"
#include "stage.cuh"

namespace cg {

#ifndef CG_RING_WARPS
#define CG_RING_WARPS 4
#endif
constexpr int RING_WARPS = CG_RING_WARPS;     // per CTA; warps are independent
#ifndef CG_RING_CTAS
#define CG_RING_CTAS 6
#endif
#ifndef CG_RING_UNROLL
#define CG_RING_UNROLL 1
#endif
constexpr int RING_UNROLL = CG_RING_UNROLL;   // steps per loop iteration
#ifndef CG_RING_NB
#define CG_RING_NB 2
#endif
constexpr int RING_LANES = 16;    // lanes (= pixels = elements per epoch) of one ring; two rings per warp
constexpr int RING_NB = CG_RING_NB;   // blocks that may start within one epoch (each needs a pixel table)
constexpr int RING_TSLOTS = 3 * RING_NB;
constexpr uint32_t RING_POS_MASK = 0x07ffffffu;   // list position within the tile; bits 27-30 table slot, bit 31 first-of-block
static_assert(RING_TSLOTS <= 16, "table slot field is 4 bits");

struct __align__(16) RingWarp {
  // (the quads are chosen so that every vector load / store of the per-step switch moves whole register quads:
  //  no register is in two of them)
  float4 eA[2][32];               // staged elements of an epoch: x, y, opacity, colour
  float4 eB[2][32];               //                              conic a, b, c, position|flags
  float4 snap[32][3];             // sums of the element a lane just finished: S0..S3 | S4 S5 S7 - | conic a b c -
  uint32_t gid[4][32];            // Gaussian index of the elements of an epoch (for the flush, two epochs later)
  float4 tQ[RING_TSLOTS][32];     // pixel table of a block that starts in an epoch: final_T, dL/dpixel, n_contrib, x
  float tY[RING_TSLOTS][32];      //   y
  float4 zero;                    // (0, 0, 0, 0): what the switch reloads the sums from
  uint32_t pre[RING_CLASSES + 1]; // blocks in the k largest size classes
};

__device__ __forceinline__ void cp_async16(void* smem, const void* g) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem)), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_async8(void* smem, const void* g) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(smem)), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_async4(void* smem, const void* g) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(smem)), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

__device__ __forceinline__ float ring_div(float a, float d) {
  // same sequence as div_rn_normal in raster_bwd.cu (tests/test_div_sequence.py)
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(d));
  r = __fmaf_rn(r, __fmaf_rn(-d, r, 1.0f), r);
  const float q = __fmul_rn(a, r);
  return __fmaf_rn(r, __fmaf_rn(-d, q, a), q);
}

__global__ void __launch_bounds__(RING_WARPS * 32, CG_RING_CTAS)
blend_bwd_ring(const uint2* __restrict__ ranges, const uint32_t* __restrict__ cls_count,
               const uint4* __restrict__ cls_list, uint32_t nblocks, uint32_t* __restrict__ work,
               const uint32_t* __restrict__ blk_cnt, int grid_x,
               const Rec* __restrict__ grec, const uint32_t* __restrict__ cand, const uint32_t* __restrict__ cand_id,
               int W, int H, float ddelx_dx, float ddely_dy, const float* __restrict__ bg,
               const float* __restrict__ final_T, const uint32_t* __restrict__ n_contrib,
               const float* __restrict__ dL_dpix, float* __restrict__ acc) {
  pdl_wait();
  __shared__ RingWarp s_ring[RING_WARPS];
  constexpr unsigned FULL = 0xffffffffu;
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  RingWarp& rw = s_ring[warp];
  // The background is consumed HERE (the empty asm makes the loaded value a plain register value from now on). Left to
  // the compiler, the first use of the load sits at the top of the step loop with a wait on the load's scoreboard -
  // a scoreboard that the candidate-position loads of every epoch boundary share, so every epoch's first step would
  // wait for loads that are meant to be consumed a whole epoch later (measured: 5.7 % of the warps' stall samples).
  float bg0 = bg[0];
  uint32_t has_bg_u = bg0 != 0.f ? 1u : 0u;
  asm volatile("" : "+f"(bg0), "+r"(has_bg_u));
  const bool has_bg = has_bg_u != 0u;   // uniform: with a black background its term is an exact zero
  // the warp runs TWO rings of 16 lanes (one 4x4 half of an 8x4 block each): hl = lane within its ring, hbase = the
  // ring's first lane. Everything the assignment code calls "uniform" is uniform per ring, and its votes / shuffles
  // name the ring's lanes only; the step loop itself is executed by both rings in lockstep.
  const uint32_t hl = lane & 15u, hbase = lane & 16u, ring = lane >> 4;
  const unsigned hmask = 0xffffu << hbase;
  const int src = int(hbase | ((hl + 15u) & 15u));

  // ---- a ring's blocks: taken one at a time from a global counter, size classes largest first (blend_fwd
  // filed every non-empty block under the half-octave of its list length): greedy list scheduling over a
  // nearly sorted list, so the rings finish within one small block of each other ----
  // s_pre[k] = number of blocks in the k largest classes
  {
    const uint32_t c = cls_count[RING_CLASSES - 1 - lane];
    uint32_t inc = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t n = __shfl_up_sync(FULL, inc, o); if (int(lane) >= o) inc += n; }
    rw.pre[lane + 1] = inc;
    if (lane == 0) { rw.pre[0] = 0; rw.zero = make_float4(0.f, 0.f, 0.f, 0.f); }
    __syncwarp();
  }
  const uint32_t n_items = rw.pre[RING_CLASSES];
  // queue of block descriptors, entry j kept by lane (j & 3); claimed and loaded four blocks ahead
  uint32_t q_bid = 0xffffffffu, q_cnt = 0, q_rx = 0, q_len = 0;
  // (the first four blocks of a warp are fixed - ranks gw, 2 NW - 1 - gw, 2 NW + gw, 4 NW - 1 - gw of the order, so that
  // no warp opens with four of the very largest; everything after that is claimed dynamically)
  const uint32_t gw = (blockIdx.x * RING_WARPS + warp) * 2u + ring, NW = gridDim.x * RING_WARPS * 2u;
  auto load_meta = [&](int fixed) {
    const uint32_t at = fixed >= 0 ? uint32_t(fixed) * NW + ((fixed & 1) ? NW - 1u - gw : gw) : 4u * NW + atomicAdd(work, 1u);
    q_bid = 0xffffffffu; q_cnt = 0; q_rx = 0; q_len = 0;
    if (at < n_items) {
      uint32_t k = 0;   // largest k with pre[k] <= at
#pragma unroll
      for (uint32_t step = RING_CLASSES / 2; step > 0; step >>= 1)
        if (rw.pre[k + step] <= at) k += step;
      // one 16-byte descriptor (written by blend_fwd when it filed the block): a chain of three dependent loads
      // (block id -> count, tile -> range) stalled the whole warp inside this function
      const uint4 d = cls_list[size_t(RING_CLASSES - 1 - k) * nblocks + (at - rw.pre[k])];
      q_bid = d.x; q_cnt = d.y; q_rx = d.z; q_len = d.w;
    }
  };
  if (hl < 4) load_meta(int(hl));
  uint32_t j_next = 0;            // next queue entry to start
  uint32_t cb_rem = 0, cb_base = 0;   // block being consumed: entries left (taken from the end), list base

  // assignment of epoch E: which list entry each lane takes; a_addr = index into cand (or ~0), a_flag = first|slot bits.
  // At most RING_NB blocks start per epoch (they need a pixel table each); empty blocks are skipped. An epoch
  // without any element therefore means that the warp has run out of blocks.
  auto assign = [&](int E, uint32_t& a_addr, uint32_t& a_flag, uint32_t (&nb)[RING_NB]) -> uint32_t {
    a_addr = 0xffffffffu; a_flag = 0;
#pragma unroll
    for (int k = 0; k < RING_NB; ++k) nb[k] = 0xffffffffu;
    uint32_t taken = min(cb_rem, uint32_t(RING_LANES));
    if (hl < taken) a_addr = cb_base + (cb_rem - 1u - hl);
    cb_rem -= taken;
    int started = 0;
    while (taken < uint32_t(RING_LANES) && started < RING_NB) {
      const int qs = int(hbase | (j_next & 3u));
      const uint32_t bid = __shfl_sync(hmask, q_bid, qs);
      if (bid == 0xffffffffu) break;
      {
        const uint32_t cnt = __shfl_sync(hmask, q_cnt, qs), rx = __shfl_sync(hmask, q_rx, qs), len = __shfl_sync(hmask, q_len, qs);
        if (int(lane) == qs) load_meta(-1);
        ++j_next;
        if (cnt != 0u) {
          const uint32_t base = 16u * rx + (bid & 15u) * len;
          const uint32_t n = min(cnt, uint32_t(RING_LANES) - taken);
          if (hl >= taken && hl < taken + n) {
            a_addr = base + (cnt - 1u - (hl - taken));
            if (hl == taken) a_flag = 0x80000000u | (uint32_t((E % 3) * RING_NB + started) << 27);
          }
#pragma unroll
          for (int k = 0; k < RING_NB; ++k) if (k == started) nb[k] = bid;
          ++started;
          cb_rem = cnt - n; cb_base = base;
          taken += n;
        }
      }
    }
    return taken;
  };

  // ---- pixel slot state (travels lane -> lane + 1 every step) ----
  float T = 0.f, Rp = 0.f, dLp = 0.f, pxf = 0.f, pyf = 0.f, Tf = 0.f;
  int nc = 0;
  // ---- element state (this lane's current instance) ----
  float ex = 0.f, ey = 0.f, ca = 0.f, cb = 0.f, cc = 0.f, eo = 0.f, col = 0.f;
  float poswf = __uint_as_float(RING_POS_MASK);   // position | flags word of the element, kept as the bits of eB.w
  float S0 = 0.f, S1 = 0.f, S2 = 0.f, S3 = 0.f, S4 = 0.f, S5 = 0.f, S7 = 0.f;
  float spad = 0.f;   // fourth register of the {S4, S5, S7, -} quad the switch stores with one vector store

  // epoch bookkeeping: A = epoch m+1 (positions loaded, records not yet), B = epoch m+2 (being assigned)
  uint32_t posA = 0, idA = 0, flagA = 0, validA = 0;
  uint32_t nbA[RING_NB];
#pragma unroll
  for (int k = 0; k < RING_NB; ++k) nbA[k] = 0xffffffffu;
  uint32_t posw_pending = RING_POS_MASK;     // position|flags word of this lane's element of the next epoch
  uint32_t real_m2 = 0, real_m1 = 0, real_0 = 0, real_p1 = 0;   // real elements of epochs m-2, m-1, m, m+1

  for (int m = -2;; ++m) {
    // ================= epoch boundary U_m =================
    if (m >= 0) {
      cp_async_wait_all();                                   // epoch m records + tables (issued at U_{m-1})
      reinterpret_cast<uint32_t*>(&rw.eB[m & 1][lane])[3] = posw_pending;
      __syncwarp();
    }
    if (m >= 2 && real_m2 != 0u) {
      // flush the sums of epoch m-2 (snapshots taken when the lanes switched during period m-1)
      const uint32_t id = rw.gid[(m - 2) & 3][lane];
      if (id != 0xffffffffu) {
        const float4 s0 = rw.snap[lane][0], s1 = rw.snap[lane][1], s2 = rw.snap[lane][2];
        const float g0 = -ddelx_dx * (s2.x * s0.y + s2.y * s0.z);
        const float g1 = -ddely_dy * (s2.z * s0.z + s2.y * s0.y);
        float* dst = acc + size_t(id) * 8;
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(g0), "f"(g1), "f"(-0.5f * s0.w), "f"(-0.5f * s1.x) : "memory");
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + 4), "f"(-0.5f * s1.y), "f"(s0.x), "f"(s1.z), "f"(0.f) : "memory");
      }
    }
    // nothing in flight and nothing left, in either ring (a ring that runs dry first idles through empty epochs)
    if (__all_sync(FULL, m >= 1 ? real_m1 == 0u : (m == 0 && real_0 == 0u))) break;
    {
      // records + Gaussian ids of epoch E = m+1, pixel tables of the blocks that start in it
      const int E = m + 1;
      if (E >= 0) {
        if (validA) {
          const Rec* r = grec + idA;                    // the per-Gaussian record, gathered by index (L2-resident table)
          cp_async16(&rw.eA[E & 1][lane], &r->x);       // x, y, opacity, colour
          cp_async16(&rw.eB[E & 1][lane], &r->ca);      // conic a, b, c (+ 1/depth, overwritten by the position word)
          rw.gid[E & 3][lane] = idA;
          posw_pending = posA | flagA;
        } else {
          rw.eA[E & 1][lane] = make_float4(0.f, 0.f, 0.f, 0.f);
          rw.eB[E & 1][lane] = make_float4(0.f, 0.f, 0.f, 0.f);
          rw.gid[E & 3][lane] = 0xffffffffu;
          posw_pending = RING_POS_MASK;
        }
#pragma unroll
        for (int k = 0; k < RING_NB; ++k) {
          const uint32_t nb = nbA[k];
          if (nb != 0xffffffffu) {
            const int ts = (E % 3) * RING_NB + k;
            const uint32_t tile = nb >> 4, b = (nb >> 1) & 7u;
            const uint32_t bx0 = (tile % uint32_t(grid_x)) * TILE_X + (b & 1u) * 8u + (nb & 1u) * 4u;
            const uint32_t by0 = (tile / uint32_t(grid_x)) * TILE_Y + (b >> 1) * 4u;
            const uint32_t px = bx0 + (hl & 3u), py = by0 + (hl >> 2);
            float* q = reinterpret_cast<float*>(&rw.tQ[ts][lane]);
            if (px < uint32_t(W) && py < uint32_t(H)) {
              const size_t pid = size_t(W) * py + px;
              cp_async4(q, final_T + pid);
              cp_async4(q + 1, dL_dpix + pid);
              cp_async4(q + 2, n_contrib + pid);
            } else {
              q[0] = 0.f; q[1] = 0.f; q[2] = 0.f;
            }
            q[3] = float(px);
            rw.tY[ts][lane] = float(py);
          }
        }
      }
    }
    {
      // assignment of epoch m+2 and its list positions (consumed at the next boundary)
      uint32_t a_addr, a_flag;
      const uint32_t real_p2 = assign(m + 2, a_addr, a_flag, nbA);
      validA = a_addr != 0xffffffffu;
      posA = validA ? cand[a_addr] : 0u;
      idA = validA ? cand_id[a_addr] : 0u;
      flagA = a_flag;
      real_m2 = real_m1; real_m1 = real_0; real_0 = real_p1; real_p1 = real_p2;
    }
    if (m < 0) continue;

    // ================= the 16 steps of period m =================
    const uint32_t stA_addr = smem_u32(&rw.eA[m & 1][lane]), stB_addr = smem_u32(&rw.eB[m & 1][lane]);
    const uint32_t snap_addr = smem_u32(&rw.snap[lane][0]);
    const uint32_t zero_addr = smem_u32(&rw.zero);
    // The element switch: pixel 0 arrives at lane u of each ring at step u, when that lane's previous element has seen
    // all 16 pixels. One lane per ring and step takes this path, so it is written as predicated vector accesses
    // straight from / into the registers that hold the sums and the element (a branch plus compiler-chosen
    // temporaries costs 8 more issue slots per step). It is issued at the END of the step before (for lane 0: ahead of
    // the loop), so that the latency of its loads runs under the next step's shuffles instead of in front of the
    // first use of the element.
#define RING_SWITCH(WHICH)                                                                                        \
    asm volatile(                                                                                                  \
        "{\n\t.reg .pred sw;\n\t"                                                                                 \
        "setp.eq.u32 sw, %16, %17;\n\t"                                                                            \
        "@sw st.shared.v4.f32 [%18], {%8, %9, %10, %11};\n\t"                                                      \
        "@sw st.shared.v4.f32 [%18+16], {%12, %13, %14, %15};\n\t"                                                 \
        "@sw st.shared.v4.f32 [%18+32], {%4, %5, %6, %7};\n\t"                                                     \
        "@sw ld.shared.v4.f32 {%0, %1, %2, %3}, [%19];\n\t"                                                        \
        "@sw ld.shared.v4.f32 {%4, %5, %6, %7}, [%20];\n\t"                                                        \
        "@sw ld.shared.v4.f32 {%8, %9, %10, %11}, [%21];\n\t" /* (zeros: two loads instead of seven moves) */      \
        "@sw ld.shared.v4.f32 {%12, %13, %14, %15}, [%21];\n\t}"                                                   \
        : "+f"(ex), "+f"(ey), "+f"(eo), "+f"(col), "+f"(ca), "+f"(cb), "+f"(cc), "+f"(poswf),                      \
          "+f"(S0), "+f"(S1), "+f"(S2), "+f"(S3), "+f"(S4), "+f"(S5), "+f"(S7), "+f"(spad)                         \
        : "r"(hl), "r"(WHICH), "r"(snap_addr), "r"(stA_addr), "r"(stB_addr), "r"(zero_addr)                        \
        : "memory")
    RING_SWITCH(0u);
#pragma unroll RING_UNROLL
    for (uint32_t u = 0; u < uint32_t(RING_LANES); ++u) {
      T = __shfl_sync(FULL, T, src);
      Rp = __shfl_sync(FULL, Rp, src);
      dLp = __shfl_sync(FULL, dLp, src);
      nc = __shfl_sync(FULL, nc, src);
      pxf = __shfl_sync(FULL, pxf, src);
      pyf = __shfl_sync(FULL, pyf, src);
      if (has_bg) Tf = __shfl_sync(FULL, Tf, src);
      const uint32_t posw = __float_as_uint(poswf);
      if (posw & 0x80000000u) {
        // first element of its block (back to front): the arriving slot becomes pixel p of that block
        const uint32_t ts = (posw >> 27) & 15u, p = (u - hl) & 15u, tp = hbase | p;
        const float4 q = rw.tQ[ts][tp];
        T = q.x;
        Tf = T;
        dLp = q.y;
        nc = __float_as_int(q.z);
        Rp = 0.f;
        pxf = q.w;
        pyf = rw.tY[ts][tp];
      }
      bool contrib = int(posw & RING_POS_MASK) < nc;
      const float dx = __fsub_rn(ex, pxf), dy = __fsub_rn(ey, pyf);
      const float power = gauss_power(ca, cb, cc, dx, dy);
      contrib = contrib && !(power > 0.0f);
      const float G = expf(power);
      const float alpha = fminf(0.99f, __fmul_rn(eo, G));
      contrib = contrib && !(alpha < 1.0f / 255.0f);
      if (contrib) {
        const float oma = 1.f - alpha;
        T = ring_div(T, oma);
        const float w = alpha * T;
        float dL_dalpha = (col - Rp) * dLp;
        Rp = __fmaf_rn(col, alpha, __fmul_rn(oma, Rp));   // accum_rec as the next contributing instance will see it
        dL_dalpha *= T;
        if (has_bg) dL_dalpha += (-Tf / oma) * (bg0 * dLp);
        const float h = G * dL_dalpha;
        S0 += h;
        const float q = eo * h;
        const float qx = q * dx, qy = q * dy;
        S1 += qx;
        S2 += qy;
        S3 = fmaf(qx, dx, S3);
        S4 = fmaf(qx, dy, S4);
        S5 = fmaf(qy, dy, S5);
        S7 = fmaf(w, dLp, S7);
      }
      RING_SWITCH(u + 1u);   // (no lane matches after the last step: lane 0 switches after the epoch boundary)
    }
#undef RING_SWITCH
  }
  // the last warp to leave re-arms the two counters for the next launch over the same forward state
  if (lane == 0) {
    __threadfence();
    if (atomicAdd(work + 1, 1u) == gridDim.x * RING_WARPS - 1u) { work[0] = 0u; work[1] = 0u; }
  }
}

}  // namespace cg
