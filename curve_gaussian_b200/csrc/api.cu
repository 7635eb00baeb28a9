// extern "C" entry points of libcurvegs.so (see include/curvegs.h).
#include <stdarg.h>
#include <string.h>
#include <atomic>
#include <mutex>
#include <vector>
#include "common.cuh"

namespace cg {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int g_use_pdl = 1;

static std::atomic<uint64_t> g_launches{0};
static std::atomic<int> g_profile{0};
struct ProfRec { int stage; cudaEvent_t e0, e1; };
static std::mutex g_prof_mu;
static std::vector<ProfRec> g_pending;
static std::vector<cudaEvent_t> g_free_events;
static double g_stage_ms[ST_COUNT];
static uint64_t g_stage_calls[ST_COUNT];

void count_launches(int n) { g_launches.fetch_add(uint64_t(n), std::memory_order_relaxed); }

static cudaEvent_t get_event() {
  if (!g_free_events.empty()) { cudaEvent_t e = g_free_events.back(); g_free_events.pop_back(); return e; }
  cudaEvent_t e = nullptr;
  cudaEventCreate(&e);
  return e;
}
StageTimer::StageTimer(int stage_, cudaStream_t st_, int kernels) : stage(stage_), st(st_), rec(nullptr) {
  count_launches(kernels);
  if (!g_profile.load(std::memory_order_relaxed)) return;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  ProfRec* r = new ProfRec{stage, get_event(), get_event()};
  cudaEventRecord(r->e0, st);
  rec = r;
}
StageTimer::~StageTimer() {
  if (!rec) return;
  ProfRec* r = reinterpret_cast<ProfRec*>(rec);
  cudaEventRecord(r->e1, st);
  std::lock_guard<std::mutex> lk(g_prof_mu);
  g_pending.push_back(*r);
  delete r;
}
static void drain_profile() {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  for (auto& r : g_pending) {
    cudaEventSynchronize(r.e1);
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, r.e0, r.e1) == cudaSuccess) { g_stage_ms[r.stage] += ms; g_stage_calls[r.stage] += 1; }
    g_free_events.push_back(r.e0);
    g_free_events.push_back(r.e1);
  }
  g_pending.clear();
}
static const char* kStageNames[ST_COUNT] = {"sample_fwd", "preprocess_fwd", "scan", "emit_keys", "radix_sort",
                                            "tile_ranges", "gather_records(unused)", "blend_fwd", "blend_bwd",
                                            "preprocess_bwd", "sample_bwd", "ssim_fwd", "ssim_bwd", "knn",
                                            "activate_fwd", "activate_bwd", "loss_fwd", "loss_bwd"};

int launch_fwd_geom(const cg_raster_settings* s, int64_t P, const float* means3D, const float* opacities,
                    const float* scales, const float* rotations, const float* cov3D_precomp, const float* colors,
                    const float* all_map, int32_t* radii, void* geom, int64_t* num_rendered, cudaStream_t st);
int launch_fwd_blend(const cg_raster_settings* s, int64_t P, int64_t R, void* geom, void* img, void* bin_keep, void* bin_scratch, float* out_color, float* out_invd,
                     float* out_map, uint32_t* nr_out, cudaStream_t st);
int launch_bwd(const cg_raster_settings* s, int64_t P, int64_t R, const float* means3D, const float* opacities,
               const float* scales, const float* rotations, const float* cov3D_precomp, const int32_t* radii,
               const void* geom, const void* img, const void* bin_keep, const float* dL_dcolor,
               const float* dL_dinvdepth, const float* dL_dall_map, void* grad_scratch, float* dL_dmeans2D,
               float* dL_dcolors, float* dL_dopacity, float* dL_dmeans3D, float* dL_dcov3D, float* dL_dscales,
               float* dL_drotations, float* dL_dall_map_in, cudaStream_t st);
int launch_mark_visible(int64_t P, const float* means3D, const float* vm, uint8_t* present, cudaStream_t st);
int launch_rebuild_keys(int64_t P, int64_t R, int W, int H, const void* geom, const void* img, const void* bin_keep,
                        const void* bin_scratch, uint64_t* dst, cudaStream_t st);

static int check_settings(const cg_raster_settings* s) {
  CG_ARG(s != nullptr, "settings");
  CG_ARG(s->image_width > 0 && s->image_height > 0, "image size");
  CG_ARG(s->image_width <= 65535 * TILE_X && s->image_height <= 65535 * TILE_Y, "image too large for packed tile rects");
  CG_ARG(s->bg && s->viewmatrix && s->projmatrix, "bg/viewmatrix/projmatrix must be device pointers");
  return CG_OK;
}

}  // namespace cg

using namespace cg;

extern "C" {

int cg_abi_version(void) { return 6; }
void cg_set_pdl(int on) { g_use_pdl = on ? 1 : 0; }
uint64_t cg_launch_count(void) { return g_launches.load(); }
void cg_profile_enable(int on) { g_profile.store(on ? 1 : 0); }
void cg_profile_reset(void) {
  drain_profile();
  for (int i = 0; i < ST_COUNT; ++i) { g_stage_ms[i] = 0.0; g_stage_calls[i] = 0; }
}
int cg_profile_stage_count(void) { return ST_COUNT; }
const char* cg_profile_stage_name(int i) { return (i >= 0 && i < ST_COUNT) ? kStageNames[i] : ""; }
int cg_profile_read(int i, double* total_ms, uint64_t* calls) {
  drain_profile();
  if (i < 0 || i >= ST_COUNT || !total_ms || !calls) return CG_ERR_ARG;
  *total_ms = g_stage_ms[i];
  *calls = g_stage_calls[i];
  return CG_OK;
}
const char* cg_last_error(void) { return g_err; }

size_t cg_raster_geom_bytes(int64_t P) {
  size_t b = 0;
  GeomState::carve(nullptr, P < 0 ? 0 : P, &b);
  return b;
}
size_t cg_raster_img_bytes(int32_t W, int32_t H) {
  size_t b = 0;
  ImgState::carve(nullptr, W, H, &b);
  return b;
}
size_t cg_raster_bin_keep_bytes(int64_t R) {
  size_t b = 0;
  BinKeep::carve(nullptr, R < 0 ? 0 : R, &b);
  return b;
}
size_t cg_raster_bin_scratch_bytes(int64_t P, int64_t R) {
  size_t b = 0;
  BinScratch::carve(nullptr, P < 0 ? 0 : P, R < 0 ? 0 : R, &b);
  return b;
}
size_t cg_raster_bwd_scratch_bytes(int64_t P) { return size_t(P < 1 ? 1 : P) * 8 * sizeof(float); }

int cg_raster_fwd_geom(const cg_raster_settings* s, int64_t P, const float* means3D, const float* opacities,
                       const float* scales, const float* rotations, const float* cov3D_precomp, const float* colors,
                       const float* all_map, int32_t* radii, void* geom, size_t geom_bytes, int64_t* num_rendered,
                       void* stream) {
  int rc = check_settings(s);
  if (rc) return rc;
  CG_ARG(num_rendered != nullptr, "num_rendered");
  *num_rendered = 0;
  if (P == 0) return CG_OK;
  CG_ARG(P > 0 && P < (int64_t(1) << 31), "P");
  CG_ARG(means3D && opacities && radii && geom && colors, "means3D/opacities/colors/radii/geom");
  CG_ARG((scales && rotations && !cov3D_precomp) || (!scales && !rotations && cov3D_precomp),
         "exactly one of scale/rotation pair or precomputed 3D covariance");
  CG_ARG(!s->render_geo || all_map, "all_map required with render_geo");
  CG_ARG(!all_map || (reinterpret_cast<uintptr_t>(all_map) & 15u) == 0, "all_map must be 16-byte aligned");
  if (geom_bytes < cg_raster_geom_bytes(P)) {
    set_error("geom buffer too small: %zu < %zu", geom_bytes, cg_raster_geom_bytes(P));
    return CG_ERR_CAPACITY;
  }
  CG_ARG((reinterpret_cast<uintptr_t>(geom) & 127u) == 0, "geom must be 128-byte aligned");
  return launch_fwd_geom(s, P, means3D, opacities, scales, rotations, cov3D_precomp, colors, all_map, radii, geom,
                         num_rendered, reinterpret_cast<cudaStream_t>(stream));
}

int cg_raster_fwd_blend(const cg_raster_settings* s, int64_t P, int64_t R, void* geom, void* img, void* bin_keep,
                        void* bin_scratch, float* out_color, float* out_invdepth, float* out_all_map, void* stream) {
  int rc = check_settings(s);
  if (rc) return rc;
  CG_ARG(P >= 0 && R >= 0, "P/R");
  // (the per-block contributor lists are addressed with 32-bit offsets of up to 16 * R)
  CG_ARG(R < (int64_t(1) << 28), "num_rendered must be below 2^28 tile instances");
  CG_ARG(img && out_color && out_invdepth, "img/out_color/out_invdepth");
  CG_ARG(!s->render_geo || out_all_map, "out_all_map required with render_geo");
  CG_ARG((reinterpret_cast<uintptr_t>(img) & 127u) == 0, "img must be 128-byte aligned");
  if (R > 0) {
    CG_ARG(geom && bin_keep && bin_scratch, "geom/bin_keep/bin_scratch");
    CG_ARG((reinterpret_cast<uintptr_t>(bin_keep) & 127u) == 0 && (reinterpret_cast<uintptr_t>(bin_scratch) & 127u) == 0,
           "bin buffers must be 128-byte aligned");
  }
  return launch_fwd_blend(s, P, R, geom, img, bin_keep, bin_scratch, out_color, out_invdepth,
                          out_all_map, nullptr, reinterpret_cast<cudaStream_t>(stream));
}

int cg_raster_fwd_capacity(const cg_raster_settings* s, int64_t P, int64_t R_cap, const float* means3D,
                           const float* opacities, const float* scales, const float* rotations,
                           const float* cov3D_precomp, const float* colors, const float* all_map, int32_t* radii,
                           void* geom, size_t geom_bytes, void* img, void* bin_keep, void* bin_scratch,
                           float* out_color, float* out_invdepth, float* out_all_map, uint32_t* num_rendered_dev,
                           void* stream) {
  int rc = check_settings(s);
  if (rc) return rc;
  CG_ARG(P > 0 && P < (int64_t(1) << 31), "P");
  CG_ARG(R_cap > 0 && R_cap < (int64_t(1) << 28), "R_cap");   // (32-bit offsets of up to 16 * R into the contributor lists)
  CG_ARG(means3D && opacities && radii && geom && img && colors, "means3D/opacities/radii/geom/img/colors");
  CG_ARG((scales && rotations && !cov3D_precomp) || (!scales && !rotations && cov3D_precomp),
         "exactly one of scale/rotation pair or precomputed 3D covariance");
  CG_ARG(bin_keep && bin_scratch && out_color && out_invdepth && num_rendered_dev,
         "bin_keep/bin_scratch/out_color/out_invdepth/num_rendered_dev");
  CG_ARG(!s->render_geo || (out_all_map && all_map), "all_map/out_all_map required with render_geo");
  if (geom_bytes < cg_raster_geom_bytes(P)) {
    set_error("geom buffer too small: %zu < %zu", geom_bytes, cg_raster_geom_bytes(P));
    return CG_ERR_CAPACITY;
  }
  CG_ARG(((reinterpret_cast<uintptr_t>(geom) | reinterpret_cast<uintptr_t>(img) | reinterpret_cast<uintptr_t>(bin_keep) |
           reinterpret_cast<uintptr_t>(bin_scratch)) & 127u) == 0, "state buffers must be 128-byte aligned");
  CG_ARG(!all_map || (reinterpret_cast<uintptr_t>(all_map) & 15u) == 0, "all_map must be 16-byte aligned");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  rc = launch_fwd_geom(s, P, means3D, opacities, scales, rotations, cov3D_precomp, colors, all_map, radii, geom, nullptr, st);
  if (rc) return rc;
  return launch_fwd_blend(s, P, R_cap, geom, img, bin_keep, bin_scratch, out_color, out_invdepth,
                          out_all_map, num_rendered_dev, st);
}

int cg_raster_bwd(const cg_raster_settings* s, int64_t P, int64_t R, const float* means3D, const float* opacities,
                  const float* scales, const float* rotations, const float* cov3D_precomp, const int32_t* radii,
                  const void* geom, const void* img, const void* bin_keep, const float* dL_dcolor,
                  const float* dL_dinvdepth, const float* dL_dall_map, void* grad_scratch, float* dL_dmeans2D,
                  float* dL_dcolors, float* dL_dopacity, float* dL_dmeans3D, float* dL_dcov3D, float* dL_dscales,
                  float* dL_drotations, float* dL_dall_map_in, void* stream) {
  int rc = check_settings(s);
  if (rc) return rc;
  if (P == 0) return CG_OK;
  CG_ARG(P > 0 && R >= 0, "P/R");
  CG_ARG(means3D && opacities && radii && geom && img && dL_dcolor && grad_scratch, "inputs");
  CG_ARG(R == 0 || bin_keep, "bin_keep");
  CG_ARG((scales && rotations && !cov3D_precomp) || (!scales && !rotations && cov3D_precomp),
         "exactly one of scale/rotation pair or precomputed 3D covariance");
  CG_ARG(dL_dmeans2D && dL_dcolors && dL_dopacity && dL_dmeans3D, "gradient outputs");
  CG_ARG(cov3D_precomp || (dL_dscales && dL_drotations), "dL_dscales/dL_drotations");
  CG_ARG((reinterpret_cast<uintptr_t>(grad_scratch) & 31u) == 0, "grad_scratch must be 32-byte aligned");
  return launch_bwd(s, P, R, means3D, opacities, scales, rotations, cov3D_precomp, radii, geom, img, bin_keep,
                    dL_dcolor, dL_dinvdepth, dL_dall_map, grad_scratch, dL_dmeans2D, dL_dcolors, dL_dopacity,
                    dL_dmeans3D, dL_dcov3D, dL_dscales, dL_drotations, dL_dall_map_in,
                    reinterpret_cast<cudaStream_t>(stream));
}

int cg_mark_visible(int64_t P, const float* means3D, const float* viewmatrix, const float* projmatrix,
                    uint8_t* present, void* stream) {
  (void)projmatrix;
  if (P == 0) return CG_OK;
  CG_ARG(P > 0 && means3D && viewmatrix && present, "mark_visible inputs");
  return launch_mark_visible(P, means3D, viewmatrix, present, reinterpret_cast<cudaStream_t>(stream));
}

int cg_raster_debug_fetch(int which, int64_t P, int64_t R, int32_t W, int32_t H, const void* geom, const void* img,
                          const void* bin_keep, const void* bin_scratch, void* dst, void* stream) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  CG_ARG(dst != nullptr, "dst");
  const size_t tiles = size_t((W + TILE_X - 1) / TILE_X) * ((H + TILE_Y - 1) / TILE_Y);
  const void* src = nullptr;
  size_t bytes = 0;
  GeomState g = GeomState::carve(const_cast<void*>(geom), P, nullptr);
  ImgState im = ImgState::carve(const_cast<void*>(img), W, H, nullptr);
  BinKeep bk = BinKeep::carve(const_cast<void*>(bin_keep), R, nullptr);
  switch (which) {
    case 0:
      // the reference's 64-bit keys are not materialised (see BinScratch); rebuild them from the sorted state
      CG_ARG(bin_scratch != nullptr && bin_keep != nullptr && geom != nullptr, "bin_scratch/bin_keep/geom");
      return launch_rebuild_keys(P, R, W, H, geom, img, bin_keep, bin_scratch, reinterpret_cast<uint64_t*>(dst), st);
    case 1: CG_ARG(bin_keep != nullptr, "bin_keep"); src = bk.point_list; bytes = size_t(R) * 4; break;
    case 2: src = im.ranges; bytes = tiles * 8; break;
    case 3: src = g.tiles; bytes = size_t(P) * 4; break;
    case 4:   // means2D (x, y): strided out of the per-Gaussian records
      CG_CUDA(cudaMemcpy2DAsync(dst, 8, &g.grec[0].x, sizeof(Rec), 8, size_t(P), cudaMemcpyDeviceToDevice, st));
      return CG_OK;
    case 5: src = g.depth; bytes = size_t(P) * 4; break;
    case 6:   // conic_opacity (conic xx, xy, yy, opacity)
      CG_CUDA(cudaMemcpy2DAsync(dst, 16, &g.grec[0].ca, sizeof(Rec), 12, size_t(P), cudaMemcpyDeviceToDevice, st));
      CG_CUDA(cudaMemcpy2DAsync(reinterpret_cast<char*>(dst) + 12, 16, &g.grec[0].o, sizeof(Rec), 4, size_t(P), cudaMemcpyDeviceToDevice, st));
      return CG_OK;
    case 7: src = im.n_contrib; bytes = size_t(W) * H * 4; break;
    case 8: src = im.final_T; bytes = size_t(W) * H * 4; break;
    case 9: src = im.blk_cnt; bytes = tiles * 16 * 4; break;
    case 10: CG_ARG(bin_keep != nullptr, "bin_keep"); src = bk.cand; bytes = size_t(R) * 16 * 4; break;
    case 11: CG_ARG(bin_keep != nullptr, "bin_keep"); src = bk.cand_id; bytes = size_t(R) * 16 * 4; break;
    default: set_error("debug_fetch: unknown selector %d", which); return CG_ERR_ARG;
  }
  CG_ARG(src != nullptr, "state buffer");
  if (bytes) CG_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, st));
  return CG_OK;
}

}  // extern "C"
