/*
 * curvegs.h — C ABI of libcurvegs.so, the B200 (sm_100a) curve-Gaussian hot path.
 *
 * Every entry point takes plain device pointers + sizes + a CUDA stream handle
 * (cudaStream_t passed as void*), owns no memory, never throws, and returns
 * 0 on success or a negative CG_ERR_* code (cg_last_error() gives the text).
 * The caller (PyTorch host code, or any FFI) allocates all outputs and the
 * opaque state buffers, exactly as the reference's pybind layer hands
 * torch-owned byte tensors to CudaRasterizer through resize callbacks
 * (reference: submodules/diff-cur-rasterization/rasterize_points.cu:27-33,
 * cuda_rasterizer/rasterizer.h:24-98).
 *
 * Reference interfaces replaced (file:line under the reference checkout):
 *   cg_raster_*        -> _C.rasterize_gaussians / _C.rasterize_gaussians_backward
 *                         (rasterize_points.cu:35-130, :133-240; rasterizer_impl.cu:198-466)
 *   cg_mark_visible    -> _C.mark_visible (rasterize_points.cu:242-260)
 *   cg_sample_*        -> GaussianCurveModel.prepare_scaling_rot + rot_to_quat_batch
 *                         (scene/gaussian_curve_model.py:70-89,180-198; utils/general_utils.py:9-86)
 *   cg_ssim_*          -> fused_ssim_cuda.fusedssim / fusedssim_backward (fused-ssim/ssim.cu:368-444)
 *   cg_knn_mean_dist2  -> simple_knn._C.distCUDA2 (simple-knn/spatial.cu:15-26, simple_knn.cu:186-222)
 *   cg_edge_ssim_loss_*, cg_curve_smooth_*, cg_endpoint_conn_* -> the caller's per-iteration loss terms
 *                         (train.py:101-107, :119-124, :133-146), optional fused forms
 *
 * All float tensors are fp32, contiguous row-major; "absent" tensors are NULL.
 */
#ifndef CURVEGS_H_
#define CURVEGS_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CG_OK 0
#define CG_ERR_ARG (-1)      /* bad argument (null / shape / alignment) */
#define CG_ERR_CUDA (-2)     /* a CUDA runtime call or launch failed */
#define CG_ERR_CAPACITY (-3) /* a caller-provided buffer is too small */

/* ABI version, bumped on any signature change. */
int cg_abi_version(void);
/* Thread-local text of the last error returned on this thread. */
const char* cg_last_error(void);

/* Instrumentation (used by bench.py): number of kernels this library has launched
 * since load, and optional per-stage device timing with CUDA events recorded on the
 * launch stream. cg_profile_read synchronizes on the recorded events. */
uint64_t cg_launch_count(void);
/* Programmatic dependent launch of the library's kernels (on by default); a debugging switch. */
void cg_set_pdl(int on);
void cg_profile_enable(int on);
void cg_profile_reset(void);
int cg_profile_stage_count(void);
const char* cg_profile_stage_name(int stage);
int cg_profile_read(int stage, double* total_ms, uint64_t* calls);

/* ------------------------------------------------------------------ */
/* Rasterizer (reference: diff_cur_rasterization._C)                   */
/* ------------------------------------------------------------------ */

/* Camera / raster settings; mirrors GaussianRasterizationSettings
 * (diff_cur_rasterization/__init__.py:153-167). Matrices are DEVICE pointers
 * to 16 floats, row-major torch layout used with row vectors
 * (auxiliary.h:70-89). bg is a DEVICE pointer; only bg[0] is read
 * (NUM_CHANNELS == 1, config.h:15). */
typedef struct cg_raster_settings {
  int32_t image_height;
  int32_t image_width;
  float tanfovx;
  float tanfovy;
  float scale_modifier;
  int32_t render_geo;   /* blend the 4 all_map channels (config.h NUM_ALL_MAP) */
  int32_t debug;        /* synchronize + check after every stage (auxiliary.h:178-185) */
  int32_t antialiasing; /* opacity compensation for the 0.3 px dilation (forward.cu:219-227) */
  const float* bg;
  const float* viewmatrix;
  const float* projmatrix;
} cg_raster_settings;

/* Sizes of the opaque state buffers the caller must allocate (bytes).
 * geom: per-Gaussian state (one 48-byte record per Gaussian that the blend kernels gather by index, depth,
 * tile counts / rects, and the depth-sort buffers of the P Gaussians), img: per-pixel / per-tile / per-4x4-block
 * state, both saved for backward. bin_keep: the point list (sorted Gaussian index per tile-instance) and the
 * per-block contributor lists the forward leaves for the ring backward, saved for backward.
 * bin_scratch: working buffers of the tile binning (sort buffers, per-chunk tile counts), only live during
 * forward. */
size_t cg_raster_geom_bytes(int64_t P);
size_t cg_raster_img_bytes(int32_t W, int32_t H);
size_t cg_raster_bin_keep_bytes(int64_t R);
size_t cg_raster_bin_scratch_bytes(int64_t P, int64_t R);

/* Forward, stage 1: per-Gaussian EWA projection + tile counts + prefix scan, then the
 * depth sort of the Gaussians and their offsets in depth order (both independent of R).
 * Writes radii[P] (int32) and the geom state; returns the number of tile-instances R
 * through *num_rendered (host int). Like rasterizer_impl.cu:287 this call waits for that
 * 4-byte read, but on an event recorded right behind the copy, so the depth sort keeps
 * the GPU busy while the host wakes up and allocates the R-sized buffers.
 * It also writes the per-Gaussian record the blend kernels later gather by Gaussian index (conic, 1/depth,
 * mean2D, opacity, colour, all_map), which is why the colour and all_map arrays arrive here.
 * Exactly one of (scales+rotations) / cov3D_precomp must be non-NULL. */
int cg_raster_fwd_geom(const cg_raster_settings* s, int64_t P,
                       const float* means3D,      /* (P,3) */
                       const float* opacities,    /* (P) */
                       const float* scales,       /* (P,3) or NULL */
                       const float* rotations,    /* (P,4) raw quaternion, or NULL */
                       const float* cov3D_precomp,/* (P,6) or NULL */
                       const float* colors,       /* (P) single colour channel (colors_precomp) */
                       const float* all_map,      /* (P,4), 16-byte aligned; NULL unless render_geo */
                       int32_t* radii,            /* out (P) */
                       void* geom, size_t geom_bytes,
                       int64_t* num_rendered,     /* out, host */
                       void* stream);

/* Forward, stage 2: depth-sort the Gaussians -> duplicate per tile -> stable radix
 * sort by tile (same permutation as the reference's 64-bit tile|depth sort) -> tile
 * ranges -> per-tile front-to-back blend (records gathered by Gaussian index from the geom state).
 * Outputs are (1,H,W), (1,H,W), (4,H,W). */
int cg_raster_fwd_blend(const cg_raster_settings* s, int64_t P, int64_t R,
                        void* geom, void* img,
                        void* bin_keep, void* bin_scratch,
                        float* out_color, float* out_invdepth, float* out_all_map,
                        void* stream);

/* Forward without the host round trip (both stages in one call; CUDA-graph capturable: no
 * host synchronisation, no host allocation). The reference reads the number of tile-instances R
 * back to size its binning buffers (rasterizer_impl.cu:283-291); here the caller sizes bin_keep and
 * bin_scratch for a CAPACITY of R_cap instances (cg_raster_bin_keep_bytes(R_cap),
 * cg_raster_bin_scratch_bytes(P, R_cap)) and the true R stays on the device: every R-sized
 * kernel is launched for the capacity and reads the count itself. num_rendered_dev (DEVICE,
 * 2 x uint32) receives {R, overflow}; with overflow == 1 (R > R_cap) the farthest R - R_cap
 * instances were dropped and the outputs are not the reference's: the caller checks the
 * counter once the stream has passed this call (it may copy it to pinned memory behind the
 * call) and repeats the step with a larger capacity. When R <= R_cap every output, the saved
 * state and the later cg_raster_bwd (called with R = R_cap) are bit-identical to the
 * cg_raster_fwd_geom + cg_raster_fwd_blend pair. */
int cg_raster_fwd_capacity(const cg_raster_settings* s, int64_t P, int64_t R_cap,
                           const float* means3D, const float* opacities, const float* scales,
                           const float* rotations, const float* cov3D_precomp,
                           const float* colors, const float* all_map,
                           int32_t* radii, void* geom, size_t geom_bytes, void* img,
                           void* bin_keep, void* bin_scratch,
                           float* out_color, float* out_invdepth, float* out_all_map,
                           uint32_t* num_rendered_dev, void* stream);

/* Backward. dL_dinvdepth / dL_dall_map may be NULL (treated as zeros, which is
 * what autograd materialises for unused outputs). Gradient outputs follow
 * rasterize_points.cu:173-193: dL_dmeans2D (P,3), dL_dcolors (P,1),
 * dL_dopacity (P,1), dL_dmeans3D (P,3), dL_dcov3D (P,6), dL_dscales (P,3),
 * dL_drotations (P,4), dL_dall_map (P,4). The caller need not zero them.
 * dL_dcov3D and dL_dall_map_in may be NULL when the caller has no use for them
 * (they are then not written). grad_scratch must hold cg_raster_bwd_scratch_bytes(P). */
size_t cg_raster_bwd_scratch_bytes(int64_t P);
int cg_raster_bwd(const cg_raster_settings* s, int64_t P, int64_t R,
                  const float* means3D, const float* opacities, const float* scales,
                  const float* rotations, const float* cov3D_precomp, const int32_t* radii,
                  const void* geom, const void* img, const void* bin_keep,
                  const float* dL_dcolor, const float* dL_dinvdepth, const float* dL_dall_map,
                  void* grad_scratch,
                  float* dL_dmeans2D, float* dL_dcolors, float* dL_dopacity,
                  float* dL_dmeans3D, float* dL_dcov3D, float* dL_dscales,
                  float* dL_drotations, float* dL_dall_map_in,
                  void* stream);

/* present[i] = (view-space z > 0.2)  (rasterizer_impl.cu:54-66). */
int cg_mark_visible(int64_t P, const float* means3D, const float* viewmatrix,
                    const float* projmatrix, uint8_t* present, void* stream);

/* Introspection for parity tests: copies of internal state as typed arrays.
 * which: 0 = sorted keys (uint64, R)   1 = sorted point list (uint32, R)
 *        2 = tile ranges (uint32 pairs, tiles)  3 = tiles_touched (uint32, P)
 *        4 = means2D (float2, P)  5 = depths (float, P)  6 = conic_opacity (float4, P)
 *        7 = n_contrib (uint32, W*H)  8 = final_T (float, W*H)
 *        9 = contributor count of every 4x4 pixel block (uint32, tiles*16; block = tile*16 + 8x4 block*2 + half)
 *        10 / 11 = contributor lists of the 4x4 blocks: tile-relative list positions / Gaussian indices
 *                  (uint32, 16*R; block b of a tile with range [x, y) starts at 16*x + b*(y-x))
 * Keys (0) are only valid between cg_raster_fwd_blend and the next call that
 * reuses bin_scratch. dst is a DEVICE pointer with room for the array. */
int cg_raster_debug_fetch(int which, int64_t P, int64_t R, int32_t W, int32_t H,
                          const void* geom, const void* img, const void* bin_keep,
                          const void* bin_scratch, void* dst, void* stream);

/* ------------------------------------------------------------------ */
/* Curve -> Gaussian sampling (reference: scene/gaussian_curve_model.py)*/
/* ------------------------------------------------------------------ */

/* Forward of prepare_scaling_rot (:180-198). curve_points (B,4,3), width (B),
 * is_bezier (B) bytes, t (n) sample parameters (torch.linspace values, :58-60),
 * half_step = 0.5/n. Outputs, Gaussian index g = b*n + m:
 *   xyz (P,3), rotation (P,4) un-normalised w>=0 quaternion, scaling (P,3).
 * norms (2 floats, device) receives the two whole-tensor Frobenius norms
 * (:190,:192) and must be kept for the backward. */
/* scratch: cg_sample_scratch_bytes(B, n) for the backward (global sums + a per-Gaussian
 * structure-of-arrays work area); the forward only uses its first 64 bytes. 8-byte aligned. */
size_t cg_sample_scratch_bytes(int64_t B, int32_t n);
int cg_sample_fwd(int64_t B, int32_t n, const float* curve_points, const float* width,
                  const uint8_t* is_bezier, const float* t, float half_step,
                  float* xyz, float* rotation, float* scaling,
                  float* norms, void* scratch, void* stream);

/* Adjoint: (dL_dxyz, dL_drotation, dL_dscaling) -> dL_dcurve_points (B,4,3),
 * dL_dwidth (B). Any of the three upstream grads may be NULL (zeros). */
int cg_sample_bwd(int64_t B, int32_t n, const float* curve_points, const float* width,
                  const uint8_t* is_bezier, const float* t, float half_step,
                  const float* norms,
                  const float* dL_dxyz, const float* dL_drotation, const float* dL_dscaling,
                  float* dL_dcurve_points, float* dL_dwidth,
                  void* scratch,
                  int32_t accumulate,   /* != 0: ADD into dL_dcurve_points / dL_dwidth (the caller's gradient
                                           buffers) instead of overwriting them */
                  void* stream);

/* Per-view activations render() applies before rasterizing
 * (gaussian_renderer/__init__.py:57-104, scene/gaussian_curve_model.py:99-122):
 *   rot_n = normalize(rotation), opacity = sigmoid(opacity_logit[b]) (x mask),
 *   scales = scaling (x mask), mask = straight-through 1[sigmoid(mask_logit) > thr]
 *   (mask_logit NULL = use_mask False), all_map = (view-space main axis facing the
 *   camera, 1). opacity_logit is (B), mask_logit (B*n); outputs are per Gaussian.
 * The backward returns gradients for rotation, scaling, opacity_logit (B) and
 * mask_logit (B*n); xyz only decides a sign and receives none. Upstream grads may be NULL. */
int cg_activate_fwd(int64_t B, int32_t n, const float* xyz, const float* rotation, const float* scaling,
                    const float* opacity_logit, const float* mask_logit, float mask_thr,
                    const float* campos, const float* viewmatrix,
                    float* rot_n, float* opacity, float* scales, float* all_map, void* stream);
int cg_activate_bwd(int64_t B, int32_t n, const float* xyz, const float* rotation, const float* scaling,
                    const float* opacity_logit, const float* mask_logit, float mask_thr,
                    const float* campos, const float* viewmatrix,
                    const float* g_rot_n, const float* g_opacity, const float* g_scales, const float* g_all_map,
                    float* g_rotation, float* g_scaling, float* g_opacity_logit, float* g_mask_logit,
                    int32_t accumulate,   /* != 0: ADD into g_opacity_logit / g_mask_logit instead of overwriting */
                    void* stream);

/* ------------------------------------------------------------------ */
/* Fused SSIM (reference: submodules/fused-ssim)                        */
/* ------------------------------------------------------------------ */

/* img1/img2 (B,CH,H,W). ssim_map always written; the three partial maps only
 * when non-NULL (train=True). */
int cg_ssim_fwd(int32_t B, int32_t CH, int32_t H, int32_t W, float C1, float C2,
                const float* img1, const float* img2, float* ssim_map,
                float* dm_dmu1, float* dm_dsigma1_sq, float* dm_dsigma12, void* stream);
int cg_ssim_bwd(int32_t B, int32_t CH, int32_t H, int32_t W, float C1, float C2,
                const float* img1, const float* img2, const float* dL_dmap,
                const float* dm_dmu1, const float* dm_dsigma1_sq, const float* dm_dsigma12,
                float* dL_dimg1, void* stream);

/* ------------------------------------------------------------------ */
/* Fused image loss of the training step (reference caller: train.py)   */
/* ------------------------------------------------------------------ */

/* loss = lambda_mse * ((1 - lambda_dssim) * edge_aware_loss(img, gt, threshold)
 *                      + lambda_dssim * (1 - mean(ssim_map(img, gt))))
 * for single-channel (H,W) images: train.py:101-107 with utils/loss_utils.py:94-115
 * (class-balanced weighted MSE) and fused_ssim (fused-ssim/ssim.cu:187-366) in ONE
 * forward and ONE backward kernel. loss_out is a DEVICE scalar (no host sync). stats
 * is cg_edge_ssim_loss_stats_bytes() of device memory kept for the backward together
 * with the three partial maps. g_loss is the DEVICE upstream scalar (NULL = 1).
 * clamp01 != 0: img is the RAW render and the clamp(0,1) render() applies
 * (gaussian_renderer/__init__.py:139) is fused in: values are clamped on load and the
 * gradient is zeroed outside [0,1], exactly like torch.clamp's adjoint. */
size_t cg_edge_ssim_loss_stats_bytes(void);
int cg_edge_ssim_loss_fwd(int32_t H, int32_t W, const float* img, const float* gt,
                          float threshold, float lambda_mse, float lambda_dssim, float C1, float C2,
                          int32_t clamp01, void* stats, float* loss_out,
                          float* dm_dmu1, float* dm_dsigma1_sq, float* dm_dsigma12, void* stream);
int cg_edge_ssim_loss_bwd(int32_t H, int32_t W, const float* img, const float* gt,
                          float threshold, float lambda_mse, float lambda_dssim, int32_t clamp01,
                          const void* stats, const float* g_loss,
                          const float* dm_dmu1, const float* dm_dsigma1_sq, const float* dm_dsigma12,
                          float* dL_dimg, void* stream);

/* out[c][i] = sum_k in[k][i] * M[c][k] over n pixels of a (3,n) planar image; M is a
 * DEVICE 3x3 with row stride ld floats (transpose != 0 applies M^T). This is the
 * view->world rotation of the direction channels in render()
 * (gaussian_renderer/__init__.py:144), a (H*W,3)x(3,3) GEMM in the reference. */
int cg_rotate_channels(int64_t n, const float* in, const float* m3x3, int32_t ld, int32_t transpose,
                       float* out, void* stream);

/* ------------------------------------------------------------------ */
/* Curve-side regularisers of the training step (reference caller: train.py) */
/* ------------------------------------------------------------------ */

/* Curve smoothness (train.py:119-124): mean over the B*(n-1) adjacent sample pairs of
 * 1 - |cos(dir[b,m], dir[b,m+1])|, dir = first column of
 * quaternion_to_matrix(normalize(rotation)) (gaussian_curve_model.py:95-97,120-122).
 * rotation is the RAW (P,4) quaternion of the curve model; the backward returns the
 * gradient with respect to it. loss_out / g_loss are DEVICE scalars (g_loss NULL = 1).
 * scratch: cg_curve_smooth_scratch_bytes() bytes, 8-byte aligned. */
size_t cg_curve_smooth_scratch_bytes(void);
int cg_curve_smooth_fwd(int64_t B, int32_t n, const float* rotation, void* scratch, float* loss_out, void* stream);
int cg_curve_smooth_bwd(int64_t B, int32_t n, const float* rotation, const float* g_loss, float* g_rotation,
                        void* stream);

/* Endpoint connectivity (train.py:133-146): over the 2B curve endpoints (first and last
 * control point of every curve), the mean Euclidean distance of all ordered pairs closer
 * than dis_thr, excluding pairs of endpoints of the same curve; 0 when there is none (the
 * reference then skips the term). The reference builds torch.cdist's dense (2B)^2 matrix;
 * here the pairs are streamed. curve_points is (B,4,3). The forward also writes v (2B,3),
 * the per-endpoint sum of unit difference vectors, which the backward scales into
 * g_curve_points (B,4,3) (rows 1 and 2 are written as zeros). scratch:
 * cg_endpoint_conn_scratch_bytes(B), 8-byte aligned, kept for the backward. */
size_t cg_endpoint_conn_scratch_bytes(int64_t B);
int cg_endpoint_conn_fwd(int64_t B, const float* curve_points, float dis_thr, void* scratch, float* v,
                         float* loss_out, void* stream);
int cg_endpoint_conn_bwd(int64_t B, const float* v, const void* scratch, const float* g_loss,
                         float* g_curve_points, void* stream);

/* ------------------------------------------------------------------ */
/* simple-knn (reference: submodules/simple-knn)                        */
/* ------------------------------------------------------------------ */

/* mean_dist2[i] = mean of the 3 smallest squared distances to other points. */
size_t cg_knn_scratch_bytes(int64_t P);
int cg_knn_mean_dist2(int64_t P, const float* points, float* mean_dist2,
                      void* scratch, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CURVEGS_H_ */
