/*
 * densepoints_cuda.h -- C ABI of the B200 (sm_100a) photometric hot path of
 * DensePoints' PMVS method.  This is the drop-in boundary: plain pointers and
 * sizes, no C++/torch types, int return codes (0 = ok, negative = dp_status),
 * never aborts, never throws.  The reference has no FFI today; each
 * entry point names the reference interface (file:line under the reference
 * root) whose per-patch loop it replaces with one batched call.  INTEGRATION.md
 * shows the reference-side binding (class OptimizationCUDA : public Optimization).
 *
 * Two layers:
 *   dp_*      host buffers in / out (what methods/pmvs would call);
 *             H2D + kernels + D2H, synchronous on return.
 *   dp_*_dev  the same launches on caller-owned DEVICE buffers and a caller
 *             stream (cudaStream_t passed as void*), asynchronous.
 *
 * A context is bound to one CUDA device and is not re-entrant (one host thread
 * drives one handle), matching SURVEY 8b "Threading".
 */
#ifndef DENSEPOINTS_CUDA_H
#define DENSEPOINTS_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DP_ABI_VERSION 1
#define DP_MAX_CELL_SIZE 32 /* cell_size (the reference's mu) supported: 2..32 */

typedef enum dp_status {
  DP_OK = 0,
  DP_ERR_INVALID_ARG = -1,
  DP_ERR_CUDA = -2,
  DP_ERR_NO_DEVICE = -3,
  DP_ERR_OOM = -4,
  DP_ERR_STATE = -5
} dp_status;

typedef struct dp_context dp_context;

/* The PMVS constants, exactly those the reference hard-codes or takes as ctor
 * defaults (SURVEY section 5 "Config"). dp_default_params() = reference defaults. */
typedef struct dp_params {
  double score_threshold;     /* Optimization ctor, optimization.h:16      (0.6)  */
  int32_t minimum_visible_image; /* Optimization ctor, optimization.h:17   (3)    */
  double visible_threshold;   /* Patch::InitRelatedImages, patch.h:56      (0.78) */
  double candidate_threshold; /* Patch::InitRelatedImages, patch.h:57      (1.04) */
  int32_t grid_scale;         /* PatchOrganizerOptions, patch_organizer.h:43 (8)  */
  int32_t max_patches_per_cell; /* PatchOrganizerOptions, patch_organizer.h:42 (1); 1..255 */
  double nm_step[3];          /* OptimizationOpenCV::Optimize, optimization_opencv.cpp:56 (0.02,0.2,0.2) */
  int32_t nm_max_evals;       /* TermCriteria maxCount, optimization_opencv.cpp:60 (500) */
  double nm_eps;              /* TermCriteria epsilon,  optimization_opencv.cpp:60 (1e-4) */
  int64_t max_pops;           /* Expand::ExpandPatches cap, expand.cpp:95 (1e7) */
} dp_params;

/* A batch of patches, structure of arrays (Patch + pcl::PointXYZRGBNormal,
 * patch.h:21-101, core/types.h:27: position/normal are stored as fp32). */
typedef struct dp_patch_soa {
  int32_t n;       /* patches */
  int32_t vstride; /* row length of vis (>= max nvis) */
  float *pos;      /* n*3  Patch::GetPosition */
  float *nrm;      /* n*3  Patch::GetNormal */
  int32_t *ref;    /* n    Patch::GetReferenceImage */
  int32_t *nvis;   /* n    GetTrullyVisibleImages().size() */
  int32_t *vis;    /* n*vstride  GetTrullyVisibleImages(), ascending view id, -1 padded */
  uint8_t *rgb;    /* n*3  r,g,b (Patch::ComputeColor); may be NULL */
} dp_patch_soa;

/* ---- context ------------------------------------------------------------- */
void dp_default_params(dp_params *p);
int dp_abi_version(void);
/* device < 0: current device.  params == NULL: defaults. */
int dp_create(dp_context **ctx, int device, const dp_params *params);
void dp_destroy(dp_context *ctx);
const char *dp_last_error(const dp_context *ctx);
int dp_set_params(dp_context *ctx, const dp_params *params);
int dp_get_params(const dp_context *ctx, dp_params *params);
int dp_sync(dp_context *ctx);
/* number of kernels this context has launched so far (bench.py's gpu_launches) */
int64_t dp_launch_count(const dp_context *ctx);

/* ---- views: replaces View::Load + View::SetProjectionMatrix -----------------
 * (core/types.cpp:7-11, 28-68; called from PMVS::AddCamera, pmvs.cpp:11-20).
 * The image (BGR u8, cv::imread layout, `stride` bytes per row) is uploaded once
 * and kept resident as packed BGRx.  xaxis/center may be NULL (the library then
 * decomposes P itself) or the host's own View::GetXAxis()/GetCameraCenter() to
 * stay bit-identical with the caller's Eigen result. */
int dp_set_num_views(dp_context *ctx, int n_views);
int dp_upload_view(dp_context *ctx, int view_id, const double P[12], const double *xaxis,
                   const double *center, const uint8_t *bgr, int width, int height, size_t stride);
int dp_num_views(const dp_context *ctx);
/* reads back the library's View state (for tests of the decomposition) */
int dp_get_view(const dp_context *ctx, int view_id, double xaxis[3], double center[3], int *width,
                int *height);

/* ---- scoring: Optimization::GetProjectedTextures + NCCScore ------------------
 * (optimization.cpp:14-56, core/error_measurements.cpp:36-60), the loop of
 * FilterByErrorMeasurement (optimization.cpp:104-110) for every patch.
 * ncc [n*vstride]: ncc[i*vstride+k] = NCCScore(texture 0, texture k), k = 1..nvis-1
 *                  (-1 for an empty texture); slot 0 and slots >= nvis are 0.
 * tex (optional) [n*vstride*s*s*3]: the s x s BGR textures; valid (optional)
 * [n*vstride]: 0 where the reference would push an empty cv::Mat. */
int dp_score(dp_context *ctx, const dp_patch_soa *patches, int cell_size, float *ncc,
             uint8_t *tex, uint8_t *valid);

/* ---- Optimization::GetProjectedTextures(normal, position, textures) (optimization.cpp:14-56),
 * the public two-argument overload the refinement objective calls
 * (optimization_opencv.cpp:17-21), + the NCC loop: `normal` / `position` (fp64, n*3 each; NULL =
 * the patch's own) only feed GetProjectedXYAxisAndScale (axes and dx, optimization.cpp:24-26);
 * the four corners stay centred on the patch's STORED position (patch.cpp:119-123).
 * Outputs as dp_score. */
int dp_score_at(dp_context *ctx, const dp_patch_soa *patches, int cell_size, const double *normal,
                const double *position, float *ncc, uint8_t *tex, uint8_t *valid);

/* ---- filter: Optimization::FilterByErrorMeasurement (optimization.cpp:98-132)
 * for every patch = body of Seed::FilterPatches (seed.cpp:110-126).
 * nvis/vis are edited in place exactly as the reference erases entries
 * (including its score/index off-by-one); keep[i] = the bool it returns. */
int dp_filter(dp_context *ctx, dp_patch_soa *patches, int cell_size, uint8_t *keep);

/* ---- refinement: OptimizationOpenCV::Optimize (optimization_opencv.cpp:44-78)
 * for every patch = body of Seed::OptimizePatches (seed.cpp:128-144).
 * pos/nrm are updated in place (fp32).  mask (optional) [n]: only patches with
 * mask[i] != 0 are refined (= the survivors of Seed::RemovePatches, seed.cpp:146-156).
 * evals (optional) [n] = function evaluations used (0 where masked out);
 * xbest (optional) [n*3] = (depth, roll, pitch) found. */
int dp_refine(dp_context *ctx, dp_patch_soa *patches, int cell_size, const uint8_t *mask,
              int32_t *evals, double *xbest);

/* ---- Seed::OptimizeAndRefinePatches (seed.cpp:88-108) in one call: FilterPatches, then
 * OptimizePatches on the survivors, with a single upload and a single download.
 * On return nvis/vis are filtered as by dp_filter, keep[i] says whether patch i survived
 * (the caller erases the others, Seed::RemovePatches), pos/nrm of the survivors are refined
 * as by dp_refine (the others are unchanged), evals (optional) as in dp_refine. */
int dp_filter_refine(dp_context *ctx, dp_patch_soa *patches, int cell_size, uint8_t *keep,
                     int32_t *evals);

/* ---- visibility: Patch::InitRelatedImages (patch.cpp:19-49) for every patch.
 * Writes nvis/vis of `patches` (visible) and, if non-NULL, ncand/cand
 * (potentially visible; cand is n*vstride). */
int dp_visibility(dp_context *ctx, dp_patch_soa *patches, int32_t *ncand, int32_t *cand);

/* ---- colour: Patch::ComputeColor (patch.cpp:51-73) for every patch -> patches->rgb */
int dp_color(dp_context *ctx, dp_patch_soa *patches);

/* ---- organizer + expansion ----------------------------------------------------
 * PatchOrganizer (patch_organizer.cpp:32-75) and Expand (expand.cpp:13-143) with
 * the patch store and the per-view occupancy grids resident on the device.
 * Order is the reference's single-thread FIFO order. */
int dp_organizer_reset(dp_context *ctx); /* AllocateViews, patch_organizer.cpp:32-40 */
/* SetSeeds (patch_organizer.cpp:70-75): TryInsert each patch in order;
 * accepted (optional) [n] = 1 where TryInsert returned non-null.  The store keeps a visible
 * set as a bit mask over the views, so the ids of every patch must be valid and strictly
 * ascending -- the order Patch::InitRelatedImages produces (patch.cpp:29-47) and
 * FilterByErrorMeasurement keeps; anything else is DP_ERR_INVALID_ARG. */
int dp_organizer_insert(dp_context *ctx, const dp_patch_soa *patches, uint8_t *accepted);
int64_t dp_organizer_size(const dp_context *ctx);
/* copies the store out; out->n / out->vstride give the capacity of the arrays */
int dp_organizer_export(dp_context *ctx, dp_patch_soa *out);
/* occupancy counts of one view's grid, row-major gh x gw (capacity bytes) */
int dp_organizer_grid(dp_context *ctx, int view_id, uint8_t *out, size_t capacity, int *gw,
                      int *gh);
/* all grids concatenated in view order; out == NULL only reports *n_cells */
int dp_organizer_grids(dp_context *ctx, uint8_t *out, size_t capacity, int64_t *n_cells);
/* Expand::ExpandPatches (expand.cpp:34-101).  max_levels < 0: until the queue is
 * empty (reference behaviour).  stats (optional) [4]: pops, candidates refined,
 * candidates that passed the filter, patches inserted. */
int dp_expand(dp_context *ctx, int cell_size, int max_levels, int64_t *stats);

/* ---- multi-GPU expansion (patches sharded by reference image) --------
 * One BFS level in three steps; the caller (one process per GPU) allgathers the
 * candidate records between steps 1 and 2 with NCCL.
 *  1. dp_expand_level_local : refine + re-derive visibility + filter the children of
 *     the frontier parents this rank owns -- owner = rank_of_view[ref] when a table is given
 *     (n_views entries), or, with rank_of_view == NULL, every world-th of 8 * world contiguous
 *     pieces of the frontier with equal work (sum of visible-view counts), the same cut on
 *     every rank because the store is replicated; survivors are
 *     written as fixed-size records straight into the send buffer `records_dev`
 *     (device pointer, capacity max_records), ascending sequence id;
 *     *n_records = how many.  Returns the record size through dp_record_bytes().
 *  2. (caller) allgather counts + records.
 *  3. dp_expand_level_commit : every rank replays TryInsert over all gathered records in
 *     sequence order -> identical grids and stores on every rank. */
/* A record is 32 bytes (seq, ref | nvis << 16, pos, nrm as fp32 bits) + the visible set as a
 * bit mask of ceil(n_views / 32) words: 40 bytes at 64 views, 64 bytes at 256 views. */
size_t dp_record_bytes(const dp_context *ctx);
int dp_expand_frontier(dp_context *ctx, int64_t *begin, int64_t *end);
/* weights [n_views] (host): per reference view, the sum of the visible-view counts of the
 * frontier parents that will expand = the work of owning that view in this level (used to
 * balance rank_of_view level by level; every rank computes the same numbers). */
int dp_expand_frontier_weights(dp_context *ctx, int64_t *weights);
int dp_expand_level_local(dp_context *ctx, int cell_size, int rank, int world,
                          const int32_t *rank_of_view, void *records_dev, int64_t max_records,
                          int64_t *n_records, void *stream);
/* candidates this rank refined in its last dp_expand_level_local call (bookkeeping) */
int64_t dp_expand_last_candidates(const dp_context *ctx);
int dp_expand_level_commit(dp_context *ctx, const void *records_dev, int64_t n_records,
                           int64_t *n_inserted, void *stream);
/* step 3 straight on the output of an allgather of padded buffers: `world` segments of
 * segment_capacity records, the first counts[r] (host array) of segment r valid. */
int dp_expand_level_commit_gathered(dp_context *ctx, const void *gathered_dev, int world,
                                    int64_t segment_capacity, const int64_t *counts,
                                    int64_t *n_inserted, void *stream);

/* ---- device-pointer layer (async on `stream`; all pointers are device memory).
 * The calls of one context share internal scratch (work counter, order table, expansion
 * buffers); calls issued on different streams are chained with events inside the library, so
 * they run one after the other, never concurrently.  Use one context per stream for overlap. */
typedef struct dp_patch_dev {
  int32_t n, vstride;
  float *pos, *nrm;
  int32_t *ref, *nvis, *vis;
  uint8_t *rgb;
} dp_patch_dev;
int dp_score_dev(dp_context *ctx, const dp_patch_dev *p, int cell_size, float *ncc, uint8_t *tex,
                 uint8_t *valid, void *stream);
int dp_score_at_dev(dp_context *ctx, const dp_patch_dev *p, int cell_size, const double *normal,
                    const double *position, float *ncc, uint8_t *tex, uint8_t *valid, void *stream);
int dp_filter_dev(dp_context *ctx, dp_patch_dev *p, int cell_size, uint8_t *keep, void *stream);
/* dp_refine_dev with cell_size > 8 and more than 4 patches per SM runs time-sliced (several
 * launches with an evaluation budget, the Nelder-Mead state of unfinished patches carried over
 * bit for bit): it reads one counter between launches, i.e. it synchronises `stream` and returns
 * with at most the last launch in flight.  Results do not depend on the slicing. */
int dp_refine_dev(dp_context *ctx, dp_patch_dev *p, int cell_size, const uint8_t *mask,
                  int32_t *evals, double *xbest, void *stream);
int dp_visibility_dev(dp_context *ctx, dp_patch_dev *p, int32_t *ncand, int32_t *cand,
                      void *stream);
int dp_color_dev(dp_context *ctx, dp_patch_dev *p, void *stream);

/* ---- the steps either side of the path -----------------------------------
 * dp_create_patches: Seed::CreatePatchesFromPoints (seed.cpp:26-54) for n triangulated
 * points (fp64, n*3): reference image = nearest camera centre (first minimum wins),
 * normal = unit viewing ray, position/normal stored as fp32, then InitRelatedImages.
 * `out` must have capacity n (out->n >= n on entry) and its vstride set; patch order = point
 * order.  ncand/cand (optional) as in dp_visibility.
 * dp_export_ply: the organizer's patch store as an ASCII PLY in the layout of the
 * reference's PMVS::PrintCloud (utils.cpp:9-50): x y z, red green blue, nx ny nz. */
int dp_create_patches(dp_context *ctx, const double *points, int n, dp_patch_soa *out,
                      int32_t *ncand, int32_t *cand);
int dp_export_ply(dp_context *ctx, const char *path);

/* ---- image pyramid (new; the reference's modules/image is an empty placeholder,
 * modules/image/Image.h:1-7).  Level l of view v = cv::pyrDown applied l times, with
 * P_l = diag(2^-l, 2^-l, 1) P.  dp_build_pyramid creates levels 1..n_levels-1 on the
 * device for every uploaded view; dp_set_level selects the level all later calls use. */
int dp_build_pyramid(dp_context *ctx, int n_levels);
int dp_set_level(dp_context *ctx, int level);
/* Per-(patch, view) level selection (north_star item 2, "at the chosen pyramid level"; no
 * reference semantics: options.h:10 `scale` is dead, modules/image/Image.h:1-7 a placeholder).
 * enable = 0: every view is read at the level dp_set_level chose (default).  enable = 1: that
 * level is the base (patch frame, visibility, proposals, grids, colours); the texture of view v
 * of a patch is taken from level base + k, k = the number of halvings that bring the longer of
 * the two sides of the patch's projected quad through corner 0 below px_per_cell * cell_size
 * base-level pixels (a texel then covers < px_per_cell pixels of the level it is sampled
 * from), capped at the coarsest level built.  The result is by definition the reference path
 * (GetProjectedTextures, optimization.cpp:14-56) run on pyrDown^k of view v with
 * P_k = diag(2^-k, 2^-k, 1) P; scoring, filter, refinement and expansion all use it. */
int dp_set_level_selection(dp_context *ctx, int enable, double px_per_cell);
int dp_download_level(dp_context *ctx, int view_id, int level, uint8_t *bgr, size_t capacity,
                      int *width, int *height);

#ifdef __cplusplus
}
#endif
#endif /* DENSEPOINTS_CUDA_H */
