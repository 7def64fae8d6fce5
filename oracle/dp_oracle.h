/*
 * dp_oracle.h -- CPU ORACLE for the DensePoints PMVS photometric hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, bench.py's
 * cpu_baseline / --impl reference legs and __graft_entry__.smoke() may link or
 * call it, and there only as the checker.  The product (densepoints_b200/) never
 * includes this header and has no CPU fallback.
 *
 * It is a plain-C restatement of the reference's algorithm (manlito/densepoints,
 * methods/pmvs + modules/core) with the OpenCV primitives the reference calls
 * (findHomography, warpPerspective, cvtColor, meanStdDev, Mat::dot,
 * DownhillSolver -- OpenCV is an un-vendored, unpinned dependency of the
 * reference; 4.13.0 is the executable copy in this image) restated from their
 * published algorithms.  Every function cites the reference file:line it
 * follows.
 *
 * Parity status:
 *   pinned   : NCCScore against the reference's own KAT
 *              (tests/core/test_error_functions.cpp:9-15); View decomposition
 *              against tests/core/test_projection_matrix_decomposition.cpp:10-36;
 *              homography / ROI / warp / gray / NCC against golden vectors made
 *              with cv2 4.13.0 (tests/golden/make_golden.py).
 *   UNPINNED : cv::DownhillSolver (not in the reference tree, not in the Python
 *              cv2 binding) -- "parity unpinned against upstream OpenCV"; pinned
 *              only by analytic-function tests of the documented decision tree.
 */
#ifndef DP_ORACLE_H
#define DP_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* One view: projection matrix + the quantities View::SetProjectionMatrix
 * derives (modules/core/types.cpp:28-68) + the BGR u8 image (cv::imread). */
typedef struct orc_view {
  double P[12];      /* 3x4 row-major */
  double xaxis[3];   /* View::GetXAxis(): row 0 of the extrinsic rotation */
  double center[3];  /* View::GetCameraCenter() */
  int width, height; /* image_.cols, image_.rows */
  const uint8_t *bgr; /* interleaved B,G,R */
  size_t stride;     /* bytes per row */
} orc_view;

/* PMVS constants (SURVEY.md section 5 "Config"); defaults = reference defaults. */
typedef struct orc_params {
  double score_threshold;       /* optimization.h:16   0.6  */
  int minimum_visible_image;    /* optimization.h:17   3    */
  double visible_threshold;     /* patch.h:56          0.78 */
  double candidate_threshold;   /* patch.h:57          1.04 */
  int grid_scale;               /* patch_organizer.h:43  8  */
  int max_patches_per_cell;     /* patch_organizer.h:42  1  */
  double nm_step[3];            /* optimization_opencv.cpp:56  0.02,0.2,0.2 */
  int nm_max_evals;             /* optimization_opencv.cpp:60  500 */
  double nm_eps;                /* optimization_opencv.cpp:60  1e-4 */
  long long max_pops;           /* expand.cpp:95  1e7 */
} orc_params;

void orc_default_params(orc_params *p);
/* 0 (default): sums in sequential order (Eigen 3.2); 1: Eigen >= 3.3's halving order -- a
 * measuring device for the unpinned Eigen version, see dp_oracle.c */
void orc_set_eigen_pairwise(int on);

/* Per-(patch, view) pyramid level (see dp_oracle.c): levels = [n_levels][n_views] view tables,
 * level 0 first (must be the array later calls pass as `views`); NULL or n_levels <= 1 = off. */
void orc_set_level_selection(const orc_view *levels, int n_levels, int n_views, double px_per_cell);
void orc_levels_batch(const orc_view *views, const float *pos, const float *nrm, const int *ref,
                      const int *nvis, const int *vis, int vstride, int n, int cell_size, int *out);
int orc_pick_level(const orc_view *v, int cell_size, const double pos[3], const double ax[3],
                   const double ay[3], double px_per_cell, int max_up);


/* ---- modules/core ---------------------------------------------------- */
/* View::SetProjectionMatrix (types.cpp:28-68): K, R (3x3 row-major), centre. */
void orc_view_decompose(const double P[12], double K[9], double R[9], double center[3]);
void orc_view_init(orc_view *v, const double P[12], const uint8_t *bgr, int width, int height,
                   size_t stride);
/* View::ProjectPoint (types.cpp:70-75) */
void orc_project(const orc_view *v, const double X[3], double uv[2]);
/* View::IsPointInside (types.cpp:77-84) */
int orc_inside(const orc_view *v, const double X[3]);
/* NCCScore (error_measurements.cpp:36-60) on s*s*3 BGR u8 textures; NULL = empty Mat. */
double orc_ncc_bgr(const uint8_t *tex_a, const uint8_t *tex_b, int n_px);
/* NCCScore, non-8UC3 branch (ToFloatMat convertTo CV_32F), for the reference KAT. */
double orc_ncc_f64(const double *a, const double *b, int n);

/* ---- OpenCV primitives restated ---------------------------------------- */
/* cv::findHomography(src, dst, 0) for exactly 4 float points (normalised DLT,
 * Jacobi eigen-solve of LtL, H /= H22).  Returns 0 if degenerate. */
int orc_find_homography4(const float src[8], const float dst[8], double H[9]);
/* cv::warpPerspective(src(roi), H, (s,s), INTER_LINEAR, BORDER_REPLICATE), 8UC3. */
void orc_warp_perspective(const uint8_t *src, size_t stride, int w, int h, const double H[9],
                          int s, uint8_t *dst);
/* Homography mode of GetProjectedTextures:
 *   0 (default) = the OpenCV procedure: findHomography's normalised DLT + Jacobi
 *       eigen-solve, then warpPerspective's 3x3 inversion.  Pinned against cv2.
 *   1 = the exact projective map cell -> quad in closed form (orc_cell_to_quad).
 * The two differ by ~1e-14 px in the source coordinates, which changes a texel only at
 * an exact tie (a coordinate of exactly k + 1/2 in 1/32-px units, reachable at texel
 * (0,0) because the quad corner is an fp32 number): there OpenCV's own result is decided
 * by the sign of its eigen-solver's rounding noise, i.e. it is not a function of the
 * inputs.  Mode 1 defines the tie by the exact value and is what the CUDA path is
 * checked against. */
void orc_set_homography_mode(int mode);
int orc_get_homography_mode(void);
int orc_cell_to_quad(const float quad[8], int s, double M[9]);
int orc_patch_quad(const orc_view *v, const double pos[3], const double ax[3], const double ay[3],
                   float pts[8], int roi[4]);
/* cv::pyrDown, CV_8UC3 (pyramid extension; pinned against cv2 golden vectors). */
void orc_pyrdown(const uint8_t *src, size_t sstride, int w, int h, uint8_t *dst, size_t dstride);
/* cv::cvtColor(BGR2GRAY) for one pixel (OpenCV 4.x 15-bit constants). */
int orc_gray(int b, int g, int r);
/* cv::DownhillSolver::minimize restated (ndim <= 8). Returns f(best); x <- best. */
typedef double (*orc_fn)(const double *x, void *user);
double orc_downhill(orc_fn f, void *user, int ndim, double *x, const double *step, int max_evals,
                    double eps, int *fcount);

/* ---- methods/pmvs: Patch / Optimization --------------------------------- */
/* Patch::GetProjectedXYAxisAndScale (patch.cpp:86-104).  */
void orc_axes_scale(const orc_view *ref, const double nrm[3], const double pos[3], double xa[3],
                    double ya[3], double *dx);
/* Patch::ComputePatchToViewHomography (patch.cpp:111-164). roi = x,y,w,h. */
int orc_patch_homography(const orc_view *v, int cell_size, const double pos[3], const double ax[3],
                         const double ay[3], double H[9], int roi[4]);
/* Optimization::GetProjectedTextures(normal, position, textures) (optimization.cpp:14-56).
 * nrm / pos = the arguments (axes and dx only); centre = patch_.GetPosition(), around which
 * ComputePatchToViewHomography builds the corners (patch.cpp:119-123).
 * tex: nvis * s*s*3 bytes; valid[k] = 0 for an empty cv::Mat. */
void orc_projected_textures(const orc_view *views, int ref, const int *vis, int nvis,
                            int cell_size, const double nrm[3], const double pos[3],
                            const double centre[3], uint8_t *tex, uint8_t *valid);
/* scores[k-1] = NCCScore(tex0, texk) k=1..nvis-1 (optimization.cpp:104-110). */
void orc_scores(const orc_view *views, int ref, const int *vis, int nvis, int cell_size,
                const float nrm[3], const float pos[3], double *scores);
/* Optimization::FilterByErrorMeasurement (optimization.cpp:98-132). vis edited in place. */
int orc_filter_by_error(const orc_view *views, int ref, int *vis, int *nvis, int cell_size,
                        const float nrm[3], const float pos[3], double thr, int min_visible);
/* Optimization::UnparametrizePatch (optimization.cpp:78-96). */
void orc_unparametrize(const orc_view *ref, const float nrm0[3], const float pos0[3], double depth,
                       double roll, double pitch, double nrm[3], double pos[3]);
/* PatchOptimizationOpenCVFunctor::calc (optimization_opencv.cpp:14-39). */
double orc_objective(const orc_view *views, int ref, const int *vis, int nvis, int cell_size,
                     const float nrm0[3], const float pos0[3], const double x[3]);
/* OptimizationOpenCV::Optimize (optimization_opencv.cpp:44-78). nrm/pos updated (fp32). */
int orc_optimize(const orc_view *views, int ref, const int *vis, int nvis, int cell_size,
                 float nrm[3], float pos[3], const orc_params *prm, int *fcount, double xbest[3]);
/* Patch::InitRelatedImages (patch.cpp:19-49). Returns via vis/cand (capacity n_views). */
void orc_init_related_images(const orc_view *views, int n_views, int ref, const float nrm[3],
                             const float pos[3], double t_vis, double t_cand, int *vis, int *nvis,
                             int *cand, int *ncand);
/* Patch::ComputeColor (patch.cpp:51-73): rgb[0]=r,[1]=g,[2]=b. */
void orc_compute_color(const orc_view *views, int n_views, const float pos[3], uint8_t rgb[3]);

/* ---- batched drivers (Seed::FilterPatches / OptimizePatches, seed.cpp:110-144) ----
 * SoA patch arrays: pos,nrm n*3 f32; ref n i32; nvis n i32; vis n*vstride i32.
 * OpenMP over patches when built with -fopenmp (these are what the CPU baseline times). */
void orc_score_batch(const orc_view *views, int n, const float *pos, const float *nrm,
                     const int *ref, const int *nvis, const int *vis, int vstride, int cell_size,
                     float *ncc /* n*vstride, [k] = score of vis[k], k>=1; [0] unused */,
                     uint8_t *tex /* optional n*vstride*s*s*3 */, uint8_t *valid /* optional */);
/* the same at trial parameters = Optimization::GetProjectedTextures(normal, position, ...)
 * (optimization.cpp:14-56): trial_nrm / trial_pos n*3 fp64, NULL = the patch's own. */
void orc_score_at_batch(const orc_view *views, int n, const float *pos, const float *nrm,
                        const int *ref, const int *nvis, const int *vis, int vstride,
                        int cell_size, const double *trial_nrm, const double *trial_pos,
                        float *ncc, uint8_t *tex, uint8_t *valid);
void orc_filter_batch(const orc_view *views, int n, const float *pos, const float *nrm,
                      const int *ref, int *nvis, int *vis, int vstride, int cell_size, double thr,
                      int min_visible, uint8_t *keep);
void orc_refine_batch(const orc_view *views, int n, float *pos, float *nrm, const int *ref,
                      const int *nvis, const int *vis, int vstride, int cell_size,
                      const orc_params *prm, int *fcount /* n */, double *xbest /* optional n*3 */);
void orc_visibility_batch(const orc_view *views, int n_views, int n, const float *pos,
                          const float *nrm, const int *ref, double t_vis, double t_cand, int *nvis,
                          int *vis, int *ncand, int *cand, int vstride);

/* Seed::CreatePatchesFromPoints (seed.cpp:26-54); patches in point order. */
void orc_create_patches(const orc_view *views, int n_views, int n, const double *points,
                        double t_vis, double t_cand, float *pos, float *nrm, int *ref, int *nvis,
                        int *vis, int vstride);

/* ---- PatchOrganizer + Expand (patch_organizer.cpp, expand.cpp), 1-thread FIFO ---- */
typedef struct orc_organizer orc_organizer;
orc_organizer *orc_organizer_create(const orc_view *views, int n_views, const orc_params *prm);
void orc_organizer_destroy(orc_organizer *o);
/* PatchOrganizer::TryInsert (patch_organizer.cpp:42-65). Returns store index or -1.
 * cells_out (optional, nvis*3): view,row,col for each cell won. */
long long orc_organizer_try_insert(orc_organizer *o, const float pos[3], const float nrm[3], int ref,
                                   const int *vis, int nvis, int *ncells_out, int *cells_out);
long long orc_organizer_size(const orc_organizer *o);
/* occupancy grid of one view: counts[h*w] (u8), dims out */
const uint8_t *orc_organizer_grid(const orc_organizer *o, int view, int *gw, int *gh);
/* copy patch store out (pos,nrm n*3; rgb n*3; ref,nvis n; vis n*vstride (-1 pad)) */
void orc_organizer_export(const orc_organizer *o, float *pos, float *nrm, uint8_t *rgb, int *ref,
                          int *nvis, int *vis, int vstride);
/* Expand::ExpandPatch (expand.cpp:103-143): up to 4 accepted children of one parent
 * (pvis/pn = the parent's visible set, copied into each child at expand.cpp:125).
 * out arrays sized 4; returns count; dir_out[k] = direction index of kth accepted. */
int orc_expand_patch(const orc_view *views, int n_views, const orc_params *prm, int cell_size,
                     const float pos[3], const float nrm[3], int ref, const int *pvis, int pn,
                     float out_pos[12], float out_nrm[12], int *out_nvis,
                     int *out_vis /* 4*n_views */, int *dir_out);
/* Expand::ExpandPatches (expand.cpp:34-101) as the literal single-thread FIFO.
 * max_pops < 0 : the reference's 1e7 cap. Returns number of pops. */
long long orc_expand_patches_fifo(orc_organizer *o, int cell_size, long long max_pops);
/* The same FIFO order evaluated level-synchronously (OpenMP over the parents of a
 * level, inserts in FIFO order).  max_levels < 0 : run until the queue is empty
 * (reference behaviour); otherwise stop after that many BFS levels.  Returns pops. */
long long orc_expand_patches(orc_organizer *o, int cell_size, int max_levels);

int orc_num_threads(void);
/* overrides OMP_NUM_THREADS (torchrun exports OMP_NUM_THREADS=1 to its workers) */
void orc_set_num_threads(int n);

#ifdef __cplusplus
}
#endif
#endif /* DP_ORACLE_H */
