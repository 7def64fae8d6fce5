// densepoints_cuda.cu -- C ABI (include/densepoints_cuda.h) over the sm_100a kernels.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -shared -Xcompiler -fPIC
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "dp_aux_kernels.cuh"
#include "dp_context.h"
#include "dp_kernels.cuh"
#include "dp_group.cuh"
#include "dp_lane.cuh"

// ---------------------------------------------------------------------------------------
// helpers

int dp_fail(dp_context *ctx, int code, const char *what, cudaError_t e) {
  if (ctx) {
    ctx->err = what ? what : "error";
    if (e != cudaSuccess) {
      ctx->err += ": ";
      ctx->err += cudaGetErrorString(e);
    }
  }
  return code;
}

int dp_scratch_acquire(dp_context *ctx, cudaStream_t st) {
  if (ctx->scratch_busy && ctx->scratch_stream != st)
    DP_CUDA(ctx, cudaStreamWaitEvent(st, ctx->scratch_event, 0));
  return DP_OK;
}
int dp_scratch_release(dp_context *ctx, cudaStream_t st) {
  DP_CUDA(ctx, cudaEventRecord(ctx->scratch_event, st));
  ctx->scratch_stream = st;
  ctx->scratch_busy = true;
  return DP_OK;
}

static int npass_for(int s) {
  int npx = s * s;
  int np = (npx + 31) / 32;
  int p = 1;
  while (p < np) p <<= 1;
  return p;
}

extern "C" void dp_default_params(dp_params *p) {
  if (!p) return;
  p->score_threshold = 0.6;
  p->minimum_visible_image = 3;
  p->visible_threshold = 0.78;
  p->candidate_threshold = 1.04;
  p->grid_scale = 8;
  p->max_patches_per_cell = 1;
  p->nm_step[0] = 0.02;
  p->nm_step[1] = 0.2;
  p->nm_step[2] = 0.2;
  p->nm_max_evals = 500;
  p->nm_eps = 0.0001;
  p->max_pops = 10000000LL;
}

extern "C" int dp_abi_version(void) { return DP_ABI_VERSION; }

static int check_params(dp_context *ctx, const dp_params *p) {
  if (p->grid_scale <= 0) return dp_fail(ctx, DP_ERR_INVALID_ARG, "grid_scale must be > 0");
  if (p->max_patches_per_cell < 1 || p->max_patches_per_cell > 255)
    return dp_fail(ctx, DP_ERR_INVALID_ARG, "max_patches_per_cell must be in [1, 255]");
  if (p->nm_max_evals < 4) return dp_fail(ctx, DP_ERR_INVALID_ARG, "nm_max_evals must be >= 4");
  return DP_OK;
}

extern "C" int dp_create(dp_context **out, int device, const dp_params *params) {
  if (!out) return DP_ERR_INVALID_ARG;
  *out = nullptr;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) return DP_ERR_NO_DEVICE;
  if (device < 0) {
    if (cudaGetDevice(&device) != cudaSuccess) return DP_ERR_NO_DEVICE;
  }
  if (device >= count) return DP_ERR_NO_DEVICE;
  DpDeviceGuard guard__(device);
  dp_context *ctx = new dp_context();
  ctx->device = device;
  if (params)
    ctx->prm = *params;
  else
    dp_default_params(&ctx->prm);
  if (check_params(ctx, &ctx->prm) != DP_OK) {
    delete ctx;
    return DP_ERR_INVALID_ARG;
  }
  if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) {
    delete ctx;
    return DP_ERR_CUDA;
  }
  if (cudaEventCreateWithFlags(&ctx->scratch_event, cudaEventDisableTiming) != cudaSuccess) {
    cudaStreamDestroy(ctx->stream);
    delete ctx;
    return DP_ERR_CUDA;
  }
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) ctx->sm_count = prop.multiProcessorCount;
  *out = ctx;
  return DP_OK;
}

static void free_views(dp_context *ctx) {
  for (auto &v : ctx->views)
    for (auto &l : v.levels)
      if (l.img) cudaFree(l.img);
  ctx->views.clear();
  ctx->views_dirty = true;
}

extern "C" void dp_destroy(dp_context *ctx) {
  if (!ctx) return;
  DpDeviceGuard guard__(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  free_views(ctx);
  DpDevBuf *bufs[] = {&ctx->d_views, &ctx->s_pos, &ctx->s_nrm, &ctx->s_ref, &ctx->s_nvis,
                      &ctx->s_vis, &ctx->s_rgb, &ctx->s_ncc, &ctx->s_tex, &ctx->s_valid,
                      &ctx->s_keep, &ctx->s_evals, &ctx->s_xbest, &ctx->s_cand, &ctx->s_ncand,
                      &ctx->s_img, &ctx->s_misc, &ctx->work_counter, &ctx->s_order, &ctx->d_views_lv, &ctx->s_nmsave, &ctx->s_pending, &ctx->e_pos, &ctx->e_nrm,
                      &ctx->e_ref, &ctx->e_nvis, &ctx->e_vis, &ctx->e_keep, &ctx->e_seq,
                      &ctx->e_cells, &ctx->e_recs, &ctx->e_flags, &ctx->e_scan, &ctx->e_count, &ctx->e_won, &ctx->org.grid,
                      &ctx->org.claim, &ctx->org.pos, &ctx->org.nrm, &ctx->org.rgb, &ctx->org.ref,
                      &ctx->org.nvis, &ctx->org.vis};
  for (DpDevBuf *b : bufs) b->release();
  for (cudaEvent_t e : ctx->pipe_events) cudaEventDestroy(e);
  if (ctx->scratch_event) cudaEventDestroy(ctx->scratch_event);
  if (ctx->stream_in) cudaStreamDestroy(ctx->stream_in);
  if (ctx->stream_out) cudaStreamDestroy(ctx->stream_out);
  cudaStreamDestroy(ctx->stream);
  delete ctx;
}

extern "C" const char *dp_last_error(const dp_context *ctx) {
  return ctx ? ctx->err.c_str() : "null context";
}

extern "C" int dp_set_params(dp_context *ctx, const dp_params *p) {
  if (!ctx || !p) return DP_ERR_INVALID_ARG;
  int rc = check_params(ctx, p);
  if (rc != DP_OK) return rc;
  if (p->grid_scale != ctx->prm.grid_scale) {
    ctx->views_dirty = true;
    ctx->org.ready = false;
  }
  ctx->prm = *p;
  return DP_OK;
}

extern "C" int dp_get_params(const dp_context *ctx, dp_params *p) {
  if (!ctx || !p) return DP_ERR_INVALID_ARG;
  *p = ctx->prm;
  return DP_OK;
}

extern "C" int dp_sync(dp_context *ctx) {
  if (!ctx) return DP_ERR_INVALID_ARG;
  DpDeviceGuard guard__(ctx->device);
  DP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return DP_OK;
}

extern "C" int64_t dp_launch_count(const dp_context *ctx) { return ctx ? ctx->launches : 0; }

// ---------------------------------------------------------------------------------------
// views

// View::SetProjectionMatrix (core/types.cpp:28-68): camera centre (null vector of P) and the
// orthogonal factor of the RQ decomposition of P[:, :3] with a positive-diagonal K; only its
// first row (View::GetXAxis) is needed on the path.
static void decompose_projection(const double P[12], double xaxis[3], double center[3]) {
  const double m00 = P[0], m01 = P[1], m02 = P[2], m10 = P[4], m11 = P[5], m12 = P[6],
               m20 = P[8], m21 = P[9], m22 = P[10];
  const double b0 = -P[3], b1 = -P[7], b2 = -P[11];
  // Cramer: M c = b
  const double c00 = m11 * m22 - m12 * m21, c01 = m12 * m20 - m10 * m22, c02 = m10 * m21 - m11 * m20;
  const double det = m00 * c00 + m01 * c01 + m02 * c02;
  const double c10 = m02 * m21 - m01 * m22, c11 = m00 * m22 - m02 * m20, c12 = m01 * m20 - m00 * m21;
  const double c20 = m01 * m12 - m02 * m11, c21 = m02 * m10 - m00 * m12, c22 = m00 * m11 - m01 * m10;
  center[0] = (c00 * b0 + c10 * b1 + c20 * b2) / det;
  center[1] = (c01 * b0 + c11 * b1 + c21 * b2) / det;
  center[2] = (c02 * b0 + c12 * b1 + c22 * b2) / det;
  // Gram-Schmidt from the last row up: r2 = m2/|m2|, r1 _|_ r2, r0 _|_ r1, r2
  double r2[3] = {m20, m21, m22};
  double n2 = sqrt(r2[0] * r2[0] + r2[1] * r2[1] + r2[2] * r2[2]);
  for (double &v : r2) v /= n2;
  double r1[3] = {m10, m11, m12};
  double d12 = r1[0] * r2[0] + r1[1] * r2[1] + r1[2] * r2[2];
  for (int j = 0; j < 3; ++j) r1[j] -= d12 * r2[j];
  double n1 = sqrt(r1[0] * r1[0] + r1[1] * r1[1] + r1[2] * r1[2]);
  for (double &v : r1) v /= n1;
  double r0[3] = {m00, m01, m02};
  double d02 = r0[0] * r2[0] + r0[1] * r2[1] + r0[2] * r2[2];
  double d01 = r0[0] * r1[0] + r0[1] * r1[1] + r0[2] * r1[2];
  for (int j = 0; j < 3; ++j) r0[j] -= d02 * r2[j] + d01 * r1[j];
  double n0 = sqrt(r0[0] * r0[0] + r0[1] * r0[1] + r0[2] * r0[2]);
  for (int j = 0; j < 3; ++j) xaxis[j] = r0[j] / n0;
}

extern "C" int dp_set_num_views(dp_context *ctx, int n_views) {
  if (!ctx || n_views < 0 || n_views > 65535) return dp_fail(ctx, DP_ERR_INVALID_ARG, "n_views");
  DpDeviceGuard guard__(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  free_views(ctx);
  ctx->views.resize(n_views);
  ctx->n_levels = 1;
  ctx->level = 0;
  ctx->org.ready = false;
  return DP_OK;
}

extern "C" int dp_num_views(const dp_context *ctx) { return ctx ? (int)ctx->views.size() : 0; }

extern "C" int dp_upload_view(dp_context *ctx, int view_id, const double P[12], const double *xaxis,
                              const double *center, const uint8_t *bgr, int width, int height,
                              size_t stride) {
  if (!ctx) return DP_ERR_INVALID_ARG;
  if (view_id < 0 || view_id >= (int)ctx->views.size())
    return dp_fail(ctx, DP_ERR_INVALID_ARG, "view_id out of range (call dp_set_num_views)");
  if (!P || !bgr || width <= 0 || height <= 0 || stride < (size_t)width * 3)
    return dp_fail(ctx, DP_ERR_INVALID_ARG, "dp_upload_view arguments");
  DpDeviceGuard guard__(ctx->device);
  DpViewHost &v = ctx->views[view_id];
  for (auto &l : v.levels)
    if (l.img) cudaFree(l.img);
  v.levels.clear();
  memcpy(v.P, P, sizeof(double) * 12);
  double xa[3], c[3];
  decompose_projection(P, xa, c);
  for (int j = 0; j < 3; ++j) {
    v.xaxis[j] = xaxis ? xaxis[j] : xa[j];
    v.center[j] = center ? center[j] : c[j];
  }
  DpLevel l;
  l.width = width;
  l.height = height;
  l.pitch_px = (width + 31) & ~31;  // 128-byte aligned rows
  // one spare row: the texel pass may read (never use) the right / lower neighbour of an edge pixel
  DP_CUDA(ctx, cudaMalloc(&l.img, (size_t)l.pitch_px * (height + 1) * sizeof(uint32_t)));
  v.levels.push_back(l);  // owned by the view from here on: freed by free_views / the next upload
  v.set = false;
  DP_CUDA(ctx, cudaMemsetAsync(l.img + (size_t)l.pitch_px * height, 0, (size_t)l.pitch_px * 4, ctx->stream));
  const size_t bytes = stride * (size_t)height;
  // (the staging buffer is reused by the next upload, hence the synchronisation below)
  DP_CUDA(ctx, ctx->s_img.ensure(bytes));
  DP_CUDA(ctx, cudaMemcpyAsync(ctx->s_img.ptr, bgr, bytes, cudaMemcpyHostToDevice, ctx->stream));
  dim3 grid((l.pitch_px + 255) / 256, height);
  dp_pack_bgrx_kernel<<<grid, 256, 0, ctx->stream>>>(ctx->s_img.as<uint8_t>(), stride, width,
                                                      height, l.img, l.pitch_px);
  ++ctx->launches;
  DP_CUDA(ctx, cudaGetLastError());
  DP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  v.set = true;
  ctx->views_dirty = true;
  ctx->org.ready = false;
  return DP_OK;
}

extern "C" int dp_get_view(const dp_context *ctx, int view_id, double xaxis[3], double center[3],
                           int *width, int *height) {
  if (!ctx || view_id < 0 || view_id >= (int)ctx->views.size() || !ctx->views[view_id].set)
    return DP_ERR_INVALID_ARG;
  const DpViewHost &v = ctx->views[view_id];
  const DpLevel &l = v.levels[std::min<int>(ctx->level, (int)v.levels.size() - 1)];
  for (int j = 0; j < 3; ++j) {
    if (xaxis) xaxis[j] = v.xaxis[j];
    if (center) center[j] = v.center[j];
  }
  if (width) *width = l.width;
  if (height) *height = l.height;
  return DP_OK;
}

// Builds the device view table for the active pyramid level.  Level l uses
// P_l = diag(2^-l, 2^-l, 1) P (exact scaling by a power of two).
int dp_sync_views(dp_context *ctx) {
  if (!ctx->views_dirty) return DP_OK;
  const int nv = (int)ctx->views.size();
  if (nv == 0) return dp_fail(ctx, DP_ERR_STATE, "no views uploaded");
  const int n_tab = ctx->auto_level ? ctx->n_levels - ctx->level : 1;  // base level + the ones above
  std::vector<DpViewDev> h((size_t)nv * n_tab);
  long long off = 0;
  for (int i = 0; i < nv; ++i) {
    const DpViewHost &v = ctx->views[i];
    if (!v.set) return dp_fail(ctx, DP_ERR_STATE, "a view was not uploaded");
    if (ctx->level + n_tab > (int)v.levels.size())
      return dp_fail(ctx, DP_ERR_STATE, "pyramid level not built");
    const double n = sqrt(v.xaxis[0] * v.xaxis[0] + v.xaxis[1] * v.xaxis[1] + v.xaxis[2] * v.xaxis[2]);
    for (int t = 0; t < n_tab; ++t) {
      DpViewDev &d = h[(size_t)t * nv + i];
      const DpLevel &l = v.levels[ctx->level + t];
      const double sc = ldexp(1.0, -(ctx->level + t));
      for (int j = 0; j < 12; ++j) d.P[j] = (j < 8) ? v.P[j] * sc : v.P[j];
      for (int j = 0; j < 3; ++j) {
        d.xa[j] = v.xaxis[j] / n;  // .normalized(), patch.cpp:95
        d.center[j] = v.center[j];
      }
      d.img = l.img;
      d.width = l.width;
      d.height = l.height;
      d.pitch_px = l.pitch_px;
      d.gw = l.width / ctx->prm.grid_scale;   // patch_organizer.cpp:35-36
      d.gh = l.height / ctx->prm.grid_scale;
      d.grid_off = off;   // (the grids belong to the base level; the tables above it are only
    }                     //  read for their image and projection)
    off += (long long)h[i].gw * h[i].gh;
  }
  DP_CUDA(ctx, ctx->d_views.ensure(sizeof(DpViewDev) * nv));
  DP_CUDA(ctx, cudaMemcpyAsync(ctx->d_views.ptr, h.data(), sizeof(DpViewDev) * nv,
                               cudaMemcpyHostToDevice, ctx->stream));
  if (n_tab > 1) {
    DP_CUDA(ctx, ctx->d_views_lv.ensure(sizeof(DpViewDev) * h.size()));
    DP_CUDA(ctx, cudaMemcpyAsync(ctx->d_views_lv.ptr, h.data(), sizeof(DpViewDev) * h.size(),
                                 cudaMemcpyHostToDevice, ctx->stream));
  }
  DP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  ctx->org.n_cells = off;
  ctx->views_dirty = false;
  return DP_OK;
}

// ---------------------------------------------------------------------------------------
// device-pointer layer

static int check_patch_dev_args(dp_context *ctx, int cell_size) {
  if (cell_size < 2 || cell_size > DP_MAX_CELL_SIZE)
    return dp_fail(ctx, DP_ERR_INVALID_ARG, "cell_size must be in [2, 32]");
  return DP_OK;
}

static int check_patch_dev(dp_context *ctx, const dp_patch_dev *p, int cell_size) {
  if (!ctx || !p) return DP_ERR_INVALID_ARG;
  if (p->n < 0 || p->vstride <= 0) return dp_fail(ctx, DP_ERR_INVALID_ARG, "patch batch shape");
  if (cell_size < 2 || cell_size > DP_MAX_CELL_SIZE)
    return dp_fail(ctx, DP_ERR_INVALID_ARG, "cell_size must be in [2, 32]");
  if (p->n > 0 && (!p->pos || !p->nrm || !p->ref || !p->nvis || !p->vis))
    return dp_fail(ctx, DP_ERR_INVALID_ARG, "null patch array");
  return DP_OK;
}

static DpPatchArgs patch_args(dp_context *ctx, const dp_patch_dev *p, int s) {
  DpPatchArgs a;
  a.views = ctx->d_views.as<DpViewDev>();
  a.n_views = (int)ctx->views.size();
  a.n = p->n;
  a.vstride = p->vstride;
  a.pos = p->pos;
  a.nrm = p->nrm;
  a.ref = p->ref;
  a.nvis = p->nvis;
  a.vis = p->vis;
  a.s = s;
  // per-(patch, view) level: only when levels above the base exist
  const int up = ctx->auto_level ? ctx->n_levels - 1 - ctx->level : 0;
  a.lv.tab = up > 0 ? ctx->d_views_lv.as<DpViewDev>() : nullptr;
  a.lv.up = up;
  const double side = ctx->level_px * (double)s;
  a.lv.thr2 = side * side;
  return a;
}

// Work order for the kernels that process several patches per warp in lockstep, and for the
// persistent refine warps: patches by descending view count (counting sort, 3 tiny kernels).
// Results do not depend on the order.  Small batches keep their natural order (*order = null).
static int build_order(dp_context *ctx, const int32_t *nvis, const uint8_t *mask, int n,
                       cudaStream_t st, const int32_t **order_out) {
  *order_out = nullptr;
  if (n < 4096) return DP_OK;
  DP_CUDA(ctx, ctx->s_order.ensure((size_t)n * 4 + DP_ORDER_BINS * 4));
  int32_t *order = ctx->s_order.as<int32_t>();
  unsigned int *hist = reinterpret_cast<unsigned int *>(order + n);
  DP_CUDA(ctx, cudaMemsetAsync(hist, 0, DP_ORDER_BINS * 4, st));
  const unsigned blocks = (unsigned)((n + 255) / 256);
  dp_order_hist_kernel<<<blocks, 256, 0, st>>>(nvis, mask, n, hist);
  dp_order_scan_kernel<<<1, 1, 0, st>>>(hist);
  dp_order_scatter_kernel<<<blocks, 256, 0, st>>>(nvis, mask, n, hist, order);
  ctx->launches += 3;
  *order_out = order;
  return DP_OK;
}

#ifndef DP_SCORE_LANE
#define DP_SCORE_LANE 1   // cells up to 8x8: one patch per lane (dp_lane.cuh)
#endif
#ifndef DP_SCORE_GROUP
#define DP_SCORE_GROUP 1  // cells up to 16x16: several patches per warp (dp_group.cuh)
#endif

// tuning / A-B runs only: DP_REFINE_KERNEL=group and DP_SCORE_KERNEL=group select the
// 4-lanes-per-patch kernels instead of the one-patch-per-lane kernels (dp_lane.cuh)
static bool env_is_group(const char *name) {
  const char *e = getenv(name);
  return e && strcmp(e, "group") == 0;
}
static bool use_lane_score() { return DP_SCORE_LANE && !env_is_group("DP_SCORE_KERNEL"); }

template <bool TEX, bool FILT>
static int launch_score(dp_context *ctx, const DpScoreArgs &a, int cell_size, cudaStream_t st) {
  if (cell_size <= DP_LGROUP_MAX_CELL && use_lane_score()) {
    const int32_t *order = nullptr;
    int rc = build_order(ctx, a.p.nvis, nullptr, a.p.n, st, &order);
    if (rc != DP_OK) return rc;
    const long long per_cta = (long long)DP_LWARPS * 32;
    const unsigned grid = (unsigned)((a.p.n + per_cta - 1) / per_cta);
#define DP_LCASE(S) case S: dp_score_lane_kernel<S, TEX, FILT><<<grid, DP_LWARPS * 32, 0, st>>>(a, order); break
    switch (cell_size) {
      DP_LCASE(2); DP_LCASE(3); DP_LCASE(4); DP_LCASE(5); DP_LCASE(6); DP_LCASE(7); DP_LCASE(8);
    }
#undef DP_LCASE
    ++ctx->launches;
    return DP_OK;
  }
  if (DP_SCORE_GROUP && cell_size <= DP_GROUP_MAX_CELL_SCORE) {
    const int32_t *order = nullptr;
    int rc = build_order(ctx, a.p.nvis, nullptr, a.p.n, st, &order);
    if (rc != DP_OK) return rc;
#define DP_GCASE(S)                                                                        \
  case S: {                                                                                \
    using C = DpCfgFor<S>;                                                                 \
    const long long per_cta = (long long)DP_GWARPS * C::GROUPS;                            \
    const unsigned grid = (unsigned)((a.p.n + per_cta - 1) / per_cta);                     \
    dp_score_group_kernel<C, TEX, FILT><<<grid, DP_GWARPS * 32, 0, st>>>(a, order);        \
  } break
    switch (cell_size) {
      DP_GCASE(2); DP_GCASE(3); DP_GCASE(4); DP_GCASE(5); DP_GCASE(6); DP_GCASE(7); DP_GCASE(8);
      DP_GCASE(9); DP_GCASE(10); DP_GCASE(11); DP_GCASE(12); DP_GCASE(13); DP_GCASE(14);
      DP_GCASE(15); DP_GCASE(16);
    }
#undef DP_GCASE
    ++ctx->launches;
    return DP_OK;
  }
  const unsigned grid = (unsigned)((a.p.n + DP_WARPS - 1) / DP_WARPS);
  switch (npass_for(cell_size)) {
    case 1: dp_score_kernel<1, TEX, FILT><<<grid, DP_WARPS * 32, 0, st>>>(a); break;
    case 2: dp_score_kernel<2, TEX, FILT><<<grid, DP_WARPS * 32, 0, st>>>(a); break;
    case 4: dp_score_kernel<4, TEX, FILT><<<grid, DP_WARPS * 32, 0, st>>>(a); break;
    case 8: dp_score_kernel<8, TEX, FILT><<<grid, DP_WARPS * 32, 0, st>>>(a); break;
    case 16: dp_score_kernel<16, TEX, FILT><<<grid, DP_WARPS * 32, 0, st>>>(a); break;
    default: dp_score_kernel<32, TEX, FILT><<<grid, DP_WARPS * 32, 0, st>>>(a); break;
  }
  ++ctx->launches;
  return DP_OK;
}

extern "C" int dp_score_at_dev(dp_context *ctx, const dp_patch_dev *p, int cell_size,
                               const double *normal, const double *position, float *ncc,
                               uint8_t *tex, uint8_t *valid, void *stream) {
  int rc = check_patch_dev(ctx, p, cell_size);
  if (rc != DP_OK) return rc;
  if (p->n == 0) return DP_OK;
  DpDeviceGuard guard__(ctx->device);
  if ((rc = dp_sync_views(ctx)) != DP_OK) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  if ((rc = dp_scratch_acquire(ctx, st)) != DP_OK) return rc;
  DpScoreArgs a;
  a.p = patch_args(ctx, p, cell_size);
  a.ncc = ncc;
  a.tex = tex;
  a.valid = valid;
  a.trial_nrm = normal;
  a.trial_pos = position;
  a.thr = 0;
  a.min_visible = 0;
  a.keep = nullptr;
  if (ncc) DP_CUDA(ctx, cudaMemsetAsync(ncc, 0, sizeof(float) * (size_t)p->n * p->vstride, st));
  if (valid) DP_CUDA(ctx, cudaMemsetAsync(valid, 0, (size_t)p->n * p->vstride, st));
  if (tex) {
    DP_CUDA(ctx, cudaMemsetAsync(tex, 0, (size_t)p->n * p->vstride * cell_size * cell_size * 3, st));
    rc = launch_score<true, false>(ctx, a, cell_size, st);
  } else {
    rc = launch_score<false, false>(ctx, a, cell_size, st);
  }
  if (rc != DP_OK) return rc;
  DP_CUDA(ctx, cudaGetLastError());
  return dp_scratch_release(ctx, st);
}

extern "C" int dp_score_dev(dp_context *ctx, const dp_patch_dev *p, int cell_size, float *ncc,
                            uint8_t *tex, uint8_t *valid, void *stream) {
  return dp_score_at_dev(ctx, p, cell_size, nullptr, nullptr, ncc, tex, valid, stream);
}

extern "C" int dp_filter_dev(dp_context *ctx, dp_patch_dev *p, int cell_size, uint8_t *keep,
                             void *stream) {
  int rc = check_patch_dev(ctx, p, cell_size);
  if (rc != DP_OK) return rc;
  if (!keep) return dp_fail(ctx, DP_ERR_INVALID_ARG, "keep is null");
  if (p->n == 0) return DP_OK;
  DpDeviceGuard guard__(ctx->device);
  if ((rc = dp_sync_views(ctx)) != DP_OK) return rc;
  if ((rc = dp_scratch_acquire(ctx, (cudaStream_t)stream)) != DP_OK) return rc;
  DpScoreArgs a;
  a.p = patch_args(ctx, p, cell_size);
  a.ncc = nullptr;
  a.tex = nullptr;
  a.valid = nullptr;
  a.trial_nrm = a.trial_pos = nullptr;
  a.thr = ctx->prm.score_threshold;
  a.min_visible = ctx->prm.minimum_visible_image;
  a.keep = keep;
  if ((rc = launch_score<false, true>(ctx, a, cell_size, (cudaStream_t)stream)) != DP_OK) return rc;
  DP_CUDA(ctx, cudaGetLastError());
  return dp_scratch_release(ctx, (cudaStream_t)stream);
}

template <int NPASS, int WPP>
static cudaError_t launch_refine_wpp(const DpRefineArgs &a, int sm_count, cudaStream_t st) {
  int per_sm = 1;
  cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(
      &per_sm, dp_refine_kernel<NPASS, WPP>, DP_RWARPS * 32, 0);
  if (e != cudaSuccess) return e;
  if (per_sm < 1) per_sm = 1;
  const long long per_cta = DP_RWARPS / WPP;
  long long want = ((long long)a.n_items + per_cta - 1) / per_cta;
  long long grid = std::max<long long>(1, std::min<long long>(want, (long long)sm_count * per_sm));
  dp_refine_kernel<NPASS, WPP><<<(unsigned)grid, DP_RWARPS * 32, 0, st>>>(a);
  return cudaGetLastError();
}
static long long env_ll(const char *name, long long dflt) {
  const char *e = getenv(name);
  return e ? atoll(e) : dflt;
}
// Warp-per-patch refinement (cells > 8), TIME-SLICED.  A launch of persistent warps lasts at
// least as long as its longest patch -- up to 500 dependent evaluations of up to ~50 views, 10-12
// ms with one warp per patch -- however few patches it has: measured on the 64-view expansion
// split over 8 GPUs (tools/shard_balance.py), levels whose ideal share was 2-8 ms took 4-14 ms.
// So a launch gives every patch a budget of evaluations; a patch that needs more is stopped with
// its Nelder-Mead state saved bit for bit (DpRefineArgs::nm_save) and continues in the next
// launch -- and the fewer patches are left, the more warps each one gets (four, then the whole
// CTA of eight: the views of a patch are dealt out to the warps, which shortens the dependent
// chain almost k-fold).  The stopped-and-resumed trajectory is the uninterrupted one, so every
// output is unchanged.  Between launches the host reads one counter (the stream is synchronised;
// the callers of this path -- the expansion levels -- synchronise anyway).
// Measured (tools/shard_balance.py, B200, the 12 levels of bench.py's 64-view expansion; budgets
// 128 / 256 evaluations for the one- / four-warp launches, thresholds 64 and 4 patches per SM):
// one GPU 705 -> 643 ms (a stopped patch restarts at the head of the next launch, which is the
// longest-first order the view count alone cannot give), slowest of 8 shards per level, summed,
// 126 -> 110 ms.  The budget is one of WORK: a patch with more than 48 visible views gets
// budget * 48 / nvis evaluations (at least 8), so that the heavy patches of a many-view scene (C5:
// up to 256 views) reach the multi-warp launches sooner (C5 level 1, slowest of 8 shards: 32.2 ->
// 29.4 ms).  DP_SLICE_B1 / _B4 / _T4 / _T8 / _VIEWS override the knobs, DP_REFINE_SLICE=0 disables,
// DP_SLICE_TRACE=1 prints one line per launch.
// DP_REFINE_WPP=1 forces one unsliced launch with one warp per patch (A-B runs).
template <int NPASS>
static int refine_sliced(dp_context *ctx, DpRefineArgs a, const int32_t *nvis, cudaStream_t st) {
  const long long n = a.p.n;
  const char *e = getenv("DP_REFINE_WPP");
  const bool allow_mw = !(e && atoi(e) == 1) && a.p.vstride <= DP_MW_MAXV;
  const long long t8 = env_ll("DP_SLICE_T8", (long long)ctx->sm_count * 4);
  const long long t4 = env_ll("DP_SLICE_T4", (long long)ctx->sm_count * 64);
  const int b1 = (int)env_ll("DP_SLICE_B1", 128), b4 = (int)env_ll("DP_SLICE_B4", 256);
  const bool slice = allow_mw && env_ll("DP_REFINE_SLICE", 1) != 0 && n > t8;
  const bool trace = env_ll("DP_SLICE_TRACE", 0) != 0;  // per-launch lines on stderr (tuning runs)
  a.budget_views = (int)env_ll("DP_SLICE_VIEWS", 48);
  a.n_items = (unsigned int)n;
  a.nm_save = nullptr;
  a.pending = nullptr;
  a.pending_count = nullptr;
  a.budget = 0;
  a.resume = 0;
  if (!slice) {
    cudaError_t ce;
    if (allow_mw && n <= t8) ce = launch_refine_wpp<NPASS, 8>(a, ctx->sm_count, st);
    else if (allow_mw && n <= t4) ce = launch_refine_wpp<NPASS, 4>(a, ctx->sm_count, st);
    else ce = launch_refine_wpp<NPASS, 1>(a, ctx->sm_count, st);
    ++ctx->launches;
    DP_CUDA(ctx, ce);
    return DP_OK;
  }
  DP_CUDA(ctx, ctx->s_nmsave.ensure((size_t)n * DP_NM_SAVE_WORDS * sizeof(double)));
  DP_CUDA(ctx, ctx->s_pending.ensure((size_t)n + 16));
  a.nm_save = ctx->s_nmsave.as<double>();
  a.pending = ctx->s_pending.as<uint8_t>();
  // the counter sits behind the flags, 8-byte aligned
  a.pending_count = reinterpret_cast<unsigned int *>(a.pending + (((size_t)n + 7) & ~(size_t)7));
  long long m = n;
  for (int phase = 0;; ++phase) {
    int wpp = 1;
    a.budget = b1;
    if (m <= t8) { wpp = 8; a.budget = 0; }
    else if (m <= t4) { wpp = 4; a.budget = b4; }
    DP_CUDA(ctx, cudaMemsetAsync(a.work_counter, 0, sizeof(unsigned int), st));
    DP_CUDA(ctx, cudaMemsetAsync(a.pending_count, 0, sizeof(unsigned int), st));
    cudaEvent_t tr0 = nullptr, tr1 = nullptr;
    if (trace) {
      cudaEventCreate(&tr0);
      cudaEventCreate(&tr1);
      cudaEventRecord(tr0, st);
    }
    cudaError_t ce;
    if (wpp == 8) ce = launch_refine_wpp<NPASS, 8>(a, ctx->sm_count, st);
    else if (wpp == 4) ce = launch_refine_wpp<NPASS, 4>(a, ctx->sm_count, st);
    else ce = launch_refine_wpp<NPASS, 1>(a, ctx->sm_count, st);
    ++ctx->launches;
    DP_CUDA(ctx, ce);
    if (trace) {
      float ms = 0.f;
      cudaEventRecord(tr1, st);
      cudaEventSynchronize(tr1);
      cudaEventElapsedTime(&ms, tr0, tr1);
      fprintf(stderr, "refine_sliced: phase %d, %lld of %lld patches, %d warp(s) per patch, budget %d: %.3f ms\n",
              phase, m, n, wpp, a.budget, ms);
      cudaEventDestroy(tr0);
      cudaEventDestroy(tr1);
    }
    if (a.budget == 0) break;
    unsigned int left = 0;
    DP_CUDA(ctx, cudaMemcpyAsync(&left, a.pending_count, sizeof(left), cudaMemcpyDeviceToHost, st));
    DP_CUDA(ctx, cudaStreamSynchronize(st));
    if (left == 0) break;
    m = left;
    // the stopped patches (all have >= 2 views) sort in front of the finished ones (key 0)
    a.mask = a.pending;
    a.resume = 1;
    int rc = build_order(ctx, nvis, a.pending, (int)n, st, &a.order);
    if (rc != DP_OK) return rc;
    a.n_items = a.order ? (unsigned int)m : (unsigned int)n;
  }
  return DP_OK;
}

#ifndef DP_REFINE_GROUP
#define DP_REFINE_GROUP 1  // cells up to 16x16: several patches per warp (dp_group.cuh)
#endif
template <typename C>
static cudaError_t launch_refine_group(const DpRefineArgs &a, int sm_count, cudaStream_t st) {
  int per_sm = 1;
  cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, dp_refine_group_kernel<C>,
                                                                DP_GWARPS * 32, 0);
  if (e != cudaSuccess) return e;
  if (per_sm < 1) per_sm = 1;
  const long long per_cta = (long long)DP_GWARPS * C::GROUPS;
  long long want = ((long long)a.p.n + per_cta - 1) / per_cta;
  long long grid = std::min<long long>(want, (long long)sm_count * per_sm);
  dp_refine_group_kernel<C><<<(unsigned)grid, DP_GWARPS * 32, 0, st>>>(a);
  return cudaGetLastError();
}

// One patch per lane (dp_lane.cuh), cells up to 8x8.
#ifndef DP_REFINE_LANE
#define DP_REFINE_LANE 1
#endif
#ifndef DP_LANE_MIN_PATCHES
#define DP_LANE_MIN_PATCHES 131072
#endif
template <int S>
static cudaError_t launch_refine_lane(const DpRefineArgs &a, int sm_count, cudaStream_t st) {
  int per_sm = 1;
  cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, dp_refine_lane_kernel<S>,
                                                                DP_LWARPS * 32, 0);
  if (e != cudaSuccess) return e;
  if (per_sm < 1) per_sm = 1;
  const long long per_cta = (long long)DP_LWARPS * 32;
  long long want = ((long long)a.p.n + per_cta - 1) / per_cta;
  long long grid = std::min<long long>(want, (long long)sm_count * per_sm);
  dp_refine_lane_kernel<S><<<(unsigned)grid, DP_LWARPS * 32, 0, st>>>(a);
  return cudaGetLastError();
}
static bool use_lane_kernel() { return DP_REFINE_LANE && !env_is_group("DP_REFINE_KERNEL"); }
// (tests lower the threshold through the environment to run the lane kernel on small batches)
static long long lane_min_patches() {
  const char *e = getenv("DP_LANE_MIN_PATCHES");
  return e ? atoll(e) : (long long)DP_LANE_MIN_PATCHES;
}

extern "C" int dp_refine_dev(dp_context *ctx, dp_patch_dev *p, int cell_size, const uint8_t *mask,
                             int32_t *evals, double *xbest, void *stream) {
  int rc = check_patch_dev(ctx, p, cell_size);
  if (rc != DP_OK) return rc;
  if (p->n == 0) return DP_OK;
  DpDeviceGuard guard__(ctx->device);
  if ((rc = dp_sync_views(ctx)) != DP_OK) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  if ((rc = dp_scratch_acquire(ctx, st)) != DP_OK) return rc;
  DP_CUDA(ctx, ctx->work_counter.ensure(sizeof(unsigned int)));
  DP_CUDA(ctx, cudaMemsetAsync(ctx->work_counter.ptr, 0, sizeof(unsigned int), st));
  DpRefineArgs a;
  a.p = patch_args(ctx, p, cell_size);
  a.evals = evals;
  a.xbest = xbest;
  for (int j = 0; j < 3; ++j) a.step[j] = ctx->prm.nm_step[j];
  a.max_evals = ctx->prm.nm_max_evals;
  a.eps = ctx->prm.nm_eps;
  a.work_counter = ctx->work_counter.as<unsigned int>();
  a.mask = mask;
  a.n_items = (unsigned int)p->n;
  a.nm_save = nullptr;
  a.pending = nullptr;
  a.pending_count = nullptr;
  a.budget = 0;
  a.budget_views = 0;
  a.resume = 0;
#ifdef DP_DEBUG_TRACE
  {
    static double *tr = nullptr;
    if (!tr) cudaMalloc(&tr, (size_t)(1 << 20) * 64 + 64);
    std::vector<double> init((size_t)p->n * 8 + 8, -7.0);
    memset(init.data() + (size_t)p->n * 8, 0, 64);
    cudaMemcpy(tr, init.data(), init.size() * 8, cudaMemcpyHostToDevice);
    a.trace = getenv("DP_REFINE_TRACE") ? tr : nullptr;
  }
#endif
  if ((rc = build_order(ctx, p->nvis, mask, p->n, st, &a.order)) != DP_OK) return rc;
  cudaError_t e;
  // One patch per lane needs several patches per lane to keep its lanes busy (a lane runs its
  // patches one after the other, ~65 evaluations each; with fewer than DP_LANE_MIN_PATCHES the
  // tail dominates and the 4-lanes-per-patch kernel is faster).
  if (cell_size <= DP_LGROUP_MAX_CELL && use_lane_kernel() && p->n >= lane_min_patches()) {
    e = cudaSuccess;
#define DP_LCASE(S) case S: e = launch_refine_lane<S>(a, ctx->sm_count, st); break
    switch (cell_size) {
      DP_LCASE(2); DP_LCASE(3); DP_LCASE(4); DP_LCASE(5); DP_LCASE(6); DP_LCASE(7); DP_LCASE(8);
    }
#undef DP_LCASE
  } else if (DP_REFINE_GROUP && cell_size <= DP_GROUP_MAX_CELL_REFINE) {
    e = cudaSuccess;
#define DP_GCASE(S) case S: e = launch_refine_group<DpCfgFor<S>>(a, ctx->sm_count, st); break
    switch (cell_size) {
      DP_GCASE(2); DP_GCASE(3); DP_GCASE(4); DP_GCASE(5); DP_GCASE(6); DP_GCASE(7); DP_GCASE(8);
#if DP_GROUP_MAX_CELL_REFINE > 8
      DP_GCASE(9); DP_GCASE(10); DP_GCASE(11); DP_GCASE(12);
#endif
    }
#undef DP_GCASE
  } else
  {
    e = cudaSuccess;
    --ctx->launches;  // refine_sliced counts its own launches
    switch (npass_for(cell_size)) {
      case 1: rc = refine_sliced<1>(ctx, a, p->nvis, st); break;
      case 2: rc = refine_sliced<2>(ctx, a, p->nvis, st); break;
      case 4: rc = refine_sliced<4>(ctx, a, p->nvis, st); break;
      case 8: rc = refine_sliced<8>(ctx, a, p->nvis, st); break;
      case 16: rc = refine_sliced<16>(ctx, a, p->nvis, st); break;
      default: rc = refine_sliced<32>(ctx, a, p->nvis, st); break;
    }
    if (rc != DP_OK) return rc;
  }
  ++ctx->launches;
  DP_CUDA(ctx, e);
#ifdef DP_DEBUG_TRACE
  if (a.trace) {
    cudaStreamSynchronize(st);
    std::vector<double> out((size_t)p->n * 8 + 8);
    cudaMemcpy(out.data(), a.trace, out.size() * 8, cudaMemcpyDeviceToHost);
    {
      const unsigned long long *c = reinterpret_cast<const unsigned long long *>(out.data() + (size_t)p->n * 8);
      fprintf(stderr, "lane census: textured %llu empty %llu fewer-views %llu tail %llu unstaged %llu of %llu slots\n",
              c[0], c[1], c[2], c[3], c[4], c[5]);
    }
    FILE *f = fopen(getenv("DP_REFINE_TRACE"), "wb");
    if (f) { fwrite(out.data(), 8, out.size(), f); fclose(f); }
  }
#endif
  return dp_scratch_release(ctx, st);
}

extern "C" int dp_visibility_dev(dp_context *ctx, dp_patch_dev *p, int32_t *ncand, int32_t *cand,
                                 void *stream) {
  if (!ctx || !p) return DP_ERR_INVALID_ARG;
  if (p->n < 0 || p->vstride <= 0) return dp_fail(ctx, DP_ERR_INVALID_ARG, "patch batch shape");
  if (p->n == 0) return DP_OK;
  if (!p->pos || !p->nrm || !p->ref || !p->nvis || !p->vis)
    return dp_fail(ctx, DP_ERR_INVALID_ARG, "null patch array");
  DpDeviceGuard guard__(ctx->device);
  int rc = dp_sync_views(ctx);
  if (rc != DP_OK) return rc;
  const long long threads = (long long)p->n * 32;
  dp_visibility_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      ctx->d_views.as<DpViewDev>(), (int)ctx->views.size(), p->n, p->pos, p->nrm, p->ref,
      ctx->prm.visible_threshold, ctx->prm.candidate_threshold, p->nvis, p->vis, ncand, cand,
      p->vstride);
  ++ctx->launches;
  DP_CUDA(ctx, cudaGetLastError());
  return DP_OK;
}

extern "C" int dp_color_dev(dp_context *ctx, dp_patch_dev *p, void *stream) {
  if (!ctx || !p) return DP_ERR_INVALID_ARG;
  if (p->n == 0) return DP_OK;
  if (p->n < 0 || !p->pos || !p->rgb) return dp_fail(ctx, DP_ERR_INVALID_ARG, "dp_color arguments");
  DpDeviceGuard guard__(ctx->device);
  int rc = dp_sync_views(ctx);
  if (rc != DP_OK) return rc;
  const long long threads = (long long)p->n * 32;
  dp_color_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      ctx->d_views.as<DpViewDev>(), (int)ctx->views.size(), p->n, p->pos, p->rgb);
  ++ctx->launches;
  DP_CUDA(ctx, cudaGetLastError());
  return DP_OK;
}

// ---------------------------------------------------------------------------------------
// host-buffer layer: H2D -> kernel -> D2H, synchronous

struct HostBatch {
  dp_patch_dev d;
};

static int upload_patches(dp_context *ctx, const dp_patch_soa *h, dp_patch_dev *d, bool need_vis) {
  if (!h) return DP_ERR_INVALID_ARG;
  if (h->n < 0 || h->vstride <= 0) return dp_fail(ctx, DP_ERR_INVALID_ARG, "patch batch shape");
  if (h->n > 0 && (!h->pos || !h->nrm || !h->ref || !h->nvis || !h->vis))
    return dp_fail(ctx, DP_ERR_INVALID_ARG, "null patch array");
  DpDeviceGuard guard__(ctx->device);
  const size_t n = (size_t)h->n, vs = (size_t)h->vstride;
  DP_CUDA(ctx, ctx->s_pos.ensure(n * 12));
  DP_CUDA(ctx, ctx->s_nrm.ensure(n * 12));
  DP_CUDA(ctx, ctx->s_ref.ensure(n * 4));
  DP_CUDA(ctx, ctx->s_nvis.ensure(n * 4));
  DP_CUDA(ctx, ctx->s_vis.ensure(n * vs * 4));
  cudaStream_t st = ctx->stream;
  DP_CUDA(ctx, cudaMemcpyAsync(ctx->s_pos.ptr, h->pos, n * 12, cudaMemcpyHostToDevice, st));
  DP_CUDA(ctx, cudaMemcpyAsync(ctx->s_nrm.ptr, h->nrm, n * 12, cudaMemcpyHostToDevice, st));
  DP_CUDA(ctx, cudaMemcpyAsync(ctx->s_ref.ptr, h->ref, n * 4, cudaMemcpyHostToDevice, st));
  if (need_vis) {
    DP_CUDA(ctx, cudaMemcpyAsync(ctx->s_nvis.ptr, h->nvis, n * 4, cudaMemcpyHostToDevice, st));
    DP_CUDA(ctx, cudaMemcpyAsync(ctx->s_vis.ptr, h->vis, n * vs * 4, cudaMemcpyHostToDevice, st));
  }
  d->n = h->n;
  d->vstride = h->vstride;
  d->pos = ctx->s_pos.as<float>();
  d->nrm = ctx->s_nrm.as<float>();
  d->ref = ctx->s_ref.as<int32_t>();
  d->nvis = ctx->s_nvis.as<int32_t>();
  d->vis = ctx->s_vis.as<int32_t>();
  d->rgb = nullptr;
  return DP_OK;
}

extern "C" int dp_score_at(dp_context *ctx, const dp_patch_soa *h, int cell_size,
                           const double *normal, const double *position, float *ncc, uint8_t *tex,
                           uint8_t *valid) {
  if (!ctx) return DP_ERR_INVALID_ARG;
  if (!ncc) return dp_fail(ctx, DP_ERR_INVALID_ARG, "ncc is null");
  dp_patch_dev d;
  int rc = upload_patches(ctx, h, &d, true);
  if (rc != DP_OK) return rc;
  if (h->n == 0) return DP_OK;
  const size_t nv = (size_t)h->n * h->vstride;
  const size_t tb = (size_t)cell_size * cell_size * 3;
  DP_CUDA(ctx, ctx->s_ncc.ensure(nv * 4));
  if (tex) DP_CUDA(ctx, ctx->s_tex.ensure(nv * tb));
  if (valid) DP_CUDA(ctx, ctx->s_valid.ensure(nv));
  const double *d_tn = nullptr, *d_tp = nullptr;
  if (normal || position) {
    const size_t b3 = (size_t)h->n * 24;
    DP_CUDA(ctx, ctx->s_xbest.ensure(2 * b3));
    if (normal) {
      DP_CUDA(ctx, cudaMemcpyAsync(ctx->s_xbest.ptr, normal, b3, cudaMemcpyHostToDevice, ctx->stream));
      d_tn = ctx->s_xbest.as<double>();
    }
    if (position) {
      DP_CUDA(ctx, cudaMemcpyAsync(ctx->s_xbest.as<char>() + b3, position, b3, cudaMemcpyHostToDevice,
                                   ctx->stream));
      d_tp = ctx->s_xbest.as<double>() + 3 * (size_t)h->n;
    }
  }
  rc = dp_score_at_dev(ctx, &d, cell_size, d_tn, d_tp, ctx->s_ncc.as<float>(),
                       tex ? ctx->s_tex.as<uint8_t>() : nullptr,
                       valid ? ctx->s_valid.as<uint8_t>() : nullptr, ctx->stream);
  if (rc != DP_OK) return rc;
  cudaStream_t st = ctx->stream;
  DP_CUDA(ctx, cudaMemcpyAsync(ncc, ctx->s_ncc.ptr, nv * 4, cudaMemcpyDeviceToHost, st));
  if (tex) DP_CUDA(ctx, cudaMemcpyAsync(tex, ctx->s_tex.ptr, nv * tb, cudaMemcpyDeviceToHost, st));
  if (valid) DP_CUDA(ctx, cudaMemcpyAsync(valid, ctx->s_valid.ptr, nv, cudaMemcpyDeviceToHost, st));
  DP_CUDA(ctx, cudaStreamSynchronize(st));
  return DP_OK;
}

extern "C" int dp_score(dp_context *ctx, const dp_patch_soa *h, int cell_size, float *ncc,
                        uint8_t *tex, uint8_t *valid) {
  return dp_score_at(ctx, h, cell_size, nullptr, nullptr, ncc, tex, valid);
}

extern "C" int dp_filter(dp_context *ctx, dp_patch_soa *h, int cell_size, uint8_t *keep) {
  if (!ctx) return DP_ERR_INVALID_ARG;
  if (!keep) return dp_fail(ctx, DP_ERR_INVALID_ARG, "keep is null");
  dp_patch_dev d;
  int rc = upload_patches(ctx, h, &d, true);
  if (rc != DP_OK) return rc;
  if (h->n == 0) return DP_OK;
  const size_t n = (size_t)h->n, vs = (size_t)h->vstride;
  DP_CUDA(ctx, ctx->s_keep.ensure(n));
  rc = dp_filter_dev(ctx, &d, cell_size, ctx->s_keep.as<uint8_t>(), ctx->stream);
  if (rc != DP_OK) return rc;
  cudaStream_t st = ctx->stream;
  DP_CUDA(ctx, cudaMemcpyAsync(keep, ctx->s_keep.ptr, n, cudaMemcpyDeviceToHost, st));
  DP_CUDA(ctx, cudaMemcpyAsync(h->nvis, d.nvis, n * 4, cudaMemcpyDeviceToHost, st));
  DP_CUDA(ctx, cudaMemcpyAsync(h->vis, d.vis, n * vs * 4, cudaMemcpyDeviceToHost, st));
  DP_CUDA(ctx, cudaStreamSynchronize(st));
  return DP_OK;
}

extern "C" int dp_refine(dp_context *ctx, dp_patch_soa *h, int cell_size, const uint8_t *mask,
                         int32_t *evals, double *xbest) {
  if (!ctx) return DP_ERR_INVALID_ARG;
  dp_patch_dev d;
  int rc = upload_patches(ctx, h, &d, true);
  if (rc != DP_OK) return rc;
  if (h->n == 0) return DP_OK;
  const size_t n = (size_t)h->n;
  if (evals) DP_CUDA(ctx, ctx->s_evals.ensure(n * 4));
  if (xbest) DP_CUDA(ctx, ctx->s_xbest.ensure(n * 24));
  if (mask) {
    DP_CUDA(ctx, ctx->s_keep.ensure(n));
    DP_CUDA(ctx, cudaMemcpyAsync(ctx->s_keep.ptr, mask, n, cudaMemcpyHostToDevice, ctx->stream));
  }
  rc = dp_refine_dev(ctx, &d, cell_size, mask ? ctx->s_keep.as<uint8_t>() : nullptr,
                     evals ? ctx->s_evals.as<int32_t>() : nullptr,
                     xbest ? ctx->s_xbest.as<double>() : nullptr, ctx->stream);
  if (rc != DP_OK) return rc;
  cudaStream_t st = ctx->stream;
  DP_CUDA(ctx, cudaMemcpyAsync(h->pos, d.pos, n * 12, cudaMemcpyDeviceToHost, st));
  DP_CUDA(ctx, cudaMemcpyAsync(h->nrm, d.nrm, n * 12, cudaMemcpyDeviceToHost, st));
  if (evals) DP_CUDA(ctx, cudaMemcpyAsync(evals, ctx->s_evals.ptr, n * 4, cudaMemcpyDeviceToHost, st));
  if (xbest) DP_CUDA(ctx, cudaMemcpyAsync(xbest, ctx->s_xbest.ptr, n * 24, cudaMemcpyDeviceToHost, st));
  DP_CUDA(ctx, cudaStreamSynchronize(st));
  return DP_OK;
}

// Seed::OptimizeAndRefinePatches in one call, as a three-stream pipeline for pinned host
// arrays: the batch is uploaded in chunks and each chunk is filtered as soon as it has landed
// (the filter kernels are short and have no tail); the refinement then runs as ONE launch over
// the whole batch -- its persistent warps end with a tail as long as the slowest patch, so
// per-chunk refinement (measured: 8 chunks, -27 %) loses more than the overlap wins -- while
// the filter's outputs (keep bits, visible sets: 2/3 of the download, read-only for the refine
// kernel) are already on their way back; the refined geometry follows.
#ifndef DP_PIPE_CHUNK
#define DP_PIPE_CHUNK 131072
#endif
extern "C" int dp_filter_refine(dp_context *ctx, dp_patch_soa *h, int cell_size, uint8_t *keep,
                                int32_t *evals) {
  if (!ctx) return DP_ERR_INVALID_ARG;
  if (!keep) return dp_fail(ctx, DP_ERR_INVALID_ARG, "keep is null");
  if (!h) return DP_ERR_INVALID_ARG;
  if (h->n < 0 || h->vstride <= 0) return dp_fail(ctx, DP_ERR_INVALID_ARG, "patch batch shape");
  if (h->n == 0) return DP_OK;
  if (!h->pos || !h->nrm || !h->ref || !h->nvis || !h->vis)
    return dp_fail(ctx, DP_ERR_INVALID_ARG, "null patch array");
  int rc = check_patch_dev_args(ctx, cell_size);
  if (rc != DP_OK) return rc;
  DpDeviceGuard guard__(ctx->device);
  if ((rc = dp_sync_views(ctx)) != DP_OK) return rc;
  const size_t n = (size_t)h->n, vs = (size_t)h->vstride;
  DP_CUDA(ctx, ctx->s_pos.ensure(n * 12));
  DP_CUDA(ctx, ctx->s_nrm.ensure(n * 12));
  DP_CUDA(ctx, ctx->s_ref.ensure(n * 4));
  DP_CUDA(ctx, ctx->s_nvis.ensure(n * 4));
  DP_CUDA(ctx, ctx->s_vis.ensure(n * vs * 4));
  DP_CUDA(ctx, ctx->s_keep.ensure(n));
  DP_CUDA(ctx, ctx->s_evals.ensure(n * 4));
  const size_t chunk = std::min<size_t>(n, DP_PIPE_CHUNK);
  // scratch of the per-chunk launches, sized once so that no reallocation (= device-wide
  // synchronisation) happens while the pipeline runs
  DP_CUDA(ctx, ctx->s_order.ensure(n * 4 + DP_ORDER_BINS * 4));
  DP_CUDA(ctx, ctx->work_counter.ensure(sizeof(unsigned int)));
  if (!ctx->stream_in) {
    DP_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->stream_in, cudaStreamNonBlocking));
    DP_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->stream_out, cudaStreamNonBlocking));
  }
  const size_t n_chunks = (n + chunk - 1) / chunk;
  while (ctx->pipe_events.size() < std::max<size_t>(2 * n_chunks, (size_t)4)) {
    cudaEvent_t e;
    DP_CUDA(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    ctx->pipe_events.push_back(e);
  }
  cudaStream_t s_in = ctx->stream_in, s_cmp = ctx->stream, s_out = ctx->stream_out;
  float *d_pos = ctx->s_pos.as<float>(), *d_nrm = ctx->s_nrm.as<float>();
  int32_t *d_ref = ctx->s_ref.as<int32_t>(), *d_nvis = ctx->s_nvis.as<int32_t>(),
          *d_vis = ctx->s_vis.as<int32_t>(), *d_evals = ctx->s_evals.as<int32_t>();
  uint8_t *d_keep = ctx->s_keep.as<uint8_t>();
  // Every failure below funnels to the common exit: no copy into or out of the caller's arrays
  // may still be in flight when this function returns.
  auto body = [&]() -> int {
    for (size_t c = 0; c < n_chunks; ++c) {
      const size_t o = c * chunk, m = std::min(chunk, n - o);
      DP_CUDA(ctx, cudaMemcpyAsync(d_pos + 3 * o, h->pos + 3 * o, m * 12, cudaMemcpyHostToDevice, s_in));
      DP_CUDA(ctx, cudaMemcpyAsync(d_nrm + 3 * o, h->nrm + 3 * o, m * 12, cudaMemcpyHostToDevice, s_in));
      DP_CUDA(ctx, cudaMemcpyAsync(d_ref + o, h->ref + o, m * 4, cudaMemcpyHostToDevice, s_in));
      DP_CUDA(ctx, cudaMemcpyAsync(d_nvis + o, h->nvis + o, m * 4, cudaMemcpyHostToDevice, s_in));
      DP_CUDA(ctx, cudaMemcpyAsync(d_vis + o * vs, h->vis + o * vs, m * vs * 4, cudaMemcpyHostToDevice, s_in));
      DP_CUDA(ctx, cudaEventRecord(ctx->pipe_events[2 * c], s_in));
    }
    dp_patch_dev d;
    d.vstride = h->vstride;
    d.rgb = nullptr;
    for (size_t c = 0; c < n_chunks; ++c) {
      const size_t o = c * chunk, m = std::min(chunk, n - o);
      d.n = (int32_t)m;
      d.pos = d_pos + 3 * o;
      d.nrm = d_nrm + 3 * o;
      d.ref = d_ref + o;
      d.nvis = d_nvis + o;
      d.vis = d_vis + o * vs;
      DP_CUDA(ctx, cudaStreamWaitEvent(s_cmp, ctx->pipe_events[2 * c], 0));
      const int frc = dp_filter_dev(ctx, &d, cell_size, d_keep + o, s_cmp);
      if (frc != DP_OK) return frc;
    }
    DP_CUDA(ctx, cudaEventRecord(ctx->pipe_events[1], s_cmp));  // filter done
    DP_CUDA(ctx, cudaStreamWaitEvent(s_out, ctx->pipe_events[1], 0));
    DP_CUDA(ctx, cudaMemcpyAsync(keep, d_keep, n, cudaMemcpyDeviceToHost, s_out));
    DP_CUDA(ctx, cudaMemcpyAsync(h->nvis, d_nvis, n * 4, cudaMemcpyDeviceToHost, s_out));
    DP_CUDA(ctx, cudaMemcpyAsync(h->vis, d_vis, n * vs * 4, cudaMemcpyDeviceToHost, s_out));
    d.n = h->n;
    d.pos = d_pos;
    d.nrm = d_nrm;
    d.ref = d_ref;
    d.nvis = d_nvis;
    d.vis = d_vis;
    const int rrc = dp_refine_dev(ctx, &d, cell_size, d_keep, evals ? d_evals : nullptr, nullptr, s_cmp);
    if (rrc != DP_OK) return rrc;
    DP_CUDA(ctx, cudaEventRecord(ctx->pipe_events[3], s_cmp));  // refine done
    DP_CUDA(ctx, cudaStreamWaitEvent(s_out, ctx->pipe_events[3], 0));
    DP_CUDA(ctx, cudaMemcpyAsync(h->pos, d_pos, n * 12, cudaMemcpyDeviceToHost, s_out));
    DP_CUDA(ctx, cudaMemcpyAsync(h->nrm, d_nrm, n * 12, cudaMemcpyDeviceToHost, s_out));
    if (evals) DP_CUDA(ctx, cudaMemcpyAsync(evals, d_evals, n * 4, cudaMemcpyDeviceToHost, s_out));
    return DP_OK;
  };
  const int first_rc = body();
  const std::string first_err = ctx->err;
  const cudaError_t e1 = cudaStreamSynchronize(s_in), e2 = cudaStreamSynchronize(s_cmp),
                    e3 = cudaStreamSynchronize(s_out);
  if (first_rc != DP_OK) {
    ctx->err = first_err;
    return first_rc;
  }
  DP_CUDA(ctx, e1);
  DP_CUDA(ctx, e2);
  DP_CUDA(ctx, e3);
  return DP_OK;
}

extern "C" int dp_visibility(dp_context *ctx, dp_patch_soa *h, int32_t *ncand, int32_t *cand) {
  if (!ctx) return DP_ERR_INVALID_ARG;
  dp_patch_dev d;
  int rc = upload_patches(ctx, h, &d, false);
  if (rc != DP_OK) return rc;
  if (h->n == 0) return DP_OK;
  const size_t n = (size_t)h->n, vs = (size_t)h->vstride;
  if (ncand) DP_CUDA(ctx, ctx->s_ncand.ensure(n * 4));
  if (cand) DP_CUDA(ctx, ctx->s_cand.ensure(n * vs * 4));
  rc = dp_visibility_dev(ctx, &d, ncand ? ctx->s_ncand.as<int32_t>() : nullptr,
                         cand ? ctx->s_cand.as<int32_t>() : nullptr, ctx->stream);
  if (rc != DP_OK) return rc;
  cudaStream_t st = ctx->stream;
  DP_CUDA(ctx, cudaMemcpyAsync(h->nvis, d.nvis, n * 4, cudaMemcpyDeviceToHost, st));
  DP_CUDA(ctx, cudaMemcpyAsync(h->vis, d.vis, n * vs * 4, cudaMemcpyDeviceToHost, st));
  if (ncand) DP_CUDA(ctx, cudaMemcpyAsync(ncand, ctx->s_ncand.ptr, n * 4, cudaMemcpyDeviceToHost, st));
  if (cand) DP_CUDA(ctx, cudaMemcpyAsync(cand, ctx->s_cand.ptr, n * vs * 4, cudaMemcpyDeviceToHost, st));
  DP_CUDA(ctx, cudaStreamSynchronize(st));
  return DP_OK;
}

extern "C" int dp_color(dp_context *ctx, dp_patch_soa *h) {
  if (!ctx || !h) return DP_ERR_INVALID_ARG;
  if (h->n < 0 || (h->n > 0 && (!h->pos || !h->rgb)))
    return dp_fail(ctx, DP_ERR_INVALID_ARG, "dp_color arguments");
  if (h->n == 0) return DP_OK;
  DpDeviceGuard guard__(ctx->device);
  const size_t n = (size_t)h->n;
  DP_CUDA(ctx, ctx->s_pos.ensure(n * 12));
  DP_CUDA(ctx, ctx->s_rgb.ensure(n * 3));
  cudaStream_t st = ctx->stream;
  DP_CUDA(ctx, cudaMemcpyAsync(ctx->s_pos.ptr, h->pos, n * 12, cudaMemcpyHostToDevice, st));
  dp_patch_dev d;
  memset(&d, 0, sizeof(d));
  d.n = h->n;
  d.vstride = 1;
  d.pos = ctx->s_pos.as<float>();
  d.rgb = ctx->s_rgb.as<uint8_t>();
  int rc = dp_color_dev(ctx, &d, st);
  if (rc != DP_OK) return rc;
  DP_CUDA(ctx, cudaMemcpyAsync(h->rgb, d.rgb, n * 3, cudaMemcpyDeviceToHost, st));
  DP_CUDA(ctx, cudaStreamSynchronize(st));
  return DP_OK;
}

#include "dp_expand.cuh"
#include "dp_pyramid.cuh"
#include "dp_seed.cuh"
