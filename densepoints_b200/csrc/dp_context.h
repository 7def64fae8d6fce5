// dp_context.h -- host-side state behind the opaque dp_context handle.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/densepoints_cuda.h"
#include "dp_device.cuh"

struct DpDevBuf {  // grow-only device scratch
  void *ptr = nullptr;
  size_t cap = 0;
  cudaError_t ensure(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (ptr) cudaFree(ptr);
    ptr = nullptr;
    cap = 0;
    size_t want = bytes + bytes / 4 + 256;
    cudaError_t e = cudaMalloc(&ptr, want);
    if (e == cudaSuccess) cap = want;
    return e;
  }
  void release() {
    if (ptr) cudaFree(ptr);
    ptr = nullptr;
    cap = 0;
  }
  template <typename T>
  T *as() const { return reinterpret_cast<T *>(ptr); }
};

// Every entry point runs on the context's device and leaves the caller's current device as it
// found it.
struct DpDeviceGuard {
  int prev = -1;
  explicit DpDeviceGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    if (prev != dev) cudaSetDevice(dev);
    else prev = -1;
  }
  ~DpDeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

struct DpLevel {  // one pyramid level of one view
  uint32_t *img = nullptr;
  int width = 0, height = 0, pitch_px = 0;
};

struct DpViewHost {
  bool set = false;
  double P[12];
  double xaxis[3];   // View::GetXAxis() (not normalised)
  double center[3];
  std::vector<DpLevel> levels;  // [0] = the uploaded image
};

// Device-resident PatchOrganizer: replicated occupancy grids + patch store.
struct DpOrganizer {
  bool ready = false;
  int level = 0;               // pyramid level the grids were allocated for
  long long n_cells = 0;
  DpDevBuf grid;               // u8 occupancy count per cell, all views concatenated
  DpDevBuf claim;              // u32 min sequence id claiming each cell in the current round
  // store (SoA), capacity `cap` patches; vstride = n_views; vis = the visible sets as bit
  // masks, ceil(n_views / 32) u32 words per patch (ascending view order is implicit)
  long long n = 0, cap = 0;
  int vstride = 0;
  DpDevBuf pos, nrm, rgb, ref, nvis, vis;
  long long frontier_begin = 0;  // store index where the current BFS level starts
  long long pops = 0;
};

struct dp_context {
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t stream_in = nullptr, stream_out = nullptr;  // copy streams of dp_filter_refine's pipeline
  std::vector<cudaEvent_t> pipe_events;
  dp_params prm;
  std::vector<DpViewHost> views;
  DpDevBuf d_views;      // DpViewDev[n_views] of the active level
  bool views_dirty = true;
  int level = 0;
  int n_levels = 1;
  bool auto_level = false;       // dp_set_level_selection: per-(patch, view) pyramid level
  double level_px = 1.5;         // ... a texel may cover this many pixels of the level it is read from
  DpDevBuf d_views_lv;           // DpViewDev[n_levels - level][n_views]: the base level and the ones above
  std::string err;
  int64_t launches = 0;
  // scratch for the host-buffer API
  DpDevBuf s_pos, s_nrm, s_ref, s_nvis, s_vis, s_rgb, s_ncc, s_tex, s_valid, s_keep, s_evals,
      s_xbest, s_cand, s_ncand, s_img, s_misc;
  DpDevBuf work_counter, s_order;
  DpDevBuf s_nmsave, s_pending;  // time-sliced refinement: stopped patches' solver state, flags
  // The dp_*_dev calls share the scratch above and below (work counter, order table, expansion
  // buffers).  Calls on different streams are therefore chained: a call first makes its stream
  // wait for the event the previous call recorded on its own stream (dp_scratch_acquire /
  // dp_scratch_release), so two asynchronous calls never use the scratch at the same time.
  cudaEvent_t scratch_event = nullptr;
  cudaStream_t scratch_stream = nullptr;
  bool scratch_busy = false;
  // expansion scratch
  DpDevBuf e_pos, e_nrm, e_ref, e_nvis, e_vis, e_keep, e_seq, e_cells, e_recs, e_flags, e_scan, e_count, e_won;
  DpOrganizer org;
  long long org_last_candidates = 0;
  int sm_count = 148;
};

// internal helpers shared by the translation units
int dp_fail(dp_context *ctx, int code, const char *what, cudaError_t e = cudaSuccess);
int dp_sync_views(dp_context *ctx);
int dp_scratch_acquire(dp_context *ctx, cudaStream_t st);
int dp_scratch_release(dp_context *ctx, cudaStream_t st);
// (dp_sync_views (re)builds the DpViewDev table for the active level)
#define DP_CUDA(ctx, call)                                             \
  do {                                                                 \
    cudaError_t e__ = (call);                                          \
    if (e__ != cudaSuccess) return dp_fail((ctx), DP_ERR_CUDA, #call, e__); \
  } while (0)
