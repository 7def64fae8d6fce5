// dp_kernels.cuh -- the score / filter / refine kernels (sm_100a), one warp per patch.
//
//   dp_score_kernel   K1+K2: GetProjectedTextures + NCCScore for every visible view
//                     (optimization.cpp:14-56, 104-110) and, fused as an epilogue,
//                     FilterByErrorMeasurement's erase loop (optimization.cpp:112-131).
//   dp_refine_kernel  K3: OptimizationOpenCV::Optimize (optimization_opencv.cpp:44-78):
//                     cv::DownhillSolver's Nelder-Mead over (depth, roll, pitch), every
//                     objective evaluation = K1 fused; persistent warps pull patches
//                     from an atomic work counter (evaluation counts vary 4..500).
#pragma once
#include "dp_device.cuh"

#define DP_WARPS 8  // warps (= patches in flight) per CTA
// Resident CTAs per SM requested through __launch_bounds__ (= register cap), measured on
// B200: the score kernel is fastest at 4 CTAs/SM (64 registers) for <= 2 texel passes, the
// refine kernel at 2 CTAs/SM (128 registers; 3 CTAs spill 600+ bytes and lose 7 %).
__host__ __device__ constexpr int dp_score_min_ctas(int npass) {
  return npass <= 2 ? 4 : (npass <= 4 ? 3 : (npass <= 8 ? 2 : 1));
}
#ifndef DP_RWARPS
#define DP_RWARPS 8  // warps per CTA of the refine kernel
#endif
#ifndef DP_RMINCTA
#define DP_RMINCTA 2
#endif
__host__ __device__ constexpr int dp_refine_min_ctas(int npass) { return npass <= 8 ? DP_RMINCTA : 1; }

struct DpPatchArgs {
  const DpViewDev *views;
  int n_views;
  int n, vstride;
  float *pos, *nrm;
  int32_t *ref, *nvis, *vis;
  int s;
  DpLevelSel lv;  // per-(patch, view) pyramid level, or tab == null
};

struct DpScoreArgs {
  DpPatchArgs p;
  float *ncc;      // n*vstride or null
  uint8_t *tex;    // n*vstride*s*s*3 or null
  uint8_t *valid;  // n*vstride or null
  // GetProjectedTextures(normal, position, textures): optional trial arguments, n*3 fp64 each
  // (null = the patch's own); the stored pos stays the corner centre (patch.cpp:119-123)
  const double *trial_nrm, *trial_pos;
  // filter epilogue
  double thr;
  int min_visible;
  uint8_t *keep;
};

// How the texel pass gets its source pixels is a per-kernel choice, measured on B200
// (profiles/r01_summary.md):
//  * score / filter kernel (STAGE = true): every patch is visited once, nothing is reused, so
//    the warp stages the ROI into one shared-memory buffer of kTilePx pixels (~4x the texel
//    count, for oblique / zoomed views) or, beyond that, not at all.
//  * refine kernel (STAGE = false): consecutive Nelder-Mead evaluations of a patch read almost
//    the same pixels, 97 % of the loads hit L1, and staging them again for every evaluation
//    costs more than it saves: the taps are gathered straight from global memory (+16 % over
//    staging; cp.async double buffering -1 %, tld4 texture gather +2 %).
// (Round 1 also staged footprints that fit a 16x16 box by TMA -- cp.async.bulk.tensor.2d from a
// per-view tensor map, mbarrier, two buffers.  Measured and retired: the TMA warp-per-patch
// score kernel reached 1.74 G evals/s against 2.62 for the cp.async group kernel that now
// serves those cell sizes, and TMA staging in the refine kernel lost 3 %; the probe that found
// the 16-byte alignment rule of the tile coordinate is kept in tools/probe/.)
template <int NPASS, bool STAGE>
struct DpTileCfg {
  static constexpr bool kDirect = !STAGE;
  static constexpr int kTilePx = kDirect ? 1 : (128 * NPASS < 768 ? 128 * NPASS : 768);
};

// Per-warp shared memory of the score / refine kernels.
template <int NPASS, bool STAGE>
struct __align__(128) DpWarpShared {
  uint32_t tile[DpTileCfg<NPASS, STAGE>::kTilePx];
  DpViewSetup recs[DP_ROUND];
};

// Evaluate all visible views of one patch at (n, p), DP_ROUND views per round:
//   phase A  batched set-up of the round's views            (dp_setup_views)
//   phase B  per view: stage ROI, warp the texels, integer moments, NCC numerator
//   phase C  one view per lane: NCCScore(texture 0, texture k) from the moments
// After each round sink(k0, kc, score) is called with lane l holding the score of view
// k0 + l (l < kc; -1 when either texture is empty, error_measurements.cpp:38-40; the entry
// of view 0 is meaningless).
template <int NPASS, bool WRITE_TEX, bool STAGE, typename Sink>
__device__ __forceinline__ void dp_eval_views(const DpViewDev *__restrict__ views, int n_views,
                                              const DpLevelSel &lv,
                                              int ref, const int32_t *vis, int nv, int s, int npx,
                                              const double n[3], const double p[3],
                                              const double pc[3],
                                              const DpTexels<NPASS> &tx, DpWarpShared<NPASS, STAGE> &ws,
                                              int lane, uint8_t *tex_base,
                                              uint8_t *valid_base, Sink sink) {
  DpViewSetup *recs = ws.recs;
  DpFrame f;
  if (ref >= 0 && ref < n_views)
    dp_make_frame(views + ref, s, n, p, pc, f);
  else
    f.ok = false;
  const double scale = 1.0 / (double)npx;  // cv::meanStdDev: mean = sum * (1/N)
  float da[NPASS];                          // centred anchor texels (texture 0)
  unsigned a1 = 0, a2 = 0;
  bool a_ok = false;
#pragma unroll 1
  for (int k0 = 0; k0 < nv; k0 += DP_ROUND) {
    const int kc = min(DP_ROUND, nv - k0);
    __syncwarp();
    dp_setup_views<32>(views, n_views, lv, vis + k0, kc, kc, s, f, recs, lane);
    __syncwarp();
    unsigned my1 = 0, my2 = 0;
    double mynum = 0.0;
    int myok = 0;
#pragma unroll 1
    for (int l = 0; l < kc; ++l) {
      const DpViewSetup &R = recs[l];
      const bool ok = R.ok != 0;  // warp-uniform
      unsigned s1 = 0, s2 = 0;
      double num = 0.0;
      if (ok) {
        bool staged = false;
        if (!DpTileCfg<NPASS, STAGE>::kDirect)
          staged = dp_stage_roi<(NPASS < 4 ? NPASS : 4)>(R, ws.tile, DpTileCfg<NPASS, STAGE>::kTilePx,
                                                         lane);
        int g[NPASS];
        uint8_t *tex_out = WRITE_TEX ? tex_base + (size_t)(k0 + l) * npx * 3 : nullptr;
        // two instantiations, so neither texel loop carries the other's addressing
        if (staged)
          dp_view_texture<NPASS, WRITE_TEX, true>(R, npx, tx, ws.tile, lane, g, tex_out);
        else
          dp_view_texture<NPASS, WRITE_TEX, false>(R, npx, tx, ws.tile, lane, g, tex_out);
        dp_moments<NPASS>(g, s1, s2);
        if (k0 + l == 0) {
          a1 = s1;
          a2 = s2;
          a_ok = true;
          dp_centre<NPASS>(g, s1, scale, npx, lane, da);
        } else if (a_ok) {
          float db[NPASS];
          dp_centre<NPASS>(g, s1, scale, npx, lane, db);
#pragma unroll
          for (int j = 0; j < NPASS; ++j) num = xadd(num, xmul((double)da[j], (double)db[j]));
          num = warp_sum_f64(num);
        }
      }
      if (lane == l) {
        my1 = s1;
        my2 = s2;
        mynum = num;
        myok = ok ? 1 : 0;
      }
    }
    double score = -1.0;
    if (lane < kc && k0 + lane >= 1 && myok && a_ok)
      score = dp_ncc_finish(a1, a2, my1, my2, mynum, scale, npx);
    if (valid_base != nullptr && lane < kc) valid_base[k0 + lane] = (uint8_t)myok;
    sink(k0, kc, score);
  }
}

template <int NPASS, bool WRITE_TEX, bool FILTER>
__global__ void __launch_bounds__(DP_WARPS * 32, dp_score_min_ctas(NPASS)) dp_score_kernel(DpScoreArgs a) {
  __shared__ DpWarpShared<NPASS, true> wsh[DP_WARPS];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long i = (long long)blockIdx.x * DP_WARPS + warp;
  if (i >= a.p.n) return;
  DpWarpShared<NPASS, true> &ws = wsh[warp];
  const int s = a.p.s, npx = s * s;
  DpTexels<NPASS> tx;
  tx.init(s, lane);
  const int nv = min(a.p.nvis[i], a.p.vstride);
  const int ref = a.p.ref[i];
  double n[3] = {(double)a.p.nrm[3 * i], (double)a.p.nrm[3 * i + 1], (double)a.p.nrm[3 * i + 2]};
  double p[3] = {(double)a.p.pos[3 * i], (double)a.p.pos[3 * i + 1], (double)a.p.pos[3 * i + 2]};
  const double pc[3] = {p[0], p[1], p[2]};
  if (a.trial_nrm) { n[0] = a.trial_nrm[3 * i]; n[1] = a.trial_nrm[3 * i + 1]; n[2] = a.trial_nrm[3 * i + 2]; }
  if (a.trial_pos) { p[0] = a.trial_pos[3 * i]; p[1] = a.trial_pos[3 * i + 1]; p[2] = a.trial_pos[3 * i + 2]; }
  int32_t *vis = a.p.vis + (size_t)i * a.p.vstride;
  float *ncc = a.ncc ? a.ncc + (size_t)i * a.p.vstride : nullptr;
  uint8_t *tex = WRITE_TEX ? a.tex + (size_t)i * a.p.vstride * npx * 3 : nullptr;
  uint8_t *valid = a.valid ? a.valid + (size_t)i * a.p.vstride : nullptr;
  // FilterByErrorMeasurement's erase loop (optimization.cpp:117-124) erases
  // visible[i - removed] when scores[i] (the score of visible[i+1]) is low; since every
  // erased entry lies before the cursor this is "drop original entry k-1 iff the score of
  // entry k is low", and the last entry always survives.  Evaluated round by round with a
  // ballot compaction, in place: a round only writes positions below the entries that later
  // rounds still have to read.
  int wcur = 0;
  const double thr = a.thr;
  const unsigned lt = (1u << lane) - 1u;
  dp_eval_views<NPASS, WRITE_TEX, true>(
      a.p.views, a.p.n_views, a.p.lv, ref, vis, nv, s, npx, n, p, pc, tx, ws, lane, tex, valid,
      [&](int k0, int kc, double score) {
        const int k = k0 + lane;
        const bool mine = lane < kc && k >= 1;
        if (ncc != nullptr && mine) ncc[k] = (float)score;
        if (FILTER) {
          const bool keepf = mine && !(score < thr);
          const int prev = mine ? vis[k - 1] : -1;
          const unsigned m = __ballot_sync(DP_FULL, keepf);
          __syncwarp();
          if (keepf) vis[wcur + __popc(m & lt)] = prev;
          wcur += __popc(m);
          __syncwarp();
        }
      });
  if (FILTER) {
    bool kept = false;
    if (nv >= 2) {  // scores.size() > 0 (optimization.cpp:113)
      if (lane == 0) vis[wcur] = vis[nv - 1];
      ++wcur;
      __syncwarp();
      for (int k = wcur + lane; k < nv; k += 32) vis[k] = -1;
      if (lane == 0) a.p.nvis[i] = wcur;
      kept = wcur >= a.min_visible;  // optimization.cpp:127
    }
    if (lane == 0) a.keep[i] = kept ? 1 : 0;
  }
}

// ------------------------------------------------------------------------------------
// K3 refine

struct DpRefineArgs {
  DpPatchArgs p;
  int32_t *evals;  // n or null
  double *xbest;   // n*3 or null
  double step[3];
  int max_evals;
  double eps;
  unsigned int *work_counter;  // zeroed before launch
  const uint8_t *mask;         // optional: refine only patches with mask[i] != 0
  const int32_t *order;        // optional: work item k = patch order[k] (longest-first schedule)
  // Time slicing (dp_refine_kernel; see refine_sliced in densepoints_cuda.cu): a launch stops a
  // patch after `budget` objective evaluations and saves its Nelder-Mead state bit for bit; a
  // later launch with resume = 1 (and mask = pending) picks it up where it stopped, so the
  // trajectory -- and every output -- is the one of an uninterrupted run.
  double *nm_save;             // n * DP_NM_SAVE_WORDS doubles, or null (no slicing)
  uint8_t *pending;            // n flags, written by every launch: 1 = stopped, to be continued
  unsigned int *pending_count; // number of patches this launch left pending (zeroed before)
  int budget;                  // evaluations per patch in this launch; 0 = unlimited
  int budget_views;            // a patch with more visible views than this gets budget * budget_views / nvis
                               // evaluations (>= 8): the budget is one of WORK, heavy patches move on sooner
  int resume;                  // 1: the patches with mask != 0 continue from nm_save
  unsigned int n_items;        // work items dp_refine_kernel hands out (<= n; the first entries of order)
#ifdef DP_DEBUG_TRACE
  double *trace;               // debug builds: objective value of the first 8 evaluations, n*8
#endif
};

// Longest-processing-time-first schedule for the persistent refine warps: the cost of a
// patch is (evaluations x visible views); the view count is known up front, so patches are
// handed out in descending view count (counting sort, 3 tiny kernels).  Results do not
// depend on the order in which patches are processed.
#define DP_ORDER_BINS 257
__device__ __forceinline__ int dp_order_key(const int32_t *__restrict__ nvis,
                                            const uint8_t *__restrict__ mask, int i) {
  return (mask && mask[i] == 0) ? 0 : min(max(nvis[i], 0), DP_ORDER_BINS - 1);
}
// A few bins receive almost all patches, so the global atomics are aggregated first: per CTA in
// shared memory for the histogram, per group of equal keys in a warp for the scatter.
__global__ void __launch_bounds__(256)
dp_order_hist_kernel(const int32_t *__restrict__ nvis, const uint8_t *__restrict__ mask, int n,
                     unsigned int *__restrict__ hist) {
  __shared__ unsigned int h[DP_ORDER_BINS];
  for (int b = threadIdx.x; b < DP_ORDER_BINS; b += blockDim.x) h[b] = 0;
  __syncthreads();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) atomicAdd(h + (DP_ORDER_BINS - 1 - dp_order_key(nvis, mask, i)), 1u);  // bin 0 = most views
  __syncthreads();
  for (int b = threadIdx.x; b < DP_ORDER_BINS; b += blockDim.x)
    if (h[b]) atomicAdd(hist + b, h[b]);
}
__global__ void dp_order_scan_kernel(unsigned int *hist) {  // exclusive scan in place, 1 thread
  unsigned int run = 0;
  for (int b = 0; b < DP_ORDER_BINS; ++b) {
    const unsigned int c = hist[b];
    hist[b] = run;
    run += c;
  }
}
__global__ void __launch_bounds__(256)
dp_order_scatter_kernel(const int32_t *__restrict__ nvis, const uint8_t *__restrict__ mask, int n,
                        unsigned int *__restrict__ cursor, int32_t *__restrict__ order) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  const bool in = i < n;
  const int bin = in ? DP_ORDER_BINS - 1 - dp_order_key(nvis, mask, i) : -1;
  const unsigned peers = __match_any_sync(DP_FULL, bin);  // lanes with my bin
  const int leader = __ffs(peers) - 1;
  unsigned int base = 0;
  if (in && lane == leader) base = atomicAdd(cursor + bin, (unsigned)__popc(peers));
  base = __shfl_sync(DP_FULL, base, leader);
  if (in) order[base + __popc(peers & ((1u << lane) - 1u))] = i;
}

// Optimization::UnparametrizePatch (optimization.cpp:78-96)
__device__ __forceinline__ void dp_unparametrize(const double *__restrict__ C, const double n0[3],
                                                 const double p0[3], double depth, double roll,
                                                 double pitch, double n[3], double p[3]) {
  const double k = xadd(1.0, depth);
#pragma unroll
  for (int j = 0; j < 3; ++j) p[j] = xadd(C[j], xmul(k, xsub(p0[j], C[j])));
  // one sincos instruction stream serves both angles: even lanes take roll, odd lanes pitch
  const int lane = threadIdx.x & 31;
  double sv, cv;
  sincos((lane & 1) ? pitch : roll, &sv, &cv);
  const double sa = __shfl_sync(DP_FULL, sv, 0), ca = __shfl_sync(DP_FULL, cv, 0);
  const double sb = __shfl_sync(DP_FULL, sv, 1), cb = __shfl_sync(DP_FULL, cv, 1);
  // rotation rows: [cb 0 -sb; sa*sb ca cb*sa; ca*sb -sa ca*cb]
  n[0] = xadd(xmul(cb, n0[0]), xmul(-sb, n0[2]));
  n[1] = xadd(xadd(xmul(xmul(sa, sb), n0[0]), xmul(ca, n0[1])), xmul(xmul(cb, sa), n0[2]));
  n[2] = xadd(xadd(xmul(xmul(ca, sb), n0[0]), xmul(-sa, n0[1])), xmul(xmul(ca, cb), n0[2]));
}

// Nelder-Mead state of one patch (cv::DownhillSolver, ndim = 3), one instance per warp in
// shared memory.  All lanes of the warp run the solver redundantly on identical values; the
// state is read with broadcast loads and written by lane 0 between two __syncwarp()s.  Compared
// with warp-uniform registers (~60 of them) this keeps the refine kernel at the register count
// of the scoring loop, and compared with shuffle-based "lane slots" it keeps the solver code
// small: the kernel is instruction-cache sensitive (profiles/r01_summary.md).
struct DpNelderMead {
  double P[4][3];   // simplex vertices
  double y[4];      // objective at the vertices
  double cs[3];     // coord_sum
  double pa[3];     // accepted reflection point
  double pt[3];     // point being evaluated
  double n0[3], p0[3], c3[3];  // patch normal / position at entry, reference camera centre
  double y_alpha, y_lo, y_nhi, y_hi;
};

// A stopped patch in global memory: the DpNelderMead words, then (state, idx, fcount, ilo, ihi).
#define DP_NM_STRUCT_WORDS ((int)(sizeof(DpNelderMead) / sizeof(double)))
#define DP_NM_SAVE_WORDS (DP_NM_STRUCT_WORDS + 5)

__device__ __forceinline__ void nm_store3(double *dst, const double v[3], int lane) {
  __syncwarp();
  if (lane == 0) { dst[0] = v[0]; dst[1] = v[1]; dst[2] = v[2]; }
  __syncwarp();
}

// coord_sum = sum of the vertices, accumulated in vertex order (updateCoordSum)
__device__ __forceinline__ void nm_coord_sum(DpNelderMead &S, int lane) {
  double t[3];
#pragma unroll
  for (int j = 0; j < 3; ++j)
    t[j] = xadd(xadd(xadd(xadd(0.0, S.P[0][j]), S.P[1][j]), S.P[2][j]), S.P[3][j]);
  nm_store3(S.cs, t, lane);
}

// tryNewPoint / replacePoint: ptry = coord_sum * (1-a)/n - p_hi * ((1-a)/n - a)  -> S.pt
__device__ __forceinline__ void nm_try_point(DpNelderMead &S, int lane, int ihi, double alpha_) {
  const double al = (1.0 - alpha_) / 3.0;
  const double be = xsub(al, alpha_);
  double pt[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) pt[j] = xsub(xmul(S.cs[j], al), xmul(S.P[ihi][j], be));
  nm_store3(S.pt, pt, lane);
}

// replacePoint: vertex ihi <- q with objective yq, then updateCoordSum
__device__ __forceinline__ void nm_replace(DpNelderMead &S, int lane, int ihi, const double q[3],
                                           double yq) {
  __syncwarp();
  if (lane == 0) {
    S.P[ihi][0] = q[0]; S.P[ihi][1] = q[1]; S.P[ihi][2] = q[2];
    S.y[ihi] = yq;
  }
  __syncwarp();
  nm_coord_sum(S, lane);
}

// vertex idx <- halfway to vertex ilo (the shrink step); also becomes the point to evaluate
__device__ __forceinline__ void nm_shrink_vertex(DpNelderMead &S, int lane, int idx, int ilo) {
  double pt[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) pt[j] = xmul(0.5, xadd(S.P[idx][j], S.P[ilo][j]));
  __syncwarp();
  if (lane == 0) {
#pragma unroll
    for (int j = 0; j < 3; ++j) { S.P[idx][j] = pt[j]; S.pt[j] = pt[j]; }
  }
  __syncwarp();
}

// WPP = warps per patch.  1: every warp refines its own patch (large batches).  4: the views of
// a patch are dealt out to four warps (view k >= 1 goes to warp (k - 1) mod 4, every warp also
// computes the anchor texture it scores against); the warps exchange their scores through shared
// memory (two named barriers per evaluation), then each of them sums the objective in view order
// and advances its own copy of the Nelder-Mead state -- identical values everywhere, no
// cross-warp solver traffic.  For batches that leave most of the machine idle (the late, small
// levels of an expansion, or one GPU's share of a level when the frontier is split over eight):
// the time of such a launch is the LATENCY of its longest patch (up to 500 dependent evaluations
// of up to ~50 views each), which four warps cut almost four-fold.
#define DP_MW_MAXV 256  // largest visible set the multi-warp form handles
__device__ __forceinline__ void dp_group_barrier(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

template <int NPASS, int WPP = 1>
__global__ void __launch_bounds__(DP_RWARPS * 32, dp_refine_min_ctas(NPASS)) dp_refine_kernel(DpRefineArgs a) {
  __shared__ DpWarpShared<NPASS, false> wsh[DP_RWARPS];
  __shared__ DpNelderMead nm[DP_RWARPS];
  __shared__ double gscore[WPP > 1 ? DP_RWARPS / WPP : 1][WPP > 1 ? DP_MW_MAXV : 1];
  __shared__ int32_t gvis[WPP > 1 ? DP_RWARPS : 1][WPP > 1 ? DP_MW_MAXV / WPP + 2 : 1];
  __shared__ unsigned int gitem[DP_RWARPS / WPP];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int grp = warp / WPP, wg = warp % WPP;  // the warp's patch slot in the CTA, its share of it
  const int s = a.p.s, npx = s * s;
  DpTexels<NPASS> tx;
  tx.init(s, lane);
  DpWarpShared<NPASS, false> &ws = wsh[warp];
  DpNelderMead &S = nm[warp];
  enum { ST_INIT, ST_REFLECT, ST_EXPAND, ST_CONTRACT, ST_SHRINK, ST_DONE };
  for (;;) {
    unsigned int iu = 0;
    if (WPP == 1) {
      if (lane == 0) iu = atomicAdd(a.work_counter, 1u);
      iu = __shfl_sync(DP_FULL, iu, 0);
    } else {
      if (wg == 0 && lane == 0) gitem[grp] = atomicAdd(a.work_counter, 1u);
      dp_group_barrier(1 + grp, WPP * 32);
      iu = gitem[grp];
      dp_group_barrier(1 + grp, WPP * 32);
    }
    if (iu >= a.n_items) break;
    const long long i = a.order ? (long long)a.order[iu] : (long long)iu;
    if (a.mask != nullptr && a.mask[i] == 0) {  // removed by Seed::RemovePatches (seed.cpp:146-156)
      if (!a.resume && lane == 0 && wg == 0) {  // (resume: finished in an earlier launch)
        if (a.evals) a.evals[i] = 0;
        if (a.pending) a.pending[i] = 0;
      }
      continue;
    }
    const int nv = min(a.p.nvis[i], a.p.vstride);
    const int ref = a.p.ref[i];
    const bool ref_ok = ref >= 0 && ref < a.p.n_views;
    const int32_t *vis = a.p.vis + (size_t)i * a.p.vstride;
    int nvw = nv;  // this warp's views: all of them, or the anchor and every WPP-th other view
    if (WPP > 1) {
      nvw = nv >= 1 ? 1 + (nv - 1 - wg + WPP - 1) / WPP : 0;
      if (nv - 1 - wg < 0) nvw = nv >= 1 ? 1 : 0;
      for (int t = lane; t < nvw; t += 32) gvis[warp][t] = t == 0 ? vis[0] : vis[1 + wg + (t - 1) * WPP];
    }
    __syncwarp();
    if (lane == 0) {
      // createInitialSimplex: v_i = x0 + step_{i-1}/2 e_{i-1}, then v_0 = x0 - step/2; x0 = 0
      const double *C = a.p.views[ref_ok ? ref : 0].center;
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        const double h = xmul(0.5, a.step[j]);
        S.P[0][j] = xsub(0.0, h);
        S.pt[j] = xsub(0.0, h);
#pragma unroll
        for (int v = 1; v < 4; ++v) S.P[v][j] = (v - 1 == j) ? xadd(0.0, h) : 0.0;
        S.n0[j] = (double)a.p.nrm[3 * i + j];
        S.p0[j] = (double)a.p.pos[3 * i + j];
        S.c3[j] = C[j];
      }
    }
    __syncwarp();
    int state = ST_INIT, idx = 0, fcount = 4;
    int ilo = 0, ihi = 0;
    if (a.resume) {  // continue a stopped patch: every warp of the group loads the same state
      const double *sv = a.nm_save + (size_t)i * DP_NM_SAVE_WORDS;
      double *Sd = reinterpret_cast<double *>(&S);
      for (int t = lane; t < DP_NM_STRUCT_WORDS; t += 32) Sd[t] = sv[t];
      state = (int)sv[DP_NM_STRUCT_WORDS];
      idx = (int)sv[DP_NM_STRUCT_WORDS + 1];
      fcount = (int)sv[DP_NM_STRUCT_WORDS + 2];
      ilo = (int)sv[DP_NM_STRUCT_WORDS + 3];
      ihi = (int)sv[DP_NM_STRUCT_WORDS + 4];
      __syncwarp();
    }
    int spent = 0;  // objective evaluations of this patch in this launch
    const int budget = (a.budget > 0 && a.budget_views > 0 && nv > a.budget_views)
                           ? max(8, (int)(((long long)a.budget * a.budget_views) / nv))
                           : a.budget;
#pragma unroll 1
    for (;;) {
      if (budget > 0 && spent >= budget && state != ST_DONE) {
        // out of budget: the state goes to global memory as it is, S.pt is the next point
        __syncwarp();
        if (wg == 0) {
          double *sv = a.nm_save + (size_t)i * DP_NM_SAVE_WORDS;
          const double *Sd = reinterpret_cast<const double *>(&S);
          for (int t = lane; t < DP_NM_STRUCT_WORDS; t += 32) sv[t] = Sd[t];
          if (lane == 0) {
            sv[DP_NM_STRUCT_WORDS] = (double)state;
            sv[DP_NM_STRUCT_WORDS + 1] = (double)idx;
            sv[DP_NM_STRUCT_WORDS + 2] = (double)fcount;
            sv[DP_NM_STRUCT_WORDS + 3] = (double)ilo;
            sv[DP_NM_STRUCT_WORDS + 4] = (double)ihi;
            a.pending[i] = 1;
            atomicAdd(a.pending_count, 1u);
          }
        }
        break;
      }
      ++spent;
      // UnparametrizePatch at the point to evaluate (or, in ST_DONE, at the best vertex)
      double n[3], p[3];
      const double p0[3] = {S.p0[0], S.p0[1], S.p0[2]};  // GetPosition(): the corner centre
      {
        const double c3[3] = {S.c3[0], S.c3[1], S.c3[2]};
        const double n0[3] = {S.n0[0], S.n0[1], S.n0[2]};
        dp_unparametrize(c3, n0, p0, S.pt[0], S.pt[1], S.pt[2], n, p);
      }
      if (state == ST_DONE) {
        // SetNormal / SetPosition store fp32 (patch.h:38-53)
        if (lane < 3 && wg == 0) {
          const double nv_ = lane == 0 ? n[0] : (lane == 1 ? n[1] : n[2]);
          const double pv_ = lane == 0 ? p[0] : (lane == 1 ? p[1] : p[2]);
          if (ref_ok) {
            a.p.nrm[3 * i + lane] = (float)nv_;
            a.p.pos[3 * i + lane] = (float)pv_;
          }
          if (a.xbest) a.xbest[3 * i + lane] = S.pt[lane];
        }
        if (lane == 0 && wg == 0) {
          if (a.evals) a.evals[i] = fcount;
          if (a.pending) a.pending[i] = 0;
        }
        break;
      }
      // ---- the single objective call site: PatchOptimizationOpenCVFunctor::calc ----------
      double fval = 2.0;  // scores.size() == 0 (optimization_opencv.cpp:30-32)
      if (nv >= 2 && ref_ok) {
        double sum = 0.0;
        if (WPP == 1) {
          dp_eval_views<NPASS, false, false>(
              a.p.views, a.p.n_views, a.p.lv, ref, vis, nv, s, npx, n, p, p0, tx, ws, lane, nullptr,
              nullptr, [&](int k0, int kc, double score) {
                // std::accumulate of (1 - NCC) in view order (optimization_opencv.cpp:24, 34)
                const double term = xsub(1.0, score);
                for (int l = (k0 == 0 ? 1 : 0); l < kc; ++l)
                  sum = xadd(sum, __shfl_sync(DP_FULL, term, l));
              });
        } else {
          dp_eval_views<NPASS, false, false>(
              a.p.views, a.p.n_views, a.p.lv, ref, gvis[warp], nvw, s, npx, n, p, p0, tx, ws, lane, nullptr,
              nullptr, [&](int k0, int kc, double score) {
                const int t = k0 + lane;  // entry of this warp's list -> view 1 + wg + (t - 1) WPP
                if (lane < kc && t >= 1) gscore[grp][1 + wg + (t - 1) * WPP] = score;
              });
          dp_group_barrier(1 + grp, WPP * 32);
          for (int k = 1; k < nv; ++k) sum = xadd(sum, xsub(1.0, gscore[grp][k]));  // view order
          dp_group_barrier(1 + grp, WPP * 32);
        }
        fval = sum / (double)(nv - 1);
      }
      // ---- consume it according to the Nelder-Mead state ---------------------------------
      bool decide = false;
      if (state == ST_INIT) {
        __syncwarp();
        if (lane == 0) S.y[idx] = fval;
        __syncwarp();
        if (++idx < 4) {
          const double q[3] = {S.P[idx][0], S.P[idx][1], S.P[idx][2]};
          nm_store3(S.pt, q, lane);
        } else {
          nm_coord_sum(S, lane);
          decide = true;
        }
      } else if (state == ST_REFLECT) {
        const double q[3] = {S.pt[0], S.pt[1], S.pt[2]};
        const double y_lo = S.y_lo, y_nhi = S.y_nhi;
        __syncwarp();
        if (lane == 0) {
          S.pa[0] = q[0]; S.pa[1] = q[1]; S.pa[2] = q[2];
          S.y_alpha = fval;
        }
        __syncwarp();
        if (fval < y_nhi) {
          if (fval < y_lo) {  // better than the best: try twice as far
            state = ST_EXPAND;
            nm_try_point(S, lane, ihi, -2.0);
            ++fcount;
          } else {
            nm_replace(S, lane, ihi, q, fval);  // replacePoint(alpha = -1)
            decide = true;
          }
        } else {
          state = ST_CONTRACT;
          nm_try_point(S, lane, ihi, 0.5);
          ++fcount;
        }
      } else if (state == ST_EXPAND) {
        const double y_alpha = S.y_alpha;
        const bool better = fval < y_alpha;
        const double q[3] = {better ? S.pt[0] : S.pa[0], better ? S.pt[1] : S.pa[1],
                             better ? S.pt[2] : S.pa[2]};
        nm_replace(S, lane, ihi, q, better ? fval : y_alpha);
        decide = true;
      } else if (state == ST_CONTRACT) {
        if (fval < S.y_hi) {
          const double q[3] = {S.pt[0], S.pt[1], S.pt[2]};
          nm_replace(S, lane, ihi, q, fval);
          decide = true;
        } else {  // shrink every vertex but the best halfway towards it
          state = ST_SHRINK;
          idx = (ilo == 0) ? 1 : 0;
          nm_shrink_vertex(S, lane, idx, ilo);
        }
      } else {  // ST_SHRINK
        __syncwarp();
        if (lane == 0) S.y[idx] = fval;
        __syncwarp();
        ++idx;
        if (idx == ilo) ++idx;
        if (idx < 4) {
          nm_shrink_vertex(S, lane, idx, ilo);
        } else {
          fcount += 3;
          nm_coord_sum(S, lane);
          decide = true;
        }
      }
      if (!decide) continue;
      // ---- find worst, next-to-worst and best vertices; stop test ------------------------
      const double yv[4] = {S.y[0], S.y[1], S.y[2], S.y[3]};
      int inhi;
      double ylo = yv[0], yhi, ynhi;
      ilo = 0;
      if (yv[0] > yv[1]) { ihi = 0; yhi = yv[0]; inhi = 1; ynhi = yv[1]; }
      else { ihi = 1; yhi = yv[1]; inhi = 0; ynhi = yv[0]; }
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        const double yc = yv[v];
        if (yc <= ylo) { ilo = v; ylo = yc; }
        if (yc > yhi) { inhi = ihi; ynhi = yhi; ihi = v; yhi = yc; }
        else if (yc > ynhi && v != ihi) { inhi = v; ynhi = yc; }
      }
      if (ilo == inhi || ilo == ihi) {
#pragma unroll
        for (int v = 3; v >= 0; --v)  // ascending search, first match wins
          if (yv[v] == ylo && v != ihi && v != inhi) ilo = v;
      }
      const double error = fabs(xsub(yhi, ylo));
      double range = 0.0;
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        double mn = S.P[0][j], mx = S.P[0][j];
#pragma unroll
        for (int v = 1; v < 4; ++v) { mn = fmin(mn, S.P[v][j]); mx = fmax(mx, S.P[v][j]); }
        range = fmax(range, fabs(xsub(mx, mn)));
      }
      if (range <= a.eps || error <= a.eps || fcount >= a.max_evals) {
        // best vertex -> x: one more trip through UnparametrizePatch, then write back
        const double q[3] = {S.P[ilo][0], S.P[ilo][1], S.P[ilo][2]};
        nm_store3(S.pt, q, lane);
        state = ST_DONE;
        continue;
      }
      __syncwarp();
      if (lane == 0) { S.y_lo = ylo; S.y_nhi = ynhi; S.y_hi = yhi; }
      __syncwarp();
      state = ST_REFLECT;  // reflect the worst point about the centroid of the others
      nm_try_point(S, lane, ihi, -1.0);
      ++fcount;
    }
  }
}
