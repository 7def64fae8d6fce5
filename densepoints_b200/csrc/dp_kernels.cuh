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

struct DpPatchArgs {
  const DpViewDev *views;
  int n_views;
  int n, vstride;
  float *pos, *nrm;
  int32_t *ref, *nvis, *vis;
  int s;
};

struct DpScoreArgs {
  DpPatchArgs p;
  float *ncc;      // n*vstride or null
  uint8_t *tex;    // n*vstride*s*s*3 or null
  uint8_t *valid;  // n*vstride or null
  // filter epilogue
  double thr;
  int min_visible;
  uint8_t *keep;
};

template <int NPASS>
struct DpTileCfg {
  // 4x the texel count covers oblique / zoomed ROIs; capped at 4 KB per warp.  A larger
  // ROI is gathered straight from global memory instead.
  static constexpr int kTilePx = (128 * NPASS < 1024) ? 128 * NPASS : 1024;
};

// Evaluate all visible views of one patch at (n, p).  For every k >= 1 calls
// sink(k, score) with NCCScore(texture 0, texture k) (-1 when either is empty).
template <int NPASS, bool WRITE_TEX, typename Sink>
__device__ __forceinline__ void dp_eval_views(const DpViewDev *__restrict__ views, int n_views,
                                              int ref, const int32_t *vis, int nv,
                                              int s, int npx, const double n[3], const double p[3],
                                              const DpTexels<NPASS> &tx, uint32_t *tile, int lane,
                                              uint8_t *tex_base, uint8_t *valid_base, Sink sink) {
  DpFrame f;
  const bool ref_ok = (ref >= 0 && ref < n_views);
  if (ref_ok)
    dp_make_frame(views + ref, s, n, p, f);
  else
    f.ok = false;
  DpAnchor<NPASS> anchor;
  anchor.valid = false;
  for (int k = 0; k < nv; ++k) {
    const int vid = vis[k];
    int g[NPASS];
    bool ok = false;
    if (f.ok && vid >= 0 && vid < n_views)
      ok = dp_view_texture<NPASS, WRITE_TEX>(views + vid, s, npx, f, tx, tile,
                                             DpTileCfg<NPASS>::kTilePx, lane, g,
                                             WRITE_TEX ? tex_base + (size_t)k * npx * 3 : nullptr);
    if (valid_base != nullptr && lane == 0) valid_base[k] = ok ? 1 : 0;
    if (k == 0) {
      anchor.valid = ok;
      if (ok) dp_set_anchor<NPASS>(g, npx, lane, anchor);
    } else {
      double score = -1.0;  // NCCScore on an empty Mat (error_measurements.cpp:38-40)
      if (ok && anchor.valid) score = dp_ncc<NPASS>(anchor, g, npx, lane);
      sink(k, score);
    }
  }
}

template <int NPASS, bool WRITE_TEX, bool FILTER>
__global__ void __launch_bounds__(DP_WARPS * 32) dp_score_kernel(DpScoreArgs a) {
  __shared__ uint32_t tiles[DP_WARPS][DpTileCfg<NPASS>::kTilePx];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long i = (long long)blockIdx.x * DP_WARPS + warp;
  if (i >= a.p.n) return;
  const int s = a.p.s, npx = s * s;
  DpTexels<NPASS> tx;
  tx.init(s, lane);
  const int nv = min(a.p.nvis[i], a.p.vstride);
  const int ref = a.p.ref[i];
  double n[3] = {(double)a.p.nrm[3 * i], (double)a.p.nrm[3 * i + 1], (double)a.p.nrm[3 * i + 2]};
  double p[3] = {(double)a.p.pos[3 * i], (double)a.p.pos[3 * i + 1], (double)a.p.pos[3 * i + 2]};
  int32_t *vis = a.p.vis + (size_t)i * a.p.vstride;
  float *ncc = a.ncc ? a.ncc + (size_t)i * a.p.vstride : nullptr;
  uint8_t *tex = WRITE_TEX ? a.tex + (size_t)i * a.p.vstride * npx * 3 : nullptr;
  uint8_t *valid = a.valid ? a.valid + (size_t)i * a.p.vstride : nullptr;
  // FilterByErrorMeasurement's erase loop (optimization.cpp:117-124) erases
  // visible[i - removed] when scores[i] (the score of visible[i+1]) is low; since every
  // erased entry lies before the cursor this is "drop original entry k-1 iff the score of
  // entry k is low", and the last entry always survives -- evaluated online here.
  int wcur = 0;
  int prev = nv > 0 ? vis[0] : -1;
  const double thr = a.thr;
  dp_eval_views<NPASS, WRITE_TEX>(
      a.p.views, a.p.n_views, ref, vis, nv, s, npx, n, p, tx, tiles[warp], lane, tex, valid,
      [&](int k, double score) {
        if (ncc != nullptr && lane == 0) ncc[k] = (float)score;
        if (FILTER) {
          if (!(score < thr)) {
            if (lane == 0) vis[wcur] = prev;  // wcur <= k-1: entry k is still unread-safe
            ++wcur;
          }
          prev = vis[k];
        }
      });
  if (FILTER) {
    bool kept = false;
    if (nv >= 2) {  // scores.size() > 0 (optimization.cpp:113)
      if (lane == 0) vis[wcur] = prev;
      ++wcur;
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
};

// Optimization::UnparametrizePatch (optimization.cpp:78-96)
__device__ __forceinline__ void dp_unparametrize(const double *__restrict__ C, const double n0[3],
                                                 const double p0[3], double depth, double roll,
                                                 double pitch, double n[3], double p[3]) {
  const double k = xadd(1.0, depth);
#pragma unroll
  for (int j = 0; j < 3; ++j) p[j] = xadd(C[j], xmul(k, xsub(p0[j], C[j])));
  double sa, ca, sb, cb;
  sincos(roll, &sa, &ca);
  sincos(pitch, &sb, &cb);
  // rotation rows: [cb 0 -sb; sa*sb ca cb*sa; ca*sb -sa ca*cb]
  n[0] = xadd(xmul(cb, n0[0]), xmul(-sb, n0[2]));
  n[1] = xadd(xadd(xmul(xmul(sa, sb), n0[0]), xmul(ca, n0[1])), xmul(xmul(cb, sa), n0[2]));
  n[2] = xadd(xadd(xmul(xmul(ca, sb), n0[0]), xmul(-sa, n0[1])), xmul(xmul(ca, cb), n0[2]));
}

// Nelder-Mead state of one patch (cv::DownhillSolver, ndim = 3).  Every lane of the
// owning warp holds the same values in registers; run-time vertex indices (ilo / ihi)
// are resolved with unrolled selects so nothing is spilled to local memory.
struct DpSimplex {
  double P[4][3];  // vertices
  double y[4];     // objective at the vertices
  double cs[3];    // coord_sum
  __device__ __forceinline__ void getP(int v, double o[3]) const {
#pragma unroll
    for (int j = 0; j < 3; ++j) o[j] = v == 0 ? P[0][j] : (v == 1 ? P[1][j] : (v == 2 ? P[2][j] : P[3][j]));
  }
  __device__ __forceinline__ void setP(int v, const double o[3]) {
#pragma unroll
    for (int w = 0; w < 4; ++w)
#pragma unroll
      for (int j = 0; j < 3; ++j) P[w][j] = (w == v) ? o[j] : P[w][j];
  }
  __device__ __forceinline__ void setY(int v, double val) {
#pragma unroll
    for (int w = 0; w < 4; ++w) y[w] = (w == v) ? val : y[w];
  }
};

__device__ __forceinline__ void dp_coord_sum(DpSimplex &S) {
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    double t = 0.0;
#pragma unroll
    for (int v = 0; v < 4; ++v) t = xadd(t, S.P[v][j]);
    S.cs[j] = t;
  }
}

// tryNewPoint / replacePoint: ptry = coord_sum * (1-a)/n - p_hi * ((1-a)/n - a)
__device__ __forceinline__ void dp_try_point(const DpSimplex &S, int ihi, double alpha_,
                                             double pt[3]) {
  const double al = (1.0 - alpha_) / 3.0;
  const double be = xsub(al, alpha_);
  double ph[3];
  S.getP(ihi, ph);
#pragma unroll
  for (int j = 0; j < 3; ++j) pt[j] = xsub(xmul(S.cs[j], al), xmul(ph[j], be));
}

template <int NPASS>
__global__ void __launch_bounds__(DP_WARPS * 32) dp_refine_kernel(DpRefineArgs a) {
  __shared__ uint32_t tiles[DP_WARPS][DpTileCfg<NPASS>::kTilePx];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int s = a.p.s, npx = s * s;
  DpTexels<NPASS> tx;
  tx.init(s, lane);
  uint32_t *tile = tiles[warp];
  DpSimplex S;
  enum { ST_INIT, ST_REFLECT, ST_EXPAND, ST_CONTRACT, ST_SHRINK };
  for (;;) {
    unsigned int iu = 0;
    if (lane == 0) iu = atomicAdd(a.work_counter, 1u);
    iu = __shfl_sync(DP_FULL, iu, 0);
    if (iu >= (unsigned int)a.p.n) break;
    const long long i = iu;
    if (a.mask != nullptr && a.mask[i] == 0) {  // removed by Seed::RemovePatches (seed.cpp:146-156)
      if (a.evals && lane == 0) a.evals[i] = 0;
      continue;
    }
    const int nv = min(a.p.nvis[i], a.p.vstride);
    const int ref = a.p.ref[i];
    const bool ref_ok = ref >= 0 && ref < a.p.n_views;
    const double n0[3] = {(double)a.p.nrm[3 * i], (double)a.p.nrm[3 * i + 1],
                          (double)a.p.nrm[3 * i + 2]};
    const double p0[3] = {(double)a.p.pos[3 * i], (double)a.p.pos[3 * i + 1],
                          (double)a.p.pos[3 * i + 2]};
    const int32_t *vis = a.p.vis + (size_t)i * a.p.vstride;
    const double *C = a.p.views[ref_ok ? ref : 0].center;
    const double c3[3] = {C[0], C[1], C[2]};

    // createInitialSimplex: v_i = x0 + step_{i-1}/2 e_{i-1}, then v_0 = x0 - step/2; x0 = 0
#pragma unroll
    for (int v = 0; v < 4; ++v)
#pragma unroll
      for (int j = 0; j < 3; ++j)
        S.P[v][j] = (v == 0) ? xsub(0.0, xmul(0.5, a.step[j]))
                             : ((v - 1 == j) ? xadd(0.0, xmul(0.5, a.step[j])) : 0.0);
    int state = ST_INIT, idx = 0, fcount = 4;
    int ilo = 0, ihi = 0;
    double y_lo = 0, y_nhi = 0, y_hi = 0, y_alpha = 0;
    double pt[3] = {S.P[0][0], S.P[0][1], S.P[0][2]}, pa[3] = {0, 0, 0};
#pragma unroll 1
    for (;;) {
      // ---- the single objective call site: PatchOptimizationOpenCVFunctor::calc ----------
      double fval = 2.0;  // scores.size() == 0 (optimization_opencv.cpp:30-32)
      if (nv >= 2 && ref_ok) {
        double n[3], p[3];
        dp_unparametrize(c3, n0, p0, pt[0], pt[1], pt[2], n, p);
        double sum = 0.0;
        dp_eval_views<NPASS, false>(a.p.views, a.p.n_views, ref, vis, nv, s, npx, n, p, tx, tile,
                                    lane, nullptr, nullptr,
                                    [&](int, double score) { sum = xadd(sum, xsub(1.0, score)); });
        fval = sum / (double)(nv - 1);
      }
      // ---- consume it according to the Nelder-Mead state ---------------------------------
      bool decide = false;
      if (state == ST_INIT) {
        S.setY(idx, fval);
        if (++idx < 4) {
          S.getP(idx, pt);
        } else {
          dp_coord_sum(S);
          decide = true;
        }
      } else if (state == ST_REFLECT) {
        y_alpha = fval;
        pa[0] = pt[0]; pa[1] = pt[1]; pa[2] = pt[2];
        if (y_alpha < y_nhi) {
          if (y_alpha < y_lo) {  // try twice as far
            state = ST_EXPAND;
            dp_try_point(S, ihi, -2.0, pt);
            ++fcount;
          } else {
            decide = true;
          }
        } else {
          state = ST_CONTRACT;
          dp_try_point(S, ihi, 0.5, pt);
          ++fcount;
        }
        if (decide) {  // replacePoint(alpha)
          S.setP(ihi, pa);
          S.setY(ihi, y_alpha);
          dp_coord_sum(S);
        }
      } else if (state == ST_EXPAND) {
        if (fval < y_alpha) {
          y_alpha = fval;
          pa[0] = pt[0]; pa[1] = pt[1]; pa[2] = pt[2];
        }
        S.setP(ihi, pa);
        S.setY(ihi, y_alpha);
        dp_coord_sum(S);
        decide = true;
      } else if (state == ST_CONTRACT) {
        if (fval < y_hi) {
          S.setP(ihi, pt);
          S.setY(ihi, fval);
          dp_coord_sum(S);
          decide = true;
        } else {  // shrink every vertex but the best halfway towards it
          state = ST_SHRINK;
          idx = (ilo == 0) ? 1 : 0;
          double pi[3], pl[3];
          S.getP(idx, pi);
          S.getP(ilo, pl);
#pragma unroll
          for (int j = 0; j < 3; ++j) pt[j] = xmul(0.5, xadd(pi[j], pl[j]));
          S.setP(idx, pt);
        }
      } else {  // ST_SHRINK
        S.setY(idx, fval);
        ++idx;
        if (idx == ilo) ++idx;
        if (idx < 4) {
          double pi[3], pl[3];
          S.getP(idx, pi);
          S.getP(ilo, pl);
#pragma unroll
          for (int j = 0; j < 3; ++j) pt[j] = xmul(0.5, xadd(pi[j], pl[j]));
          S.setP(idx, pt);
        } else {
          fcount += 3;
          dp_coord_sum(S);
          decide = true;
        }
      }
      if (!decide) continue;
      // ---- find worst, next-to-worst and best vertices; stop test ------------------------
      int inhi;
      double ylo = S.y[0], yhi, ynhi;
      ilo = 0;
      if (S.y[0] > S.y[1]) { ihi = 0; yhi = S.y[0]; inhi = 1; ynhi = S.y[1]; }
      else { ihi = 1; yhi = S.y[1]; inhi = 0; ynhi = S.y[0]; }
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        const double yv = S.y[v];
        if (yv <= ylo) { ilo = v; ylo = yv; }
        if (yv > yhi) { inhi = ihi; ynhi = yhi; ihi = v; yhi = yv; }
        else if (yv > ynhi && v != ihi) { inhi = v; ynhi = yv; }
      }
      if (ilo == inhi || ilo == ihi) {
#pragma unroll
        for (int v = 3; v >= 0; --v)  // ascending search, first match wins
          if (S.y[v] == ylo && v != ihi && v != inhi) ilo = v;
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
      if (range <= a.eps || error <= a.eps || fcount >= a.max_evals) break;
      y_lo = ylo; y_nhi = ynhi; y_hi = yhi;
      state = ST_REFLECT;  // reflect the worst point about the centroid of the others
      dp_try_point(S, ihi, -1.0, pt);
      ++fcount;
    }
    // best vertex -> x; UnparametrizePatch; SetNormal / SetPosition store fp32
    double xb[3];
    S.getP(ilo, xb);
    double n[3], p[3];
    dp_unparametrize(c3, n0, p0, xb[0], xb[1], xb[2], n, p);
    if (lane < 3) {
      const double nv_ = lane == 0 ? n[0] : (lane == 1 ? n[1] : n[2]);
      const double pv_ = lane == 0 ? p[0] : (lane == 1 ? p[1] : p[2]);
      if (ref_ok) {
        a.p.nrm[3 * i + lane] = (float)nv_;
        a.p.pos[3 * i + lane] = (float)pv_;
      }
      if (a.xbest) a.xbest[3 * i + lane] = lane == 0 ? xb[0] : (lane == 1 ? xb[1] : xb[2]);
    }
    if (a.evals && lane == 0) a.evals[i] = fcount;
  }
}
