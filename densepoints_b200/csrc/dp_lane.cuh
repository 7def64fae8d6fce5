// dp_lane.cuh -- refine kernel with ONE PATCH PER LANE (cells up to 8x8).
//
// The group kernels (dp_group.cuh) give a patch 4 lanes; everything that is scalar per patch or
// per (patch, view) -- UnparametrizePatch, the patch frame, the Nelder-Mead bookkeeping, the
// four corner projections, the cell -> quad map, the NCC finish -- is still computed four times
// over, the texels of a view are spread over lanes that then have to meet again through
// shuffles and group barriers, and the round-1 profile put only 112 of the 215 warp
// instructions per patch-view evaluation into texel work.  Here a lane owns a patch outright
// and the warp advances 32 patches in lockstep:
//   * no instruction is redundant: every lane of every "scalar" instruction works for its own
//     patch, the per-view set-up is plain per-lane code (no ballots, no shuffles, no records);
//   * a lane walks its s x s texels itself, row by row (the row is unrolled, so x is an
//     immediate and the row terms of the projective map are hoisted); integer moments and the
//     NCC numerator are lane-private sums in the reference's own order -- no reductions;
//   * the only cooperation left is the staging copy (lane pairs split a tile's 16-byte pieces
//     so that neighbouring pieces travel in one request) and one __syncwarp() per view.
// Shared memory per lane: a staging tile of DP_LTILE pixels (its ROI, row pitch = the ROI width
// rounded up to 4 pixels), two rows-of-bytes gray buffers (anchor / current view) and the
// 23 doubles of the Nelder-Mead state, all laid out [item][lane] so that a lane always stays in
// its own banks for the private data; the tiles of different lanes are staggered.
//
// The arithmetic is the one of dp_group.cuh / dp_device.cuh, expression for expression (same
// functions where they are shared); the NCC numerator is accumulated texel by texel in row-major
// order, i.e. exactly as the oracle does.
#pragma once
#include "dp_group.cuh"

#ifndef DP_LWARPS
#define DP_LWARPS 2       // warps per CTA
#endif
#ifndef DP_LMINCTA
#define DP_LMINCTA 4      // resident CTAs per SM asked for
#endif
#ifndef DP_LTILE
#define DP_LTILE 96       // pixels of a lane's staging tile (s = 7: ROIs are <= 12 x 7 with the
#endif                    // alignment offset on BASELINE configs[1]; larger ROIs gather from L2)
#define DP_LTSTR (DP_LTILE + 4)   // tile stride between lanes: 16-byte aligned, staggers the banks
#define DP_LGROUP_MAX_CELL 8
#define DP_NM_WORDS 23    // P[4][3], y[4], pt[3], pa[3], y_alpha
#ifndef DP_LANE_NM_LOCAL
#define DP_LANE_NM_LOCAL 0  // 1: the Nelder-Mead words live in (L1-cached) local memory
#endif
#if DP_LANE_NM_LOCAL
#define DP_NM_STRIDE 1
#else
#define DP_NM_STRIDE 32
#endif

template <int S>
struct DpLaneShared {
  // +32 words: the neighbour taps of an edge pixel (weight 0) may read past the last tile
  __align__(16) uint32_t tile[DP_LWARPS][32 * DP_LTSTR + 32];
  uint2 gray[DP_LWARPS][2][S][32];          // [0] anchor texture, [1] current view: a row of bytes
#if !DP_LANE_NM_LOCAL
  double nm[DP_LWARPS][DP_NM_WORDS][32];
#endif
};

// Nelder-Mead state of the lane's patch: word k at b[32 k] (shared memory, [word][lane]).
struct DpNmLane {
  double *b;
  __device__ __forceinline__ double &P(int v, int j) const { return b[DP_NM_STRIDE * (3 * v + j)]; }
  __device__ __forceinline__ double &y(int v) const { return b[DP_NM_STRIDE * (12 + v)]; }
  __device__ __forceinline__ double &pt(int j) const { return b[DP_NM_STRIDE * (16 + j)]; }
  __device__ __forceinline__ double &pa(int j) const { return b[DP_NM_STRIDE * (19 + j)]; }
  __device__ __forceinline__ double &y_alpha() const { return b[DP_NM_STRIDE * 22]; }
};

// tryNewPoint: ptry = coord_sum * (1-a)/n - p_hi * ((1-a)/n - a) -> pt; coord_sum is the sum of
// the vertices in vertex order (updateCoordSum), a pure function of the simplex.
__device__ __forceinline__ void nml_try_point(const DpNmLane &S, int ihi, double alpha_) {
  const double al = (1.0 - alpha_) / 3.0;
  const double be = xsub(al, alpha_);
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    const double cs = xadd(xadd(xadd(xadd(0.0, S.P(0, j)), S.P(1, j)), S.P(2, j)), S.P(3, j));
    S.pt(j) = xsub(xmul(cs, al), xmul(S.P(ihi, j), be));
  }
}
__device__ __forceinline__ void nml_replace(const DpNmLane &S, int ihi, double q0, double q1,
                                            double q2, double yq) {
  S.P(ihi, 0) = q0; S.P(ihi, 1) = q1; S.P(ihi, 2) = q2;
  S.y(ihi) = yq;
}
__device__ __forceinline__ void nml_shrink_vertex(const DpNmLane &S, int idx, int ilo) {
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    const double v = xmul(0.5, xadd(S.P(idx, j), S.P(ilo, j)));
    S.P(idx, j) = v;
    S.pt(j) = v;
  }
}

// One step of cv::DownhillSolver's state machine -- the decision tree of nmg_step
// (dp_group.cuh) on lane-private state: consumes the objective value of pt, leaves the next
// point to evaluate in pt; true when the solver stops (pt = best vertex).
__device__ __forceinline__ bool nml_step(const DpNmLane &S, int &state, int &idx, int &fcount,
                                         int &ilo, int &ihi, double &y_lo, double &y_nhi,
                                         double &y_hi, double fval, double eps, int max_evals,
                                         const double step[3]) {
  bool decide = false;
  if (state == NMG_INIT) {
    S.y(idx) = fval;
    const int nx = idx + 1;
    idx = nx;
    if (nx < 4) {
      // vertex nx of the initial simplex, recomputed instead of copied from P(nx, .): ptxas
      // 12.9 folded the row offset of that dynamically indexed shared-memory load twice
      // (LDS [base + (idx+1)*768 + 768]: it evaluated vertex idx + 2; found with a trace build)
#pragma unroll
      for (int j = 0; j < 3; ++j) S.pt(j) = (nx - 1 == j) ? xadd(0.0, xmul(0.5, step[j])) : 0.0;
    } else {
      decide = true;
    }
  } else if (state == NMG_REFLECT) {
    const double q0 = S.pt(0), q1 = S.pt(1), q2 = S.pt(2);
    S.pa(0) = q0; S.pa(1) = q1; S.pa(2) = q2;
    S.y_alpha() = fval;
    if (fval < y_nhi) {
      if (fval < y_lo) {  // better than the best: try twice as far
        state = NMG_EXPAND;
        nml_try_point(S, ihi, -2.0);
        ++fcount;
      } else {
        nml_replace(S, ihi, q0, q1, q2, fval);  // replacePoint(alpha = -1)
        decide = true;
      }
    } else {
      state = NMG_CONTRACT;
      nml_try_point(S, ihi, 0.5);
      ++fcount;
    }
  } else if (state == NMG_EXPAND) {
    const double ya = S.y_alpha();
    const bool better = fval < ya;
    nml_replace(S, ihi, better ? S.pt(0) : S.pa(0), better ? S.pt(1) : S.pa(1),
                better ? S.pt(2) : S.pa(2), better ? fval : ya);
    decide = true;
  } else if (state == NMG_CONTRACT) {
    if (fval < y_hi) {
      nml_replace(S, ihi, S.pt(0), S.pt(1), S.pt(2), fval);
      decide = true;
    } else {  // shrink every vertex but the best halfway towards it
      state = NMG_SHRINK;
      idx = (ilo == 0) ? 1 : 0;
      nml_shrink_vertex(S, idx, ilo);
    }
  } else {  // NMG_SHRINK
    S.y(idx) = fval;
    ++idx;
    if (idx == ilo) ++idx;
    if (idx < 4) {
      nml_shrink_vertex(S, idx, ilo);
    } else {
      fcount += 3;
      decide = true;
    }
  }
  if (!decide) return false;
  // ---- find worst, next-to-worst and best vertices; stop test ------------------------------
  const double yv[4] = {S.y(0), S.y(1), S.y(2), S.y(3)};
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
    double mn = S.P(0, j), mx = mn;
#pragma unroll
    for (int v = 1; v < 4; ++v) { const double pv = S.P(v, j); mn = fmin(mn, pv); mx = fmax(mx, pv); }
    range = fmax(range, fabs(xsub(mx, mn)));
  }
  if (range <= eps || error <= eps || fcount >= max_evals) {
#pragma unroll
    for (int j = 0; j < 3; ++j) S.pt(j) = S.P(ilo, j);  // best vertex -> x
    return true;
  }
  y_lo = ylo; y_nhi = ynhi; y_hi = yhi;
  state = NMG_REFLECT;  // reflect the worst point about the centroid of the others
  nml_try_point(S, ihi, -1.0);
  ++fcount;
  return false;
}

// Optimization::UnparametrizePatch (optimization.cpp:78-96), every lane for its own patch.
// Out of line and with one rolled sincos site: it runs once per evaluation (not per view) from
// two places, and the refine kernel is instruction-cache sensitive (ncu r2l: 16 % of the warp
// samples waited for instructions).
__device__ __noinline__ void dp_unparametrize_l(const double C[3], const double n0[3],
                                                const double p0[3], const double x[3], double n[3],
                                                double p[3]) {
  const double k = xadd(1.0, x[0]);
#pragma unroll
  for (int j = 0; j < 3; ++j) p[j] = xadd(C[j], xmul(k, xsub(p0[j], C[j])));
  double sn[2], cs[2];
#pragma unroll 1
  for (int a = 0; a < 2; ++a) sincos(x[1 + a], &sn[a], &cs[a]);
  const double sa = sn[0], ca = cs[0], sb = sn[1], cb = cs[1];
  n[0] = xadd(xmul(cb, n0[0]), xmul(-sb, n0[2]));
  n[1] = xadd(xadd(xmul(xmul(sa, sb), n0[0]), xmul(ca, n0[1])), xmul(xmul(cb, sa), n0[2]));
  n[2] = xadd(xadd(xmul(xmul(ca, sb), n0[0]), xmul(-sa, n0[1])), xmul(xmul(ca, cb), n0[2]));
}

#ifndef DP_LANE_PROJECT_OOL
#define DP_LANE_PROJECT_OOL 1
#endif
// View::ProjectPoint out of line for the lane kernels' set-up (4 call sites, two fp64 divisions
// each): the refine kernel is instruction-cache sensitive.
__device__ __noinline__ void dp_project_ool(const double *__restrict__ P, double X0, double X1,
                                            double X2, double *uv) {
  dp_project(P, X0, X1, X2, uv[0], uv[1]);
}

// Per-lane set-up of one view (patch.cpp:111-164 up to the homography): what dp_setup_views
// computes with four lanes and shuffles, here straight-line code of one lane.
struct DpLaneView {
  double M0, M1, M2, M3, M4, M5, M6, M7;  // see DpViewSetup
  const uint32_t *src;                    // first pixel of the ROI
  int pitch, rw, rh;
  bool ok;
  bool tame;  // every source coordinate of the cell is finite and far inside the int32 range:
              // the texel loop may round with the magic-number add instead of cvt.rni (below)
};

__device__ __forceinline__ void dp_lane_setup(const DpViewDev *__restrict__ views, int n_views,
                                              const DpLevelSel &lv,
                                              int vid, bool active, int s, double inv_s,
                                              const DpFrame &f, DpLaneView &R) {
  const bool inr = active && f.ok && vid >= 0 && vid < n_views;
  const DpViewDev *V = views + (inr ? vid : 0);
  double u[4], v[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const double sgx = (c == 1 || c == 2) ? 1.0 : -1.0;  // corners (-,-) (+,-) (+,+) (-,+),
    const double sgy = (c >= 2) ? 1.0 : -1.0;            // patch.cpp:119-123
    const double X0 = xadd(xadd(f.p[0], sgx * f.ax[0]), sgy * f.ay[0]);
    const double X1 = xadd(xadd(f.p[1], sgx * f.ax[1]), sgy * f.ay[1]);
    const double X2 = xadd(xadd(f.p[2], sgx * f.ax[2]), sgy * f.ay[2]);
#if DP_LANE_PROJECT_OOL
    double uv[2];
    dp_project_ool(V->P, X0, X1, X2, uv);
    u[c] = uv[0];
    v[c] = uv[1];
#else
    dp_project(V->P, X0, X1, X2, u[c], v[c]);
#endif
  }
  if (lv.tab != nullptr) {  // read this view at the level its footprint asks for (dp_pick_level)
    const int up = dp_pick_level(xsub(u[1], u[0]), xsub(v[1], v[0]), xsub(u[3], u[0]),
                                 xsub(v[3], v[0]), lv.thr2, lv.up);
    if (up > 0) {  // P_l = diag(2^-l, 2^-l, 1) P: the projection scales exactly
      const double sc = __hiloint2double((1023 - up) << 20, 0);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        u[c] = xmul(u[c], sc);
        v[c] = xmul(v[c], sc);
      }
      V = lv.tab + ((size_t)up * n_views + (inr ? vid : 0));
    }
  }
  const int W = V->width, H = V->height;
  int tlx = W, tly = H, brx = 0, bry = 0;  // patch.cpp:126
  bool all_in = inr;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    all_in = all_in && (u[c] > 0) && (u[c] < (double)W) && (v[c] > 0) && (v[c] < (double)H);
    // ROI: tl = min ceil, br = max floor over the 4 corners (patch.cpp:137-140)
    tlx = min(tlx, __double2int_ru(u[c]));
    tly = min(tly, __double2int_ru(v[c]));
    brx = max(brx, __double2int_rd(u[c]));
    bry = max(bry, __double2int_rd(v[c]));
  }
  const int rw = brx - tlx, rh = bry - tly;
  double q[8];
#pragma unroll
  for (int c = 0; c < 4; ++c) {  // cv::Point2f, then `-= roi.x` in fp32 (patch.cpp:134, 148-151)
    q[2 * c] = (double)__fsub_rn((float)u[c], (float)tlx);
    q[2 * c + 1] = (double)__fsub_rn((float)v[c], (float)tly);
  }
  double m0, m1, m3, m4, m6, m7;
  const bool mapped = dp_quad_map(q[0], q[1], q[2], q[3], q[4], q[5], q[6], q[7], inv_s, m0, m1, m3,
                                  m4, m6, m7);
  const bool ok = all_in && rw > 0 && rh > 0 && mapped;  // optimization.cpp:45
  R.M0 = m0; R.M1 = m1; R.M2 = 32.0 * q[0]; R.M3 = m3; R.M4 = m4; R.M5 = 32.0 * q[1];
  R.M6 = m6; R.M7 = m7;
  R.src = V->img + (ok ? (size_t)tly * V->pitch_px + tlx : 0);
  R.pitch = V->pitch_px;
  R.rw = rw;
  R.rh = rh;
  R.ok = ok;
  // W = m6 x + m7 y + 1 is linear over the cell [0, s-1]^2, so its minimum is at a corner; the
  // numerators are bounded by their coefficient sums.  W >= 2^-6 and |num| < 2^22 give
  // |coordinate| < 2^28: rint() by adding 1.5 * 2^52 is then exact and nothing saturates.
  const double e = (double)(s - 1);
  const double wmin = 1.0 + fmin(0.0, m6 * e) + fmin(0.0, m7 * e);
  const double nx = (fabs(m0) + fabs(m1)) * e + fabs(R.M2), ny = (fabs(m3) + fabs(m4)) * e + fabs(R.M5);
  R.tame = ok && (wmin >= 0.015625) && (nx < 4194304.0) && (ny < 4194304.0);
}

// Blend of the four taps with OpenCV's 15-bit weights and BGR -> gray, as dp_texel_blend, with
// the horizontal pass as two-way dot products: one PRMT gathers (B0, B1, G0, G1) of a row's tap
// pair, IDP.2A multiplies by (32 - ax, ax) in 16-bit lanes.  Exact integers throughout.
#ifndef DP_LANE_DP2A
#define DP_LANE_DP2A 1
#endif
__device__ __forceinline__ int dp_lane_blend(uint32_t p00, uint32_t p01, uint32_t p10, uint32_t p11,
                                             unsigned wx1, unsigned wy1, unsigned &Bo, unsigned &Go,
                                             unsigned &Ro) {
#if DP_LANE_DP2A
  const unsigned wxp = wx1 * 0xffffu + 32u;  // (32 - wx1) | wx1 << 16
  const unsigned wy0 = 32u - wy1;
  const unsigned t0 = __byte_perm(p00, p01, 0x5140), u0 = __byte_perm(p00, p01, 0x0062);
  const unsigned t1 = __byte_perm(p10, p11, 0x5140), u1 = __byte_perm(p10, p11, 0x0062);
  const unsigned hb0 = __dp2a_lo(wxp, t0, 0u), hg0 = __dp2a_hi(wxp, t0, 0u), hr0 = __dp2a_lo(wxp, u0, 0u);
  const unsigned hb1 = __dp2a_lo(wxp, t1, 0u), hg1 = __dp2a_hi(wxp, t1, 0u), hr1 = __dp2a_lo(wxp, u1, 0u);
  const unsigned B = (hb0 * wy0 + hb1 * wy1 + 512u) >> 10;
  const unsigned G = (hg0 * wy0 + hg1 * wy1 + 512u) >> 10;
  const unsigned Rr = (hr0 * wy0 + hr1 * wy1 + 512u) >> 10;
  Bo = B; Go = G; Ro = Rr;
  return (int)((3735u * B + 19235u * G + 9798u * Rr + (1u << 14)) >> 15);
#else
  DpTaps t;
  t.p00 = p00; t.p01 = p01; t.p10 = p10; t.p11 = p11; t.wx1 = wx1; t.wy1 = wy1;
  uint32_t B, G, Rr;
  const int g = dp_texel_blend(t, B, G, Rr);
  Bo = B; Go = G; Ro = Rr;
  return g;
#endif
}

// The s x s texture of one view, walked by one lane: gray bytes row by row into grow[32 * y],
// integer moments in ma / mb.  STAGED: taps from the lane's tile (tp = tile + column offset,
// tpitch = its row pitch); else from the image itself (tp = ROI origin, tpitch = image pitch),
// rolled -- only for ROIs that do not fit the tile.
template <int S, bool STAGED, bool WRITE_TEX = false>
__device__ __forceinline__ void dp_lane_texels(const DpLaneView &R, const uint32_t *tp, int tpitch,
                                               uint2 *grow, unsigned &ma, unsigned &mb,
                                               uint8_t *__restrict__ tex_out = nullptr) {
  const int xmax = (R.rw - 1) << 5, ymax = (R.rh - 1) << 5;
  ma = 0;
  mb = 0;
  double yd = 0.0;
#pragma unroll 1
  for (int y = 0; y < S; ++y, yd += 1.0) {
    // the row's terms of source = (M0 x + M1 y + M2, M3 x + M4 y + M5) / (M6 x + M7 y + 1)
    const double A = fma(R.M1, yd, R.M2), B = fma(R.M4, yd, R.M5), Cw = fma(R.M7, yd, 1.0);
    unsigned glo = 0, ghi = 0;
#pragma unroll(STAGED ? S : 1)
    for (int x = 0; x < S; ++x) {
      const double xd = (double)x;
      const double Wd = fma(R.M6, xd, Cw);
      const double r = dp_rcp(Wd);  // W == 0 -> NaN coordinates -> cvt gives 0 (see dp_texel_fetch)
      const double fX = fma(R.M0, xd, A) * r;
      const double fY = fma(R.M3, xd, B) * r;
      // saturate_cast<int>(cvRound(.)), half to even.  The staged loop only runs for "tame" views
      // (DpLaneView::tame: |coordinate| < 2^28, finite), where adding 1.5 * 2^52 leaves the
      // round-to-nearest-even integer in the low word -- one FP64 add instead of a conversion
      // on the (quarter-rate) XU pipe; the generic loop keeps the saturating cvt.rni.
      int Xi, Yi;
      if (STAGED) {
        Xi = __double2loint(fX + 6755399441055744.0);
        Yi = __double2loint(fY + 6755399441055744.0);
      } else {
        Xi = __double2int_rn(fX);
        Yi = __double2int_rn(fY);
      }
      const int Xc = max(min(Xi, xmax), 0), Yc = max(min(Yi, ymax), 0);  // BORDER_REPLICATE
      const int x0 = Xc >> 5, y0 = Yc >> 5;
      const uint32_t *r0 = tp + (unsigned)(y0 * tpitch + x0), *r1 = r0 + tpitch;
      uint32_t p00, p01, p10, p11;
      if (STAGED) {
        p00 = r0[0]; p01 = r0[1]; p10 = r1[0]; p11 = r1[1];
      } else {
        p00 = __ldg(r0); p01 = __ldg(r0 + 1); p10 = __ldg(r1); p11 = __ldg(r1 + 1);
      }
      unsigned Bc, Gc, Rc;
      const unsigned gray = (unsigned)dp_lane_blend(p00, p01, p10, p11, (unsigned)(Xc & 31),
                                                    (unsigned)(Yc & 31), Bc, Gc, Rc);
      if (WRITE_TEX) {
        uint8_t *t = tex_out + 3 * (y * S + x);
        t[0] = (uint8_t)Bc; t[1] = (uint8_t)Gc; t[2] = (uint8_t)Rc;
      }
      ma += gray;
      mb += gray * gray;
      if (x < 4) glo |= gray << (8 * (x & 3));
      else ghi |= gray << (8 * (x & 3));
    }
    grow[32 * y] = make_uint2(glo, ghi);
  }
}

// sum_i fl32(a_i - fl32(mean_a)) * fl32(b_i - fl32(mean_b)) in fp64, texel by texel in row-major
// order (error_measurements.cpp:54 on CV_32F operands, Mat::dot's order).
// (Measured and dropped: an fp64-only variant for textures whose fp32 centring is provably exact
// -- no I2F / F2F on the XU pipe -- with this loop as the fallback: 5 % of the lane-views fail
// the exactness test, so nearly every warp ran both loops; refine 3.31 -> 3.09 G evals/s.)
template <int S>
__device__ __forceinline__ double dp_lane_numerator(const uint2 *ga, const uint2 *gb, float mfa,
                                                    float mfb) {
  double num = 0.0;
#pragma unroll 1
  for (int y = 0; y < S; ++y) {
    const uint2 a = ga[32 * y], b = gb[32 * y];
#pragma unroll
    for (int x = 0; x < S; ++x) {
      const unsigned av = ((x < 4 ? a.x : a.y) >> (8 * (x & 3))) & 0xffu;
      const unsigned bv = ((x < 4 ? b.x : b.y) >> (8 * (x & 3))) & 0xffu;
      const float da = __fsub_rn((float)av, mfa), db = __fsub_rn((float)bv, mfb);
      num = fma((double)da, (double)db, num);  // the product of two floats is exact in fp64
    }
  }
  return num;
}

// Stage the lane's ROI into its tile: one 16-byte cp.async per 4-pixel piece of a row, from the
// ROI origin rounded down to 4 pixels (16-byte alignment; the taps carry the 0-3 pixel offset).
// The tile row pitch is 4 * pieces pixels.  Returns whether the ROI fits the tile (else the taps
// come straight from the image).  Completion: cp.async.wait_all (the tile is lane-private).
// (Measured and dropped: lane pairs splitting each other's rows so that neighbouring pieces
// share an L2 request -- 4 shuffles and two nested loops per view cost 17.7 of 172 warp
// instructions per evaluation, ncu r2l; this form is ~2.)
__device__ __forceinline__ bool dp_lane_stage(const DpLaneView &R, bool want, uint32_t *tile,
                                              int &tpitch, int &xoff) {
  // image rows are 128-byte aligned: the pixel offset of the ROI inside its 16-byte quad is
  // visible in the pointer
  xoff = (int)((reinterpret_cast<uintptr_t>(R.src) >> 2) & 3u);
  const int pieces = (R.rw + xoff + 3) >> 2;
  tpitch = 4 * pieces;
  const bool fits = want && tpitch * R.rh <= DP_LTILE;
  if (fits) {
    const uint32_t *g = R.src - xoff;
    uint32_t *d = tile;
    for (int r = 0; r < R.rh; ++r, g += R.pitch, d += tpitch) {
      dp_cp_async16(d, g);
      if (pieces > 1) dp_cp_async16(d + 4, g + 4);
      if (pieces > 2) dp_cp_async16(d + 8, g + 8);
      for (int q = 3; q < pieces; ++q) dp_cp_async16(d + 4 * q, g + 4 * q);
    }
  }
  return fits;
}

// K1 + K2 with one patch per lane: GetProjectedTextures + NCCScore for every visible view and,
// fused, FilterByErrorMeasurement's erase loop (optimization.cpp:98-132) -- "erase visible[k-1]
// iff the score of visible[k] is low; the last entry always survives" (see dp_score_kernel),
// which a lane applies to its own patch in place, sequentially.
template <int S, bool WRITE_TEX, bool FILTER>
__global__ void __launch_bounds__(DP_LWARPS * 32, DP_LMINCTA)
dp_score_lane_kernel(DpScoreArgs a, const int32_t *__restrict__ order) {
  __shared__ __align__(16) uint32_t s_tile[DP_LWARPS][32 * DP_LTSTR + 32];
  __shared__ uint2 s_gray[DP_LWARPS][2][S][32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int npx = S * S;
  const double scale = 1.0 / (double)npx;
  const double inv_s = 1.0 / (double)S;
  uint32_t *warp_tiles = s_tile[warp];
  uint2 *gA = &s_gray[warp][0][0][lane], *gB = &s_gray[warp][1][0][lane];
  const long long slot = ((long long)blockIdx.x * DP_LWARPS + warp) * 32 + lane;
  const bool have = slot < a.p.n;
  const long long i = have ? (order ? (long long)order[slot] : slot) : 0;
  const int nv = have ? min(a.p.nvis[i], a.p.vstride) : 0;
  const int ref = a.p.ref[i];
  const bool ref_ok = ref >= 0 && ref < a.p.n_views;
  double n[3] = {(double)a.p.nrm[3 * i], (double)a.p.nrm[3 * i + 1], (double)a.p.nrm[3 * i + 2]};
  double p[3] = {(double)a.p.pos[3 * i], (double)a.p.pos[3 * i + 1], (double)a.p.pos[3 * i + 2]};
  const double pc[3] = {p[0], p[1], p[2]};
  if (a.trial_nrm) { n[0] = a.trial_nrm[3 * i]; n[1] = a.trial_nrm[3 * i + 1]; n[2] = a.trial_nrm[3 * i + 2]; }
  if (a.trial_pos) { p[0] = a.trial_pos[3 * i]; p[1] = a.trial_pos[3 * i + 1]; p[2] = a.trial_pos[3 * i + 2]; }
  int32_t *vis = a.p.vis + (size_t)i * a.p.vstride;
  float *ncc = a.ncc ? a.ncc + (size_t)i * a.p.vstride : nullptr;
  uint8_t *tex = WRITE_TEX ? a.tex + (size_t)i * a.p.vstride * npx * 3 : nullptr;
  uint8_t *valid = a.valid ? a.valid + (size_t)i * a.p.vstride : nullptr;
  DpFrame f;
  dp_make_frame(a.p.views + (ref_ok ? ref : 0), S, n, p, pc, f);
  if (!ref_ok) f.ok = false;  // every texture empty (optimization.cpp:45)
  const int nvmax = __reduce_max_sync(DP_FULL, nv);
  unsigned a1 = 0, a2 = 0;
  bool a_ok = false;
  float mfa = 0.f;
  int wcur = 0, prev = -1;
  const double thr = a.thr;
#pragma unroll 1
  for (int k = 0; k < nvmax; ++k) {
    const bool active = k < nv;
    const int vid = active ? vis[k] : -1;
    DpLaneView R;
    dp_lane_setup(a.p.views, a.p.n_views, a.p.lv, vid, active, S, inv_s, f, R);
    int tpitch, xoff;
    const bool staged = dp_lane_stage(R, R.ok && R.tame, warp_tiles + lane * DP_LTSTR, tpitch, xoff);
    dp_cp_async_wait_all();
    unsigned s1 = 0, s2 = 0;
    uint2 *g = (k == 0) ? gA : gB;
    if (R.ok) {
      uint8_t *to = WRITE_TEX ? tex + (size_t)k * npx * 3 : nullptr;
      if (staged)
        dp_lane_texels<S, true, WRITE_TEX>(R, warp_tiles + lane * DP_LTSTR + xoff, tpitch, g, s1, s2, to);
      else
        dp_lane_texels<S, false, WRITE_TEX>(R, R.src, R.pitch, g, s1, s2, to);
    }
    if (valid != nullptr && active) valid[k] = R.ok ? 1 : 0;
    if (k == 0) {
      a1 = s1;
      a2 = s2;
      a_ok = R.ok;
      mfa = (float)xmul((double)a1, scale);
    } else if (active) {
      double score = -1.0;  // an empty texture (error_measurements.cpp:38-40)
      if (R.ok && a_ok) {
        const float mfb = (float)xmul((double)s1, scale);
        const double num = dp_lane_numerator<S>(gA, gB, mfa, mfb);
        score = dp_ncc_finish(a1, a2, s1, s2, num, scale, npx);
      }
      if (ncc != nullptr) ncc[k] = (float)score;
      if (FILTER && !(score < thr)) vis[wcur++] = prev;  // keep original entry k-1
    }
    if (active) prev = vid;  // original entry k, read before position k-1 or above is written
  }
  if (FILTER && have) {
    bool kept = false;
    if (nv >= 2) {  // scores.size() > 0 (optimization.cpp:113)
      vis[wcur++] = prev;  // the last entry always survives
      for (int k = wcur; k < nv; ++k) vis[k] = -1;
      a.p.nvis[i] = wcur;
      kept = wcur >= a.min_visible;  // optimization.cpp:127
    }
    a.keep[i] = kept ? 1 : 0;
  }
}

template <int S>
__global__ void __launch_bounds__(DP_LWARPS * 32, DP_LMINCTA) dp_refine_lane_kernel(DpRefineArgs a) {
  __shared__ DpLaneShared<S> sh;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int npx = S * S;
  const double scale = 1.0 / (double)npx;  // cv::meanStdDev: mean = sum * (1/N)
  const double inv_s = 1.0 / (double)S;
#if DP_LANE_NM_LOCAL
  double nm_words[DP_NM_WORDS];
  const DpNmLane NM{nm_words};
#else
  const DpNmLane NM{&sh.nm[warp][0][lane]};
#endif
  uint32_t *warp_tiles = sh.tile[warp];
  uint2 *gA = &sh.gray[warp][0][0][lane], *gB = &sh.gray[warp][1][0][lane];
#pragma unroll 1
  for (int k = 0; k < DP_NM_WORDS; ++k) NM.b[DP_NM_STRIDE * k] = 0.0;  // defined values for idle lanes
  bool have = false, exhausted = false;
  long long i = 0;
  int nv = 0, ref = 0;
  bool ref_ok = false;
  const int32_t *vis = a.p.vis;
  int state = NMG_INIT, idx = 0, fcount = 4, ilo = 0, ihi = 0;
  double y_lo = 0.0, y_nhi = 0.0, y_hi = 0.0;
  double n0[3] = {0, 0, 0}, p0[3] = {0, 0, 0}, c3[3] = {0, 0, 0};
#pragma unroll 1
  for (;;) {
    // ---- 1. lanes without a patch take the next ones from the work counter --------------------
    // (a.order hands the patches out by descending view count, so the lanes of a warp mostly
    // evaluate the same number of views.  Measured and dropped: one queue per view-count class
    // with every warp staying inside a class -- no lane ever idles through a neighbour's extra
    // view, but the longest-first order is lost and the classes with many views finish last:
    // refine 3.61 -> 3.35 G evals/s.)
    for (;;) {
      const unsigned need = __ballot_sync(DP_FULL, !have && !exhausted);
      if (need == 0) break;
      const int leader = __ffs(need) - 1;
      unsigned int base = 0;
      if (lane == leader) base = atomicAdd(a.work_counter, (unsigned)__popc(need));
      base = __shfl_sync(DP_FULL, base, leader);
      if (!have && !exhausted) {
        const unsigned int iu = base + __popc(need & ((1u << lane) - 1u));
        if (iu >= (unsigned int)a.p.n) {
          exhausted = true;
        } else {
          i = a.order ? (long long)a.order[iu] : (long long)iu;
          if (a.mask != nullptr && a.mask[i] == 0) {  // removed by Seed::RemovePatches
            if (a.evals) a.evals[i] = 0;
          } else {
            nv = min(a.p.nvis[i], a.p.vstride);
            ref = a.p.ref[i];
            ref_ok = ref >= 0 && ref < a.p.n_views;
            vis = a.p.vis + (size_t)i * a.p.vstride;
            const double *C = a.p.views[ref_ok ? ref : 0].center;
            // createInitialSimplex: v_i = x0 + step_{i-1}/2 e_{i-1}, then v_0 = x0 - step/2; x0 = 0
#pragma unroll
            for (int j = 0; j < 3; ++j) {
              const double h = xmul(0.5, a.step[j]);
              NM.P(0, j) = xsub(0.0, h);
              NM.pt(j) = xsub(0.0, h);
#pragma unroll
              for (int v = 1; v < 4; ++v) NM.P(v, j) = (v - 1 == j) ? xadd(0.0, h) : 0.0;
              n0[j] = (double)a.p.nrm[3 * i + j];
              p0[j] = (double)a.p.pos[3 * i + j];
              c3[j] = C[j];
            }
            state = NMG_INIT;
            idx = 0;
            fcount = 4;
            ilo = ihi = 0;
            have = true;
          }
        }
      }
    }
    if (!__any_sync(DP_FULL, have)) break;
    // ---- 2. one objective evaluation per lane, in lockstep -------------------------------------
    // PatchOptimizationOpenCVFunctor::calc (optimization_opencv.cpp:14-39)
    double n[3], p[3];
    {
      const double x[3] = {NM.pt(0), NM.pt(1), NM.pt(2)};
      dp_unparametrize_l(c3, n0, p0, x, n, p);
    }
    const int nv_eval = (have && nv >= 2 && ref_ok) ? nv : 0;
    DpFrame f;
    dp_make_frame(a.p.views + (ref_ok ? ref : 0), S, n, p, p0, f);  // corners stay around p0
    const int nvmax = __reduce_max_sync(DP_FULL, nv_eval);
    double sum = 0.0;
    unsigned a1 = 0, a2 = 0;
    bool a_ok = false;
    float mfa = 0.f;
#pragma unroll 1
    for (int k = 0; k < nvmax; ++k) {
      const bool active = k < nv_eval;
      DpLaneView R;
      dp_lane_setup(a.p.views, a.p.n_views, a.p.lv, active ? vis[k] : -1, active, S, inv_s, f, R);
      int tpitch, xoff;
      const bool staged = dp_lane_stage(R, R.ok && R.tame, warp_tiles + lane * DP_LTSTR, tpitch, xoff);
      dp_cp_async_wait_all();
      unsigned s1 = 0, s2 = 0;
      uint2 *g = (k == 0) ? gA : gB;
#ifdef DP_DEBUG_TRACE
      if (a.trace) {  // lane-slot census of this view step: [0] textured, [1] empty texture,
                      // [2] patch has fewer views, [3] tail (no patch left), [4] unstaged
        const unsigned m0 = __ballot_sync(DP_FULL, R.ok), m1 = __ballot_sync(DP_FULL, active && !R.ok),
                       m2 = __ballot_sync(DP_FULL, have && !active), m3 = __ballot_sync(DP_FULL, !have),
                       m4 = __ballot_sync(DP_FULL, R.ok && !staged);
        if (lane == 0) {
          unsigned long long *c = reinterpret_cast<unsigned long long *>(a.trace) + (size_t)8 * a.p.n;
          atomicAdd(c + 0, (unsigned long long)__popc(m0));
          atomicAdd(c + 1, (unsigned long long)__popc(m1));
          atomicAdd(c + 2, (unsigned long long)__popc(m2));
          atomicAdd(c + 3, (unsigned long long)__popc(m3));
          atomicAdd(c + 4, (unsigned long long)__popc(m4));
          atomicAdd(c + 5, 32ull);
        }
      }
#endif
      if (R.ok) {
        if (staged)
          dp_lane_texels<S, true>(R, warp_tiles + lane * DP_LTSTR + xoff, tpitch, g, s1, s2);
        else
          dp_lane_texels<S, false>(R, R.src, R.pitch, g, s1, s2);
      }
      if (k == 0) {
        a1 = s1;
        a2 = s2;
        a_ok = R.ok;
        mfa = (float)xmul((double)a1, scale);
      } else if (active) {
        double score = -1.0;  // an empty texture (error_measurements.cpp:38-40)
        if (R.ok && a_ok) {
          const float mfb = (float)xmul((double)s1, scale);
          const double num = dp_lane_numerator<S>(gA, gB, mfa, mfb);
          score = dp_ncc_finish(a1, a2, s1, s2, num, scale, npx);
        }
        sum = xadd(sum, xsub(1.0, score));  // std::accumulate in view order
      }
    }
    const double fval = nv_eval ? sum / (double)(nv - 1) : 2.0;  // scores.size() == 0 -> 2
#ifdef DP_DEBUG_TRACE
    if (have && a.trace) {
      int e = 0;
      while (e < 8 && a.trace[8 * i + e] != -7.0) ++e;
      if (e < 8) a.trace[8 * i + e] = fval;
    }
#endif
    // ---- 3. Nelder-Mead bookkeeping of each lane (diverges by solver state, short) -------------
    if (have) {
      if (nml_step(NM, state, idx, fcount, ilo, ihi, y_lo, y_nhi, y_hi, fval, a.eps, a.max_evals,
                   a.step)) {
        // best vertex -> one more trip through UnparametrizePatch, then write back;
        // SetNormal / SetPosition store fp32 (patch.h:38-53)
        double nb[3], pb[3];
        const double xb[3] = {NM.pt(0), NM.pt(1), NM.pt(2)};
        dp_unparametrize_l(c3, n0, p0, xb, nb, pb);
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          if (ref_ok) {
            a.p.nrm[3 * i + j] = (float)nb[j];
            a.p.pos[3 * i + j] = (float)pb[j];
          }
          if (a.xbest) a.xbest[3 * i + j] = NM.pt(j);
        }
        if (a.evals) a.evals[i] = fcount;
        have = false;
      }
    }
  }
}
