// dp_group.cuh -- score / filter / refine kernels with SEVERAL patches per warp (s <= 16).
//
// With one warp per patch (dp_kernels.cuh) more than half of the warp instructions of an
// objective evaluation are not texel work: UnparametrizePatch, the patch frame and the
// Nelder-Mead state machine are scalar (all 32 lanes compute the same value), the per-view
// set-up fills 8 lane slots of which ~5 are used, the reductions run over 32 lanes for 49
// texels, and the texel passes themselves use 49 of 64 lane slots.  Here a patch owns a group
// of GL lanes (4 for s <= 8) and a warp advances 32 / GL patches in lockstep:
//   * scalar work is done once per instruction for 8 patches;
//   * a set-up pass handles one view of each of the 8 patches (4 corners x 8 slots, all used);
//   * a 7x7 texture takes 13 passes of 4 lanes (52 slots for 49 texels instead of 64);
//   * reductions are 2-step shuffle trees inside the group.
// The warp's control flow stays uniform: every iteration of the refine kernel's main loop is
// one objective evaluation for each of the warp's patches (at its own simplex point), the view
// loop runs to the largest view count in the warp (patches are handed out sorted by view
// count, so they usually agree), and only the short Nelder-Mead bookkeeping diverges per
// group.  A group whose patch has converged writes it back and takes the next patch from the
// work counter.
//
// The arithmetic is the one of dp_kernels.cuh / dp_device.cuh (same functions); only the
// order of the fp64 partial sums of the NCC numerator differs (group tree), ~1e-16: scores,
// keep bits, visible sets, evaluation counts and refined fp32 geometry are bit-identical.
#pragma once
#include "dp_kernels.cuh"

// Group geometry per cell size: GL lanes per patch, NP = ceil(s^2 / GL) texel passes, and the
// staging tile of a group (TW x TH pixels).  Measured on B200 at s = 7: 4 lanes per patch
// 2.46 G evals/s, 8 lanes 2.31, 16 lanes 2.12; larger cells take wider groups so that the
// per-lane texel arrays and the tiles stay small.
template <int GL_, int NP_, int TW_, int TH_>
struct DpGroupCfg {
  static constexpr int GL = GL_;            // lanes per patch
  static constexpr int NP = NP_;            // texel passes
  static constexpr int TW = TW_, TH = TH_;  // staging tile, pixels (TW a multiple of 4)
  static constexpr int GROUPS = 32 / GL;    // patches per warp
  static constexpr int GROUND = GL;         // views per round of a group: one per lane in phase C
  static constexpr int TSTRIDE = TW * TH + 4;  // +4 words: stagger the groups over the banks
};
// s <= 8: 4 lanes, 16x8 tile; s <= 12: 8 lanes, 16x16; s <= 16: 16 lanes, 32x24
template <int S>
using DpCfgFor = DpGroupCfg<(S <= 8 ? 4 : (S <= 12 ? 8 : 16)),
                            (S * S + (S <= 8 ? 4 : (S <= 12 ? 8 : 16)) - 1) / (S <= 8 ? 4 : (S <= 12 ? 8 : 16)),
                            (S <= 12 ? 16 : 32), (S <= 8 ? 8 : (S <= 12 ? 16 : 24))>;
// Largest cells served by the group kernels.  Measured on B200, group vs warp-per-patch kernel,
// ~7 views per patch: s = 11 score +18 %, refine +7 %; s = 16 score +33 %, refine -1 %; with
// 24-49 views per patch (64-view scene, expansion at s = 11) the refine group kernel is 13 %
// slower (nothing scalar left to amortise, lockstep over unequal view counts), the score
// kernel neutral.  Hence: score / filter up to 16, refine up to 8.
#ifndef DP_GROUP_MAX_CELL_SCORE
#define DP_GROUP_MAX_CELL_SCORE 16
#endif
#ifndef DP_GROUP_MAX_CELL_REFINE
#define DP_GROUP_MAX_CELL_REFINE 8
#endif

#ifndef DP_GWARPS
#define DP_GWARPS 4              // warps per CTA of the group kernels
#endif
#ifndef DP_GMINCTA
#define DP_GMINCTA 4             // resident CTAs per SM asked for (register cap 128)
#endif

struct DpGroupLane {
  int sub;         // lane index inside the group
  int base;        // first lane of the group
  unsigned mask;   // the group's lanes
  bool leader;     // sub == 0
};

// Texel coordinates (x, y) come as doubles from a per-CTA table in shared memory: texel
// sub + GL*j of a lane is one LDS.128 at a constant offset from the lane's base pointer.

template <int GL, typename T>
__device__ __forceinline__ T dp_group_sum(T v, unsigned mask) {
#pragma unroll
  for (int o = GL / 2; o > 0; o >>= 1) v += __shfl_xor_sync(mask, v, o);
  return v;
}

// ---- Nelder-Mead helpers, group flavour (state of the group's patch in shared memory, all
// lanes of the group compute, the leader writes between two group barriers) ---------------
__device__ __forceinline__ void nmg_store3(double *dst, const double v[3], const DpGroupLane &L) {
  __syncwarp(L.mask);
  if (L.leader) { dst[0] = v[0]; dst[1] = v[1]; dst[2] = v[2]; }
  __syncwarp(L.mask);
}
__device__ __forceinline__ void nmg_coord_sum(DpNelderMead &S, const DpGroupLane &L) {
  double t[3];
#pragma unroll
  for (int j = 0; j < 3; ++j)
    t[j] = xadd(xadd(xadd(xadd(0.0, S.P[0][j]), S.P[1][j]), S.P[2][j]), S.P[3][j]);
  nmg_store3(S.cs, t, L);
}
__device__ __forceinline__ void nmg_try_point(DpNelderMead &S, const DpGroupLane &L, int ihi,
                                              double alpha_) {
  const double al = (1.0 - alpha_) / 3.0;
  const double be = xsub(al, alpha_);
  double pt[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) pt[j] = xsub(xmul(S.cs[j], al), xmul(S.P[ihi][j], be));
  nmg_store3(S.pt, pt, L);
}
__device__ __forceinline__ void nmg_replace(DpNelderMead &S, const DpGroupLane &L, int ihi,
                                            const double q[3], double yq) {
  __syncwarp(L.mask);
  if (L.leader) {
    S.P[ihi][0] = q[0]; S.P[ihi][1] = q[1]; S.P[ihi][2] = q[2];
    S.y[ihi] = yq;
  }
  __syncwarp(L.mask);
  nmg_coord_sum(S, L);
}
__device__ __forceinline__ void nmg_shrink_vertex(DpNelderMead &S, const DpGroupLane &L, int idx,
                                                  int ilo) {
  double pt[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) pt[j] = xmul(0.5, xadd(S.P[idx][j], S.P[ilo][j]));
  __syncwarp(L.mask);
  if (L.leader) {
#pragma unroll
    for (int j = 0; j < 3; ++j) { S.P[idx][j] = pt[j]; S.pt[j] = pt[j]; }
  }
  __syncwarp(L.mask);
}

enum { NMG_INIT, NMG_REFLECT, NMG_EXPAND, NMG_CONTRACT, NMG_SHRINK };

// One step of cv::DownhillSolver's state machine (the same decision tree as in
// dp_refine_kernel): consumes the objective value of S.pt, leaves the next point to
// evaluate in S.pt; returns true when the solver stops (S.pt = best vertex).
__device__ __forceinline__ bool nmg_step(DpNelderMead &S, const DpGroupLane &L, int &state, int &idx,
                                         int &fcount, int &ilo, int &ihi, double fval, double eps,
                                         int max_evals) {
  bool decide = false;
  if (state == NMG_INIT) {
    __syncwarp(L.mask);
    if (L.leader) S.y[idx] = fval;
    __syncwarp(L.mask);
    if (++idx < 4) {
      const double q[3] = {S.P[idx][0], S.P[idx][1], S.P[idx][2]};
      nmg_store3(S.pt, q, L);
    } else {
      nmg_coord_sum(S, L);
      decide = true;
    }
  } else if (state == NMG_REFLECT) {
    const double q[3] = {S.pt[0], S.pt[1], S.pt[2]};
    const double y_lo = S.y_lo, y_nhi = S.y_nhi;
    __syncwarp(L.mask);
    if (L.leader) {
      S.pa[0] = q[0]; S.pa[1] = q[1]; S.pa[2] = q[2];
      S.y_alpha = fval;
    }
    __syncwarp(L.mask);
    if (fval < y_nhi) {
      if (fval < y_lo) {  // better than the best: try twice as far
        state = NMG_EXPAND;
        nmg_try_point(S, L, ihi, -2.0);
        ++fcount;
      } else {
        nmg_replace(S, L, ihi, q, fval);  // replacePoint(alpha = -1)
        decide = true;
      }
    } else {
      state = NMG_CONTRACT;
      nmg_try_point(S, L, ihi, 0.5);
      ++fcount;
    }
  } else if (state == NMG_EXPAND) {
    const double y_alpha = S.y_alpha;
    const bool better = fval < y_alpha;
    const double q[3] = {better ? S.pt[0] : S.pa[0], better ? S.pt[1] : S.pa[1],
                         better ? S.pt[2] : S.pa[2]};
    nmg_replace(S, L, ihi, q, better ? fval : y_alpha);
    decide = true;
  } else if (state == NMG_CONTRACT) {
    if (fval < S.y_hi) {
      const double q[3] = {S.pt[0], S.pt[1], S.pt[2]};
      nmg_replace(S, L, ihi, q, fval);
      decide = true;
    } else {  // shrink every vertex but the best halfway towards it
      state = NMG_SHRINK;
      idx = (ilo == 0) ? 1 : 0;
      nmg_shrink_vertex(S, L, idx, ilo);
    }
  } else {  // NMG_SHRINK
    __syncwarp(L.mask);
    if (L.leader) S.y[idx] = fval;
    __syncwarp(L.mask);
    ++idx;
    if (idx == ilo) ++idx;
    if (idx < 4) {
      nmg_shrink_vertex(S, L, idx, ilo);
    } else {
      fcount += 3;
      nmg_coord_sum(S, L);
      decide = true;
    }
  }
  if (!decide) return false;
  // ---- find worst, next-to-worst and best vertices; stop test ----------------------------
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
  if (range <= eps || error <= eps || fcount >= max_evals) {
    const double q[3] = {S.P[ilo][0], S.P[ilo][1], S.P[ilo][2]};
    nmg_store3(S.pt, q, L);  // best vertex -> x
    return true;
  }
  __syncwarp(L.mask);
  if (L.leader) { S.y_lo = ylo; S.y_nhi = ynhi; S.y_hi = yhi; }
  __syncwarp(L.mask);
  state = NMG_REFLECT;  // reflect the worst point about the centroid of the others
  nmg_try_point(S, L, ihi, -1.0);
  ++fcount;
  return false;
}

// Optimization::UnparametrizePatch (optimization.cpp:78-96); the group's lanes 0 and 1 compute
// the sincos of roll and pitch.  `mask` must cover every lane that executes the call.
__device__ __forceinline__ void dp_unparametrize_g(const double C[3], const double n0[3],
                                                   const double p0[3], double depth, double roll,
                                                   double pitch, double n[3], double p[3],
                                                   const DpGroupLane &L, unsigned mask) {
  const double k = xadd(1.0, depth);
#pragma unroll
  for (int j = 0; j < 3; ++j) p[j] = xadd(C[j], xmul(k, xsub(p0[j], C[j])));
  double sv, cv;
  sincos((L.sub & 1) ? pitch : roll, &sv, &cv);
  const double sa = __shfl_sync(mask, sv, L.base), ca = __shfl_sync(mask, cv, L.base);
  const double sb = __shfl_sync(mask, sv, L.base + 1), cb = __shfl_sync(mask, cv, L.base + 1);
  n[0] = xadd(xmul(cb, n0[0]), xmul(-sb, n0[2]));
  n[1] = xadd(xadd(xmul(xmul(sa, sb), n0[0]), xmul(ca, n0[1])), xmul(xmul(cb, sa), n0[2]));
  n[2] = xadd(xadd(xmul(xmul(ca, sb), n0[0]), xmul(-sa, n0[1])), xmul(xmul(ca, cb), n0[2]));
}

// ---- rolled, software-pipelined texel pass ---------------------------------------------
// The fully unrolled pass of dp_view_texture is NP x ~85 instructions per view (18 KB of
// straight-line code at NP = 13): ncu showed 14 % of the warp samples stalled on instruction
// fetch and 23 % on the tap loads.  Here the pass is a loop: the taps of texel j+1 are
// requested before texel j is blended (their latency hides behind ~45 instructions of
// arithmetic), the gray values go to a byte per (pass, lane) in shared memory (they are
// needed again once the mean is known) and the integer moments accumulate on the fly.
// Per-group staging tile: TH rows of TW pixels, filled with 16-byte cp.async copies (one per
// row and 4-pixel piece).  The window starts at the ROI origin rounded down to 4 pixels
// (16-byte alignment), the taps carry the 0-3 pixel offset.
#ifndef DP_GROUP_STAGE
#define DP_GROUP_STAGE 1
#endif

__device__ __forceinline__ void dp_cp_async16(uint32_t *dst, const uint32_t *src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dp_smem_u32(dst)), "l"(src)
               : "memory");
}
__device__ __forceinline__ void dp_cp_async_wait_all() {
  asm volatile("cp.async.wait_all;" ::: "memory");
}

struct DpTaps {
  uint32_t p00, p01, p10, p11;
  unsigned wx1, wy1;
};

struct DpWarpConsts {  // one view, from its set-up record
  double M0, M1, M2, M3, M4, M5, M6, M7;
  const uint32_t *src;
  int pitch, xmax, ymax;
};

template <bool STAGED, int TW>
__device__ __forceinline__ void dp_texel_fetch(const DpWarpConsts &c, const double2 xy, DpTaps &t) {
  const double x = xy.x, y = xy.y;
  const double Wd = fma(c.M6, x, fma(c.M7, y, 1.0));
  // W ? INTER_TAB_SIZE / W : 0 -- without a select: W == 0 makes r (inf, then NaN after the
  // Newton steps), fX and fY NaN, and cvt.rni.s32.f64 turns NaN into 0 = what the reference's
  // zero scale gives (the numerators are finite, dp_setup_views checks the map)
  const double r = dp_rcp(Wd);
  const double fX = fma(c.M0, x, fma(c.M1, y, c.M2)) * r;
  const double fY = fma(c.M3, x, fma(c.M4, y, c.M5)) * r;
  const int Xi = __double2int_rn(fX);  // saturate_cast<int>(cvRound), half to even
  const int Yi = __double2int_rn(fY);
  // BORDER_REPLICATE as a clamp of the 1/32-px coordinate (see dp_view_texture)
  // (xmax, ymax >= 0; min first, then max with 0: one VIMNMX.RELU per coordinate)
  const int Xc = max(min(Xi, c.xmax), 0), Yc = max(min(Yi, c.ymax), 0);
  const int x0 = Xc >> 5, y0 = Yc >> 5;  // INTER_BITS = 5
  t.wx1 = (unsigned)(Xc & 31);
  t.wy1 = (unsigned)(Yc & 31);
  if (STAGED) {  // c.src = the group's tile (+ column offset), rows of TW pixels
    const uint32_t *r0 = c.src + (y0 * TW + x0);
    t.p00 = r0[0]; t.p01 = r0[1];
    t.p10 = r0[TW]; t.p11 = r0[TW + 1];
  } else {
    const uint32_t *r0 = c.src + (unsigned)(y0 * c.pitch + x0), *r1 = r0 + c.pitch;
    t.p00 = __ldg(r0); t.p01 = __ldg(r0 + 1);
    t.p10 = __ldg(r1); t.p11 = __ldg(r1 + 1);
  }
}

__device__ __forceinline__ int dp_texel_blend(const DpTaps &t, uint32_t &Bo, uint32_t &Go,
                                              uint32_t &Ro) {
  const uint32_t wx1 = t.wx1, wx0 = 32u - wx1, wy1 = t.wy1, wy0 = 32u - wy1;
  const uint32_t br0 = (t.p00 & 0x00ff00ffu) * wx0 + (t.p01 & 0x00ff00ffu) * wx1;  // B | R<<16
  const uint32_t br1 = (t.p10 & 0x00ff00ffu) * wx0 + (t.p11 & 0x00ff00ffu) * wx1;
  // G << 8 = the blend of the whole pixel word minus its B | R part: the x byte of a pixel is 0
  // and every partial sum is <= 255 * 32, so the word blend is < 2^30 and the difference is
  // exact (two masks less per tap pair).  A tap with weight 0 may be any word.
  const uint32_t g0 = (t.p00 * wx0 + t.p01 * wx1) - br0;
  const uint32_t g1 = (t.p10 * wx0 + t.p11 * wx1) - br1;
  const uint32_t B = ((br0 & 0xffffu) * wy0 + (br1 & 0xffffu) * wy1 + 512u) >> 10;
  const uint32_t Rr = ((br0 >> 16) * wy0 + (br1 >> 16) * wy1 + 512u) >> 10;
  const uint32_t G = (g0 * wy0 + g1 * wy1 + (512u << 8)) >> 18;
  Bo = B;
  Go = G;
  Ro = Rr;
  // cv::cvtColor(BGR2GRAY), 8U: 15-bit fixed point
  return (int)((3735u * B + 19235u * G + 9798u * Rr + (1u << 14)) >> 15);
}

#ifndef DP_TEXEL_UNROLL
#define DP_TEXEL_UNROLL 2  // measured with the staged tile: 1: 2.62, 2: 2.76, 3: 2.53, 4: 2.65, 6: 2.55 G evals/s
#endif
// (Measured and rejected: prefetch.global.L1 of the next view's ROI rows while the current
// view is computed, -5 %.)

// gs: this lane's column of the warp's gray buffer, gs[32 * j] = texel j (0 past the patch).
template <typename C, bool STAGED, bool WRITE_TEX>
__device__ __forceinline__ void dp_texel_loop(const DpWarpConsts &c, int npx, const double2 *txy,
                                              int sub, uint8_t *gs, unsigned &ma, unsigned &mb,
                                              uint8_t *__restrict__ tex_out) {
  constexpr int NP = C::NP, GL = C::GL;
  DpTaps cur;
  dp_texel_fetch<STAGED, C::TW>(c, txy[0], cur);
  constexpr int kUnroll = DP_TEXEL_UNROLL;
  // passes 0 .. NP-2 hold texels of the patch in every lane (GL * (NP - 1) < npx); only the last
  // pass needs the mask and it has nothing to prefetch, so it is peeled off the loop
#pragma unroll kUnroll
  for (int j = 0; j < NP - 1; ++j) {
    txy += GL;
    DpTaps nxt;
    dp_texel_fetch<STAGED, C::TW>(c, txy[0], nxt);
    uint32_t B, G, Rr;
    const int gray = dp_texel_blend(cur, B, G, Rr);
    if (WRITE_TEX) {
      const int i = sub + GL * j;
      tex_out[3 * i + 0] = (uint8_t)B;
      tex_out[3 * i + 1] = (uint8_t)G;
      tex_out[3 * i + 2] = (uint8_t)Rr;
    }
    *gs = (uint8_t)gray;
    gs += 32;
    ma += (unsigned)gray;
    mb += (unsigned)(gray * gray);
    cur = nxt;
  }
  {
    uint32_t B, G, Rr;
    int gray = dp_texel_blend(cur, B, G, Rr);
    const int i = sub + GL * (NP - 1);
    gray = (i < npx) ? gray : 0;
    if (WRITE_TEX && i < npx) {
      tex_out[3 * i + 0] = (uint8_t)B;
      tex_out[3 * i + 1] = (uint8_t)G;
      tex_out[3 * i + 2] = (uint8_t)Rr;
    }
    *gs = (uint8_t)gray;
    ma += (unsigned)gray;
    mb += (unsigned)(gray * gray);
  }
}

// tile: the group's staging tile (DP_GROUP_STAGE) -- every evaluation reads its ROIs from L2
// again (128 patches x 5 views per SM do not stay in L1), and with 8 patches per warp nearly
// every tap load had at least one lane missing L1; staged, the ROI is requested once, all rows
// at the same time, and the 4 x NP taps per lane are shared-memory reads.
// (Measured and rejected: requesting the next view's ROI right after the texel loop of the
// current view, so that its latency hides behind the reductions: -6 %; two tiles per group with
// the next view's ROI requested a whole view ahead, cp.async groups, 56 KB of shared memory
// per CTA: -4 %.  The staging wait is not what the warps stall on.  The rarely used unstaged
// texel loop out of line, to make the hot path's code smaller: -2 %.)

// Requests the ROI of one view into the group's tile; false when it does not fit the tile (the
// taps then come straight from global memory).  The caller has made sure that the group is done
// reading the tile.  Completion: dp_cp_async_wait_all() + __syncwarp(group).
template <typename C>
__device__ __forceinline__ bool dp_stage_issue(const DpViewSetupG &R, uint32_t *tile,
                                               const DpGroupLane &L) {
#if DP_GROUP_STAGE
  constexpr int PR = C::TW / 4;    // 16-byte pieces per tile row
  constexpr int RS = C::GL / PR;   // rows copied per step by the group
  static_assert(PR * RS == C::GL, "tile width / group size");
  // image rows are 128-byte aligned, so the pixel offset of the ROI inside its 16-byte
  // quad is visible in the pointer
  const int xoff = (int)((reinterpret_cast<uintptr_t>(R.src) >> 2) & 3u);
  const int wv = (R.rw + xoff + 3) >> 2;  // 16-byte pieces per ROI row
  if (wv > PR || R.rh > C::TH) return false;  // uniform inside the group
  const int piece = L.sub % PR, row0 = L.sub / PR;
  if (piece < wv) {
    const uint32_t *g = R.src - xoff + 4 * piece;
    uint32_t *d = tile + 4 * piece;
#pragma unroll
    for (int r = 0; r < C::TH; r += RS)
      if (r + row0 < R.rh)
        dp_cp_async16(d + (r + row0) * C::TW, g + (size_t)(r + row0) * R.pitch);
  }
  return true;
#else
  return false;
#endif
}

template <typename C, bool WRITE_TEX>
__device__ __forceinline__ void dp_view_texture_rolled(const DpViewSetupG &R, int npx,
                                                       const double2 *txy, uint32_t *tile,
                                                       bool staged, const DpGroupLane &L,
                                                       uint8_t *gs, unsigned &ma, unsigned &mb,
                                                       uint8_t *__restrict__ tex_out) {
  DpWarpConsts c;
  c.M0 = R.M0; c.M1 = R.M1; c.M2 = (double)R.M2; c.M3 = R.M3;
  c.M4 = R.M4; c.M5 = (double)R.M5; c.M6 = R.M6; c.M7 = R.M7;
  c.src = R.src;
  c.pitch = R.pitch;
  c.xmax = (R.rw - 1) << 5;
  c.ymax = (R.rh - 1) << 5;
  ma = 0;
  mb = 0;
  if (staged) {
    dp_cp_async_wait_all();
    __syncwarp(L.mask);
    c.src = tile + (int)((reinterpret_cast<uintptr_t>(R.src) >> 2) & 3u);
    dp_texel_loop<C, true, WRITE_TEX>(c, npx, txy, L.sub, gs, ma, mb, tex_out);
  } else {
    dp_texel_loop<C, false, WRITE_TEX>(c, npx, txy, L.sub, gs, ma, mb, tex_out);
  }
}

// Evaluate the visible views of the warp's patches in lockstep, GROUND views per round:
//   phase A  set-up of the round's views (one view of each patch per pass)
//   phase B  per view: stage the ROI, warp the texels, integer moments, NCC numerator
//   phase C  one view per lane of the group: NCCScore(texture 0, texture k)
// After each round sink(k0, kc, kcmax, score) runs in warp-uniform code: lane `sub` of a group
// holds the score of the group's view k0 + sub (sub < kc; -1 when either texture is empty,
// error_measurements.cpp:38-40; the entry of view 0 is meaningless); kcmax is the largest kc in
// the warp.  nv = 0 marks a group that does not evaluate.  Must be called by the whole warp.
template <typename C, bool WRITE_TEX, typename Sink>
__device__ __forceinline__ void dp_eval_views_g(const DpViewDev *__restrict__ views, int n_views,
                                                const DpLevelSel &lv,
                                                int ref, bool ref_ok, const int32_t *vis, int nv,
                                                int s, int npx, const double n[3], const double p[3],
                                                const double pc[3],
                                                DpViewSetupG *recs, const double2 *txy, uint8_t *gs,
                                                uint32_t *tile, int lane, const DpGroupLane &L,
                                                uint8_t *tex_base, uint8_t *valid_base, Sink sink) {
  DpFrame f;
  dp_make_frame(views + (ref_ok ? ref : 0), s, n, p, pc, f);
  if (!ref_ok) f.ok = false;  // every texture empty (optimization.cpp:45)
  constexpr int NP = C::NP, GL = C::GL, GROUND = C::GROUND;
  const double scale = 1.0 / (double)npx;  // cv::meanStdDev: mean = sum * (1/N)
  float da[NP];                             // centred anchor texels (texture 0)
  unsigned a1 = 0, a2 = 0;
  bool a_ok = false;
  const int nvmax = __reduce_max_sync(DP_FULL, nv);
#pragma unroll 1
  for (int k0 = 0; k0 < nvmax; k0 += GROUND) {
    const int kc = min(max(nv - k0, 0), GROUND);   // this group's views in the round
    const int kcmax = min(GROUND, nvmax - k0);     // warp-uniform loop bound
    __syncwarp();
    dp_setup_views<GL, DpViewSetupG>(views, n_views, lv, vis + k0, kc, kcmax, s, f, recs, lane);
    __syncwarp();
    unsigned my1 = 0, my2 = 0;
    double mynum = 0.0;
    int myok = 0;
#pragma unroll 1
    for (int l = 0; l < kcmax; ++l) {
      const DpViewSetupG &R = recs[l];
      const bool ok = l < kc && R.ok != 0;  // uniform inside the group
      unsigned s1 = 0, s2 = 0;
      double num = 0.0;
      if (ok) {
        unsigned ma = 0, mb = 0;
        __syncwarp(L.mask);  // the previous view's taps are done with the tile
        const bool staged = dp_stage_issue<C>(R, tile, L);
        dp_view_texture_rolled<C, WRITE_TEX>(
            R, npx, txy, tile, staged, L, gs, ma, mb,
            WRITE_TEX ? tex_base + (size_t)(k0 + l) * npx * 3 : nullptr);
        s1 = dp_group_sum<GL>(ma, L.mask);  // exact integer moments (cv::meanStdDev's sums)
        s2 = dp_group_sum<GL>(mb, L.mask);
        // fl32(g_i - fl32(mean)): `Mat - scalar` on CV_32F (error_measurements.cpp:54); the
        // lane reads back its own column of the gray buffer: no barrier needed
        const float mf = (float)xmul((double)s1, scale);
        if (k0 + l == 0) {
          a1 = s1;
          a2 = s2;
          a_ok = true;
          // only the last pass can hold texels past the patch (GL * (NP - 1) < npx)
#pragma unroll
          for (int j = 0; j < NP; ++j)
            da[j] = (j < NP - 1 || L.sub + GL * j < npx) ? __fsub_rn((float)gs[32 * j], mf) : 0.f;
        } else if (a_ok) {
#pragma unroll
          for (int j = 0; j < NP; ++j) {
            const float db = (j < NP - 1 || L.sub + GL * j < npx) ? __fsub_rn((float)gs[32 * j], mf) : 0.f;
            // the product of two floats is exact in fp64, so the fused form rounds exactly
            // like num + da * db
            num = fma((double)da[j], (double)db, num);
          }
          num = dp_group_sum<GL>(num, L.mask);
        }
      }
      if (L.sub == l) {
        my1 = s1;
        my2 = s2;
        mynum = num;
        myok = ok ? 1 : 0;
      }
    }
    // phase C, one view per lane of the group
    double score = -1.0;
    if (L.sub < kc && k0 + L.sub >= 1 && myok && a_ok)
      score = dp_ncc_finish(a1, a2, my1, my2, mynum, scale, npx);
    if (valid_base != nullptr && L.sub < kc) valid_base[k0 + L.sub] = (uint8_t)myok;
    sink(k0, kc, kcmax, score);
  }
}

// PatchOptimizationOpenCVFunctor::calc for all patches of the warp at once: mean of (1 - NCC)
// over the visible views in view order (optimization_opencv.cpp:17-35).
template <typename C>
__device__ __forceinline__ double dp_objective_g(const DpViewDev *__restrict__ views, int n_views,
                                                 const DpLevelSel &lv,
                                                 int ref, const int32_t *vis, int nv, int s, int npx,
                                                 const double n[3], const double p[3],
                                                 const double pc[3],
                                                 DpViewSetupG *recs, const double2 *txy,
                                                 uint8_t *gs, uint32_t *tile, int lane,
                                                 const DpGroupLane &L) {
  double sum = 0.0;
  dp_eval_views_g<C, false>(
      views, n_views, lv, ref, true, vis, nv, s, npx, n, p, pc, recs, txy, gs, tile, lane, L, nullptr, nullptr,
      [&](int k0, int kc, int kcmax, double score) {
        // std::accumulate of (1 - NCC) in view order (optimization_opencv.cpp:24, 34)
        const double term = xsub(1.0, score);
        for (int l = (k0 == 0 ? 1 : 0); l < kcmax; ++l) {
          const double t = __shfl_sync(DP_FULL, term, L.base + l);
          if (l < kc) sum = xadd(sum, t);
        }
      });
  return nv >= 2 ? sum / (double)(nv - 1) : 2.0;  // scores.size() == 0 -> 2
}

// Shared memory of a CTA of the group kernels.
template <typename C>
struct DpGroupShared {
  DpViewSetupG recs[DP_GWARPS][C::GROUPS][C::GROUND];
  double2 txy[C::NP * C::GL];
  // +32 words: the neighbour taps of an edge pixel (weight 0) may read past the last tile
  __align__(16) uint32_t tile[DP_GROUP_STAGE ? DP_GWARPS * C::GROUPS * C::TSTRIDE + 32 : 4];
  uint8_t gray[DP_GWARPS][C::NP][32];
  __device__ __forceinline__ void init_texels(int s, int npx) {
    for (int t = threadIdx.x; t < C::NP * C::GL; t += blockDim.x) {
      const int tt = t < npx ? t : 0;  // lanes past the last texel work on texel 0, masked later
      const int yy = tt / s;
      txy[t] = make_double2((double)(tt - yy * s), (double)yy);
    }
  }
  __device__ __forceinline__ uint32_t *group_tile(int warp, int grp) {
    return tile + (DP_GROUP_STAGE ? (warp * C::GROUPS + grp) * C::TSTRIDE : 0);
  }
};

template <int GL>
__device__ __forceinline__ DpGroupLane dp_group_lane(int lane) {
  DpGroupLane L;
  L.sub = lane & (GL - 1);
  L.base = lane & ~(GL - 1);
  L.mask = (GL == 32) ? DP_FULL : (((1u << GL) - 1u) << L.base);
  L.leader = L.sub == 0;
  return L;
}

// K1+K2 for cells up to 16x16: GetProjectedTextures + NCCScore for every visible view and,
// fused, FilterByErrorMeasurement's erase loop -- the warp-per-patch dp_score_kernel with
// C::GROUPS patches per warp.  Work item `slot` = patch order[slot] (patches sorted by view
// count so that the patches of a warp run the same number of view steps), or patch `slot`.
template <typename C, bool WRITE_TEX, bool FILTER>
__global__ void __launch_bounds__(DP_GWARPS * 32, DP_GMINCTA)
dp_score_group_kernel(DpScoreArgs a, const int32_t *__restrict__ order) {
  constexpr int GL = C::GL;
  __shared__ DpGroupShared<C> sh;
  const int s = a.p.s, npx = s * s;
  sh.init_texels(s, npx);
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const DpGroupLane L = dp_group_lane<GL>(lane);
  const int grp = lane / GL;
  const long long slot = ((long long)blockIdx.x * DP_GWARPS + warp) * C::GROUPS + grp;
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
  int wcur = 0;
  const double thr = a.thr;
  const unsigned lt = (1u << L.sub) - 1u;
  dp_eval_views_g<C, WRITE_TEX>(
      a.p.views, a.p.n_views, a.p.lv, ref, ref_ok, vis, nv, s, npx, n, p, pc,
      sh.recs[warp][grp], sh.txy + L.sub, &sh.gray[warp][0][lane], sh.group_tile(warp, grp), lane, L,
      tex, valid, [&](int k0, int kc, int kcmax, double score) {
        const int k = k0 + L.sub;
        const bool mine = L.sub < kc && k >= 1;
        if (ncc != nullptr && mine) ncc[k] = (float)score;
        if (FILTER) {
          // "drop original entry k-1 iff the score of entry k is low" (see dp_score_kernel),
          // compacted in place with the group's bits of a warp ballot
          const bool keepf = mine && !(score < thr);
          const int prev = mine ? vis[k - 1] : -1;
          const unsigned m = (__ballot_sync(DP_FULL, keepf) >> L.base) & ((GL == 32) ? DP_FULL : ((1u << GL) - 1u));
          __syncwarp();
          if (keepf) vis[wcur + __popc(m & lt)] = prev;
          wcur += __popc(m);
          __syncwarp();
        }
      });
  if (FILTER && have) {
    bool kept = false;
    if (nv >= 2) {  // scores.size() > 0 (optimization.cpp:113)
      if (L.leader) vis[wcur] = vis[nv - 1];
      ++wcur;
      __syncwarp(L.mask);
      for (int k = wcur + L.sub; k < nv; k += GL) vis[k] = -1;
      if (L.leader) a.p.nvis[i] = wcur;
      kept = wcur >= a.min_visible;  // optimization.cpp:127
    }
    if (L.leader) a.keep[i] = kept ? 1 : 0;
  }
}

template <typename C>
__global__ void __launch_bounds__(DP_GWARPS * 32, DP_GMINCTA) dp_refine_group_kernel(DpRefineArgs a) {
  constexpr int GL = C::GL;
  __shared__ DpGroupShared<C> sh;
  __shared__ DpNelderMead nm_s[DP_GWARPS][C::GROUPS];
  const int s = a.p.s, npx = s * s;
  sh.init_texels(s, npx);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const DpGroupLane L = dp_group_lane<GL>(lane);
  const int grp = lane / GL;
  DpViewSetupG *recs = sh.recs[warp][grp];
  DpNelderMead &S = nm_s[warp][grp];
  if (L.leader) {  // defined values for the lockstep evaluations of a group without a patch
    double *z = reinterpret_cast<double *>(&S);
    for (int j = 0; j < (int)(sizeof(DpNelderMead) / sizeof(double)); ++j) z[j] = 0.0;
  }
  __syncthreads();
  const double2 *txy = sh.txy + L.sub;
  bool have = false, exhausted = false;
  long long i = 0;
  int nv = 0, ref = 0;
  bool ref_ok = false;
  const int32_t *vis = a.p.vis;
  int state = NMG_INIT, idx = 0, fcount = 4, ilo = 0, ihi = 0;
#pragma unroll 1
  for (;;) {
    // ---- 1. a group without a patch takes the next one from the work counter ----------------
    if (!have && !exhausted) {
      for (;;) {
        unsigned int iu = 0;
        if (L.leader) iu = atomicAdd(a.work_counter, 1u);
        iu = __shfl_sync(L.mask, iu, L.base);
        if (iu >= (unsigned int)a.p.n) {
          exhausted = true;
          break;
        }
        i = a.order ? (long long)a.order[iu] : (long long)iu;
        if (a.mask != nullptr && a.mask[i] == 0) {  // removed by Seed::RemovePatches
          if (a.evals && L.leader) a.evals[i] = 0;
          continue;
        }
        nv = min(a.p.nvis[i], a.p.vstride);
        ref = a.p.ref[i];
        ref_ok = ref >= 0 && ref < a.p.n_views;
        vis = a.p.vis + (size_t)i * a.p.vstride;
        __syncwarp(L.mask);
        if (L.leader) {
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
        __syncwarp(L.mask);
        state = NMG_INIT;
        idx = 0;
        fcount = 4;
        ilo = ihi = 0;
        have = true;
        break;
      }
    }
    __syncwarp();
    if (!__any_sync(DP_FULL, have)) break;
    // ---- 2. one objective evaluation per group, in lockstep ---------------------------------
    double n[3], p[3];
    const double p0[3] = {S.p0[0], S.p0[1], S.p0[2]};  // GetPosition(): the corner centre
    {
      const double c3[3] = {S.c3[0], S.c3[1], S.c3[2]};
      const double n0[3] = {S.n0[0], S.n0[1], S.n0[2]};
      dp_unparametrize_g(c3, n0, p0, S.pt[0], S.pt[1], S.pt[2], n, p, L, DP_FULL);
    }
    const int nv_eval = (have && nv >= 2 && ref_ok) ? nv : 0;
    const double fobj = dp_objective_g<C>(a.p.views, a.p.n_views, a.p.lv, ref_ok ? ref : 0, vis, nv_eval, s,
                                           npx, n, p, p0, recs, txy, &sh.gray[warp][0][lane],
                                           sh.group_tile(warp, grp), lane, L);
    const double fval = nv_eval ? fobj : 2.0;  // scores.size() == 0 (optimization_opencv.cpp:30-32)
    // ---- 3. Nelder-Mead bookkeeping of each group (diverges by solver state, short) ---------
    if (have) {
      if (nmg_step(S, L, state, idx, fcount, ilo, ihi, fval, a.eps, a.max_evals)) {
        // best vertex -> one more trip through UnparametrizePatch, then write back;
        // SetNormal / SetPosition store fp32 (patch.h:38-53)
        const double c3[3] = {S.c3[0], S.c3[1], S.c3[2]};
        const double n0[3] = {S.n0[0], S.n0[1], S.n0[2]};
        double nb[3], pb[3];
        dp_unparametrize_g(c3, n0, p0, S.pt[0], S.pt[1], S.pt[2], nb, pb, L, L.mask);
        if (L.sub < 3) {
          const double nv_ = L.sub == 0 ? nb[0] : (L.sub == 1 ? nb[1] : nb[2]);
          const double pv_ = L.sub == 0 ? pb[0] : (L.sub == 1 ? pb[1] : pb[2]);
          if (ref_ok) {
            a.p.nrm[3 * i + L.sub] = (float)nv_;
            a.p.pos[3 * i + L.sub] = (float)pv_;
          }
          if (a.xbest) a.xbest[3 * i + L.sub] = S.pt[L.sub];
        }
        if (a.evals && L.leader) a.evals[i] = fcount;
        have = false;
      }
    }
  }
}
