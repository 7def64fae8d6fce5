// dp_device.cuh -- device-side building blocks of the PMVS photometric path (sm_100a).
//
// One warp owns one patch.  The warp walks the patch's visible views in order; for
// every view it (1) projects the four patch corners (one corner per lane quad),
// derives the ROI and the patch->ROI homography in closed form, (2) stages the ROI
// pixels (packed BGRx) into a shared-memory tile with row-coalesced loads, (3) warps
// the s x s texel grid through the homography with OpenCV's exact fixed-point
// bilinear arithmetic (1/32 px coordinates, 15-bit weights, u8 rounding, 15-bit gray),
// one texel per lane per pass, and (4) scores zero-mean NCC against the anchor (first
// visible) texture, which stays in registers, with warp-shuffle reductions.
//
// Reference semantics implemented here (paths under the reference root):
//   View::ProjectPoint / IsPointInside            modules/core/types.cpp:70-84
//   Patch::GetProjectedXYAxisAndScale             methods/pmvs/patch.cpp:86-104
//   Patch::ComputePatchToViewHomography           methods/pmvs/patch.cpp:111-164
//   Optimization::GetProjectedTextures            methods/pmvs/optimization.cpp:14-56
//   cv::warpPerspective(INTER_LINEAR, BORDER_REPLICATE) + cv::cvtColor(BGR2GRAY)
//   NCCScore                                      modules/core/error_measurements.cpp:36-60
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

struct DpViewDev {
  double P[12];      // View::GetProjectionMatrix(), row-major 3x4
  double xa[3];      // View::GetXAxis().normalized()
  double center[3];  // View::GetCameraCenter()
  const uint32_t *img;  // packed BGRx, pitch_px pixels per row
  int width, height, pitch_px;
  int gw, gh;              // PatchGrid dims: width / grid_scale, height / grid_scale
  long long grid_off;      // offset of this view's grid in the occupancy array
};

#define DP_FULL 0xffffffffu

// Per-(patch, view) pyramid level (dp_set_level_selection).  tab = the view tables
// of the base level (index 0 = the table the kernels get as `views`) and of the `up` coarser
// levels above it, [up + 1][n_views]; null = every view is read at the base level.
struct DpLevelSel {
  const DpViewDev *tab;
  int up;
  double thr2;  // (px_per_cell * s)^2: squared side of the projected quad beyond which the next level is used
};

// fp64 operations that must not be contracted into FMAs: the set-up chain
// (axes -> corners -> projection -> fp32 points -> ROI) mirrors the reference's
// unfused evaluation order so the fp32-rounded quad and the integer ROI agree.
__device__ __forceinline__ double xmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double xadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double xsub(double a, double b) { return __dsub_rn(a, b); }

__device__ __forceinline__ void dp_project(const double *__restrict__ P, double X0, double X1,
                                           double X2, double &u, double &v) {
  double x = xadd(xadd(xadd(xmul(P[0], X0), xmul(P[1], X1)), xmul(P[2], X2)), P[3]);
  double y = xadd(xadd(xadd(xmul(P[4], X0), xmul(P[5], X1)), xmul(P[6], X2)), P[7]);
  double w = xadd(xadd(xadd(xmul(P[8], X0), xmul(P[9], X1)), xmul(P[10], X2)), P[11]);
  u = x / w;
  v = y / w;
}

// 1/W to ~1 ulp without the IEEE-division slow path: hardware seed (20 bits) plus two
// Newton steps.  The result only feeds a coordinate that is then quantised to 1/32 px.
#ifndef DP_RCP3
#define DP_RCP3 1
#endif
__device__ __forceinline__ double dp_rcp(double w) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(w));
#if DP_RCP3
  // r0 (1 + e + e^2), e = 1 - w r0: relative error e^3 ~ 2^-60 in three FMAs
  const double e = fma(-w, r, 1.0);
  const double t = fma(e, e, e);
  return fma(r, t, r);
#else
  double e = fma(-w, r, 1.0);
  r = fma(r, e, r);
  e = fma(-w, r, 1.0);
  r = fma(r, e, r);
  return r;
#endif
}

__device__ __forceinline__ double warp_sum_f64(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(DP_FULL, v, o);
  return v;
}

// How many levels above the base a view is read at: the longer of the two quad sides through
// corner 0, in base-level pixels, is halved until it is shorter than px_per_cell * s (so a texel
// of the s x s cell covers less than px_per_cell pixels of the level it is sampled from).
// (du1, dv1) = corner 1 - corner 0, (du3, dv3) = corner 3 - corner 0.  NaN sides pick level 0.
__device__ __forceinline__ int dp_pick_level(double du1, double dv1, double du3, double dv3,
                                             double thr2, int up) {
  const double a = xadd(xmul(du1, du1), xmul(dv1, dv1));
  const double b = xadd(xmul(du3, du3), xmul(dv3, dv3));
  const double d2 = (b > a) ? b : a;
  int k = 0;
  double t = thr2;
  while (k < up && d2 >= t) {
    ++k;
    t = xmul(t, 4.0);
  }
  return k;
}

// Per-patch frame: scaled patch axes in world units (optimization.cpp:19-30).
struct DpFrame {
  double p[3];   // centre of the four corners = patch_.GetPosition(), the STORED position
                 // (patch.cpp:119-123), whatever (normal, position) GetProjectedTextures got
  double ax[3];  // scale * x_axis
  double ay[3];  // scale * y_axis (y = n x x, not normalised, patch.cpp:96)
  bool ok;       // false when dx == 0 (LOG(FATAL) in the reference, optimization.cpp:27)
};

// n, p: the (normal, position) ARGUMENTS of GetProjectedTextures -- they only feed
// GetProjectedXYAxisAndScale (optimization.cpp:24-26: y axis and dx); pc: the stored position.
// During Optimize() p is the trial position, so a trial depth rescales the quad around pc.
__device__ __forceinline__ void dp_make_frame(const DpViewDev *__restrict__ ref, int s,
                                              const double n[3], const double p[3],
                                              const double pc[3], DpFrame &f) {
  const double xa0 = ref->xa[0], xa1 = ref->xa[1], xa2 = ref->xa[2];
  // y_axis = normal.cross(x_axis)
  double ya0 = xsub(xmul(n[1], xa2), xmul(n[2], xa1));
  double ya1 = xsub(xmul(n[2], xa0), xmul(n[0], xa2));
  double ya2 = xsub(xmul(n[0], xa1), xmul(n[1], xa0));
  double cu, cv, qu, qv;
  dp_project(ref->P, p[0], p[1], p[2], cu, cv);
  dp_project(ref->P, xadd(p[0], xa0), xadd(p[1], xa1), xadd(p[2], xa2), qu, qv);
  double du = xsub(qu, cu), dv = xsub(qv, cv);
  double dx = sqrt(xadd(xmul(du, du), xmul(dv, dv)));
  f.ok = (dx != 0.0) && isfinite(dx);
  double scale = (double)(s / 2) / dx;  // integer cell_size / 2 (optimization.cpp:30)
  f.p[0] = pc[0]; f.p[1] = pc[1]; f.p[2] = pc[2];
  f.ax[0] = xmul(scale, xa0); f.ax[1] = xmul(scale, xa1); f.ax[2] = xmul(scale, xa2);
  f.ay[0] = xmul(scale, ya0); f.ay[1] = xmul(scale, ya1); f.ay[2] = xmul(scale, ya2);
}

// Per-lane texel grid: texel i = lane + 32*j of the s x s destination, row-major.
// Up to 8 passes (s <= 16) the coordinates live in registers; above that they are
// recomputed per texel to keep the register count bounded.
template <int NPASS, bool PRE = (NPASS <= 8)>
struct DpTexels {
  double x[NPASS], y[NPASS];
  __device__ __forceinline__ void init(int s, int lane) {
#pragma unroll
    for (int j = 0; j < NPASS; ++j) {
      int i = lane + 32 * j;
      int yy = i / s;
      x[j] = (double)(i - yy * s);
      y[j] = (double)yy;
    }
  }
  __device__ __forceinline__ void get(int j, int, double &xo, double &yo) const {
    xo = x[j];
    yo = y[j];
  }
};
template <int NPASS>
struct DpTexels<NPASS, false> {
  int s_;
  float inv_s_;
  __device__ __forceinline__ void init(int s, int) {
    s_ = s;
    inv_s_ = 1.0f / (float)s;
  }
  __device__ __forceinline__ void get(int, int i, double &xo, double &yo) const {
    int yy = (int)(((float)i + 0.5f) * inv_s_);
    xo = (double)(i - yy * s_);
    yo = (double)yy;
  }
};

// ---------------------------------------------------------------------------------------
// Phase A: per-view set-up, batched across the warp.
//
// Everything that happens once per (patch, view) -- four corner projections, inside test,
// ROI, the cell -> quad projective map -- is scalar work.  Doing it with all 32 lanes for one
// view at a time wastes 31/32 of the machine, so it is batched instead: lane = 4*slot +
// corner handles one corner of one of 8 views per pass, the four lanes of a slot share
// their results with width-4 shuffles and derive the view's map, and lane 4*slot writes a
// DpViewSetup record to shared memory.  Phase B then only reads one record per view.

struct __align__(16) DpViewSetup {
  double M[8];          // source = (M0 x + M1 y + M2, M3 x + M4 y + M5) / (M6 x + M7 y + 1),
                        // in 1/32-px units (pre-scaled by INTER_TAB_SIZE), relative to the ROI
  const uint32_t *src;  // first pixel of the ROI (packed BGRx)
  int pitch, rw, rh;    // image pitch (pixels), ROI width / height
  int ok;               // 0 where the reference pushes an empty cv::Mat
  int lgp;              // log2 of the staged tile's row pitch (next power of two >= rw)
  int pad_;
};
static_assert(sizeof(DpViewSetup) == 96, "DpViewSetup layout");

// The same record without the TMA / generic-staging fields, for the group kernels (dp_group.cuh):
// 80 bytes.  M2 and M5 are 32 x (an fp32 number), exactly representable in fp32.
struct __align__(16) DpViewSetupG {
  double M0, M1, M3, M4, M6, M7;
  const uint32_t *src;
  float M2, M5;
  int pitch, rw, rh, ok;
};
static_assert(sizeof(DpViewSetupG) == 80, "DpViewSetupG layout");

__device__ __forceinline__ uint32_t dp_smem_u32(const void *p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}

__device__ __forceinline__ void dp_write_setup(DpViewSetup &R, double m0, double m1, double m2,
                                               double m3, double m4, double m5, double m6, double m7,
                                               const uint32_t *src, int pitch, int rw, int rh, bool ok) {
  R.M[0] = m0; R.M[1] = m1; R.M[2] = m2; R.M[3] = m3;
  R.M[4] = m4; R.M[5] = m5; R.M[6] = m6; R.M[7] = m7;
  R.src = src;
  R.pitch = pitch;
  R.rw = rw;
  R.rh = rh;
  R.ok = ok ? 1 : 0;
  R.lgp = 32 - __clz(max(rw, 1) - 1);  // ceil(log2(rw))
  R.pad_ = 0;
}
__device__ __forceinline__ void dp_write_setup(DpViewSetupG &R, double m0, double m1, double m2,
                                               double m3, double m4, double m5, double m6, double m7,
                                               const uint32_t *src, int pitch, int rw, int rh, bool ok) {
  R.M0 = m0; R.M1 = m1; R.M3 = m3; R.M4 = m4; R.M6 = m6; R.M7 = m7;
  R.M2 = (float)m2;
  R.M5 = (float)m5;
  R.src = src;
  R.pitch = pitch;
  R.rw = rw;
  R.rh = rh;
  R.ok = ok ? 1 : 0;
}

// cv::findHomography on 4 points is the exact projective map quad -> [0,s]^2 and
// cv::warpPerspective uses its inverse; that inverse (cell -> quad) has the closed form below
// (unit square -> quadrilateral): no 9x9 eigen-solve, no 3x3 inversion.  Outputs the six
// non-trivial coefficients of source = (m0 x + m1 y + 32 qx0, m3 x + m4 y + 32 qy0) /
// (m6 x + m7 y + 1) in 1/32-px units; false for a degenerate quad or non-finite coefficients.
__device__ __forceinline__ bool dp_quad_map(double qx0, double qy0, double qx1, double qy1,
                                            double qx2, double qy2, double qx3, double qy3,
                                            double inv_s, double &m0, double &m1, double &m3,
                                            double &m4, double &m6, double &m7) {
  const double sxq = qx0 - qx1 + qx2 - qx3, syq = qy0 - qy1 + qy2 - qy3;
  const double dx1 = qx1 - qx2, dx2 = qx3 - qx2, dy1 = qy1 - qy2, dy2 = qy3 - qy2;
  const double den = dx1 * dy2 - dx2 * dy1;
  const double rden = 1.0 / den;
  const double gq = (sxq * dy2 - dx2 * syq) * rden;
  const double hq = (dx1 * syq - sxq * dy1) * rden;
  m0 = 32.0 * (qx1 - qx0 + gq * qx1) * inv_s;
  m1 = 32.0 * (qx3 - qx0 + hq * qx3) * inv_s;
  m3 = 32.0 * (qy1 - qy0 + gq * qy1) * inv_s;
  m4 = 32.0 * (qy3 - qy0 + hq * qy3) * inv_s;
  m6 = gq * inv_s;
  m7 = hq * inv_s;
  const bool fin = isfinite(m0) && isfinite(m1) && isfinite(m3) && isfinite(m4) && isfinite(m6) &&
                   isfinite(m7);
  return fin && (den != 0.0);
}

#define DP_ROUND 16  // views whose set-up records are resident at once (per warp)

// GL = lanes that share one patch (32: the whole warp; 8: four patches per warp, each group of
// 8 lanes sets up GL/4 views of its own patch per pass and writes to its own `recs`).  kcount
// is this group's number of views in the round, kcmax the largest kcount in the warp (the loop
// bound must be warp-uniform: the shuffles below are full-warp).
template <int GL = 32, typename REC = DpViewSetup>
__device__ __forceinline__ void dp_setup_views(const DpViewDev *__restrict__ views, int n_views,
                                               const DpLevelSel &lv, const int32_t *vis, int kcount, int kcmax, int s,
                                               const DpFrame &f, REC *recs, int lane) {
  const int c = lane & 3, slot = (lane & (GL - 1)) >> 2;
  const double sgx = (c == 1 || c == 2) ? 1.0 : -1.0;  // corners (-,-) (+,-) (+,+) (-,+),
  const double sgy = (c >= 2) ? 1.0 : -1.0;            // patch.cpp:119-123
  const double X0 = xadd(xadd(f.p[0], sgx * f.ax[0]), sgy * f.ay[0]);
  const double X1 = xadd(xadd(f.p[1], sgx * f.ax[1]), sgy * f.ay[1]);
  const double X2 = xadd(xadd(f.p[2], sgx * f.ax[2]), sgy * f.ay[2]);
  const double inv_s = 1.0 / (double)s;
#pragma unroll 1
  for (int base = 0; base < kcmax; base += GL / 4) {
    const int k = base + slot;
    const bool active = k < kcount;
    const int vid = active ? vis[k] : -1;
    const bool inr = active && f.ok && vid >= 0 && vid < n_views;
    const DpViewDev *V = views + (inr ? vid : 0);
    double u, v;
    dp_project(V->P, X0, X1, X2, u, v);
    if (lv.tab != nullptr) {  // read this view at the level its footprint asks for
      const double u0 = __shfl_sync(DP_FULL, u, 0, 4), v0 = __shfl_sync(DP_FULL, v, 0, 4);
      const double u1 = __shfl_sync(DP_FULL, u, 1, 4), v1 = __shfl_sync(DP_FULL, v, 1, 4);
      const double u3 = __shfl_sync(DP_FULL, u, 3, 4), v3 = __shfl_sync(DP_FULL, v, 3, 4);
      const int up = dp_pick_level(xsub(u1, u0), xsub(v1, v0), xsub(u3, u0), xsub(v3, v0), lv.thr2, lv.up);
      if (up > 0) {  // P_l = diag(2^-l, 2^-l, 1) P: the projection scales exactly
        const double sc = __hiloint2double((1023 - up) << 20, 0);
        u = xmul(u, sc);
        v = xmul(v, sc);
        V = lv.tab + ((size_t)up * n_views + (inr ? vid : 0));
      }
    }
    const int W = V->width, H = V->height;
    const bool in = inr && (u > 0) && (u < (double)W) && (v > 0) && (v < (double)H);
    const unsigned inm = __ballot_sync(DP_FULL, in);
    const bool all_in = ((inm >> (lane & ~3)) & 0xfu) == 0xfu;  // any corner outside -> empty
    // ROI: tl = min ceil, br = max floor over the 4 corners (patch.cpp:126-147)
    int tlx = __double2int_ru(u), tly = __double2int_ru(v);
    int brx = __double2int_rd(u), bry = __double2int_rd(v);
#pragma unroll
    for (int o = 1; o <= 2; o <<= 1) {
      tlx = min(tlx, __shfl_xor_sync(DP_FULL, tlx, o));
      tly = min(tly, __shfl_xor_sync(DP_FULL, tly, o));
      brx = max(brx, __shfl_xor_sync(DP_FULL, brx, o));
      bry = max(bry, __shfl_xor_sync(DP_FULL, bry, o));
    }
    tlx = min(tlx, W); tly = min(tly, H); brx = max(brx, 0); bry = max(bry, 0);
    const int rw = brx - tlx, rh = bry - tly;
    // cv::Point2f, then `-= roi.x` in fp32 (patch.cpp:134, 148-151)
    const float fx = __fsub_rn((float)u, (float)tlx);
    const float fy = __fsub_rn((float)v, (float)tly);
    const double qx0 = (double)__shfl_sync(DP_FULL, fx, 0, 4), qy0 = (double)__shfl_sync(DP_FULL, fy, 0, 4);
    const double qx1 = (double)__shfl_sync(DP_FULL, fx, 1, 4), qy1 = (double)__shfl_sync(DP_FULL, fy, 1, 4);
    const double qx2 = (double)__shfl_sync(DP_FULL, fx, 2, 4), qy2 = (double)__shfl_sync(DP_FULL, fy, 2, 4);
    const double qx3 = (double)__shfl_sync(DP_FULL, fx, 3, 4), qy3 = (double)__shfl_sync(DP_FULL, fy, 3, 4);
    double m0, m1, m3, m4, m6, m7;
    const bool mapped = dp_quad_map(qx0, qy0, qx1, qy1, qx2, qy2, qx3, qy3, inv_s, m0, m1, m3, m4, m6, m7);
    if (c == 0 && active) {
      const bool ok = all_in && rw > 0 && rh > 0 && mapped;  // optimization.cpp:45
      dp_write_setup(recs[k], m0, m1, 32.0 * qx0, m3, m4, 32.0 * qy0, m6, m7,
                     V->img + (ok ? (size_t)tly * V->pitch_px + tlx : 0), V->pitch_px, rw, rh, ok);
    }
  }
}

// ---------------------------------------------------------------------------------------
// Phase B: the texture of one view from its set-up record: gray value of every texel owned
// by this lane (g[j], 0..255), optionally the BGR texels themselves.  `tile` is this warp's
// shared-memory staging buffer of tile_cap pixels (row pitch 2^R.lgp).

// Generic staging: tile rows padded to a power-of-two pitch so the flat index splits with a
// shift and a mask; each 32-lane step loads 32/pitch whole rows with 32-bit loads.
template <int NFAST>
__device__ __forceinline__ bool dp_stage_roi(const DpViewSetup &R, uint32_t *tile, int tile_cap,
                                             int lane) {
  const int lgp = R.lgp, rw = R.rw, pitch = R.pitch;
  const int area = R.rh << lgp;
  if (area > tile_cap) return false;
  const uint32_t *__restrict__ src = R.src;
  const int cmask = (1 << lgp) - 1;
  // The first 32*NFAST tile entries (the whole ROI in the normal case) without a branch: all
  // loads are issued before the first store, lanes past the ROI load nothing and store 0.
  uint32_t v[NFAST];
#pragma unroll
  for (int u = 0; u < NFAST; ++u) {
    const int t = lane + 32 * u;
    const int r = t >> lgp, c = t & cmask;
    v[u] = (t < area && c < rw) ? __ldg(src + (unsigned)(r * pitch + c)) : 0u;
  }
#pragma unroll
  for (int u = 0; u < NFAST; ++u) tile[lane + 32 * u] = v[u];
  if (area > 32 * NFAST)
    for (int t = 32 * NFAST + lane; t < area; t += 32) {
      const int r = t >> lgp, c = t & cmask;
      if (c < rw) tile[t] = __ldg(src + (unsigned)(r * pitch + c));
    }
  __syncwarp();
  return true;
}

// `lane` is the lane's index inside its group of GL lanes; it owns texels lane + GL*j.
template <int NPASS, bool WRITE_TEX, bool staged, int GL = 32, typename TX = DpTexels<NPASS>>
__device__ __forceinline__ void dp_view_texture(const DpViewSetup &R, int npx, const TX &tx,
                                                const uint32_t *tile0, int lane, int (&g)[NPASS],
                                                uint8_t *__restrict__ tex_out) {
  const uint32_t *tile = tile0;
  const double M0 = R.M[0], M1 = R.M[1], M2 = R.M[2], M3 = R.M[3], M4 = R.M[4], M5 = R.M[5],
               M6 = R.M[6], M7 = R.M[7];
  const uint32_t *__restrict__ src = R.src;
  const int pitch = R.pitch, lgp = R.lgp;
  const int xmax = (R.rw - 1) << 5, ymax = (R.rh - 1) << 5;
  // ---- warp the texel grid ------------------------------------------------------------------
  // Branch-free: lanes past the last texel compute on a clamped (harmless) coordinate and are
  // masked at the end, so the NPASS independent passes can be interleaved by the scheduler.
#pragma unroll
  for (int j = 0; j < NPASS; ++j) {
    const int i = lane + GL * j;
    double x, y;
    tx.get(j, i, x, y);
    const double Wd = fma(M6, x, fma(M7, y, 1.0));
    // W ? INTER_TAB_SIZE / W : 0 without a select: W == 0 gives r = NaN after the Newton steps,
    // NaN coordinates, and cvt.rni.s32.f64 turns NaN into 0 like the reference's zero scale
    const double r = dp_rcp(Wd);
    const double fX = fma(M0, x, fma(M1, y, M2)) * r;
    const double fY = fma(M3, x, fma(M4, y, M5)) * r;
    const int Xi = __double2int_rn(fX);  // saturate_cast<int>(cvRound), half to even
    const int Yi = __double2int_rn(fY);
    // BORDER_REPLICATE at the ROI edge: a clamped tap pair reads the same pixel twice, so the
    // result is that pixel whatever the weight.  Clamping the 1/32-px coordinate itself to
    // [0, 32 (rw-1)] gives the same blend -- inside nothing changes, outside the coordinate
    // lands exactly on the edge pixel with weight 0 for its right / lower neighbour -- and the
    // taps are simply (x0, y0) + {0,1}^2.  The neighbour of an edge pixel may lie outside the
    // ROI (never outside the allocation: images carry one spare row); its weight is 0.
    const int Xc = min(max(Xi, 0), xmax), Yc = min(max(Yi, 0), ymax);
    const int x0 = Xc >> 5, axw = Xc & 31;  // INTER_BITS = 5
    const int y0 = Yc >> 5, ayw = Yc & 31;
    uint32_t p00, p01, p10, p11;
    if (staged) {
      const uint32_t *t0 = tile + ((y0 << lgp) + x0), *t1 = t0 + (1 << lgp);
      p00 = t0[0]; p01 = t0[1];
      p10 = t1[0]; p11 = t1[1];
    } else {
      // unsigned 32-bit element offset from the ROI origin: one IMAD.WIDE.U32 per row
      const uint32_t *r0 = src + (unsigned)(y0 * pitch + x0), *r1 = r0 + pitch;
      p00 = __ldg(r0); p01 = __ldg(r0 + 1);
      p10 = __ldg(r1); p11 = __ldg(r1 + 1);
    }
    // separable form of the 15-bit weights (32-ax)(32-ay)*32 ...: exact in integers,
    // (sum*32 + 2^14) >> 15 == (sum + 2^9) >> 10.  B and R share one multiply per tap pair
    // (B | R<<16, each partial sum <= 255*32).
    const uint32_t wx1 = (uint32_t)axw, wx0 = 32u - wx1, wy1 = (uint32_t)ayw, wy0 = 32u - wy1;
    const uint32_t br0 = (p00 & 0x00ff00ffu) * wx0 + (p01 & 0x00ff00ffu) * wx1;  // B | R<<16
    const uint32_t br1 = (p10 & 0x00ff00ffu) * wx0 + (p11 & 0x00ff00ffu) * wx1;
    // G stays in place (bits 8..15, partial sums <= 26 bits): the blend of the whole pixel word
    // (x byte = 0, < 2^30) minus its B | R part -- two masks less per tap pair, exact
    const uint32_t g0 = (p00 * wx0 + p01 * wx1) - br0;
    const uint32_t g1 = (p10 * wx0 + p11 * wx1) - br1;
    const uint32_t B = ((br0 & 0xffffu) * wy0 + (br1 & 0xffffu) * wy1 + 512u) >> 10;
    const uint32_t Rr = ((br0 >> 16) * wy0 + (br1 >> 16) * wy1 + 512u) >> 10;
    const uint32_t G = (g0 * wy0 + g1 * wy1 + (512u << 8)) >> 18;
    // cv::cvtColor(BGR2GRAY), 8U: 15-bit fixed point
    const int gray = (int)((3735u * B + 19235u * G + 9798u * Rr + (1u << 14)) >> 15);
    g[j] = (i < npx) ? gray : 0;
    if (WRITE_TEX && i < npx) {
      tex_out[3 * i + 0] = (uint8_t)B;
      tex_out[3 * i + 1] = (uint8_t)G;
      tex_out[3 * i + 2] = (uint8_t)Rr;
    }
  }
  if (staged) __syncwarp();  // the tile may be overwritten by the next view
}

// Integer moments of the gray texels held by the warp (exact; cv::meanStdDev's sums).
template <int NPASS>
__device__ __forceinline__ void dp_moments(const int (&g)[NPASS], unsigned &s1, unsigned &s2) {
  unsigned a = 0, b = 0;
#pragma unroll
  for (int j = 0; j < NPASS; ++j) {
    a += (unsigned)g[j];
    b += (unsigned)(g[j] * g[j]);
  }
  s1 = __reduce_add_sync(DP_FULL, a);
  s2 = __reduce_add_sync(DP_FULL, b);
}

// fl32(g_i - fl32(mean)): `Mat - scalar` on CV_32F (error_measurements.cpp:54)
template <int NPASS>
__device__ __forceinline__ void dp_centre(const int (&g)[NPASS], unsigned s1, double scale, int npx,
                                          int lane, float (&d)[NPASS]) {
  const float mf = (float)xmul((double)s1, scale);
#pragma unroll
  for (int j = 0; j < NPASS; ++j) d[j] = (lane + 32 * j < npx) ? __fsub_rn((float)g[j], mf) : 0.f;
}

// Phase C, one view per lane: NCCScore from the exact moments and the numerator
// (error_measurements.cpp:47-59): population sigma, clamp 0.1, (num / den) / N.
__device__ __forceinline__ double dp_ncc_finish(unsigned a1, unsigned a2, unsigned b1, unsigned b2,
                                                double num, double scale, int npx) {
  const double ma = xmul((double)a1, scale), mb = xmul((double)b1, scale);
  const double va = xsub(xmul((double)a2, scale), xmul(ma, ma));
  const double vb = xsub(xmul((double)b2, scale), xmul(mb, mb));
  const double sa = sqrt(va > 0.0 ? va : 0.0), sb = sqrt(vb > 0.0 ? vb : 0.0);
  double den = xmul(sa, sb);
  den = den > 1e-1 ? den : 1e-1;
  return (num / den) / (double)npx;
}
