/*
 * dp_oracle.c -- CPU ORACLE (test infrastructure; see dp_oracle.h for the rules).
 *
 * Restates, function by function, the reference's photometric hot path
 * (manlito/densepoints methods/pmvs + modules/core) and the OpenCV primitives it
 * calls.  Compile with -ffp-contract=off so every fp64 expression is evaluated
 * exactly as written (no FMA contraction).
 */
#include "dp_oracle.h"

#include <float.h>
#include <limits.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

int orc_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

void orc_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

void orc_default_params(orc_params *p) {
  p->score_threshold = 0.6;       /* optimization.h:16 */
  p->minimum_visible_image = 3;   /* optimization.h:17 */
  p->visible_threshold = 0.78;    /* patch.h:56 */
  p->candidate_threshold = 1.04;  /* patch.h:57 */
  p->grid_scale = 8;              /* patch_organizer.h:43 */
  p->max_patches_per_cell = 1;    /* patch_organizer.h:42 */
  p->nm_step[0] = 0.02;           /* optimization_opencv.cpp:56 */
  p->nm_step[1] = 0.2;
  p->nm_step[2] = 0.2;
  p->nm_max_evals = 500;          /* optimization_opencv.cpp:60 */
  p->nm_eps = 0.0001;
  p->max_pops = 10000000LL;       /* expand.cpp:95 */
}

/* ------------------------------------------------------------------------ */
/* small vector helpers                                                      */
/* Evaluation order of Eigen's small fixed-size sums: 0 = sequential
 * ((a0+a1)+a2)+a3, what Eigen 3.2's unrolled coefficient products do and what the oracle and the
 * kernels assume; 1 = the halving order of Eigen >= 3.3's redux unroller, (a0+a1)+(a2+a3) for four
 * terms and a0+(a1+a2) for three.  Mode 1 exists to MEASURE what the unpinned Eigen version can
 * change (tests/test_oracle_logic.py); the product path is mode 0. */
static int g_eigen_pairwise = 0;
void orc_set_eigen_pairwise(int on) { g_eigen_pairwise = on; }

static double dot3(const double a[3], const double b[3]) {
  if (g_eigen_pairwise) return a[0] * b[0] + (a[1] * b[1] + a[2] * b[2]);
  return a[0] * b[0] + a[1] * b[1] + a[2] * b[2];
}
static double norm3(const double a[3]) { return sqrt(dot3(a, a)); }
static void cross3(const double a[3], const double b[3], double c[3]) {
  c[0] = a[1] * b[2] - a[2] * b[1];
  c[1] = a[2] * b[0] - a[0] * b[2];
  c[2] = a[0] * b[1] - a[1] * b[0];
}

/* ------------------------------------------------------------------------ */
/* modules/core/types.cpp                                                    */

/* View::SetProjectionMatrix (types.cpp:28-68).  The reference takes the camera
 * centre from the SVD null vector of P and K,R from an RQ decomposition of
 * P[:, :3] with the diagonal of K forced positive (types.cpp:57-66).  With a
 * positive diagonal the orthogonal factor is unique, so it is restated here as
 * bottom-up Gram-Schmidt on the rows of M; the centre as C = -M^-1 p4 (the same
 * null vector).  Only row 0 of R (GetXAxis) and C feed the hot path. */
void orc_view_decompose(const double P[12], double K[9], double R[9], double center[3]) {
  double m[3][3], p4[3];
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) m[i][j] = P[i * 4 + j];
    p4[i] = P[i * 4 + 3];
  }
  /* centre: solve M C = -p4 by the adjugate */
  double c0[3], c1[3], c2[3];
  cross3(m[1], m[2], c0);
  cross3(m[2], m[0], c1);
  cross3(m[0], m[1], c2);
  double det = dot3(m[0], c0);
  for (int j = 0; j < 3; ++j)
    center[j] = -(c0[j] * p4[0] + c1[j] * p4[1] + c2[j] * p4[2]) / det;
  /* rows of R, bottom-up */
  double r[3][3];
  double n2 = norm3(m[2]);
  for (int j = 0; j < 3; ++j) r[2][j] = m[2][j] / n2;
  double k12 = dot3(m[1], r[2]);
  double t[3];
  for (int j = 0; j < 3; ++j) t[j] = m[1][j] - k12 * r[2][j];
  double k11 = norm3(t);
  for (int j = 0; j < 3; ++j) r[1][j] = t[j] / k11;
  double k02 = dot3(m[0], r[2]);
  double k01 = dot3(m[0], r[1]);
  for (int j = 0; j < 3; ++j) t[j] = m[0][j] - k02 * r[2][j] - k01 * r[1][j];
  double k00 = norm3(t);
  for (int j = 0; j < 3; ++j) r[0][j] = t[j] / k00;
  double k22 = n2;
  double Kt[9] = {k00, k01, k02, 0, k11, k12, 0, 0, k22};
  for (int i = 0; i < 9; ++i) K[i] = Kt[i] / k22; /* types.cpp:65 */
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) R[i * 3 + j] = r[i][j];
}

void orc_view_init(orc_view *v, const double P[12], const uint8_t *bgr, int width, int height,
                   size_t stride) {
  double K[9], R[9];
  memcpy(v->P, P, sizeof(double) * 12);
  orc_view_decompose(P, K, R, v->center);
  v->xaxis[0] = R[0]; /* View::GetXAxis, types.cpp:86-89 */
  v->xaxis[1] = R[1];
  v->xaxis[2] = R[2];
  v->width = width;
  v->height = height;
  v->bgr = bgr;
  v->stride = stride;
}

/* View::ProjectPoint (types.cpp:70-75) */
void orc_project(const orc_view *v, const double X[3], double uv[2]) {
  const double *P = v->P;
  double x, y, w;
  if (g_eigen_pairwise) {
    x = (P[0] * X[0] + P[1] * X[1]) + (P[2] * X[2] + P[3]);
    y = (P[4] * X[0] + P[5] * X[1]) + (P[6] * X[2] + P[7]);
    w = (P[8] * X[0] + P[9] * X[1]) + (P[10] * X[2] + P[11]);
  } else {
    x = P[0] * X[0] + P[1] * X[1] + P[2] * X[2] + P[3];
    y = P[4] * X[0] + P[5] * X[1] + P[6] * X[2] + P[7];
    w = P[8] * X[0] + P[9] * X[1] + P[10] * X[2] + P[11];
  }
  uv[0] = x / w;
  uv[1] = y / w;
}

/* View::IsPointInside (types.cpp:77-84): strict, no depth-sign test */
int orc_inside(const orc_view *v, const double X[3]) {
  double uv[2];
  orc_project(v, X, uv);
  return (uv[0] > 0 && uv[0] < v->width && uv[1] > 0 && uv[1] < v->height) ? 1 : 0;
}

/* ------------------------------------------------------------------------ */
/* modules/core/error_measurements.cpp                                       */

/* cv::cvtColor(BGR2GRAY), 8U, OpenCV 4.x: 15-bit fixed point. */
int orc_gray(int b, int g, int r) { return (3735 * b + 19235 * g + 9798 * r + (1 << 14)) >> 15; }

/* NCCScore body after ToFloatMat (error_measurements.cpp:47-59):
 * cv::meanStdDev on CV_32F (fp64 sums, population sigma), `Mat - double` on
 * CV_32F (scalar rounded to fp32, subtraction in fp32), Mat::dot, clamp 0.1. */
static double ncc_float(const float *a, const float *b, int n) {
  double sa = 0, sb = 0, qa = 0, qb = 0;
  for (int i = 0; i < n; ++i) {
    sa += a[i];
    qa += (double)a[i] * a[i];
    sb += b[i];
    qb += (double)b[i] * b[i];
  }
  double scale = 1.0 / n;
  double mean_a = sa * scale, mean_b = sb * scale;
  double var_a = qa * scale - mean_a * mean_a;
  double var_b = qb * scale - mean_b * mean_b;
  double std_a = sqrt(var_a > 0 ? var_a : 0);
  double std_b = sqrt(var_b > 0 ? var_b : 0);
  float ma = (float)mean_a, mb = (float)mean_b;
  double numerator = 0;
  for (int i = 0; i < n; ++i) {
    float da = a[i] - ma;
    float db = b[i] - mb;
    numerator += (double)da * (double)db;
  }
  double denominator = std_a * std_b;
  denominator = denominator > 1e-1 ? denominator : 1e-1; /* :57 */
  return (numerator / denominator) / (double)n;          /* :58 */
}

/* NCCScore (error_measurements.cpp:36-60), CV_8UC3 inputs */
double orc_ncc_bgr(const uint8_t *ta, const uint8_t *tb, int n) {
  if (!ta || !tb) return -1; /* :38-40 empty Mat */
  float a[1024], b[1024];
  if (n > 1024) return -1;
  for (int i = 0; i < n; ++i) {
    a[i] = (float)orc_gray(ta[3 * i], ta[3 * i + 1], ta[3 * i + 2]);
    b[i] = (float)orc_gray(tb[3 * i], tb[3 * i + 1], tb[3 * i + 2]);
  }
  return ncc_float(a, b, n);
}

/* NCCScore on CV_64F inputs (tests/core/test_error_functions.cpp:9-15) */
double orc_ncc_f64(const double *a, const double *b, int n) {
  float fa[1024], fb[1024];
  if (n > 1024) return -1;
  for (int i = 0; i < n; ++i) {
    fa[i] = (float)a[i];
    fb[i] = (float)b[i];
  }
  return ncc_float(fa, fb, n);
}

/* ------------------------------------------------------------------------ */
/* OpenCV primitives                                                         */

/* cyclic Jacobi eigen-solver for a symmetric n x n matrix (n <= 9); V rows are
 * eigenvectors, sorted by descending eigenvalue (cv::eigen contract). */
static void jacobi_eigen(double *A, int n, double *W, double *V) {
  for (int i = 0; i < n; ++i) {
    for (int j = 0; j < n; ++j) V[i * n + j] = (i == j);
    W[i] = A[i * n + i];
  }
  for (int sweep = 0; sweep < 64; ++sweep) {
    double off = 0;
    for (int i = 0; i < n; ++i)
      for (int j = i + 1; j < n; ++j) off += A[i * n + j] * A[i * n + j];
    if (off < 1e-300) break;
    for (int p = 0; p < n; ++p) {
      for (int q = p + 1; q < n; ++q) {
        double apq = A[p * n + q];
        if (fabs(apq) < 1e-300) continue;
        double theta = (A[q * n + q] - A[p * n + p]) / (2.0 * apq);
        double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < n; ++k) { /* A <- A J */
          double akp = A[k * n + p], akq = A[k * n + q];
          A[k * n + p] = c * akp - s * akq;
          A[k * n + q] = s * akp + c * akq;
        }
        for (int k = 0; k < n; ++k) { /* A <- J^T A */
          double apk = A[p * n + k], aqk = A[q * n + k];
          A[p * n + k] = c * apk - s * aqk;
          A[q * n + k] = s * apk + c * aqk;
        }
        for (int k = 0; k < n; ++k) { /* eigenvectors as rows */
          double vpk = V[p * n + k], vqk = V[q * n + k];
          V[p * n + k] = c * vpk - s * vqk;
          V[q * n + k] = s * vpk + c * vqk;
        }
      }
    }
  }
  for (int i = 0; i < n; ++i) W[i] = A[i * n + i];
  for (int i = 0; i < n - 1; ++i) { /* sort descending */
    int m = i;
    for (int j = i + 1; j < n; ++j)
      if (W[j] > W[m]) m = j;
    if (m != i) {
      double tw = W[i];
      W[i] = W[m];
      W[m] = tw;
      for (int k = 0; k < n; ++k) {
        double tv = V[i * n + k];
        V[i * n + k] = V[m * n + k];
        V[m * n + k] = tv;
      }
    }
  }
}

static void mat3_mul(const double a[9], const double b[9], double c[9]) {
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j)
      c[i * 3 + j] = a[i * 3] * b[j] + a[i * 3 + 1] * b[3 + j] + a[i * 3 + 2] * b[6 + j];
}

/* cv::findHomography(src, dst, 0) with 4 points (called at patch.cpp:161): with
 * 4 points OpenCV runs only HomographyEstimatorCallback::runKernel -- Hartley
 * normalisation (centroid, mean absolute deviation per axis), LtL accumulation,
 * eigenvector of the smallest eigenvalue, de-normalisation, division by H22.
 * No RANSAC and no LM refinement (npoints == 4). */
int orc_find_homography4(const float M[8], const float m[8], double H[9]) {
  const int count = 4;
  double cMx = 0, cMy = 0, cmx = 0, cmy = 0, sMx = 0, sMy = 0, smx = 0, smy = 0;
  for (int i = 0; i < count; ++i) {
    cmx += m[2 * i];
    cmy += m[2 * i + 1];
    cMx += M[2 * i];
    cMy += M[2 * i + 1];
  }
  cmx /= count; cmy /= count; cMx /= count; cMy /= count;
  for (int i = 0; i < count; ++i) {
    smx += fabs(m[2 * i] - cmx);
    smy += fabs(m[2 * i + 1] - cmy);
    sMx += fabs(M[2 * i] - cMx);
    sMy += fabs(M[2 * i + 1] - cMy);
  }
  if (fabs(smx) < DBL_EPSILON || fabs(smy) < DBL_EPSILON || fabs(sMx) < DBL_EPSILON ||
      fabs(sMy) < DBL_EPSILON)
    return 0;
  smx = count / smx; smy = count / smy; sMx = count / sMx; sMy = count / sMy;
  double invHnorm[9] = {1. / smx, 0, cmx, 0, 1. / smy, cmy, 0, 0, 1};
  double Hnorm2[9] = {sMx, 0, -cMx * sMx, 0, sMy, -cMy * sMy, 0, 0, 1};
  double LtL[81];
  memset(LtL, 0, sizeof(LtL));
  for (int i = 0; i < count; ++i) {
    double x = (m[2 * i] - cmx) * smx, y = (m[2 * i + 1] - cmy) * smy;
    double X = (M[2 * i] - cMx) * sMx, Y = (M[2 * i + 1] - cMy) * sMy;
    double Lx[9] = {X, Y, 1, 0, 0, 0, -x * X, -x * Y, -x};
    double Ly[9] = {0, 0, 0, X, Y, 1, -y * X, -y * Y, -y};
    for (int j = 0; j < 9; ++j)
      for (int k = j; k < 9; ++k) LtL[j * 9 + k] += Lx[j] * Lx[k] + Ly[j] * Ly[k];
  }
  for (int j = 0; j < 9; ++j)
    for (int k = 0; k < j; ++k) LtL[j * 9 + k] = LtL[k * 9 + j];
  double W[9], V[81];
  jacobi_eigen(LtL, 9, W, V);
  const double *H0 = &V[8 * 9];
  double Htemp[9], Hd[9];
  mat3_mul(invHnorm, H0, Htemp);
  mat3_mul(Htemp, Hnorm2, Hd);
  if (!(fabs(Hd[8]) > 0) || !isfinite(Hd[8])) return 0;
  double inv = 1. / Hd[8];
  for (int i = 0; i < 9; ++i) H[i] = Hd[i] * inv;
  for (int i = 0; i < 9; ++i)
    if (!isfinite(H[i])) return 0;
  return 1;
}

/* cv::invert(3x3, DECOMP_LU): adjugate / determinant; zero matrix if singular. */
static void invert3(const double S[9], double D[9]) {
  double d = S[0] * (S[4] * S[8] - S[5] * S[7]) - S[1] * (S[3] * S[8] - S[5] * S[6]) +
             S[2] * (S[3] * S[7] - S[4] * S[6]);
  if (d != 0.) {
    d = 1. / d;
    D[0] = (S[4] * S[8] - S[5] * S[7]) * d;
    D[1] = (S[2] * S[7] - S[1] * S[8]) * d;
    D[2] = (S[1] * S[5] - S[2] * S[4]) * d;
    D[3] = (S[5] * S[6] - S[3] * S[8]) * d;
    D[4] = (S[0] * S[8] - S[2] * S[6]) * d;
    D[5] = (S[2] * S[3] - S[0] * S[5]) * d;
    D[6] = (S[3] * S[7] - S[4] * S[6]) * d;
    D[7] = (S[1] * S[6] - S[0] * S[7]) * d;
    D[8] = (S[0] * S[4] - S[1] * S[3]) * d;
  } else {
    memset(D, 0, sizeof(double) * 9);
  }
}

static int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* cv::warpPerspective(src, dst, H, (s,s), INTER_LINEAR, BORDER_REPLICATE) for
 * CV_8UC3 (called at optimization.cpp:53 on image(roi)).  Exact arithmetic of
 * OpenCV's WarpPerspectiveInvoker + remap fixed-point bilinear (SURVEY a4,
 * Appendix A): M = H^-1 in fp64; source coordinates quantised to 1/32 px with
 * round-half-even; taps clamped to the ROI; 15-bit integer weights. */
static void warp_with_M(const uint8_t *src, size_t stride, int w, int h, const double M[9], int s,
                        uint8_t *dst) {
  for (int y = 0; y < s; ++y) {
    double X0 = M[1] * y + M[2];
    double Y0 = M[4] * y + M[5];
    double W0 = M[7] * y + M[8];
    for (int x = 0; x < s; ++x) {
      double W = W0 + M[6] * x;
      W = W ? 32. / W : 0; /* INTER_TAB_SIZE = 32 */
      double fX = (X0 + M[0] * x) * W;
      double fY = (Y0 + M[3] * x) * W;
      fX = fX < (double)INT_MIN ? (double)INT_MIN : (fX > (double)INT_MAX ? (double)INT_MAX : fX);
      fY = fY < (double)INT_MIN ? (double)INT_MIN : (fY > (double)INT_MAX ? (double)INT_MAX : fY);
      int X = (int)nearbyint(fX); /* cvRound: round half to even */
      int Y = (int)nearbyint(fY);
      int sx = clampi(X >> 5, -32768, 32767), ax = X & 31; /* INTER_BITS = 5, short */
      int sy = clampi(Y >> 5, -32768, 32767), ay = Y & 31;
      int x0 = clampi(sx, 0, w - 1), x1 = clampi(sx + 1, 0, w - 1); /* BORDER_REPLICATE */
      int y0 = clampi(sy, 0, h - 1), y1 = clampi(sy + 1, 0, h - 1);
      int w00 = (32 - ax) * (32 - ay) * 32, w01 = ax * (32 - ay) * 32;
      int w10 = (32 - ax) * ay * 32, w11 = ax * ay * 32; /* sum = 1 << 15 */
      const uint8_t *p00 = src + (size_t)y0 * stride + 3 * x0;
      const uint8_t *p01 = src + (size_t)y0 * stride + 3 * x1;
      const uint8_t *p10 = src + (size_t)y1 * stride + 3 * x0;
      const uint8_t *p11 = src + (size_t)y1 * stride + 3 * x1;
      for (int c = 0; c < 3; ++c) {
        int v = p00[c] * w00 + p01[c] * w01 + p10[c] * w10 + p11[c] * w11;
        dst[(y * s + x) * 3 + c] = (uint8_t)((v + (1 << 14)) >> 15); /* COEF_BITS = 15 */
      }
    }
  }
}

void orc_warp_perspective(const uint8_t *src, size_t stride, int w, int h, const double H[9],
                          int s, uint8_t *dst) {
  double M[9];
  invert3(H, M);
  warp_with_M(src, stride, w, h, M, s, dst);
}

/* The exact projective map cell [0,s]^2 -> quad, i.e. what
 * invert(findHomography(quad -> cell)) is mathematically (closed form of the
 * unit-square-to-quadrilateral mapping).  Used in homography mode 1, see
 * orc_set_homography_mode. Returns 0 for a degenerate quad. */
int orc_cell_to_quad(const float q[8], int s, double M[9]) {
  double qx0 = q[0], qy0 = q[1], qx1 = q[2], qy1 = q[3], qx2 = q[4], qy2 = q[5], qx3 = q[6],
         qy3 = q[7];
  double sx = qx0 - qx1 + qx2 - qx3, sy = qy0 - qy1 + qy2 - qy3;
  double dx1 = qx1 - qx2, dx2 = qx3 - qx2, dy1 = qy1 - qy2, dy2 = qy3 - qy2;
  double den = dx1 * dy2 - dx2 * dy1;
  if (!(den != 0.0)) return 0;
  double rden = 1.0 / den;
  double g = (sx * dy2 - dx2 * sy) * rden;
  double h = (dx1 * sy - sx * dy1) * rden;
  double inv_s = 1.0 / (double)s;
  M[0] = (qx1 - qx0 + g * qx1) * inv_s;
  M[1] = (qx3 - qx0 + h * qx3) * inv_s;
  M[2] = qx0;
  M[3] = (qy1 - qy0 + g * qy1) * inv_s;
  M[4] = (qy3 - qy0 + h * qy3) * inv_s;
  M[5] = qy0;
  M[6] = g * inv_s;
  M[7] = h * inv_s;
  M[8] = 1.0;
  for (int i = 0; i < 8; ++i)
    if (!isfinite(M[i])) return 0;
  return 1;
}

static int g_homography_mode = 0;
void orc_set_homography_mode(int mode) { g_homography_mode = mode; }
int orc_get_homography_mode(void) { return g_homography_mode; }

/* cv::pyrDown on CV_8UC3 (the build's pyramid extension; the reference has no pyramid,
 * modules/image/Image.h:1-7): separable [1 4 6 4 1]/16, BORDER_REFLECT_101, output
 * ((w+1)/2, (h+1)/2), exact integer accumulation, (sum + 128) >> 8. */
static int reflect101(int i, int n) {
  if (n == 1) return 0;
  while (i < 0 || i >= n) i = i < 0 ? -i : 2 * n - 2 - i;
  return i;
}
void orc_pyrdown(const uint8_t *src, size_t sstride, int w, int h, uint8_t *dst, size_t dstride) {
  static const int k[5] = {1, 4, 6, 4, 1};
  int dw = (w + 1) / 2, dh = (h + 1) / 2;
  for (int y = 0; y < dh; ++y)
    for (int x = 0; x < dw; ++x)
      for (int c = 0; c < 3; ++c) {
        int sum = 0;
        for (int dy = -2; dy <= 2; ++dy) {
          const uint8_t *row = src + (size_t)reflect101(2 * y + dy, h) * sstride;
          int r = 0;
          for (int dx = -2; dx <= 2; ++dx) r += k[dx + 2] * row[3 * reflect101(2 * x + dx, w) + c];
          sum += k[dy + 2] * r;
        }
        dst[(size_t)y * dstride + 3 * x + c] = (uint8_t)((sum + 128) >> 8);
      }
}

/* cv::DownhillSolver::minimize (OpenCV >= 3.0 core/downhill_simplex.cpp), called
 * at optimization_opencv.cpp:63.  PARITY UNPINNED against upstream: the solver's
 * source is neither in the reference tree nor in this image; this restates the
 * published algorithm. */
static void nm_coord_sum(const double *p, int ndim, double *cs) {
  for (int j = 0; j < ndim; ++j) cs[j] = 0.;
  for (int i = 0; i <= ndim; ++i)
    for (int j = 0; j < ndim; ++j) cs[j] += p[i * ndim + j];
}
static double nm_try(orc_fn f, void *user, const double *p, const double *cs, int ndim, int ihi,
                     double alpha_, double *ptry, int *fcount) {
  double alpha = (1.0 - alpha_) / ndim;
  double beta = alpha - alpha_;
  for (int j = 0; j < ndim; ++j) ptry[j] = cs[j] * alpha - p[ihi * ndim + j] * beta;
  ++*fcount;
  return f(ptry, user);
}
static void nm_replace(double *p, double *cs, double *y, int ndim, int ihi, double alpha_,
                       double ytry) {
  double alpha = (1.0 - alpha_) / ndim;
  double beta = alpha - alpha_;
  for (int j = 0; j < ndim; ++j) p[ihi * ndim + j] = cs[j] * alpha - p[ihi * ndim + j] * beta;
  y[ihi] = ytry;
  nm_coord_sum(p, ndim, cs);
}

double orc_downhill(orc_fn f, void *user, int ndim, double *x, const double *step, int nmax,
                    double eps, int *fcount_out) {
  double p[9 * 8], y[9], cs[8], buf[8];
  int i, j;
  if (ndim > 8) return NAN;
  /* createInitialSimplex: v_i = x0 + step_{i-1}/2 e_{i-1}; then v_0 = x0 - step/2 */
  for (j = 0; j < ndim; ++j) p[j] = x[j];
  for (i = 1; i <= ndim; ++i) {
    for (j = 0; j < ndim; ++j) p[i * ndim + j] = p[j];
    p[i * ndim + (i - 1)] += 0.5 * step[i - 1];
  }
  for (j = 0; j < ndim; ++j) p[j] -= 0.5 * step[j];

  int fcount = ndim + 1;
  for (i = 0; i <= ndim; ++i) y[i] = f(&p[i * ndim], user);
  nm_coord_sum(p, ndim, cs);

  for (;;) {
    int ilo = 0, ihi, inhi;
    if (y[0] > y[1]) {
      ihi = 0; inhi = 1;
    } else {
      ihi = 1; inhi = 0;
    }
    for (i = 0; i <= ndim; ++i) {
      double yval = y[i];
      if (yval <= y[ilo]) ilo = i;
      if (yval > y[ihi]) {
        inhi = ihi;
        ihi = i;
      } else if (yval > y[inhi] && i != ihi)
        inhi = i;
    }
    if (ilo == inhi || ilo == ihi) {
      for (i = 0; i <= ndim; ++i) {
        double yval = y[i];
        if (yval == y[ilo] && i != ihi && i != inhi) {
          ilo = i;
          break;
        }
      }
    }
    double error = fabs(y[ihi] - y[ilo]);
    double range = 0;
    for (j = 0; j < ndim; ++j) {
      double minval = p[j], maxval = p[j];
      for (i = 1; i <= ndim; ++i) {
        double pval = p[i * ndim + j];
        minval = pval < minval ? pval : minval;
        maxval = pval > maxval ? pval : maxval;
      }
      double r = fabs(maxval - minval);
      range = r > range ? r : range;
    }
    if (range <= eps || error <= eps || fcount >= nmax) {
      double ty = y[0];
      y[0] = y[ilo];
      y[ilo] = ty;
      for (j = 0; j < ndim; ++j) {
        double tp = p[j];
        p[j] = p[ilo * ndim + j];
        p[ilo * ndim + j] = tp;
      }
      break;
    }
    double y_lo = y[ilo], y_nhi = y[inhi], y_hi = y[ihi];
    double alpha = -1.0;
    double y_alpha = nm_try(f, user, p, cs, ndim, ihi, alpha, buf, &fcount);
    if (y_alpha < y_nhi) {
      if (y_alpha < y_lo) {
        double beta = -2.0;
        double y_beta = nm_try(f, user, p, cs, ndim, ihi, beta, buf, &fcount);
        if (y_beta < y_alpha) {
          alpha = beta;
          y_alpha = y_beta;
        }
      }
      nm_replace(p, cs, y, ndim, ihi, alpha, y_alpha);
    } else {
      double gamma = 0.5;
      double y_gamma = nm_try(f, user, p, cs, ndim, ihi, gamma, buf, &fcount);
      if (y_gamma < y_hi)
        nm_replace(p, cs, y, ndim, ihi, gamma, y_gamma);
      else {
        for (i = 0; i <= ndim; ++i) {
          if (i != ilo) {
            for (j = 0; j < ndim; ++j)
              p[i * ndim + j] = 0.5 * (p[i * ndim + j] + p[ilo * ndim + j]);
            y[i] = f(&p[i * ndim], user);
          }
        }
        fcount += ndim;
        nm_coord_sum(p, ndim, cs);
      }
    }
  }
  for (j = 0; j < ndim; ++j) x[j] = p[j];
  if (fcount_out) *fcount_out = fcount;
  return y[0];
}

/* ------------------------------------------------------------------------ */
/* methods/pmvs/patch.cpp                                                    */

/* Patch::GetProjectedXYAxisAndScale (patch.cpp:86-104) */
void orc_axes_scale(const orc_view *ref, const double nrm[3], const double pos[3], double xa[3],
                    double ya[3], double *dx) {
  double n = norm3(ref->xaxis);
  for (int j = 0; j < 3; ++j) xa[j] = ref->xaxis[j] / n; /* .normalized() */
  cross3(nrm, xa, ya);                                    /* NOT normalised */
  double c[2], px[2], q[3];
  orc_project(ref, pos, c);
  for (int j = 0; j < 3; ++j) q[j] = pos[j] + xa[j];
  orc_project(ref, q, px);
  double du = px[0] - c[0], dv = px[1] - c[1];
  *dx = sqrt(du * du + dv * dv);
}

/* Patch::ComputePatchToViewHomography (patch.cpp:111-151): projected quad relative to
 * the ROI (fp32 cv::Point2f) and the ROI itself. */
int orc_patch_quad(const orc_view *v, const double pos[3], const double ax[3], const double ay[3],
                   float pts[8], int roi[4]) {
  static const double sgn[4][2] = {{-1, -1}, {+1, -1}, {+1, +1}, {-1, +1}}; /* :119-123 */
  int tlx = v->width, tly = v->height, brx = 0, bry = 0;                     /* :126 */
  for (int i = 0; i < 4; ++i) {
    double X[3];
    for (int j = 0; j < 3; ++j) X[j] = pos[j] + sgn[i][0] * ax[j] + sgn[i][1] * ay[j];
    if (!orc_inside(v, X)) return 0; /* :130-132 */
    double uv[2];
    orc_project(v, X, uv);
    pts[2 * i] = (float)uv[0]; /* cv::Point2f, :134 */
    pts[2 * i + 1] = (float)uv[1];
    int cx = (int)ceil(uv[0]), cy = (int)ceil(uv[1]);
    int fx = (int)floor(uv[0]), fy = (int)floor(uv[1]);
    tlx = cx < tlx ? cx : tlx;
    tly = cy < tly ? cy : tly;
    brx = fx > brx ? fx : brx;
    bry = fy > bry ? fy : bry;
  }
  roi[0] = tlx;
  roi[1] = tly;
  roi[2] = brx - tlx;
  roi[3] = bry - tly;
  for (int i = 0; i < 4; ++i) { /* fp32 subtraction, :148-151 */
    pts[2 * i] -= (float)roi[0];
    pts[2 * i + 1] -= (float)roi[1];
  }
  return 1;
}

/* Patch::ComputePatchToViewHomography (patch.cpp:111-164) */
int orc_patch_homography(const orc_view *v, int cell_size, const double pos[3], const double ax[3],
                         const double ay[3], double H[9], int roi[4]) {
  float pts[8];
  if (!orc_patch_quad(v, pos, ax, ay, pts, roi)) return 0;
  float cs = (float)cell_size;
  float cell[8] = {0, 0, cs, 0, cs, cs, 0, cs};
  if (!orc_find_homography4(pts, cell, H)) return -1; /* empty H: OpenCV would throw */
  return 1;
}

/* ---- per-(patch, view) pyramid level ------------------------------------------
 * The reference has no semantics for this (modules/image/Image.h:1-7 is a placeholder,
 * options.h:10 `scale` is dead); the definition is: view v of a patch is read from
 * pyrDown^k(image_v) with P_k = diag(2^-k, 2^-k, 1) P, i.e. the reference path
 * (optimization.cpp:14-56) on the level views, k = the number of halvings that bring the longer
 * of the two quad sides through corner 0 (patch.cpp:119-123 corners, projected with the base
 * view) below px_per_cell * cell_size pixels.  Level tables are registered once
 * (orc_set_level_selection); level 0 of the table must be the `views` array the calls get. */
static const orc_view *g_lv_views = NULL;
static int g_lv_levels = 0, g_lv_nviews = 0;
static double g_lv_px = 1.5;
void orc_set_level_selection(const orc_view *levels, int n_levels, int n_views, double px_per_cell) {
  g_lv_views = (levels && n_levels > 1) ? levels : NULL;
  g_lv_levels = n_levels;
  g_lv_nviews = n_views;
  g_lv_px = px_per_cell;
}
int orc_pick_level(const orc_view *v, int cell_size, const double pos[3], const double ax[3],
                   const double ay[3], double px_per_cell, int max_up) {
  static const double sgn[3][2] = {{-1, -1}, {+1, -1}, {-1, +1}}; /* corners 0, 1, 3 */
  double uv[3][2];
  for (int i = 0; i < 3; ++i) {
    double X[3];
    for (int j = 0; j < 3; ++j) X[j] = pos[j] + sgn[i][0] * ax[j] + sgn[i][1] * ay[j];
    orc_project(v, X, uv[i]);
  }
  double du1 = uv[1][0] - uv[0][0], dv1 = uv[1][1] - uv[0][1];
  double du3 = uv[2][0] - uv[0][0], dv3 = uv[2][1] - uv[0][1];
  double a = du1 * du1 + dv1 * dv1, b = du3 * du3 + dv3 * dv3;
  double d2 = (b > a) ? b : a;
  double side = px_per_cell * (double)cell_size;
  double t = side * side;
  int k = 0;
  while (k < max_up && d2 >= t) {
    ++k;
    t = t * 4.0;
  }
  return k;
}

/* Optimization::GetProjectedTextures(normal, position, textures) (optimization.cpp:14-56).
 * `nrm` / `pos` are the ARGUMENTS: they only feed GetProjectedXYAxisAndScale (:24-26), i.e.
 * the patch axes and dx.  The four corners are built around `centre` = patch_.GetPosition(),
 * the STORED position (patch.cpp:119-123) -- so while Optimize() evaluates a trial depth, the
 * quad stays centred on the stored point and the depth only changes its scale. */
void orc_projected_textures(const orc_view *views, int ref, const int *vis, int nvis,
                            int cell_size, const double nrm[3], const double pos[3],
                            const double centre[3], uint8_t *tex, uint8_t *valid) {
  double xa[3], ya[3], dx;
  orc_axes_scale(&views[ref], nrm, pos, xa, ya, &dx);
  int s = cell_size;
  if (dx == 0 || !isfinite(dx)) { /* :27 LOG(FATAL); mapped to "all textures empty" */
    for (int k = 0; k < nvis; ++k) valid[k] = 0;
    return;
  }
  double scale = (double)(cell_size / 2) / dx; /* :30 integer division */
  double ax[3], ay[3];
  for (int j = 0; j < 3; ++j) {
    ax[j] = scale * xa[j];
    ay[j] = scale * ya[j];
  }
  for (int k = 0; k < nvis; ++k) {
    const orc_view *v = &views[vis[k]];
    if (g_lv_views) { /* the level this view is read at */
      int up = orc_pick_level(v, cell_size, centre, ax, ay, g_lv_px, g_lv_levels - 1);
      v = &g_lv_views[(size_t)up * g_lv_nviews + vis[k]];
    }
    double H[9], M[9];
    int roi[4];
    int ok;
    if (g_homography_mode == 0) {
      ok = orc_patch_homography(v, cell_size, centre, ax, ay, H, roi);
    } else {
      float pts[8];
      ok = orc_patch_quad(v, centre, ax, ay, pts, roi);
      if (ok > 0 && roi[2] > 0 && roi[3] > 0) ok = orc_cell_to_quad(pts, s, M) ? 1 : -1;
    }
    if (ok <= 0 || roi[2] <= 0 || roi[3] <= 0) { /* :45-48 */
      valid[k] = 0;
      continue;
    }
    const uint8_t *src = v->bgr + (size_t)roi[1] * v->stride + 3 * (size_t)roi[0];
    uint8_t *dst = tex + (size_t)k * s * s * 3;
    if (g_homography_mode == 0)
      orc_warp_perspective(src, v->stride, roi[2], roi[3], H, s, dst);
    else
      warp_with_M(src, v->stride, roi[2], roi[3], M, s, dst);
    valid[k] = 1;
  }
}

/* the level every (patch, visible view) pair is read at under the registered selection
 * (test statistics: out[i * vstride + k], -1 where there is no view or no selection) */
void orc_levels_batch(const orc_view *views, const float *pos, const float *nrm, const int *ref,
                      const int *nvis, const int *vis, int vstride, int n, int cell_size, int *out) {
  for (int i = 0; i < n; ++i) {
    double nn[3] = {nrm[3 * i], nrm[3 * i + 1], nrm[3 * i + 2]};
    double pp[3] = {pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]};
    double xa[3], ya[3], dx;
    for (int k = 0; k < vstride; ++k) out[(size_t)i * vstride + k] = -1;
    if (!g_lv_views) continue;
    orc_axes_scale(&views[ref[i]], nn, pp, xa, ya, &dx);
    if (dx == 0 || !isfinite(dx)) continue;
    double scale = (double)(cell_size / 2) / dx, ax[3], ay[3];
    for (int j = 0; j < 3; ++j) {
      ax[j] = scale * xa[j];
      ay[j] = scale * ya[j];
    }
    for (int k = 0; k < nvis[i] && k < vstride; ++k)
      out[(size_t)i * vstride + k] = orc_pick_level(&views[vis[(size_t)i * vstride + k]], cell_size,
                                                    pp, ax, ay, g_lv_px, g_lv_levels - 1);
  }
}

#define ORC_MAX_S 32
#define ORC_TEX_BYTES (ORC_MAX_S * ORC_MAX_S * 3)

static void f2d3(const float a[3], double b[3]) {
  b[0] = a[0];
  b[1] = a[1];
  b[2] = a[2];
}

/* scores of optimization.cpp:104-110 at (nrm, pos) in fp64, corners around `centre` */
static void scores_at(const orc_view *views, int ref, const int *vis, int nvis, int cell_size,
                      const double nrm[3], const double pos[3], const double centre[3],
                      double *scores, uint8_t *tex_out, uint8_t *valid_out) {
  int s = cell_size, tb = s * s * 3;
  uint8_t *tex = tex_out ? tex_out : (uint8_t *)malloc((size_t)(nvis > 0 ? nvis : 1) * tb);
  uint8_t *valid = valid_out ? valid_out : (uint8_t *)malloc((size_t)(nvis > 0 ? nvis : 1));
  orc_projected_textures(views, ref, vis, nvis, cell_size, nrm, pos, centre, tex, valid);
  for (int k = 1; k < nvis; ++k)
    scores[k - 1] = orc_ncc_bgr(valid[0] ? tex : NULL, valid[k] ? tex + (size_t)k * tb : NULL,
                                s * s);
  if (!tex_out) free(tex);
  if (!valid_out) free(valid);
}

void orc_scores(const orc_view *views, int ref, const int *vis, int nvis, int cell_size,
                const float nrm[3], const float pos[3], double *scores) {
  double n[3], p[3];
  f2d3(nrm, n);
  f2d3(pos, p);
  scores_at(views, ref, vis, nvis, cell_size, n, p, p, scores, NULL, NULL);
}

/* Optimization::FilterByErrorMeasurement (optimization.cpp:98-132), including the
 * off-by-one of :117-124: scores[i] belongs to visible[i+1] but visible[i - removed]
 * is erased. */
int orc_filter_by_error(const orc_view *views, int ref, int *vis, int *nvis_io, int cell_size,
                        const float nrm[3], const float pos[3], double thr, int min_visible) {
  int nvis = *nvis_io;
  int nscores = nvis > 0 ? nvis - 1 : 0;
  if (nscores == 0) return 0; /* :113-115 */
  double *scores = (double *)malloc(sizeof(double) * nscores);
  orc_scores(views, ref, vis, nvis, cell_size, nrm, pos, scores);
  int removed = 0;
  for (int i = 0; i < nscores; ++i) {
    if (scores[i] < thr) {
      int idx = i - removed; /* RemoveTrullyVisibleImage(score_index - removed_images) */
      for (int k = idx; k + 1 < nvis; ++k) vis[k] = vis[k + 1];
      --nvis;
      ++removed;
    }
  }
  free(scores);
  *nvis_io = nvis;
  return nvis >= min_visible ? 1 : 0; /* :127-131 */
}

/* Optimization::UnparametrizePatch (optimization.cpp:78-96) */
void orc_unparametrize(const orc_view *ref, const float nrm0[3], const float pos0[3], double depth,
                       double roll, double pitch, double nrm[3], double pos[3]) {
  double n0[3], p0[3];
  f2d3(nrm0, n0);
  f2d3(pos0, p0);
  for (int j = 0; j < 3; ++j) pos[j] = ref->center[j] + (1 + depth) * (p0[j] - ref->center[j]);
  double ca = cos(roll), sa = sin(roll), cb = cos(pitch), sb = sin(pitch);
  double R[9] = {cb, 0, -sb, sa * sb, ca, cb * sa, ca * sb, -sa, ca * cb};
  for (int i = 0; i < 3; ++i) nrm[i] = R[i * 3] * n0[0] + R[i * 3 + 1] * n0[1] + R[i * 3 + 2] * n0[2];
}

/* PatchOptimizationOpenCVFunctor::calc (optimization_opencv.cpp:14-39) */
double orc_objective(const orc_view *views, int ref, const int *vis, int nvis, int cell_size,
                     const float nrm0[3], const float pos0[3], const double x[3]) {
  double nrm[3], pos[3], centre[3];
  orc_unparametrize(&views[ref], nrm0, pos0, x[0], x[1], x[2], nrm, pos);
  f2d3(pos0, centre); /* patch_.GetPosition(): untouched until Optimize() returns (:69-70) */
  int nscores = nvis > 0 ? nvis - 1 : 0;
  if (nscores == 0) return 2; /* :30-32 */
  double sc[256];
  double *scores = nscores <= 256 ? sc : (double *)malloc(sizeof(double) * nscores);
  scores_at(views, ref, vis, nvis, cell_size, nrm, pos, centre, scores, NULL, NULL);
  double sum = 0.0;
  for (int k = 0; k < nscores; ++k) sum += 1 - scores[k]; /* :24, std::accumulate */
  if (scores != sc) free(scores);
  return sum / nscores;
}

typedef struct {
  const orc_view *views;
  int ref;
  const int *vis;
  int nvis;
  int cell_size;
  const float *nrm0, *pos0;
} obj_ctx;

static double obj_thunk(const double *x, void *user) {
  obj_ctx *c = (obj_ctx *)user;
  return orc_objective(c->views, c->ref, c->vis, c->nvis, c->cell_size, c->nrm0, c->pos0, x);
}

/* OptimizationOpenCV::Optimize (optimization_opencv.cpp:44-78) */
int orc_optimize(const orc_view *views, int ref, const int *vis, int nvis, int cell_size,
                 float nrm[3], float pos[3], const orc_params *prm, int *fcount, double xbest[3]) {
  obj_ctx c = {views, ref, vis, nvis, cell_size, nrm, pos};
  double x[3] = {0, 0, 0}; /* :51-52 */
  int fc = 0;
  orc_downhill(obj_thunk, &c, 3, x, prm->nm_step, prm->nm_max_evals, prm->nm_eps, &fc);
  double n[3], p[3];
  orc_unparametrize(&views[ref], nrm, pos, x[0], x[1], x[2], n, p); /* :66-68 */
  for (int j = 0; j < 3; ++j) { /* SetNormal/SetPosition: double -> fp32 (patch.h:38-53) */
    nrm[j] = (float)n[j];
    pos[j] = (float)p[j];
  }
  if (fcount) *fcount = fc;
  if (xbest) {
    xbest[0] = x[0];
    xbest[1] = x[1];
    xbest[2] = x[2];
  }
  return 1; /* :77 always true */
}

/* Patch::InitRelatedImages (patch.cpp:19-49) */
void orc_init_related_images(const orc_view *views, int n_views, int ref, const float nrm[3],
                             const float pos[3], double t_vis, double t_cand, int *vis, int *nvis,
                             int *cand, int *ncand) {
  double n[3], p[3];
  f2d3(nrm, n);
  f2d3(pos, p);
  int nv = 0, nc = 0;
  for (int v = 0; v < n_views; ++v) {
    if (v == ref) continue;
    if (!orc_inside(&views[v], p)) continue;
    double d[3] = {p[0] - views[v].center[0], p[1] - views[v].center[1],
                   p[2] - views[v].center[2]};
    double angle = acos(dot3(n, d) / norm3(d));
    if (angle < t_vis) {
      vis[nv++] = v;
    } else if (angle < t_cand) {
      if (cand) cand[nc] = v;
      nc++;
    }
  }
  *nvis = nv;
  if (ncand) *ncand = nc;
}

/* Patch::ComputeColor (patch.cpp:51-73).  No containing view => 0/0 in the
 * reference (UB cast); defined here as 0. */
void orc_compute_color(const orc_view *views, int n_views, const float pos[3], uint8_t rgb[3]) {
  double p[3], sum[3] = {0, 0, 0};
  f2d3(pos, p);
  int count = 0;
  for (int v = 0; v < n_views; ++v) {
    if (!orc_inside(&views[v], p)) continue;
    double uv[2];
    orc_project(&views[v], p, uv);
    const uint8_t *px = views[v].bgr + (size_t)((int)uv[1]) * views[v].stride + 3 * (size_t)((int)uv[0]);
    sum[0] += px[0];
    sum[1] += px[1];
    sum[2] += px[2];
    ++count;
  }
  if (count == 0) {
    rgb[0] = rgb[1] = rgb[2] = 0;
    return;
  }
  rgb[0] = (uint8_t)(sum[2] / count); /* r */
  rgb[1] = (uint8_t)(sum[1] / count); /* g */
  rgb[2] = (uint8_t)(sum[0] / count); /* b */
}

/* Seed::CreatePatchesFromPoints (seed.cpp:26-54), patches in point order. */
void orc_create_patches(const orc_view *views, int n_views, int n, const double *points,
                        double t_vis, double t_cand, float *pos, float *nrm, int *ref, int *nvis,
                        int *vis, int vstride) {
#pragma omp parallel for schedule(static)
  for (int i = 0; i < n; ++i) {
    const double *pt = points + 3 * i;
    double d0[3] = {pt[0] - views[0].center[0], pt[1] - views[0].center[1], pt[2] - views[0].center[2]};
    double min_distance = norm3(d0);
    int min_index = 0;
    for (int c = 1; c < n_views; ++c) {
      double d[3] = {pt[0] - views[c].center[0], pt[1] - views[c].center[1], pt[2] - views[c].center[2]};
      double distance = norm3(d);
      if (distance < min_distance) {
        min_index = c;
        min_distance = distance;
      }
    }
    double ptc[3] = {pt[0] - views[min_index].center[0], pt[1] - views[min_index].center[1],
                     pt[2] - views[min_index].center[2]};
    double nn = norm3(ptc);
    for (int j = 0; j < 3; ++j) {
      pos[3 * i + j] = (float)pt[j];
      nrm[3 * i + j] = (float)(ptc[j] / nn);
    }
    ref[i] = min_index;
    int *vi = vis + (size_t)i * vstride;
    for (int k = 0; k < vstride; ++k) vi[k] = -1;
    int nv, nc;
    orc_init_related_images(views, n_views, min_index, nrm + 3 * i, pos + 3 * i, t_vis, t_cand, vi,
                            &nv, NULL, &nc);
    nvis[i] = nv;
  }
}

/* ------------------------------------------------------------------------ */
/* batched drivers (Seed::FilterPatches / OptimizePatches, seed.cpp:110-144)  */

/* GetProjectedTextures(normal, position, textures) + the NCC loop for every patch, at trial
 * parameters: trial_nrm / trial_pos (fp64, n*3, either may be NULL = the patch's own) are the
 * arguments of optimization.cpp:14, the stored pos stays the corner centre (patch.cpp:119-123). */
void orc_score_at_batch(const orc_view *views, int n, const float *pos, const float *nrm,
                        const int *ref, const int *nvis, const int *vis, int vstride,
                        int cell_size, const double *trial_nrm, const double *trial_pos,
                        float *ncc, uint8_t *tex, uint8_t *valid) {
  int tb = cell_size * cell_size * 3;
#pragma omp parallel for schedule(dynamic, 64)
  for (int i = 0; i < n; ++i) {
    double sc[256];
    int nv = nvis[i];
    double nn[3], pp[3], cc[3];
    f2d3(nrm + 3 * i, nn);
    f2d3(pos + 3 * i, pp);
    f2d3(pos + 3 * i, cc);
    if (trial_nrm) memcpy(nn, trial_nrm + 3 * (size_t)i, sizeof(nn));
    if (trial_pos) memcpy(pp, trial_pos + 3 * (size_t)i, sizeof(pp));
    uint8_t *t = tex ? tex + (size_t)i * vstride * tb : NULL;
    uint8_t *vl = valid ? valid + (size_t)i * vstride : NULL;
    if (vl) memset(vl, 0, vstride);
    scores_at(views, ref[i], vis + (size_t)i * vstride, nv, cell_size, nn, pp, cc, sc, t, vl);
    ncc[(size_t)i * vstride] = 0.f;
    for (int k = 1; k < vstride; ++k)
      ncc[(size_t)i * vstride + k] = k < nv ? (float)sc[k - 1] : 0.f;
  }
}

void orc_score_batch(const orc_view *views, int n, const float *pos, const float *nrm,
                     const int *ref, const int *nvis, const int *vis, int vstride, int cell_size,
                     float *ncc, uint8_t *tex, uint8_t *valid) {
  orc_score_at_batch(views, n, pos, nrm, ref, nvis, vis, vstride, cell_size, NULL, NULL, ncc, tex,
                     valid);
}

void orc_filter_batch(const orc_view *views, int n, const float *pos, const float *nrm,
                      const int *ref, int *nvis, int *vis, int vstride, int cell_size, double thr,
                      int min_visible, uint8_t *keep) {
#pragma omp parallel for schedule(dynamic, 64)
  for (int i = 0; i < n; ++i) {
    int nv = nvis[i];
    int *vi = vis + (size_t)i * vstride;
    keep[i] = (uint8_t)orc_filter_by_error(views, ref[i], vi, &nv, cell_size, nrm + 3 * i,
                                           pos + 3 * i, thr, min_visible);
    for (int k = nv; k < nvis[i]; ++k) vi[k] = -1;
    nvis[i] = nv;
  }
}

void orc_refine_batch(const orc_view *views, int n, float *pos, float *nrm, const int *ref,
                      const int *nvis, const int *vis, int vstride, int cell_size,
                      const orc_params *prm, int *fcount, double *xbest) {
#pragma omp parallel for schedule(dynamic, 8)
  for (int i = 0; i < n; ++i) {
    int fc;
    double xb[3];
    orc_optimize(views, ref[i], vis + (size_t)i * vstride, nvis[i], cell_size, nrm + 3 * i,
                 pos + 3 * i, prm, &fc, xb);
    if (fcount) fcount[i] = fc;
    if (xbest) {
      xbest[3 * i] = xb[0];
      xbest[3 * i + 1] = xb[1];
      xbest[3 * i + 2] = xb[2];
    }
  }
}

void orc_visibility_batch(const orc_view *views, int n_views, int n, const float *pos,
                          const float *nrm, const int *ref, double t_vis, double t_cand, int *nvis,
                          int *vis, int *ncand, int *cand, int vstride) {
#pragma omp parallel for schedule(static)
  for (int i = 0; i < n; ++i) {
    int *vi = vis + (size_t)i * vstride;
    int *ci = cand ? cand + (size_t)i * vstride : NULL;
    int nv, nc;
    for (int k = 0; k < vstride; ++k) {
      vi[k] = -1;
      if (ci) ci[k] = -1;
    }
    orc_init_related_images(views, n_views, ref[i], nrm + 3 * i, pos + 3 * i, t_vis, t_cand, vi,
                            &nv, ci, &nc);
    nvis[i] = nv;
    if (ncand) ncand[i] = nc;
  }
}

/* ------------------------------------------------------------------------ */
/* methods/pmvs/patch_organizer.cpp + expand.cpp                             */

typedef struct {
  float pos[3], nrm[3];
  uint8_t rgb[3];
  int ref, nvis;
  int *vis;
} orc_patch;

struct orc_organizer {
  const orc_view *views;
  int n_views;
  orc_params prm;
  int *gw, *gh;
  uint8_t **grid; /* occupancy count per cell (only the count is ever read) */
  orc_patch *store;
  long long n, cap;
};

/* PatchOrganizer::AllocateViews (patch_organizer.cpp:32-40) */
orc_organizer *orc_organizer_create(const orc_view *views, int n_views, const orc_params *prm) {
  orc_organizer *o = (orc_organizer *)calloc(1, sizeof(*o));
  o->views = views;
  o->n_views = n_views;
  o->prm = *prm;
  o->gw = (int *)malloc(sizeof(int) * n_views);
  o->gh = (int *)malloc(sizeof(int) * n_views);
  o->grid = (uint8_t **)malloc(sizeof(uint8_t *) * n_views);
  for (int v = 0; v < n_views; ++v) {
    o->gw[v] = views[v].width / prm->grid_scale;
    o->gh[v] = views[v].height / prm->grid_scale;
    o->grid[v] = (uint8_t *)calloc((size_t)o->gw[v] * o->gh[v] + 1, 1);
  }
  o->cap = 1024;
  o->store = (orc_patch *)malloc(sizeof(orc_patch) * o->cap);
  return o;
}

void orc_organizer_destroy(orc_organizer *o) {
  if (!o) return;
  for (int v = 0; v < o->n_views; ++v) free(o->grid[v]);
  for (long long i = 0; i < o->n; ++i) free(o->store[i].vis);
  free(o->grid);
  free(o->gw);
  free(o->gh);
  free(o->store);
  free(o);
}

long long orc_organizer_size(const orc_organizer *o) { return o->n; }

const uint8_t *orc_organizer_grid(const orc_organizer *o, int view, int *gw, int *gh) {
  *gw = o->gw[view];
  *gh = o->gh[view];
  return o->grid[view];
}

/* PatchOrganizer::TryInsert + PatchGrid::TryInsert (patch_organizer.cpp:42-65, 15-30).
 * static_cast<size_t>(q) truncates toward zero: a quotient in (-1, 0) is cell 0 (defined
 * behaviour -- a refined seed can project to u or v in (-grid_scale, 0) in a view it kept
 * from the pre-refinement InitRelatedImages).  q <= -1, NaN and q >= 2^63 are UB in the
 * reference; they are defined here as out of bounds (x86-64's cvttsd2si gives 2^63 for
 * them, which fails the bounds test too). */
long long orc_organizer_try_insert(orc_organizer *o, const float pos[3], const float nrm[3], int ref,
                                   const int *vis, int nvis, int *ncells_out, int *cells_out) {
  double p[3];
  f2d3(pos, p);
  int ncells = 0;
  for (int k = 0; k < nvis; ++k) {
    int v = vis[k];
    double uv[2];
    orc_project(&o->views[v], p, uv);
    double qr = uv[1] / (double)o->prm.grid_scale, qc = uv[0] / (double)o->prm.grid_scale;
    if (!(qr > -1.0) || !(qc > -1.0) || !(qr < 2147483647.0) || !(qc < 2147483647.0)) continue;
    long long row = (long long)qr, col = (long long)qc; /* truncation toward zero */
    if (col < o->gw[v] && row < o->gh[v]) {
      uint8_t *cell = &o->grid[v][row * o->gw[v] + col];
      if (*cell < o->prm.max_patches_per_cell) {
        ++*cell; /* consumed even if the patch is rejected below */
        if (cells_out) {
          cells_out[3 * ncells] = v;
          cells_out[3 * ncells + 1] = (int)row;
          cells_out[3 * ncells + 2] = (int)col;
        }
        ++ncells;
      }
    }
  }
  if (ncells_out) *ncells_out = ncells;
  if (ncells > 1) { /* :58 */
    if (o->n == o->cap) {
      o->cap *= 2;
      o->store = (orc_patch *)realloc(o->store, sizeof(orc_patch) * o->cap);
    }
    orc_patch *q = &o->store[o->n];
    memcpy(q->pos, pos, sizeof(float) * 3);
    memcpy(q->nrm, nrm, sizeof(float) * 3);
    q->ref = ref;
    q->nvis = nvis;
    q->vis = (int *)malloc(sizeof(int) * (nvis > 0 ? nvis : 1));
    memcpy(q->vis, vis, sizeof(int) * nvis);
    orc_compute_color(o->views, o->n_views, pos, q->rgb); /* :60 */
    return o->n++;
  }
  return -1;
}

void orc_organizer_export(const orc_organizer *o, float *pos, float *nrm, uint8_t *rgb, int *ref,
                          int *nvis, int *vis, int vstride) {
  for (long long i = 0; i < o->n; ++i) {
    const orc_patch *q = &o->store[i];
    memcpy(pos + 3 * i, q->pos, sizeof(float) * 3);
    memcpy(nrm + 3 * i, q->nrm, sizeof(float) * 3);
    memcpy(rgb + 3 * i, q->rgb, 3);
    ref[i] = q->ref;
    nvis[i] = q->nvis;
    for (int k = 0; k < vstride; ++k) vis[i * vstride + k] = k < q->nvis ? q->vis[k] : -1;
  }
}

/* Expand::ExpandPatch (expand.cpp:103-143) */
int orc_expand_patch(const orc_view *views, int n_views, const orc_params *prm, int cell_size,
                     const float pos[3], const float nrm[3], int ref, const int *pvis, int pn,
                     float out_pos[12], float out_nrm[12], int *out_nvis, int *out_vis,
                     int *dir_out) {
  double n[3], p[3], xa[3], ya[3], dx;
  f2d3(nrm, n);
  f2d3(pos, p);
  orc_axes_scale(&views[ref], n, p, xa, ya, &dx); /* :108-110 */
  double scale = (double)prm->grid_scale / dx;     /* :112 */
  double dirs[4][3];
  for (int j = 0; j < 3; ++j) { /* :114-116 */
    dirs[0][j] = xa[j];
    dirs[1][j] = -xa[j];
    dirs[2][j] = ya[j];
    dirs[3][j] = -ya[j];
  }
  int count = 0;
  for (int d = 0; d < 4; ++d) {
    float cp[3], cn[3];
    for (int j = 0; j < 3; ++j) {
      cp[j] = (float)(p[j] + scale * dirs[d][j]); /* :123-127 SetPosition -> fp32 */
      cn[j] = nrm[j];
    }
    int fc;
    orc_optimize(views, ref, pvis, pn, cell_size, cn, cp, prm, &fc, NULL); /* :129-130 */
    int *cv = out_vis + (size_t)count * n_views;
    int nv, nc;
    orc_init_related_images(views, n_views, ref, cn, cp, prm->visible_threshold,
                            prm->candidate_threshold, cv, &nv, NULL, &nc); /* :132 */
    if (orc_filter_by_error(views, ref, cv, &nv, cell_size, cn, cp, prm->score_threshold,
                            prm->minimum_visible_image)) { /* :133-136 */
      memcpy(out_pos + 3 * count, cp, sizeof(float) * 3);
      memcpy(out_nrm + 3 * count, cn, sizeof(float) * 3);
      out_nvis[count] = nv;
      if (dir_out) dir_out[count] = d;
      ++count;
    }
  }
  return count;
}

typedef struct {
  int count;
  float pos[12], nrm[12];
  int nvis[4];
  int *vis; /* 4 * n_views */
} orc_children;

static void expand_one(const orc_organizer *o, int cell_size, long long idx, orc_children *ch) {
  const orc_patch *q = &o->store[idx];
  ch->count = 0;
  if (q->nvis < 2) return; /* expand.cpp:69 */
  ch->count = orc_expand_patch(o->views, o->n_views, &o->prm, cell_size, q->pos, q->nrm, q->ref,
                               q->vis, q->nvis, ch->pos, ch->nrm, ch->nvis, ch->vis, NULL);
}

/* Expand::ExpandPatches (expand.cpp:34-101), literal single-thread FIFO. */
long long orc_expand_patches_fifo(orc_organizer *o, int cell_size, long long max_pops) {
  long long qcap = 1024, qh = 0, qt = 0;
  long long *queue = (long long *)malloc(sizeof(long long) * qcap);
  for (long long i = 0; i < o->n; ++i) { /* :45-48 */
    if (qt == qcap) queue = (long long *)realloc(queue, sizeof(long long) * (qcap *= 2));
    queue[qt++] = i;
  }
  orc_children ch;
  ch.vis = (int *)malloc(sizeof(int) * 4 * (o->n_views > 0 ? o->n_views : 1));
  long long pops = 0;
  long long limit = max_pops >= 0 ? max_pops : o->prm.max_pops;
  while (qh < qt) {
    long long idx = queue[qh++];
    int ref = o->store[idx].ref;
    expand_one(o, cell_size, idx, &ch);
    for (int c = 0; c < ch.count; ++c) {
      long long ins = orc_organizer_try_insert(o, ch.pos + 3 * c, ch.nrm + 3 * c, ref,
                                               ch.vis + (size_t)c * o->n_views, ch.nvis[c], NULL,
                                               NULL);
      if (ins >= 0) {
        if (qt == qcap) queue = (long long *)realloc(queue, sizeof(long long) * (qcap *= 2));
        queue[qt++] = ins;
      }
    }
    ++pops;
    if (pops >= limit) break; /* :95-97 */
  }
  free(ch.vis);
  free(queue);
  return pops;
}

/* The same FIFO order evaluated level-synchronously: ExpandPatch never reads the
 * grids (expand.cpp:103-143), so the children of one BFS level can be computed
 * in parallel and inserted afterwards in (parent order, direction order) --
 * exactly the order the 1-thread FIFO produces. */
long long orc_expand_patches(orc_organizer *o, int cell_size, int max_levels) {
  long long begin = 0, pops = 0;
  int level = 0;
  while (begin < o->n && (max_levels < 0 || level < max_levels)) {
    long long end = o->n, nf = end - begin;
    if (pops + nf > o->prm.max_pops) nf = o->prm.max_pops - pops;
    orc_children *ch = (orc_children *)malloc(sizeof(orc_children) * (size_t)nf);
    for (long long i = 0; i < nf; ++i)
      ch[i].vis = (int *)malloc(sizeof(int) * 4 * (o->n_views > 0 ? o->n_views : 1));
#pragma omp parallel for schedule(dynamic, 4)
    for (long long i = 0; i < nf; ++i) expand_one(o, cell_size, begin + i, &ch[i]);
    for (long long i = 0; i < nf; ++i) {
      int ref = o->store[begin + i].ref;
      for (int c = 0; c < ch[i].count; ++c)
        orc_organizer_try_insert(o, ch[i].pos + 3 * c, ch[i].nrm + 3 * c, ref,
                                 ch[i].vis + (size_t)c * o->n_views, ch[i].nvis[c], NULL, NULL);
      free(ch[i].vis);
    }
    free(ch);
    pops += nf;
    if (pops >= o->prm.max_pops) break;
    begin = end;
    ++level;
  }
  return pops;
}
