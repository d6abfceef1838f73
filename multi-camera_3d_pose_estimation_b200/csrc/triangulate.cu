// Batched multi-view DLT triangulation for sm_100a.
//
// What it computes (per joint): the eigenvector h of the smallest eigenvalue of B = A^T A, where A
// stacks, for each view v, the rows  w_v (y_v P_v[2] - P_v[1])  and  w_v (P_v[0] - x_v P_v[2])
// (reference utils.py:21-28), and returns h[0:3] / h[3] (utils.py:34).  MODE_TOP2 first picks the two
// highest-score views per joint and undistorts them (pose_estimation.py:35-52, utils.py:1314-1315).
//
// How: one thread per joint; a persistent CTA streams 256-joint tiles of the keypoint array through a
// multi-stage shared-memory ring filled by 1-D TMA bulk copies (cp.async.bulk + mbarrier), accumulates the
// 10 unique entries of B in registers (double), and finds the smallest eigenpair with a secular-equation
// Newton / Rayleigh-quotient iteration on the (X,1) parametrisation:
//      (M - lam I) X = -b,   lam <- lam + (c - lam + b.X) / (1 + X.X),      B = [[M, b], [b^T, c]]
// (cubic convergence, 2 LDL^T solves for ordinary rigs).  Joints where that parametrisation breaks down
// (pivot loss, no convergence, point at infinity) fall back to a cyclic 4x4 Jacobi eigensolver in
// registers.  Results leave through a shared-memory tile and a TMA bulk store.
#include "mc3d_common.cuh"
#include <math.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>
#include <type_traits>

namespace mc3d {

constexpr int TRI_TILE = 256;          // joints per tile == threads per CTA

struct TriParams {
    double P[MC3D_MAX_VIEWS][12];
    double K[MC3D_MAX_VIEWS][9];
    double dist[MC3D_MAX_VIEWS][5];
    int n_views;
    int undistort;
    int flags;
    int layout;
    // mixed-precision path (float storage)
    mc3d_tri_start_pair start[MC3D_TRI_MAX_START];     // closed-form two-view starting points (fill_start_pairs)
    int n_start;
    float rig2;                       // mean squared distance of the camera centres from the world origin
};

// ---- small dense helpers ----------------------------------------------------------------------

// Solve (M) z = r for SPD 3x3 M (lower triangle m00 m10 m11 m20 m21 m22) by LDL^T.
// Returns false when a pivot is not safely positive.  dmin = smallest pivot.
__device__ __forceinline__ bool ldl3_solve(double m00, double m10, double m11, double m20, double m21, double m22,
                                           double r0, double r1, double r2, double &z0, double &z1, double &z2,
                                           double &dmin) {
    const double tiny = 1e-13;
    if (!(m00 > 0.0)) return false;
    const double i0 = 1.0 / m00;
    const double l10 = m10 * i0, l20 = m20 * i0;
    const double d1 = fma(-l10, m10, m11);
    if (!(d1 > tiny * m11)) return false;
    const double i1 = 1.0 / d1;
    const double t21 = fma(-l20, m10, m21);
    const double l21 = t21 * i1;
    const double d2 = fma(-l21, t21, fma(-l20, m20, m22));
    if (!(d2 > tiny * m22)) return false;
    const double i2 = 1.0 / d2;
    const double y1 = fma(-l10, r0, r1);
    const double y2 = fma(-l21, y1, fma(-l20, r0, r2));
    z2 = y2 * i2;
    z1 = fma(y1, i1, -l21 * z2);
    z0 = fma(r0, i0, -fma(l10, z1, l20 * z2));
    dmin = fmin(m00, fmin(d1, d2));
    return true;
}

// Cyclic Jacobi on the symmetric 4x4 B (10 unique entries, order 00 10 11 20 21 22 30 31 32 33);
// returns the eigenvector of the smallest eigenvalue de-homogenised.  Fallback path: kept out of line.
__device__ __noinline__ void jacobi4_smallest(const double *Bsym, double &X0, double &X1, double &X2) {
    double a[4][4], v[4][4];
    a[0][0] = Bsym[0];
    a[1][0] = a[0][1] = Bsym[1];
    a[1][1] = Bsym[2];
    a[2][0] = a[0][2] = Bsym[3];
    a[2][1] = a[1][2] = Bsym[4];
    a[2][2] = Bsym[5];
    a[3][0] = a[0][3] = Bsym[6];
    a[3][1] = a[1][3] = Bsym[7];
    a[3][2] = a[2][3] = Bsym[8];
    a[3][3] = Bsym[9];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) v[i][j] = (i == j) ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 24; ++sweep) {
        bool done = true;
#pragma unroll
        for (int p = 0; p < 3; ++p) {
#pragma unroll
            for (int q = p + 1; q < 4; ++q) {
                const double apq = a[p][q];
                // relative criterion suits graded positive semi-definite matrices
                if (!(fabs(apq) > 1e-17 * sqrt(fabs(a[p][p] * a[q][q])) && apq != 0.0)) continue;
                done = false;
                const double theta = (a[q][q] - a[p][p]) / (2.0 * apq);
                const double t = copysign(1.0, theta) / (fabs(theta) + sqrt(fma(theta, theta, 1.0)));
                const double c = rsqrt(fma(t, t, 1.0));
                const double s = t * c;
                a[p][p] = fma(-t, apq, a[p][p]);
                a[q][q] = fma(t, apq, a[q][q]);
                a[p][q] = a[q][p] = 0.0;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    if (k != p && k != q) {
                        const double akp = a[k][p], akq = a[k][q];
                        a[k][p] = a[p][k] = fma(c, akp, -s * akq);
                        a[k][q] = a[q][k] = fma(s, akp, c * akq);
                    }
                    const double vkp = v[k][p], vkq = v[k][q];
                    v[k][p] = fma(c, vkp, -s * vkq);
                    v[k][q] = fma(s, vkp, c * vkq);
                }
            }
        }
        if (done) break;
    }
    int best = 0;
    double ev = a[0][0];
#pragma unroll
    for (int i = 1; i < 4; ++i)
        if (a[i][i] < ev) { ev = a[i][i]; best = i; }
    double h0 = v[0][0], h1 = v[1][0], h2 = v[2][0], h3 = v[3][0];
#pragma unroll
    for (int i = 1; i < 4; ++i)
        if (best == i) { h0 = v[0][i]; h1 = v[1][i]; h2 = v[2][i]; h3 = v[3][i]; }
    const double inv = 1.0 / h3;
    X0 = h0 * inv;
    X1 = h1 * inv;
    X2 = h2 * inv;
}

// Smallest eigenpair of B via secular Newton; false -> caller should use Jacobi.
__device__ __forceinline__ bool secular_newton(const double *B, double &X0, double &X1, double &X2) {
    const double m00 = B[0], m10 = B[1], m11 = B[2], m20 = B[3], m21 = B[4], m22 = B[5];
    const double b0 = B[6], b1 = B[7], b2 = B[8], c = B[9];
    double lam = 0.0;
#pragma unroll 1
    for (int it = 0; it < 12; ++it) {
        double z0, z1, z2, dmin;
        if (!ldl3_solve(m00 - lam, m10, m11 - lam, m20, m21, m22 - lam, -b0, -b1, -b2, z0, z1, z2, dmin)) return false;
        const double f = (c - lam) + fma(b0, z0, fma(b1, z1, b2 * z2));
        const double dl = f / (1.0 + fma(z0, z0, fma(z1, z1, z2 * z2)));
        X0 = z0; X1 = z1; X2 = z2;
        if (!(fabs(dl) <= 1.0e300)) return false;            // NaN / inf
        if (fabs(dl) <= 1e-13 * dmin) return true;            // X solved with lam accurate to |dl|
        lam += dl;
    }
    return false;
}

// cv.undistortPoints(pt, K, dist, None, K): 5 fixed-point iterations then re-projection with K.
__device__ __forceinline__ void undistort_px(double &u, double &v, const double *K, const double *d) {
    const double fx = K[0], fy = K[4], cx = K[2], cy = K[5];
    const double k1 = d[0], k2 = d[1], p1 = d[2], p2 = d[3], k3 = d[4];
    const double x0 = (u - cx) / fx, y0 = (v - cy) / fy;
    double x = x0, y = y0;
#pragma unroll 1
    for (int j = 0; j < 5; ++j) {
        const double r2 = fma(x, x, y * y);
        const double icdist = 1.0 / fma(fma(fma(k3, r2, k2), r2, k1), r2, 1.0);
        if (icdist < 0.0) { x = x0; y = y0; break; }
        const double dx = fma(2.0 * p1 * x, y, p2 * fma(2.0 * x, x, r2));
        const double dy = fma(p1, fma(2.0 * y, y, r2), 2.0 * p2 * x * y);
        x = (x0 - dx) * icdist;
        y = (y0 - dy) * icdist;
    }
    const double ww = 1.0 / fma(K[6], x, fma(K[7], y, K[8]));
    u = fma(K[0], x, fma(K[1], y, K[2])) * ww;
    v = fma(K[3], x, fma(K[4], y, K[5])) * ww;
}

__device__ __forceinline__ void accumulate_view(double *B, double x, double y, double w, const double *p) {
    const double wy = w * y, wx = w * x;
    double a[4], c[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        a[k] = fma(wy, p[8 + k], -(w * p[4 + k]));
        c[k] = fma(-wx, p[8 + k], w * p[k]);
    }
    B[0] = fma(a[0], a[0], fma(c[0], c[0], B[0]));
    B[1] = fma(a[1], a[0], fma(c[1], c[0], B[1]));
    B[2] = fma(a[1], a[1], fma(c[1], c[1], B[2]));
    B[3] = fma(a[2], a[0], fma(c[2], c[0], B[3]));
    B[4] = fma(a[2], a[1], fma(c[2], c[1], B[4]));
    B[5] = fma(a[2], a[2], fma(c[2], c[2], B[5]));
    B[6] = fma(a[3], a[0], fma(c[3], c[0], B[6]));
    B[7] = fma(a[3], a[1], fma(c[3], c[1], B[7]));
    B[8] = fma(a[3], a[2], fma(c[3], c[2], B[8]));
    B[9] = fma(a[3], a[3], fma(c[3], c[3], B[9]));
}

__device__ __forceinline__ float rcp_fast(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// All-double solve of one joint straight from its shared-memory row (cold path of the mixed kernel).
__device__ __noinline__ void solve_double_from_row(const TriParams &prm, const float *row, int nv, bool l3v, double &X0,
                                                   double &X1, double &X2) {
    double B[10];
    for (int i = 0; i < 10; ++i) B[i] = 0.0;
    int n_used = 0;
    for (int v = 0; v < nv; ++v) {
        const double x = (double)(l3v ? row[v] : row[3 * v]);
        const double y = (double)(l3v ? row[nv + v] : row[3 * v + 1]);
        const double w = (double)(l3v ? row[2 * nv + v] : row[3 * v + 2]);
        n_used += (w != 0.0);
        accumulate_view(B, x, y, w, prm.P[v]);
    }
    X0 = X1 = X2 = NAN;
    if (!(fabs((B[0] + B[2]) + (B[5] + B[9])) <= 1.0e300) || n_used < 2) return;
    if (!secular_newton(B, X0, X1, X2)) jacobi4_smallest(B, X0, X1, X2);
}

// x, y, w of views [vg, vg + G) of one joint from its shared-memory row; 128-bit loads when the row allows it.
template <int V, int LAYOUT, int G>
__device__ __forceinline__ void load_group(const float *row, int vg, float (&gx)[G], float (&gy)[G], float (&gw)[G]) {
    if constexpr (G == 4 && V % 4 == 0) {
        if constexpr (LAYOUT == MC3D_LAYOUT_3V) {
            const float4 a = *reinterpret_cast<const float4 *>(row + vg);
            const float4 b = *reinterpret_cast<const float4 *>(row + V + vg);
            const float4 c = *reinterpret_cast<const float4 *>(row + 2 * V + vg);
            gx[0] = a.x; gx[1] = a.y; gx[2] = a.z; gx[3] = a.w;
            gy[0] = b.x; gy[1] = b.y; gy[2] = b.z; gy[3] = b.w;
            gw[0] = c.x; gw[1] = c.y; gw[2] = c.z; gw[3] = c.w;
        } else {
            const float4 a = *reinterpret_cast<const float4 *>(row + 3 * vg);
            const float4 b = *reinterpret_cast<const float4 *>(row + 3 * vg + 4);
            const float4 c = *reinterpret_cast<const float4 *>(row + 3 * vg + 8);
            gx[0] = a.x; gy[0] = a.y; gw[0] = a.z;
            gx[1] = a.w; gy[1] = b.x; gw[1] = b.y;
            gx[2] = b.z; gy[2] = b.w; gw[2] = c.x;
            gx[3] = c.y; gy[3] = c.z; gw[3] = c.w;
        }
    } else {
#pragma unroll
        for (int i = 0; i < G; ++i) {
            const int v = vg + i;
            gx[i] = LAYOUT == MC3D_LAYOUT_3V ? row[v] : row[3 * v];
            gy[i] = LAYOUT == MC3D_LAYOUT_3V ? row[V + v] : row[3 * v + 1];
            gw[i] = LAYOUT == MC3D_LAYOUT_3V ? row[2 * V + v] : row[3 * v + 2];
        }
    }
}

// ---- mixed-precision kernel (float storage) -----------------------------------------------------------------------
// Forming B in double costs >= 304 DFMA per joint at 8 views -- more than the FP64 pipe delivers at the HBM rate -- and
// forming it in float loses the answer (up to 1e-3 mm).  The float kernel therefore touches the full view set ONCE:
//   start  a closed-form two-view point Xp: the ray of view A cut by the plane that view B's pixel spans along its
//          epipolar direction (18 float operations and one reciprocal, all constants prepared on the host in double).
//          It lies a few millimetres from the minimiser (the pixel noise of two views); a second pair of views serves
//          the joints whose first pair has a zero-weight view, and a joint without a usable pair starts from the world
//          origin (first pass = float solve over all views, the second pass finishes);
//   pass   ONE loop over all V views accumulates the float normal matrix M~ = sum w^2 (a a^T + c c^T) (packed FFMA2: the
//          two rows of a view ride in one float2) and the gradient g = A_m^T A (Xp,1) with the residual A (Xp,1)
//          evaluated in DOUBLE (only the cancellation y (P2.X) - P1.X needs it: 11 DFMA per joint-view) and rounded to float;
//   solve  the correction solves the eigen-equations to first order in lam/M with one float LDL^T factorisation:
//             e1 = -M~^-1 g,   lam = (|r|^2 + g.e1) / (1 + |Xp + e1|^2),   e = e1 + lam M~^-1 (Xp + e1)
//          (exact fixed point (M - lam) X = -b; the neglected term is (lam/mu_min)^2 |X|, checked per joint).
// The float solve is accurate to ~3e-7 |e|, so a correction of up to a few centimetres lands within 1e-5 mm; larger
// corrections take another pass from the updated point, and whatever cannot be handled (failed pivots, non-finite
// input, lam not small, no convergence) goes to the all-double solver.  M~ and the float rows only precondition the
// iteration: its fixed point is set by the double residuals.
struct __align__(16) CamF {
    double P[12];        // projection rows as given (double residuals)
    float2 p2pm[3];      // (-P2_k, P2_k), k < 3
    float2 p01[3];       // (P0_k, -P1_k), k < 3:  (x, y) * p2pm + p01 = (P0_k - x P2_k, y P2_k - P1_k) = (c_k, a_k)
};
static_assert(sizeof(CamF) == 144, "CamF layout");
static_assert(sizeof(mc3d_tri_start_pair) == 96 && offsetof(mc3d_tri_start_pair, C) == 36 && offsetof(mc3d_tri_start_pair, ua) == 48 &&
                  offsetof(mc3d_tri_start_pair, alpha) == 72 && offsetof(mc3d_tri_start_pair, ka) == 80 &&
                  offsetof(mc3d_tri_start_pair, view_a) == 88,
              "pair_start reads mc3d_tri_start_pair as 24 consecutive four-byte fields");

#ifndef MC3D_TRI_PIVOT
#define MC3D_TRI_PIVOT 1.0e-3f
#endif
// Two joints per thread; every float2 below holds (joint 0, joint 1) unless it is a row pair (c_k, a_k).
struct Ldl3f2 {
    float2 i0, l10, l20, i1, l21, i2;
};
__device__ __forceinline__ float2 neg2(float2 a) { return make_float2(-a.x, -a.y); }
__device__ __forceinline__ float2 rcp_fast2(float2 a) { return make_float2(rcp_fast(a.x), rcp_fast(a.y)); }
__device__ __forceinline__ float2 bcast2(float a) { return make_float2(a, a); }
__device__ __forceinline__ float2 dot3_2(float2 a0, float2 a1, float2 a2, float2 b0, float2 b1, float2 b2) {
    return __ffma2_rn(a0, b0, __ffma2_rn(a1, b1, __fmul2_rn(a2, b2)));
}
// LDL^T of the SPD 3x3 systems of both joints (lower triangles m00 m10 m11 m20 m21 m22).  Approximate reciprocals are
// fine: the solve only preconditions an iteration whose fixed point is set by residuals evaluated in double.
__device__ __forceinline__ void ldl3_factor_f2(float2 m00, float2 m10, float2 m11, float2 m20, float2 m21, float2 m22,
                                               Ldl3f2 &f, bool &ok0, bool &ok1) {
    f.i0 = rcp_fast2(m00);
    f.l10 = __fmul2_rn(m10, f.i0);
    f.l20 = __fmul2_rn(m20, f.i0);
    const float2 d1 = __ffma2_rn(neg2(f.l10), m10, m11);
    f.i1 = rcp_fast2(d1);
    const float2 t21 = __ffma2_rn(neg2(f.l20), m10, m21);
    f.l21 = __fmul2_rn(t21, f.i1);
    const float2 d2 = __ffma2_rn(neg2(f.l21), t21, __ffma2_rn(neg2(f.l20), m20, m22));
    f.i2 = rcp_fast2(d2);
    // a pivot below 1e-3 of its diagonal entry means cond(M~) >~ 1e3: two nearly collinear rays, or ONE usable view (rank 2,
    // the pivot is rounding noise ~1e-7) -- the all-double solver takes those joints (and counts their usable views)
    const float2 t1 = __fmul2_rn(bcast2(MC3D_TRI_PIVOT), m11), t2 = __fmul2_rn(bcast2(MC3D_TRI_PIVOT), m22);
    ok0 = (m00.x > 0.f) & (d1.x > t1.x) & (d2.x > t2.x);
    ok1 = (m00.y > 0.f) & (d1.y > t1.y) & (d2.y > t2.y);
}
__device__ __forceinline__ void ldl3_apply_f2(const Ldl3f2 &f, float2 r0, float2 r1, float2 r2, float2 &z0, float2 &z1,
                                              float2 &z2) {
    const float2 y1 = __ffma2_rn(neg2(f.l10), r0, r1);
    const float2 y2 = __ffma2_rn(neg2(f.l21), y1, __ffma2_rn(neg2(f.l20), r0, r2));
    z2 = __fmul2_rn(y2, f.i2);
    z1 = __ffma2_rn(y1, f.i1, __fmul2_rn(neg2(f.l21), z2));
    z0 = __ffma2_rn(r0, f.i0, neg2(__ffma2_rn(f.l10, z1, __fmul2_rn(f.l20, z2))));
}
// (a[0][i].x + a[0][i].y, a[1][i].x + a[1][i].y): the two halves of a row-packed accumulator, per joint
#define MC3D_HSUM2(a, i) make_float2((a)[0][i].x + (a)[0][i].y, (a)[1][i].x + (a)[1][i].y)

// x, y, w of one (runtime) view of a joint's shared-memory row
template <int V, int LAYOUT>
__device__ __forceinline__ void load_view(const float *row, int v, float &x, float &y, float &w) {
    if constexpr (LAYOUT == MC3D_LAYOUT_3V) {
        x = row[v]; y = row[V + v]; w = row[2 * V + v];
    } else {
        x = row[3 * v]; y = row[3 * v + 1]; w = row[3 * v + 2];
    }
}

// Closed-form two-view start (constants: fill_start_pairs).  X(s) = C_A + s D_A with D_A = H_A (x_A, y_A, 1) is the ray of
// view A; view B contributes the one linear constraint along its epipolar direction (alpha, beta):
//     [(alpha P0 + beta P1) - z P2]_B . (X(s), 1) = 0,   z = alpha x_B + beta y_B
// whose solution is s = (z kb - ka) / (ua.q - z ub.q), q = (x_A, y_A, 1).  wmin = min(w_A, w_B).
template <int V, int LAYOUT>
__device__ __forceinline__ void pair_start(const mc3d_tri_start_pair *pc, int va, int vb, const float *row, float &X0,
                                           float &X1, float &X2, float &wmin) {
    const float *pf = reinterpret_cast<const float *>(pc);              // 24 four-byte fields, 16-byte aligned
    const float4 h0 = *reinterpret_cast<const float4 *>(pf);            // H00 H01 H02 H10
    const float4 h1 = *reinterpret_cast<const float4 *>(pf + 4);        // H11 H12 H20 H21
    const float4 h2 = *reinterpret_cast<const float4 *>(pf + 8);        // H22 C0 C1 C2
    const float4 u0 = *reinterpret_cast<const float4 *>(pf + 12);       // ua0 ua1 ua2 ub0
    const float4 u1 = *reinterpret_cast<const float4 *>(pf + 16);       // ub1 ub2 alpha beta
    const float2 kk = *reinterpret_cast<const float2 *>(pf + 20);       // ka kb
    float xa, ya, wa, xb, yb, wb;
    load_view<V, LAYOUT>(row, va, xa, ya, wa);
    load_view<V, LAYOUT>(row, vb, xb, yb, wb);
    const float z = fmaf(u1.z, xb, u1.w * yb);
    const float ta = fmaf(u0.x, xa, fmaf(u0.y, ya, u0.z));
    const float tb = fmaf(u0.w, xa, fmaf(u1.x, ya, u1.y));
    const float s = fmaf(z, kk.y, -kk.x) * rcp_fast(fmaf(-z, tb, ta));
    X0 = fmaf(s, fmaf(h0.x, xa, fmaf(h0.y, ya, h0.z)), h2.y);
    X1 = fmaf(s, fmaf(h0.w, xa, fmaf(h1.x, ya, h1.y)), h2.z);
    X2 = fmaf(s, fmaf(h1.z, xa, fmaf(h1.w, ya, h2.x)), h2.w);
    wmin = fminf(wa, wb);
}

#ifndef MC3D_TRI_UNR_LIMIT
#define MC3D_TRI_UNR_LIMIT 16           // view loop fully unrolled (no spills at 128 registers; +5 % at V = 8, +6 % at V = 16)
#endif
// The float solve leaves an error of ~cond(M~) 1e-7 |e|, so a correction is final when cond |e| is small against the range:
//     |e|^2 (tr(M~) / smallest pivot)^2 <= MC3D_TRI_ACCEPT (|X|^2 + rig scale^2)
// tr / pivot is ~6 for eight views on a ring (|e| <~ 3 % of the range, the limit of the first-order treatment of lam) and
// up to ~3e3 for the worst pair of views the pivot test lets through (|e| <~ 1e-4 of the range); either way the error stays
// below ~2e-8 of the range.  Larger corrections take another pass from the updated point.
#ifndef MC3D_TRI_ACCEPT
#define MC3D_TRI_ACCEPT 4.0e-2f
#endif

// All NJ = 2 NP joints of a thread (the solve runs on NP packed pairs).  state: 0 = Xo holds the result, 1 = the all-double solver has to take the joint.
template <int V, int LAYOUT, int NP>
__device__ __forceinline__ void solve_mixed(const CamF *__restrict__ cam, const mc3d_tri_start_pair *__restrict__ pairs,
                                            const int4 sv, int n_pairs, float rig2, const float *const (&rows)[2 * NP],
                                            float (&Xo)[2 * NP][3], int (&state)[2 * NP]) {
    constexpr int NJ = 2 * NP;
    constexpr int G = (V % 4 == 0) ? 4 : V;                     // views per load group
    double Xd[NJ][3];                                           // float-valued in the first pass
    bool own[NJ];                                               // started from its own first two usable views (below)
#pragma unroll
    for (int j = 0; j < NJ; ++j) own[j] = false;
    // ---- start ---------------------------------------------------------------------------------------------------
    {
        const float lim = 1.0e4f * (rig2 + 1.f);                // a start farther out than 100 rig radii is not one
        bool have[NJ];
        float Xf[NJ][3];
#pragma unroll
        for (int j = 0; j < NJ; ++j) { have[j] = false; Xf[j][0] = Xf[j][1] = Xf[j][2] = 0.f; }
        if (n_pairs > 0) {
#pragma unroll
            for (int j = 0; j < NJ; ++j) {
                float s0, s1, s2, wm;
                pair_start<V, LAYOUT>(&pairs[0], sv.x, sv.y, rows[j], s0, s1, s2, wm);
                if (wm > 0.f && fmaf(s0, s0, fmaf(s1, s1, s2 * s2)) <= lim) { Xf[j][0] = s0; Xf[j][1] = s1; Xf[j][2] = s2; have[j] = true; }
            }
            // warp-uniform: clean input never evaluates the later pairs; a joint takes the first pair it sees with both views
            // (the later pairs' views come from their shared-memory records: as a register array indexed in this loop they
            // went to local memory, and the cascade -- which most warps run as soon as 1 % of the views are unusable -- waited on it)
#pragma unroll 1
            for (int p = 1; p < n_pairs; ++p) {
                bool all_have = true;
#pragma unroll
                for (int j = 0; j < NJ; ++j) all_have = all_have && have[j];
                if (!__any_sync(0xffffffffu, !all_have)) break;
                // (pair 1 -- the one most warps reach -- has its views in registers like pair 0)
                const int va = p == 1 ? sv.z : pairs[p].view_a, vb = p == 1 ? sv.w : pairs[p].view_b;
#pragma unroll
                for (int j = 0; j < NJ; ++j) {
                    if (!__any_sync(0xffffffffu, !have[j])) continue;   // this half of the tile has all its starts
                    float s0, s1, s2, wm;
                    pair_start<V, LAYOUT>(&pairs[p], va, vb, rows[j], s0, s1, s2, wm);
                    if (!have[j] && wm > 0.f && fmaf(s0, s0, fmaf(s1, s1, s2 * s2)) <= lim) { Xf[j][0] = s0; Xf[j][1] = s1; Xf[j][2] = s2; have[j] = true; }
                }
            }
        }
        // A joint that sees none of the planned pairs with both views (typically one left with two or three usable views)
        // would start from the origin, and the residual there trips the rounding guard: ~1 000 warp-instructions in the all-
        // double solver for what is an ordinary two-view point.  It gets a two-view start from its OWN first two usable views
        // instead: the weighted normal equations of their four rows in double (3 x 3, LDL^T) -- for a joint with exactly two
        // usable views that is the answer up to the eigenvalue term.  Such a joint keeps its float result only when those
        // views are well apart (tr / pivot <= 100 below, at least as strict as a pivot above 1 % of its diagonal): with two
        // nearly opposite views the float correction is off by up to 3e-3 mm, and those go to the double solver as before.
        if constexpr (V <= 8) {                                    // (sixteen views: the instantiation would spill at 128 registers)
            bool all_have = true;
#pragma unroll
            for (int j = 0; j < NJ; ++j) all_have = all_have && have[j];
            if (__any_sync(0xffffffffu, !all_have)) {
#pragma unroll
                for (int j = 0; j < NJ; ++j) {
                    if (have[j]) continue;
                    double m00 = 0.0, m10 = 0.0, m11 = 0.0, m20 = 0.0, m21 = 0.0, m22 = 0.0, b0 = 0.0, b1 = 0.0, b2 = 0.0;
                    int cnt = 0;
#pragma unroll 1
                    for (int v = 0; v < V && cnt < 2; ++v) {
                        float x, y, w;
                        load_view<V, LAYOUT>(rows[j], v, x, y, w);
                        if (!(w > 0.f)) continue;
                        const double *Pv = cam[v].P;
                        const double xd = (double)x, yd = (double)y, wd = (double)w;
                        double rc[4], ra[4];
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            rc[k] = wd * fma(-xd, Pv[8 + k], Pv[k]);         // w (P0 - x P2)
                            ra[k] = wd * fma(yd, Pv[8 + k], -Pv[4 + k]);     // w (y P2 - P1)
                        }
                        m00 = fma(rc[0], rc[0], fma(ra[0], ra[0], m00)); m10 = fma(rc[1], rc[0], fma(ra[1], ra[0], m10));
                        m11 = fma(rc[1], rc[1], fma(ra[1], ra[1], m11)); m20 = fma(rc[2], rc[0], fma(ra[2], ra[0], m20));
                        m21 = fma(rc[2], rc[1], fma(ra[2], ra[1], m21)); m22 = fma(rc[2], rc[2], fma(ra[2], ra[2], m22));
                        b0 = fma(rc[0], rc[3], fma(ra[0], ra[3], b0)); b1 = fma(rc[1], rc[3], fma(ra[1], ra[3], b1));
                        b2 = fma(rc[2], rc[3], fma(ra[2], ra[3], b2));
                        ++cnt;
                    }
                    double z0, z1, z2, dmin;
                    if (cnt == 2 && ldl3_solve(m00, m10, m11, m20, m21, m22, -b0, -b1, -b2, z0, z1, z2, dmin)) {
                        const float f0 = (float)z0, f1 = (float)z1, f2 = (float)z2;
                        if (fmaf(f0, f0, fmaf(f1, f1, f2 * f2)) <= lim) { Xf[j][0] = f0; Xf[j][1] = f1; Xf[j][2] = f2; own[j] = true; }
                    }
                }
            }
        }
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            state[j] = 2;
#pragma unroll
            for (int k = 0; k < 3; ++k) { Xd[j][k] = (double)Xf[j][k]; Xo[j][k] = NAN; }
        }
    }
    // ---- pass + solve from Xp, repeated only when the correction was large ----------------------------------------
    constexpr int UNR = (V > MC3D_TRI_UNR_LIMIT) ? 1 : (V / G);
#pragma unroll 1
    for (int it = 0; it < 6; ++it) {
        bool any = false;
#pragma unroll
        for (int j = 0; j < NJ; ++j) any = any || state[j] == 2;
        if (!any) break;
        float2 aM[NJ][6], ag[NJ][3], arr[NJ];
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            arr[j] = make_float2(0.f, 0.f);
#pragma unroll
            for (int i = 0; i < 6; ++i) aM[j][i] = make_float2(0.f, 0.f);
#pragma unroll
            for (int i = 0; i < 3; ++i) ag[j][i] = make_float2(0.f, 0.f);
        }
#pragma unroll UNR
        for (int vg = 0; vg < V; vg += G) {
            float gx[NJ][G], gy[NJ][G], gw[NJ][G];
#pragma unroll
            for (int j = 0; j < NJ; ++j) load_group<V, LAYOUT, G>(rows[j], vg, gx[j], gy[j], gw[j]);
#pragma unroll
            for (int i = 0; i < G; ++i) {
                const CamF &c = cam[vg + i];
                double P[12];
#pragma unroll
                for (int k = 0; k < 6; ++k) {
                    const double2 t = *reinterpret_cast<const double2 *>(&c.P[2 * k]);
                    P[2 * k] = t.x; P[2 * k + 1] = t.y;
                }
                float2 p2pm[3], p01[3];
                {
                    const float4 q0 = *reinterpret_cast<const float4 *>(&c.p2pm[0]);
                    const float4 q1 = *reinterpret_cast<const float4 *>(&c.p2pm[2]);
                    const float4 q2 = *reinterpret_cast<const float4 *>(&c.p01[1]);
                    p2pm[0] = make_float2(q0.x, q0.y); p2pm[1] = make_float2(q0.z, q0.w); p2pm[2] = make_float2(q1.x, q1.y);
                    p01[0] = make_float2(q1.z, q1.w); p01[1] = make_float2(q2.x, q2.y); p01[2] = make_float2(q2.z, q2.w);
                }
#pragma unroll
                for (int j = 0; j < NJ; ++j) {
                    const float x = gx[j][i], y = gy[j][i], w = gw[j][i];
                    const float2 xy = make_float2(x, y), ww = make_float2(w, w);
                    float2 ac[3];                        // weighted rows (c_k, a_k) = w (P0_k - x P2_k, y P2_k - P1_k)
#pragma unroll
                    for (int k = 0; k < 3; ++k) ac[k] = __fmul2_rn(ww, __ffma2_rn(xy, p2pm[k], p01[k]));
                    aM[j][0] = __ffma2_rn(ac[0], ac[0], aM[j][0]);
                    aM[j][1] = __ffma2_rn(ac[1], ac[0], aM[j][1]);
                    aM[j][2] = __ffma2_rn(ac[1], ac[1], aM[j][2]);
                    aM[j][3] = __ffma2_rn(ac[2], ac[0], aM[j][3]);
                    aM[j][4] = __ffma2_rn(ac[2], ac[1], aM[j][4]);
                    aM[j][5] = __ffma2_rn(ac[2], ac[2], aM[j][5]);
                    const double d0 = fma(P[0], Xd[j][0], fma(P[1], Xd[j][1], fma(P[2], Xd[j][2], P[3])));
                    const double d1 = fma(P[4], Xd[j][0], fma(P[5], Xd[j][1], fma(P[6], Xd[j][2], P[7])));
                    const double d2 = fma(P[8], Xd[j][0], fma(P[9], Xd[j][1], fma(P[10], Xd[j][2], P[11])));
                    const float rc = (float)fma(-(double)x, d2, d0);    // P0.X - x (P2.X)
                    const float ra = (float)fma((double)y, d2, -d1);    // y (P2.X) - P1.X : the cancellation is in double
                    const float2 t = __fmul2_rn(ww, make_float2(rc, ra));
#pragma unroll
                    for (int k = 0; k < 3; ++k) ag[j][k] = __ffma2_rn(t, ac[k], ag[j][k]);
                    arr[j] = __ffma2_rn(t, t, arr[j]);
                }
            }
        }
#pragma unroll
        for (int q = 0; q < NP; ++q) {
            Ldl3f2 f;
            bool ok[2];
            const float2 m00 = MC3D_HSUM2(aM + 2 * q, 0), m11 = MC3D_HSUM2(aM + 2 * q, 2), m22 = MC3D_HSUM2(aM + 2 * q, 5);
            const float2 tr2 = __fadd2_rn(__fadd2_rn(m00, m11), m22);
            ldl3_factor_f2(m00, MC3D_HSUM2(aM + 2 * q, 1), m11, MC3D_HSUM2(aM + 2 * q, 3), MC3D_HSUM2(aM + 2 * q, 4), m22, f, ok[0], ok[1]);
            const float2 g0 = MC3D_HSUM2(ag + 2 * q, 0), g1 = MC3D_HSUM2(ag + 2 * q, 1), g2 = MC3D_HSUM2(ag + 2 * q, 2);
            float2 e0, e1, e2;
            ldl3_apply_f2(f, neg2(g0), neg2(g1), neg2(g2), e0, e1, e2);
            const float2 rs = make_float2(arr[2 * q].x + arr[2 * q].y, arr[2 * q + 1].x + arr[2 * q + 1].y);           // |A (Xp,1)|^2
            const float2 rr = __fadd2_rn(rs, dot3_2(g0, g1, g2, e0, e1, e2));
            float Xf[2][3];                                       // float(Xd): exact in the first pass
#pragma unroll
            for (int j = 0; j < 2; ++j)
#pragma unroll
                for (int k = 0; k < 3; ++k) Xf[j][k] = (float)Xd[2 * q + j][k];
            const float2 x0 = __fadd2_rn(make_float2(Xf[0][0], Xf[1][0]), e0);
            const float2 x1 = __fadd2_rn(make_float2(Xf[0][1], Xf[1][1]), e1);
            const float2 x2 = __fadd2_rn(make_float2(Xf[0][2], Xf[1][2]), e2);
            const float2 nx2 = dot3_2(x0, x1, x2, x0, x1, x2);
            const float2 lam2 = __fmul2_rn(rr, rcp_fast2(__fadd2_rn(bcast2(1.f), nx2)));
            float2 h0, h1, h2;
            ldl3_apply_f2(f, __fmul2_rn(lam2, x0), __fmul2_rn(lam2, x1), __fmul2_rn(lam2, x2), h0, h1, h2);
            e0 = __fadd2_rn(e0, h0); e1 = __fadd2_rn(e1, h1); e2 = __fadd2_rn(e2, h2);
            const float2 ne2 = dot3_2(e0, e1, e2, e0, e1, e2);
            const float2 lim2 = __fmul2_rn(bcast2(MC3D_TRI_ACCEPT), __fadd2_rn(nx2, bcast2(rig2)));
            // first-order treatment of lam needs lam << smallest pivot of M~
            const float imax[2] = {fmaxf(f.i0.x, fmaxf(f.i1.x, f.i2.x)), fmaxf(f.i0.y, fmaxf(f.i1.y, f.i2.y))};
            const float2 kap = __fmul2_rn(tr2, make_float2(imax[0], imax[1]));          // ~cond(M~)
            const float2 nk2 = __fmul2_rn(ne2, __fmul2_rn(kap, kap));
            // g is a float sum of terms ~|r| |a| that cancel: its rounding error ~1e-7 sqrt(|r|^2 tr) moves the solution by that
            // times 1 / (smallest pivot), whatever the size of e -- no further pass can repair it, the double solver has to
            const float2 gn2 = __fmul2_rn(__fmul2_rn(rs, kap), make_float2(imax[0], imax[1]));   // |r|^2 tr / pivot^2
            const float ne[2] = {ne2.x, ne2.y}, nk[2] = {nk2.x, nk2.y}, gn[2] = {gn2.x, gn2.y}, lam[2] = {lam2.x, lam2.y}, lim[2] = {lim2.x, lim2.y};
            const float kp[2] = {kap.x, kap.y};
            const float es[2][3] = {{e0.x, e1.x, e2.x}, {e0.y, e1.y, e2.y}};
#pragma unroll
            for (int jj = 0; jj < 2; ++jj) {
                const int j = 2 * q + jj;
                if (state[j] != 2) continue;
                if (!ok[jj] || !(ne[jj] <= 3.0e38f) || !(fabsf(lam[jj]) * imax[jj] <= 3.0e-5f) || !(gn[jj] <= lim[jj]) ||
                    (own[j] && !(kp[jj] <= 100.f))) { state[j] = 1; continue; }
                if (nk[jj] <= lim[jj]) {
                    if (it == 0) {                                   // Xd is float-valued: float(Xd + e) is one float addition
                        Xo[j][0] = Xf[jj][0] + es[jj][0]; Xo[j][1] = Xf[jj][1] + es[jj][1]; Xo[j][2] = Xf[jj][2] + es[jj][2];
                    } else {
                        Xo[j][0] = (float)(Xd[j][0] + (double)es[jj][0]); Xo[j][1] = (float)(Xd[j][1] + (double)es[jj][1]);
                        Xo[j][2] = (float)(Xd[j][2] + (double)es[jj][2]);
                    }
                    state[j] = 0;
                } else {                                             // another pass from the updated point
                    Xd[j][0] += (double)es[jj][0]; Xd[j][1] += (double)es[jj][1]; Xd[j][2] += (double)es[jj][2];
                }
            }
        }
    }
#pragma unroll
    for (int j = 0; j < NJ; ++j)
        if (state[j] == 2) state[j] = 1;                         // did not converge in 6 passes
}

#ifndef MC3D_TRI_MBLOCKS
#define MC3D_TRI_MBLOCKS 4
#endif
constexpr int TRI_MWARPS = 4;                   // warps per CTA of the mixed kernel
constexpr int TRI_MTHREADS = 32 * TRI_MWARPS;
#ifndef MC3D_TRI_NP
#define MC3D_TRI_NP 1
#endif
constexpr int TRI_NP = MC3D_TRI_NP;             // packed joint pairs per thread
constexpr int TRI_NJ = 2 * TRI_NP;              // joints per thread: lane l owns joints l, l + 32, ... of its warp's tile
constexpr int TRI_WTILE = 32 * TRI_NJ;          // joints per warp tile

// Resident CTAs per SM the mixed kernel is compiled for.  Measured at V = 8: 4 CTAs x 128 registers and 3 CTAs x 162
// registers run at the same speed (the extra registers buy the compiler what the fourth CTA's warps would hide), so the
// instantiations that spill at 128 registers -- the (N, 3, V) layout, whose x / y pairs need moves, and V = 4 -- get 3.
#ifndef MC3D_TRI_V16_STAGES
#define MC3D_TRI_V16_STAGES 1
#endif
// Input stages per warp: two, except for 16 views, where a 64-joint tile is 12 KB and two stages per warp would hold the SM
// to 8 warps; with one stage 16 warps are resident.
__host__ __device__ constexpr int tri_mixed_stages(int V) { return (V * TRI_NP >= 16) ? MC3D_TRI_V16_STAGES : 2; }
__host__ __device__ constexpr int tri_mixed_blocks(int V, int layout) {
    return (V * TRI_NP >= 16) ? (tri_mixed_stages(V) == 1 ? (layout == MC3D_LAYOUT_3V ? MC3D_TRI_MBLOCKS - 1 : MC3D_TRI_MBLOCKS) : 2)
                              : ((layout == MC3D_LAYOUT_3V && V >= 4) || V == 4) ? MC3D_TRI_MBLOCKS - 1 : MC3D_TRI_MBLOCKS;
}

__device__ __noinline__ void mixed_cold_fix(const TriParams &prm, int nv, const float *row, bool l3v, float *o) {
    double f0, f1, f2;
    solve_double_from_row(prm, row, nv, l3v, f0, f1, f2);
    o[0] = (float)f0; o[1] = (float)f1; o[2] = (float)f2;
}

// Mixed-precision kernel (float storage, weighted mode, V in {2,3,4,8,16}, no undistortion).  Every WARP is its own
// pipeline: 64-joint tiles of the keypoint array stream through the warp's private shared-memory ring, filled by 1-D TMA
// bulk copies that lane 0 issues (cp.async.bulk + mbarrier complete_tx), and the 64 results leave through the warp's
// private output tile and a TMA bulk store.  There is no block-wide barrier after the set-up (the tile loop of the
// block-cooperative version spent 7 % of its issue slots' time in one, and a warp that needs a second pass -- an
// unusable view -- now delays nobody else).  Lane l solves joints l, l + 32, ... of the tile together (packed float2
// arithmetic, the camera constants of a view fetched once for all of them); the ragged tail (less than a tile) is one
// more tile of the next warp in line, filled with plain loads.
template <int V, int LAYOUT>
__global__ void __launch_bounds__(TRI_MTHREADS, tri_mixed_blocks(V, LAYOUT))
triangulate_mixed_kernel(const float *__restrict__ kpts, float *__restrict__ out, long long n,
                         const __grid_constant__ TriParams prm) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int row_elems = 3 * V;
    constexpr uint32_t stage_bytes = (uint32_t)(TRI_WTILE * row_elems * sizeof(float));
    constexpr uint32_t otile_bytes = (uint32_t)(TRI_WTILE * 3 * sizeof(float));
    constexpr int NST = tri_mixed_stages(V);           // input stages per warp
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // layout: [warp][NST] input stages | [warp][2] output tiles | [warp][2] mbarriers | cameras | start pairs
    unsigned char *ring = smem_raw + (size_t)warp * NST * stage_bytes;
    float *otile = reinterpret_cast<float *>(smem_raw + (size_t)TRI_MWARPS * NST * stage_bytes + (size_t)warp * 2 * otile_bytes);
    unsigned char *fixed = smem_raw + (size_t)TRI_MWARPS * (NST * stage_bytes + 2 * otile_bytes);
    uint64_t *full = reinterpret_cast<uint64_t *>(fixed) + warp * 2;
    CamF *cam = reinterpret_cast<CamF *>(fixed + TRI_MWARPS * 2 * sizeof(uint64_t));
    mc3d_tri_start_pair *pairs = reinterpret_cast<mc3d_tri_start_pair *>(cam + V);
    if (lane == 0) {
        mbar_init(&full[0], 1);
        mbar_init(&full[1], 1);
        fence_mbar_init();
    }
    if (tid < V) {
        CamF c;
#pragma unroll
        for (int k = 0; k < 12; ++k) c.P[k] = prm.P[tid][k];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            c.p2pm[k] = make_float2(-(float)prm.P[tid][8 + k], (float)prm.P[tid][8 + k]);
            c.p01[k] = make_float2((float)prm.P[tid][k], -(float)prm.P[tid][4 + k]);
        }
        cam[tid] = c;
    } else if (tid < V + MC3D_TRI_MAX_START) {
        pairs[tid - V] = prm.start[tid - V];
    }
    __syncthreads();                                   // the only block-wide barrier

    const unsigned n_tiles = (unsigned)(n / TRI_WTILE);           // full tiles (the host keeps n / 64 below 2^31)
    const int tail = (int)(n - (long long)n_tiles * TRI_WTILE);
    const unsigned gw = blockIdx.x * TRI_MWARPS + warp, gstride = gridDim.x * TRI_MWARPS;
    const unsigned my_tiles = (gw < n_tiles) ? (n_tiles - gw + gstride - 1) / gstride : 0;
    // shared-window addresses and running global pointers (lane 0 is the only one that uses them)
    const uint32_t ring_sa = smem_u32(ring), full_sa = smem_u32(full), ot_sa = smem_u32(otile);
    const unsigned char *src = reinterpret_cast<const unsigned char *>(kpts) + (size_t)gw * stage_bytes;    // tile k + 1
    unsigned char *dst = reinterpret_cast<unsigned char *>(out) + (size_t)gw * otile_bytes;                 // tile k
    const uint32_t src_step = gstride * stage_bytes, dst_step = gstride * otile_bytes;                      // < 2^32
    if (lane == 0 && my_tiles > 0) {
        mbar_arrive_expect_tx_sa(full_sa, stage_bytes);
        bulk_g2s_sa(ring_sa, src, stage_bytes, full_sa);
    }
    src += src_step;
    const int n_pairs = prm.n_start;
    // views of the starting pairs straight from the parameter bank (uniform): no shared-memory load in front of the row loads
    const int4 sv = make_int4(prm.start[0].view_a, prm.start[0].view_b, prm.start[1].view_a, prm.start[1].view_b);
    for (unsigned k = 0; k < my_tiles; ++k) {
        const uint32_t b = k & 1u;                    // output buffer; input stage when there are two
        const uint32_t sb = NST == 2 ? b : 0u;
        // two stages: refill the stage of iteration k - 1 (every lane left it before that iteration's last __syncwarp)
        if (NST == 2 && lane == 0 && k + 1 < my_tiles) {
            mbar_arrive_expect_tx_sa(full_sa + 8u * (b ^ 1u), stage_bytes);
            bulk_g2s_sa(ring_sa + (b ^ 1u) * stage_bytes, src, stage_bytes, full_sa + 8u * (b ^ 1u));
        }
        if (NST == 2) src += src_step;
        const float *stage = reinterpret_cast<const float *>(ring + sb * stage_bytes);
        mbar_wait_sa(full_sa + 8u * sb, NST == 2 ? ((k >> 1) & 1u) : (k & 1u));
        const float *rows[TRI_NJ];
#pragma unroll
        for (int j = 0; j < TRI_NJ; ++j) rows[j] = stage + (lane + 32 * j) * row_elems;
        float Xo[TRI_NJ][3];
        int state[TRI_NJ];
        solve_mixed<V, LAYOUT, TRI_NP>(cam, pairs, sv, n_pairs, prm.rig2, rows, Xo, state);
        float *ot = otile + b * (TRI_WTILE * 3);
        if (lane == 0) bulk_wait_read<1>();           // the store of iteration k - 2 has left this output buffer
        __syncwarp();
#pragma unroll
        for (int j = 0; j < TRI_NJ; ++j) {
            float *o = ot + (lane + 32 * j) * 3;
            o[0] = Xo[j][0]; o[1] = Xo[j][1]; o[2] = Xo[j][2];
            if (state[j] == 1) mixed_cold_fix(prm, V, rows[j], LAYOUT == MC3D_LAYOUT_3V, o);
        }
        fence_proxy_async_smem();
        __syncwarp();                                 // the stage is consumed by every lane; output tile complete
        if (lane == 0) {
            if (NST == 1 && k + 1 < my_tiles) {       // one stage (16 views: twice the resident warps instead of a second stage):
                mbar_arrive_expect_tx_sa(full_sa, stage_bytes);                       // the refill overlaps the output epilogue only,
                bulk_g2s_sa(ring_sa, src, stage_bytes, full_sa);                      // the other warps of the SM cover its latency
            }
            bulk_s2g_sa(dst, ot_sa + b * otile_bytes, otile_bytes);
            bulk_commit();
        }
        if (NST == 1) src += src_step;
        dst += dst_step;
    }
    if (tail > 0 && gw == n_tiles % gstride) {        // ragged tail: the next warp in line, plain loads and stores
        float *stage = reinterpret_cast<float *>(ring);            // every bulk load of this warp has been consumed
        const float *tsrc = kpts + (size_t)n_tiles * TRI_WTILE * row_elems;
        for (int i = lane; i < tail * row_elems; i += 32) stage[i] = tsrc[i];
        __syncwarp();
        const float *rows[TRI_NJ];
#pragma unroll
        for (int j = 0; j < TRI_NJ; ++j) rows[j] = stage + (lane + 32 * j < tail ? lane + 32 * j : 0) * row_elems;   // idle slots repeat joint 0
        float Xo[TRI_NJ][3];
        int state[TRI_NJ];
        solve_mixed<V, LAYOUT, TRI_NP>(cam, pairs, sv, n_pairs, prm.rig2, rows, Xo, state);
        float *tdst = out + (size_t)n_tiles * TRI_WTILE * 3;
#pragma unroll
        for (int j = 0; j < TRI_NJ; ++j) {
            const int slot = lane + 32 * j;
            if (slot < tail) {
                float fix[3] = {Xo[j][0], Xo[j][1], Xo[j][2]};
                if (state[j] == 1) mixed_cold_fix(prm, V, rows[j], LAYOUT == MC3D_LAYOUT_3V, fix);
                tdst[slot * 3 + 0] = fix[0]; tdst[slot * 3 + 1] = fix[1]; tdst[slot * 3 + 2] = fix[2];
            }
        }
    }
    if (lane == 0) bulk_wait_all<0>();
}

// Shared-memory rows of the generic kernel.  One thread reads one joint's row, so a row of r 16-byte units with r a multiple
// of 8 puts every lane of a warp on the same banks (double storage, V = 16: r = 24 -> 32-way conflicts; `short_scoreboard`
// 5 warps per issue, 39 % of the roofline).  Such rows are copied one by one (each thread issues the bulk copy of its own
// row; all of them complete on the stage's mbarrier) into slots of r + 1 units, which is conflict-free: +45 % for that
// kernel.  For r = 12 (double, V = 8: 8-way) the grouped variant (two rows per copy, one padding unit per group) removes
// the conflicts too, but the 128 small TMA copies per tile cost as much as they save, so it keeps the single contiguous
// bulk copy.
struct RowPlan {
    int group;        // rows per bulk copy (0: one contiguous copy for the whole tile)
    int slot_elems;   // elements per group slot (group * row_elems + padding)
};
__host__ __device__ __forceinline__ RowPlan tri_row_plan(int row_elems, int esize) {
    const int bytes = row_elems * esize;
    if (bytes % 16 != 0) return RowPlan{0, 0};
    const int r = bytes / 16;
    if (r % 8 != 0) return RowPlan{0, 0};
    return RowPlan{1, row_elems + 16 / esize};
}
__host__ __device__ __forceinline__ size_t tri_stage_elems(int tile, int row_elems, int esize) {
    const RowPlan p = tri_row_plan(row_elems, esize);
    return p.group ? (size_t)(tile / p.group) * p.slot_elems : (size_t)tile * row_elems;
}
// element offset of row `slot` of a tile inside its stage
__device__ __forceinline__ size_t tri_row_offset(const RowPlan &p, int slot, int row_elems) {
    return p.group ? (size_t)(slot / p.group) * p.slot_elems + (size_t)(slot % p.group) * row_elems : (size_t)slot * row_elems;
}

// ---- generic kernel -----------------------------------------------------------------------------------
// V > 0: number of views known at compile time (fully unrolled); V == 0: runtime prm.n_views.
template <typename T, int V, int MODE, bool UNDISTORT>
__global__ void __launch_bounds__(TRI_TILE, (V > 0 && V <= 8) ? 3 : 2)
triangulate_kernel(const T *__restrict__ kpts, T *__restrict__ out, long long n, int n_stages,
                   const __grid_constant__ TriParams prm) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int nv = (V > 0) ? V : prm.n_views;
    const int row_elems = 3 * nv;
    const RowPlan plan = tri_row_plan(row_elems, (int)sizeof(T));
    const uint32_t row_bytes = (uint32_t)(row_elems * sizeof(T));
    const uint32_t stage_bytes = (uint32_t)(tri_stage_elems(TRI_TILE, row_elems, (int)sizeof(T)) * sizeof(T));
    // layout: [n_stages][TILE*row_stride] input ring | [TILE*3] output tile | mbarriers | (TOP2) camera tables
    T *ring = reinterpret_cast<T *>(smem_raw);
    T *otile = reinterpret_cast<T *>(smem_raw + (size_t)n_stages * stage_bytes);
    uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw + (size_t)n_stages * stage_bytes + TRI_TILE * 3 * sizeof(T));
    double *cam = reinterpret_cast<double *>(full + 8);       // TOP2 only: [V][12+9+5]

    const int tid = threadIdx.x;
    const long long n_tiles = (n + TRI_TILE - 1) / TRI_TILE;
    const long long first = blockIdx.x;
    const long long stride = gridDim.x;
    const long long my_tiles = (first < n_tiles) ? (n_tiles - first + stride - 1) / stride : 0;

    if (tid == 0) {
        for (int s = 0; s < n_stages; ++s) mbar_init(&full[s], 1);
        fence_mbar_init();
    }
    if (MODE == MC3D_TRI_TOP2) {
        for (int i = tid; i < nv * 26; i += TRI_TILE) {
            const int v = i / 26, k = i % 26;
            cam[i] = (k < 12) ? prm.P[v][k] : (k < 21 ? prm.K[v][k - 12] : prm.dist[v][k - 21]);
        }
    }
    __syncthreads();

    auto tile_is_full = [&](long long tile) { return (tile + 1) * TRI_TILE <= n; };
    auto issue_load = [&](long long k, int s) {   // every thread; s = k % n_stages
        const long long tile = first + k * stride;
        if (k < my_tiles && tile_is_full(tile)) {
            unsigned char *dst = reinterpret_cast<unsigned char *>(ring) + (size_t)s * stage_bytes;
            const T *src = kpts + tile * TRI_TILE * (long long)row_elems;
            if (plan.group) {
                if (tid == 0) mbar_arrive_expect_tx(&full[s], TRI_TILE * row_bytes);
                if (tid % plan.group == 0)
                    bulk_g2s(dst + tri_row_offset(plan, tid, row_elems) * sizeof(T), src + (size_t)tid * row_elems,
                             plan.group * row_bytes, &full[s]);
            } else if (tid == 0) {
                mbar_arrive_expect_tx(&full[s], stage_bytes);
                bulk_g2s(dst, src, stage_bytes, &full[s]);
            }
        }
    };
    for (int k = 0; k < n_stages - 1; ++k) issue_load(k, k);

    int s = 0;                     // stage of iteration k
    int s_next = n_stages - 1;     // stage of iteration k + n_stages - 1
    uint32_t parity = 0;
    for (long long k = 0; k < my_tiles; ++k) {
        const long long tile = first + k * stride;
        const bool full_tile = tile_is_full(tile);
        T *stage = reinterpret_cast<T *>(reinterpret_cast<unsigned char *>(ring) + (size_t)s * stage_bytes);
        issue_load(k + n_stages - 1, s_next);       // refills the stage consumed in iteration k-1
        if (tid == 0) bulk_wait_read<0>();          // the output tile of iteration k-1 has left shared memory
        const long long joint = tile * TRI_TILE + tid;
        const bool active = joint < n;
        if (full_tile) {
            mbar_wait(&full[s], parity);
        } else {                               // ragged last tile: plain cooperative copy
            const long long base = tile * TRI_TILE * (long long)row_elems;
            const long long cnt = (n - tile * TRI_TILE) * row_elems;
            for (long long i = tid; i < cnt; i += TRI_TILE) stage[tri_row_offset(plan, (int)(i / row_elems), row_elems) + (i % row_elems)] = kpts[base + i];
            __syncthreads();
        }

        double X0 = NAN, X1 = NAN, X2 = NAN;
        const bool solved = false;
        double B[10];
#pragma unroll
        for (int i = 0; i < 10; ++i) B[i] = 0.0;
        int n_used = 0;
        if (active && !solved) {
            const T *row = stage + tri_row_offset(plan, tid, row_elems);
            const bool l3v = prm.layout == MC3D_LAYOUT_3V;
            if (MODE == MC3D_TRI_WEIGHTED && std::is_same<T, double>::value && V > 0 && V % 2 == 0 && !UNDISTORT) {
                // double storage, even view count: two views per three 128-bit loads (half the shared-memory
                // instructions, and a quarter of the bank-conflict wavefronts of 64-bit loads at these row strides)
                const double *rowd = reinterpret_cast<const double *>(row);
#pragma unroll
                for (int v = 0; v < ((V > 0) ? V : 2); v += 2) {
                    double x0, y0, w0, x1, y1, w1;
                    if (l3v) {
                        const double2 qx = *reinterpret_cast<const double2 *>(rowd + v);
                        const double2 qy = *reinterpret_cast<const double2 *>(rowd + V + v);
                        const double2 qw = *reinterpret_cast<const double2 *>(rowd + 2 * V + v);
                        x0 = qx.x; x1 = qx.y; y0 = qy.x; y1 = qy.y; w0 = qw.x; w1 = qw.y;
                    } else {
                        const double2 q0 = *reinterpret_cast<const double2 *>(rowd + 3 * v);
                        const double2 q1 = *reinterpret_cast<const double2 *>(rowd + 3 * v + 2);
                        const double2 q2 = *reinterpret_cast<const double2 *>(rowd + 3 * v + 4);
                        x0 = q0.x; y0 = q0.y; w0 = q1.x; x1 = q1.y; y1 = q2.x; w1 = q2.y;
                    }
                    n_used += (w0 != 0.0) + (w1 != 0.0);
                    accumulate_view(B, x0, y0, w0, prm.P[v]);
                    accumulate_view(B, x1, y1, w1, prm.P[v + 1]);
                }
            } else if (MODE == MC3D_TRI_WEIGHTED) {
#pragma unroll
                for (int v = 0; v < ((V > 0) ? V : MC3D_MAX_VIEWS); ++v) {
                    if (V == 0 && v >= nv) break;
                    double x = (double)(l3v ? row[v] : row[3 * v]);
                    double y = (double)(l3v ? row[nv + v] : row[3 * v + 1]);
                    const double w = (double)(l3v ? row[2 * nv + v] : row[3 * v + 2]);
                    if (UNDISTORT) undistort_px(x, y, prm.K[v], prm.dist[v]);
                    n_used += (w != 0.0);
                    accumulate_view(B, x, y, w, prm.P[v]);
                }
            } else {
                // two highest scores; np.argsort(conf)[-2:] semantics: ties -> higher index, NaN sorts last
                int i1 = -1, i0 = -1;
                double k1 = 0.0, k0 = 0.0;
                for (int v = 0; v < nv; ++v) {
                    double sc = (double)(l3v ? row[2 * nv + v] : row[3 * v + 2]);
                    if (sc != sc) sc = INFINITY;
                    if (i1 < 0 || sc >= k1) { i0 = i1; k0 = k1; i1 = v; k1 = sc; }
                    else if (i0 < 0 || sc >= k0) { i0 = v; k0 = sc; }
                }
                const int sel[2] = {i0, i1};
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    const int v = sel[q];
                    double x = (double)(l3v ? row[v] : row[3 * v]);
                    double y = (double)(l3v ? row[nv + v] : row[3 * v + 1]);
                    const double *cv = cam + v * 26;
                    if (UNDISTORT) undistort_px(x, y, cv + 12, cv + 21);
                    accumulate_view(B, x, y, 1.0, cv);
                }
                n_used = 2;
            }
        }
        __syncthreads();                       // [A] every thread has consumed stage s

        // non-finite pixels or weights poison the (non-negative) diagonal of B
        const bool finite_in = fabs((B[0] + B[2]) + (B[5] + B[9])) <= 1.0e300;
        if (active && !solved && finite_in && n_used >= 2) {
            bool ok = false;
            if (!(prm.flags & MC3D_TRI_FLAG_JACOBI)) ok = secular_newton(B, X0, X1, X2);
            if (!ok) {                             // cold path: only here does B go to local memory
                double Bl[10];
#pragma unroll
                for (int i = 0; i < 10; ++i) Bl[i] = B[i];
                jacobi4_smallest(Bl, X0, X1, X2);
            }
        }

        T *ot = otile;
        if (full_tile) {
            ot[tid * 3 + 0] = (T)X0;
            ot[tid * 3 + 1] = (T)X1;
            ot[tid * 3 + 2] = (T)X2;
            fence_proxy_async_smem();
            __syncthreads();                   // [B] tile complete and visible to the async proxy
            if (tid == 0) {
                bulk_s2g(out + tile * TRI_TILE * 3, ot, (uint32_t)(TRI_TILE * 3 * sizeof(T)));
                bulk_commit();
            }
        } else {
            if (active) {
                out[joint * 3 + 0] = (T)X0;
                out[joint * 3 + 1] = (T)X1;
                out[joint * 3 + 2] = (T)X2;
            }
            if (tid == 0) bulk_commit();       // keep one group per iteration for wait_group accounting
        }
        if (++s == n_stages) { s = 0; parity ^= 1u; }
        if (++s_next == n_stages) s_next = 0;
    }
    if (tid == 0) bulk_wait_all<0>();
}

// ---- host side --------------------------------------------------------------------------------
namespace {
struct ViewGeom {
    bool ok;             // finite camera: the left 3x3 block of P is invertible
    double H[9];         // its inverse
    double C[3];         // camera centre: P (C,1) = 0
    double axis[3];      // unit viewing direction (third row of the block, towards positive depth)
};

bool invert3(const double *m, double *inv, double &det) {
    const double a = m[0], b = m[1], c = m[2], d = m[3], e = m[4], f = m[5], g = m[6], h = m[7], i = m[8];
    det = a * (e * i - f * h) - b * (d * i - f * g) + c * (d * h - e * g);
    const double scale = fabs(a) + fabs(b) + fabs(c) + fabs(d) + fabs(e) + fabs(f) + fabs(g) + fabs(h) + fabs(i);
    if (!(fabs(det) > 1e-30 * scale * scale * scale) || !(fabs(det) <= 1.0e300)) return false;
    const double r = 1.0 / det;
    inv[0] = (e * i - f * h) * r; inv[1] = (c * h - b * i) * r; inv[2] = (b * f - c * e) * r;
    inv[3] = (f * g - d * i) * r; inv[4] = (a * i - c * g) * r; inv[5] = (c * d - a * f) * r;
    inv[6] = (d * h - e * g) * r; inv[7] = (b * g - a * h) * r; inv[8] = (a * e - b * d) * r;
    return true;
}

ViewGeom view_geometry(const double *P) {
    ViewGeom g;
    const double M[9] = {P[0], P[1], P[2], P[4], P[5], P[6], P[8], P[9], P[10]};
    double det = 0.0;
    g.ok = invert3(M, g.H, det);
    if (!g.ok) return g;
    for (int r = 0; r < 3; ++r) g.C[r] = -(g.H[3 * r] * P[3] + g.H[3 * r + 1] * P[7] + g.H[3 * r + 2] * P[11]);
    const double n = sqrt(P[8] * P[8] + P[9] * P[9] + P[10] * P[10]);
    const double sgn = det < 0 ? -1.0 : 1.0;
    for (int k = 0; k < 3; ++k) g.axis[k] = n > 0 ? sgn * P[8 + k] / n : 0.0;
    for (int k = 0; k < 3; ++k) g.ok = g.ok && fabs(g.C[k]) <= 1.0e300;
    return g;
}
}  // namespace

// Starting-point plan of the mixed-precision kernel: up to two pairs of views (A, B), chosen for the widest angle
// between their rays to a nominal scene point, with the constants of the closed-form two-view point (pair_start).
// Returns the number of pairs (0 when no two finite cameras exist: every joint then starts from the world origin).
// Only the speed of the kernel depends on the choice -- its fixed point does not.
int fill_start_pairs(const double (*P)[12], int n_views, mc3d_tri_start_pair *out) {
    ViewGeom geo[MC3D_MAX_VIEWS];
    int idx[MC3D_MAX_VIEWS], m = 0;
    for (int v = 0; v < n_views; ++v) {
        geo[v] = view_geometry(P[v]);
        if (geo[v].ok) idx[m++] = v;
    }
    if (m < 2) return 0;
    // nominal scene point: closest to every optical axis, pulled (weakly) towards a point in front of the rig so that
    // parallel axes do not leave it undetermined
    double Cm[3] = {0, 0, 0}, am[3] = {0, 0, 0};
    for (int i = 0; i < m; ++i)
        for (int k = 0; k < 3; ++k) { Cm[k] += geo[idx[i]].C[k] / m; am[k] += geo[idx[i]].axis[k] / m; }
    double L = 0.0;
    for (int i = 0; i < m; ++i)
        for (int k = 0; k < 3; ++k) L += (geo[idx[i]].C[k] - Cm[k]) * (geo[idx[i]].C[k] - Cm[k]) / m;
    L = sqrt(L);
    if (!(L > 0.0)) L = 1.0;
    const double an = sqrt(am[0] * am[0] + am[1] * am[1] + am[2] * am[2]);
    double Xreg[3];
    for (int k = 0; k < 3; ++k) Xreg[k] = Cm[k] + (an > 1e-6 ? am[k] / an : 0.0) * L;
    const double eps = 1e-3 * m;
    double A[9] = {eps, 0, 0, 0, eps, 0, 0, 0, eps}, b[3] = {eps * Xreg[0], eps * Xreg[1], eps * Xreg[2]};
    for (int i = 0; i < m; ++i) {
        const ViewGeom &g = geo[idx[i]];
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c) {
                const double q = (r == c ? 1.0 : 0.0) - g.axis[r] * g.axis[c];
                A[3 * r + c] += q;
                b[r] += q * g.C[c];
            }
    }
    double Ai[9], det = 0.0, X0[3] = {Xreg[0], Xreg[1], Xreg[2]};
    if (invert3(A, Ai, det)) {
        double cand[3];
        for (int r = 0; r < 3; ++r) cand[r] = Ai[3 * r] * b[0] + Ai[3 * r + 1] * b[1] + Ai[3 * r + 2] * b[2];
        bool in_front = true;               // diverging axes meet BEHIND the cameras: keep the point in front of the rig
        for (int i = 0; i < m; ++i) {
            const ViewGeom &g = geo[idx[i]];
            const double depth = g.axis[0] * (cand[0] - g.C[0]) + g.axis[1] * (cand[1] - g.C[1]) + g.axis[2] * (cand[2] - g.C[2]);
            in_front = in_front && depth > 0.0;
        }
        if (in_front) for (int k = 0; k < 3; ++k) X0[k] = cand[k];
    }
    auto score = [&](int a, int c) {        // |sin| of the angle between the rays from the two centres to X0
        double da[3], db[3];
        for (int k = 0; k < 3; ++k) { da[k] = X0[k] - geo[a].C[k]; db[k] = X0[k] - geo[c].C[k]; }
        const double cx = da[1] * db[2] - da[2] * db[1], cy = da[2] * db[0] - da[0] * db[2], cz = da[0] * db[1] - da[1] * db[0];
        const double na = sqrt(da[0] * da[0] + da[1] * da[1] + da[2] * da[2]), nb = sqrt(db[0] * db[0] + db[1] * db[1] + db[2] * db[2]);
        return (na > 0 && nb > 0) ? sqrt(cx * cx + cy * cy + cz * cz) / (na * nb) : 0.0;
    };
    // Up to MC3D_TRI_MAX_START pairs, widest angle first; a pair shares no view with the earlier ones while the rig allows it
    // (a joint takes the first pair whose two views it sees: with disjoint pairs one unusable view costs one pair, and only
    // joints that lose every pair need the second pass from the origin), then pairs that differ from the earlier ones.
    int pa[MC3D_TRI_MAX_START], pb[MC3D_TRI_MAX_START], np_ = 0;
    bool used[MC3D_MAX_VIEWS] = {false};
    for (int want = 0; want < MC3D_TRI_MAX_START; ++want) {
        int ba = -1, bb = -1;
        double best = -1.0;
        for (int pass = 0; pass < 2 && ba < 0; ++pass)
            for (int i = 0; i < m; ++i)
                for (int j = i + 1; j < m; ++j) {
                    const int a = idx[i], c = idx[j];
                    if (pass == 0 && (used[a] || used[c])) continue;
                    bool dup = false;
                    for (int q = 0; q < np_; ++q) dup = dup || (pa[q] == a && pb[q] == c) || (pa[q] == c && pb[q] == a);
                    if (dup) continue;
                    const double s = score(a, c);
                    if (s > best) { best = s; ba = a; bb = c; }
                }
        if (ba < 0) break;
        pa[np_] = ba; pb[np_] = bb; ++np_;
        used[ba] = used[bb] = true;
    }
    if (np_ == 1) { pa[1] = pb[0]; pb[1] = pa[0]; np_ = 2; }   // two cameras: the same pair with the roles exchanged
    for (int p = 0; p < np_; ++p) {
        const ViewGeom &ga = geo[pa[p]];
        const double *Pb = P[pb[p]];
        // epipolar direction in view B: the image of a step along the ray from C_A to X0
        double dir[3], len = 0.0;
        for (int k = 0; k < 3; ++k) { dir[k] = X0[k] - ga.C[k]; len += dir[k] * dir[k]; }
        len = sqrt(len);
        auto project_b = [&](const double *X, double &u, double &v) {
            const double h0 = Pb[0] * X[0] + Pb[1] * X[1] + Pb[2] * X[2] + Pb[3];
            const double h1 = Pb[4] * X[0] + Pb[5] * X[1] + Pb[6] * X[2] + Pb[7];
            const double h2 = Pb[8] * X[0] + Pb[9] * X[1] + Pb[10] * X[2] + Pb[11];
            u = h0 / h2; v = h1 / h2;
        };
        double X1[3], u0, v0, u1, v1;
        for (int k = 0; k < 3; ++k) X1[k] = X0[k] + 0.01 * dir[k];
        project_b(X0, u0, v0);
        project_b(X1, u1, v1);
        double alpha = u1 - u0, beta = v1 - v0;
        const double nl = sqrt(alpha * alpha + beta * beta);
        if (nl > 0.0 && nl <= 1.0e300 && len > 0.0) { alpha /= nl; beta /= nl; } else { alpha = 1.0; beta = 0.0; }
        double Pa[4];
        for (int k = 0; k < 4; ++k) Pa[k] = alpha * Pb[k] + beta * Pb[4 + k];
        mc3d_tri_start_pair &o = out[p];
        for (int k = 0; k < 9; ++k) o.H[k] = (float)ga.H[k];
        for (int k = 0; k < 3; ++k) {
            o.C[k] = (float)ga.C[k];
            o.ua[k] = (float)(ga.H[k] * Pa[0] + ga.H[3 + k] * Pa[1] + ga.H[6 + k] * Pa[2]);        // H^T Pa
            o.ub[k] = (float)(ga.H[k] * Pb[8] + ga.H[3 + k] * Pb[9] + ga.H[6 + k] * Pb[10]);       // H^T P2_B
        }
        o.alpha = (float)alpha;
        o.beta = (float)beta;
        o.ka = (float)(Pa[0] * ga.C[0] + Pa[1] * ga.C[1] + Pa[2] * ga.C[2] + Pa[3]);
        o.kb = (float)(Pb[8] * ga.C[0] + Pb[9] * ga.C[1] + Pb[10] * ga.C[2] + Pb[11]);
        o.view_a = pa[p];
        o.view_b = pb[p];
    }
    return np_;
}

static int fill_params(TriParams &prm, const mc3d_rig *rig, int layout, int mode, int flags) {
    if (!rig || !rig->P) { set_error("rig / rig->P is NULL"); return MC3D_ERR_INVALID_ARGUMENT; }
    if (rig->n_views < 2 || rig->n_views > MC3D_MAX_VIEWS) {
        set_error("n_views=%d outside [2, %d]", rig->n_views, MC3D_MAX_VIEWS);
        return MC3D_ERR_INVALID_ARGUMENT;
    }
    if (layout != MC3D_LAYOUT_V3 && layout != MC3D_LAYOUT_3V) { set_error("bad layout %d", layout); return MC3D_ERR_INVALID_ARGUMENT; }
    if (mode != MC3D_TRI_WEIGHTED && mode != MC3D_TRI_TOP2) { set_error("bad mode %d", mode); return MC3D_ERR_INVALID_ARGUMENT; }
    if ((rig->K == nullptr) != (rig->dist == nullptr)) {
        set_error("rig->K and rig->dist must both be given or both be NULL");
        return MC3D_ERR_INVALID_ARGUMENT;
    }
    memset(&prm, 0, sizeof(prm));
    prm.n_views = rig->n_views;
    prm.undistort = rig->K != nullptr;
    prm.flags = flags;
    prm.layout = layout;
    for (int v = 0; v < rig->n_views; ++v) {
        for (int k = 0; k < 12; ++k) prm.P[v][k] = rig->P[v * 12 + k];
        if (rig->K) {
            for (int k = 0; k < 9; ++k) prm.K[v][k] = rig->K[v * 9 + k];
            for (int k = 0; k < 5; ++k) prm.dist[v][k] = rig->dist[v * 5 + k];
        }
    }
    prm.n_start = fill_start_pairs(prm.P, rig->n_views, prm.start);
    double acc = 0.0;                         // rig scale: mean squared distance of the camera centres from the world origin
    int cnt = 0;
    for (int v = 0; v < rig->n_views; ++v) {
        const ViewGeom g = view_geometry(prm.P[v]);
        if (!g.ok) continue;
        const double n2 = g.C[0] * g.C[0] + g.C[1] * g.C[1] + g.C[2] * g.C[2];
        if (n2 <= 1.0e300) { acc += n2; ++cnt; }
    }
    prm.rig2 = cnt ? (float)fmin(acc / cnt, 1.0e30) : 0.f;
    return MC3D_OK;
}

// ---- warp-pipelined kernel for double storage -----------------------------------------------------------------------------
// The weighted, undistortion-free, compile-time-even-V case of the generic kernel above with the data path of the mixed
// kernel: every warp is its own pipeline (private two-stage ring of 32-joint tiles filled by 1-D TMA bulk copies, private
// output tile, TMA bulk store), one joint per lane, no block-wide barrier after the set-up and none of the generic kernel's
// per-tile bookkeeping (~170 of its ~1 030 instructions per joint).  Same accumulation order, same solver: results are
// bit-identical to the generic kernel's.  The ragged tail (< 32 joints) is one more tile of the next warp in line.
constexpr int TRI_W64_WARPS = 4;
constexpr int TRI_W64_TILE = 32;

// Rows whose size is a multiple of 128 bytes (16 views: 384 B) would put every lane of a warp on the same shared-memory
// banks: each lane then bulk-copies its OWN row into a slot padded by 16 bytes (conflict-free), all 32 copies completing on
// the warp's mbarrier, and -- the slots being 12.8 KB per tile -- the warp keeps ONE input stage and one output tile so that
// 16 warps stay resident.  Measured: V = 16 8.8e9 -> 1.26e10 points/s (80 % of the HBM roofline).  Padding the 192-byte rows of
// V = 8 the same way (4-way conflicts) costs more in small TMA copies than it saves: 2.18e10 -> 1.75e10, not done.
__host__ __device__ constexpr bool tri_w64_padded(int V) { return (3 * V * sizeof(double)) % 128 == 0; }
__host__ __device__ constexpr int tri_w64_slot_elems(int V) { return 3 * V + (tri_w64_padded(V) ? 2 : 0); }
__host__ __device__ constexpr int tri_w64_stages(int V) { return tri_w64_padded(V) ? 1 : 2; }
__host__ __device__ constexpr size_t tri_w64_smem(int V) {
    return (size_t)TRI_W64_WARPS * (tri_w64_stages(V) * (size_t)TRI_W64_TILE * tri_w64_slot_elems(V) * sizeof(double) +
                                    tri_w64_stages(V) * (size_t)TRI_W64_TILE * 3 * sizeof(double) + 2 * sizeof(uint64_t));
}

template <int V, int LAYOUT>
__global__ void __launch_bounds__(32 * TRI_W64_WARPS, 4)
triangulate_warp64_kernel(const double *__restrict__ kpts, double *__restrict__ out, long long n,
                          const __grid_constant__ TriParams prm) {
    static_assert(V > 0 && V % 2 == 0, "even compile-time view count");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int row_elems = 3 * V;
    constexpr bool PADDED = tri_w64_padded(V);
    constexpr int NST = tri_w64_stages(V);             // input stages and output tiles per warp
    constexpr int slot_elems = tri_w64_slot_elems(V);
    constexpr uint32_t row_bytes = (uint32_t)(row_elems * sizeof(double));
    constexpr uint32_t tile_bytes = (uint32_t)(TRI_W64_TILE * row_elems * sizeof(double));      // bytes a tile has in global memory
    constexpr uint32_t stage_bytes = (uint32_t)(TRI_W64_TILE * slot_elems * sizeof(double));     // ... and in shared memory
    constexpr uint32_t otile_bytes = (uint32_t)(TRI_W64_TILE * 3 * sizeof(double));
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // layout: [warp][NST] input stages | [warp][NST] output tiles | [warp][2] mbarriers
    unsigned char *ring = smem_raw + (size_t)warp * NST * stage_bytes;
    double *otile = reinterpret_cast<double *>(smem_raw + (size_t)TRI_W64_WARPS * NST * stage_bytes + (size_t)warp * NST * otile_bytes);
    uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw + (size_t)TRI_W64_WARPS * NST * (stage_bytes + otile_bytes)) + warp * 2;
    if (lane == 0) {
        mbar_init(&full[0], 1);
        mbar_init(&full[1], 1);
        fence_mbar_init();
    }
    __syncwarp();
    const unsigned n_tiles = (unsigned)(n / TRI_W64_TILE);
    const int tail = (int)(n - (long long)n_tiles * TRI_W64_TILE);
    const unsigned gw = blockIdx.x * TRI_W64_WARPS + warp, gstride = gridDim.x * TRI_W64_WARPS;
    const unsigned my_tiles = (gw < n_tiles) ? (n_tiles - gw + gstride - 1) / gstride : 0;
    const uint32_t ring_sa = smem_u32(ring), full_sa = smem_u32(full), ot_sa = smem_u32(otile);
    const unsigned char *src = reinterpret_cast<const unsigned char *>(kpts) + (size_t)gw * tile_bytes;
    unsigned char *dst = reinterpret_cast<unsigned char *>(out) + (size_t)gw * otile_bytes;
    const uint32_t src_step = gstride * tile_bytes, dst_step = gstride * otile_bytes;
    // one tile from global memory into stage `st`: one contiguous bulk copy, or one copy per lane into its padded slot
    auto load_tile = [&](const unsigned char *from, uint32_t st) {
        if (lane == 0) mbar_arrive_expect_tx_sa(full_sa + 8u * st, tile_bytes);
        if (PADDED) {
            __syncwarp();                             // the expected byte count is registered before any copy can complete
            bulk_g2s_sa(ring_sa + st * stage_bytes + lane * (uint32_t)(slot_elems * sizeof(double)), from + (size_t)lane * row_bytes,
                        row_bytes, full_sa + 8u * st);
        } else if (lane == 0) {
            bulk_g2s_sa(ring_sa + st * stage_bytes, from, tile_bytes, full_sa + 8u * st);
        }
    };
    if (my_tiles > 0) load_tile(src, 0);
    src += src_step;
    auto solve_row = [&](const double *rowd, double &X0, double &X1, double &X2) {
        double B[10];
#pragma unroll
        for (int i = 0; i < 10; ++i) B[i] = 0.0;
        int n_used = 0;
#pragma unroll
        for (int v = 0; v < V; v += 2) {
            double x0, y0, w0, x1, y1, w1;
            if (LAYOUT == MC3D_LAYOUT_3V) {
                const double2 qx = *reinterpret_cast<const double2 *>(rowd + v);
                const double2 qy = *reinterpret_cast<const double2 *>(rowd + V + v);
                const double2 qw = *reinterpret_cast<const double2 *>(rowd + 2 * V + v);
                x0 = qx.x; x1 = qx.y; y0 = qy.x; y1 = qy.y; w0 = qw.x; w1 = qw.y;
            } else {
                const double2 q0 = *reinterpret_cast<const double2 *>(rowd + 3 * v);
                const double2 q1 = *reinterpret_cast<const double2 *>(rowd + 3 * v + 2);
                const double2 q2 = *reinterpret_cast<const double2 *>(rowd + 3 * v + 4);
                x0 = q0.x; y0 = q0.y; w0 = q1.x; x1 = q1.y; y1 = q2.x; w1 = q2.y;
            }
            n_used += (w0 != 0.0) + (w1 != 0.0);
            accumulate_view(B, x0, y0, w0, prm.P[v]);
            accumulate_view(B, x1, y1, w1, prm.P[v + 1]);
        }
        X0 = X1 = X2 = NAN;
        const bool finite_in = fabs((B[0] + B[2]) + (B[5] + B[9])) <= 1.0e300;
        if (finite_in && n_used >= 2) {
            bool ok = false;
            if (!(prm.flags & MC3D_TRI_FLAG_JACOBI)) ok = secular_newton(B, X0, X1, X2);
            if (!ok) {
                double Bl[10];
#pragma unroll
                for (int i = 0; i < 10; ++i) Bl[i] = B[i];
                jacobi4_smallest(Bl, X0, X1, X2);
            }
        }
    };
    for (unsigned k = 0; k < my_tiles; ++k) {
        const uint32_t b = NST == 2 ? (k & 1u) : 0u;
        if (NST == 2) {                               // refills the stage of iteration k - 1
            if (k + 1 < my_tiles) load_tile(src, b ^ 1u);
            src += src_step;
        }
        const double *rowd = reinterpret_cast<const double *>(ring + b * stage_bytes) + lane * slot_elems;
        mbar_wait_sa(full_sa + 8u * b, NST == 2 ? ((k >> 1) & 1u) : (k & 1u));
        double X0, X1, X2;
        solve_row(rowd, X0, X1, X2);
        double *ot = otile + b * (TRI_W64_TILE * 3);
        if (lane == 0) {                              // the store that last used this output buffer has left it
            if (NST == 2) bulk_wait_read<1>(); else bulk_wait_read<0>();
        }
        __syncwarp();                                 // ... and every lane has consumed the stage
        if (NST == 1) {                               // one stage: the refill overlaps the epilogue, the other warps cover its latency
            if (k + 1 < my_tiles) load_tile(src, 0);
            src += src_step;
        }
        ot[lane * 3 + 0] = X0; ot[lane * 3 + 1] = X1; ot[lane * 3 + 2] = X2;
        fence_proxy_async_smem();
        __syncwarp();                                 // output tile complete
        if (lane == 0) {
            bulk_s2g_sa(dst, ot_sa + b * otile_bytes, otile_bytes);
            bulk_commit();
        }
        dst += dst_step;
    }
    if (tail > 0 && gw == n_tiles % gstride) {        // ragged tail: plain loads and stores
        double *stage = reinterpret_cast<double *>(ring);
        const double *tsrc = kpts + (size_t)n_tiles * TRI_W64_TILE * row_elems;
        for (int i = lane; i < tail * row_elems; i += 32) stage[(i / row_elems) * slot_elems + i % row_elems] = tsrc[i];
        __syncwarp();
        double X0, X1, X2;
        solve_row(stage + (lane < tail ? lane : 0) * slot_elems, X0, X1, X2);
        if (lane < tail) {
            double *tdst = out + ((size_t)n_tiles * TRI_W64_TILE + lane) * 3;
            tdst[0] = X0; tdst[1] = X1; tdst[2] = X2;
        }
    }
    if (lane == 0) bulk_wait_all<0>();
}

template <typename T, int V, int MODE, bool UNDISTORT>
static int launch_one(const T *d_kpts, long long n, const TriParams &prm, T *d_out, cudaStream_t stream) {
    const int nv = prm.n_views;
    const size_t stage_bytes = tri_stage_elems(TRI_TILE, 3 * nv, (int)sizeof(T)) * sizeof(T);
    auto kern = triangulate_kernel<T, V, MODE, UNDISTORT>;
    { const int as = func_max_smem_once((const void *)kern, 227 * 1024); if (as != MC3D_OK) return as; }
    // The kernel is bound by fp64 latency, not by bytes in flight: prefer more resident CTAs (warps) over a
    // deeper ring; 2 stages already cover the HBM latency at this arithmetic intensity.
    const size_t fixed = TRI_TILE * 3 * sizeof(T) + 8 * sizeof(uint64_t) +
                         (MODE == MC3D_TRI_TOP2 ? (size_t)nv * 26 * sizeof(double) : 0);
    int n_stages = 2, per_sm = 0;
    size_t smem = 0;
    for (int st = 4; st >= 2; --st) {
        const size_t sm_bytes = st * stage_bytes + fixed;
        if (sm_bytes > 227 * 1024) continue;
        int occ = 0;
        MC3D_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, TRI_TILE, sm_bytes));
        if (occ > per_sm) { per_sm = occ; n_stages = st; smem = sm_bytes; }
    }
    if (per_sm < 1) { set_error("triangulate kernel does not fit in shared memory (views=%d)", nv); return MC3D_ERR_UNSUPPORTED; }
    // weighted double storage with a compile-time even view count: the warp-pipelined kernel
    if constexpr (std::is_same<T, double>::value && MODE == MC3D_TRI_WEIGHTED && !UNDISTORT && V > 0 && V % 2 == 0) {
        if (n / TRI_W64_TILE < 0x7fffffffLL) {
            constexpr size_t smem64 = tri_w64_smem(V);
            static_assert(smem64 <= 227 * 1024, "warp64 kernel: shared memory");
            const long long warp_tiles = (n + TRI_W64_TILE - 1) / TRI_W64_TILE;
            const long long ctas = (warp_tiles + TRI_W64_WARPS - 1) / TRI_W64_WARPS;
            auto launch = [&](auto wk) -> int {
                { const int as = func_max_smem_once((const void *)wk, 227 * 1024); if (as != MC3D_OK) return as; }
                int occ = 0;
                MC3D_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, wk, 32 * TRI_W64_WARPS, smem64));
                if (occ < 1) { set_error("warp64 triangulate kernel does not fit on an SM"); return MC3D_ERR_UNSUPPORTED; }
                long long wgrid = (long long)sm_count() * occ;
                if (wgrid > ctas) wgrid = ctas;
                wk<<<(unsigned)wgrid, 32 * TRI_W64_WARPS, smem64, stream>>>(d_kpts, d_out, n, prm);
                count_launch();
                MC3D_CUDA_TRY(cudaGetLastError());
                return MC3D_OK;
            };
            if (prm.layout == MC3D_LAYOUT_3V) return launch(triangulate_warp64_kernel<V, MC3D_LAYOUT_3V>);
            return launch(triangulate_warp64_kernel<V, MC3D_LAYOUT_V3>);
        }
    }
    const long long n_tiles = (n + TRI_TILE - 1) / TRI_TILE;
    long long grid = (long long)sm_count() * per_sm;      // persistent: a whole number of waves
    if (grid > n_tiles) grid = n_tiles;
    kern<<<(unsigned)grid, TRI_TILE, smem, stream>>>(d_kpts, d_out, n, n_stages, prm);
    count_launch();
    MC3D_CUDA_TRY(cudaGetLastError());
    return MC3D_OK;
}

template <int V, int LAYOUT>
static int launch_mixed(const float *d_kpts, long long n, const TriParams &prm, float *d_out, cudaStream_t stream) {
    constexpr size_t stage_bytes = (size_t)TRI_WTILE * 3 * V * sizeof(float);
    constexpr size_t otile_bytes = (size_t)TRI_WTILE * 3 * sizeof(float);
    // two stages per warp (one for 16 views): a warp needs ~3 us per tile, which covers the HBM latency, and more resident
    // warps beat a deeper ring
    constexpr size_t smem = TRI_MWARPS * (tri_mixed_stages(V) * stage_bytes + 2 * otile_bytes + 2 * sizeof(uint64_t)) + V * sizeof(CamF) +
                            MC3D_TRI_MAX_START * sizeof(mc3d_tri_start_pair);
    static_assert(smem <= 227 * 1024, "mixed kernel: shared memory");
    auto kern = triangulate_mixed_kernel<V, LAYOUT>;
    { const int as = func_max_smem_once((const void *)kern, 227 * 1024); if (as != MC3D_OK) return as; }
    static int per_sm_cache[64] = {0};                                     // per device (this instantiation)
    int &per_sm = per_sm_cache[current_device_slot()];
    if (per_sm == 0) {
        int occ = 0;
        MC3D_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, TRI_MTHREADS, smem));
        if (occ < 1) { set_error("mixed triangulate kernel does not fit on an SM"); return MC3D_ERR_UNSUPPORTED; }
        per_sm = occ;
    }
    if (n / TRI_WTILE >= 0x7fffffffLL) { set_error("n=%lld: more than 2^31 tiles in one launch", n); return MC3D_ERR_UNSUPPORTED; }
    const long long warp_tiles = (n + TRI_WTILE - 1) / TRI_WTILE;          // the ragged tail is one more warp tile
    long long grid = (long long)sm_count() * per_sm;                       // persistent: a whole number of waves
    const long long ctas = (warp_tiles + TRI_MWARPS - 1) / TRI_MWARPS;
    if (grid > ctas) grid = ctas;
    kern<<<(unsigned)grid, TRI_MTHREADS, smem, stream>>>(d_kpts, d_out, n, prm);
    count_launch();
    MC3D_CUDA_TRY(cudaGetLastError());
    return MC3D_OK;
}

template <int V>
static int launch_mixed_layout(const float *d_kpts, long long n, const TriParams &prm, float *d_out, cudaStream_t stream) {
    if (prm.layout == MC3D_LAYOUT_3V) return launch_mixed<V, MC3D_LAYOUT_3V>(d_kpts, n, prm, d_out, stream);
    return launch_mixed<V, MC3D_LAYOUT_V3>(d_kpts, n, prm, d_out, stream);
}

template <typename T>
int triangulate_device(const T *d_kpts, long long n, const mc3d_rig *rig, int layout, int mode, int flags, T *d_out,
                       cudaStream_t stream) {
    TriParams prm;
    int st = fill_params(prm, rig, layout, mode, flags);
    if (st != MC3D_OK) return st;
    if (n < 0) { set_error("n=%lld < 0", n); return MC3D_ERR_INVALID_ARGUMENT; }
    if (n == 0) return MC3D_OK;
    if (!d_kpts || !d_out) { set_error("NULL device pointer"); return MC3D_ERR_INVALID_ARGUMENT; }
    if (!aligned16(d_kpts) || !aligned16(d_out)) {
        set_error("device pointers must be 16-byte aligned (kpts=%p out=%p)", (const void *)d_kpts, (void *)d_out);
        return MC3D_ERR_MISALIGNED;
    }
    if (mode == MC3D_TRI_TOP2) {
        if (prm.undistort) return launch_one<T, 0, MC3D_TRI_TOP2, true>(d_kpts, n, prm, d_out, stream);
        return launch_one<T, 0, MC3D_TRI_TOP2, false>(d_kpts, n, prm, d_out, stream);
    }
    if (prm.undistort) return launch_one<T, 0, MC3D_TRI_WEIGHTED, true>(d_kpts, n, prm, d_out, stream);
    // float storage, common rigs: mixed-precision kernel (float accumulation + double residuals)
    if constexpr (std::is_same<T, float>::value) {
        if (!(flags & (MC3D_TRI_FLAG_JACOBI | MC3D_TRI_FLAG_FP64))) {
            switch (prm.n_views) {
                case 2: return launch_mixed_layout<2>(d_kpts, n, prm, d_out, stream);
                case 3: return launch_mixed_layout<3>(d_kpts, n, prm, d_out, stream);
                case 4: return launch_mixed_layout<4>(d_kpts, n, prm, d_out, stream);
                case 8: return launch_mixed_layout<8>(d_kpts, n, prm, d_out, stream);
                case 16: return launch_mixed_layout<16>(d_kpts, n, prm, d_out, stream);
                default: break;
            }
        }
    }
    switch (prm.n_views) {      // fully unrolled view loops for the common rigs
        case 2: return launch_one<T, 2, MC3D_TRI_WEIGHTED, false>(d_kpts, n, prm, d_out, stream);
        case 3: return launch_one<T, 3, MC3D_TRI_WEIGHTED, false>(d_kpts, n, prm, d_out, stream);
        case 4: return launch_one<T, 4, MC3D_TRI_WEIGHTED, false>(d_kpts, n, prm, d_out, stream);
        case 8: return launch_one<T, 8, MC3D_TRI_WEIGHTED, false>(d_kpts, n, prm, d_out, stream);
        case 16: return launch_one<T, 16, MC3D_TRI_WEIGHTED, false>(d_kpts, n, prm, d_out, stream);
        default: return launch_one<T, 0, MC3D_TRI_WEIGHTED, false>(d_kpts, n, prm, d_out, stream);
    }
}

template int triangulate_device<float>(const float *, long long, const mc3d_rig *, int, int, int, float *, cudaStream_t);
template int triangulate_device<double>(const double *, long long, const mc3d_rig *, int, int, int, double *, cudaStream_t);

}  // namespace mc3d

extern "C" {

int mc3d_triangulate_start_plan(const mc3d_rig *rig, mc3d_tri_start_pair *pairs, int32_t *n_pairs) {
    if (!rig || !rig->P || !pairs || !n_pairs) { mc3d::set_error("mc3d_triangulate_start_plan: NULL argument"); return MC3D_ERR_INVALID_ARGUMENT; }
    if (rig->n_views < 2 || rig->n_views > MC3D_MAX_VIEWS) {
        mc3d::set_error("n_views=%d outside [2, %d]", rig->n_views, MC3D_MAX_VIEWS);
        return MC3D_ERR_INVALID_ARGUMENT;
    }
    double P[MC3D_MAX_VIEWS][12];
    for (int v = 0; v < rig->n_views; ++v)
        for (int k = 0; k < 12; ++k) P[v][k] = rig->P[v * 12 + k];
    memset(pairs, 0, MC3D_TRI_MAX_START * sizeof(mc3d_tri_start_pair));
    *n_pairs = mc3d::fill_start_pairs(P, rig->n_views, pairs);
    return MC3D_OK;
}

int mc3d_triangulate_f32(const float *d_kpts, int64_t n, const mc3d_rig *rig, int layout, int mode, int flags,
                         float *d_out, void *stream) {
    return mc3d::triangulate_device<float>(d_kpts, n, rig, layout, mode, flags, d_out, (cudaStream_t)stream);
}

int mc3d_triangulate_f64(const double *d_kpts, int64_t n, const mc3d_rig *rig, int layout, int mode, int flags,
                         double *d_out, void *stream) {
    return mc3d::triangulate_device<double>(d_kpts, n, rig, layout, mode, flags, d_out, (cudaStream_t)stream);
}

}  // extern "C"
