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
    // mixed-precision path (float storage): rows re-centred on each view's principal point
    //   a = (y - cy) P2 - P1',  c = P0' - (x - cx) P2,   P0' = P0 - cx P2,  P1' = P1 - cy P2   (same rows, smaller numbers)
    float2 p2[MC3D_MAX_VIEWS][4];     // (P2_k, P2_k)
    float2 p10[MC3D_MAX_VIEWS][4];    // (-P1'_k, P0'_k)
    float2 cxy[MC3D_MAX_VIEWS];       // (cx, cy) as floats
    double cxyd[MC3D_MAX_VIEWS][2];   // the same values as doubles
    double Pc[MC3D_MAX_VIEWS][12];    // P0', P1', P2 in double
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

// ---- mixed-precision solver (float storage) ---------------------------------------------------------------------
// float LDL^T of a 3x3 SPD system (ldl3_factor_f / ldl3_apply_f below); approximate reciprocals are fine: the solve
// only preconditions an iteration whose fixed point is set by residuals evaluated in double.
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

// ---- merged-pass mixed-precision solver (float storage) -----------------------------------------------------------
// The two-pass solver above touches every view twice (float normal equations, then double residuals).  This one
// touches the full view set once:
//   E. a weighted float DLT over a fixed subset of (at most) three well-spread views gives a starting point Xp that
//      is a few millimetres from the minimiser (the subset's own noise); a subset without two usable views starts
//      from the world origin instead and takes a second pass;
//   M. ONE pass over all V views accumulates, per view, the float normal matrix M~ = sum w^2 (a a^T + c c^T) and the
//      gradient g = A_m^T A (Xp,1) with the residual A (Xp,1) evaluated in DOUBLE (the cancellation
//      y (P2.X) - P1.X needs it) and rounded to float;
//   S. the correction solves the eigen-equations to first order in lam/M:
//         e1 = -M~^-1 g,   lam = (|r|^2 + g.e1) / (1 + |Xp + e1|^2),   e = e1 + lam M~^-1 (Xp + e1)
//      (exact fixed point (M - lam) X = -b; the neglected term is (lam/mu_min)^2 |X|, checked per joint).
// The float solve is accurate to ~3e-7 |e|, so a correction of up to a few centimetres lands within 1e-5 mm;
// larger corrections take another pass from the updated point, and whatever cannot be handled (failed pivots,
// non-finite input, lam not small, no convergence) goes to the all-double solver.
struct __align__(16) CamC {
    double Pc[12];       // P0', P1', P2 (rows re-centred on the principal point)
    double cxd, cyd;
    float2 p10[4];       // (-P1'_k, P0'_k)
    float4 p2;           // P2_k
    float cx, cy, pad0, pad1;
};
static_assert(sizeof(CamC) == 176, "CamC layout");

struct Ldl3f {
    float i0, l10, l20, i1, l21, i2;
};

__device__ __forceinline__ bool ldl3_factor_f(float m00, float m10, float m11, float m20, float m21, float m22, Ldl3f &f) {
    f.i0 = rcp_fast(m00);
    f.l10 = m10 * f.i0;
    f.l20 = m20 * f.i0;
    const float d1 = fmaf(-f.l10, m10, m11);
    f.i1 = rcp_fast(d1);
    const float t21 = fmaf(-f.l20, m10, m21);
    f.l21 = t21 * f.i1;
    const float d2 = fmaf(-f.l21, t21, fmaf(-f.l20, m20, m22));
    f.i2 = rcp_fast(d2);
    return (m00 > 0.f) & (d1 > 1e-6f * m11) & (d2 > 1e-6f * m22);
}

__device__ __forceinline__ void ldl3_apply_f(const Ldl3f &f, float r0, float r1, float r2, float &z0, float &z1, float &z2) {
    const float y1 = fmaf(-f.l10, r0, r1);
    const float y2 = fmaf(-f.l21, y1, fmaf(-f.l20, r0, r2));
    z2 = y2 * f.i2;
    z1 = fmaf(y1, f.i1, -f.l21 * z2);
    z0 = fmaf(r0, f.i0, -fmaf(f.l10, z1, f.l20 * z2));
}

// ---- the same 3x3 solve for TWO joints at once (MC3D_TRI_PACKED_SOLVE) -----------------------------------------------
// Every float2 holds (joint 0, joint 1) of one thread.  Operation by operation the arithmetic is that of ldl3_factor_f /
// ldl3_apply_f (one FFMA2 / FMUL2 / FADD2 = two independent IEEE operations), so results are bit-identical to the scalar
// code; what changes is the issue-slot count of the two solve phases (scalar: ~130 float instructions per joint).
#ifndef MC3D_TRI_PACKED_SOLVE
#define MC3D_TRI_PACKED_SOLVE 1
#endif
struct Ldl3f2 {
    float2 i0, l10, l20, i1, l21, i2;
};
__device__ __forceinline__ float2 neg2(float2 a) { return make_float2(-a.x, -a.y); }
__device__ __forceinline__ float2 rcp_fast2(float2 a) { return make_float2(rcp_fast(a.x), rcp_fast(a.y)); }
__device__ __forceinline__ float2 dot3_2(float2 a0, float2 a1, float2 a2, float2 b0, float2 b1, float2 b2) {
    return __ffma2_rn(a0, b0, __ffma2_rn(a1, b1, __fmul2_rn(a2, b2)));       // fmaf(a0, b0, fmaf(a1, b1, a2 * b2))
}

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
    const float2 lim = make_float2(1e-6f, 1e-6f);
    const float2 t1 = __fmul2_rn(lim, m11), t2 = __fmul2_rn(lim, m22);
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

// views of the starting-point subset: (i * V) / NE for i < NE, NE = min(V, 3)
__host__ __device__ constexpr int est_count(int V) { return V < 3 ? V : 3; }
__host__ __device__ constexpr bool is_est_view(int V, int v) {
    for (int i = 0; i < est_count(V); ++i)
        if ((i * V) / est_count(V) == v) return true;
    return false;
}
__host__ __device__ constexpr bool group_has_est_view(int V, int vg, int G) {
    for (int i = 0; i < G; ++i)
        if (is_est_view(V, vg + i)) return true;
    return false;
}

#ifndef MC3D_TRI_ROWS_E
#define MC3D_TRI_ROWS_E 1
#endif
// tuning build: the double residuals use the projection rows as given (P0, P1, P2) instead of the rows re-centred on
// the principal point -- the same residuals in exact arithmetic (re-centring only matters for the FLOAT rows), two
// DADD and one shared-memory load fewer per joint-view; equal to the shipped kernel up to float-rounding ties
#ifndef MC3D_TRI_RAW_RESID
#define MC3D_TRI_RAW_RESID 1
#endif
// weighted float rows of one view, packed (a_k, c_k), k < NK
template <int NK>
__device__ __forceinline__ void float_rows(float x, float y, float w, float cx, float cy, const float (&p2)[4],
                                           const float2 (&p10)[4], float2 (&ac)[NK]) {
#if MC3D_TRI_ROWS_E
    // tuning build: w (yx P2 + p10) as the starting-point phase forms its rows -- two instructions fewer per view than
    // (w yx) P2 + w p10, different rounding (not bit-identical to the shipped kernel; the fixed point is unchanged)
    const float2 yx = make_float2(y - cy, cx - x);
    const float2 ww = make_float2(w, w);
#pragma unroll
    for (int k = 0; k < NK; ++k) ac[k] = __fmul2_rn(ww, __ffma2_rn(yx, make_float2(p2[k], p2[k]), p10[k]));
#else
    const float xc = x - cx, yc = y - cy;
    const float2 sv = make_float2(w * yc, -(w * xc));
    const float2 ww = make_float2(w, w);
#pragma unroll
    for (int k = 0; k < NK; ++k) ac[k] = __ffma2_rn(sv, make_float2(p2[k], p2[k]), __fmul2_rn(ww, p10[k]));
#endif
}

#ifndef MC3D_TRI_UNR_LIMIT
#define MC3D_TRI_UNR_LIMIT 4            // full unrolling of 8+ views hoists loads into spills
#endif
#ifndef MC3D_TRI_ACCEPT
#define MC3D_TRI_ACCEPT 1.0e-3f      // accept a correction with |e|^2 <= this * (|X|^2 + rig scale^2): |e| <~ 3 % of the range
#endif

// tuning build (with MC3D_TRI_PACKED_SOLVE, two joints per thread): the accepted result leaves the solver as FLOATS.  In the
// first pass the starting point is float-valued, so the result float(Xp + e) is one float addition; the conversions
// float(Xp), double(e), float(X) and the double additions of the shipped tail (9 F2F + 3 DADD per joint) are then only
// executed by joints that need another pass.  Same value as the shipped kernel up to double-rounding ties.
#ifndef MC3D_TRI_FLOAT_TAIL
#define MC3D_TRI_FLOAT_TAIL 1
#endif
#if MC3D_TRI_FLOAT_TAIL && !MC3D_TRI_PACKED_SOLVE
#error "MC3D_TRI_FLOAT_TAIL needs MC3D_TRI_PACKED_SOLVE"
#endif
#if MC3D_TRI_FLOAT_TAIL
#define MC3D_XO_PARAM , float (&Xo)[NJ][3]
#define MC3D_XO_ARG , Xo
#else
#define MC3D_XO_PARAM
#define MC3D_XO_ARG
#endif

template <int V, int LAYOUT, int NJ>
__device__ __forceinline__ void solve_merged(const CamC *__restrict__ cam, float rig2, const float *const (&rows)[NJ],
                                             const bool (&act)[NJ], double (&X)[NJ][3], int (&state)[NJ] MC3D_XO_PARAM) {
#if MC3D_TRI_FLOAT_TAIL
    static_assert(NJ == 2, "MC3D_TRI_FLOAT_TAIL is written for two joints per thread");
    float Xf[NJ][3];                                                // float(Xd): exact in the first pass
#endif
    constexpr int G = (V % 4 == 0) ? 4 : V;                     // views per load group
    double Xd[NJ][3];
    // ---- E: starting point from the subset ---------------------------------------------------------------
    {
        float2 aM[NJ][6], ab[NJ][3];
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
#pragma unroll
            for (int i = 0; i < 6; ++i) aM[j][i] = make_float2(0.f, 0.f);
#pragma unroll
            for (int i = 0; i < 3; ++i) ab[j][i] = make_float2(0.f, 0.f);
        }
#pragma unroll
        for (int vg = 0; vg < V; vg += G) {
            if (!group_has_est_view(V, vg, G)) continue;
            float gx[NJ][G], gy[NJ][G], gw[NJ][G];
#pragma unroll
            for (int j = 0; j < NJ; ++j) load_group<V, LAYOUT, G>(rows[j], vg, gx[j], gy[j], gw[j]);
#pragma unroll
            for (int i = 0; i < G; ++i) {
                const int v = vg + i;
                if (!is_est_view(V, v)) continue;
                const CamC &c = cam[v];
                const float4 p2v = c.p2;
                const float p2[4] = {p2v.x, p2v.y, p2v.z, p2v.w};
                float2 p10[4];
                {
                    const float4 a = *reinterpret_cast<const float4 *>(&c.p10[0]);
                    const float4 b = *reinterpret_cast<const float4 *>(&c.p10[2]);
                    p10[0] = make_float2(a.x, a.y); p10[1] = make_float2(a.z, a.w);
                    p10[2] = make_float2(b.x, b.y); p10[3] = make_float2(b.z, b.w);
                }
                const float2 cxy = *reinterpret_cast<const float2 *>(&c.cx);
#pragma unroll
                for (int j = 0; j < NJ; ++j) {
                    // rows w (y P2 - P1'), w (P0' - x P2) of this view: a zero-weight view drops out of the start and a
                    // low-confidence one barely moves it, so unusable detections among the starting views do not cost
                    // the warp a second pass (measured at 1 % unusable views: 2.8e10 joints/s against 1.9e10 with
                    // unweighted rows, which are 3 % faster on clean input)
                    float2 ac[4];
                    const float2 yx = make_float2(gy[j][i] - cxy.y, cxy.x - gx[j][i]);
                    const float2 ww = make_float2(gw[j][i], gw[j][i]);
#pragma unroll
                    for (int k = 0; k < 4; ++k) ac[k] = __fmul2_rn(ww, __ffma2_rn(yx, make_float2(p2[k], p2[k]), p10[k]));
                    aM[j][0] = __ffma2_rn(ac[0], ac[0], aM[j][0]);
                    aM[j][1] = __ffma2_rn(ac[1], ac[0], aM[j][1]);
                    aM[j][2] = __ffma2_rn(ac[1], ac[1], aM[j][2]);
                    aM[j][3] = __ffma2_rn(ac[2], ac[0], aM[j][3]);
                    aM[j][4] = __ffma2_rn(ac[2], ac[1], aM[j][4]);
                    aM[j][5] = __ffma2_rn(ac[2], ac[2], aM[j][5]);
                    ab[j][0] = __ffma2_rn(ac[3], ac[0], ab[j][0]);
                    ab[j][1] = __ffma2_rn(ac[3], ac[1], ab[j][1]);
                    ab[j][2] = __ffma2_rn(ac[3], ac[2], ab[j][2]);
                }
            }
        }
#if MC3D_TRI_PACKED_SOLVE
        if constexpr (NJ == 2) {
            Ldl3f2 f;
            bool ok[2];
            ldl3_factor_f2(MC3D_HSUM2(aM, 0), MC3D_HSUM2(aM, 1), MC3D_HSUM2(aM, 2), MC3D_HSUM2(aM, 3), MC3D_HSUM2(aM, 4),
                           MC3D_HSUM2(aM, 5), f, ok[0], ok[1]);
            float2 z0, z1, z2;
            ldl3_apply_f2(f, neg2(MC3D_HSUM2(ab, 0)), neg2(MC3D_HSUM2(ab, 1)), neg2(MC3D_HSUM2(ab, 2)), z0, z1, z2);
            const float2 nz2 = dot3_2(z0, z1, z2, z0, z1, z2);
            const float2 tr2 = __fadd2_rn(__fadd2_rn(MC3D_HSUM2(aM, 0), MC3D_HSUM2(aM, 2)), MC3D_HSUM2(aM, 5));
            const float nz[2] = {nz2.x, nz2.y}, tr[2] = {tr2.x, tr2.y};
            const float zs[2][3] = {{z0.x, z1.x, z2.x}, {z0.y, z1.y, z2.y}};
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                X[j][0] = X[j][1] = X[j][2] = NAN;
                Xd[j][0] = Xd[j][1] = Xd[j][2] = 0.0;
                state[j] = 0;
                if (!act[j]) continue;
                if (!(tr[j] <= 3.0e38f)) { state[j] = 1; continue; }
                state[j] = 2;
                if (ok[j] && nz[j] <= 3.0e38f) { Xd[j][0] = (double)zs[j][0]; Xd[j][1] = (double)zs[j][1]; Xd[j][2] = (double)zs[j][2]; }
            }
#if MC3D_TRI_FLOAT_TAIL
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const bool started = act[j] && tr[j] <= 3.0e38f && ok[j] && nz[j] <= 3.0e38f;
#pragma unroll
                for (int k = 0; k < 3; ++k) { Xf[j][k] = started ? zs[j][k] : 0.f; Xo[j][k] = NAN; }
            }
#endif
        } else
#endif
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            X[j][0] = X[j][1] = X[j][2] = NAN;
            Xd[j][0] = Xd[j][1] = Xd[j][2] = 0.0;
            state[j] = 0;
            if (!act[j]) continue;
            Ldl3f f;
            const bool ok = ldl3_factor_f(aM[j][0].x + aM[j][0].y, aM[j][1].x + aM[j][1].y, aM[j][2].x + aM[j][2].y,
                                          aM[j][3].x + aM[j][3].y, aM[j][4].x + aM[j][4].y, aM[j][5].x + aM[j][5].y, f);
            float z0, z1, z2;
            ldl3_apply_f(f, -(ab[j][0].x + ab[j][0].y), -(ab[j][1].x + ab[j][1].y), -(ab[j][2].x + ab[j][2].y), z0, z1, z2);
            const float nz = fmaf(z0, z0, fmaf(z1, z1, z2 * z2));
            const float tr = (aM[j][0].x + aM[j][0].y) + (aM[j][2].x + aM[j][2].y) + (aM[j][5].x + aM[j][5].y);
            if (!(tr <= 3.0e38f)) { state[j] = 1; continue; }         // non-finite input: the all-double path classifies it
            state[j] = 2;
            // fewer than two usable views in the subset (zero weights): start from the world origin -- the first pass
            // is then a float solve over all views, the second one finishes
            if (ok && nz <= 3.0e38f) { Xd[j][0] = (double)z0; Xd[j][1] = (double)z1; Xd[j][2] = (double)z2; }
        }
    }
    // ---- M + S: merged pass from Xp, repeated only when the correction was large ---------------------------------
    constexpr int UNR = (V > MC3D_TRI_UNR_LIMIT) ? 1 : (V / G);
#pragma unroll 1
    for (int it = 0; it < 6; ++it) {
        bool any = false;
#pragma unroll
        for (int j = 0; j < NJ; ++j) any = any || (state[j] == 2);
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
                const CamC &c = cam[vg + i];
                double Pc[12];
#pragma unroll
                for (int k = 0; k < 6; ++k) {
                    const double2 t = *reinterpret_cast<const double2 *>(&c.Pc[2 * k]);
                    Pc[2 * k] = t.x; Pc[2 * k + 1] = t.y;
                }
#if !MC3D_TRI_RAW_RESID
                const double2 cxyd = *reinterpret_cast<const double2 *>(&c.cxd);
#endif
                const float4 p2v = c.p2;
                const float p2[4] = {p2v.x, p2v.y, p2v.z, p2v.w};
                float2 p10[4];
                {
                    const float4 a = *reinterpret_cast<const float4 *>(&c.p10[0]);
                    const float2 b = c.p10[2];
                    p10[0] = make_float2(a.x, a.y); p10[1] = make_float2(a.z, a.w); p10[2] = b;
                    p10[3] = make_float2(0.f, 0.f);
                }
                const float2 cxy = *reinterpret_cast<const float2 *>(&c.cx);
#pragma unroll
                for (int j = 0; j < NJ; ++j) {
                    const float x = gx[j][i], y = gy[j][i], w = gw[j][i];
                    float2 ac[3];
                    float_rows<3>(x, y, w, cxy.x, cxy.y, p2, p10, ac);
                    aM[j][0] = __ffma2_rn(ac[0], ac[0], aM[j][0]);
                    aM[j][1] = __ffma2_rn(ac[1], ac[0], aM[j][1]);
                    aM[j][2] = __ffma2_rn(ac[1], ac[1], aM[j][2]);
                    aM[j][3] = __ffma2_rn(ac[2], ac[0], aM[j][3]);
                    aM[j][4] = __ffma2_rn(ac[2], ac[1], aM[j][4]);
                    aM[j][5] = __ffma2_rn(ac[2], ac[2], aM[j][5]);
                    const double d0 = fma(Pc[0], Xd[j][0], fma(Pc[1], Xd[j][1], fma(Pc[2], Xd[j][2], Pc[3])));
                    const double d1 = fma(Pc[4], Xd[j][0], fma(Pc[5], Xd[j][1], fma(Pc[6], Xd[j][2], Pc[7])));
                    const double d2 = fma(Pc[8], Xd[j][0], fma(Pc[9], Xd[j][1], fma(Pc[10], Xd[j][2], Pc[11])));
#if MC3D_TRI_RAW_RESID
                    const double xcd = (double)x, ycd = (double)y;                      // Pc holds P0, P1, P2 themselves
#else
                    const double xcd = (double)x - cxyd.x, ycd = (double)y - cxyd.y;
#endif
                    const float r1 = (float)fma(ycd, d2, -d1);   // y (P2.X) - P1.X : the cancellation is in double
                    const float r2 = (float)fma(-xcd, d2, d0);   // P0.X - x (P2.X)
                    const float2 t = __fmul2_rn(make_float2(w, w), make_float2(r1, r2));
#pragma unroll
                    for (int k = 0; k < 3; ++k) ag[j][k] = __ffma2_rn(t, ac[k], ag[j][k]);
                    arr[j] = __ffma2_rn(t, t, arr[j]);
                }
            }
        }
#if MC3D_TRI_PACKED_SOLVE
        if constexpr (NJ == 2) {
            Ldl3f2 f;
            bool ok[2];
            ldl3_factor_f2(MC3D_HSUM2(aM, 0), MC3D_HSUM2(aM, 1), MC3D_HSUM2(aM, 2), MC3D_HSUM2(aM, 3), MC3D_HSUM2(aM, 4),
                           MC3D_HSUM2(aM, 5), f, ok[0], ok[1]);
            const float2 g0 = MC3D_HSUM2(ag, 0), g1 = MC3D_HSUM2(ag, 1), g2 = MC3D_HSUM2(ag, 2);
            float2 e0, e1, e2;
            ldl3_apply_f2(f, neg2(g0), neg2(g1), neg2(g2), e0, e1, e2);
            const float2 rr = __fadd2_rn(make_float2(arr[0].x + arr[0].y, arr[1].x + arr[1].y), dot3_2(g0, g1, g2, e0, e1, e2));
#if MC3D_TRI_FLOAT_TAIL
            const float2 x0 = __fadd2_rn(make_float2(Xf[0][0], Xf[1][0]), e0);
            const float2 x1 = __fadd2_rn(make_float2(Xf[0][1], Xf[1][1]), e1);
            const float2 x2 = __fadd2_rn(make_float2(Xf[0][2], Xf[1][2]), e2);
#else
            const float2 x0 = __fadd2_rn(make_float2((float)Xd[0][0], (float)Xd[1][0]), e0);
            const float2 x1 = __fadd2_rn(make_float2((float)Xd[0][1], (float)Xd[1][1]), e1);
            const float2 x2 = __fadd2_rn(make_float2((float)Xd[0][2], (float)Xd[1][2]), e2);
#endif
            const float2 nx2 = dot3_2(x0, x1, x2, x0, x1, x2);
            const float2 lam2 = __fmul2_rn(rr, rcp_fast2(__fadd2_rn(make_float2(1.f, 1.f), nx2)));
            float2 h0, h1, h2;
            ldl3_apply_f2(f, __fmul2_rn(lam2, x0), __fmul2_rn(lam2, x1), __fmul2_rn(lam2, x2), h0, h1, h2);
            e0 = __fadd2_rn(e0, h0); e1 = __fadd2_rn(e1, h1); e2 = __fadd2_rn(e2, h2);
            const float2 ne2 = dot3_2(e0, e1, e2, e0, e1, e2);
            const float2 lim2 = __fmul2_rn(make_float2(MC3D_TRI_ACCEPT, MC3D_TRI_ACCEPT), __fadd2_rn(nx2, make_float2(rig2, rig2)));
            const float ne[2] = {ne2.x, ne2.y}, lam[2] = {lam2.x, lam2.y}, lim[2] = {lim2.x, lim2.y};
            const float imax[2] = {fmaxf(f.i0.x, fmaxf(f.i1.x, f.i2.x)), fmaxf(f.i0.y, fmaxf(f.i1.y, f.i2.y))};
            const float es[2][3] = {{e0.x, e1.x, e2.x}, {e0.y, e1.y, e2.y}};
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                if (state[j] != 2) continue;
                if (!ok[j] || !(ne[j] <= 3.0e38f) || !(fabsf(lam[j]) * imax[j] <= 3.0e-5f)) { state[j] = 1; continue; }
#if MC3D_TRI_FLOAT_TAIL
                if (ne[j] <= lim[j]) {
                    if (it == 0) {                                   // Xd is float-valued: float(Xd + e) is one float addition
                        Xo[j][0] = Xf[j][0] + es[j][0]; Xo[j][1] = Xf[j][1] + es[j][1]; Xo[j][2] = Xf[j][2] + es[j][2];
                    } else {
                        Xo[j][0] = (float)(Xd[j][0] + (double)es[j][0]); Xo[j][1] = (float)(Xd[j][1] + (double)es[j][1]);
                        Xo[j][2] = (float)(Xd[j][2] + (double)es[j][2]);
                    }
                    state[j] = 0;
                } else {                                             // another pass from the updated point
                    Xd[j][0] += (double)es[j][0]; Xd[j][1] += (double)es[j][1]; Xd[j][2] += (double)es[j][2];
                    Xf[j][0] = (float)Xd[j][0]; Xf[j][1] = (float)Xd[j][1]; Xf[j][2] = (float)Xd[j][2];
                }
#else
                Xd[j][0] += (double)es[j][0]; Xd[j][1] += (double)es[j][1]; Xd[j][2] += (double)es[j][2];
                if (ne[j] <= lim[j]) {
                    X[j][0] = Xd[j][0]; X[j][1] = Xd[j][1]; X[j][2] = Xd[j][2];
                    state[j] = 0;
                }
#endif
            }
        } else
#endif
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            if (state[j] != 2) continue;
            Ldl3f f;
            const bool ok = ldl3_factor_f(aM[j][0].x + aM[j][0].y, aM[j][1].x + aM[j][1].y, aM[j][2].x + aM[j][2].y,
                                          aM[j][3].x + aM[j][3].y, aM[j][4].x + aM[j][4].y, aM[j][5].x + aM[j][5].y, f);
            const float g0 = ag[j][0].x + ag[j][0].y, g1 = ag[j][1].x + ag[j][1].y, g2 = ag[j][2].x + ag[j][2].y;
            float e0, e1, e2;
            ldl3_apply_f(f, -g0, -g1, -g2, e0, e1, e2);
            const float rr = (arr[j].x + arr[j].y) + fmaf(g0, e0, fmaf(g1, e1, g2 * e2));
            const float x0 = (float)Xd[j][0] + e0, x1 = (float)Xd[j][1] + e1, x2 = (float)Xd[j][2] + e2;
            const float nx = fmaf(x0, x0, fmaf(x1, x1, x2 * x2));
            const float lam = rr * rcp_fast(1.f + nx);
            float h0, h1, h2;
            ldl3_apply_f(f, lam * x0, lam * x1, lam * x2, h0, h1, h2);
            e0 += h0; e1 += h1; e2 += h2;
            const float ne = fmaf(e0, e0, fmaf(e1, e1, e2 * e2));
            // first-order treatment of lam needs lam << smallest pivot of M~
            const float imax = fmaxf(f.i0, fmaxf(f.i1, f.i2));
            if (!ok || !(ne <= 3.0e38f) || !(fabsf(lam) * imax <= 3.0e-5f)) { state[j] = 1; continue; }
            Xd[j][0] += (double)e0; Xd[j][1] += (double)e1; Xd[j][2] += (double)e2;
            if (ne <= MC3D_TRI_ACCEPT * (nx + rig2)) {
                X[j][0] = Xd[j][0]; X[j][1] = Xd[j][1]; X[j][2] = Xd[j][2];
                state[j] = 0;
            }
        }
    }
#pragma unroll
    for (int j = 0; j < NJ; ++j)
        if (state[j] == 2) state[j] = 1;                         // did not converge in 6 passes
}

// Shared-memory rows.  One thread reads one joint's row, so a row of r 16-byte units with r a multiple of 8 puts every
// lane of a warp on the same banks (double storage, V = 16: r = 24 -> 32-way conflicts; `short_scoreboard` 5 warps per
// issue, 39 % of the roofline).  Such rows are copied one by one (each thread issues the bulk copy of its own row;
// all of them complete on the stage's mbarrier) into slots of r + 1 units, which is conflict-free: +45 % for that
// kernel.  For r = 12 (double, V = 8: 8-way) and r = 6 (float, V = 8: 2-way) the grouped variant (G = 8 / gcd(r, 8)
// rows per copy, one padding unit per group) removes the conflicts too, but the 64-128 small TMA copies per tile cost
// as much as they save (double V = 8: +-0 %, float V = 8: -12 %), so those keep the single contiguous bulk copy.
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

#ifndef MC3D_TRI_NJ
#define MC3D_TRI_NJ 2
#endif
#ifndef MC3D_TRI_MTHREADS
#define MC3D_TRI_MTHREADS 128
#endif
#ifndef MC3D_TRI_MBLOCKS
#define MC3D_TRI_MBLOCKS 4
#endif
constexpr int TRI_NJ = MC3D_TRI_NJ;             // joints per thread in the mixed kernel
constexpr int TRI_MTHREADS = MC3D_TRI_MTHREADS; // threads per CTA of the mixed kernel
constexpr int TRI_MTILE = TRI_MTHREADS * TRI_NJ;    // joints per tile

// Mixed-precision kernel (float storage, weighted mode, V in {2,3,4,8,16}, no undistortion).  Same TMA ring as the
// generic kernel; 128 threads x 2 joints per thread (thread t owns joints t and t + 128 of a 256-joint tile), view
// loops unrolled four views at a time, camera constants in shared memory (128-bit broadcast loads), 4 CTAs per SM
// (3 for 16 views): the fastest on B200 of the (joints/thread, threads, CTAs/SM, unroll) variants timed
// (profiles/README.md).
template <int V, int LAYOUT>
__global__ void __launch_bounds__(TRI_MTHREADS, (V >= 16) ? MC3D_TRI_MBLOCKS - 1 : MC3D_TRI_MBLOCKS)
triangulate_mixed_kernel(const float *__restrict__ kpts, float *__restrict__ out, long long n, int n_stages,
                         const __grid_constant__ TriParams prm) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int row_elems = 3 * V;
    const RowPlan plan = tri_row_plan(row_elems, (int)sizeof(float));
    constexpr uint32_t row_bytes = (uint32_t)(row_elems * sizeof(float));
    const uint32_t stage_bytes = (uint32_t)(tri_stage_elems(TRI_MTILE, row_elems, (int)sizeof(float)) * sizeof(float));
    float *ring = reinterpret_cast<float *>(smem_raw);
    float *otile = reinterpret_cast<float *>(smem_raw + (size_t)n_stages * stage_bytes);
    uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw + (size_t)n_stages * stage_bytes + 2 * TRI_MTILE * 3 * sizeof(float));
    CamC *cam = reinterpret_cast<CamC *>(full + 8);          // per-view constants, read with 128-bit loads
    const int tid = threadIdx.x;
    const long long n_tiles = (n + TRI_MTILE - 1) / TRI_MTILE;
    const long long first = blockIdx.x, stride = gridDim.x;
    const long long my_tiles = (first < n_tiles) ? (n_tiles - first + stride - 1) / stride : 0;
    if (tid == 0) {
        for (int s = 0; s < n_stages; ++s) mbar_init(&full[s], 1);
        fence_mbar_init();
    }
    if (tid < V) {
        CamC c;
#pragma unroll
        for (int k = 0; k < 12; ++k) c.Pc[k] = MC3D_TRI_RAW_RESID ? prm.P[tid][k] : prm.Pc[tid][k];
        c.cxd = prm.cxyd[tid][0];
        c.cyd = prm.cxyd[tid][1];
#pragma unroll
        for (int k = 0; k < 4; ++k) c.p10[k] = prm.p10[tid][k];
        c.p2 = make_float4(prm.p2[tid][0].x, prm.p2[tid][1].x, prm.p2[tid][2].x, prm.p2[tid][3].x);
        c.cx = prm.cxy[tid].x;
        c.cy = prm.cxy[tid].y;
        c.pad0 = c.pad1 = 0.f;
        cam[tid] = c;
    }
    __syncthreads();
    auto tile_is_full = [&](long long tile) { return (tile + 1) * TRI_MTILE <= n; };
    auto issue_load = [&](long long k, int s) {   // every thread
        const long long tile = first + k * stride;
        if (k < my_tiles && tile_is_full(tile)) {
            unsigned char *dst = reinterpret_cast<unsigned char *>(ring) + (size_t)s * stage_bytes;
            const float *src = kpts + tile * TRI_MTILE * (long long)row_elems;
            if (plan.group) {
                if (tid == 0) mbar_arrive_expect_tx(&full[s], TRI_MTILE * row_bytes);
#pragma unroll
                for (int j = 0; j < TRI_NJ; ++j) {
                    const int slot = tid + j * TRI_MTHREADS;
                    if (slot % plan.group == 0)
                        bulk_g2s(dst + tri_row_offset(plan, slot, row_elems) * sizeof(float), src + (size_t)slot * row_elems,
                                 plan.group * row_bytes, &full[s]);
                }
            } else if (tid == 0) {
                mbar_arrive_expect_tx(&full[s], stage_bytes);
                bulk_g2s(dst, src, stage_bytes, &full[s]);
            }
        }
    };
    for (int k = 0; k < n_stages - 1; ++k) issue_load(k, k);
    int s = 0, s_next = n_stages - 1;
    uint32_t parity = 0;
    for (long long k = 0; k < my_tiles; ++k) {
        const long long tile = first + k * stride;
        const bool full_tile = tile_is_full(tile);
        float *stage = reinterpret_cast<float *>(reinterpret_cast<unsigned char *>(ring) + (size_t)s * stage_bytes);
        issue_load(k + n_stages - 1, s_next);
        if (tid == 0) bulk_wait_read<1>();
        if (full_tile) {
            mbar_wait(&full[s], parity);
        } else {
            const long long base = tile * TRI_MTILE * (long long)row_elems;
            const long long cnt = (n - tile * TRI_MTILE) * row_elems;
            for (long long i = tid; i < cnt; i += TRI_MTHREADS) stage[tri_row_offset(plan, (int)(i / row_elems), row_elems) + (i % row_elems)] = kpts[base + i];
            __syncthreads();
        }
        const float *rows[TRI_NJ];
        bool act[TRI_NJ];
        double X[TRI_NJ][3];
        int state[TRI_NJ];
#pragma unroll
        for (int j = 0; j < TRI_NJ; ++j) {
            const int slot = tid + j * TRI_MTHREADS;
            act[j] = tile * TRI_MTILE + slot < n;
            rows[j] = stage + tri_row_offset(plan, act[j] ? slot : tid, row_elems);      // inactive slots read a valid row
        }
#if MC3D_TRI_FLOAT_TAIL
        float Xo[TRI_NJ][3];
#endif
        solve_merged<V, LAYOUT, TRI_NJ>(cam, prm.rig2, rows, act, X, state MC3D_XO_ARG);
#pragma unroll
        for (int j = 0; j < TRI_NJ; ++j)
            if (act[j] && state[j] == 1) {
                double f0, f1, f2;                   // separate scalars: X[][] must not have its address taken
                solve_double_from_row(prm, rows[j], V, LAYOUT == MC3D_LAYOUT_3V, f0, f1, f2);
                X[j][0] = f0; X[j][1] = f1; X[j][2] = f2;
            }
#if MC3D_TRI_FLOAT_TAIL
#pragma unroll
        for (int j = 0; j < TRI_NJ; ++j)             // accepted joints left the solver as floats; widen for the common tail
            if (!(act[j] && state[j] == 1)) { X[j][0] = (double)Xo[j][0]; X[j][1] = (double)Xo[j][1]; X[j][2] = (double)Xo[j][2]; }
#endif
        __syncthreads();                       // [A] every thread has consumed stage s
        float *ot = otile + (size_t)(k & 1) * TRI_MTILE * 3;
        if (full_tile) {
#pragma unroll
            for (int j = 0; j < TRI_NJ; ++j) {
                const int slot = tid + j * TRI_MTHREADS;
                ot[slot * 3 + 0] = (float)X[j][0];
                ot[slot * 3 + 1] = (float)X[j][1];
                ot[slot * 3 + 2] = (float)X[j][2];
            }
            fence_proxy_async_smem();
            __syncthreads();                   // [B]
            if (tid == 0) {
                bulk_s2g(out + tile * TRI_MTILE * 3, ot, (uint32_t)(TRI_MTILE * 3 * sizeof(float)));
                bulk_commit();
            }
        } else {
#pragma unroll
            for (int j = 0; j < TRI_NJ; ++j) {
                const long long joint = tile * TRI_MTILE + tid + j * TRI_MTHREADS;
                if (act[j]) {
                    out[joint * 3 + 0] = (float)X[j][0];
                    out[joint * 3 + 1] = (float)X[j][1];
                    out[joint * 3 + 2] = (float)X[j][2];
                }
            }
            if (tid == 0) bulk_commit();
        }
        if (++s == n_stages) { s = 0; parity ^= 1u; }
        if (++s_next == n_stages) s_next = 0;
    }
    if (tid == 0) bulk_wait_all<0>();
}


// ---- lean tile loop for the mixed kernel (MC3D_TRI_LEAN, tuning build) ----------------------------------------------------
// Same ring, same solver, same results; what differs is the per-tile bookkeeping, which costs the kernel above ~75 of its
// ~675 instructions per joint: this one processes FULL tiles only (the host launches the kernel above on the ragged tail),
// so there is no activity mask and no ragged path in the loop; global addresses are running pointers instead of 64-bit
// products; only thread 0 evaluates the refill; the output tile is written straight from the solver's registers with the
// rare all-double fallback patching its slot afterwards (no register shuffling around an out-of-line call on the hot path);
// and one block barrier per tile instead of two (thread 0 waits for the previous bulk store to have been read BEFORE the
// barrier, which publishes that the other output buffer is free again).
#ifndef MC3D_TRI_LEAN
#define MC3D_TRI_LEAN 1
#endif
#if MC3D_TRI_LEAN
template <int V>
__device__ __noinline__ void mixed_cold_fix(const TriParams &prm, const float *row0, const float *row1, int st0, int st1,
                                            bool l3v, float *ot0, float *ot1) {
    if (st0 == 1) {
        double f0, f1, f2;
        solve_double_from_row(prm, row0, V, l3v, f0, f1, f2);
        ot0[0] = (float)f0; ot0[1] = (float)f1; ot0[2] = (float)f2;
    }
    if (st1 == 1) {
        double f0, f1, f2;
        solve_double_from_row(prm, row1, V, l3v, f0, f1, f2);
        ot1[0] = (float)f0; ot1[1] = (float)f1; ot1[2] = (float)f2;
    }
}

template <int V, int LAYOUT>
__global__ void __launch_bounds__(TRI_MTHREADS, (V >= 16) ? MC3D_TRI_MBLOCKS - 1 : MC3D_TRI_MBLOCKS)
triangulate_mixed_lean_kernel(const float *__restrict__ kpts, float *__restrict__ out, unsigned n_tiles, int n_stages,
                              const __grid_constant__ TriParams prm) {
    static_assert(TRI_NJ == 2, "the lean loop is written for two joints per thread");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int row_elems = 3 * V;
    constexpr uint32_t stage_bytes = (uint32_t)(TRI_MTILE * row_elems * sizeof(float));      // contiguous tiles only
    constexpr uint32_t otile_bytes = (uint32_t)(TRI_MTILE * 3 * sizeof(float));
    static_assert((row_elems * sizeof(float)) % 128 != 0, "padded row plans use the kernel above");
    unsigned char *ring = smem_raw;
    float *otile = reinterpret_cast<float *>(smem_raw + (size_t)n_stages * stage_bytes);
    uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw + (size_t)n_stages * stage_bytes + 2 * otile_bytes);
    CamC *cam = reinterpret_cast<CamC *>(full + 8);
    const int tid = threadIdx.x;
    const unsigned first = blockIdx.x, stride = gridDim.x;
    const unsigned my_tiles = (first < n_tiles) ? (n_tiles - first + stride - 1) / stride : 0;
    if (tid == 0) {
        for (int s = 0; s < n_stages; ++s) mbar_init(&full[s], 1);
        fence_mbar_init();
    }
    if (tid < V) {
        CamC c;
#pragma unroll
        for (int k = 0; k < 12; ++k) c.Pc[k] = MC3D_TRI_RAW_RESID ? prm.P[tid][k] : prm.Pc[tid][k];
        c.cxd = prm.cxyd[tid][0];
        c.cyd = prm.cxyd[tid][1];
#pragma unroll
        for (int k = 0; k < 4; ++k) c.p10[k] = prm.p10[tid][k];
        c.p2 = make_float4(prm.p2[tid][0].x, prm.p2[tid][1].x, prm.p2[tid][2].x, prm.p2[tid][3].x);
        c.cx = prm.cxy[tid].x;
        c.cy = prm.cxy[tid].y;
        c.pad0 = c.pad1 = 0.f;
        cam[tid] = c;
    }
    __syncthreads();
    // thread 0 only: global addresses of the tiles this CTA loads and stores (computed inside its branches, so the other
    // warps neither execute the 64-bit arithmetic nor hold the pointers in registers)
    auto load_tile = [&](unsigned i, int st) {        // i-th tile of this CTA into stage st
        const unsigned char *src = reinterpret_cast<const unsigned char *>(kpts) + ((size_t)first + (size_t)i * stride) * stage_bytes;
        mbar_arrive_expect_tx(&full[st], stage_bytes);
        bulk_g2s(ring + (size_t)st * stage_bytes, src, stage_bytes, &full[st]);
    };
    if (tid == 0)
        for (unsigned i = 0; i < (unsigned)(n_stages - 1) && i < my_tiles; ++i) load_tile(i, (int)i);
    int s = 0, s_next = n_stages - 1;
    uint32_t parity = 0;
    for (unsigned k = 0; k < my_tiles; ++k) {
        if (tid == 0 && k + (unsigned)(n_stages - 1) < my_tiles) load_tile(k + (unsigned)(n_stages - 1), s_next);   // refills the stage of iteration k - 1
        const float *stage = reinterpret_cast<const float *>(ring + (size_t)s * stage_bytes);
        mbar_wait(&full[s], parity);
        const float *rows[TRI_NJ] = {stage + tid * row_elems, stage + (tid + TRI_MTHREADS) * row_elems};
        const bool act[TRI_NJ] = {true, true};
        double X[TRI_NJ][3];
        int state[TRI_NJ];
        float *ot = otile + (size_t)(k & 1) * TRI_MTILE * 3 + tid * 3;
#if MC3D_TRI_FLOAT_TAIL
        float Xo[TRI_NJ][3];
        solve_merged<V, LAYOUT, TRI_NJ>(cam, prm.rig2, rows, act, X, state, Xo);
        ot[0] = Xo[0][0]; ot[1] = Xo[0][1]; ot[2] = Xo[0][2];
        ot[TRI_MTHREADS * 3 + 0] = Xo[1][0]; ot[TRI_MTHREADS * 3 + 1] = Xo[1][1]; ot[TRI_MTHREADS * 3 + 2] = Xo[1][2];
#else
        solve_merged<V, LAYOUT, TRI_NJ>(cam, prm.rig2, rows, act, X, state);
        ot[0] = (float)X[0][0]; ot[1] = (float)X[0][1]; ot[2] = (float)X[0][2];
        ot[TRI_MTHREADS * 3 + 0] = (float)X[1][0]; ot[TRI_MTHREADS * 3 + 1] = (float)X[1][1]; ot[TRI_MTHREADS * 3 + 2] = (float)X[1][2];
#endif
        if (state[0] == 1 || state[1] == 1)
            mixed_cold_fix<V>(prm, stage + tid * row_elems, stage + (tid + TRI_MTHREADS) * row_elems, state[0], state[1],
                              LAYOUT == MC3D_LAYOUT_3V, ot, ot + TRI_MTHREADS * 3);
        fence_proxy_async_smem();
        if (tid == 0) bulk_wait_read<0>();            // the store of iteration k - 1 has left the OTHER output buffer
        __syncthreads();                              // stage s consumed by everyone; output tile complete
        if (tid == 0) {
            bulk_s2g(reinterpret_cast<unsigned char *>(out) + ((size_t)first + (size_t)k * stride) * otile_bytes,
                     otile + (size_t)(k & 1) * TRI_MTILE * 3, otile_bytes);
            bulk_commit();
        }
        if (++s == n_stages) { s = 0; parity ^= 1u; }
        if (++s_next == n_stages) s_next = 0;
    }
    if (tid == 0) bulk_wait_all<0>();
}
#endif  // MC3D_TRI_LEAN


// ---- kernel -----------------------------------------------------------------------------------
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
    for (int v = 0; v < rig->n_views; ++v) {
        const double *P0 = prm.P[v], *P1 = prm.P[v] + 4, *P2 = prm.P[v] + 8;
        const double n2 = P2[0] * P2[0] + P2[1] * P2[1] + P2[2] * P2[2];
        // principal point of the view: any value keeps the rows identical, this one keeps their entries small
        const float cx = n2 > 0 ? (float)((P0[0] * P2[0] + P0[1] * P2[1] + P0[2] * P2[2]) / n2) : 0.f;
        const float cy = n2 > 0 ? (float)((P1[0] * P2[0] + P1[1] * P2[1] + P1[2] * P2[2]) / n2) : 0.f;
        prm.cxy[v] = make_float2(cx, cy);
        prm.cxyd[v][0] = (double)cx;
        prm.cxyd[v][1] = (double)cy;
        for (int k = 0; k < 4; ++k) {
            const double p0c = P0[k] - (double)cx * P2[k], p1c = P1[k] - (double)cy * P2[k];
            prm.Pc[v][k] = p0c; prm.Pc[v][4 + k] = p1c; prm.Pc[v][8 + k] = P2[k];
            prm.p2[v][k] = make_float2((float)P2[k], (float)P2[k]);
            prm.p10[v][k] = make_float2((float)-p1c, (float)p0c);
        }
    }
    double acc = 0.0;
    int cnt = 0;
    for (int v = 0; v < rig->n_views; ++v) {                 // camera centre C: P (C,1) = 0
        const double *p = prm.P[v];
        const double a = p[0], b = p[1], c = p[2], d = p[4], e = p[5], f = p[6], g = p[8], h = p[9], i = p[10];
        const double det = a * (e * i - f * h) - b * (d * i - f * g) + c * (d * h - e * g);
        if (!(fabs(det) > 0.0)) continue;
        const double r0 = -p[3], r1 = -p[7], r2 = -p[11];
        const double c0 = (r0 * (e * i - f * h) - b * (r1 * i - f * r2) + c * (r1 * h - e * r2)) / det;
        const double c1 = (a * (r1 * i - f * r2) - r0 * (d * i - f * g) + c * (d * r2 - r1 * g)) / det;
        const double c2 = (a * (e * r2 - r1 * h) - b * (d * r2 - r1 * g) + r0 * (d * h - e * g)) / det;
        const double n2 = c0 * c0 + c1 * c1 + c2 * c2;
        if (n2 <= 1.0e300) { acc += n2; ++cnt; }
    }
    prm.rig2 = cnt ? (float)fmin(acc / cnt, 1.0e30) : 0.f;
    return MC3D_OK;
}

// ---- lean tile loop for double storage (MC3D_TRI_LEAN64, tuning build) -------------------------------------------------------
// The generic kernel above spends ~170 of its ~1 030 instructions per joint on per-tile bookkeeping (one joint per thread,
// so nothing amortises it).  This one is the weighted, undistortion-free, compile-time-V, contiguous-tile case only: full
// tiles (the host launches the generic kernel on the ragged tail), layout as a template parameter, global addresses computed
// in thread 0's branches.  Same accumulation order, same solver: bit-identical results.
#ifndef MC3D_TRI_LEAN64
#define MC3D_TRI_LEAN64 1
#endif
#if MC3D_TRI_LEAN64
template <int V, int LAYOUT>
__global__ void __launch_bounds__(TRI_TILE, (V <= 8) ? 3 : 2)
triangulate_lean64_kernel(const double *__restrict__ kpts, double *__restrict__ out, unsigned n_tiles, int n_stages,
                          const __grid_constant__ TriParams prm) {
    static_assert(V > 0 && V % 2 == 0, "even compile-time view count");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int row_elems = 3 * V;
    constexpr uint32_t stage_bytes = (uint32_t)(TRI_TILE * row_elems * sizeof(double));
    constexpr uint32_t otile_bytes = (uint32_t)(TRI_TILE * 3 * sizeof(double));
    static_assert((row_elems * sizeof(double)) % 128 != 0, "padded row plans use the generic kernel");
    unsigned char *ring = smem_raw;
    double *otile = reinterpret_cast<double *>(smem_raw + (size_t)n_stages * stage_bytes);
    uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw + (size_t)n_stages * stage_bytes + otile_bytes);
    const int tid = threadIdx.x;
    const unsigned first = blockIdx.x, stride = gridDim.x;
    const unsigned my_tiles = (first < n_tiles) ? (n_tiles - first + stride - 1) / stride : 0;
    if (tid == 0) {
        for (int s = 0; s < n_stages; ++s) mbar_init(&full[s], 1);
        fence_mbar_init();
    }
    __syncthreads();
    auto load_tile = [&](unsigned i, int st) {
        const unsigned char *src = reinterpret_cast<const unsigned char *>(kpts) + ((size_t)first + (size_t)i * stride) * stage_bytes;
        mbar_arrive_expect_tx(&full[st], stage_bytes);
        bulk_g2s(ring + (size_t)st * stage_bytes, src, stage_bytes, &full[st]);
    };
    if (tid == 0)
        for (unsigned i = 0; i < (unsigned)(n_stages - 1) && i < my_tiles; ++i) load_tile(i, (int)i);
    int s = 0, s_next = n_stages - 1;
    uint32_t parity = 0;
    for (unsigned k = 0; k < my_tiles; ++k) {
        if (tid == 0) {
            if (k + (unsigned)(n_stages - 1) < my_tiles) load_tile(k + (unsigned)(n_stages - 1), s_next);
            bulk_wait_read<0>();                   // the output tile of iteration k - 1 has left shared memory
        }
        const double *rowd = reinterpret_cast<const double *>(ring + (size_t)s * stage_bytes) + tid * row_elems;
        mbar_wait(&full[s], parity);
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
        __syncthreads();                           // [A] stage s consumed; thread 0's wait on the previous store is published
        double X0 = NAN, X1 = NAN, X2 = NAN;
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
        otile[tid * 3 + 0] = X0;
        otile[tid * 3 + 1] = X1;
        otile[tid * 3 + 2] = X2;
        fence_proxy_async_smem();
        __syncthreads();                           // [B] tile complete and visible to the async proxy
        if (tid == 0) {
            bulk_s2g(reinterpret_cast<unsigned char *>(out) + ((size_t)first + (size_t)k * stride) * otile_bytes, otile, otile_bytes);
            bulk_commit();
        }
        if (++s == n_stages) { s = 0; parity ^= 1u; }
        if (++s_next == n_stages) s_next = 0;
    }
    if (tid == 0) bulk_wait_all<0>();
}
#endif  // MC3D_TRI_LEAN64

template <typename T, int V, int MODE, bool UNDISTORT>
static int launch_one(const T *d_kpts, long long n, const TriParams &prm, T *d_out, cudaStream_t stream) {
    const int nv = prm.n_views;
    const size_t stage_bytes = tri_stage_elems(TRI_TILE, 3 * nv, (int)sizeof(T)) * sizeof(T);
    auto kern = triangulate_kernel<T, V, MODE, UNDISTORT>;
    static bool attr_done = false;     // per instantiation
    if (!attr_done) {
        MC3D_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr_done = true;
    }
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
#if MC3D_TRI_LEAN64
    // tuning build: weighted double storage with a compile-time even view count -- full tiles through the lean loop, the
    // ragged tail through the generic kernel
    if constexpr (std::is_same<T, double>::value && MODE == MC3D_TRI_WEIGHTED && !UNDISTORT && V > 0 && V % 2 == 0 &&
                  (3 * V * sizeof(double)) % 128 != 0) {              // rows of a multiple of 128 bytes use padded slots
        if (tri_row_plan(3 * V, (int)sizeof(double)).group == 0 && n / TRI_TILE > 0 && n / TRI_TILE < 0x7fffffffLL) {
            const long long n_full = n / TRI_TILE, tail = n - n_full * TRI_TILE;
            long long lgrid = (long long)sm_count() * per_sm;
            if (lgrid > n_full) lgrid = n_full;
            static bool lean_attr_done[2] = {false, false};
            const int li = prm.layout == MC3D_LAYOUT_3V ? 1 : 0;
            if (li) {
                auto lean = triangulate_lean64_kernel<V, MC3D_LAYOUT_3V>;
                if (!lean_attr_done[li]) { MC3D_CUDA_TRY(cudaFuncSetAttribute(lean, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)); lean_attr_done[li] = true; }
                lean<<<(unsigned)lgrid, TRI_TILE, smem, stream>>>(d_kpts, d_out, (unsigned)n_full, n_stages, prm);
            } else {
                auto lean = triangulate_lean64_kernel<V, MC3D_LAYOUT_V3>;
                if (!lean_attr_done[li]) { MC3D_CUDA_TRY(cudaFuncSetAttribute(lean, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)); lean_attr_done[li] = true; }
                lean<<<(unsigned)lgrid, TRI_TILE, smem, stream>>>(d_kpts, d_out, (unsigned)n_full, n_stages, prm);
            }
            count_launch();
            MC3D_CUDA_TRY(cudaGetLastError());
            if (tail > 0) {
                kern<<<1, TRI_TILE, smem, stream>>>(d_kpts + n_full * TRI_TILE * 3 * V, d_out + n_full * TRI_TILE * 3, tail, n_stages, prm);
                count_launch();
                MC3D_CUDA_TRY(cudaGetLastError());
            }
            return MC3D_OK;
        }
    }
#endif
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
    const size_t stage_bytes = tri_stage_elems(TRI_MTILE, 3 * V, (int)sizeof(float)) * sizeof(float);
    auto kern = triangulate_mixed_kernel<V, LAYOUT>;
    static bool attr_done = false;
    if (!attr_done) {
        MC3D_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr_done = true;
    }
    const size_t fixed = 2 * TRI_MTILE * 3 * sizeof(float) + 8 * sizeof(uint64_t) + V * sizeof(CamC);
    int n_stages = 2, per_sm = 0;
    size_t smem = 0;
    for (int st = 3; st >= 2; --st) {
        const size_t sm_bytes = st * stage_bytes + fixed;
        if (sm_bytes > 227 * 1024) continue;
        int occ = 0;
        MC3D_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, TRI_MTHREADS, sm_bytes));
        if (occ > per_sm) { per_sm = occ; n_stages = st; smem = sm_bytes; }
    }
    if (per_sm < 1) { set_error("mixed triangulate kernel does not fit in shared memory"); return MC3D_ERR_UNSUPPORTED; }
#if MC3D_TRI_LEAN
    // tuning build: full tiles through the lean loop, the ragged tail (< one tile) through the kernel above
    if (tri_row_plan(3 * V, (int)sizeof(float)).group == 0 && n / TRI_MTILE > 0 && n / TRI_MTILE < 0x7fffffffLL) {
        auto lean = triangulate_mixed_lean_kernel<V, LAYOUT>;
        static bool lean_attr_done = false;
        if (!lean_attr_done) {
            MC3D_CUDA_TRY(cudaFuncSetAttribute(lean, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
            lean_attr_done = true;
        }
        const long long n_full = n / TRI_MTILE, tail = n - n_full * TRI_MTILE;
        long long lgrid = (long long)sm_count() * per_sm;
        if (lgrid > n_full) lgrid = n_full;
        lean<<<(unsigned)lgrid, TRI_MTHREADS, smem, stream>>>(d_kpts, d_out, (unsigned)n_full, n_stages, prm);
        count_launch();
        MC3D_CUDA_TRY(cudaGetLastError());
        if (tail > 0) {
            kern<<<1, TRI_MTHREADS, smem, stream>>>(d_kpts + n_full * TRI_MTILE * 3 * V, d_out + n_full * TRI_MTILE * 3, tail,
                                                    n_stages, prm);
            count_launch();
            MC3D_CUDA_TRY(cudaGetLastError());
        }
        return MC3D_OK;
    }
#endif
    const long long n_tiles = (n + TRI_MTILE - 1) / TRI_MTILE;
    long long grid = (long long)sm_count() * per_sm;
    if (grid > n_tiles) grid = n_tiles;
    kern<<<(unsigned)grid, TRI_MTHREADS, smem, stream>>>(d_kpts, d_out, n, n_stages, prm);
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

int mc3d_triangulate_f32(const float *d_kpts, int64_t n, const mc3d_rig *rig, int layout, int mode, int flags,
                         float *d_out, void *stream) {
    return mc3d::triangulate_device<float>(d_kpts, n, rig, layout, mode, flags, d_out, (cudaStream_t)stream);
}

int mc3d_triangulate_f64(const double *d_kpts, int64_t n, const mc3d_rig *rig, int layout, int mode, int flags,
                         double *d_out, void *stream) {
    return mc3d::triangulate_device<double>(d_kpts, n, rig, layout, mode, flags, d_out, (cudaStream_t)stream);
}

}  // extern "C"
