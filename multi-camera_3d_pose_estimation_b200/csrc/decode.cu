// Heatmap -> (keypoint, score) + Gaussian moments, one pass over HBM, for sm_100a.
//
// Per heatmap (H x W floats):
//   * argmax (first maximum), score = max, quarter-pixel shift towards the larger neighbour
//     (the mmpose MSRA decode the reference gets from mmpose_pose_estimation.py:253-259);
//   * values < thr -> 0, then sum / mean / central second moments of the normalised map:
//     [mean_x, mean_y, var_x, cov_xy, cov_xy, var_y]   (reference mmpose_pose_estimation.py:163-215).
//
// One warp owns one heatmap at a time.  Each warp runs its own ring of shared-memory stages filled by
// 1-D TMA bulk copies (a 64x48 map is one contiguous 12 KB copy) so several maps per warp are in flight
// while the lanes make two passes over the resident map with conflict-free 128-bit shared loads and
// warp-shuffle reductions.  HBM is read exactly once; 72 output bytes per map.
#include "mc3d_common.cuh"
#include <math.h>

namespace mc3d {

struct DecodeParams {
    int H, W;
    float thr;
    int write_back;        // store the thresholded map back to global memory (reference quirk Q7)
    int kpt_layout;        // MC3D_KPT_PLAIN / _NV3 / _N3V
    int views, joints;     // for the transposing layouts: maps are ordered (T, C=views, J=joints)
    int has_affine;
    int affine_group;      // maps per affine entry
};

struct MapStats {
    double s, mx, my, vx, vy, cxy;
    float best;
    int best_idx;
};

__device__ __forceinline__ void argmax_merge(float &bv, int &bi, float ov, int oi) {
    if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
}

// Two passes over one resident map.  VEC4: W % 4 == 0 and src 16-byte aligned.
template <bool VEC4>
__device__ __forceinline__ MapStats map_stats(const float *src, float *wb, int H, int W, float thr, int lane) {
    const int HW = H * W;
    const float invW = 1.0f / (float)W;
    float s = 0.f, sx = 0.f, sy = 0.f;
    float best = -INFINITY;
    int best_idx = 0x7fffffff;
    if (VEC4) {
        const float4 *src4 = reinterpret_cast<const float4 *>(src);
        for (int i4 = lane; i4 < (HW >> 2); i4 += 32) {
            const float4 v = src4[i4];
            const int e = i4 << 2;
            const int y = (int)(((float)e + 0.5f) * invW);
            const int x = e - y * W;
            if (v.x > best) { best = v.x; best_idx = e; }
            if (v.y > best) { best = v.y; best_idx = e + 1; }
            if (v.z > best) { best = v.z; best_idx = e + 2; }
            if (v.w > best) { best = v.w; best_idx = e + 3; }
            const float t0 = v.x < thr ? 0.f : v.x, t1 = v.y < thr ? 0.f : v.y;
            const float t2 = v.z < thr ? 0.f : v.z, t3 = v.w < thr ? 0.f : v.w;
            if (wb) reinterpret_cast<float4 *>(wb)[i4] = make_float4(t0, t1, t2, t3);
            const float q = (t0 + t1) + (t2 + t3);
            const float tx = fmaf(3.f, t3, fmaf(2.f, t2, t1));
            s += q;
            sx += fmaf((float)x, q, tx);
            sy = fmaf((float)y, q, sy);
        }
    } else {
        for (int e = lane; e < HW; e += 32) {
            const float v = src[e];
            const int y = (int)(((float)e + 0.5f) * invW);
            const int x = e - y * W;
            if (v > best) { best = v; best_idx = e; }
            const float t = v < thr ? 0.f : v;
            if (wb) wb[e] = t;
            s += t;
            sx = fmaf((float)x, t, sx);
            sy = fmaf((float)y, t, sy);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, best_idx, o);
        argmax_merge(best, best_idx, ov, oi);
    }
    MapStats r;
    r.best = best;
    r.best_idx = best_idx;
    r.s = warp_sum((double)s);
    const double dsx = warp_sum((double)sx), dsy = warp_sum((double)sy);
    r.mx = r.my = r.vx = r.vy = r.cxy = 0.0;
    if (r.s == 0.0) return r;                           // all-zero map: six zeros (:191-193)
    r.mx = dsx / r.s;
    r.my = dsy / r.s;
    const float fmx = (float)r.mx, fmy = (float)r.my;
    // residual of the float means, applied analytically so the central moments are about the double means
    float sxx = 0.f, syy = 0.f, sxy = 0.f, sdx = 0.f, sdy = 0.f;
    if (VEC4) {
        const float4 *src4 = reinterpret_cast<const float4 *>(src);
        for (int i4 = lane; i4 < (HW >> 2); i4 += 32) {
            const float4 v = src4[i4];
            const int e = i4 << 2;
            const int y = (int)(((float)e + 0.5f) * invW);
            const int x = e - y * W;
            const float t0 = v.x < thr ? 0.f : v.x, t1 = v.y < thr ? 0.f : v.y;
            const float t2 = v.z < thr ? 0.f : v.z, t3 = v.w < thr ? 0.f : v.w;
            const float dx0 = (float)x - fmx, dy = (float)y - fmy;
            const float dx1 = dx0 + 1.f, dx2 = dx0 + 2.f, dx3 = dx0 + 3.f;
            const float q = (t0 + t1) + (t2 + t3);
            const float m1 = fmaf(dx0, t0, fmaf(dx1, t1, fmaf(dx2, t2, dx3 * t3)));
            sxx += fmaf(dx0 * dx0, t0, fmaf(dx1 * dx1, t1, fmaf(dx2 * dx2, t2, dx3 * dx3 * t3)));
            sdx += m1;
            sxy = fmaf(dy, m1, sxy);
            syy = fmaf(dy * dy, q, syy);
            sdy = fmaf(dy, q, sdy);
        }
    } else {
        for (int e = lane; e < HW; e += 32) {
            const float v = src[e];
            const int y = (int)(((float)e + 0.5f) * invW);
            const int x = e - y * W;
            const float t = v < thr ? 0.f : v;
            const float dx = (float)x - fmx, dy = (float)y - fmy;
            sxx = fmaf(dx * dx, t, sxx);
            syy = fmaf(dy * dy, t, syy);
            sxy = fmaf(dx * dy, t, sxy);
            sdx = fmaf(dx, t, sdx);
            sdy = fmaf(dy, t, sdy);
        }
    }
    const double inv = 1.0 / r.s;
    const double ex = warp_sum((double)sdx) * inv, ey = warp_sum((double)sdy) * inv;   // E[x - fmx], E[y - fmy]
    r.vx = warp_sum((double)sxx) * inv - ex * ex;
    r.vy = warp_sum((double)syy) * inv - ey * ey;
    r.cxy = warp_sum((double)sxy) * inv - ex * ey;
    return r;
}

__device__ __forceinline__ long long div_small(long long a, int b);
__device__ __forceinline__ void store_keypoint(float kx, float ky, float score, long long map, const DecodeParams &p, float *__restrict__ kpt);

// map index / small divisor: 32-bit when the index allows (a 64-bit division is ~100 instructions, and the transposing layouts
// with an affine need two per map -- 10 % of the moments kernel)
__device__ __forceinline__ long long div_small(long long a, int b) {
    return (a >> 31) == 0 ? (long long)((unsigned)a / (unsigned)b) : a / b;
}

__device__ __forceinline__ void write_outputs(const MapStats &r, const float *src, long long map, const DecodeParams &p,
                                              const float *__restrict__ affine, float *__restrict__ kpt,
                                              double *__restrict__ moments, int lane) {
    if (moments && lane < 6) {
        const double val = lane == 0 ? r.mx : lane == 1 ? r.my : lane == 2 ? r.vx : lane == 5 ? r.vy : r.cxy;
        moments[map * 6 + lane] = val;
    }
    if (kpt && lane == 0) {
        float kx = -1.f, ky = -1.f;
        const float score = r.best;
        if (score > 0.f) {                              // mmpose: maxima <= 0 are "not found" (-1, -1)
            const int py = r.best_idx / p.W, px = r.best_idx - py * p.W;
            kx = (float)px;
            ky = (float)py;
            if (px > 1 && px < p.W - 1 && py > 1 && py < p.H - 1) {
                const float dx = src[py * p.W + px + 1] - src[py * p.W + px - 1];
                const float dy = src[(py + 1) * p.W + px] - src[(py - 1) * p.W + px];
                kx += dx > 0.f ? 0.25f : (dx < 0.f ? -0.25f : 0.f);
                ky += dy > 0.f ? 0.25f : (dy < 0.f ? -0.25f : 0.f);
            }
            if (p.has_affine) {
                const float *a = affine + div_small(map, p.affine_group) * 4;
                kx = fmaf(kx, a[0], a[2]);
                ky = fmaf(ky, a[1], a[3]);
            }
        }
        store_keypoint(kx, ky, score, map, p, kpt);
    }
}


__device__ __forceinline__ void store_keypoint(float kx, float ky, float score, long long map, const DecodeParams &p, float *__restrict__ kpt) {
    if (p.kpt_layout == MC3D_KPT_PLAIN) {
        kpt[map * 3 + 0] = kx; kpt[map * 3 + 1] = ky; kpt[map * 3 + 2] = score;
    } else {
        const int cj = p.views * p.joints;
        const long long t = div_small(map, cj);
        const int rem = (int)(map - t * cj);
        const int c = rem / p.joints, j = rem - c * p.joints;
        const long long tj = t * p.joints + j;
        if (p.kpt_layout == MC3D_KPT_NV3) {
            float *o = kpt + (tj * p.views + c) * 3;
            o[0] = kx; o[1] = ky; o[2] = score;
        } else {
            float *o = kpt + tj * 3 * p.views + c;
            o[0] = kx; o[p.views] = ky; o[2 * p.views] = score;
        }
    }
}

// ---- register-resident kernel for 64 x 48 maps (the reference's heatmap size, BASELINE config 3) ----------------
// One warp per map; every lane pulls its 24 float4 (element e = 4 lane + 128 k, k = 0..23) with streaming 128-bit
// loads straight into registers -- 12 KB in flight per warp, 16 warps per SM -- and both passes run from registers:
// no shared memory, no re-read.  With W = 48 the lane's column quad repeats with period 3 in k and its row is
// 8 (k / 3) + const, so coordinates cost one add per float4.
__device__ __forceinline__ float4 ldg_stream(const float4 *p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}

#ifndef MC3D_DEC_BLOCKS
#define MC3D_DEC_BLOCKS 4        // CTAs per SM of the unstaged moments kernel (128 registers, 240 B spilled: occupancy beats spills)
#endif
#ifndef MC3D_DEC_STAGED_BLOCKS
#define MC3D_DEC_STAGED_BLOCKS 3 // CTAs per SM of the staged one: 168 registers hold the map (measured 0.83 of the HBM peak; 4 CTAs x
#endif                           // 128 registers with their spills 0.46, 2 CTAs x 255 0.75, two stages per warp 0.71)
#ifndef MC3D_DEC_STAGES
#define MC3D_DEC_STAGES 1
#endif
// MOMENTS = false (keypoints only, e.g. decode -> triangulate): the thresholding, both moment passes and six of the
// eight warp reductions disappear and the kernel is a pure streaming max.
// STAGED (the moments variant): the two passes cost ~1 850 issue slots per map and 16 resident warps cannot hide the ~3 000
// cycles a map takes to arrive behind them, so each warp also owns ONE 12 KB shared-memory stage: its lane 0 starts the 1-D TMA
// bulk copy of the warp's NEXT map as soon as the current one has been lifted from the stage into registers (24 conflict-free
// LDS.128 per lane), and the copy travels while the current map is reduced.  16 warps x 12 KB = 192 KB per SM in flight.
template <bool MOMENTS, bool STAGED = false>
__global__ void __launch_bounds__(128, STAGED ? MC3D_DEC_STAGED_BLOCKS : (MOMENTS ? MC3D_DEC_BLOCKS : 4))
decode_reg6448_kernel(const float *__restrict__ hm, float *__restrict__ hm_wb, long long n_maps,
                      const float *__restrict__ affine, float *__restrict__ kpt, double *__restrict__ moments,
                      const __grid_constant__ DecodeParams p) {
    constexpr int W = 48, HW = 64 * 48, NK = HW / 128;          // 24 float4 per lane
    constexpr int NST = MC3D_DEC_STAGES;
    extern __shared__ __align__(128) unsigned char dec_stage_raw[];
    __shared__ __align__(8) uint64_t dec_bars[4 * NST];
    const int lane = threadIdx.x & 31;
    const float4 *stage = reinterpret_cast<const float4 *>(dec_stage_raw) + (threadIdx.x >> 5) * (NST * HW / 4);
    uint64_t *bar = &dec_bars[(threadIdx.x >> 5) * NST];
    uint32_t uses = 0;
    const long long gwarp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long stride = ((long long)gridDim.x * blockDim.x) >> 5;
    // lane constants: k = 3 j + r  ->  x0 = X[r], y = 8 j + Y[r]
    float Xr[3], Yr[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        const int u = 4 * lane + 32 * r;
        Xr[r] = (float)(u % W);
        Yr[r] = (float)(2 * r + u / W);
    }
    const float thr = p.thr;
    if (STAGED) {
        if (lane == 0) {
            for (int q = 0; q < NST; ++q) mbar_init(bar + q, 1);
            fence_mbar_init();
            for (int q = 0; q < NST; ++q)
                if (gwarp + q * stride < n_maps) {
                    mbar_arrive_expect_tx(bar + q, HW * 4);
                    bulk_g2s(const_cast<float4 *>(stage) + q * (HW / 4), hm + (gwarp + q * stride) * HW, HW * 4, bar + q);
                }
        }
        __syncwarp();
    }
    for (long long map = gwarp; map < n_maps; map += stride) {
        const float4 *src = reinterpret_cast<const float4 *>(hm + map * HW) + lane;
        float4 v[NK];
        if (STAGED) {
            const uint32_t q = uses % NST;
            mbar_wait(bar + q, (uses / NST) & 1u);
            ++uses;
            const float4 *st = stage + q * (HW / 4);
#pragma unroll
            for (int k = 0; k < NK; ++k) v[k] = st[lane + 32 * k];
            __syncwarp();                                        // every lane has lifted its share: the stage is free
            if (lane == 0 && map + NST * stride < n_maps) {
                mbar_arrive_expect_tx(bar + q, HW * 4);
                bulk_g2s(const_cast<float4 *>(st), hm + (map + NST * stride) * HW, HW * 4, bar + q);
            }
        } else {
#pragma unroll
            for (int k = 0; k < NK; ++k) v[k] = ldg_stream(src + 32 * k);
        }
        // pass 1: argmax on the raw values, threshold in place, zeroth and first moments
        float best = -INFINITY, s = 0.f, sx = 0.f, sy = 0.f;
        int bk = 0;
#pragma unroll
        for (int k = 0; k < NK; ++k) {
            const float m4 = fmaxf(fmaxf(v[k].x, v[k].y), fmaxf(v[k].z, v[k].w));
            if (m4 > best) { best = m4; bk = k; }
            if (!MOMENTS) continue;
            v[k].x = v[k].x < thr ? 0.f : v[k].x;
            v[k].y = v[k].y < thr ? 0.f : v[k].y;
            v[k].z = v[k].z < thr ? 0.f : v[k].z;
            v[k].w = v[k].w < thr ? 0.f : v[k].w;
            const float q = (v[k].x + v[k].y) + (v[k].z + v[k].w);
            const float tx = fmaf(3.f, v[k].w, fmaf(2.f, v[k].z, v[k].y));
            const float yk = Yr[k % 3] + (float)(8 * (k / 3));
            s += q;
            sx += fmaf(Xr[k % 3], q, tx);
            sy = fmaf(yk, q, sy);
        }
        // NaN anywhere: fmaxf drops it from the argmax (as the smem kernel's `>` does) but the sums carry it
        int best_idx = 4 * lane + 128 * bk;                  // base of the float4 holding this lane's maximum
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, best_idx, o);
            argmax_merge(best, best_idx, ov, oi);
        }
        // The keypoint needs raw values around the maximum -- the winning float4 (which of its four elements is the first
        // maximum), the elements left and right of it and the float4s one row above and below (the quarter-pixel shift) -- and the
        // map's affine.  Fourteen lanes fetch one scalar each NOW (the map has just come through L2), four more the affine, so
        // that the loads travel behind the moment reductions instead of forming a chain of L2 round trips in lane 0 at the end
        // of every map (measured with moments: plain layout 0.88 -> 0.90 of the HBM peak, transposed layout with affine 0.80 -> 0.82; 1.01 without keypoints).
        float nbv = 0.f, affv = 0.f;
        if (kpt) {
            const float *raw = hm + map * HW;
            const int a = lane < 4 ? best_idx + lane : lane < 8 ? best_idx - W + (lane - 4) : lane < 12 ? best_idx + W + (lane - 8)
                        : lane == 12 ? best_idx - 1 : lane == 13 ? best_idx + 4 : -1;
            if (a >= 0 && a < HW) nbv = raw[a];
            if (p.has_affine && lane >= 16 && lane < 20) affv = affine[div_small(map, p.affine_group) * 4 + (lane - 16)];
        }
        MapStats r;
        r.best = best;
        r.best_idx = best_idx;
        r.s = r.mx = r.my = r.vx = r.vy = r.cxy = 0.0;
        double dsx = 0.0, dsy = 0.0;
        if (MOMENTS) {
            r.s = warp_sum((double)s);
            dsx = warp_sum((double)sx);
            dsy = warp_sum((double)sy);
        }
        if (MOMENTS && r.s != 0.0) {
            r.mx = dsx / r.s;
            r.my = dsy / r.s;
            const float fmx = (float)r.mx, fmy = (float)r.my;
            // the four column offsets of a float4 as two packed pairs per k % 3 (12 registers instead of three adds per float4);
            // products, their sums and the second x-moment in packed FMUL2 / FADD2 / FFMA2: 14 instead of 22 issue slots per float4.
            // Each product dx_i v_i is still formed on its own, so a one-pixel map keeps its exactly zero variance.
            float2 dXa[3], dXb[3];
            float dY[3];
#pragma unroll
            for (int q = 0; q < 3; ++q) {
                const float d0 = Xr[q] - fmx;
                dXa[q] = make_float2(d0, d0 + 1.f);
                dXb[q] = make_float2(d0 + 2.f, d0 + 3.f);
                dY[q] = Yr[q] - fmy;
            }
            float2 sxx2 = make_float2(0.f, 0.f);
            float syy = 0.f, sxy = 0.f, sdx = 0.f, sdy = 0.f;
#pragma unroll
            for (int k = 0; k < NK; ++k) {
                const float dy = dY[k % 3] + (float)(8 * (k / 3));
                const float2 lo = make_float2(v[k].x, v[k].y), hi = make_float2(v[k].z, v[k].w);
                const float2 pa = __fmul2_rn(dXa[k % 3], lo), pb = __fmul2_rn(dXb[k % 3], hi);
                const float2 ps = __fadd2_rn(pa, pb), qs = __fadd2_rn(lo, hi);
                const float m1 = ps.x + ps.y, q = qs.x + qs.y;
                sxx2 = __ffma2_rn(pa, dXa[k % 3], sxx2);
                sxx2 = __ffma2_rn(pb, dXb[k % 3], sxx2);
                sdx += m1;
                sxy = fmaf(dy, m1, sxy);
                const float dq = dy * q;
                syy = fmaf(dy, dq, syy);
                sdy += dq;
            }
            const float sxx = sxx2.x + sxx2.y;
            const double inv = 1.0 / r.s;
            const double ex = warp_sum((double)sdx) * inv, ey = warp_sum((double)sdy) * inv;
            r.vx = warp_sum((double)sxx) * inv - ex * ex;
            r.vy = warp_sum((double)syy) * inv - ey * ey;
            r.cxy = warp_sum((double)sxy) * inv - ex * ey;
        }
        if (moments && lane < 6) {
            const double val = lane == 0 ? r.mx : lane == 1 ? r.my : lane == 2 ? r.vx : lane == 5 ? r.vy : r.cxy;
            moments[map * 6 + lane] = val;
        }
        if (kpt) {
            // the element inside the winning float4 (first maximum) and its four neighbours, from the lanes that fetched them
            const float q0 = __shfl_sync(0xffffffffu, nbv, 0), q1 = __shfl_sync(0xffffffffu, nbv, 1);
            const float q2 = __shfl_sync(0xffffffffu, nbv, 2), q3 = __shfl_sync(0xffffffffu, nbv, 3);
            const int i = q0 == best ? 0 : q1 == best ? 1 : q2 == best ? 2 : 3;
            const float up = __shfl_sync(0xffffffffu, nbv, 4 + i), down = __shfl_sync(0xffffffffu, nbv, 8 + i);
            const float l12 = __shfl_sync(0xffffffffu, nbv, 12), r13 = __shfl_sync(0xffffffffu, nbv, 13);
            const float left = i == 0 ? l12 : i == 1 ? q0 : i == 2 ? q1 : q2;
            const float right = i == 0 ? q1 : i == 1 ? q2 : i == 2 ? q3 : r13;
            const float a0 = __shfl_sync(0xffffffffu, affv, 16), a1 = __shfl_sync(0xffffffffu, affv, 17);
            const float a2 = __shfl_sync(0xffffffffu, affv, 18), a3 = __shfl_sync(0xffffffffu, affv, 19);
            if (lane == 0) {
                float kx = -1.f, ky = -1.f;
                if (best > 0.f) {                               // mmpose: maxima <= 0 are "not found" (-1, -1)
                    const int e = r.best_idx + i;
                    const int py = e / W, px = e - py * W;
                    kx = (float)px;
                    ky = (float)py;
                    if (px > 1 && px < W - 1 && py > 1 && py < 64 - 1) {
                        const float dx = right - left, dy = down - up;
                        kx += dx > 0.f ? 0.25f : (dx < 0.f ? -0.25f : 0.f);
                        ky += dy > 0.f ? 0.25f : (dy < 0.f ? -0.25f : 0.f);
                    }
                    if (p.has_affine) { kx = fmaf(kx, a0, a2); ky = fmaf(ky, a1, a3); }
                }
                store_keypoint(kx, ky, best, map, p, kpt);
            }
        }
        if (MOMENTS && p.write_back) {                       // upstream's in-place thresholding, after the raw reads
            __syncwarp();
            float4 *dst = reinterpret_cast<float4 *>(hm_wb + map * HW) + lane;
#pragma unroll
            for (int k = 0; k < NK; ++k) dst[32 * k] = v[k];
        }
    }
}

// TMA path: W % 4 == 0, map bytes % 16 == 0, base 16-byte aligned.
__global__ void __launch_bounds__(256, 1)
decode_tma_kernel(const float *__restrict__ hm, float *__restrict__ hm_wb, long long n_maps, int warps_per_cta,
                  int n_stages, const float *__restrict__ affine, float *__restrict__ kpt,
                  double *__restrict__ moments, const __grid_constant__ DecodeParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int HW = p.H * p.W;
    const uint32_t map_bytes = (uint32_t)HW * 4u;
    float *ring = reinterpret_cast<float *>(smem_raw) + (size_t)warp * n_stages * HW;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw + (size_t)warps_per_cta * n_stages * map_bytes) + warp * 8;
    if (lane == 0) {
        for (int s = 0; s < n_stages; ++s) mbar_init(&bars[s], 1);
        fence_mbar_init();
    }
    __syncwarp();
    const long long gwarp = (long long)blockIdx.x * warps_per_cta + warp;
    const long long stride = (long long)gridDim.x * warps_per_cta;
    const long long mine = gwarp < n_maps ? (n_maps - gwarp + stride - 1) / stride : 0;

    auto issue = [&](long long k) {                      // lane 0
        if (k < mine) {
            const int s = (int)(k % n_stages);
            mbar_arrive_expect_tx(&bars[s], map_bytes);
            bulk_g2s(ring + (size_t)s * HW, hm + (gwarp + k * stride) * HW, map_bytes, &bars[s]);
        }
    };
    if (lane == 0)
        for (int k = 0; k < n_stages - 1; ++k) issue(k);
    for (long long k = 0; k < mine; ++k) {
        const long long map = gwarp + k * stride;
        const int s = (int)(k % n_stages);
        if (lane == 0) issue(k + n_stages - 1);          // the stage read in iteration k-1 (guarded by __syncwarp below)
        mbar_wait(&bars[s], (uint32_t)((k / n_stages) & 1));
        const float *src = ring + (size_t)s * HW;
        const MapStats r = map_stats<true>(src, p.write_back ? hm_wb + map * HW : nullptr, p.H, p.W, p.thr, lane);
        write_outputs(r, src, map, p, affine, kpt, moments, lane);
        __syncwarp();
    }
}

// Generic path (any W, any alignment): two passes straight from global memory (the second hits L1/L2).
__global__ void __launch_bounds__(256)
decode_generic_kernel(const float *__restrict__ hm, float *__restrict__ hm_wb, long long n_maps,
                      const float *__restrict__ affine, float *__restrict__ kpt, double *__restrict__ moments,
                      const __grid_constant__ DecodeParams p) {
    const int lane = threadIdx.x & 31;
    const long long gwarp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long stride = ((long long)gridDim.x * blockDim.x) >> 5;
    const int HW = p.H * p.W;
    for (long long map = gwarp; map < n_maps; map += stride) {
        const float *src = hm + map * HW;
        const MapStats r = map_stats<false>(src, nullptr, p.H, p.W, p.thr, lane);
        write_outputs(r, src, map, p, affine, kpt, moments, lane);
        if (p.write_back) {                              // after the neighbour reads of write_outputs
            __syncwarp();
            for (int e = lane; e < HW; e += 32) {
                const float v = src[e];
                hm_wb[map * HW + e] = v < p.thr ? 0.f : v;
            }
        }
    }
}

int decode_device(const float *d_hm, long long n_maps, int H, int W, float thr, int flags, int kpt_layout, int views,
                  int joints, const float *d_affine, int affine_group, float *d_kpt, double *d_moments,
                  cudaStream_t stream) {
    if (n_maps < 0 || H <= 0 || W <= 0) { set_error("bad heatmap shape n=%lld H=%d W=%d", n_maps, H, W); return MC3D_ERR_INVALID_ARGUMENT; }
    if ((long long)H * W > (1 << 22)) { set_error("heatmap %dx%d too large", H, W); return MC3D_ERR_UNSUPPORTED; }
    if (kpt_layout != MC3D_KPT_PLAIN && kpt_layout != MC3D_KPT_NV3 && kpt_layout != MC3D_KPT_N3V) {
        set_error("bad kpt_layout %d", kpt_layout);
        return MC3D_ERR_INVALID_ARGUMENT;
    }
    if (kpt_layout != MC3D_KPT_PLAIN && (views <= 0 || joints <= 0 || (long long)views * joints > (1LL << 30) ||
                                         n_maps % ((long long)views * joints) != 0)) {
        set_error("transposing layouts need views, joints > 0 and n_maps %% (views*joints) == 0");
        return MC3D_ERR_INVALID_ARGUMENT;
    }
    if (d_affine && affine_group <= 0) { set_error("affine_group must be > 0"); return MC3D_ERR_INVALID_ARGUMENT; }
    if (n_maps == 0) return MC3D_OK;
    if (!d_hm || (!d_kpt && !d_moments)) { set_error("NULL heatmap pointer or no output requested"); return MC3D_ERR_INVALID_ARGUMENT; }
    DecodeParams p;
    p.H = H; p.W = W; p.thr = thr;
    p.write_back = (flags & MC3D_DECODE_FLAG_WRITE_BACK) ? 1 : 0;
    p.kpt_layout = kpt_layout; p.views = views; p.joints = joints;
    p.has_affine = d_affine != nullptr; p.affine_group = affine_group > 0 ? affine_group : 1;
    float *wb = const_cast<float *>(d_hm);
    const size_t map_bytes = (size_t)H * W * 4;
    const bool tma_ok = (W % 4 == 0) && aligned16(d_hm) && map_bytes <= 96 * 1024 && !(flags & MC3D_DECODE_FLAG_GENERIC);
    const bool reg_ok = H == 64 && W == 48 && aligned16(d_hm) && !(flags & (MC3D_DECODE_FLAG_GENERIC | MC3D_DECODE_FLAG_TMA));
    if (reg_ok) {
        const bool staged = (d_moments || p.write_back) && !(flags & MC3D_DECODE_FLAG_NO_STAGE);
        long long grid = (long long)sm_count() * (staged ? MC3D_DEC_STAGED_BLOCKS : 4);   // persistent: one wave of resident CTAs
        const long long need = (n_maps + 3) / 4;
        if (grid > need) grid = need;
        if ((d_moments || p.write_back) && !(flags & MC3D_DECODE_FLAG_NO_STAGE)) {
            auto kern = decode_reg6448_kernel<true, true>;
            const size_t smem = 4 * MC3D_DEC_STAGES * map_bytes;    // the warps' stages
            { const int as = func_max_smem_once((const void *)kern, 112 * 1024); if (as != MC3D_OK) return as; }
            kern<<<(unsigned)grid, 128, smem, stream>>>(d_hm, wb, n_maps, d_affine, d_kpt, d_moments, p);
        } else if (d_moments || p.write_back)
            decode_reg6448_kernel<true><<<(unsigned)grid, 128, 0, stream>>>(d_hm, wb, n_maps, d_affine, d_kpt, d_moments, p);
        else
            decode_reg6448_kernel<false><<<(unsigned)grid, 128, 0, stream>>>(d_hm, wb, n_maps, d_affine, d_kpt, nullptr, p);
    } else if (tma_ok) {
        // choose warps x stages to keep as many bytes in flight as fit in ~200 KB of shared memory
        // The two passes are issue-bound, not latency-bound: warps (instruction streams) matter more than ring depth.
        // Two stages per warp already overlap the next map's copy with the current map's arithmetic.
        int warps = 8, stages = 2;
        const size_t budget = 216 * 1024;
        while (warps > 1 && (size_t)warps * stages * map_bytes > budget) warps >>= 1;
        while ((size_t)warps * (stages + 1) * map_bytes <= budget && stages < 4) ++stages;
        const size_t smem = (size_t)warps * stages * map_bytes + (size_t)warps * 8 * sizeof(uint64_t);
        { const int as = func_max_smem_once((const void *)decode_tma_kernel, 227 * 1024); if (as != MC3D_OK) return as; }
        long long grid = sm_count();
        const long long need = (n_maps + warps - 1) / warps;
        if (grid > need) grid = need;
        decode_tma_kernel<<<(unsigned)grid, warps * 32, smem, stream>>>(d_hm, wb, n_maps, warps, stages, d_affine, d_kpt,
                                                                       d_moments, p);
    } else {
        long long grid = (long long)sm_count() * 8;
        const long long need = (n_maps + 7) / 8;
        if (grid > need) grid = need;
        decode_generic_kernel<<<(unsigned)grid, 256, 0, stream>>>(d_hm, wb, n_maps, d_affine, d_kpt, d_moments, p);
    }
    count_launch();
    MC3D_CUDA_TRY(cudaGetLastError());
    return MC3D_OK;
}

}  // namespace mc3d

extern "C" int mc3d_decode_heatmaps_f32(const float *d_heatmaps, int64_t n_maps, int H, int W, float threshold,
                                        int flags, int kpt_layout, int views, int joints, const float *d_affine,
                                        int affine_group, float *d_kpt, double *d_moments, void *stream) {
    return mc3d::decode_device(d_heatmaps, n_maps, H, W, threshold, flags, kpt_layout, views, joints, d_affine,
                               affine_group, d_kpt, d_moments, (cudaStream_t)stream);
}
