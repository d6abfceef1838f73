// Learning one camera's extrinsics from sampled 3D points, for sm_100a: the optimize_trajectory=False /
// extrinsic_optimization_IDs=[id] branch of the reference's sgd_optimize (pose_refinement.py:915-1091).
//
// The samples (N per frame and joint, drawn from the two ground-truth cameras' Gaussians and triangulated,
// pose_refinement.py:684-706, :811) are fixed; each iteration evaluates
//     cost(R, T) = mean over finite samples of 0.5 d^T S d,   d = pi(R X + T) - mean[t, j]          (:800-831)
// and its gradient with respect to the 9 entries of R (upstream optimises the live 3x3 matrix entry-wise) and the 3 of
// T, then clip_grad_norm_(1.0) + Adam on those 12 numbers and the same running-mean early stopping as the trajectory
// optimiser.  Two kernels per iteration: a grid-stride pass over the samples (14 sums: cost, count, dR, dT -- the
// gradient of a sample w.r.t. R is the outer product of its camera-frame gradient with the point) and a one-warp
// step kernel; whole runs replay from a CUDA graph.
#include "mc3d_common.cuh"
#include <math.h>

namespace mc3d {

constexpr int EX_THREADS = 256;
constexpr int EX_ACC = 0;        // + 16 * parity : cost_sum count dR[9] dT[3]
constexpr int EX_STATE = 32;     // + 16 * parity : step run_sum run_cnt best no_improve stopped iters_done improved
constexpr int EX_HIST = 64;      // + 2 * step    : sample cost, total cost

template <typename C> __device__ __forceinline__ bool finite_x(C v);
template <> __device__ __forceinline__ bool finite_x<float>(float v) { return fabsf(v) <= 3.0e38f; }
template <> __device__ __forceinline__ bool finite_x<double>(double v) { return fabs(v) <= 1.0e300; }

// 0.5 d^T S d for one point in the camera frame, and the gradient w.r.t. that camera-frame point
// (project_points_torch, pose_refinement.py:134-174, differentiated by hand).
template <typename C>
__device__ __forceinline__ C sample_term(const C *K, const C *D, bool ignore_dist, C xc, C yc, C zc, C mx, C my, C s00, C s01,
                                         C s11, C (&g)[3]) {
    const C one = (C)1, two = (C)2;
    const C iz = one / zc;
    const C a = xc * iz, b = yc * iz;
    C xd = a, yd = b, j00 = one, j01 = (C)0, j11 = one;
    if (!ignore_dist) {
        const C k1 = D[0], k2 = D[1], p1 = D[2], p2 = D[3], k3 = D[4];
        const C r2 = fma(a, a, b * b);
        const C rad = fma(fma(fma(k3, r2, k2), r2, k1), r2, one);
        const C drad = fma(fma((C)3 * k3, r2, two * k2), r2, k1);
        xd = fma(a, rad, fma(two * p1 * a, b, p2 * fma(two * a, a, r2)));
        yd = fma(b, rad, fma(p1, fma(two * b, b, r2), two * p2 * a * b));
        j00 = rad + two * a * a * drad + two * p1 * b + (C)6 * p2 * a;
        j01 = two * a * b * drad + two * p1 * a + two * p2 * b;
        j11 = rad + two * b * b * drad + (C)6 * p1 * b + two * p2 * a;
    }
    const C u = fma(K[0], xd, fma(K[1], yd, K[2]));
    const C v = fma(K[3], xd, fma(K[4], yd, K[5]));
    const C is = one / fma(K[6], xd, fma(K[7], yd, K[8]));
    const C px = u * is, py = v * is;
    const C dx = px - mx, dy = py - my;
    const C sdx = fma(s00, dx, s01 * dy), sdy = fma(s01, dx, s11 * dy);
    const C gxd = (sdx * (K[0] - px * K[6]) + sdy * (K[3] - py * K[6])) * is;
    const C gyd = (sdx * (K[1] - px * K[7]) + sdy * (K[4] - py * K[7])) * is;
    const C ga = fma(j00, gxd, j01 * gyd), gb = fma(j01, gxd, j11 * gyd);
    g[0] = ga * iz;
    g[1] = gb * iz;
    g[2] = -(a * ga + b * gb) * iz;
    return (C)0.5 * fma(dx, sdx, dy * sdy);
}

template <typename T>
__global__ void __launch_bounds__(EX_THREADS)
extrinsic_costgrad_kernel(const __grid_constant__ mc3d_extrinsic_problem pb, int parity) {
    __shared__ double red[8 * 14];
    double *ctrl = pb.ctrl;
    if (ctrl[EX_STATE + 16 * parity + 5] != 0.0) return;          // stopped
    T R[9], Tv[3], K[9], D[5];
#pragma unroll
    for (int i = 0; i < 9; ++i) { R[i] = (T)pb.params[i]; K[i] = (T)pb.K[i]; }
#pragma unroll
    for (int i = 0; i < 3; ++i) Tv[i] = (T)pb.params[9 + i];
#pragma unroll
    for (int i = 0; i < 5; ++i) D[i] = (T)pb.dist[i];
    const T *X = (const T *)pb.samples3d, *mean = (const T *)pb.mean, *S = (const T *)pb.S;
    const long long n = pb.n_frames * pb.n_joints * (long long)pb.n_samples;
    const bool ign = pb.ignore_distortions != 0;
    double acc[14];
#pragma unroll
    for (int i = 0; i < 14; ++i) acc[i] = 0.0;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
        const long long tj = e / pb.n_samples;
        const T x0 = X[e * 3], x1 = X[e * 3 + 1], x2 = X[e * 3 + 2];
        const T xc = fma(R[0], x0, fma(R[1], x1, fma(R[2], x2, Tv[0])));
        const T yc = fma(R[3], x0, fma(R[4], x1, fma(R[5], x2, Tv[1])));
        const T zc = fma(R[6], x0, fma(R[7], x1, fma(R[8], x2, Tv[2])));
        T g[3];
        const T q = sample_term<T>(K, D, ign, xc, yc, zc, mean[tj * 2], mean[tj * 2 + 1], S[tj * 3], S[tj * 3 + 1], S[tj * 3 + 2], g);
        if (finite_x(q)) {                                       // nan_mean: non-finite samples drop out of cost and gradient
            acc[0] += (double)q;
            acc[1] += 1.0;
            const bool gok = finite_x(g[0]) && finite_x(g[1]) && finite_x(g[2]);
            if (gok) {
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    acc[2 + 3 * i + 0] += (double)(g[i] * x0);
                    acc[2 + 3 * i + 1] += (double)(g[i] * x1);
                    acc[2 + 3 * i + 2] += (double)(g[i] * x2);
                    acc[11 + i] += (double)g[i];
                }
            }
        }
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < 14; ++i) acc[i] = warp_sum(acc[i]);
    if (lane == 0)
#pragma unroll
        for (int i = 0; i < 14; ++i) red[warp * 14 + i] = acc[i];
    __syncthreads();
    if (threadIdx.x < 14) {
        double s = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += red[w * 14 + threadIdx.x];
        if (s != 0.0) atomicAdd(ctrl + EX_ACC + 16 * parity + threadIdx.x, s);
    }
}

// One warp: gradient mean, clip_grad_norm_(1.0) over the 12 entries, Adam, running-mean early stopping (Q5), best snapshot.
template <typename T>
__global__ void extrinsic_step_kernel(const __grid_constant__ mc3d_extrinsic_problem pb, int parity) {
    double *ctrl = pb.ctrl;
    double *acc = ctrl + EX_ACC + 16 * parity, *st = ctrl + EX_STATE + 16 * parity, *nx = ctrl + EX_STATE + 16 * (parity ^ 1);
    const int i = threadIdx.x;
    if (st[5] != 0.0) {
        if (i < 16) nx[i] = st[i];
        if (i < 16) ctrl[EX_ACC + 16 * (parity ^ 1) + i] = 0.0;
        return;
    }
    const double count = acc[1];
    const double cost = acc[0] / count, total = cost + pb.const_cost;
    const double gi = i < 12 ? acc[2 + i] / count : 0.0;
    const double gnorm = sqrt(warp_sum(gi * gi));
    const double clip = fmin(1.0, 1.0 / (gnorm + 1e-6));
    const double step = st[0] + 1.0;
    double run_sum = st[1] + total, run_cnt = st[2] + 1.0, best = st[3], no_imp = st[4], iters = st[6];
    const double mean = run_sum / run_cnt;                         // running mean over costs AND earlier means (Q5)
    run_sum += mean; run_cnt += 1.0;
    const bool improved = mean < best - pb.tolerance;
    if (improved) { best = mean; no_imp = 0.0; } else { no_imp += 1.0; }
    iters += 1.0;
    const bool stop = (no_imp >= (double)pb.patience) || (iters > (double)pb.max_iter);
    if (i < 12) {
        double *p = pb.params + i, *m = pb.params + 12 + i, *v = pb.params + 24 + i, *b = pb.params + 36 + i;
        const double g = !(clip == clip) ? NAN : gi * clip;
        const double mi = *m + (g - *m) * (1.0 - pb.beta1);
        const double vi = *v * pb.beta2 + (1.0 - pb.beta2) * g * g;
        const double denom = sqrt(vi) / sqrt(1.0 - pow(pb.beta2, step)) + pb.eps;
        const double pn = *p - (pb.lr / (1.0 - pow(pb.beta1, step))) * (mi / denom);
        *p = (double)(T)pn;                                        // the parameters live in the state dtype upstream
        *m = (double)(T)mi;
        *v = (double)(T)vi;
        if (improved) *b = *p;
    }
    if (i < 16) ctrl[EX_ACC + 16 * (parity ^ 1) + i] = 0.0;
    if (i == 0) {
        nx[0] = step; nx[1] = run_sum; nx[2] = run_cnt; nx[3] = best; nx[4] = no_imp;
        nx[5] = stop ? 1.0 : 0.0; nx[6] = iters; nx[7] = improved ? 1.0 : 0.0;
        const long long hs = (long long)(step - 1.0);
        if (hs < pb.hist_capacity) { ctrl[EX_HIST + 2 * hs] = cost; ctrl[EX_HIST + 2 * hs + 1] = total; }
    }
}

template <typename T>
int extrinsic_run(const mc3d_extrinsic_problem *pb, long long first_step, long long n_iters, cudaStream_t stream) {
    if (!pb) { set_error("NULL problem"); return MC3D_ERR_INVALID_ARGUMENT; }
    if (pb->n_frames < 1 || pb->n_joints < 1 || pb->n_samples < 1) { set_error("empty sample set"); return MC3D_ERR_INVALID_ARGUMENT; }
    if (!pb->samples3d || !pb->mean || !pb->S || !pb->params || !pb->ctrl) { set_error("NULL device pointer in extrinsic problem"); return MC3D_ERR_INVALID_ARGUMENT; }
    const long long n = pb->n_frames * pb->n_joints * (long long)pb->n_samples;
    long long grid = (n + EX_THREADS - 1) / EX_THREADS;
    if (grid > (long long)sm_count() * 4) grid = (long long)sm_count() * 4;
    auto one = [&](long long step, cudaStream_t s) -> int {
        const int parity = (int)(step & 1);
        extrinsic_costgrad_kernel<T><<<(unsigned)grid, EX_THREADS, 0, s>>>(*pb, parity);
        extrinsic_step_kernel<T><<<1, 32, 0, s>>>(*pb, parity);
        MC3D_CUDA_TRY(cudaGetLastError());
        return MC3D_OK;
    };
    long long done = 0;
    int st = MC3D_OK;
    if ((first_step & 1) && n_iters > 0) { st = one(first_step, stream); if (st != MC3D_OK) return st; count_launch(2); done = 1; }
    const long long pairs = (n_iters - done) / 2;
    if (pairs >= 4 && !stream_is_capturing(stream)) {          // a capturing caller gets plain launches (they join its capture)
        static thread_local cudaStream_t cap_stream = nullptr;
        if (!cap_stream) MC3D_CUDA_TRY(cudaStreamCreateWithFlags(&cap_stream, cudaStreamNonBlocking));
        cudaGraph_t graph = nullptr;
        cudaGraphExec_t exec = nullptr;
        MC3D_CUDA_TRY(cudaStreamBeginCapture(cap_stream, cudaStreamCaptureModeThreadLocal));
        for (int u = 0; u < 2 && st == MC3D_OK; ++u) st = one(first_step + done + u, cap_stream);
        cudaError_t ce = cudaStreamEndCapture(cap_stream, &graph);
        if (st != MC3D_OK) { if (graph) cudaGraphDestroy(graph); return st; }
        if (ce != cudaSuccess) { if (graph) cudaGraphDestroy(graph); return cuda_fail(ce, "cudaStreamEndCapture"); }
        GraphGuard guard{graph, nullptr};                          // destroys the graph and its exec on every exit path
        MC3D_CUDA_TRY(cudaGraphInstantiate(&exec, graph, 0));
        guard.exec = exec;
        for (long long i = 0; i < pairs; ++i) MC3D_CUDA_TRY(cudaGraphLaunch(exec, stream));
        count_launch((int)(pairs * 4));
        done += pairs * 2;
        MC3D_CUDA_TRY(cudaStreamSynchronize(stream));
    }
    for (; done < n_iters; ++done) { st = one(first_step + done, stream); if (st != MC3D_OK) return st; count_launch(2); }
    return MC3D_OK;
}

// ---- cameras and trajectory learnt together (extrinsic_optimization_IDs with optimize_trajectory=True) ----------------
// The trajectory's step is csrc/refine.cu's three phases; between its gradient phase and its Adam phase this kernel
// (one block, one warp per learnt camera) turns each camera's raw sums -- extrinsic_costgrad_kernel over the
// trajectory points, one "sample" per (frame, joint) -- into its gradient of the likelihood cost (divide by the
// N_lik of ALL cameras, pose_refinement.py:889), adds the cameras' squared norms to the trajectory's so that ONE
// clip_grad_norm_ covers every learnable parameter (:1047), and applies Adam to the cameras with that clip factor.
//   refine_ctrl : the trajectory's control block (sums of this step at [16 p ...]: N_lik at 1, |g_x|^2 at 7; Adam step at 32 + 16 p)
//   cam_ctrl    : per learnt camera a 64-double control block as in mc3d_extrinsic_problem.ctrl; [0..13] = cost, count,
//                 dR[9], dT[3] from extrinsic_costgrad_kernel, zeroed here (the rest stays zero)
//   cam_params  : per learnt camera 36 doubles: R[9] T[3] | m[12] | v[12]
template <typename T>
__global__ void extrinsic_joint_step_kernel(double *refine_ctrl, double *cam_ctrl, double *cam_params, int n_learn, int parity,
                                            double lr, double beta1, double beta2, double eps) {
    __shared__ double sq[MC3D_MAX_VIEWS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double *acc = refine_ctrl + 16 * parity;
    const double *st = refine_ctrl + 32 + 16 * parity;
    if (st[5] != 0.0) return;                                      // stopped
    const double n_lik = acc[1];
    double g = 0.0;
    if (warp < n_learn && lane < 12) g = cam_ctrl[64 * warp + 2 + lane] / n_lik;
    const double s = warp_sum(g * g);
    if (warp < n_learn && lane == 0) sq[warp] = s;
    __syncthreads();
    double total = acc[7];
    for (int c = 0; c < n_learn; ++c) total += sq[c];
    __syncthreads();
    if (threadIdx.x == 0) acc[7] = total;                          // the trajectory's Adam phase clips with the joint norm
    const double clip = fmin(1.0, 1.0 / (sqrt(total) + 1e-6));
    const double step = st[0] + 1.0;
    if (warp < n_learn && lane < 12) {
        double *p = cam_params + 36 * warp + lane, *m = p + 12, *v = p + 24;
        const double gc = !(clip == clip) ? NAN : g * clip;
        const double mi = *m + (gc - *m) * (1.0 - beta1);
        const double vi = *v * beta2 + (1.0 - beta2) * gc * gc;
        const double pn = *p - (lr / (1.0 - pow(beta1, step))) * (mi / (sqrt(vi) / sqrt(1.0 - pow(beta2, step)) + eps));
        *p = (double)(T)pn;
        *m = (double)(T)mi;
        *v = (double)(T)vi;
    }
    if (warp < n_learn && lane < 16) cam_ctrl[64 * warp + lane] = 0.0;
}

template <typename T>
int extrinsic_costgrad(const mc3d_extrinsic_problem *pb, cudaStream_t stream) {
    if (!pb || !pb->samples3d || !pb->mean || !pb->S || !pb->params || !pb->ctrl) { set_error("NULL pointer in extrinsic problem"); return MC3D_ERR_INVALID_ARGUMENT; }
    const long long n = pb->n_frames * pb->n_joints * (long long)pb->n_samples;
    if (n <= 0) return MC3D_OK;
    long long grid = (n + EX_THREADS - 1) / EX_THREADS;
    if (grid > (long long)sm_count() * 4) grid = (long long)sm_count() * 4;
    extrinsic_costgrad_kernel<T><<<(unsigned)grid, EX_THREADS, 0, stream>>>(*pb, 0);
    count_launch();
    MC3D_CUDA_TRY(cudaGetLastError());
    return MC3D_OK;
}

template <typename T>
int extrinsic_joint_step(double *refine_ctrl, double *cam_ctrl, double *cam_params, int n_learn, long long step_index, double lr,
                         double beta1, double beta2, double eps, cudaStream_t stream) {
    if (!refine_ctrl || !cam_ctrl || !cam_params || n_learn < 1 || n_learn > MC3D_MAX_VIEWS) { set_error("bad joint-step arguments"); return MC3D_ERR_INVALID_ARGUMENT; }
    extrinsic_joint_step_kernel<T><<<1, 32 * n_learn, 0, stream>>>(refine_ctrl, cam_ctrl, cam_params, n_learn, (int)(step_index & 1), lr,
                                                                beta1, beta2, eps);
    count_launch();
    MC3D_CUDA_TRY(cudaGetLastError());
    return MC3D_OK;
}

}  // namespace mc3d

extern "C" {
int mc3d_extrinsic_costgrad_f32(const mc3d_extrinsic_problem *pb, void *stream) { return mc3d::extrinsic_costgrad<float>(pb, (cudaStream_t)stream); }
int mc3d_extrinsic_costgrad_f64(const mc3d_extrinsic_problem *pb, void *stream) { return mc3d::extrinsic_costgrad<double>(pb, (cudaStream_t)stream); }
int mc3d_extrinsic_joint_step_f32(double *refine_ctrl, double *cam_ctrl, double *cam_params, int n_learn, int64_t step_index, double lr,
                                  double beta1, double beta2, double eps, void *stream) {
    return mc3d::extrinsic_joint_step<float>(refine_ctrl, cam_ctrl, cam_params, n_learn, step_index, lr, beta1, beta2, eps, (cudaStream_t)stream);
}
int mc3d_extrinsic_joint_step_f64(double *refine_ctrl, double *cam_ctrl, double *cam_params, int n_learn, int64_t step_index, double lr,
                                  double beta1, double beta2, double eps, void *stream) {
    return mc3d::extrinsic_joint_step<double>(refine_ctrl, cam_ctrl, cam_params, n_learn, step_index, lr, beta1, beta2, eps, (cudaStream_t)stream);
}
int mc3d_extrinsic_problem_size(void) { return (int)sizeof(mc3d_extrinsic_problem); }
int mc3d_extrinsic_run_f32(const mc3d_extrinsic_problem *pb, int64_t first_step, int64_t n_iters, void *stream) {
    return mc3d::extrinsic_run<float>(pb, first_step, n_iters, (cudaStream_t)stream);
}
int mc3d_extrinsic_run_f64(const mc3d_extrinsic_problem *pb, int64_t first_step, int64_t n_iters, void *stream) {
    return mc3d::extrinsic_run<double>(pb, first_step, n_iters, (cudaStream_t)stream);
}
}
