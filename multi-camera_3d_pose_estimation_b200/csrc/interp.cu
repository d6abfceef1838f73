// Windowed outlier filter + local line fit of a trajectory (reference pose_refinement.py:15-84,
// `linear_interpolation`): the default refinement of the reference CLI and the optional initialiser of the SGD.
// One thread per (frame, joint, dim) scalar; the <= 33-sample window is read with frame stride (coalesced across
// the threads of a warp, which hold consecutive (joint, dim) scalars) and kept in local arrays.
#include "mc3d_common.cuh"
#include <math.h>

namespace mc3d {

constexpr int INTERP_MAX_WINDOW = 65;

// numpy's pairwise summation for n < 128 (the order np.mean / np.std use), so that the <= comparisons of the
// outlier test see the same rounding as upstream.
__device__ __forceinline__ double np_sum(const double *w, int n) {
    if (n < 8) {
        double s = 0.0;
        for (int i = 0; i < n; ++i) s += w[i];
        return s;
    }
    double r[8];
    for (int j = 0; j < 8; ++j) r[j] = w[j];
    int i = 8;
    for (; i < n - (n % 8); i += 8)
        for (int j = 0; j < 8; ++j) r[j] += w[i + j];
    double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
    for (; i < n; ++i) res += w[i];
    return res;
}

__device__ __forceinline__ double median_of(double *s, int n) {      // sorts s in place
    for (int i = 1; i < n; ++i) {
        const double v = s[i];
        int j = i - 1;
        while (j >= 0 && s[j] > v) { s[j + 1] = s[j]; --j; }
        s[j + 1] = v;
    }
    return (n & 1) ? s[n / 2] : 0.5 * (s[n / 2 - 1] + s[n / 2]);
}

__global__ void __launch_bounds__(128)
interp_kernel(const double *__restrict__ pts, double *__restrict__ out, long long T, long long PD, int k, double k_std,
              double median_std, int rolling, int filter_median) {
    const long long total = T * PD;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const long long t = idx / PD, c = idx - t * PD;
        const long long lo = (t - k / 2 > 0) ? t - k / 2 : 0;
        const long long hi = (t + k / 2 + 1 < T) ? t + k / 2 + 1 : T;
        const int n = (int)(hi - lo);
        double w[INTERP_MAX_WINDOW], tmp[INTERP_MAX_WINDOW];
        for (int i = 0; i < n; ++i) w[i] = pts[(lo + i) * PD + c];
        const double mean = np_sum(w, n) / n;
        for (int i = 0; i < n; ++i) { const double d = w[i] - mean; tmp[i] = d * d; }
        const double sd = sqrt(np_sum(tmp, n) / n);
        for (int i = 0; i < n; ++i) tmp[i] = w[i];
        const double med = median_of(tmp, n);
        for (int i = 0; i < n; ++i) tmp[i] = fabs(w[i] - med);
        const double mad = median_of(tmp, n);
        int cnt = 0;
        double vals[INTERP_MAX_WINDOW], times[INTERP_MAX_WINDOW];
        for (int i = 0; i < n; ++i) {
            bool ok = fabs(w[i] - mean) <= k_std * sd;
            if (filter_median) ok = ok && (fabs(w[i] - med) <= median_std * mad);
            if (ok) { vals[cnt] = w[i]; times[cnt] = (double)(lo + i); ++cnt; }
        }
        double res = 0.0;                                   // upstream leaves 0 when fewer than two samples survive
        if (cnt >= 2) {
            const double vm = np_sum(vals, cnt) / cnt;
            if (rolling) {
                res = vm;
            } else {
                const double tm = np_sum(times, cnt) / cnt;
                double sxx = 0.0, sxy = 0.0;
                for (int i = 0; i < cnt; ++i) { const double dt = times[i] - tm; sxx += dt * dt; sxy += dt * (vals[i] - vm); }
                res = vm + (sxy / sxx) * ((double)t - tm);
            }
        }
        out[idx] = res;
    }
}

int interp_device(const double *d_pts, long long T, long long PD, int k, double k_std, double median_std, int rolling,
                  int filter_median, double *d_out, cudaStream_t stream) {
    if (T < 0 || PD < 0 || k < 0) { set_error("bad shape"); return MC3D_ERR_INVALID_ARGUMENT; }
    if (2 * (k / 2) + 1 > INTERP_MAX_WINDOW) { set_error("window k=%d too large (max %d)", k, INTERP_MAX_WINDOW - 1); return MC3D_ERR_UNSUPPORTED; }
    if (T * PD == 0) return MC3D_OK;
    if (!d_pts || !d_out) { set_error("NULL device pointer"); return MC3D_ERR_INVALID_ARGUMENT; }
    long long grid = (T * PD + 127) / 128;
    if (grid > (long long)sm_count() * 16) grid = (long long)sm_count() * 16;
    interp_kernel<<<(unsigned)grid, 128, 0, stream>>>(d_pts, d_out, T, PD, k, k_std, median_std, rolling, filter_median);
    count_launch();
    MC3D_CUDA_TRY(cudaGetLastError());
    return MC3D_OK;
}

}  // namespace mc3d

extern "C" int mc3d_linear_interpolation_f64(const double *d_points, int64_t n_frames, int64_t scalars_per_frame, int k,
                                             double k_std, double median_std, int use_rolling_average,
                                             int filter_distance_from_median, double *d_out, void *stream) {
    return mc3d::interp_device(d_points, n_frames, scalars_per_frame, k, k_std, median_std, use_rolling_average,
                               filter_distance_from_median, d_out, (cudaStream_t)stream);
}
