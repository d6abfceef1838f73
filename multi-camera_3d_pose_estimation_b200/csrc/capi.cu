// Library-level C ABI: errors, device info, launch counter, and the host-buffer pipelines.
#include "mc3d_common.cuh"
#include <atomic>
#include <mutex>
#include <set>
#include <utility>
#include <string.h>

namespace mc3d {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char *what) {
    set_error("CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
    return e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver ? MC3D_ERR_NO_DEVICE : MC3D_ERR_CUDA;
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int sm_count() {
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (cached[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}

int current_device_slot() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 0;
    return dev;
}

int func_max_smem_once(const void *func, int bytes) {
    static std::mutex mu;
    static std::set<std::pair<int, const void *>> done;
    const int dev = current_device_slot();
    std::lock_guard<std::mutex> lock(mu);
    if (done.count({dev, func})) return MC3D_OK;
    MC3D_CUDA_TRY(cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    done.insert({dev, func});
    return MC3D_OK;
}

template <typename T>
int triangulate_device(const T *d_kpts, long long n, const mc3d_rig *rig, int layout, int mode, int flags, T *d_out,
                       cudaStream_t stream);
int decode_device(const float *d_hm, long long n_maps, int H, int W, float thr, int flags, int kpt_layout, int views,
                  int joints, const float *d_affine, int affine_group, float *d_kpt, double *d_moments,
                  cudaStream_t stream);

// ---- host-buffer pipeline ------------------------------------------------------------------------
// Three in-flight chunks, one stream each: H2D(i+1) and D2H(i-1) overlap the kernel of chunk i
// (PCIe is full duplex, the copy engines run beside the SMs).
namespace {
constexpr int NBUF = 3;
constexpr int MAX_DEVICES = 64;
struct HostPipe {
    bool ready = false;
    cudaStream_t stream[NBUF] = {nullptr, nullptr, nullptr};
    void *d_in[NBUF] = {nullptr, nullptr, nullptr};
    void *d_out[NBUF] = {nullptr, nullptr, nullptr};
    size_t in_bytes = 0, out_bytes = 0;
};
std::mutex g_pipe_mutex;
HostPipe g_pipes[MAX_DEVICES];          // one staging pipeline per device ordinal, grown on demand, kept for the process

// The host-buffer entry points work on the device the CALLER names, and hand the calling thread back on the device it
// was on: under torchrun every rank has its own current device, and a library call must not move it.
struct DeviceGuard {
    int prev = -1;
    bool moved = false;
    int enter(int device) {
        MC3D_CUDA_TRY(cudaGetDevice(&prev));
        if (prev != device) {
            MC3D_CUDA_TRY(cudaSetDevice(device));
            moved = true;
        }
        return MC3D_OK;
    }
    ~DeviceGuard() {
        if (moved) cudaSetDevice(prev);
    }
};

int ensure_pipe(int device, size_t in_bytes, size_t out_bytes, HostPipe **out) {
    HostPipe &pipe = g_pipes[device];
    if (!pipe.ready) {
        for (int b = 0; b < NBUF; ++b)
            if (!pipe.stream[b]) MC3D_CUDA_TRY(cudaStreamCreateWithFlags(&pipe.stream[b], cudaStreamNonBlocking));
        pipe.ready = true;
    }
    if (in_bytes > pipe.in_bytes) {
        pipe.in_bytes = 0;
        for (int b = 0; b < NBUF; ++b) {
            if (pipe.d_in[b]) MC3D_CUDA_TRY(cudaFree(pipe.d_in[b]));
            pipe.d_in[b] = nullptr;
            MC3D_CUDA_TRY(cudaMalloc(&pipe.d_in[b], in_bytes));
        }
        pipe.in_bytes = in_bytes;
    }
    if (out_bytes > pipe.out_bytes) {
        pipe.out_bytes = 0;
        for (int b = 0; b < NBUF; ++b) {
            if (pipe.d_out[b]) MC3D_CUDA_TRY(cudaFree(pipe.d_out[b]));
            pipe.d_out[b] = nullptr;
            MC3D_CUDA_TRY(cudaMalloc(&pipe.d_out[b], out_bytes));
        }
        pipe.out_bytes = out_bytes;
    }
    *out = &pipe;
    return MC3D_OK;
}

// Wait for everything queued on the pipeline.  Called on every exit path: the copies read and write the CALLER's host
// buffers, so nothing may still be in flight when the entry point returns, error or not.
int drain_pipe(HostPipe &pipe, int status) {
    for (int i = 0; i < NBUF; ++i) {
        const cudaError_t e = cudaStreamSynchronize(pipe.stream[i]);
        if (e != cudaSuccess && status == MC3D_OK) status = cuda_fail(e, "cudaStreamSynchronize (host pipeline)");
    }
    return status;
}
}  // namespace

template <typename T>
static int triangulate_host_chunks(HostPipe &pipe, const T *h_kpts, long long n, long long chunk, const mc3d_rig *rig,
                                   int layout, int mode, int flags, T *h_out) {
    const size_t row_bytes = (size_t)3 * rig->n_views * sizeof(T);
    int b = 0;
    for (long long off = 0; off < n; off += chunk, b = (b + 1) % NBUF) {
        const long long m = (n - off < chunk) ? (n - off) : chunk;
        cudaStream_t s = pipe.stream[b];           // chunk i + NBUF re-uses buffer b in stream order
        MC3D_CUDA_TRY(cudaMemcpyAsync(pipe.d_in[b], h_kpts + off * 3 * rig->n_views, (size_t)m * row_bytes,
                                      cudaMemcpyHostToDevice, s));
        const int st = triangulate_device<T>((const T *)pipe.d_in[b], m, rig, layout, mode, flags, (T *)pipe.d_out[b], s);
        if (st != MC3D_OK) return st;
        MC3D_CUDA_TRY(cudaMemcpyAsync(h_out + off * 3, pipe.d_out[b], (size_t)m * 3 * sizeof(T), cudaMemcpyDeviceToHost, s));
    }
    return MC3D_OK;
}

template <typename T>
int triangulate_host(const T *h_kpts, long long n, const mc3d_rig *rig, int layout, int mode, int flags, T *h_out,
                     int device) {
    if (n < 0 || !rig) { set_error("bad arguments"); return MC3D_ERR_INVALID_ARGUMENT; }
    if (n == 0) return MC3D_OK;
    if (!h_kpts || !h_out) { set_error("NULL host pointer"); return MC3D_ERR_INVALID_ARGUMENT; }
    if (rig->n_views < 2 || rig->n_views > MC3D_MAX_VIEWS) { set_error("n_views=%d out of range", rig->n_views); return MC3D_ERR_INVALID_ARGUMENT; }
    if (device < 0 || device >= MAX_DEVICES) { set_error("device ordinal %d out of range", device); return MC3D_ERR_INVALID_ARGUMENT; }
    std::lock_guard<std::mutex> lock(g_pipe_mutex);
    DeviceGuard guard;                                                      // restores the caller's device on every exit path
    { const int gs = guard.enter(device); if (gs != MC3D_OK) return gs; }
    const size_t row_bytes = (size_t)3 * rig->n_views * sizeof(T);
    long long chunk = (long long)((64u << 20) / row_bytes) / 256 * 256;     // ~64 MiB of keypoints per chunk
    if (chunk > n) chunk = (n + 255) / 256 * 256;
    HostPipe *pipe = nullptr;
    const int st = ensure_pipe(device, (size_t)chunk * row_bytes, (size_t)chunk * 3 * sizeof(T), &pipe);
    if (st != MC3D_OK) return st;
    return drain_pipe(*pipe, triangulate_host_chunks<T>(*pipe, h_kpts, n, chunk, rig, layout, mode, flags, h_out));
}

static int decode_host_chunks(HostPipe &pipe, const float *h_hm, long long n_maps, long long chunk, size_t kpt_bytes, int H,
                              int W, float thr, float *h_kpt, double *h_moments) {
    const size_t map_bytes = (size_t)H * W * sizeof(float);
    int b = 0;
    for (long long off = 0; off < n_maps; off += chunk, b = (b + 1) % NBUF) {
        const long long m = (n_maps - off < chunk) ? (n_maps - off) : chunk;
        cudaStream_t s = pipe.stream[b];
        float *d_kpt = (float *)pipe.d_out[b];
        double *d_mom = (double *)((char *)pipe.d_out[b] + kpt_bytes);
        MC3D_CUDA_TRY(cudaMemcpyAsync(pipe.d_in[b], h_hm + off * H * W, (size_t)m * map_bytes, cudaMemcpyHostToDevice, s));
        const int st = decode_device((const float *)pipe.d_in[b], m, H, W, thr, 0, MC3D_KPT_PLAIN, 0, 0, nullptr, 0,
                                     h_kpt ? d_kpt : nullptr, h_moments ? d_mom : nullptr, s);
        if (st != MC3D_OK) return st;
        if (h_kpt) MC3D_CUDA_TRY(cudaMemcpyAsync(h_kpt + off * 3, d_kpt, (size_t)m * 3 * sizeof(float), cudaMemcpyDeviceToHost, s));
        if (h_moments) MC3D_CUDA_TRY(cudaMemcpyAsync(h_moments + off * 6, d_mom, (size_t)m * 6 * sizeof(double), cudaMemcpyDeviceToHost, s));
    }
    return MC3D_OK;
}

int decode_host(const float *h_hm, long long n_maps, int H, int W, float thr, float *h_kpt, double *h_moments,
                int device) {
    if (n_maps < 0 || H <= 0 || W <= 0) { set_error("bad heatmap shape"); return MC3D_ERR_INVALID_ARGUMENT; }
    if (n_maps == 0) return MC3D_OK;
    if (!h_hm || (!h_kpt && !h_moments)) { set_error("NULL host pointer"); return MC3D_ERR_INVALID_ARGUMENT; }
    if (device < 0 || device >= MAX_DEVICES) { set_error("device ordinal %d out of range", device); return MC3D_ERR_INVALID_ARGUMENT; }
    std::lock_guard<std::mutex> lock(g_pipe_mutex);
    DeviceGuard guard;                                                      // restores the caller's device on every exit path
    { const int gs = guard.enter(device); if (gs != MC3D_OK) return gs; }
    const size_t map_bytes = (size_t)H * W * sizeof(float);
    long long chunk = (long long)((64u << 20) / map_bytes);
    if (chunk < 1) chunk = 1;
    if (chunk > n_maps) chunk = n_maps;
    const size_t kpt_bytes = ((size_t)chunk * 3 * sizeof(float) + 255) / 256 * 256;   // moments start 256-B aligned
    HostPipe *pipe = nullptr;
    const int st = ensure_pipe(device, (size_t)chunk * map_bytes, kpt_bytes + (size_t)chunk * 6 * sizeof(double), &pipe);
    if (st != MC3D_OK) return st;
    return drain_pipe(*pipe, decode_host_chunks(*pipe, h_hm, n_maps, chunk, kpt_bytes, H, W, thr, h_kpt, h_moments));
}

}  // namespace mc3d

extern "C" {

int mc3d_decode_heatmaps_host_f32(const float *h_heatmaps, int64_t n_maps, int H, int W, float threshold,
                                  float *h_kpt, double *h_moments, int device) {
    return mc3d::decode_host(h_heatmaps, n_maps, H, W, threshold, h_kpt, h_moments, device);
}

int mc3d_version(void) { return MC3D_VERSION; }

const char *mc3d_last_error(void) { return mc3d::g_err; }

const char *mc3d_status_string(int status) {
    switch (status) {
        case MC3D_OK: return "ok";
        case MC3D_ERR_INVALID_ARGUMENT: return "invalid argument";
        case MC3D_ERR_MISALIGNED: return "misaligned pointer";
        case MC3D_ERR_CUDA: return "CUDA error";
        case MC3D_ERR_NO_DEVICE: return "no CUDA device";
        case MC3D_ERR_UNSUPPORTED: return "unsupported";
        default: return "unknown status";
    }
}

int64_t mc3d_launch_count(void) { return mc3d::g_launches.load(); }

int mc3d_device_info(char *name, int name_len, int *sm_count, int *cc_major, int *cc_minor) {
    int dev = 0;
    MC3D_CUDA_TRY(cudaGetDevice(&dev));
    cudaDeviceProp prop;
    MC3D_CUDA_TRY(cudaGetDeviceProperties(&prop, dev));
    if (name && name_len > 0) {
        strncpy(name, prop.name, (size_t)name_len - 1);
        name[name_len - 1] = 0;
    }
    if (sm_count) *sm_count = prop.multiProcessorCount;
    if (cc_major) *cc_major = prop.major;
    if (cc_minor) *cc_minor = prop.minor;
    return MC3D_OK;
}

int mc3d_triangulate_host_f32(const float *h_kpts, int64_t n, const mc3d_rig *rig, int layout, int mode, int flags,
                              float *h_out, int device) {
    return mc3d::triangulate_host<float>(h_kpts, n, rig, layout, mode, flags, h_out, device);
}

int mc3d_triangulate_host_f64(const double *h_kpts, int64_t n, const mc3d_rig *rig, int layout, int mode, int flags,
                              double *h_out, int device) {
    return mc3d::triangulate_host<double>(h_kpts, n, rig, layout, mode, flags, h_out, device);
}

// ---- peer memory (in-kernel refinement exchange) ------------------------------------------------------------
int mc3d_peer_alloc(int64_t bytes, void **d_ptr, unsigned char *handle) {
    if (bytes <= 0 || !d_ptr || !handle) { mc3d::set_error("mc3d_peer_alloc: bad arguments"); return MC3D_ERR_INVALID_ARGUMENT; }
    static_assert(sizeof(cudaIpcMemHandle_t) <= MC3D_IPC_HANDLE_BYTES, "IPC handle size");
    void *p = nullptr;
    MC3D_CUDA_TRY(cudaMalloc(&p, (size_t)bytes));
    cudaError_t e = cudaMemset(p, 0, (size_t)bytes);
    cudaIpcMemHandle_t h;
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { cudaFree(p); return mc3d::cuda_fail(e, "mc3d_peer_alloc"); }
    memset(handle, 0, MC3D_IPC_HANDLE_BYTES);
    memcpy(handle, &h, sizeof(h));
    *d_ptr = p;
    return MC3D_OK;
}

int mc3d_peer_open(const unsigned char *handle, void **d_ptr) {
    if (!handle || !d_ptr) { mc3d::set_error("mc3d_peer_open: bad arguments"); return MC3D_ERR_INVALID_ARGUMENT; }
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof(h));
    void *p = nullptr;
    MC3D_CUDA_TRY(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    *d_ptr = p;
    return MC3D_OK;
}

int mc3d_peer_close(void *d_ptr) {
    if (!d_ptr) return MC3D_OK;
    MC3D_CUDA_TRY(cudaIpcCloseMemHandle(d_ptr));
    return MC3D_OK;
}

int mc3d_peer_free(void *d_ptr) {
    if (!d_ptr) return MC3D_OK;
    MC3D_CUDA_TRY(cudaFree(d_ptr));
    return MC3D_OK;
}

}  // extern "C"
