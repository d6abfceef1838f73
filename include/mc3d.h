/*
 * mc3d.h -- C ABI of libmc3d.so, the B200 (sm_100a) hot path of
 * sashapersonxyz/Multi-camera_3D_Pose_Estimation:
 *
 *   1. multi-view DLT triangulation          (reference utils.py:19-34 `DLT`,
 *                                              utils.py:1277-1336 `triangulate_points`,
 *                                              pose_estimation.py:11-65 `get_pose_3D`)
 *   2. heatmap -> keypoint / Gaussian decode (reference mmpose_pose_estimation.py:163-215
 *                                              `get_heatmap_means_cov`; mmpose argmax decode
 *                                              called at mmpose_pose_estimation.py:253-259)
 *   3. refinement loss + gradient + Adam     (reference pose_refinement.py:836-889 costs,
 *                                              :894-1096 `sgd_optimize`)
 *
 * The reference is pure Python and has no FFI of its own; these entry points are what
 * a ctypes binding for that path binds (INTEGRATION.md shows the stub).  Plain pointers
 * and sizes only.  Every function returns an mc3d_status; mc3d_last_error() returns the
 * text for the calling thread's last failure.
 *
 * Pointer conventions
 *   d_*  device pointer (caller-owned, e.g. torch.Tensor.data_ptr()); 16-byte aligned.
 *   h_*  host pointer (pinned memory gives full PCIe rate; pageable also works).
 *   stream  a cudaStream_t passed as void* (NULL = legacy default stream).  Device-pointer
 *           entry points only enqueue work on `stream`; they never synchronise.
 *   Camera parameters are always HOST double arrays; they travel as kernel parameters.
 */
#ifndef MC3D_H
#define MC3D_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MC3D_VERSION 100          /* 0.1.0 */
#define MC3D_MAX_VIEWS 16         /* cameras per rig handled by one launch */
#define MC3D_MAX_JOINTS 133       /* COCO-WholeBody */
#define MC3D_MAX_BONES 64
#define MC3D_MAX_PEERS 16         /* ranks of one NVLink domain in the in-kernel refinement exchange */
#define MC3D_IPC_HANDLE_BYTES 64

typedef enum {
    MC3D_OK = 0,
    MC3D_ERR_INVALID_ARGUMENT = 1,   /* NULL pointer, n < 0, n_views out of range, bad enum */
    MC3D_ERR_MISALIGNED = 2,         /* a device pointer is not 16-byte aligned */
    MC3D_ERR_CUDA = 3,               /* a CUDA runtime call failed; see mc3d_last_error() */
    MC3D_ERR_NO_DEVICE = 4,          /* no sm_100 device visible */
    MC3D_ERR_UNSUPPORTED = 5
} mc3d_status;

/* Keypoint memory layout of one joint (3*V scalars, contiguous). */
typedef enum {
    MC3D_LAYOUT_V3 = 0,   /* (N, V, 3): [x,y,w] per view                   */
    MC3D_LAYOUT_3V = 1    /* (N, 3, V): [x_0..x_V-1, y.., w..]     (reference kpts_2d (T,J,3,C), */
                          /*                                         pose_estimation.py:135)      */
} mc3d_layout;

/* What the triangulation solves per joint. */
typedef enum {
    MC3D_TRI_WEIGHTED = 0,  /* all V views, rows scaled by w_v: utils.py:21-28 generalised      */
    MC3D_TRI_TOP2 = 1       /* the two highest-score views, unweighted: pose_estimation.py:35-52 */
} mc3d_tri_mode;

enum {
    MC3D_TRI_FLAG_JACOBI = 1,  /* solve every joint with the 4x4 Jacobi eigensolver (test hook) */
    MC3D_TRI_FLAG_FP64 = 2     /* float storage: use the all-double solver instead of the mixed-precision one */
};

/* Per-rig camera description for triangulation (host memory, double).
 *   P      [V][12]  row-major 3x4 projection matrices K[R|T]        (utils.py:433-435)
 *   K      [V][9]   row-major intrinsics, or NULL  } both non-NULL => points are undistorted
 *   dist   [V][5]   k1,k2,p1,p2,k3, or NULL        } like cv.undistortPoints(.., P=K), utils.py:1314
 */
typedef struct {
    int32_t n_views;
    const double *P;
    const double *K;
    const double *dist;
} mc3d_rig;

/* ---- library ---------------------------------------------------------------------------- */
int mc3d_version(void);
const char *mc3d_last_error(void);
const char *mc3d_status_string(int status);
/* Number of kernels this library has launched in the calling process (for bench.py). */
int64_t mc3d_launch_count(void);
/* name, SM count, compute capability of the current device. */
int mc3d_device_info(char *name, int name_len, int *sm_count, int *cc_major, int *cc_minor);

/* ---- 1. triangulation ---------------------------------------------------------------------
 * d_kpts: n joints x 3V scalars in `layout`; d_out: n x 3 (X,Y,Z).
 * Replaces: the per-(frame,joint) loop of pose_estimation.py:27-54 and utils.py:19-34.
 * f32: float storage; a closed-form two-view start, then the float normal matrix and the gradient of ONE pass over the
 *      views with residuals evaluated in double, so the error is one output rounding (MC3D_TRI_FLAG_FP64 selects
 *      all-double arithmetic).
 * Degenerate joints (fewer than two views with non-zero weight, non-finite input) give NaN. */
int mc3d_triangulate_f32(const float *d_kpts, int64_t n, const mc3d_rig *rig, int layout, int mode,
                         int flags, float *d_out, void *stream);
int mc3d_triangulate_f64(const double *d_kpts, int64_t n, const mc3d_rig *rig, int layout, int mode,
                         int flags, double *d_out, void *stream);

#ifndef MC3D_TRI_MAX_START
#define MC3D_TRI_MAX_START 4
#endif
/* Starting-point plan of the float-storage kernel (host-only, no device work): up to MC3D_TRI_MAX_START pairs of views (A, B)
 * -- widest angle first, sharing no view with the earlier pairs while the rig allows it; a joint starts from the first pair
 * whose two views it sees -- whose
 * closed-form two-view point X = C + s H (x_A, y_A, 1),  s = (z kb - ka) / (ua.q - z ub.q),  q = (x_A, y_A, 1),
 * z = alpha x_B + beta y_B  starts the iteration (the ray of view A cut by the plane that view B's pixel spans along
 * its epipolar direction).  Exposed so that the plan can be checked without a GPU; the kernel's result does not
 * depend on it, only its speed.  24 four-byte fields, no padding. */
typedef struct {
    float H[9];            /* inverse of the left 3x3 block of P_A (row-major) */
    float C[3];            /* centre of camera A */
    float ua[3], ub[3];    /* H^T (alpha P0 + beta P1)_B[:3],  H^T P2_B[:3] */
    float alpha, beta;     /* unit epipolar direction in view B */
    float ka, kb;          /* (alpha P0 + beta P1)_B . (C,1),  P2_B . (C,1) */
    int32_t view_a, view_b;
} mc3d_tri_start_pair;
int mc3d_triangulate_start_plan(const mc3d_rig *rig, mc3d_tri_start_pair *pairs /* [MC3D_TRI_MAX_START] */, int32_t *n_pairs);

/* Host-buffer variants: chunked H2D -> kernel -> D2H pipeline on the library's own streams;
 * returns after the last byte of h_out is written.  `device` = CUDA ordinal.  One staging pipeline
 * (three streams, three pairs of device buffers for chunks of <= 64 MiB) per device ordinal, created
 * on first use and kept for the life of the process; calls are serialised by a mutex; on an error
 * every stream is drained before the call returns, so nothing still touches the caller's buffers. */
int mc3d_triangulate_host_f32(const float *h_kpts, int64_t n, const mc3d_rig *rig, int layout, int mode,
                              int flags, float *h_out, int device);
int mc3d_triangulate_host_f64(const double *h_kpts, int64_t n, const mc3d_rig *rig, int layout, int mode,
                              int flags, double *h_out, int device);


/* ---- 2. heatmap decode ----------------------------------------------------------------------
 * d_heatmaps: n_maps x H x W floats.  Per map, in one pass over HBM:
 *   d_kpt     [x, y, score]: flat argmax (first maximum), +-0.25 px towards the larger neighbour,
 *             (-1,-1) when the maximum is <= 0 -- the mmpose MSRA decode behind
 *             mmpose_pose_estimation.py:253-259 (third-party upstream; parity unpinned, see
 *             oracle/decode.py); optional affine x*a0+a2, y*a1+a3 per `affine_group` maps.
 *   d_moments [mean_x, mean_y, var_x, cov_xy, cov_xy, var_y] of the map with values < threshold
 *             zeroed (0.01 upstream), six zeros for an all-zero map: get_heatmap_means_cov,
 *             mmpose_pose_estimation.py:163-215.
 * Either output may be NULL.  kpt_layout PLAIN writes d_kpt[map*3..]; NV3 / N3V read the map index
 * as (t, view, joint) -- the reference's heatmap order (T, C, J), pose_estimation.py:110,190 -- and
 * write the keypoint where the triangulation kernel expects it: (T*J, V, 3) or (T*J, 3, V). */
typedef enum { MC3D_KPT_PLAIN = 0, MC3D_KPT_NV3 = 1, MC3D_KPT_N3V = 2 } mc3d_kpt_layout;
enum {
    MC3D_DECODE_FLAG_WRITE_BACK = 1,  /* also store the thresholded maps in place (upstream mutates its input) */
    MC3D_DECODE_FLAG_GENERIC = 2,     /* force the plain-load kernel (test hook) */
    MC3D_DECODE_FLAG_TMA = 4,         /* force the TMA / shared-memory kernel even for 64x48 maps (test hook) */
    MC3D_DECODE_FLAG_NO_STAGE = 8     /* 64x48 maps with moments: plain streaming loads instead of the per-warp TMA stage (A/B) */
};
int mc3d_decode_heatmaps_f32(const float *d_heatmaps, int64_t n_maps, int H, int W, float threshold, int flags,
                             int kpt_layout, int views, int joints, const float *d_affine, int affine_group,
                             float *d_kpt, double *d_moments, void *stream);
/* Host-buffer variant (PLAIN layout, no affine): chunked H2D -> kernel -> D2H. */
int mc3d_decode_heatmaps_host_f32(const float *h_heatmaps, int64_t n_maps, int H, int W, float threshold,
                                  float *h_kpt, double *h_moments, int device);


/* ---- 3. trajectory refinement -----------------------------------------------------------------
 * One optimiser step of Optimized_3d_Pose_Estimation.sgd_optimize (pose_refinement.py:1006-1050) is three
 * phases on the caller's stream:
 *   phase 0  costs    : likelihood (:863-889, camera-0 Gaussians), smoothness (:836-845), bone length
 *                       (:848-860) -> 7 global sums in ctrl
 *   phase 1  gradient : closed-form gradient (replaces total_cost.backward(), :1044) -> g, sum g^2 in ctrl
 *   phase 2  step     : clip_grad_norm_(1.0) (:1047), Adam (:1050), running-mean early stopping (:1069-1089),
 *                       best_trajectory snapshot (:1075), cost history (:1052-1054)
 * All state is caller-owned device memory (state dtype = float or double per entry-point suffix):
 *   x        (n_frames + 4, J, 3)  trajectory with TWO HALO FRAMES at each end (multi-GPU: neighbours' frames)
 *   m, v, best, g  (n_frames, J, 3)  Adam moments, best snapshot, gradient scratch
 *   mu0 (n_frames, J, 2), S (n_frames, J, 3)   from mc3d_refine_prepare_*: camera-0 means and the symmetric
 *                                              inverse of (cov + 1e-6 I) as [s00, s01, s11] (:663-668)
 *   term_ok  (n_frames + 4) bytes  validity of the smoothness term ending at each frame (halo-extended); written
 *                                  once per run by mc3d_refine_flags_* after the x halo is in place
 *   ctrl     doubles, >= 64 + 4 * hist_capacity, zero-filled except ctrl[32+3] = ctrl[48+3] = +inf:
 *            [0..7]+16p   sums of the step with parity p: S_lik N_lik S_smooth N_smooth a.b b.b a.a gnorm^2
 *            [32..39]+16p state entering a step of parity p: adam_step run_sum run_cnt best no_improve
 *                         stopped iterations improved
 *            [64+4s ..]   cost history of step s: total likelihood smoothness body_length
 * Multi-GPU (frames sharded contiguously, one process per GPU), two ways:
 *   host-driven: the driver exchanges the x halo after phase 2 and all-reduces ctrl[0..6]+16p after phase 0 and
 *                ctrl[7]+16p after phase 1 with any collective library (that is the whole exchange; every rank then
 *                takes the same decisions);
 *   in-kernel  : `xchg` holds every rank's exchange block (mc3d_peer_alloc / mc3d_peer_open, NVLink peer memory) and
 *                the kernels do the exchange themselves: the last block of phase 0 / 1 stores this rank's partial
 *                sums into every peer's block followed by a sequence flag, phases 1 / 2 spin on their own block's
 *                flags and add the partial sums in rank order (bitwise identical on every rank), and phase 2 stores
 *                its boundary frames straight into the neighbours' halo frames.  No NCCL call, no host round trip,
 *                and the whole step sequence can be replayed from one CUDA graph on every rank.
 *                In this mode `x` must live inside the exchange allocation at byte MC3D_XCHG_X_OFFSET. */
typedef struct {
    int32_t n_joints, n_cams, n_bones, ignore_distortions;
    int32_t patience, max_iter;
    int64_t n_frames;        /* local frames (without halo) */
    int64_t frame_offset;    /* global index of local frame 0 */
    int64_t win_begin, win_end;   /* global frame window the costs are evaluated on (a batch, :786-796) */
    int64_t hist_capacity;
    int64_t total_frames;    /* frames of the whole (unsharded) trajectory */
    double lr, beta1, beta2, eps, lambda_smooth, lambda_body, tolerance;
    double aa;               /* ||a||^2 = window length x sum of squared target bone lengths (:857) */
    double cams[MC3D_MAX_VIEWS][26];      /* per camera: K[9] R[9] T[3] dist[5] (pose_refinement.py:94) */
    double bone_len[MC3D_MAX_BONES];
    int32_t bone_start[MC3D_MAX_BONES], bone_end[MC3D_MAX_BONES];
    int32_t adj_start[MC3D_MAX_JOINTS + 3];          /* CSR joint -> incident bones */
    int32_t adj_bone[2 * MC3D_MAX_BONES], adj_sign[2 * MC3D_MAX_BONES];   /* sign +1: joint is the bone's end */
    void *x, *m, *v, *best, *g, *mu0, *S;
    uint8_t *term_ok;
    double *ctrl;
    /* in-kernel exchange (all zero / NULL = off) */
    void *gc;                /* 4 gradient components of the two-phase step, each n_frames*J*3 scalars at a stride rounded
                              * up to a multiple of 4 scalars (16-byte aligned starts); NULL selects the three phases */
    int32_t rank, world;
    int64_t n_frames_left;   /* local frame count of rank - 1 (locates its right halo) */
    int64_t spin_timeout_ns; /* a wait on a peer gives up after this long and sets mc3d_refine_xchg.error */
    void *xchg[MC3D_MAX_PEERS];   /* exchange blocks of all ranks as mapped in this process; xchg[rank] is local */
    /* 0: camera 0's means and inverse covariances are used for EVERY camera (what upstream computes: pose_refinement.py:663,
     * :885, quirk Q1) and mu0 / S are (n_frames, J, 2 / 3).  n_frames * n_joints: per-camera Gaussians (the form of the
     * superseded Trajectory_Optimization, :499) -- mu0 / S are (n_cams, n_frames, J, 2 / 3), one mc3d_refine_prepare_* call
     * per camera. */
    int64_t gauss_cam_stride;
    /* NULL: the cameras are the `cams` above (kernel parameters).  Otherwise device memory holding n_cams x 26 doubles in
     * the same order, read by every launch: cameras that are learnt together with the trajectory (pose_refinement.py:931-961)
     * are updated on the device between the phases of a step, with no host round trip. */
    const double *cams_dev;
    /* Test hook (0 in production).  Bit 0: the persistent kernel treats the counts of finite terms as changed at every fourth
     * optimiser step, which sends it through the repetition of pass 1 that a value turning non-finite would cause. */
    int64_t test_flags;
} mc3d_refine_problem;

/* Exchange block at the start of each rank's peer allocation (zero-filled by mc3d_peer_alloc). */
#define MC3D_XCHG_X_OFFSET 32768
#define MC3D_XCHG_BLOCK_FLAGS 960 /* blocks of the persistent kernel that can announce their range edges */
#define MC3D_REFINE_SUMS2 17     /* sums of the two-phase step: 7 cost sums + 10 gradient-component dot products */
typedef struct {
    double sums[2][MC3D_MAX_PEERS][8];      /* [step parity][source rank]: S_lik N_lik S_s N_s a.b b.b a.a | gnorm^2 */
    int64_t seq_costs[2][MC3D_MAX_PEERS];   /* adam step + 1 of the cost sums stored in that slot */
    int64_t seq_grad[2][MC3D_MAX_PEERS];    /* the same for gnorm^2 */
    int64_t halo_seq[2];                    /* adam step count the left / right halo frames belong to */
    int64_t ticket[4];                      /* block tickets of phases 0, 1, 2 (local) */
    int64_t gen[4];                         /* persistent kernel: grid-barrier generation of phases 0, 1, 2 (local) */
    int64_t error;                          /* != 0: a wait timed out (results are invalid) */
    /* two-phase step (mc3d_refine_run_*): one exchange of MC3D_REFINE_SUMS2 sums per step */
    double acc2[2][24];                     /* [parity]: this rank's sums (atomic targets of its blocks) */
    double sums2[2][MC3D_MAX_PEERS][24];    /* [parity][source rank] */
    int64_t seq2[2][MC3D_MAX_PEERS];
    int64_t ll[2][MC3D_MAX_PEERS][40];      /* persistent kernel: LL words, (step number << 32) | 32 data bits; 2 per sum */
    int64_t ll_retry[MC3D_MAX_PEERS][40];   /* the same for a repeated pass 1 of a step (counts of finite terms changed) */
    int64_t blk_seq[MC3D_XCHG_BLOCK_FLAGS]; /* fused sweep: Adam step count the two edges of block b's item range hold (local) */
} mc3d_refine_xchg;

/* Pinhole + 5-coefficient Brown projection of n points (n,3) -> (n,2) for one camera given as
 * cam26 = K[9] R[9] T[3] dist[5] (host doubles): project_points_torch, pose_refinement.py:94-179. */
int mc3d_project_points_f32(const float *d_points, int64_t n, const double *cam26, int ignore_distortions, float *d_out, void *stream);
int mc3d_project_points_f64(const double *d_points, int64_t n, const double *cam26, int ignore_distortions, double *d_out, void *stream);

int mc3d_refine_prepare_f32(const float *d_gaussians, int64_t n_frames, int n_cams, int n_joints, int cam, double eps,
                            float *d_mu0, float *d_S, void *stream);
int mc3d_refine_prepare_f64(const double *d_gaussians, int64_t n_frames, int n_cams, int n_joints, int cam, double eps,
                            double *d_mu0, double *d_S, void *stream);
/* term_ok[s + 2] = frames s, s-1, s-2 all finite and inside the trajectory, for local s in [0, n_frames + 2). */
int mc3d_refine_flags_f32(const mc3d_refine_problem *pb, void *stream);
int mc3d_refine_flags_f64(const mc3d_refine_problem *pb, void *stream);
/* sizeof(mc3d_refine_problem), so that a binding can check its struct layout. */
int mc3d_refine_problem_size(void);
/* Host mirror of the fused sweep's work split (the function the kernel calls): the items [lo, hi) of the (frame, joint) order
 * that thread block `block` of `grid` owns on rank `rank` of `world`.  Boundaries are multiples of 4 items; a block next to a
 * neighbour rank gets one trip (256 threads x 8 / elem_size items) less than the others.  For tests and capacity planning. */
int mc3d_refine_sweep_range(int64_t n_items, int grid, int n_joints, int rank, int world, int elem_size, int block, int64_t *lo, int64_t *hi);
int mc3d_refine_phase_f32(const mc3d_refine_problem *pb, int phase, int64_t step_index, int end_of_iteration, void *stream);
int mc3d_refine_phase_f64(const mc3d_refine_problem *pb, int phase, int64_t step_index, int end_of_iteration, void *stream);
/* n_iters whole-window iterations.
 * Two-phase step (needs `gc` and the exchange block `xchg`; any world size): the gradient does not wait for the
 * global sums.  It is linear in three scalars that depend on them,
 *     g = alpha g1 + sigma gs + beta (G2 - mu G3),   alpha = 1/N_lik, sigma = 2 lambda_s/N_s, beta = -2 lambda_b mu/(a.a),
 * so ONE pass computes the costs, the component vectors (g1 likelihood, gs smoothness, G2' = G2 - mu_prev G3 and
 * G3 bone length; mu_prev = last step's mu keeps the cancelling pair small) and their mutual dot products.  After
 * ONE reduction every thread knows alpha, sigma, beta, mu and |g|^2 (a quadratic form in them), and the Adam pass
 * combines the components, clips and applies Adam.  Per step: one cross-rank exchange of the sums and the halo stores
 * (the three-phase step needs three passes and two exchanges).
 * ctrl[32 + 16p + 8] carries mu_prev, [.. + 9], [.. + 10] the counts N_lik, N_s the step found.
 * All iterations run inside one persistent cooperative kernel, which stores THREE components: gA = alpha g1 + sigma gs is
 * folded with the previous step's counts (they change only when a value turns non-finite; the step's totals are checked
 * and pass 1 is repeated once if they differ), 13 sums.  Float state: the FUSED SWEEP -- Adam of step s and pass 1 of step
 * s + 1 in one sweep over block-owned item ranges (cp.async-staged Adam operands in flight during pass 1, range edges
 * announced through blk_seq), ONE grid-wide meeting per step, any shard size below 2^30 joint-frames.  Double state:
 * the two passes one after the other (two grid barriers per step), any shard size as well (the float two-pass form --
 * MC3D_REFINE_SWEEP=0 -- up to ~150 000 frames x 17 joints, beyond that a CUDA graph of the three
 * phases).  MC3D_REFINE_FUSED=0 selects the graph of two kernels (four components, 17 sums).
 * Without `gc`: phases 0,1,2 replayed from a CUDA graph of the three kernels (with the in-kernel exchange between them
 * when world > 1).  MC3D_REFINE_TWO_PHASE / MC3D_REFINE_FUSED / MC3D_REFINE_SWEEP = 0 / 1 in the environment force a
 * variant (tests, A/B timing).  Every rank must call it with the same arguments.  The best-trajectory snapshot is complete
 * when the call returns (inside the persistent kernel it is deferred while consecutive steps improve). */
/* Text describing what mc3d_refine_run_* launches for this problem on the current device (static string). */
const char *mc3d_refine_plan(const mc3d_refine_problem *pb);
int mc3d_refine_run_f32(const mc3d_refine_problem *pb, int64_t first_step, int64_t n_iters, void *stream);
int mc3d_refine_run_f64(const mc3d_refine_problem *pb, int64_t first_step, int64_t n_iters, void *stream);

/* Peer memory for the in-kernel exchange: a zero-filled device allocation on the current device plus its CUDA IPC
 * handle (MC3D_IPC_HANDLE_BYTES bytes) for the other ranks of the node; mc3d_peer_open maps another rank's
 * allocation into this process with peer access enabled (NVLink / NVSwitch on an HGX B200 board). */
int mc3d_peer_alloc(int64_t bytes, void **d_ptr, unsigned char *handle);
int mc3d_peer_open(const unsigned char *handle, void **d_ptr);
int mc3d_peer_close(void *d_ptr);
int mc3d_peer_free(void *d_ptr);


/* ---- 3b. one camera's extrinsics from sampled points ---------------------------------------------------------
 * The optimize_trajectory=False / extrinsic_optimization_IDs=[id] branch of sgd_optimize (pose_refinement.py:915-1091,
 * cost :800-831): N samples per (frame, joint) -- drawn from the two ground-truth cameras' Gaussians (:684-706) and
 * triangulated (mc3d_triangulate_*) -- are projected with the learnt camera and scored with 0.5 d^T S d against one
 * Gaussian per (frame, joint); the mean over the finite samples is minimised over the 9 entries of R and the 3 of T
 * with clip_grad_norm_(1.0) + Adam and the same running-mean early stopping as the trajectory optimiser.
 *   samples3d (T, J, N, 3), mean (T, J, 2), S (T, J, 3) [s00 s01 s11]: device, state dtype
 *   params   48 device doubles: R[9] T[3] | Adam m[12] | v[12] | best R, T[12]
 *   ctrl     device doubles >= 64 + 2 * hist_capacity, zero-filled except ctrl[32+3] = ctrl[48+3] = +inf:
 *            [0..13]+16p sums (cost, count, dR[9], dT[3]); [32..39]+16p state as in mc3d_refine_problem;
 *            [64+2s..] history of step s: sample cost, total cost (= sample cost + const_cost)            */
typedef struct {
    int64_t n_frames;
    int32_t n_joints, n_samples, ignore_distortions, patience, max_iter, reserved;
    int64_t hist_capacity;
    double lr, beta1, beta2, eps, tolerance;
    double const_cost;       /* smoothness + bone-length cost of the fixed trajectory (:984-986) */
    double K[9], dist[5];    /* intrinsics of the learnt camera */
    const void *samples3d, *mean, *S;
    double *params, *ctrl;
} mc3d_extrinsic_problem;
int mc3d_extrinsic_problem_size(void);
int mc3d_extrinsic_run_f32(const mc3d_extrinsic_problem *pb, int64_t first_step, int64_t n_iters, void *stream);
int mc3d_extrinsic_run_f64(const mc3d_extrinsic_problem *pb, int64_t first_step, int64_t n_iters, void *stream);
/* Cameras and trajectory learnt together (extrinsic_optimization_IDs with optimize_trajectory=True, pose_refinement.py:931-961):
 * the trajectory steps through mc3d_refine_phase_* 0, 1, 2; between phases 1 and 2, per learnt camera,
 * mc3d_extrinsic_costgrad_* accumulates the raw sums (cost, count, dR, dT) over the trajectory points (problem with
 * n_samples = 1, samples3d = the trajectory, sums into ctrl[0..13], no step), and mc3d_extrinsic_joint_step_* turns them
 * into gradients (divided by the N_lik of all cameras), adds their squared norms to refine_ctrl's |g|^2 so that phase 2
 * clips with the norm over every learnable parameter (:1047), applies Adam to the cameras and zeroes cam_ctrl.
 *   cam_ctrl 64 zero-filled doubles per camera (that camera's mc3d_extrinsic_problem.ctrl); cam_params 36 doubles per
 *   camera: R[9] T[3] | m[12] | v[12] (that camera's mc3d_extrinsic_problem.params). */
int mc3d_extrinsic_costgrad_f32(const mc3d_extrinsic_problem *pb, void *stream);
int mc3d_extrinsic_costgrad_f64(const mc3d_extrinsic_problem *pb, void *stream);
int mc3d_extrinsic_joint_step_f32(double *d_refine_ctrl, double *d_cam_ctrl, double *d_cam_params, int n_learn, int64_t step_index,
                                  double lr, double beta1, double beta2, double eps, void *stream);
int mc3d_extrinsic_joint_step_f64(double *d_refine_ctrl, double *d_cam_ctrl, double *d_cam_params, int n_learn, int64_t step_index,
                                  double lr, double beta1, double beta2, double eps, void *stream);


/* ---- 4. linear interpolation (pose_refinement.py:15-84) ---------------------------------------------------
 * d_points / d_out: (n_frames, scalars_per_frame) doubles, scalars_per_frame = joints x dims.  Per scalar and frame:
 * window of k//2 frames each side, outliers outside mean +- k_std*std and (optionally) median +- median_std*MAD are
 * dropped, the rest is fitted by a least-squares line evaluated at the frame (or averaged); fewer than two survivors
 * give 0 as upstream.  k <= 64. */
int mc3d_linear_interpolation_f64(const double *d_points, int64_t n_frames, int64_t scalars_per_frame, int k, double k_std,
                                  double median_std, int use_rolling_average, int filter_distance_from_median,
                                  double *d_out, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* MC3D_H */
