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
    MC3D_TRI_FLAG_JACOBI = 1   /* solve every joint with the 4x4 Jacobi eigensolver (test hook) */
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
 * f32: float storage, double arithmetic inside (error = one output rounding).
 * Degenerate joints (fewer than two views with non-zero weight, non-finite input) give NaN. */
int mc3d_triangulate_f32(const float *d_kpts, int64_t n, const mc3d_rig *rig, int layout, int mode,
                         int flags, float *d_out, void *stream);
int mc3d_triangulate_f64(const double *d_kpts, int64_t n, const mc3d_rig *rig, int layout, int mode,
                         int flags, double *d_out, void *stream);

/* Host-buffer variants: chunked H2D -> kernel -> D2H pipeline on the library's own streams;
 * returns after the last byte of h_out is written.  `device` = CUDA ordinal. */
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
    MC3D_DECODE_FLAG_GENERIC = 2      /* force the non-TMA kernel (test hook) */
};
int mc3d_decode_heatmaps_f32(const float *d_heatmaps, int64_t n_maps, int H, int W, float threshold, int flags,
                             int kpt_layout, int views, int joints, const float *d_affine, int affine_group,
                             float *d_kpt, double *d_moments, void *stream);
/* Host-buffer variant (PLAIN layout, no affine): chunked H2D -> kernel -> D2H. */
int mc3d_decode_heatmaps_host_f32(const float *h_heatmaps, int64_t n_maps, int H, int W, float threshold,
                                  float *h_kpt, double *h_moments, int device);

#ifdef __cplusplus
}
#endif
#endif /* MC3D_H */
