/* Plain C (C99) consumer of include/mc3d.h: proves the header is valid C, that every entry point links from C, and that
 * the struct layouts seen by a C compiler are the ones the library was built with.  No compute call is made (it runs on
 * machines without a GPU).  Built and run by tests/test_abi.py. */
#include <stdio.h>
#include <string.h>
#include "mc3d.h"

int main(void) {
    /* taking the address of every entry point makes the linker resolve all of them */
    typedef void (*fn_t)(void);
    const fn_t entry[] = {
        (fn_t)mc3d_version, (fn_t)mc3d_last_error, (fn_t)mc3d_status_string,
        (fn_t)mc3d_launch_count, (fn_t)mc3d_device_info,
        (fn_t)mc3d_triangulate_f32, (fn_t)mc3d_triangulate_f64,
        (fn_t)mc3d_triangulate_host_f32, (fn_t)mc3d_triangulate_host_f64, (fn_t)mc3d_triangulate_start_plan,
        (fn_t)mc3d_decode_heatmaps_f32, (fn_t)mc3d_decode_heatmaps_host_f32,
        (fn_t)mc3d_project_points_f32, (fn_t)mc3d_project_points_f64,
        (fn_t)mc3d_refine_prepare_f32, (fn_t)mc3d_refine_prepare_f64,
        (fn_t)mc3d_refine_problem_size, (fn_t)mc3d_refine_plan, (fn_t)mc3d_refine_sweep_range,
        (fn_t)mc3d_refine_flags_f32, (fn_t)mc3d_refine_flags_f64,
        (fn_t)mc3d_refine_phase_f32, (fn_t)mc3d_refine_phase_f64,
        (fn_t)mc3d_refine_run_f32, (fn_t)mc3d_refine_run_f64,
        (fn_t)mc3d_extrinsic_problem_size,
        (fn_t)mc3d_extrinsic_run_f32, (fn_t)mc3d_extrinsic_run_f64,
        (fn_t)mc3d_extrinsic_costgrad_f32, (fn_t)mc3d_extrinsic_costgrad_f64,
        (fn_t)mc3d_extrinsic_joint_step_f32, (fn_t)mc3d_extrinsic_joint_step_f64,
        (fn_t)mc3d_peer_alloc, (fn_t)mc3d_peer_open, (fn_t)mc3d_peer_close, (fn_t)mc3d_peer_free,
        (fn_t)mc3d_linear_interpolation_f64,
    };
    const int n_entry = (int)(sizeof(entry) / sizeof(entry[0]));
    int i, bad = 0;
    for (i = 0; i < n_entry; ++i) bad += entry[i] == (fn_t)0;
    if (mc3d_version() != MC3D_VERSION) { printf("version mismatch\n"); return 2; }
    if (mc3d_refine_problem_size() != (int)sizeof(mc3d_refine_problem)) { printf("mc3d_refine_problem: C sees %d bytes, library %d\n", (int)sizeof(mc3d_refine_problem), mc3d_refine_problem_size()); return 3; }
    if (mc3d_extrinsic_problem_size() != (int)sizeof(mc3d_extrinsic_problem)) { printf("mc3d_extrinsic_problem size mismatch\n"); return 4; }
    if (strcmp(mc3d_status_string(MC3D_OK), "ok") != 0) return 5;
    {   /* argument validation happens before any CUDA call: a NULL rig is refused without touching a device */
        float dummy = 0.f;
        if (mc3d_triangulate_host_f32(&dummy, 1, NULL, MC3D_LAYOUT_V3, MC3D_TRI_WEIGHTED, 0, &dummy, 0) != MC3D_ERR_INVALID_ARGUMENT) return 6;
    }
    printf("ok %d entry points, refine problem %d bytes, extrinsic problem %d bytes\n", n_entry - bad,
           (int)sizeof(mc3d_refine_problem), (int)sizeof(mc3d_extrinsic_problem));
    return bad ? 1 : 0;
}
