/*
 * igi_b200 — C-ABI of the B200-native observation hot path of
 * FactoryTaskInsertionTactile (osheraz/IsaacGymInsertion).
 *
 * Contract: every pointer is a DEVICE pointer owned by the caller
 * (torch tensors on the Python side); the callee never allocates, frees or
 * synchronises; kernels are enqueued on `stream` (a cudaStream_t passed as
 * void*); the return value is 0 on success, <0 on error (igi_last_error() holds
 * the message).  There is no CPU fallback.
 *
 * The reference is pure Python and has no FFI of its own; each entry point cites
 * the reference Python function (file:line under /root/reference) it replaces.
 * INTEGRATION.md shows the ctypes binding a maintainer adds on the reference side.
 */
#ifndef IGI_B200_H
#define IGI_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define IGI_B200_VERSION 100 /* major*100 + minor */

int igi_version(void);
const char* igi_last_error(void);

/* --------------------------------------------------------------------------
 * (P) external-camera point cloud
 * -------------------------------------------------------------------------- */

/* K4  depth+seg -> world points -> box filter -> ORDERED compaction, per env and
 * per segmentation class.
 * Replaces: seg masking  tasks/factory_tactile/factory_task_insertion.py:956-959,975
 *           PointCloudGenerator.convert          tasks/utils/pcl_utils.py:62-90
 *           filter_pts                           factory_task_insertion.py:66-77
 *           CameraPointCloud.get_ptd_cuda        tasks/utils/pcl_utils.py:203-212
 *   depth      (n_envs, H*W) f32, negative metric depth, -inf on ray miss
 *   seg        (n_envs, H*W) i32 or NULL (no masking: one class, seg_ids ignored)
 *   seg_ids    n_classes host ints (e.g. {2,3} = plug, socket)
 *   uvx        (n_envs, W) f32  first  component of _uv_one_in_cam[0, u]
 *   uvy        (n_envs, H) f32  second component of _uv_one_in_cam[v, 0]
 *   uvz        (n_envs)    f32  third  component (1.0 up to torch.inverse rounding)
 *   ext        (n_envs, 16) f32 row-major inverse(view_matrix)      (row-vector convention)
 *   e2g_inv    (n_envs, 16) f32 row-major inverse(env_to_global)    (applied as pts @ M^T)
 *   depth_max  valid = depth > -depth_max; pass a negative value to disable (depth_max=None)
 *   box        6 host floats {x_lo,x_hi,y_lo,y_hi,z_lo,z_hi}, inclusive; NULL = no filter
 *   out_pts    (n_envs, n_classes, H*W, 3) f32 compacted points, row-major pixel order
 *   out_count  (n_envs, n_classes) i32 number of points kept
 *   out_any    (n_envs, n_classes) i32 1 iff any kept coordinate is non-zero
 *              (`all_pts[env_id].any()` pcl_utils.py:179)
 */
int igi_pcl_compact(const float* depth, const int32_t* seg, const int32_t* seg_ids, int n_classes,
                    const float* uvx, const float* uvy, const float* uvz, const float* ext,
                    const float* e2g_inv, int n_envs, int H, int W, float depth_max, const float* box,
                    float* out_pts, int32_t* out_count, int32_t* out_any, void* stream);

/* K5A  reference sampler: ids = raw % count, with replacement.
 * Replaces: CameraPointCloud.get_point_cloud + sample_n   pcl_utils.py:168-184,195-201
 * The reference draws torch.randint(0, count, (m,)) from the CPU MT19937 only for
 * envs whose cloud is non-empty, in env order; `raw` is that generator's 32-bit
 * output stream (n_envs*m words, enough for the all-non-empty case) and the kernel
 * consumes it at offset m * (#non-empty envs before this one).
 *   pts/count/any   as written by igi_pcl_compact, for ONE class: class stride given
 *   raw             (n_envs*m) u32
 *   out             (n_envs, m, 3) f32 (zeros for empty envs), row stride out_stride floats
 *   out_idx         (n_envs, m) i32 or NULL
 *   out_consumed    (1) i32: number of raw words consumed (m * #non-empty), or NULL
 *   scratch_offsets (n_envs) i32 caller-owned scratch (per-env stream offsets)
 */
int igi_pcl_sample_gather(const float* pts, const int32_t* count, const int32_t* any, int n_classes,
                          int cls, int cap, const uint32_t* raw, int n_envs, int m, float* out,
                          int64_t out_stride, int32_t* out_idx, int32_t* out_consumed,
                          int32_t* scratch_offsets, void* stream);

/* K5B  farthest-point sampling (north-star sampler).
 * Semantics of pointnet2_ops.furthest_point_sample, the only FPS in the reference
 * (algo/models/transformer/point_mae.py:14-21): start at index 0, temp=1e10,
 * points with |p|^2 <= 1e-3 never become candidates, arg-max ties resolved as the
 * upstream block reduction does (oracle/fps.py states the rule).
 *   pts     task t's points at pts + t*task_stride floats, (n_t, 3) f32
 *   count   (n_tasks) i32 per-task n_t (device) or NULL => every task has n_fixed points
 *   any     (n_tasks) i32 or NULL: tasks with any==0 produce zeros (pcl_utils.py:179-183)
 *   out_pts (n_tasks, m, 3) f32 gathered points or NULL, row stride out_stride floats
 *   out_idx (n_tasks, m) i32 or NULL
 */
int igi_fps(const float* pts, int64_t task_stride, const int32_t* count, const int32_t* any,
            int64_t count_stride, int n_fixed, int n_tasks, int m, float* out_pts, int64_t out_stride,
            int32_t* out_idx, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* IGI_B200_H */
