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
/* Number of kernels this library has launched so far in the process (bench.py's gpu_launches). */
long long igi_launch_count(void);

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
 *   flags   0, or IGI_FPS_NO_CLUSTER: tasks of 1025..8192 points, which normally run on thread-block clusters
 *           (4 CTAs x 256 threads, candidates exchanged through distributed shared memory), go to the one-CTA
 *           kernel instead (same results; for A/B timing and tests)
 */
#define IGI_FPS_NO_CLUSTER 1
int igi_fps(const float* pts, int64_t task_stride, const int32_t* count, const int32_t* any,
            int64_t count_stride, int n_fixed, int n_tasks, int m, float* out_pts, int64_t out_stride,
            int32_t* out_idx, int flags, void* stream);

/* K5B with a size-ordered schedule.  Same results as igi_fps for device-side counts; the tasks
 * are first counting-sorted by point count (descending) and persistent CTAs pull them longest
 * first, so the SMs finish together although neighbouring envs' clouds differ several-fold in
 * size.  One call can cover several classes of igi_pcl_compact's output at once: with
 * task_stride = cap*3, count_stride = 1 and n_tasks = n_envs*n_classes, task t = env*n_classes
 * + class and out_pts is (n_envs, n_classes, m, 3) - for the task that is its packed
 * [plug | socket] cloud row (factory_task_insertion.py:1014-1027), no concatenation pass.
 *   scratch  (n_tasks + 8) i32 caller-owned: schedule counters + the ordered task list
 */
int igi_fps_balanced(const float* pts, int64_t task_stride, const int32_t* count, const int32_t* any,
                     int64_t count_stride, int n_tasks, int m, float* out_pts, int64_t out_stride,
                     int32_t* out_idx, int32_t* scratch, int flags, void* stream);

/* --------------------------------------------------------------------------
 * (T) allsight tactile renderer
 * -------------------------------------------------------------------------- */

/* Sensor constants (host memory).  They travel with every call and reach the kernels as a kernel parameter, so
 * the library keeps NO sensor state: engines with different sensor yamls, or on different devices, can share a
 * process.  Values come from the sensor yaml the reference loads at tacto/renderer.py:86-87
 * (config_allsight_white.yml) and from tacto_allsight_wrapper/allsight_wrapper.py:100-174 (spot lights),
 * expressed in the CAMERA frame (camera at the origin looking down -z, OpenGL convention). */
typedef struct IgiSensorParams {
  int32_t width, height;        /* 224 x 224 (FactoryTaskInsertionTactile.yaml:31-33) */
  float znear;                  /* yml camera.znear */
  float dxp_first, dxp_last;    /* first / last entry of the ray-slope table of the columns (IgiTactileStatic.dxp) */
  float dyp_first, dyp_last;    /* first / last entry of the ray-slope table of the rows    (IgiTactileStatic.dyp) */
  int32_t n_lights;
  const float* light_pos;       /* (L,3) */
  const float* light_dir;       /* (L,3) unit spot direction */
  const float* light_col;       /* (L,3) */
  const float* light_int;       /* (L)   */
  const float* light_las;       /* (L) 1/max(.001, cos(inner)-cos(outer)) */
  const float* light_lao;       /* (L) -cos(outer)*las */
  int32_t inverse_square;       /* 1: radiance / d^2 (default, pyrender); 0: none (yaml lights.falloff, DESIGN.md 'light model') */
  float base_color[3], metallic, roughness;
  double cam_R[9], cam_p[3];    /* camera zero pose in the sensor frame (renderer.py:305-311) */
  double max_force, max_deformation; /* yml force.range_force[1], force.max_deformation */
  float calib_scale, clip_lo, clip_hi; /* yml bg_calibration */
  int32_t blur_ksize;           /* 7 */
  float gauss[7];               /* cv2.getGaussianKernel(7, sigma) */
  float grid_org[3], grid_h, grid_slack; /* conservative gel-interior distance grid */
  int32_t grid_n[3];
  float depth0_max;
  int32_t hiz_levels, hiz_off[10], hiz_w[10]; /* depth0 max-pyramid layout (IgiTactileStatic.hiz) */
} IgiSensorParams;

/* Mesh table (device): all plug meshes concatenated, faces grouped into clusters. */
typedef struct IgiTactileMeshes {
  const float* verts;           /* (nv,3) object frame, x,y pre-scaled (allsight_render.py:101-107) */
  const float* vnorm;           /* (nv,3) vertex normals */
  const int32_t* faces;         /* (nf,3) vertex ids into verts, cluster order */
  const int32_t* face_orig;     /* (nf) original face index inside its mesh (depth-tie order) */
  const int32_t* meshes;        /* (n_meshes,4) face_off, n_faces, cluster_off, n_clusters */
  const void* clusters;         /* (n_clusters) {float cx,cy,cz,r; int first,count,pad,pad} */
} IgiTactileMeshes;

/* Static per-sensor-type images (device). */
typedef struct IgiTactileStatic {
  const float* depth0;          /* (H,W)   gel depth, get_background_sim renderer.py:165-168 */
  const uint8_t* bg_sim;        /* (H,W,3) raw gel render */
  const uint8_t* bg_real;       /* (n_bg,H,W,3) real reference frames, renderer.py:555-558 */
  const float* obs_empty;       /* (2048)  observation of a no-contact frame */
  const float* grid;            /* (nz,ny,nx) distance grid */
  const float* hiz;             /* max-pyramid of depth0, level l at hiz + hiz_off[l], row length hiz_w[l] */
  const float* dxp;             /* (W) ray slope per column: ((px+.5)/W*2-1)*tan(yfov/2)*aspect */
  const float* dyp;             /* (H) ray slope per row:    (1-(py+.5)/H*2)*tan(yfov/2)        */
} IgiTactileStatic;

/* Per-step inputs (device): poses exactly as update_tactile gathers them
 * (factory_task_insertion.py:481-484), frame f = env*sensors_per_env + sensor. */
typedef struct IgiTactileFrames {
  int32_t n_envs, sensors_per_env;
  const float* finger_pos;      /* (n_envs*S,3), or NULL: sensor n's positions at finger_pos_n[n] + env*finger_pos_stride */
  const float* finger_quat;     /* (n_envs*S,4) xyzw, or NULL: finger_quat_n[n] + env*finger_quat_stride */
  const float* plug_pos;        /* (n_envs,3) */
  const float* plug_quat;       /* (n_envs,4) xyzw */
  const float* force;           /* (n_envs*S) normal force or NULL -> force_const (70, task :535) */
  float force_const;
  const uint8_t* update;        /* (n_envs) update_freq & update_delay or NULL (all) (task :523) */
  const uint8_t* update2;       /* (n_envs) or NULL: second mask ANDed with `update`, so update_freq and update_delay can be
                                   passed as they are (no logical_and launch) */
  const float* finger_pos_n[8]; /* used when finger_pos == NULL: the reference's per-fingertip state views
                                   (left/right/middle_finger_pos, factory_task_insertion.py:481-483) without a stack copy */
  const float* finger_quat_n[8];
  int64_t finger_pos_stride, finger_quat_stride; /* floats between consecutive envs of one sensor's view */
  const int32_t* mesh_id;       /* (n_envs) */
  const int32_t* bg_id;         /* (n_envs*S) index into bg_real */
  int32_t stage_mask;           /* 0 = whole pipeline (= 8|4).  bits: 1 geometry alone, 2 standalone fill,
                                   4 contact, 8 geometry with the fill fused in (excludes 1 and 2).  Single
                                   stages exist for profiling: later ones reuse the scratch of an earlier run */
  int32_t region_budget;        /* 0 = built-in.  Test hook: cap the pixels (interior + blur halo) one shared-memory
                                   region of the contact kernel may hold, so small inputs exercise the multi-region path */
  int32_t fill_split;           /* 0 = built-in.  Tuning hook: parts mask + 1 of the no-contact result (1 colour, 2
                                   gel_depth, 4 obs) the geometry kernel writes; the contact kernel writes the others.
                                   Results do not depend on the split */
} IgiTactileFrames;

typedef struct IgiTactileScratch {
  float* M;                     /* (F,12) object->camera matrices */
  void* setups;                 /* (F,kmax) 64-byte triangle setup records */
  void* normals;                /* (F,kmax) 48-byte records: camera-frame vertex normals of the triangle in
                                   the same slot of `setups` (interpolated per fragment when shading) */
  int32_t* counts;              /* (F) */
  int32_t* bbox;                /* (F,4) */
  int32_t* worklist;            /* (F) */
  int32_t* counters;            /* (4): work_n, cursor, overflow flag (sticky: a frame produced more than kmax candidate
                                   triangles and the rest was dropped), largest per-frame candidate count seen (sticky).
                                   The caller reads [2], [3] when it likes (an asynchronous 8-byte copy per step needs
                                   no host sync) and clears them */
  int32_t kmax;                 /* <= 4096 */
} IgiTactileScratch;

typedef struct IgiTactileOut {
  uint8_t* color;               /* (F,H,W,3) calibrated tactile image, AllSightRenderer.render()[0] */
  float* gel_depth;             /* (F,H,W)   depth0 - depth,           AllSightRenderer.render()[1] */
  float* obs;                   /* frame (e,n) at obs + e*obs_env_stride + n*obs_sensor_stride:
                                   (2048) f32 = tactile_imgs[e,n] (task :574); strides in floats, % 4 == 0,
                                   so the rows can live inside a packed [tactile|pcl] send buffer */
  int64_t obs_env_stride, obs_sensor_stride;
} IgiTactileOut;

/* K0: depth0 + bg_sim of the static gel.  Replaces Renderer.__init__/_init_camera/_init_light
 * (tacto/renderer.py:65-163,291-325; allsight_wrapper.py:100-174) and get_background_sim (renderer.py:165-168).
 * dxp (W) / dyp (H) device ray-slope tables; gel_tris (n,3,3) f32 sensor frame; scratch_zbuf (H*W) u64. */
int igi_tactile_gel_precompute(const IgiSensorParams* sensor, const float* dxp, const float* dyp, const float* gel_tris,
                               int n_tris, uint64_t* scratch_zbuf, float* depth0, uint8_t* bg_sim, void* stream);

/* K1+K2+K3 for every env x sensor frame.  Replaces the hot loop of _render_tactile
 * (factory_task_insertion.py:515-583): update_pose_given_sim_pose (allsight_render.py:168-172),
 * AllSightRenderer.render (:179-212) -> Renderer.render/adjust_with_force/pyrender draw
 * (tacto/renderer.py:560-648), _calibrate (allsight_wrapper.py:57-98), depth0-depth, remove_bg,
 * mask, flipud, crop, INTER_AREA resize, gray (task :546-574). */
int igi_tactile_render(const IgiSensorParams* sensor, const IgiTactileMeshes* meshes, const IgiTactileStatic* st,
                       const IgiTactileFrames* frames, const IgiTactileScratch* scratch, const IgiTactileOut* out,
                       void* stream);

/* K3 alone: color (F,H,W,3) u8 -> obs.  Replaces factory_task_insertion.py:546-574. */
int igi_tactile_obs(const uint8_t* color, const uint8_t* bg_real, const int32_t* bg_id, int n_frames, float* obs,
                    int64_t obs_stride, void* stream);

/* --------------------------------------------------------------------------
 * (S) stages next to the hot path: image observations of the external camera,
 * RNG-defined noise, history queues, and the student's first preprocessing step.
 * Random numbers are Philox4x32-10 keyed by `seed`, counter = (global element index, step,
 * stream id): results depend on the GLOBAL env id (env0 + local env), not on how envs are
 * sharded over GPUs.  The reference draws from torch's global generators over compacted rows
 * (distribution-level parity only); oracle/student.py restates the same Philox so
 * kernel == oracle bit for bit on the integer parts.
 * -------------------------------------------------------------------------- */

/* S1  depth / segmentation image observations.
 * Replaces: update_external_cam depth_cam + seg_cam branches  factory_task_insertion.py:925-943
 *           DepthImageProcessor.process_depth_image / normalize_depth_image / add_seg_noise
 *                                                         tasks/factory_tactile/factory_utils.py:23-37,55-72
 *   image_buf[e] = ((-clip(depth[e] + dis_noise*2*(U-0.5), -far, -near)) - near) / (far - near)  where update[e]
 *   seg_buf[e]   = seg[e]                                                               where update_seg[e]
 *   seg_buf[e][(seg > 0) & (U' < flip_prob)] = 0                       where seg_noise[e] & update_seg[e]
 *   depth (n_envs,npix) f32, seg (n_envs,npix) i32; update / update_seg / seg_noise (n_envs) u8 or NULL
 *   (NULL update* = every row, NULL seg_noise = none); image_buf / seg_buf may be NULL (stage off).
 *   npix % 4 == 0, rows 16-byte aligned. */
int igi_cam_image_obs(const float* depth, const int32_t* seg, const uint8_t* update, const uint8_t* update_seg,
                      const uint8_t* seg_noise, int n_envs, int npix, long long env0, double dis_noise,
                      double far_clip, double near_clip, float flip_prob, uint64_t seed, uint32_t step,
                      float* image_buf, int32_t* seg_buf, void* stream);

/* S2  PointCloudAugmentations.random_noise, in place, under a per-env mask.
 * Replaces: factory_utils.py:93-100 as called at factory_task_insertion.py:966-969,979-982
 *   p += clamp(N(0,1)*sigma, +-clip) * (U < noise_prob);  p += clamp(pcl_noise[e]*const_noise, +-clip)
 *   pts: env e's (n_pts,3) f32 at pts + e*env_stride floats; mask (n_envs) u8 or NULL (all);
 *   pcl_noise (n_envs,3) f32 = the task's pcl_pos_noise. */
int igi_pcl_noise(float* pts, int64_t env_stride, int n_envs, int n_pts, const uint8_t* mask, const float* pcl_noise,
                  long long env0, float sigma, float noise_clip, float const_noise, float noise_prob, uint64_t seed,
                  uint32_t step, void* stream);

/* S3  RunningMeanStd.forward on (rows, channels) f32, per_channel=False.
 * Replaces: algo/models/running_mean_std.py:60-93 as called by ExtrinsicAdapt.process_obs
 *           (algo/ext_adapt/ext_adapt.py:405: pcl.reshape(-1, 3); :421: student_obs)
 *   training != 0: running_mean/var/count (f64, device) are first updated with this batch's mean and
 *   unbiased variance (parallel-variance update :47-57), then used.
 *   mode 0: clamp((x - mean.float()) / sqrt(var.float() + eps), -5, 5);  1 (norm_only): x / sqrt(...);
 *   2 (unnorm=True): sqrt(...) * clamp(x, -5, 5) + mean.  y may alias x.
 *   scratch: igi_rms_scratch_bytes(channels) bytes, zeroed once by the caller, needed when training. */
long long igi_rms_scratch_bytes(int channels);
int igi_rms_forward(const float* x, long long rows, int channels, double* running_mean, double* running_var,
                    double* count, float epsilon, int training, int mode, float* y, void* scratch, void* stream);

/* S4  seg / depth-image masking of the student's inputs.
 * Replaces: ExtrinsicAdapt.process_obs  algo/ext_adapt/ext_adapt.py:391-396
 *   valid = (seg == obj_id) | (seg == socket_id);  seg_out = distinct ? seg*valid : valid;  img_out = img*valid
 *   n elements (multiple of 4), f32; img / img_out may be NULL; outputs may alias inputs. */
int igi_seg_valid_mask(const float* seg, const float* img, long long n, float obj_id, float socket_id, int distinct,
                       float* seg_out, float* img_out, void* stream);

/* S5  history queue push: queue[:, 1:] = queue[:, :-1]; queue[:, 0] = x (converted to f32).
 * Replaces: factory_task_insertion.py:512-513 (tactile_queue), :1046-1056 (pcl / img / seg queues)
 *   queue (n_envs, hist_len, row_len) f32; x row e at x + e*x_stride elements, f32 or (x_is_int32) i32. */
int igi_queue_push(float* queue, const void* x, int x_is_int32, int64_t x_stride, int n_envs, int hist_len,
                   long long row_len, void* stream);

/* ----------------------------------------------------------------------------
 * (L) trajectory logger buffers: device side of DataLoggerSim
 * (algo/ppo/experience.py:352-490) as used by SimLogger.log_trajectory_data (:634-746) for the
 * observation rows this library produces (tactile (3,2048), img, seg) and any other (n_envs, L) row.
 * -------------------------------------------------------------------------- */

/* L1  log[e, counter[e], :] = float(x[e, :]).
 * Replaces: `self.log_data[key][self.env_ids, self.env_step_counter, ...] = value...` experience.py:426-434
 *   log (n_envs, episode_len, row_len) f32 (128-bit stores when row_len % 4 == 0); x row e at x + e*x_stride
 *   elements, f32 or (x_is_int32) i32; x NULL logs zeros (a None value, :430-431); counter (n_envs) i64.
 *   A counter outside [0, episode_len) sets *overflow = 1 and skips the row (the reference raises). */
int igi_traj_append(float* log, const void* x, int x_is_int32, int64_t x_stride, const long long* counter, int n_envs,
                    int episode_len, long long row_len, int32_t* overflow, void* stream);

/* L2  done_log[e, counter[e]] = done[e]; counter[e] += 1; done_ids[0..*n_done) = envs with done set, env order.
 * Replaces: experience.py:436-445.  done_log rows of done_pitch >= episode_len bytes; done (n_envs) u8 or NULL
 * (none); done_ids (n_envs) i32; n_done (1) i32. */
int igi_traj_step(uint8_t* done_log, int done_pitch, const uint8_t* done, long long* counter, int n_envs, int episode_len,
                  int32_t* done_ids, int32_t* n_done, int32_t* overflow, void* stream);

/* L3  staging[j, :] = buf[ids[j], :] for j < min(*n_ids, max_ids); zero_after != 0 then clears buf[ids[j], :]
 *     (staging NULL: clear only).  Rows of row_bytes (multiple of 4; 128-bit copies when a multiple of 16) bytes.
 * Replaces: the per-env `.clone().cpu()` of experience.py:448-455 and _reset_buffers :417-420. */
int igi_traj_gather(void* buf, const int32_t* ids, const int32_t* n_ids, int max_ids, long long row_bytes, void* staging,
                    int zero_after, void* stream);

/* counter[ids[j]] = 0 for j < min(*n_ids, max_ids)   (_reset_buffers, experience.py:420). */
int igi_traj_reset_counters(long long* counter, const int32_t* ids, const int32_t* n_ids, int max_ids, void* stream);

/* dst[r, :] = src[r, :] for the rows r whose flag[r] != 0 equals want != 0; rows of row_bytes (multiple of 16) bytes at
 * dst + r*dst_stride_bytes / src + r*src_stride_bytes.
 * Replaces: `self.seg_buf[update_seg] = seg[update_seg]` factory_task_insertion.py:934-940 (and any masked row write). */
int igi_copy_rows_where(void* dst, long long dst_stride_bytes, const void* src, long long src_stride_bytes,
                        const uint8_t* flag, int want, long long rows, long long row_bytes, void* stream);

/* The per-env masks of one update_external_cam call in ONE launch (bool tensors are 1 byte per env):
 *   upd_seg   = update_freq & seg_update_delay                                 factory_task_insertion.py:934-940 (NULL: skip)
 *   restarted = socket_pending & (got_socket == 0); got_socket[restarted] = 1   :981,:987
 *               (socket_pending == 2: every env counts as restarted, e.g. after a reset of all envs)
 *   upd_pcl   = (update_freq & update_delay) | restarted                        :896-897,:988-989 */
int igi_cam_masks(const uint8_t* update_freq, const uint8_t* update_delay, const uint8_t* seg_update_delay,
                  int32_t* got_socket, int socket_pending, uint8_t* upd_seg, uint8_t* upd_pcl, uint8_t* restarted,
                  int n_envs, void* stream);

/* pcl[e, :] = src[e, :] where update[e] (NULL: every row), then the history push pcl_queue[:, 1:] = pcl_queue[:, :-1];
 * pcl_queue[:, 0] = pcl, one pass.  Replaces factory_task_insertion.py:1027 and :1046-1048.
 *   src row e at src + e*src_stride floats, pcl row e at pcl + e*pcl_stride floats (the pcl part of the packed
 *   observation row); queue (n_envs, hist_len, row_len) f32 or NULL; row_len % 4 == 0, rows 16-byte aligned. */
int igi_pcl_assemble(const float* src, int64_t src_stride, float* pcl, int64_t pcl_stride, const uint8_t* update,
                     float* queue, int n_envs, int hist_len, long long row_len, void* stream);

/* ----------------------------------------------------------------------------
 * (G) gather of the packed observation rows onto the learner rank without SM work.
 * The reference has no counterpart: its ranks exchange gradients only (algo/ext_adapt/ext_adapt.py:833-851,
 * isaacgyminsertion/train.py:58-64 one process per GPU); the north star adds "all-gather observations onto
 * the learner rank".  isaacgyminsertion_b200/dist.py `ObsGather(transport="p2p")` is the host side: every
 * rank copies its rows straight into the learner's buffer (a cudaMalloc block shared over CUDA IPC) with a
 * copy-engine peer copy and publishes the step number with a second 4-byte copy; waits are stream memory
 * operations on the waiter's OWN memory.  No kernel is launched by any of these entry points.
 * -------------------------------------------------------------------------- */

/* G1  cudaMalloc a zeroed block and export its 64-byte CUDA IPC handle (the one place this library allocates:
 *     IPC handles exist only for whole allocations).  Release with igi_peer_free. */
int igi_peer_alloc(unsigned long long bytes, void** ptr_out, unsigned char* handle64_out);
int igi_peer_free(void* ptr);
/* G2  map another process's block (cudaIpcMemLazyEnablePeerAccess) / unmap it. */
int igi_peer_open(const unsigned char* handle64, void** ptr_out);
int igi_peer_close(void* ptr);
/* G3  dst <- src (bytes) on `stream`, devices resolved by unified addressing: a copy-engine transfer over NVLink. */
int igi_peer_copy_async(void* dst, const void* src, unsigned long long bytes, void* stream);
/* G4  stream memory operations (cuStreamWriteValue32 / cuStreamWaitValue32 with CU_STREAM_WAIT_VALUE_GEQ, cyclic
 *     comparison) on a 4-byte-aligned device address of the calling device. */
int igi_stream_write_value32(void* stream, void* addr, unsigned int value);
int igi_stream_wait_value32_geq(void* stream, void* addr, unsigned int value);

#ifdef __cplusplus
}
#endif
#endif /* IGI_B200_H */
