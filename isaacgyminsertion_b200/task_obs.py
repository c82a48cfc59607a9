"""Observation members of `FactoryTaskInsertionTactile`, batched.

Mirror of the reference task's observation buffers and functions
(isaacgyminsertion/tasks/factory_tactile/factory_task_insertion.py):
  buffers                         :273-338   tactile_imgs, tactile_queue, pcl, pcl_queue, seg_buf,
                                             image_buf, socket_pcl, got_socket
  update_tactile / _render_tactile :479-583
  update_external_cam              :896-1056
  reset bookkeeping                :1753-1777
  obs_dict assembly                :2126-2144
Same method names and arguments; IsaacGym state (fingertip / plug poses, camera image
tensors) is whatever the caller assigns to the attributes the reference reads
(`left_finger_pos`, ..., `cam_renders`, `seg_renders`), e.g. `synthetic` tensors.

tactile_imgs and pcl are two views of ONE packed row [3*2048 | 2400] per env
(`obs_packed`), so the multi-GPU all-gather (dist.py) sends the kernels' output buffer
directly, with no pack step.
"""
import ctypes as _c

import numpy as np
import torch
from scipy.spatial.transform import Rotation as R

from . import _lib
from .allsight_render import BatchedAllSight, OBS_LEN
from .pcl_utils import CameraPointCloud, filter_pts

TACTILE_FLOATS = 3 * OBS_LEN


class PointCloudAugmentations:
    """factory_utils.py:83-166; only `random_noise` is live (`augment` :157-166)."""

    def __init__(self, num_points=400, sigma=0.001, noise_clip=0.001):
        self.num_points = num_points
        self.sigma = sigma
        self.const_noise = 0.001
        self.noise_clip = noise_clip

    def random_noise(self, pointcloud_batch, pcl_noise, noise_prob=0.3):
        B, N, _ = pointcloud_batch.shape
        pointwise = torch.clamp(torch.randn_like(pointcloud_batch) * self.sigma, -self.noise_clip, self.noise_clip)
        mask = (torch.rand(B, N, device=pointcloud_batch.device) < noise_prob).unsqueeze(-1).float()
        pointcloud_batch = pointcloud_batch + pointwise * mask
        const = torch.clamp(pcl_noise * self.const_noise, -self.noise_clip, self.noise_clip)
        return pointcloud_batch + const

    def augment(self, pointcloud_batch, angle, axes, pcl_noise, dropout_ratio=0.2):
        if not pointcloud_batch.shape[0]:
            return pointcloud_batch
        return self.random_noise(pointcloud_batch, pcl_noise)


class FactoryTaskInsertionTactileObs:
    def __init__(self, num_envs, gym, mesh_ids, bg_ids=None, device="cuda", num_points=400, num_points_socket=400,
                 tact_hist_len=1, pcl_hist_len=1, sampler="reference", tactile=True, pcl_cam=True, kmax=1024,
                 strict_rng=True, pcl_noise_enabled=False, overlap_streams=True, include_all_pcl=False,
                 total_points=2048, falloff=None, global_env_offset=0, total_envs=None):
        self.device = torch.device(device)
        self.num_envs = num_envs
        # env sharding: this object owns global envs [global_env_offset, global_env_offset + num_envs)
        # of a job of total_envs; results do not depend on how the envs are sharded
        self.global_env_offset = int(global_env_offset)
        self.total_envs = int(total_envs) if total_envs is not None else int(num_envs)
        self.fingertips = ["finger_1_3", "finger_2_3", "finger_3_3"]   # factory_env_insertion.py:748
        self.num_points, self.num_points_socket = num_points, num_points_socket
        # pcl row = [plug | socket | all-scene] (pcl_components order, factory_task_insertion.py:1014-1027);
        # include_all_pcl / total_points: FactoryTaskInsertionTactile.yaml:118,124 (off in the shipped config)
        self.include_all_pcl, self.total_points = bool(include_all_pcl), int(total_points)
        self.pcl_floats = (num_points + num_points_socket + (self.total_points if self.include_all_pcl else 0)) * 3
        self.sampler = sampler
        self.strict_rng = strict_rng
        self.pcl_noise_enabled = pcl_noise_enabled   # RNG-defined augmentation
        self.overlap_streams = overlap_streams       # compute_observations: pcl path on a side stream
        self._side = None
        dev = self.device
        N = num_envs
        # packed observation rows: [tactile (3*2048) | pcl (2400)]
        self.obs_packed = torch.zeros((N, TACTILE_FLOATS + self.pcl_floats), dtype=torch.float32, device=dev)
        self.tactile_imgs = self.obs_packed[:, :TACTILE_FLOATS].view(N, 3, OBS_LEN)
        self.pcl = self.obs_packed[:, TACTILE_FLOATS:]
        self.tactile_queue = torch.zeros((N, tact_hist_len, 3, OBS_LEN), dtype=torch.float32, device=dev)
        self.pcl_queue = torch.zeros((N, pcl_hist_len, self.pcl_floats), dtype=torch.float32, device=dev)
        # plug and socket clouds live side by side, [plug | socket] per env = the layout of the merged pcl
        # row (:1014-1027), so one all-class FPS launch writes both and no concatenation is needed
        self._both_pts = (torch.zeros((N, 2, num_points, 3), dtype=torch.float32, device=dev)
                          if num_points == num_points_socket else None)
        self.socket_pcl = (self._both_pts[:, 1] if self._both_pts is not None else
                           torch.zeros((N, num_points_socket, 3), dtype=torch.float32, device=dev))
        self.got_socket = torch.zeros((N, 1), dtype=torch.int32, device=dev)
        self._all_pts = (torch.zeros((N, self.total_points, 3), dtype=torch.float32, device=dev)
                         if self.include_all_pcl else None)
        self._socket_pending = True        # host mirror of `not self.got_socket.all()` (no sync)
        self._socket_force = False         # invalidate_socket_cache(): treat every env as restarted
        self._lib = _lib.load()
        self._upd_seg = torch.zeros(N, dtype=torch.uint8, device=dev)
        self._upd_pcl = torch.zeros(N, dtype=torch.uint8, device=dev)
        self._restarted = torch.zeros(N, dtype=torch.uint8, device=dev)
        self.pcl_pos_noise = torch.randn(N, 1, 3, device=dev)
        self.rot_pcl_angle = torch.zeros(N, device=dev)
        self.rot_axes = torch.zeros(N, dtype=torch.long, device=dev)
        self.pcl_process = PointCloudAugmentations(num_points=num_points)
        self.pcl_rot = 30.0                # cfg_task.randomize.pcl_rot, degrees (FactoryTaskInsertionTactile.yaml:154)
        self.res = [gym.width, gym.height]
        self.seg_buf = torch.zeros(N, self.res[1] * self.res[0], dtype=torch.int32, device=dev)
        self.tactile = tactile
        self.pcl_cam = pcl_cam
        if tactile:
            self.tactile_engine = BatchedAllSight(N, mesh_ids, bg_ids, device=dev, kmax=kmax, falloff=falloff)
            self.tactile_handles = None    # built lazily by `handles()`
        if pcl_cam:
            self.pcl_generator = CameraPointCloud(None, gym, gym.envs, gym.camera_handles, gym.camera_props,
                                                  sample_num=num_points, filter_func=filter_pts, pt_in_local=True,
                                                  graphics_device=dev, compute_device=dev, sampler=sampler)
            self.pcl_generator.engine.sharded = self.total_envs != self.num_envs
            self._plug_pts = (self._both_pts[:, 0] if self._both_pts is not None else
                              torch.zeros((N, num_points, 3), dtype=torch.float32, device=dev))
        # state the reference reads from gym tensors
        z3 = torch.zeros((N, 3), device=dev)
        q4 = torch.tensor([0, 0, 0, 1.0], device=dev).repeat(N, 1)
        self.left_finger_pos, self.right_finger_pos, self.middle_finger_pos = z3.clone(), z3.clone(), z3.clone()
        self.left_finger_quat, self.right_finger_quat, self.middle_finger_quat = q4.clone(), q4.clone(), q4.clone()
        self.plug_pos, self.plug_quat = z3.clone(), q4.clone()
        self.finger_normalized_forces = torch.zeros((N, 3), device=dev)
        self.tactile_wrt_force = False
        self.cam_renders = None            # (N,H,W) f32 depth  (torch.stack(self.cam_renders) in the reference)
        self.seg_renders = None            # (N,H,W) i32 segmentation

    # ------------------------------------------------------------------ tactile
    def handles(self):
        if self.tactile_handles is None:
            self.tactile_handles = self.tactile_engine.handles()
        return self.tactile_handles

    @torch.no_grad()
    def update_tactile(self, update_freq, update_delay):
        """factory_task_insertion.py:479-513, poses stay on the device in f32.  Two launches of this library's
        kernels + the queue push; the fingertip views and the two masks are passed as they are (no stack /
        logical_and / copy launches)."""
        fpos = (self.left_finger_pos, self.right_finger_pos, self.middle_finger_pos)
        fquat = (self.left_finger_quat, self.right_finger_quat, self.middle_finger_quat)
        force = (100 * self.finger_normalized_forces) if self.tactile_wrt_force else None   # :532-535
        self.tactile_engine.render(fpos, fquat, self._dense(self.plug_pos), self._dense(self.plug_quat),
                                   force=force, update=update_freq, update2=update_delay, obs_out=self.tactile_imgs)
        # tactile_queue[:, 1:] = tactile_queue[:, :-1]; tactile_queue[:, 0] = tactile_imgs   (:512-513)
        with _lib.on_device(self.device):
            _lib.check(self._lib.igi_queue_push(
                _c.c_void_p(self.tactile_queue.data_ptr()), _c.c_void_p(self.tactile_imgs.data_ptr()), _c.c_int(0),
                _c.c_int64(self.obs_packed.stride(0)), _c.c_int(self.num_envs), _c.c_int(self.tactile_queue.shape[1]),
                _c.c_longlong(TACTILE_FLOATS), _lib.stream_ptr(self.device)), "igi_queue_push")

    @staticmethod
    def _dense(t):
        return t if t.is_contiguous() and t.dtype == torch.float32 else t.contiguous().float()

    @staticmethod
    def _u8(m):
        """bool (1 byte per env) / uint8 mask -> contiguous uint8 tensor without a copy when possible."""
        if m.dtype == torch.bool and m.is_contiguous():
            return m.view(torch.uint8)
        return m.to(torch.uint8).contiguous()

    @torch.no_grad()
    def _render_tactile(self, left_finger_pose, right_finger_pose, middle_finger_pose, object_pose, update_freq,
                        update_delay):
        """factory_task_insertion.py:515-583 with its (N,4,4) host matrices."""
        def pq(T):
            T = np.asarray(T, dtype=np.float64)
            return T[:, :3, 3].astype(np.float32), R.from_matrix(T[:, :3, :3]).as_quat().astype(np.float32)
        ps, qs = zip(*(pq(T) for T in (left_finger_pose, right_finger_pose, middle_finger_pose)))
        op, oq = pq(object_pose)
        dev = self.device
        update = torch.logical_and(torch.as_tensor(update_freq, device=dev), torch.as_tensor(update_delay, device=dev))
        force = (100 * self.finger_normalized_forces) if self.tactile_wrt_force else None
        self.tactile_engine.render(torch.from_numpy(np.stack(ps, 1)).to(dev), torch.from_numpy(np.stack(qs, 1)).to(dev),
                                   torch.from_numpy(op).to(dev), torch.from_numpy(oq).to(dev), force=force,
                                   update=update, obs_out=self.tactile_imgs)

    # ------------------------------------------------------------------ external camera
    @torch.no_grad()
    def update_external_cam(self, update_freq, update_delay, seg_update_delay, seg_noise, pcl_noise):
        """factory_task_insertion.py:896-1056 (cam_type 'd', seg_cam + pcl_cam, merge_socket_pcl,
        include_plug_pcl — the shipped student configuration).  Steady state = 5 launches of this library's
        kernels (masks, seg_buf rows, compaction, FPS, assemble + queue) and no eager torch op."""
        depth, seg = self.cam_renders, self.seg_renders
        N, dev, lib = self.num_envs, self.device, self._lib
        st = _lib.stream_ptr(dev)
        compute_socket = bool(self._socket_pending)
        f, d, sd = self._u8(update_freq), self._u8(update_delay), self._u8(seg_update_delay)
        # upd_seg = freq & seg_delay; restarted = socket pending & got_socket == 0 (-> 1); upd_pcl = freq & delay | restarted
        with _lib.on_device(dev):
            _lib.check(lib.igi_cam_masks(
                _c.c_void_p(f.data_ptr()), _c.c_void_p(d.data_ptr()), _c.c_void_p(sd.data_ptr()),
                _c.c_void_p(self.got_socket.data_ptr()), _c.c_int((2 if self._socket_force else 1) if compute_socket else 0),
                _c.c_void_p(self._upd_seg.data_ptr()), _c.c_void_p(self._upd_pcl.data_ptr()),
                _c.c_void_p(self._restarted.data_ptr()), _c.c_int(N), st), "igi_cam_masks")
        segf = seg.reshape(N, -1)
        if segf.dtype != torch.int32 or not segf.is_contiguous():
            segf = segf.to(torch.int32).contiguous()
        row = self.seg_buf.shape[1] * 4
        with _lib.on_device(dev):
            _lib.check(lib.igi_copy_rows_where(                                                        # :934-940
                _c.c_void_p(self.seg_buf.data_ptr()), _c.c_longlong(row), _c.c_void_p(segf.data_ptr()), _c.c_longlong(row),
                _c.c_void_p(self._upd_seg.data_ptr()), _c.c_int(1), _c.c_longlong(N), _c.c_longlong(row), st),
                "igi_copy_rows_where")
        gen = self.pcl_generator
        box = filter_pts.box
        all_pts = None
        if self.include_all_pcl:                                                                    # :946-954
            # the whole scene: unmasked depth, same box filter, total_points samples.  The reference draws
            # this cloud BEFORE the plug and socket clouds, so the index stream is consumed in that order.
            a_pts, a_cnt, a_any = gen.engine.compact(depth, None, (0,), box, tag="all")
            self._sample(a_pts, a_cnt, a_any, 0, self.total_points, self._all_pts)
            all_pts = self._all_pts
            if self.pcl_noise_enabled:
                all_pts = torch.where(pcl_noise[:, None, None], self.pcl_process.augment(
                    all_pts, self.rot_pcl_angle, self.rot_axes, self.pcl_pos_noise), all_pts)
        pts, cnt, any_ = gen.engine.compact(depth, seg, (2, 3) if compute_socket else (2,), box)     # :956-959,975
        fused = compute_socket and self.sampler == "fps" and self._both_pts is not None
        if fused:    # plug and socket tasks in one size-ordered launch
            gen.engine.sample_fps(pts, cnt, any_, None, self.num_points, out=self._both_pts)
        else:
            self._sample(pts, cnt, any_, 0, self.num_points, self._plug_pts)                         # :961-964
        plug_pts = self._plug_pts
        if self.pcl_noise_enabled:       # RNG-defined augmentation: torch ops, off in the shipped path
            plug_pts = torch.where(pcl_noise[:, None, None], self.pcl_process.augment(
                plug_pts, self.rot_pcl_angle, self.rot_axes, self.pcl_pos_noise), plug_pts)          # :966-969
        if compute_socket:                                                                          # :972-989
            if not fused:
                self._sample(pts, cnt, any_, 1, self.num_points_socket, self.socket_pcl)
            if self.pcl_noise_enabled:
                noisy = pcl_noise | self._restarted.bool()
                self.socket_pcl.copy_(torch.where(noisy[:, None, None], self.pcl_process.augment(
                    self.socket_pcl, self.rot_pcl_angle, self.rot_axes, self.pcl_pos_noise), self.socket_pcl))
            self._socket_pending = False
            self._socket_force = False
        if all_pts is not None:
            merged = torch.cat([plug_pts, self.socket_pcl, all_pts], dim=1).flatten(start_dim=1)     # :1014-1027
        elif plug_pts is self._plug_pts and self._both_pts is not None:
            merged = self._both_pts.view(N, -1)
        else:
            merged = torch.cat([plug_pts, self.socket_pcl], dim=1).flatten(start_dim=1)
        # pcl[update] = merged[update] (:1027); pcl_queue[:, 1:] = pcl_queue[:, :-1]; pcl_queue[:, 0] = pcl (:1046-1048)
        with _lib.on_device(dev):
            _lib.check(lib.igi_pcl_assemble(
                _c.c_void_p(merged.data_ptr()), _c.c_int64(merged.stride(0)), _c.c_void_p(self.pcl.data_ptr()),
                _c.c_int64(self.pcl.stride(0)), _c.c_void_p(self._upd_pcl.data_ptr()), _c.c_void_p(self.pcl_queue.data_ptr()),
                _c.c_int(N), _c.c_int(self.pcl_queue.shape[1]), _c.c_longlong(self.pcl_floats), st), "igi_pcl_assemble")

    def invalidate_socket_cache(self):
        """Every env recomputes its socket cloud on the next step (what a reset of all envs does to got_socket),
        without a fill launch: the masks kernel is told to treat every env as restarted."""
        self._socket_pending = True
        self._socket_force = True

    # ------------------------------------------------------------------ both parts of one env step
    @torch.no_grad()
    def compute_observations(self, tactile_update_freq, tactile_update_delay, img_update_freq, img_update,
                             seg_update, seg_add_noise, pcl_add_noise):
        """Observation part of `compute_observations` (factory_task_insertion.py:862-887): `update_tactile`
        then `update_external_cam` with the masks the reference draws there.  The two parts share no
        data, so the point-cloud kernels are enqueued on a side stream and share the SMs with the tactile
        kernels (FPS is issue-bound, `tac_geom` / `tac_contact` latency-bound: together they use issue slots
        either leaves idle); the current stream waits for both before returning, so callers see the
        reference's ordering."""
        both = self.tactile and self.pcl_cam and self.overlap_streams
        if not both:
            if self.tactile:
                self.update_tactile(tactile_update_freq, tactile_update_delay)
            if self.pcl_cam:
                self.update_external_cam(img_update_freq, img_update, seg_update, seg_add_noise, pcl_add_noise)
            return self.obs_packed
        cur = torch.cuda.current_stream(self.device)
        if self._side is None:
            # same priority as the caller's stream: CTAs of the two paths then interleave on every SM.  A high-priority
            # side stream lets the FPS kernel take every CTA slot while it runs, which serialises the two paths
            # (measured at 4096 envs: 2.52 ms/step with priority -1 or without the side stream, 2.36 ms with priority 0)
            self._side = torch.cuda.Stream(device=self.device)
            self._ev_fork, self._ev_join = torch.cuda.Event(), torch.cuda.Event()
        self._ev_fork.record(cur)
        self._side.wait_event(self._ev_fork)
        with torch.cuda.stream(self._side):
            self.update_external_cam(img_update_freq, img_update, seg_update, seg_add_noise, pcl_add_noise)
            self._ev_join.record(self._side)
        self.update_tactile(tactile_update_freq, tactile_update_delay)
        cur.wait_event(self._ev_join)
        return self.obs_packed

    def _sample(self, pts, cnt, any_, cls, m, out):
        eng = self.pcl_generator.engine
        if self.sampler == "fps":
            eng.sample_fps(pts, cnt, any_, cls, m, out=out)
        else:
            eng.sample_reference(pts, cnt, any_, cls, m, out=out, strict_rng=self.strict_rng)

    # ------------------------------------------------------------------ reset / obs
    def reset_idx(self, env_ids):
        """Observation part of the reset (factory_task_insertion.py:1753-1777): queues AND the current
        buffers of the reset envs are zeroed, so an env whose update flag is off on the next step shows
        zeros (the reference copies `tactile_queue[e, 0]`, which is 0 after a reset, :578-579), and the
        per-env point-cloud augmentation state is redrawn with the reference's three calls in its order
        (CPU generator: uniform angles; CUDA generator: position noise, rotation axes)."""
        if self.tactile:
            self.tactile_queue[env_ids] = 0
            self.tactile_imgs[env_ids] = 0.
        self.seg_buf[env_ids] = 0
        if self.pcl_cam:
            N, dev = self.num_envs, self.device
            rand_angles = torch.FloatTensor(N).uniform_(-self.pcl_rot, self.pcl_rot).to(dev)
            self.rot_pcl_angle[env_ids] = torch.deg2rad(rand_angles)[env_ids]
            self.pcl_pos_noise[env_ids] = torch.randn(N, 1, 3, device=dev)[env_ids]
            self.rot_axes[env_ids] = torch.randint(0, 3, (N,), device=dev)[env_ids]
            self.pcl_queue[env_ids] = 0
            self.pcl[env_ids] = 0
            self.got_socket[env_ids] = 0
            self.socket_pcl[env_ids] = 0
            self._socket_pending = True

    def obs_dict(self, rl_device=None):
        """factory_task_insertion.py:2126-2144."""
        dev = rl_device or self.device
        return {"tactile": self.tactile_queue.clone().to(dev), "pcl": self.pcl_queue.clone().to(dev)}
