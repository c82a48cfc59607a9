"""Drop-in for the reference's `isaacgyminsertion/tasks/utils/pcl_utils.py`.

Same class names, constructor arguments and method signatures
(`PointCloudGenerator` pcl_utils.py:28-97, `CameraPointCloud` pcl_utils.py:100-220);
the per-env Python loops (`get_ptd_cuda` :203-212, `get_point_cloud` :178-183)
are replaced by batched sm_100a kernels behind the C-ABI (include/igi_b200.h).
`BatchedPointCloud` is the batched engine both classes sit on.

No CPU fallback: tensors must live on a CUDA device.
"""
import ctypes

import numpy as np
import torch

from . import _lib

_c = ctypes


class BoxFilter:
    """Axis-aligned inclusive workspace box; calling it filters a (M,3) tensor like
    the reference's module-level `filter_pts` (factory_task_insertion.py:66-77)."""

    def __init__(self, x=(0.1, 0.7), y=(-0.4, 0.4), z=(0.001, 0.6)):
        self.box = (float(x[0]), float(x[1]), float(y[0]), float(y[1]), float(z[0]), float(z[1]))

    @torch.no_grad()
    def __call__(self, pts):
        b = self.box
        x, y, z = pts[:, 0], pts[:, 1], pts[:, 2]
        valid1 = (z >= b[4]) & (z <= b[5])
        valid2 = (x >= b[0]) & (x <= b[1])
        valid3 = (y >= b[2]) & (y <= b[3])
        return pts[valid1 & valid3 & valid2]


#: same name and arity as factory_task_insertion.py:65 `filter_pts(pts)`
filter_pts = BoxFilter()


def _box_of(filter_func):
    if filter_func is None:
        return None
    box = getattr(filter_func, "box", None)
    if box is None:
        raise NotImplementedError(
            "only box filters (BoxFilter / filter_pts) run on the batched kernel path; "
            "wrap the workspace limits in isaacgyminsertion_b200.pcl_utils.BoxFilter")
    return tuple(float(v) for v in box)


class TorchCpuRandintStream:
    """The 32-bit word stream torch's default CPU generator (MT19937) feeds to
    `torch.randint(0, n, (m,))` (pcl_utils.py:199): ATen computes `random() % n`
    per element for n < 2**32 (verified in tests against real torch.randint).

    `peek(k)` returns the next k raw words without advancing torch's generator;
    `commit(k)` advances it by k words, so after a batched call the generator is in
    exactly the state the reference's per-env loop would have left it in.
    """
    _OFF_LEFT = 8
    _OFF_NEXT = 16
    _OFF_STATE = 24
    _N = 624

    def __init__(self, generator=None):
        self.generator = generator

    def _get(self):
        st = (self.generator.get_state() if self.generator is not None else torch.get_rng_state()).numpy().copy()
        left = int(st[self._OFF_LEFT:self._OFF_LEFT + 4].view(np.int32)[0])
        key = st[self._OFF_STATE:self._OFF_STATE + 8 * self._N].view(np.uint64).astype(np.uint32)
        bg = np.random.MT19937()
        bg.state = {"bit_generator": "MT19937", "state": {"key": key, "pos": 625 - left}}
        return st, bg

    def peek(self, k):
        _, bg = self._get()
        return bg.random_raw(int(k)).astype(np.uint32)

    def commit(self, k):
        if k <= 0:
            return
        st, bg = self._get()
        bg.random_raw(int(k))
        s = bg.state["state"]
        pos = int(s["pos"])
        st[self._OFF_STATE:self._OFF_STATE + 8 * self._N] = s["key"].astype(np.uint64).view(np.uint8)
        st[self._OFF_LEFT:self._OFF_LEFT + 4] = np.array([625 - pos], dtype=np.int32).view(np.uint8)
        st[self._OFF_NEXT:self._OFF_NEXT + 8] = np.array([pos], dtype=np.uint64).view(np.uint8)
        t = torch.from_numpy(st)
        if self.generator is not None:
            self.generator.set_state(t)
        else:
            torch.set_rng_state(t)


def uv_table(proj_matrix, width, height):
    """`_uv_one_in_cam` exactly as PointCloudGenerator.__init__ builds it
    (pcl_utils.py:37-57), with the same torch ops on the host: (H,W,3) f32."""
    proj_matrix = np.asarray(proj_matrix)
    fu = 2 / proj_matrix[0, 0]
    fv = 2 / proj_matrix[1, 1]
    fu = width / fu
    fv = height / fv
    cu = width / 2.
    cv = height / 2.
    int_mat = torch.Tensor([[-fu, 0, cu], [0, fv, cv], [0, 0, 1]])
    int_mat_T_inv = torch.inverse(int_mat.T)
    x, y = torch.meshgrid(torch.arange(height), torch.arange(width), indexing="ij")
    uv_one = torch.stack((y, x, torch.ones_like(x)), dim=-1).float()
    return uv_one @ int_mat_T_inv


class BatchedPointCloud:
    """Batched engine: per-env camera tables on the device + the K4/K5 kernels."""

    def __init__(self, proj_matrices, view_matrices, env_to_globals, width, height, depth_max=1.0,
                 device="cuda"):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("BatchedPointCloud needs a CUDA device (no CPU fallback)")
        self.lib = _lib.load()
        self.num_envs = len(view_matrices)
        self.W, self.H = int(width), int(height)
        self.depth_max = depth_max
        n = self.num_envs
        uvx = torch.empty(n, self.W)
        uvy = torch.empty(n, self.H)
        uvz = torch.empty(n)
        ext = torch.empty(n, 4, 4)
        e2g = torch.empty(n, 4, 4)
        cache = {}
        for i in range(n):
            P = np.asarray(proj_matrices[i])
            key = (float(P[0, 0]), float(P[1, 1]))
            tab = cache.get(key)
            if tab is None:
                tab = uv_table(P, self.W, self.H)
                # the table is separable and its third component constant: keep the factors
                if not (torch.equal(tab[:, :, 0], tab[0:1, :, 0].expand(self.H, -1))
                        and torch.equal(tab[:, :, 1], tab[:, 0:1, 1].expand(-1, self.W))
                        and bool((tab[:, :, 2] == tab[0, 0, 2]).all())):
                    raise RuntimeError("uv table is not separable; unsupported intrinsics")
                cache[key] = tab
            uvx[i] = tab[0, :, 0]
            uvy[i] = tab[:, 0, 1]
            uvz[i] = tab[0, 0, 2]
            ext[i] = torch.inverse(torch.Tensor(np.asarray(view_matrices[i])))          # pcl_utils.py:48
            e2g[i] = torch.inverse(torch.Tensor(np.asarray(env_to_globals[i])))         # pcl_utils.py:84
        self.uvx = uvx.to(self.device).contiguous()
        self.uvy = uvy.to(self.device).contiguous()
        self.uvz = uvz.to(self.device).contiguous()
        self.ext = ext.reshape(n, 16).to(self.device).contiguous()
        self.e2g_inv = e2g.reshape(n, 16).to(self.device).contiguous()
        self._scratch = {}
        self.rng_stream = TorchCpuRandintStream()
        # True when this engine holds one rank's contiguous env slice of a larger job (torch.distributed
        # initialised): the reference sampler then takes its index words by GLOBAL env order.
        self.sharded = False

    # -- scratch ---------------------------------------------------------------
    def _buf(self, name, shape, dtype):
        key = (name, tuple(shape), dtype)
        t = self._scratch.get(key)
        if t is None:
            t = torch.empty(shape, dtype=dtype, device=self.device)
            self._scratch[key] = t
        return t

    def _tables(self, env_ids):
        if env_ids is None:
            return self.uvx, self.uvy, self.uvz, self.ext, self.e2g_inv
        idx = torch.as_tensor(env_ids, device=self.device, dtype=torch.long)
        return (self.uvx[idx].contiguous(), self.uvy[idx].contiguous(), self.uvz[idx].contiguous(),
                self.ext[idx].contiguous(), self.e2g_inv[idx].contiguous())

    # -- K4 ----------------------------------------------------------------------
    @torch.no_grad()
    def compact(self, depth, seg=None, seg_ids=(0,), box=None, env_ids=None, tag="c"):
        """depth (n,H,W)|(n,H*W) f32, seg same shape i32 or None ->
        (pts (n,C,H*W,3) f32, count (n,C) i32, any (n,C) i32); rows beyond count are stale."""
        n = depth.shape[0]
        npix = self.H * self.W
        depth = depth.reshape(n, npix)
        if seg is not None:
            seg = seg.reshape(n, npix)
        C = len(seg_ids)
        uvx, uvy, uvz, ext, e2g = self._tables(env_ids)
        if uvx.shape[0] != n:
            raise RuntimeError(f"depth batch {n} != number of cameras {uvx.shape[0]}")
        pts = self._buf(tag + "_pts", (n, C, npix, 3), torch.float32)
        cnt = self._buf(tag + "_cnt", (n, C), torch.int32)
        any_ = self._buf(tag + "_any", (n, C), torch.int32)
        box_arr = _lib.carr(_c.c_float, box) if box is not None else None
        with _lib.on_device(self.device):
            rc = self.lib.igi_pcl_compact(
                _lib.dptr(depth, torch.float32, "depth"), _lib.dptr(seg, torch.int32, "seg"),
                _lib.carr(_c.c_int32, [int(s) for s in seg_ids]), _c.c_int(C),
                _lib.dptr(uvx), _lib.dptr(uvy), _lib.dptr(uvz), _lib.dptr(ext), _lib.dptr(e2g),
                _c.c_int(n), _c.c_int(self.H), _c.c_int(self.W),
                _c.c_float(-1.0 if self.depth_max is None else float(self.depth_max)),
                box_arr, _lib.dptr(pts), _lib.dptr(cnt), _lib.dptr(any_), _lib.stream_ptr(self.device))
        _lib.check(rc, "igi_pcl_compact")
        return pts, cnt, any_

    # -- K5A ---------------------------------------------------------------------
    @torch.no_grad()
    def sample_reference(self, pts, cnt, any_, cls, m, out=None, return_idx=False, strict_rng=True):
        """ids = torch.randint(0, count, (m,)) per non-empty env in env order, drawn from
        torch's default CPU generator exactly as pcl_utils.py:178-183,195-201 does."""
        n, C, cap, _ = pts.shape
        if out is None:
            out = torch.empty((n, m, 3), dtype=torch.float32, device=self.device)
        assert out.stride(-1) == 1 and out.stride(-2) == 3
        skip = total_words = 0
        if self.sharded:
            # The reference draws m words per NON-EMPTY env in global env order (pcl_utils.py:178-183), so a rank
            # starts m x (non-empty envs on lower ranks) words into the stream and the call consumes m x (non-empty
            # envs of the whole job).  Every rank's CPU generator must be in the same state (same torch.manual_seed),
            # as it is in the single-process run.  One tiny all-gather + host read: this sampler is host-synchronous
            # by construction (its indices come from the CPU generator).
            import torch.distributed as dist
            mine = (any_[:, cls] != 0).sum().to(torch.int64).reshape(1)
            every = torch.empty(dist.get_world_size(), dtype=torch.int64, device=self.device)
            dist.all_gather_into_tensor(every, mine)
            from .dist import shard_word_window
            skip, total_words = shard_word_window(every.cpu().tolist(), dist.get_rank(), m)
        raw_np = self.rng_stream.peek(skip + n * m)[skip:]
        raw = torch.from_numpy(raw_np.view(np.int32).copy()).to(self.device, non_blocking=False)
        idx = torch.empty((n, m), dtype=torch.int32, device=self.device) if return_idx else None
        consumed = self._buf("consumed", (1,), torch.int32)
        offs = self._buf("offs", (n,), torch.int32)
        with _lib.on_device(self.device):
            rc = self.lib.igi_pcl_sample_gather(
                _lib.dptr(pts, torch.float32), _lib.dptr(cnt, torch.int32), _lib.dptr(any_, torch.int32),
                _c.c_int(C), _c.c_int(cls), _c.c_int(cap), _lib.dptr(raw), _c.c_int(n), _c.c_int(m),
                _c.c_void_p(out.data_ptr()), _c.c_int64(out.stride(0)), _lib.dptr(idx), _lib.dptr(consumed),
                _lib.dptr(offs), _lib.stream_ptr(self.device))
        _lib.check(rc, "igi_pcl_sample_gather")
        if strict_rng:
            self.rng_stream.commit(total_words if self.sharded else int(consumed.item()))
        return (out, idx) if return_idx else out

    # -- K5B ---------------------------------------------------------------------
    @torch.no_grad()
    def sample_fps(self, pts, cnt, any_, cls, m, out=None, return_idx=False, flags=0):
        """FPS of class `cls` (int) or of every class at once (cls=None: out is (n, C, m, 3)).
        Size-ordered schedule (igi_fps_balanced)."""
        n, C, cap, _ = pts.shape
        if cls is None:
            shape, n_tasks = (n, C, m, 3), n * C
            p0, c0, a0 = pts, cnt, any_
            task_stride, count_stride = pts.stride(1), 1
        else:
            shape, n_tasks = (n, m, 3), n
            p0, c0, a0 = pts[:, cls], cnt[:, cls], any_[:, cls]
            task_stride, count_stride = pts.stride(0), C
        if out is None:
            out = torch.empty(shape, dtype=torch.float32, device=self.device)
        assert tuple(out.shape) == shape and out.stride(-1) == 1 and out.stride(-2) == 3
        out_stride = out.stride(-3)
        if cls is None:
            assert out.stride(0) == C * out_stride, "all-class output must be dense over (env, class)"
        idx = torch.empty(shape[:-1], dtype=torch.int32, device=self.device) if return_idx else None
        scratch = self._buf("fps_sched", (n_tasks + 8,), torch.int32)
        with _lib.on_device(self.device):
            rc = self.lib.igi_fps_balanced(
                _c.c_void_p(p0.data_ptr()), _c.c_int64(task_stride),
                _c.c_void_p(c0.data_ptr()), _c.c_void_p(a0.data_ptr()), _c.c_int64(count_stride),
                _c.c_int(n_tasks), _c.c_int(m), _c.c_void_p(out.data_ptr()), _c.c_int64(out_stride),
                _lib.dptr(idx), _lib.dptr(scratch), _c.c_int(flags), _lib.stream_ptr(self.device))
        _lib.check(rc, "igi_fps_balanced")
        return (out, idx) if return_idx else out

    @torch.no_grad()
    def sample_fps_static(self, pts, cnt, any_, cls, m, out=None, return_idx=False):
        """igi_fps: same results, static task assignment (no scratch)."""
        n, C, cap, _ = pts.shape
        if out is None:
            out = torch.empty((n, m, 3), dtype=torch.float32, device=self.device)
        assert out.stride(-1) == 1 and out.stride(-2) == 3
        idx = torch.empty((n, m), dtype=torch.int32, device=self.device) if return_idx else None
        p0 = pts[:, cls]
        with _lib.on_device(self.device):
            rc = self.lib.igi_fps(
                _c.c_void_p(p0.data_ptr()), _c.c_int64(pts.stride(0)),
                _c.c_void_p(cnt[:, cls].data_ptr()), _c.c_void_p(any_[:, cls].data_ptr()), _c.c_int64(C),
                _c.c_int(0), _c.c_int(n), _c.c_int(m), _c.c_void_p(out.data_ptr()), _c.c_int64(out.stride(0)),
                _lib.dptr(idx), _c.c_int(0), _lib.stream_ptr(self.device))
        _lib.check(rc, "igi_fps")
        return (out, idx) if return_idx else out


@torch.no_grad()
def furthest_point_sample(xyz, npoint, flags=0):
    """pointnet2_ops.pointnet2_utils.furthest_point_sample(xyz (B,N,3) f32 cuda, npoint) -> (B,npoint) i32
    (the call behind `fps()` in algo/models/transformer/point_mae.py:14-21)."""
    lib = _lib.load()
    B, N, _ = xyz.shape
    xyz = xyz.contiguous()
    idx = torch.empty((B, npoint), dtype=torch.int32, device=xyz.device)
    with _lib.on_device(xyz.device):
        rc = lib.igi_fps(_lib.dptr(xyz, torch.float32, "xyz"), _c.c_int64(N * 3), None, None, _c.c_int64(1),
                         _c.c_int(N), _c.c_int(B), _c.c_int(npoint), None, _c.c_int64(npoint * 3),
                         _lib.dptr(idx), _c.c_int(flags), _lib.stream_ptr(xyz.device))
    _lib.check(rc, "igi_fps")
    return idx


class PointCloudGenerator:
    """pcl_utils.py:28-97 — one camera.  `convert` runs the batched kernel with n=1."""

    def __init__(self, proj_matrix, view_matrix, env_to_global, camera_props=None,
                 height=None, width=None, sample_num=None,
                 depth_max=None, device='cpu'):
        self.cam_width = camera_props.width if camera_props is not None else width
        self.cam_height = camera_props.height if camera_props is not None else height
        self.env_to_global = env_to_global
        self.depth_max = depth_max
        self.sample_num = sample_num
        self.device = device
        self._engine = BatchedPointCloud([proj_matrix], [view_matrix], [env_to_global],
                                         self.cam_width, self.cam_height, depth_max=depth_max, device=device)
        self.ext_mat = self._engine.ext.reshape(4, 4)

    @torch.no_grad()
    def convert(self, depth_buffer):
        pts, cnt, _ = self._engine.compact(depth_buffer.reshape(1, -1).contiguous(), None, (0,), None)
        k = int(cnt[0, 0].item())
        return pts[0, 0, :k].clone()

    @torch.no_grad()
    def sample_n(self, pts):
        num = pts.shape[0]
        ids = torch.randint(0, num, size=(self.sample_num,))
        return pts[ids.to(pts.device)]


class CameraPointCloud:
    """pcl_utils.py:100-220.  `isc_gym` only needs get_camera_view_matrix,
    get_camera_proj_matrix and get_env_origin (isaacgym's gym object or
    `synthetic.SyntheticGym`).  `sampler` selects the K5 mode: 'reference'
    (torch.randint stream, default) or 'fps'."""

    def __init__(self, isc_sim, isc_gym, envs, camera_handles,
                 camera_props, sample_num=4000,
                 filter_func=None, pt_in_local=False,
                 depth_max=1.0, graphics_device='cpu',
                 compute_device='cpu', sampler='reference'):
        self.sim = isc_sim
        self.gym = isc_gym
        self.envs = envs
        self.camera_handles = camera_handles
        assert pt_in_local
        self.filter_func = filter_func
        self.camera_props = camera_props
        self.graphics_device = graphics_device
        self.compute_device = compute_device
        self.sample_num = sample_num
        self.num_envs = len(self.envs)
        self.sampler = sampler

        views, projs, e2gs = [], [], []
        for idx in range(len(envs)):
            views.append(np.asarray(self.gym.get_camera_view_matrix(self.sim, envs[idx], camera_handles[idx])))
            env_position = self.gym.get_env_origin(envs[idx])
            env_to_global = np.identity(4)
            env_to_global[:3, 3] = np.array([env_position.x, env_position.y, env_position.z])
            e2gs.append(env_to_global)
            projs.append(np.asarray(self.gym.get_camera_proj_matrix(self.sim, envs[idx], camera_handles[idx])))
        self.engine = BatchedPointCloud(projs, views, e2gs, camera_props[0].width, camera_props[0].height,
                                        depth_max=depth_max, device=self.graphics_device)

    @torch.no_grad()
    def compute_view_matrix(self, local_transform):
        p, q = local_transform.p, local_transform.r
        T = torch.eye(4)
        T[:3, :3] = torch.tensor([
            [1 - 2 * q.y ** 2 - 2 * q.z ** 2, 2 * q.x * q.y - 2 * q.w * q.z, 2 * q.x * q.z + 2 * q.w * q.y],
            [2 * q.x * q.y + 2 * q.w * q.z, 1 - 2 * q.x ** 2 - 2 * q.z ** 2, 2 * q.y * q.z - 2 * q.w * q.x],
            [2 * q.x * q.z - 2 * q.w * q.y, 2 * q.y * q.z + 2 * q.w * q.x, 1 - 2 * q.x ** 2 - 2 * q.y ** 2],
        ], dtype=torch.float32)
        T[:3, 3] = torch.tensor([p.x, p.y, p.z], dtype=torch.float32)
        return torch.inverse(T)

    @torch.no_grad()
    def get_point_cloud(self, depths, env_ids=None, filter_func=None, sample_num=None, seg=None, seg_id=None,
                        return_idx=False):
        """(n,H,W) depth -> (n, sample_num, 3).  Extra keywords `seg`/`seg_id` fuse the
        task's `depth * (seg == id)` masking (factory_task_insertion.py:956-959) into the kernel."""
        if filter_func is None:
            filter_func = self.filter_func
        sample_num = self.sample_num if sample_num is None else sample_num
        box = _box_of(filter_func)
        depths = depths.to(self.graphics_device)
        if env_ids is not None:
            sel = torch.as_tensor(env_ids, device=depths.device, dtype=torch.long)
            depths = depths[sel]
            seg = seg[sel] if seg is not None else None
        n = depths.shape[0]
        pts, cnt, any_ = self.engine.compact(depths.reshape(n, -1).contiguous().float(),
                                             None if seg is None else seg.reshape(n, -1).contiguous(),
                                             (int(seg_id) if seg_id is not None else 0,), box, env_ids=env_ids)
        if self.sampler == 'fps':
            res = self.engine.sample_fps(pts, cnt, any_, 0, sample_num, return_idx=return_idx)
        else:
            res = self.engine.sample_reference(pts, cnt, any_, 0, sample_num, return_idx=return_idx)
        if return_idx:
            return res[0].to(self.compute_device), res[1]
        return res.to(self.compute_device).detach()

    @torch.no_grad()
    def get_ptd_cuda(self, depth_imgs, env_ids=None, filter_func=None):
        """List of per-env (M_e,3) clouds (pcl_utils.py:203-212); one batched launch + one count read-back."""
        if filter_func is None:
            filter_func = self.filter_func
        box = _box_of(filter_func)
        depth_imgs = depth_imgs.to(self.graphics_device)
        if env_ids is not None:
            depth_imgs = depth_imgs[torch.as_tensor(env_ids, device=depth_imgs.device, dtype=torch.long)]
        n = depth_imgs.shape[0]
        pts, cnt, _ = self.engine.compact(depth_imgs.reshape(n, -1).contiguous().float(), None, (0,), box,
                                          env_ids=env_ids)
        counts = cnt[:, 0].tolist()
        return [pts[i, 0, :counts[i]].clone() for i in range(n)]

    @torch.no_grad()
    def _proc_pts(self, env_id, depth_images, filter_func=None):
        """pcl_utils.py:186-193: ONE env's depth image (H,W) with that env's camera -> (M,3)."""
        if filter_func is None:
            filter_func = self.filter_func
        box = _box_of(filter_func)
        d = depth_images.to(self.graphics_device).reshape(1, -1).contiguous().float()
        pts, cnt, _ = self.engine.compact(d, None, (0,), box, env_ids=[int(env_id)])
        return pts[0, 0, :int(cnt[0, 0].item())].clone()

    @torch.no_grad()
    def sample_n(self, pts, sample_num=None):
        sample_num = self.sample_num if sample_num is None else sample_num
        num = pts.shape[0]
        ids = torch.randint(0, num, size=(sample_num,))
        return pts[ids.to(pts.device)]

    @torch.no_grad()
    def clone_img_tensor(self, img_tensors, env_ids=None):
        env_iter = range(len(self.envs)) if env_ids is None else env_ids
        return torch.stack([torch.stack(img_tensors[i]) for i in env_iter])
