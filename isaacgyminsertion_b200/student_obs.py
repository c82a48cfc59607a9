"""Stages next to the hot path, same names and signatures as the reference's:

  RunningMeanStd         algo/models/running_mean_std.py:22-93      (student normaliser)
  TactileTransform       algo/models/transformer/utils.py:131-156   (batched, identity-size short cut)
  process_obs            algo/ext_adapt/ext_adapt.py:383-435        (ExtrinsicAdapt.process_obs)
  DepthImageProcessor    tasks/factory_tactile/factory_utils.py:12-72
  PointCloudAugmentations.random_noise   factory_utils.py:83-100
  CamImageObs            depth_cam / seg_cam branches of update_external_cam
                         (factory_task_insertion.py:925-943) + img / seg history queues (:1050-1056)
  queue_push             q[:, 1:] = q[:, :-1].clone(); q[:, 0] = x  (:512-513, :1046-1056)

All compute goes through the C-ABI (csrc/student.cu); CPU tensors raise.  The RNG-defined stages use
a counter-based Philox keyed by (seed, step, global element index), so a sharded run reproduces the
single-GPU result; the reference draws from torch's global generator (distribution-level parity).
"""
import ctypes as _c

import torch

from . import _lib


def _stream(dev):
    return _lib.stream_ptr(dev)


def _u8(mask, n, dev, name):
    if mask is None:
        return None
    m = mask.to(device=dev, dtype=torch.uint8).contiguous()
    if m.numel() != n:
        raise RuntimeError(f"{name}: expected {n} entries, got {m.numel()}")
    return m


def queue_push(queue, x):
    """queue (N, T, L) f32 <- x (N, L) f32 or i32, rows may be strided (a view of a packed buffer)."""
    lib = _lib.load()
    N, T, L = queue.shape
    if x.shape != (N, L) or x.stride(1) != 1:
        raise RuntimeError("queue_push: x must be (N, L) with unit inner stride")
    if x.dtype not in (torch.float32, torch.int32):
        raise RuntimeError("queue_push: x must be float32 or int32")
    if not x.is_cuda:
        raise RuntimeError("queue_push: expected a CUDA tensor (no CPU fallback)")
    rc = lib.igi_queue_push(_lib.dptr(queue, torch.float32, "queue"), _c.c_void_p(x.data_ptr()),
                            _c.c_int(1 if x.dtype == torch.int32 else 0), _c.c_int64(x.stride(0)), _c.c_int(N),
                            _c.c_int(T), _c.c_longlong(L), _stream(queue.device))
    _lib.check(rc, "igi_queue_push")
    return queue


class RunningMeanStd:
    """running_mean_std.py:22-93 for per_channel=False inputs of shape (rows, C), C = insize <= 256."""

    def __init__(self, insize, epsilon=1e-05, per_channel=False, norm_only=False, device="cuda"):
        if per_channel:
            raise NotImplementedError("per_channel=True is not used on the student path (ext_adapt.py:120-135)")
        self.insize = insize
        C = int(insize[0] if isinstance(insize, (tuple, list)) else insize)
        self.channels = C
        self.epsilon = epsilon
        self.norm_only = norm_only
        self.per_channel = per_channel
        self.training = True
        dev = torch.device(device)
        self.running_mean = torch.zeros(C, dtype=torch.float64, device=dev)
        self.running_var = torch.ones(C, dtype=torch.float64, device=dev)
        self.count = torch.ones((), dtype=torch.float64, device=dev)
        lib = _lib.load()
        nbytes = lib.igi_rms_scratch_bytes(C)
        if nbytes < 0:
            raise RuntimeError("RunningMeanStd: insize must be 1..256")
        self._scratch = torch.zeros(nbytes, dtype=torch.uint8, device=dev)

    def train(self, mode=True):
        self.training = bool(mode)
        return self

    def eval(self):
        return self.train(False)

    def state_dict(self):
        return {"running_mean": self.running_mean, "running_var": self.running_var, "count": self.count}

    def load_state_dict(self, sd):
        self.running_mean.copy_(sd["running_mean"])
        self.running_var.copy_(sd["running_var"])
        self.count.copy_(sd["count"])

    @torch.no_grad()
    def forward(self, input, unnorm=False, out=None):
        lib = _lib.load()
        C = self.channels
        if input.shape[-1] != C or input.dim() != 2:
            raise RuntimeError(f"RunningMeanStd: expected (rows, {C}), got {tuple(input.shape)}")
        x = input if input.is_contiguous() else input.contiguous()
        y = torch.empty_like(x) if out is None else out
        mode = 2 if unnorm else (1 if self.norm_only else 0)
        rc = lib.igi_rms_forward(_lib.dptr(x, torch.float32, "input"), _c.c_longlong(x.shape[0]), _c.c_int(C),
                                 _lib.dptr(self.running_mean), _lib.dptr(self.running_var), _lib.dptr(self.count),
                                 _c.c_float(self.epsilon), _c.c_int(1 if self.training else 0), _c.c_int(mode),
                                 _lib.dptr(y, torch.float32, "out"), _lib.dptr(self._scratch), _stream(x.device))
        _lib.check(rc, "igi_rms_forward")
        return y

    __call__ = forward


@torch.no_grad()
def process_obs(obs, pcl_mean_std=None, stud_obs_mean_std=None, obj_id=2, socket_id=3, distinct=True):
    """ExtrinsicAdapt.process_obs (ext_adapt.py:383-435): seg / img masking, pcl (and student_obs)
    normalisation.  Keys that are absent stay None, as with the *_info switches of the reference."""
    lib = _lib.load()
    student_obs = obs.get("student_obs")
    tactile = obs.get("tactile")
    img, seg, pcl = obs.get("img"), obs.get("seg"), obs.get("pcl")
    if seg is not None:
        seg_c = seg.contiguous()
        img_c = img.contiguous() if img is not None else None
        seg_out = torch.empty_like(seg_c)
        img_out = torch.empty_like(img_c) if img_c is not None else None
        rc = lib.igi_seg_valid_mask(_lib.dptr(seg_c, torch.float32, "seg"), _lib.dptr(img_c, torch.float32, "img"),
                                    _c.c_longlong(seg_c.numel()), _c.c_float(obj_id), _c.c_float(socket_id),
                                    _c.c_int(1 if distinct else 0), _lib.dptr(seg_out), _lib.dptr(img_out),
                                    _stream(seg_c.device))
        _lib.check(rc, "igi_seg_valid_mask")
        seg, img = seg_out, img_out
    if pcl is not None and pcl_mean_std is not None:
        # [B, T, N*3] -> [B*T*N, 3] -> normalise -> [B, T*N, 3]
        B = pcl.shape[0]
        pcl = pcl_mean_std(pcl.reshape(-1, 3)).reshape((B, -1, 3))
    if student_obs is not None and stud_obs_mean_std is not None:
        student_obs = stud_obs_mean_std(student_obs)
    return {"student_obs": student_obs, "tactile": tactile, "img": img, "seg": seg, "pcl": pcl}


class TactileTransform:
    """algo/models/transformer/utils.py:131-156 without the per-image Python loop (3*N*T iterations per step).

    The student's eval transform is Resize((w, h), bilinear) -> CenterCrop((cw, ch)) (`define_tactile_transforms`,
    utils.py:217-274).  Both are pure per-image functions, so applying them once to the flattened
    (B*T*F, C, H, W) batch gives the same pixels as the reference's loop; at the shipped sizes (32 x 64 in,
    32 x 64 out; FactoryTaskInsertionTactile.yaml:39-46) they are the identity and nothing is launched at all.
    `identity_size=(H, W)`: the (resize == crop) size for which the transform may be skipped."""

    def __init__(self, tactile_transform=None, identity_size=None):
        self.tactile_transform = tactile_transform
        self.identity_size = tuple(identity_size) if identity_size is not None else None

    def __call__(self, tac_input):
        B, T, F, C, H, W = tac_input.shape
        if self.tactile_transform is None or (H, W) == self.identity_size:
            return tac_input
        out = self.tactile_transform(tac_input.reshape(-1, C, H, W))
        return out.view(B, T, F, C, *out.shape[2:])


class DepthImageProcessor:
    """factory_utils.py:12-72.  `step` advances once per call so successive calls draw fresh numbers."""

    def __init__(self, cfg, dis_noise, far_clip, near_clip, seed=0, env0=0):
        self.cfg = cfg
        self.dis_noise, self.far_clip, self.near_clip = dis_noise, far_clip, near_clip
        self.seed, self.env0, self.step = int(seed), int(env0), 0

    def _call(self, depth, seg, update, update_seg, seg_noise, image_buf, seg_buf, flip_prob):
        lib = _lib.load()
        ref = depth if depth is not None else seg
        n = ref.shape[0]
        npix = ref[0].numel()
        dev = ref.device
        # the converted masks must stay referenced until the launch is enqueued: a temporary freed between
        # two conversions hands the same block of the caching allocator to the next one
        m_upd, m_seg, m_noise = (_u8(update, n, dev, "update"), _u8(update_seg, n, dev, "update_seg"),
                                 _u8(seg_noise, n, dev, "seg_noise"))
        rc = lib.igi_cam_image_obs(
            _lib.dptr(depth, torch.float32, "depth"), _lib.dptr(seg, torch.int32, "seg"),
            _lib.dptr(m_upd), _lib.dptr(m_seg),
            _lib.dptr(m_noise), _c.c_int(n), _c.c_int(npix), _c.c_longlong(self.env0),
            _c.c_double(self.dis_noise), _c.c_double(self.far_clip), _c.c_double(self.near_clip), _c.c_float(flip_prob),
            _c.c_uint64(self.seed), _c.c_uint32(self.step), _lib.dptr(image_buf, torch.float32, "image_buf"),
            _lib.dptr(seg_buf, torch.int32, "seg_buf"), _stream(dev))
        _lib.check(rc, "igi_cam_image_obs")
        self.step += 1

    @torch.no_grad()
    def process_depth_image(self, depth_images):
        """noise, clip to [-far, -near], normalise to [0, 1] (:55-72); returns a new tensor."""
        d = depth_images.contiguous()
        out = torch.empty_like(d)
        self._call(d.reshape(d.shape[0], -1), None, None, None, None, out.view(d.shape[0], -1), None, 0.0)
        return out.squeeze(0) if out.size(0) == 1 else out

    @torch.no_grad()
    def add_seg_noise(self, seg_images_to_noise, flip_prob=0.1):
        """Object pixels (> 0) flip to background with probability flip_prob, in place (:23-37)."""
        s = seg_images_to_noise
        if not s.is_contiguous():
            raise RuntimeError("add_seg_noise: tensor must be contiguous (it is modified in place)")
        n = s.shape[0]
        ones = torch.ones(n, dtype=torch.uint8, device=s.device)
        # depth pointer is required by the entry point but unused when image_buf is NULL
        self._call(s.view(n, -1).view(torch.float32), s.view(n, -1), None, ones, ones, None, s.view(n, -1), flip_prob)
        return s

    def normalize_depth_image(self, depth_images):
        depth_images = depth_images * -1
        return (depth_images - self.near_clip) / (self.far_clip - self.near_clip)


class PointCloudAugmentations:
    """factory_utils.py:83-166; only `random_noise` is live (`augment` :157-166)."""

    def __init__(self, num_points=400, sigma=0.001, noise_clip=0.001, seed=0, env0=0):
        self.num_points = num_points
        self.sigma = sigma
        self.const_noise = 0.001
        self.noise_clip = noise_clip
        self.seed, self.env0, self.step = int(seed), int(env0), 0

    @torch.no_grad()
    def random_noise(self, pointcloud_batch, pcl_noise, noise_prob=0.3, mask=None):
        """In place on (B, N, 3); `mask` (B,) restricts it to the envs the reference indexes with
        `pts[pcl_noise]` (factory_task_insertion.py:966-969)."""
        lib = _lib.load()
        p = pointcloud_batch
        B, N, _ = p.shape
        if p.stride(2) != 1 or p.stride(1) != 3:
            raise RuntimeError("random_noise: points must be (B, N, 3) with packed rows")
        noise = pcl_noise.reshape(B, 3).to(torch.float32).contiguous()
        m = _u8(mask, B, p.device, "mask")   # kept referenced until the launch is enqueued
        rc = lib.igi_pcl_noise(_c.c_void_p(p.data_ptr()), _c.c_int64(p.stride(0)), _c.c_int(B), _c.c_int(N),
                               _lib.dptr(m), _lib.dptr(noise), _c.c_longlong(self.env0),
                               _c.c_float(self.sigma), _c.c_float(self.noise_clip), _c.c_float(self.const_noise),
                               _c.c_float(noise_prob), _c.c_uint64(self.seed), _c.c_uint32(self.step),
                               _stream(p.device))
        _lib.check(rc, "igi_pcl_noise")
        self.step += 1
        return p

    def augment(self, pointcloud_batch, angle, axes, pcl_noise, dropout_ratio=0.2, mask=None):
        if not pointcloud_batch.shape[0]:
            return pointcloud_batch
        return self.random_noise(pointcloud_batch, pcl_noise, mask=mask)


class CamImageObs:
    """image_buf / seg_buf and their history queues (factory_task_insertion.py:329-338, 925-943, 1050-1056)."""

    def __init__(self, num_envs, npix, img_hist_len=1, dis_noise=0.001, far_clip=0.5, near_clip=0.1, device="cuda",
                 seed=0, env0=0, flip_prob=0.1):
        dev = torch.device(device)
        self.num_envs, self.npix = num_envs, npix
        self.image_buf = torch.zeros(num_envs, npix, device=dev)
        self.seg_buf = torch.zeros(num_envs, npix, dtype=torch.int32, device=dev)
        self.img_queue = torch.zeros((num_envs, img_hist_len, npix), dtype=torch.float32, device=dev)
        self.seg_queue = torch.zeros((num_envs, img_hist_len, npix), dtype=torch.float32, device=dev)
        self.depth_process = DepthImageProcessor(None, dis_noise, far_clip, near_clip, seed=seed, env0=env0)
        self.flip_prob = flip_prob

    @torch.no_grad()
    def update(self, depth, seg, update, update_seg, seg_noise):
        """One launch for :925-943, then the two queue pushes (:1050-1056)."""
        n = self.num_envs
        self.depth_process._call(depth.reshape(n, -1), seg.reshape(n, -1), update, update_seg, seg_noise,
                                 self.image_buf, self.seg_buf, self.flip_prob)
        queue_push(self.img_queue, self.image_buf)
        queue_push(self.seg_queue, self.seg_buf)
        return self.image_buf, self.seg_buf
