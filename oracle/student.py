"""CPU oracle for the stages next to the hot path — TEST INFRASTRUCTURE ONLY.

Restates, with CPU torch / numpy in the reference's own op order:
  DepthImageProcessor.process_depth_image / normalize_depth_image / add_seg_noise
                                   isaacgyminsertion/tasks/factory_tactile/factory_utils.py:23-37,55-72
  update_external_cam depth_cam / seg_cam branches     factory_task_insertion.py:925-943
  PointCloudAugmentations.random_noise                  factory_utils.py:93-100
  RunningMeanStd.forward                                algo/models/running_mean_std.py:47-93
  ExtrinsicAdapt.process_obs                            algo/ext_adapt/ext_adapt.py:383-435
  history queues                                        factory_task_insertion.py:512-513,1046-1056

The random draws of the reference come from torch's global generators; here every draw is an explicit
argument (`u`, `z`, ...) so the arithmetic can be pinned, and `philox_*` restates the counter-based
generator of csrc/student.cu (Philox4x32-10, Salmon et al. 2011) that produces those draws on the GPU.

Parity pin: tests/golden/student_golden.npz holds outputs of the REAL reference code
(RunningMeanStd, DepthImageProcessor, PointCloudAugmentations and process_obs executed from
/root/reference with torch.rand* patched to return the Philox draws; tools/make_golden_student.py);
tests/test_oracle_student.py checks this restatement against them bit-for-bit.
"""
import numpy as np
import torch

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)


def philox(idx, step, stream, seed):
    """Philox4x32-10 blocks for counters (idx_lo, idx_hi, step, stream), key = seed -> (n, 4) u32."""
    idx = np.asarray(idx, dtype=np.uint64)
    c0 = idx & MASK
    c1 = idx >> np.uint64(32)
    c2 = np.full_like(c0, np.uint64(step))
    c3 = np.full_like(c0, np.uint64(stream))
    k0, k1 = int(seed) & 0xFFFFFFFF, (int(seed) >> 32) & 0xFFFFFFFF
    for _ in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK
        c0, c1, c2, c3 = hi1 ^ c1 ^ np.uint64(k0), lo1, hi0 ^ c3 ^ np.uint64(k1), lo0
        k0, k1 = (k0 + W0) & 0xFFFFFFFF, (k1 + W1) & 0xFFFFFFFF
    return np.stack([c0, c1, c2, c3], axis=-1).astype(np.uint32)


def u01(r):
    return ((r >> np.uint32(8)).astype(np.float32)) * np.float32(1.0 / 16777216.0)


def u01_open(r):
    return (((r >> np.uint32(8)) + np.uint32(1)).astype(np.float32)) * np.float32(1.0 / 16777216.0)


def philox_uniform_image(n_envs, npix, step, stream, seed, env0=0):
    """(n_envs, npix) uniforms as cam_image_obs_kernel draws them: one block per 4 consecutive pixels."""
    q = npix // 4
    idx = (np.arange(n_envs, dtype=np.uint64)[:, None] + np.uint64(env0)) * np.uint64(q) + np.arange(q, dtype=np.uint64)
    return u01(philox(idx.reshape(-1), step, stream, seed)).reshape(n_envs, npix)


def philox_pcl_draws(n_envs, n_pts, step, seed, env0=0):
    """(z (n_envs,n_pts,3) standard normals, u (n_envs,n_pts) gate uniforms) as pcl_noise_kernel draws them."""
    idx = ((np.arange(n_envs, dtype=np.uint64)[:, None] + np.uint64(env0)) * np.uint64(n_pts)
           + np.arange(n_pts, dtype=np.uint64)).reshape(-1)
    r0, r1 = philox(idx, step, 2, seed), philox(idx, step, 3, seed)

    def bm(a, b):
        r = np.sqrt(np.float32(-2.0) * np.log(u01_open(a)), dtype=np.float32)
        ang = np.float32(2.0) * u01(b)
        return r * np.cos(np.pi * ang.astype(np.float64)).astype(np.float32), \
            r * np.sin(np.pi * ang.astype(np.float64)).astype(np.float32)
    z0, z1 = bm(r0[:, 0], r0[:, 1])
    z2, _ = bm(r0[:, 2], r0[:, 3])
    z = np.stack([z0, z1, z2], -1).reshape(n_envs, n_pts, 3)
    return z, u01(r1[:, 0]).reshape(n_envs, n_pts)


# ---- DepthImageProcessor (factory_utils.py:12-72) -----------------------------------------------
def process_depth_image(depth_images, u, dis_noise, far_clip, near_clip):
    """:55-72 with the uniform draw `u` (= torch.rand(depth_images.shape)) passed in."""
    depth_images = depth_images.clone()
    noise = dis_noise * 2 * (u - 0.5)
    depth_images += noise
    depth_images = torch.clip(depth_images, -far_clip, -near_clip)
    depth_images = depth_images * -1
    depth_images = (depth_images - near_clip) / (far_clip - near_clip)
    return depth_images.squeeze(0) if depth_images.size(0) == 1 else depth_images


def add_seg_noise(seg_images_to_noise, u, flip_prob=0.1):
    """:23-37 with the uniform draw `u` (= torch.rand_like(seg, dtype=float)) passed in; in place."""
    object_mask = seg_images_to_noise > 0
    flip_mask = u < flip_prob
    seg_images_to_noise[object_mask & flip_mask] = 0
    return seg_images_to_noise


def cam_image_obs(depth, seg, update, update_seg, seg_noise, image_buf, seg_buf, u_depth, u_seg, dis_noise,
                  far_clip, near_clip, flip_prob=0.1):
    """factory_task_insertion.py:902-943 (cam_type 'd', depth_cam + seg_cam); the draws are full-size
    (N, npix) arrays indexed by env, where the reference draws compacted arrays for the updated rows."""
    seg_noise = torch.logical_and(seg_noise, update_seg)
    depth = depth.flatten(start_dim=1)
    seg = seg.flatten(start_dim=1)
    if update.any():
        out = process_depth_image(depth[update], u_depth[update], dis_noise, far_clip, near_clip)
        image_buf[update] = out.reshape(int(update.sum()), -1)
    if update_seg.any():
        seg_buf[update_seg] = seg[update_seg]
    if seg_noise.any():
        seg_buf[seg_noise] = add_seg_noise(seg_buf[seg_noise], u_seg[seg_noise], flip_prob)
    return image_buf, seg_buf


# ---- PointCloudAugmentations.random_noise (factory_utils.py:93-100) -----------------------------
def random_noise(pointcloud_batch, pcl_noise, z, u, sigma=0.001, noise_clip=0.001, const_noise=0.001, noise_prob=0.3):
    """z = torch.randn_like(points), u = torch.rand(B, N)."""
    pointcloud_batch = pointcloud_batch.clone()
    pointwise_noise = torch.clamp(z * sigma, -noise_clip, noise_clip)
    noise_mask = (u < noise_prob).unsqueeze(-1).float()
    pointcloud_batch += pointwise_noise * noise_mask
    constant_noise = torch.clamp(pcl_noise * const_noise, -noise_clip, noise_clip)
    return pointcloud_batch + constant_noise


# ---- RunningMeanStd (running_mean_std.py:22-93), per_channel=False --------------------------------
class RunningMeanStd:
    def __init__(self, insize, epsilon=1e-05, norm_only=False):
        self.epsilon, self.norm_only, self.training = epsilon, norm_only, True
        self.running_mean = torch.zeros(insize, dtype=torch.float64)
        self.running_var = torch.ones(insize, dtype=torch.float64)
        self.count = torch.ones((), dtype=torch.float64)

    def forward(self, input, unnorm=False):
        if self.training:
            mean = input.mean([0])
            var = input.var([0])
            batch_count = input.size()[0]
            delta = mean - self.running_mean
            tot_count = self.count + batch_count
            new_mean = self.running_mean + delta * batch_count / tot_count
            m_a = self.running_var * self.count
            m_b = var * batch_count
            M2 = m_a + m_b + delta ** 2 * self.count * batch_count / tot_count
            self.running_mean, self.running_var, self.count = new_mean, M2 / tot_count, tot_count
        current_mean, current_var = self.running_mean, self.running_var
        if unnorm:
            y = torch.clamp(input, min=-5.0, max=5.0)
            y = torch.sqrt(current_var.float() + self.epsilon) * y + current_mean.float()
        elif self.norm_only:
            y = input / torch.sqrt(current_var.float() + self.epsilon)
        else:
            y = (input - current_mean.float()) / torch.sqrt(current_var.float() + self.epsilon)
            y = torch.clamp(y, min=-5.0, max=5.0)
        return y

    __call__ = forward


# ---- ExtrinsicAdapt.process_obs (ext_adapt.py:383-435) ------------------------------------------
def process_obs(obs, pcl_mean_std=None, stud_obs_mean_std=None, obj_id=2, socket_id=3, distinct=True):
    student_obs, tactile = obs.get("student_obs"), obs.get("tactile")
    img, seg, pcl = obs.get("img"), obs.get("seg"), obs.get("pcl")
    if seg is not None:
        valid_mask = ((seg == obj_id) | (seg == socket_id)).float()
        seg = seg * valid_mask if distinct else valid_mask
        if img is not None:
            img = img * valid_mask
    if pcl is not None and pcl_mean_std is not None:
        pcl = pcl_mean_std(pcl.reshape(-1, 3)).reshape((obs["pcl"].shape[0], -1, 3))
    if student_obs is not None and stud_obs_mean_std is not None:
        student_obs = stud_obs_mean_std(student_obs)
    return {"student_obs": student_obs, "tactile": tactile, "img": img, "seg": seg, "pcl": pcl}


def queue_push(queue, x):
    """factory_task_insertion.py:1046-1056."""
    queue[:, 1:] = queue[:, :-1].clone().detach()
    queue[:, 0, ...] = x
    return queue
