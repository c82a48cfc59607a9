#!/usr/bin/env python
"""Generate tests/golden/student_golden.npz by running the REAL reference code.

Build-container only (needs /root/reference).
  * algo/models/running_mean_std.py is pure torch and is imported as is.
  * tasks/factory_tactile/factory_utils.py imports isaacgym / pytorch3d / torch_jit_utils at the top
    (unused by DepthImageProcessor and PointCloudAugmentations); they are stubbed.
  * ExtrinsicAdapt.process_obs (algo/ext_adapt/ext_adapt.py:383-435) is exec'd from its source text with
    a namespace object as `self` (the module itself needs isaacgym, hydra, wandb ...).
  * The reference draws its noise from torch's global generators.  torch.rand / rand_like / randn_like
    are patched for the duration of each call to return the draws of the oracle's Philox restatement,
    so the golden outputs pin the reference's ARITHMETIC given those draws.
"""
import ast
import importlib.util
import os
import sys
import types
from unittest import mock

import numpy as np
import torch

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
REF = "/root/reference"

from oracle import student as ost  # noqa: E402
from isaacgyminsertion_b200 import synthetic  # noqa: E402


def load(path, name, stubs=()):
    for s in stubs:
        sys.modules.setdefault(s, mock.MagicMock())
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


ref_rms = load(f"{REF}/algo/models/running_mean_std.py", "ref_rms")
ref_fu = load(f"{REF}/isaacgyminsertion/tasks/factory_tactile/factory_utils.py", "ref_factory_utils",
              stubs=["isaacgym", "isaacgym.torch_utils", "isaacgyminsertion", "isaacgyminsertion.utils",
                     "isaacgyminsertion.utils.torch_jit_utils", "pytorch3d", "pytorch3d.transforms"])

src = open(f"{REF}/algo/ext_adapt/ext_adapt.py").read()
cls = [n for n in ast.parse(src).body if isinstance(n, ast.ClassDef) and n.name == "ExtrinsicAdapt"][0]
fn = [n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name == "process_obs"][0]
ns = {"torch": torch}
exec(compile(ast.Module(body=[fn], type_ignores=[]), "ref_process_obs", "exec"), ns)
ref_process_obs = ns["process_obs"]


class Draws:
    """Patches torch.rand / rand_like / randn_like to hand out queued tensors."""

    def __init__(self, **q):
        self.q = q

    def __enter__(self):
        self.p = [mock.patch.object(torch, "rand", lambda *a, **k: self.q.pop("rand")),
                  mock.patch.object(torch, "rand_like", lambda *a, **k: self.q.pop("rand_like")),
                  mock.patch.object(torch, "randn_like", lambda *a, **k: self.q.pop("randn_like"))]
        for p in self.p:
            p.start()
        return self

    def __exit__(self, *a):
        for p in self.p:
            p.stop()
        assert not self.q, f"unused draws: {list(self.q)}"


def main():
    N, seed = 12, 7
    gym = synthetic.SyntheticGym(N, seed=5)
    pp, pq, sp = synthetic.scene_poses(N, seed=5)
    depth, seg = synthetic.external_camera_frames(gym, pp, pq, sp, seed=5)
    H, W = depth.shape[1:]
    npix = H * W
    depth[3, :4, :4] = np.nan          # NaN propagates through clip
    depth_t, seg_t = torch.from_numpy(depth), torch.from_numpy(seg)
    rng = np.random.default_rng(0)
    update = torch.from_numpy(rng.random(N) < 0.7)
    update_seg = torch.from_numpy(rng.random(N) < 0.7)
    seg_noise = torch.from_numpy(rng.random(N) < 0.6)
    update[0], update_seg[0], seg_noise[0] = True, True, True
    update[1], update_seg[1], seg_noise[1] = False, False, True     # noise without update: ignored
    dis_noise, far_clip, near_clip, flip_prob, step = 0.001, 0.5, 0.1, 0.1, 3

    # ---- image observations: the reference's own lines (factory_task_insertion.py:902-943) on its classes
    proc = ref_fu.DepthImageProcessor(cfg=None, dis_noise=dis_noise, far_clip=far_clip, near_clip=near_clip)
    u_depth = torch.from_numpy(ost.philox_uniform_image(N, npix, step, 0, seed))
    u_seg = torch.from_numpy(ost.philox_uniform_image(N, npix, step, 1, seed))
    image_buf = torch.full((N, npix), -7.0)
    seg_buf = torch.full((N, npix), -7, dtype=torch.int32)
    sn = torch.logical_and(seg_noise, update_seg)
    with Draws(rand=u_depth[update].reshape(-1, H, W)):
        image_buf[update] = proc.process_depth_image(depth_t[update]).flatten(start_dim=1)
    seg_buf[update_seg] = seg_t[update_seg].flatten(start_dim=1)
    with Draws(rand_like=u_seg[sn]):
        seg_buf[sn] = proc.add_seg_noise(seg_buf[sn])

    # ---- PointCloudAugmentations.random_noise
    B, P = 6, 400
    pts = torch.from_numpy(rng.uniform(-0.3, 0.7, (B, P, 3)).astype(np.float32))
    pcl_pos_noise = torch.from_numpy(rng.standard_normal((B, 1, 3)).astype(np.float32) * 1.5)
    z, u = ost.philox_pcl_draws(B, P, step, seed)
    aug = ref_fu.PointCloudAugmentations(num_points=P)
    with Draws(randn_like=torch.from_numpy(z), rand=torch.from_numpy(u)):
        noisy = aug.random_noise(pts.clone(), pcl_pos_noise)

    # ---- RunningMeanStd: three training batches then one eval batch, on pcl-shaped rows
    torch.manual_seed(0)
    m = ref_rms.RunningMeanStd(3)
    m.train()
    batches = [torch.from_numpy(rng.normal([0.45, 0.02, 0.08], [0.05, 0.03, 0.02], (4000, 3)).astype(np.float32))
               for _ in range(4)]
    rms_out = [m(b).numpy() for b in batches[:3]]
    rms_state = [m.running_mean.numpy().copy(), m.running_var.numpy().copy(), m.count.numpy().copy()]
    m.eval()
    rms_eval = m(batches[3]).numpy()
    rms_unnorm = m(torch.from_numpy(rms_eval), unnorm=True).numpy()

    # ---- process_obs with seg, img, pcl
    T, NP = 2, 3
    segq = torch.from_numpy(rng.integers(0, 5, (NP, T, npix)).astype(np.float32))
    imgq = torch.from_numpy(rng.random((NP, T, npix)).astype(np.float32))
    pclq = torch.from_numpy(rng.normal(0.3, 0.1, (NP, T, 2400)).astype(np.float32))
    stud = torch.from_numpy(rng.normal(0, 1, (NP, 18)).astype(np.float32))
    self_ns = types.SimpleNamespace(obs_info=True, tactile_info=False, img_info=True, seg_info=True, pcl_info=True,
                                    display_obs=False, stats=None, train_config=types.SimpleNamespace(from_offline=False),
                                    pcl_mean_std=ref_rms.RunningMeanStd(3), stud_obs_mean_std=ref_rms.RunningMeanStd(18))
    po = ref_process_obs(self_ns, {"student_obs": stud, "img": imgq, "seg": segq, "pcl": pclq})
    po_nd = ref_process_obs(types.SimpleNamespace(**{**vars(self_ns), "obs_info": False, "pcl_info": False}),
                            {"img": imgq, "seg": segq}, distinct=False)

    # ---- Philox known answers (Random123 kat_vectors: philox4x32-10)
    out = dict(
        seed=np.int64(seed), step=np.int64(step), dis_noise=dis_noise, far_clip=far_clip, near_clip=near_clip,
        flip_prob=flip_prob, depth=depth, seg=seg, update=update.numpy(), update_seg=update_seg.numpy(),
        seg_noise=seg_noise.numpy(), image_buf=image_buf.numpy(), seg_buf=seg_buf.numpy(),
        pts=pts.numpy(), pcl_pos_noise=pcl_pos_noise.numpy(), noisy=noisy.numpy(),
        rms_batches=np.stack([b.numpy() for b in batches]), rms_out=np.stack(rms_out), rms_mean=rms_state[0],
        rms_var=rms_state[1], rms_count=rms_state[2], rms_eval=rms_eval, rms_unnorm=rms_unnorm,
        po_seg_in=segq.numpy(), po_img_in=imgq.numpy(), po_pcl_in=pclq.numpy(), po_stud_in=stud.numpy(),
        po_seg=po["seg"].numpy(), po_img=po["img"].numpy(), po_pcl=po["pcl"].numpy(), po_stud=po["student_obs"].numpy(),
        po_nd_seg=po_nd["seg"].numpy(), po_nd_img=po_nd["img"].numpy(),
    )
    path = os.path.join(ROOT, "tests", "golden", "student_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
