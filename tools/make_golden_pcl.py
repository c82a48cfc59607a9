#!/usr/bin/env python
"""Generate tests/golden/pcl_golden.npz by running the REAL reference classes.

Build-container only (needs /root/reference).  The reference module
isaacgyminsertion/tasks/utils/pcl_utils.py imports isaacgym and matplotlib at the
top (only used for gym handles / plotting); both are stubbed so that the module's own
PointCloudGenerator / CameraPointCloud code runs unmodified on CPU torch.  `filter_pts`
is exec'd from the source text of factory_task_insertion.py:65-77 (that module cannot be
imported: hydra, isaacgym, TkAgg).
"""
import ast
import importlib.util
import os
import sys
from unittest import mock

import numpy as np
import torch

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
REF = "/root/reference"

for name in ["isaacgym", "isaacgym.gymapi", "isaacgym.gymtorch", "matplotlib", "matplotlib.pyplot",
             "mpl_toolkits", "mpl_toolkits.mplot3d"]:
    sys.modules[name] = mock.MagicMock()

spec = importlib.util.spec_from_file_location(
    "ref_pcl_utils", f"{REF}/isaacgyminsertion/tasks/utils/pcl_utils.py")
ref = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref)

src = open(f"{REF}/isaacgyminsertion/tasks/factory_tactile/factory_task_insertion.py").read()
fn = [n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == "filter_pts"][0]
ns = {"torch": torch}
exec(compile(ast.Module(body=[fn], type_ignores=[]), "ref_filter_pts", "exec"), ns)
ref_filter_pts = ns["filter_pts"]

from isaacgyminsertion_b200 import synthetic  # noqa: E402


def main():
    N, m = 24, 400
    gym = synthetic.SyntheticGym(N, seed=3)
    plug_pos, plug_quat, socket_pos = synthetic.scene_poses(N, seed=3)
    depth, seg = synthetic.external_camera_frames(gym, plug_pos, plug_quat, socket_pos, seed=3)
    # edge cases: env 5 has no plug pixels, env 6 is all ray misses, env 7 has no socket pixels,
    # env 8: every pixel is plug (max-size cloud before the box filter)
    seg[5][seg[5] == 2] = 1
    depth[6][:] = -np.inf
    seg[7][seg[7] == 3] = 0
    seg[8][:] = 2
    depth_t = torch.from_numpy(depth)
    seg_t = torch.from_numpy(seg)

    gen = ref.CameraPointCloud(isc_sim=None, isc_gym=gym, envs=gym.envs, camera_handles=gym.camera_handles,
                               camera_props=gym.camera_props, sample_num=m, pt_in_local=True,
                               graphics_device="cpu", compute_device="cpu")
    torch.manual_seed(42)
    plug_depth = depth_t.flatten(start_dim=1) * (seg_t.flatten(start_dim=1) == 2)
    plug = gen.get_point_cloud(depths=plug_depth.reshape(N, gym.height, gym.width),
                               filter_func=ref_filter_pts, sample_num=m)
    socket_depth = depth_t.flatten(start_dim=1) * (seg_t.flatten(start_dim=1) == 3)
    socket = gen.get_point_cloud(depths=socket_depth.reshape(N, gym.height, gym.width),
                                 filter_func=ref_filter_pts, sample_num=m)
    probe = torch.randint(0, 1 << 20, (8,))  # generator position after the two calls

    plug_all = gen.get_ptd_cuda(plug_depth.reshape(N, gym.height, gym.width), filter_func=ref_filter_pts)
    socket_all = gen.get_ptd_cuda(socket_depth.reshape(N, gym.height, gym.width), filter_func=ref_filter_pts)
    unfiltered0 = gen.pt_generators[0].convert(depth_t[0])

    out = dict(
        seed=np.int64(3), depth=depth, seg=seg,
        plug=plug.numpy(), socket=socket.numpy(), rng_probe=probe.numpy(),
        plug_counts=np.array([len(p) for p in plug_all]), socket_counts=np.array([len(p) for p in socket_all]),
        plug_all=torch.cat(plug_all).numpy(), socket_all=torch.cat(socket_all).numpy(),
        unfiltered0=unfiltered0.numpy(),
        uv_table0=gen.pt_generators[0]._uv_one_in_cam.numpy(),
        ext0=gen.pt_generators[0].ext_mat.numpy(),
    )
    path = os.path.join(ROOT, "tests", "golden", "pcl_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) / 1e3, "kB")
    print("plug counts", out["plug_counts"], "socket counts", out["socket_counts"])


if __name__ == "__main__":
    main()
