#!/usr/bin/env python
"""Per-phase cycle shares of tac_contact (library built with -DCT_PROFILE)."""
import ctypes, sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch
import bench
from isaacgyminsertion_b200 import _lib
from isaacgyminsertion_b200.task_obs import FactoryTaskInsertionTactileObs
E = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
gym, P, depth, seg = bench.make_inputs(E, 0, E)
task = FactoryTaskInsertionTactileObs(E, gym, P["mesh_id"], P["bg_id"], device="cuda:0", sampler="fps", strict_rng=False, pcl_cam=False)
dev = task.device
fp = torch.from_numpy(P["finger_pos"]).to(dev); fq = torch.from_numpy(P["finger_quat"]).to(dev)
pp = torch.from_numpy(P["plug_pos"]).to(dev); pq = torch.from_numpy(P["plug_quat"]).to(dev)
lib = _lib.load()
out = (ctypes.c_ulonglong * 16)()
for i in range(3):
    task.tactile_engine.render(fp, fq, pp, pq, obs_out=task.tactile_imgs)
lib.igi_debug_read_prof(out, 1)
task.tactile_engine.render(fp, fq, pp, pq, obs_out=task.tactile_imgs)
lib.igi_debug_read_prof(out, 1)
names = ["fetch/idle", "zinit", "rows+scan", "raster", "shade", "hblur", "vblur+store", "(loop end)", "obs"]
v = np.array(list(out)[:9], dtype=np.float64)
for n, x in zip(names, v):
    print(f"{n:12s} {100*x/v.sum():5.1f}%   {x/1e6:9.1f} Mcycles")
