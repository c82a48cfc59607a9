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
task = FactoryTaskInsertionTactileObs(E, gym, P["mesh_id"], P["bg_id"], device="cuda:0", sampler="fps", strict_rng=False, pcl_cam=False,
                                      falloff=os.environ.get("IGI_FALLOFF") or None)
print("falloff", task.tactile_engine.cfg.falloff)
dev = task.device
fp = torch.from_numpy(P["finger_pos"]).to(dev); fq = torch.from_numpy(P["finger_quat"]).to(dev)
pp = torch.from_numpy(P["plug_pos"]).to(dev); pq = torch.from_numpy(P["plug_quat"]).to(dev)
lib = _lib.load()
out = (ctypes.c_ulonglong * 32)()
for i in range(3):
    task.tactile_engine.render(fp, fq, pp, pq, obs_out=task.tactile_imgs)
lib.igi_debug_read_prof(out, 1)
task.tactile_engine.render(fp, fq, pp, pq, obs_out=task.tactile_imgs)
lib.igi_debug_read_prof(out, 1)
names = ["fetch/idle", "zinit", "rows+scan", "raster", "shade", "hblur", "vblur+store", "(loop end)", "obs"]
v = np.array(list(out)[:9], dtype=np.float64)
for n, x in zip(names, v):
    print(f"{n:12s} {100*x/v.sum():5.1f}%   {x/1e6:9.1f} Mcycles")
c = list(out)
nf = int((task.tactile_engine.contact_counts() > 0).sum().item())
print(f"frames {nf}: per frame rows {c[9]/nf:.0f}, span pixels {c[10]/nf:.0f}, shaded px {c[11]/nf:.0f}, "
      f"regions x chunks {c[12]/nf:.2f}, region px {c[13]/nf:.0f}, tris {c[14]/nf:.0f}")
print(f"frames with hits {c[15]} of {nf}; per hit frame: changed box {c[23]/max(c[15],1):.0f} px, dirty window {c[24]/max(c[15],1):.0f} px")
F = 3 * E
print(f"geom per frame (all {F}): clusters kept {c[16]/F:.1f} of {c[17]/F:.1f}, faces in {c[18]/F:.0f}, front-facing {c[19]/F:.0f}, "
      f"near gel {c[20]/F:.0f}, on screen {c[21]/F:.0f}, pass hi-z {c[22]/F:.0f} (incl. warp-queue recomputation), emitted {float(task.tactile_engine.contact_counts().clamp(min=0).float().mean()):.0f}")
