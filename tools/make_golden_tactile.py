#!/usr/bin/env python
"""Golden vectors for the tactile path, produced by the CPU oracle (oracle/tactile.py + raster.c).

The reference ships no golden image and pyrender cannot be installed here, so these
pin the ORACLE: its cv2 / scipy stages are the real libraries the reference calls, its raster stage
is this repo's statement of pyrender's behaviour (parity unpinned).  The fixtures detect drift of
either (a cv2 upgrade changing INTER_AREA, an edit of raster.c) and give the CUDA path a fixed
target that does not need the oracle at test time.

Two files, one per light model (DESIGN.md "light model"): tactile_golden.npz is the shipped default
(`falloff: inverse_square`: every fragment saturates, the colour image equals the background and only
gel_depth carries the contact), tactile_golden_none.npz the documented deviation (`falloff: none`), whose
unsaturated images exercise the shading + calibration arithmetic.

    python tools/make_golden_tactile.py     ->  tests/golden/tactile_golden{,_none}.npz
"""
import os
import sys

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)

from isaacgyminsertion_b200 import synthetic  # noqa: E402
from oracle import tactile as ot  # noqa: E402

N_ENVS = 14   # two envs per peg: 42 frames with the 50/25/25 contact mix


def main():
    for falloff, name in (("inverse_square", "tactile_golden.npz"), ("none", "tactile_golden_none.npz")):
        make(falloff, name)


def make(falloff, name):
    model = ot.SensorModel(falloff=falloff)
    P = synthetic.tactile_poses(N_ENVS, model.assets, seed=0)
    obj_tf = ot.xyzquat_to_tf_numpy(np.concatenate([P["plug_pos"], P["plug_quat"]], 1))
    M, gd, delta, obs = [], [], [], []
    for e in range(N_ENVS):
        for n in range(3):
            h = ot.OracleAllSight(model, int(P["mesh_id"][e]), int(P["bg_id"][e, n]))
            ftf = ot.xyzquat_to_tf_numpy(np.concatenate([P["finger_pos"][e, n], P["finger_quat"][e, n]]))[0]
            h.update_pose_given_sim_pose(ftf, obj_tf[e])
            color, gel_depth, raw, kind, m = h.render(obj_tf[e], 70, return_raw=True)
            M.append(m)
            gd.append(gel_depth)
            delta.append(color.astype(np.int16) - h.bg_img.astype(np.int16))   # sparse: compresses well
            obs.append(ot.tactile_obs(color, h.bg_img, h.mask))
    out = os.path.join(ROOT, "tests", "golden", name)
    np.savez_compressed(
        out, n_envs=N_ENVS, seed=0, force=70.0, falloff=falloff,
        finger_pos=P["finger_pos"], finger_quat=P["finger_quat"], plug_pos=P["plug_pos"], plug_quat=P["plug_quat"],
        mesh_id=P["mesh_id"], bg_id=P["bg_id"], M=np.stack(M).astype(np.float32), gel_depth=np.stack(gd),
        color_delta=np.stack(delta), obs=np.stack(obs).astype(np.float32),
        depth0=model.depth0, bg_sim=model.bg_sim)
    print(out, os.path.getsize(out) / 1e6, "MB;", sum(int((g != 0).any()) for g in gd), "frames with contact")


if __name__ == "__main__":
    main()
