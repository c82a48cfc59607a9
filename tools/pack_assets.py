#!/usr/bin/env python
"""Pack the reference's DATA assets the hot path needs into one small npz.

Run in the build container (needs /root/reference, which does not exist on the
GPU box):  python tools/pack_assets.py

Nothing here is reference source code: it reads meshes / JPEGs / npz pose
tables and stores derived arrays.

What is packed (and which reference line consumes the original):
  gel_tris   (G,3,3) f32  allsight_fine01.obj triangle soup, x,y * 1.02
                          (tacto/renderer.py:195-204), minus triangles that lie
                          entirely behind the in-gel camera's near plane
                          (x <= 0.011 m, cannot produce a fragment).
  peg_<i>_v / _f / _vn    the 7 default plug meshes
                          (FactoryEnvInsertionTactile.yaml:47-55,
                          factory_env_insertion.py:1037-1053): vertices merged on
                          position AND file normal (trimesh.load -> merge_vertices
                          defaults: merge_norm=False, 8 / 2 digits), x,y * asset scale
                          (allsight_render.py:101-107), angle-weighted vertex
                          normals recomputed after scaling (SURVEY T2 decision).
  bg_real    (8,224,224,3) u8  ref_frame_white{12..19}.jpg, cv2.resize to
                          224x224 then BGR->RGB (tacto/renderer.py:555-558).
  grasp_*    per subassembly, first 512 rows of <sub>_noise.npz
                          (factory_env_insertion.py:541-607): plug/socket poses.
"""
import os
import sys

import cv2
import numpy as np
import yaml

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(__file__), "..", "isaacgyminsertion_b200", "data", "assets.npz")

SUBASSEMBLIES = [
    "hexagon", "ellipse", "trapez", "small_triangle",
    "red_round_peg_1_5in", "yellow_round_peg_2in", "square_peg_hole_32mm_loose",
]


sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
from isaacgyminsertion_b200.assets import load_obj, merge_vertices, angle_weighted_normals  # noqa: E402


def main():
    out = {}
    info = yaml.safe_load(open(f"{REF}/assets/factory/yaml/factory_asset_info_insertion.yaml"))

    # --- gel ---
    V, F = load_obj(f"{REF}/assets/urdf/kuka_openhand_description/meshes/allsight/allsight_fine01.obj")
    V[:, 0] *= 1.02
    V[:, 1] *= 1.02
    tri = V[F]
    keep = tri[:, :, 0].max(axis=1) > 0.011  # camera x=0.01, znear 0.001
    out["gel_tris"] = tri[keep].astype(np.float32)
    out["gel_tri_count_full"] = np.int64(len(F))
    print("gel:", len(F), "tris ->", int(keep.sum()), "kept")

    # --- pegs ---
    names, scales = [], []
    for i, sub in enumerate(SUBASSEMBLIES):
        comps = list(info[sub].keys())
        plug = info[sub][comps[0]]
        f = plug["urdf_path"]
        f += "_subdiv_3x.obj" if ("rectangular" in f or "square" in f) else ".obj"
        V, F, VNfile = load_obj(f"{REF}/assets/factory/mesh/factory_insertion/{f}", with_normals=True)
        V, F = merge_vertices(V, F, VNfile)
        s = float(plug["scale"])
        V[:, 0] *= s
        V[:, 1] *= s
        VN = angle_weighted_normals(V, F)
        out[f"peg_{i}_v"] = V.astype(np.float32)
        out[f"peg_{i}_f"] = F.astype(np.int32)
        out[f"peg_{i}_vn"] = VN.astype(np.float32)
        names.append(sub)
        scales.append(s)
        print(f"peg {i} {sub}: {f} scale {s} verts {len(V)} faces {len(F)} "
              f"bbox {V.min(0).round(4)} {V.max(0).round(4)}")
        g = np.load(f"{REF}/isaacgyminsertion/initial_grasp_data/{sub}_noise.npz")
        for k in ("plug_pos", "plug_quat", "socket_pos", "socket_quat"):
            out[f"grasp_{i}_{k}"] = g[k][:512].astype(np.float32)
    out["peg_names"] = np.array(names)
    out["peg_scales"] = np.array(scales, dtype=np.float64)

    # --- backgrounds ---
    bgs = []
    for bg_id in range(12, 20):
        img = cv2.imread(f"{REF}/isaacgyminsertion/allsight/experiments/conf/ref/ref_frame_white{bg_id}.jpg")
        assert img is not None
        img = cv2.resize(img, (224, 224))[:, :, ::-1]
        bgs.append(np.ascontiguousarray(img))
    out["bg_real"] = np.stack(bgs).astype(np.uint8)

    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    np.savez_compressed(OUT, **out)
    print("wrote", os.path.abspath(OUT), os.path.getsize(OUT) / 1e6, "MB")


if __name__ == "__main__":
    sys.exit(main())
