"""Mesh / image assets of the tactile path: OBJ loading the way the reference's
`trimesh.load` + in-place scaling does it (allsight_render.py:101-107), and the packed
asset file `data/assets.npz` written by tools/pack_assets.py."""
import os

import numpy as np

DATA_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")
ASSETS_NPZ = os.path.join(DATA_DIR, "assets.npz")
SENSOR_YML = os.path.join(DATA_DIR, "sensor_allsight_white.yml")


def load_obj(path):
    """All-triangle OBJ -> (V f64 (nv,3), F i64 (nf,3)); vn/vt records are ignored."""
    vs, fs = [], []
    with open(path) as fh:
        for line in fh:
            if line.startswith("v "):
                p = line.split()
                vs.append((float(p[1]), float(p[2]), float(p[3])))
            elif line.startswith("f "):
                idx = [int(t.split("/")[0]) for t in line.split()[1:]]
                if len(idx) != 3:
                    raise ValueError(f"non-triangle face in {path}")
                fs.append(idx)
    V = np.asarray(vs, dtype=np.float64)
    F = np.asarray(fs, dtype=np.int64)
    F = np.where(F > 0, F - 1, F + len(V))
    return V, F


def merge_vertices(V, F, digits=8):
    """Merge vertices on rounded position (trimesh.load -> merge_vertices, tol 1e-8)."""
    key = np.round(V, digits)
    _, first, inverse = np.unique(key, axis=0, return_index=True, return_inverse=True)
    order = np.argsort(first)
    rank = np.empty_like(order)
    rank[order] = np.arange(len(order))
    return V[first[order]], rank[inverse.reshape(-1)][F]


def angle_weighted_normals(V, F):
    """Vertex normals as trimesh computes them after the vertices were edited:
    face normals averaged with the corner angle as weight."""
    tri = V[F]
    fn = np.cross(tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0])
    ln = np.linalg.norm(fn, axis=1, keepdims=True)
    fn = np.where(ln > 0, fn / np.maximum(ln, 1e-300), 0.0)
    vn = np.zeros_like(V)
    for k in range(3):
        a = tri[:, (k + 1) % 3] - tri[:, k]
        b = tri[:, (k + 2) % 3] - tri[:, k]
        c = np.einsum("ij,ij->i", a, b) / np.maximum(np.linalg.norm(a, axis=1) * np.linalg.norm(b, axis=1), 1e-300)
        np.add.at(vn, F[:, k], fn * np.arccos(np.clip(c, -1.0, 1.0))[:, None])
    ln = np.linalg.norm(vn, axis=1, keepdims=True)
    return np.where(ln > 0, vn / np.maximum(ln, 1e-300), 0.0)


def load_peg_from_obj(path, scale):
    """(V f32, VN f32, F i32) of a plug OBJ with x,y scaled by `scale`."""
    V, F = merge_vertices(*load_obj(path))
    V[:, 0] *= scale
    V[:, 1] *= scale
    return V.astype(np.float32), angle_weighted_normals(V, F).astype(np.float32), F.astype(np.int32)


def load_packed(path=ASSETS_NPZ):
    if not os.path.exists(path):
        raise FileNotFoundError(f"{path} missing: run tools/pack_assets.py in the build container")
    return np.load(path)
