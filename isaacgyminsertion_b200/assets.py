"""Mesh / image assets of the tactile path: OBJ loading the way the reference's
`trimesh.load` + in-place scaling does it (allsight_render.py:101-107), and the packed
asset file `data/assets.npz` written by tools/pack_assets.py."""
import os

import numpy as np

DATA_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")
ASSETS_NPZ = os.path.join(DATA_DIR, "assets.npz")
SENSOR_YML = os.path.join(DATA_DIR, "sensor_allsight_white.yml")


def load_obj(path, with_normals=False):
    """All-triangle OBJ -> (V f64 (nv,3), F i64 (nf,3)) [, VN f64 (nv,3) | None].

    With `with_normals`, a face corner is a (position, normal) pair like in trimesh's OBJ loader:
    corners that share a position index but name different `vn` records become different
    vertices (first-occurrence order), and VN holds the FILE normal of every vertex; VN is None
    when the file has no `vn` records or a corner omits its normal index.  `vt` is ignored."""
    vs, vns, corners = [], [], []
    with open(path) as fh:
        for line in fh:
            if line.startswith("v "):
                p = line.split()
                vs.append((float(p[1]), float(p[2]), float(p[3])))
            elif line.startswith("vn "):
                p = line.split()
                vns.append((float(p[1]), float(p[2]), float(p[3])))
            elif line.startswith("f "):
                toks = line.split()[1:]
                if len(toks) != 3:
                    raise ValueError(f"non-triangle face in {path}")
                for t in toks:
                    q = t.split("/")
                    corners.append((int(q[0]), int(q[2]) if len(q) > 2 and q[2] else 0))
    V = np.asarray(vs, dtype=np.float64).reshape(-1, 3)
    C = np.asarray(corners, dtype=np.int64).reshape(-1, 2)
    vi = np.where(C[:, 0] > 0, C[:, 0] - 1, C[:, 0] + len(V))          # negative = relative to the end
    if not with_normals:
        return V, vi.reshape(-1, 3)
    VNf = np.asarray(vns, dtype=np.float64).reshape(-1, 3)
    if len(VNf) == 0 or (C[:, 1] == 0).any():
        return V, vi.reshape(-1, 3), None
    ni = np.where(C[:, 1] > 0, C[:, 1] - 1, C[:, 1] + len(VNf))
    if len(VNf) == len(V) and np.array_equal(vi, ni):                  # Meshlab export: one normal per vertex
        return V, vi.reshape(-1, 3), VNf
    pair = vi * (len(VNf) + 1) + ni
    _, first, inverse = np.unique(pair, return_index=True, return_inverse=True)
    order = np.argsort(first)
    rank = np.empty_like(order)
    rank[order] = np.arange(len(order))
    keep = first[order]
    return V[vi[keep]], rank[inverse.reshape(-1)].reshape(-1, 3), VNf[ni[keep]]


def merge_vertices(V, F, VN=None, digits=8, digits_norm=2):
    """`trimesh.load(...)` -> `Trimesh.process()` -> `merge_vertices()` with its defaults: referenced
    vertices are merged when their position agrees to 1e-8 AND (when the file carries vertex
    normals; `merge_norm=False`) their file normal agrees to 2 decimals, so crease edges that the
    exporter split stay split.  Unreferenced vertices are dropped.  First-occurrence order.
    Returns (V', F')."""
    F = np.asarray(F, dtype=np.int64)
    ref = np.zeros(len(V), dtype=bool)
    ref[F.reshape(-1)] = True
    idx = np.nonzero(ref)[0]
    cols = [np.round(V[idx] * 10 ** digits)]
    if VN is not None and np.shape(VN) == np.shape(V):
        cols.append(np.round(np.asarray(VN)[idx] * 10 ** digits_norm))
    key = np.column_stack(cols).astype(np.int64)
    _, first, inverse = np.unique(key, axis=0, return_index=True, return_inverse=True)
    order = np.argsort(first)
    rank = np.empty_like(order)
    rank[order] = np.arange(len(order))
    remap = np.full(len(V), -1, dtype=np.int64)
    remap[idx] = rank[inverse.reshape(-1)]
    return V[idx[first[order]]], remap[F]


def angle_weighted_normals(V, F):
    """Vertex normals as trimesh (>= 3.9, `geometry.weighted_vertex_normals`) recomputes them once the
    vertices were edited in place (allsight_render.py:105-106 drops the cached file normals): face
    normals summed with the corner angle as weight, then normalised.  Degenerate faces add nothing."""
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
    """(V f32, VN f32, F i32) of a plug OBJ the way allsight_render.py:101-109 prepares it: trimesh.load
    (merge rule above), x,y scaled in place by `scale`, smooth vertex normals recomputed AFTER the scale."""
    V, F, VN = load_obj(path, with_normals=True)
    V, F = merge_vertices(V, F, VN)
    V[:, 0] *= scale
    V[:, 1] *= scale
    return V.astype(np.float32), angle_weighted_normals(V, F).astype(np.float32), F.astype(np.int32)


def load_packed(path=ASSETS_NPZ):
    if not os.path.exists(path):
        raise FileNotFoundError(f"{path} missing: run tools/pack_assets.py in the build container")
    return np.load(path)
