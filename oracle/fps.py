"""CPU oracle for farthest-point sampling — TEST INFRASTRUCTURE ONLY.

The reference's only FPS is `fps()` in algo/models/transformer/point_mae.py:14-21, which
calls `pointnet2_ops.pointnet2_utils.furthest_point_sample` — a third-party CUDA op
(erikwijmans/Pointnet2_PyTorch, pointnet2_ops_lib, unpinned: requirements list it as a
git dependency) whose source is NOT under /root/reference.  PARITY UNPINNED: this file
restates the published algorithm of that kernel (furthest_point_sampling_kernel in
sampling_gpu.cu) from memory:

  idxs[0] = 0; temp[:] = 1e10
  repeat m-1 times with `old` = last pick:
     for every k:  skip if |p_k|^2 <= 1e-3
                   d  = (x_k-x_old)^2 + (y_k-y_old)^2 + (z_k-z_old)^2
                   d2 = min(d, temp[k]); temp[k] = d2
     pick arg-max of d2.  Upstream threads stride over k with B = largest power of two
     <= min(n, 512) threads, keep the first maximum they meet (strict >), and the
     pairwise tree reduction keeps the lower thread on ties; the survivor among tied
     candidates is therefore the one with the smallest bit-reversed (k mod B), then the
     smallest k.  No candidate at all -> index 0.

Distances are f32 with every multiply/add rounded separately (no FMA contraction) — a
decision of this repo, stated here and in DESIGN.md; the CUDA kernel uses the same
non-contracted operations so indices can be compared bit-for-bit.
"""
import numpy as np


def _bitrev(v, bits):
    r = 0
    for _ in range(bits):
        r = (r << 1) | (v & 1)
        v >>= 1
    return r


def furthest_point_sample(pts, m):
    """pts (n,3) f32 -> (m,) int32 indices."""
    pts = np.ascontiguousarray(pts, dtype=np.float32)
    n = pts.shape[0]
    idx = np.zeros(m, dtype=np.int32)
    if n == 0:
        return idx
    lg = min(int(np.floor(np.log2(n))), 9)
    B = 1 << lg
    k = np.arange(n)
    tiekey = np.array([(_bitrev(int(i) & (B - 1), lg) << 16) | int(i) for i in k], dtype=np.int64)
    x, y, z = pts[:, 0], pts[:, 1], pts[:, 2]
    mag = (x * x + y * y) + z * z
    cand = mag > np.float32(1e-3)
    temp = np.full(n, np.float32(1e10), dtype=np.float32)
    old = 0
    for j in range(1, m):
        dx = x - x[old]
        dy = y - y[old]
        dz = z - z[old]
        d = (dx * dx + dy * dy) + dz * dz
        d2 = np.minimum(d, temp)
        temp = np.where(cand, d2, temp)
        if not cand.any():
            old = 0
        else:
            best = d2[cand].max()
            tied = cand & (d2 == best)
            old = int(k[tied][np.argmin(tiekey[tied])])
        idx[j] = old
    return idx


def fps_batch(pts_list, m):
    """Per-task FPS with the get_point_cloud emptiness rule (pcl_utils.py:179-183)."""
    out = np.zeros((len(pts_list), m, 3), dtype=np.float32)
    idx = np.zeros((len(pts_list), m), dtype=np.int32)
    for t, p in enumerate(pts_list):
        p = np.asarray(p, dtype=np.float32)
        if p.shape[0] and p.any():
            idx[t] = furthest_point_sample(p, m)
            out[t] = p[idx[t]]
    return out, idx
