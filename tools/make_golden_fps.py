#!/usr/bin/env python
"""Generate tests/golden/fps_golden.npz: FPS index vectors for the sizes SURVEY.md 8c lists (N = 1, 37, 400, 5184,
duplicate-point / identical-point / never-candidate / lattice cases included).

The indices are produced by the LITERAL thread-by-thread emulation of the published pointnet2_ops kernel
(`_fps_literal` in tests/test_oracle_pcl.py: B threads striding over the points, strict `>`, shared-memory tree where the
lower position wins ties; sums of squares as nvcc contracts them, evaluated with libm's fmaf) - code that shares nothing
with oracle/fps.c's sorted-key formulation, so the committed vectors pin both the oracle and the CUDA kernels.
Takes about a minute (pure Python loops).
"""
import os
import sys

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_oracle_pcl import _fps_literal  # noqa: E402


def cases():
    rng = np.random.default_rng(2024)
    out = {}
    for n, m in ((1, 8), (37, 64), (400, 400), (5184, 48)):
        base = (rng.random((n, 3)) * 0.5 + 0.1).astype(np.float32)
        out[f"rand_{n}"] = (base, m)
        if n >= 37:
            dup = base.copy(); dup[n // 2:] = dup[: n - n // 2]
            out[f"dup_{n}"] = (dup, m)
            zer = base.copy(); zer[::3] = 0.0
            out[f"zero_{n}"] = (zer, m)
        if n == 37:
            same = base.copy(); same[:] = same[0]
            out["same_37"] = (same, m)
        if n == 400:
            g = np.stack(np.meshgrid(np.arange(8), np.arange(8), np.arange(8)), -1).reshape(-1, 3)[:n]
            out["lattice_400"] = ((g * 0.125 + 0.25).astype(np.float32), m)
    return out


def main():
    data = {}
    for name, (pts, m) in cases().items():
        idx = _fps_literal(pts, m)
        data[f"{name}_pts"] = pts
        data[f"{name}_idx"] = idx.astype(np.int32)
        print(name, pts.shape, m, idx[:6])
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "fps_golden.npz"), **data)


if __name__ == "__main__":
    main()
