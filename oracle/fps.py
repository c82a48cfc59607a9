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

Arithmetic: f32, in the form nvcc gives the published expressions under its default
-fmad=true — `x*x + y*y + z*z` becomes mul(y,y), fma(x,x,.), fma(z,z,.) (checked here by
compiling the two source lines with nvcc 12.9 and reading the PTX).  Python has no f32
fma, so the loop lives in oracle/fps.c (explicit fmaf(), built with -ffp-contract=off by
oracle/Makefile); this module is its loader plus the task-level emptiness rule.  The CUDA
kernels use __fmul_rn/__fmaf_rn in the same order, so indices compare bit-for-bit.
tests/test_oracle_pcl.py checks the C loop against a literal thread-by-thread emulation
of the published kernel.
"""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
FPS_SO = os.path.join(HERE, "liboracle_fps.so")
_lib = None


def _load():
    global _lib
    if _lib is None:
        if not os.path.exists(FPS_SO):
            subprocess.run(["make", "-C", HERE, "liboracle_fps.so"], check=True, stdout=subprocess.DEVNULL)
        _lib = ctypes.CDLL(FPS_SO)
        _lib.igi_oracle_fps.restype = ctypes.c_int
        _lib.igi_oracle_fps.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
    return _lib


def furthest_point_sample(pts, m):
    """pts (n,3) f32 -> (m,) int32 indices."""
    pts = np.ascontiguousarray(pts, dtype=np.float32).reshape(-1, 3)
    idx = np.zeros(m, dtype=np.int32)
    if pts.shape[0] == 0 or m == 0:
        return idx
    rc = _load().igi_oracle_fps(pts.ctypes.data, pts.shape[0], m, idx.ctypes.data)
    if rc != 0:
        raise MemoryError("igi_oracle_fps")
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
