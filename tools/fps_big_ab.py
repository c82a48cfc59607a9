#!/usr/bin/env python
"""Big-task FPS (n > 1024): thread-block-cluster kernel vs the one-CTA kernel, CUDA-event times."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from isaacgyminsertion_b200 import _lib
from isaacgyminsertion_b200.pcl_utils import furthest_point_sample
lib = _lib.load()
g = torch.Generator(device="cuda").manual_seed(0)
for B, n, m in ((8, 5184, 2048), (64, 5184, 2048), (512, 5184, 400), (4096, 2048, 400)):
    pts = torch.rand((B, n, 3), device="cuda", generator=g) * 0.5 + 0.1
    res = {}
    for mode in (1, 0):
        fl = 0 if mode else 1      # IGI_FPS_NO_CLUSTER
        for _ in range(2):
            idx = furthest_point_sample(pts, m, flags=fl)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for _ in range(3):
            idx = furthest_point_sample(pts, m, flags=fl)
        e1.record(); torch.cuda.synchronize()
        res[mode] = (e0.elapsed_time(e1) / 3, idx)
    print(f"tasks {B:5d} x {n} points, m={m}: cluster {res[1][0]:8.3f} ms   one-CTA {res[0][0]:8.3f} ms   equal {bool(torch.equal(res[1][1], res[0][1]))}")
