#!/usr/bin/env python
"""Pinned host<->device copy rates at the e2e leg's sizes (170 MB up, 140 MB down), alone and together."""
import torch
up_b, dn_b = 170328064, 139984896
h_up = torch.empty(up_b, dtype=torch.uint8).pin_memory(); d_up = torch.empty(up_b, dtype=torch.uint8, device="cuda")
h_dn = torch.empty(dn_b, dtype=torch.uint8).pin_memory(); d_dn = torch.empty(dn_b, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(up, dn, n=10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    s1.wait_event(e0); s2.wait_event(e0)
    for _ in range(n):
        if up:
            with torch.cuda.stream(s1): d_up.copy_(h_up, non_blocking=True)
        if dn:
            with torch.cuda.stream(s2): h_dn.copy_(d_dn, non_blocking=True)
    torch.cuda.current_stream().wait_stream(s1); torch.cuda.current_stream().wait_stream(s2)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
run(True, True, 3)
a, b, c = run(True, False), run(False, True), run(True, True)
print(f"H2D alone {a:.3f} ms ({up_b/a/1e6:.1f} GB/s)  D2H alone {b:.3f} ms ({dn_b/b/1e6:.1f} GB/s)  both {c:.3f} ms per step-pair")
