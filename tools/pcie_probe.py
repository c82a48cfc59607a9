#!/usr/bin/env python
"""Pinned host<->device copy rates at the e2e leg's sizes (170 MB up, 140 MB down), alone and together.
Under torchrun every rank probes ITS GPU at the same time (rendezvous barrier before each leg), bound to the
GPU's NUMA node unless --no-numa: the sum over ranks is the host-side ceiling of the multi-GPU e2e leg."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import torch.distributed as dist
from isaacgyminsertion_b200 import dist as igdist
rank, local_rank, world = igdist.init_from_env()
torch.cuda.set_device(local_rank)
numa = None if "--no-numa" in sys.argv else igdist.bind_to_gpu_numa(local_rank)
up_b, dn_b = 170328064, 139984896
h_up = torch.empty(up_b, dtype=torch.uint8).pin_memory(); d_up = torch.empty(up_b, dtype=torch.uint8, device="cuda")
h_dn = torch.empty(dn_b, dtype=torch.uint8).pin_memory(); d_dn = torch.empty(dn_b, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(up, dn, n=10):
    if world > 1:
        dist.barrier()
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
msg = (f"rank {rank}/{world} H2D alone {a:.3f} ms ({up_b/a/1e6:.1f} GB/s)  D2H alone {b:.3f} ms ({dn_b/b/1e6:.1f} GB/s)  "
       f"both {c:.3f} ms per step-pair ({(up_b+dn_b)/c/1e6:.1f} GB/s)  numa {numa}")
if world > 1:
    out = [None] * world
    dist.all_gather_object(out, (msg, c))
    if rank == 0:
        for m, _ in out:
            print(m)
        worst = max(x for _, x in out)
        print(f"H2D+D2H all {world} ranks at once: slowest rank {worst:.3f} ms per step-pair, aggregate {(up_b+dn_b)*world/worst/1e6:.1f} GB/s")
    dist.barrier(); dist.destroy_process_group()
else:
    print(msg)
