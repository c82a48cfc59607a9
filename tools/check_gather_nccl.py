#!/usr/bin/env python
"""NCCL check: every rank renders its env slice, one all-gather of the packed rows;
rank 0 recomputes ALL envs alone and compares bit-for-bit.  Launch with torchrun --nproc-per-node G."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import torch.distributed as dist
import bench
from isaacgyminsertion_b200 import dist as igdist
from isaacgyminsertion_b200.task_obs import FactoryTaskInsertionTactileObs

E = int(sys.argv[1]) if len(sys.argv) > 1 else 64
rank, local_rank, world = igdist.init_from_env()
torch.cuda.set_device(local_rank)
dev = torch.device("cuda", local_rank)


def run(n, offset, total):
    gym, P, depth, seg = bench.make_inputs(n, offset, total)
    task = FactoryTaskInsertionTactileObs(n, gym, P["mesh_id"], P["bg_id"], device=dev, sampler="fps", strict_rng=False)
    t = lambda a: torch.from_numpy(a).to(dev)
    fp, fq = t(P["finger_pos"]), t(P["finger_quat"])
    task.left_finger_pos, task.right_finger_pos, task.middle_finger_pos = fp[:, 0], fp[:, 1], fp[:, 2]
    task.left_finger_quat, task.right_finger_quat, task.middle_finger_quat = fq[:, 0], fq[:, 1], fq[:, 2]
    task.plug_pos, task.plug_quat = t(P["plug_pos"]), t(P["plug_quat"])
    task.cam_renders, task.seg_renders = t(depth), t(seg)
    ones = torch.ones(n, dtype=torch.bool, device=dev)
    zeros = torch.zeros(n, dtype=torch.bool, device=dev)
    task.update_tactile(ones, ones)
    task.update_external_cam(ones, ones, ones, zeros, zeros)
    return task.obs_packed


total = E * world
mine = run(E, rank * E, total)
got = igdist.gather_observations(mine, total_envs=total)
torch.cuda.synchronize()
if rank == 0:
    want = run(total, 0, total)
    same = bool(torch.equal(got, want))
    print(f"nccl gather world={world} total_envs={total}: gathered == single-GPU bit-for-bit: {same}; "
          f"nonzero rows {(got.abs().sum(1) > 0).sum().item()}")
    assert same
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
