"""GPU, N > 1: the sharded observation step + gather equals the single-GPU step bit-for-bit (SURVEY 8e,
BASELINE config 4).  Spawns one process per GPU (2 ranks, then every visible GPU) over NCCL; skipped on a
one-GPU box (run with `gpurun --gpus 2|4|8 -- python -m pytest tests/test_multigpu.py -m gpu`; the log of that run is
kept under profiles/).  Both samplers: FPS (deterministic) and the reference `torch.randint` stream, whose index
words are assigned by GLOBAL env order (pcl_utils.BatchedPointCloud.sample_reference), and every gather
transport (`dist.ObsGather`: NCCL all-gather, copy-engine peer copies into the learner's buffer).
"""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu

ENVS_PER_RANK = 48


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _obs(n, offset, total, dev, sampler, seed_rng):
    import bench
    from isaacgyminsertion_b200.task_obs import FactoryTaskInsertionTactileObs
    gym, P, depth, seg = bench.make_inputs(n, offset, total)
    task = FactoryTaskInsertionTactileObs(n, gym, P["mesh_id"], P["bg_id"], device=dev, sampler=sampler,
                                          falloff="none", global_env_offset=offset, total_envs=total)
    t = lambda a: torch.from_numpy(a).to(dev)
    fp, fq = t(P["finger_pos"]), t(P["finger_quat"])
    task.left_finger_pos, task.right_finger_pos, task.middle_finger_pos = fp[:, 0], fp[:, 1], fp[:, 2]
    task.left_finger_quat, task.right_finger_quat, task.middle_finger_quat = fq[:, 0], fq[:, 1], fq[:, 2]
    task.plug_pos, task.plug_quat = t(P["plug_pos"]), t(P["plug_quat"])
    task.cam_renders, task.seg_renders = t(depth), t(seg)
    ones = torch.ones(n, dtype=torch.bool, device=dev)
    zeros = torch.zeros(n, dtype=torch.bool, device=dev)
    torch.manual_seed(seed_rng)          # the reference sampler reads torch's CPU generator
    task.compute_observations(ones, ones, ones, ones, ones, zeros, zeros)
    return task


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    import torch.distributed as dist
    from isaacgyminsertion_b200 import dist as igdist
    r, lr, w = igdist.init_from_env(backend="nccl")
    dev = torch.device("cuda", lr)
    total = ENVS_PER_RANK * world
    results = {}
    for sampler in ("fps", "reference"):
        task = _obs(ENVS_PER_RANK, rank * ENVS_PER_RANK, total, dev, sampler, 123)
        for transport in ("nccl", "p2p"):
            g = igdist.ObsGather(task.obs_packed, total_envs=total, transport=transport)
            for step in range(3):        # several steps: buffer rotation and the step flags of the p2p transport
                got = g.gather(task.obs_packed)
                got = g.wait(got)
            torch.cuda.synchronize()
            if rank == 0:
                want = _obs(total, 0, total, dev, sampler, 123).obs_packed
                results[(sampler, transport)] = (bool(torch.equal(got, want)),
                                                 int((got.abs().sum(1) > 0).sum().item()))
            g.close()
            dist.barrier()
    if rank == 0:
        q.put(results)
    dist.barrier()
    dist.destroy_process_group()


def _run(world):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = q.get(timeout=600)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    return res


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs (gpurun --gpus N)")
@pytest.mark.parametrize("world", [2, 0])     # 0 = every visible GPU
def test_gathered_rows_equal_single_gpu_rows(built_lib, world):
    if world == 0 and torch.cuda.device_count() == 2:
        pytest.skip("two GPUs: covered by the world=2 case")
    world = world or torch.cuda.device_count()
    res = _run(world)
    total = ENVS_PER_RANK * world
    for key, (same, nonzero) in res.items():
        assert same, f"{key}: gathered rows differ from the single-GPU rows at world={world}"
        assert nonzero == total, f"{key}: {nonzero} of {total} rows populated"
    assert set(res) == {(s, t) for s in ("fps", "reference") for t in ("nccl", "p2p")}


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs (gpurun --gpus N)")
def test_engines_on_two_devices_in_one_process(built_lib):
    """One process, current device cuda:0, task objects on cuda:0 and cuda:1: same rows on both (function attributes
    are set per device, launches run under a device guard, no library state is shared)."""
    import numpy as np
    import bench
    from isaacgyminsertion_b200.task_obs import FactoryTaskInsertionTactileObs
    n = 12
    gym, P, depth, seg = bench.make_inputs(n, 0, n)
    torch.cuda.set_device(0)
    rows = []
    for dev in ("cuda:1", "cuda:0", "cuda:1"):
        task = FactoryTaskInsertionTactileObs(n, gym, P["mesh_id"], P["bg_id"], device=dev, sampler="fps", falloff="none")
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
        fp, fq = t(P["finger_pos"]), t(P["finger_quat"])
        task.left_finger_pos, task.right_finger_pos, task.middle_finger_pos = fp[:, 0], fp[:, 1], fp[:, 2]
        task.left_finger_quat, task.right_finger_quat, task.middle_finger_quat = fq[:, 0], fq[:, 1], fq[:, 2]
        task.plug_pos, task.plug_quat = t(P["plug_pos"]), t(P["plug_quat"])
        task.cam_renders, task.seg_renders = t(depth), t(seg)
        ones = torch.ones(n, dtype=torch.bool, device=dev)
        zeros = torch.zeros(n, dtype=torch.bool, device=dev)
        task.compute_observations(ones, ones, ones, ones, ones, zeros, zeros)
        task.tactile_engine.check_overflow()
        rows.append(task.obs_packed.cpu())
        assert torch.cuda.current_device() == 0
    assert torch.equal(rows[0], rows[1]) and torch.equal(rows[0], rows[2])
    assert float(rows[0].abs().sum(1).min()) > 0
