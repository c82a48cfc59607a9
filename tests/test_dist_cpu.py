"""Multi-process host logic of the observation gather (gloo, world_size 2, CPU)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from isaacgyminsertion_b200 import dist as igdist


def test_env_slices_partition_the_envs():
    for total in (1, 7, 8, 4096, 4099):
        for world in (1, 2, 3, 8):
            parts = [igdist.env_slice(total, r, world) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == total
            assert all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in parts]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, total, row, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    r, lr, w = igdist.init_from_env(backend="gloo")
    assert (r, w) == (rank, world)
    lo, hi = igdist.env_slice(total, rank, world)
    # packed rows carry the global env id so order and ownership can be checked after the gather
    local = torch.arange(lo, hi, dtype=torch.float32)[:, None] * 10 + torch.arange(row, dtype=torch.float32)[None]
    got = igdist.gather_observations(local, total_envs=total)
    want = torch.arange(total, dtype=torch.float32)[:, None] * 10 + torch.arange(row, dtype=torch.float32)[None]
    q.put((rank, bool(torch.equal(got, want)), tuple(got.shape)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("total", [8, 9])   # equal slices, and ragged slices (padded gather)
def test_gather_observations_world2_gloo(total):
    world, row = 2, 12
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, total, row, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok, shape in res:
        assert ok, f"rank {rank}: gathered rows are not in global env order"
        assert shape == (total, row)


def _worker_obsgather(rank, world, port, total, row, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    igdist.init_from_env(backend="gloo")
    lo, hi = igdist.env_slice(total, rank, world)
    ok = True
    g = igdist.ObsGather(torch.zeros(hi - lo, row), total_envs=total, transport="nccl")
    held = None
    for step in range(1, 5):      # slot rotation: the result of step s stays intact until gather(s + 1) is called
        local = torch.arange(lo, hi, dtype=torch.float32)[:, None] * 10 + step * 1000 + torch.arange(row, dtype=torch.float32)[None]
        want = torch.arange(total, dtype=torch.float32)[:, None] * 10 + step * 1000 + torch.arange(row, dtype=torch.float32)[None]
        got = g.wait(g.gather(local))
        ok = ok and bool(torch.equal(got, want))
        if held is not None:
            ok = ok and bool(torch.equal(held[0], held[1]))      # the previous step's tensor was not overwritten
        held = (got, want.clone())
    g.close()
    q.put((rank, ok))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("total", [8, 9])
def test_obsgather_collective_transport_world2_gloo(total):
    world, row = 2, 12
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_obsgather, args=(r, world, port, total, row, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok in res)


def test_sharded_reference_sampler_word_windows():
    """m words per non-empty env in GLOBAL env order: the ranks' windows tile the job's words without gap or overlap."""
    m, counts = 400, [3, 0, 5, 2]
    wins = [igdist.shard_word_window(counts, r, m) for r in range(len(counts))]
    assert [w[0] for w in wins] == [0, 1200, 1200, 3200] and all(w[1] == 4000 for w in wins)
    for r in range(len(counts) - 1):
        assert wins[r][0] + m * counts[r] == wins[r + 1][0]
    assert igdist.shard_word_window([7], 0, 400) == (0, 2800)


def test_single_process_gather_is_identity():
    x = torch.randn(5, 7)
    assert igdist.gather_observations(x) is x


def test_sharded_synthetic_inputs_equal_the_single_gpu_inputs():
    """bench.make_inputs(E, rank*E, total) over the ranks == make_inputs(total, 0, total): poses, mesh / background
    ids, depth and segmentation frames are functions of the GLOBAL env id, so the work is the same however the envs
    are sharded (the NCCL run then gathers rows that equal the single-GPU rows bit-for-bit, tools/check_gather_nccl.py)."""
    import numpy as np
    import bench
    total, world = 6, 2
    _, P, depth, seg = bench.make_inputs(total, 0, total)
    E = total // world
    for r in range(world):
        _, Pr, dr, sr = bench.make_inputs(E, r * E, total)
        sl = slice(r * E, (r + 1) * E)
        for k in ("finger_pos", "finger_quat", "plug_pos", "plug_quat", "mesh_id", "bg_id"):
            assert np.array_equal(Pr[k], P[k][sl]), k
        assert np.array_equal(sr, seg[sl])
        assert np.array_equal(dr, depth[sl], equal_nan=True)
