"""GPU: task-level observation functions, host pipeline, multi-region path and full-size properties.

The full-size cases (BASELINE configs: 1024 / 4096 envs) are checked through size-independent
properties: the no-contact invariant, determinism, fused-vs-standalone observation equality, the
update mask, FPS index validity and the box constraint of every sampled point.
"""
import os

import numpy as np
import pytest
import torch

from isaacgyminsertion_b200 import _lib, assets, synthetic

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _inputs(n, seed=0):
    packed = assets.load_packed()
    gym = synthetic.SyntheticGym(n, seed=seed)
    P = synthetic.tactile_poses(n, packed, seed=seed)
    _, _, socket_pos = synthetic.scene_poses(n, seed=seed, assets=packed)
    depth, seg = synthetic.external_camera_frames(gym, P["plug_pos"].astype(np.float64),
                                                  P["plug_quat"].astype(np.float64), socket_pos, seed=seed)
    return gym, P, depth, seg


def _load(task, P, depth, seg):
    dev = task.device
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    fp, fq = t(P["finger_pos"]), t(P["finger_quat"])
    task.left_finger_pos, task.right_finger_pos, task.middle_finger_pos = fp[:, 0], fp[:, 1], fp[:, 2]
    task.left_finger_quat, task.right_finger_quat, task.middle_finger_quat = fq[:, 0], fq[:, 1], fq[:, 2]
    task.plug_pos, task.plug_quat = t(P["plug_pos"]), t(P["plug_quat"])
    task.cam_renders, task.seg_renders = t(depth), t(seg)


def _task(n, gym, P, **kw):
    from isaacgyminsertion_b200.task_obs import FactoryTaskInsertionTactileObs
    return FactoryTaskInsertionTactileObs(n, gym, P["mesh_id"], P["bg_id"], device=DEV, **kw)


# --------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["tactile_golden.npz", "tactile_golden_none.npz"])
def test_cuda_matches_golden_fixture(built_lib, golden_dir, name):
    """CUDA path against tests/golden/tactile_golden{,_none}.npz (one per light model); needs no oracle at run time."""
    from isaacgyminsertion_b200.allsight_render import BatchedAllSight
    g = np.load(os.path.join(golden_dir, name))
    n = int(g["n_envs"])
    eng = BatchedAllSight(n, g["mesh_id"], g["bg_id"], device=DEV, falloff=str(g["falloff"]))
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(DEV)
    eng.render(t(g["finger_pos"]), t(g["finger_quat"]), t(g["plug_pos"]), t(g["plug_quat"]), force=float(g["force"]))
    eng.check_overflow()
    assert np.array_equal(eng.depth0.cpu().numpy(), g["depth0"])
    assert np.array_equal(eng._M.cpu().numpy().reshape(-1, 3, 4), g["M"])
    gd = eng.gel_depth.reshape(-1, 224, 224).cpu().numpy()
    assert np.array_equal(gd, g["gel_depth"]), "gel depth must be bit-exact"
    bg = eng.bg_real.cpu().numpy()[g["bg_id"].reshape(-1) - 12].astype(np.int16)
    delta = eng.color.reshape(-1, 224, 224, 3).cpu().numpy().astype(np.int16) - bg
    d = np.abs(delta - g["color_delta"])
    assert d.max() <= 1, "tactile image more than 1/255 off"
    assert (d > 0).mean() < 2e-3
    assert np.abs(eng.obs.reshape(-1, 2048).cpu().numpy() - g["obs"]).max() <= 1.0 / 255 + 1e-6


def test_task_observation_functions_match_oracle(built_lib):
    from oracle import pcl as opcl
    from oracle import tactile as ot
    n = 6
    gym, P, depth, seg = _inputs(n, seed=3)
    task = _task(n, gym, P, sampler="reference")
    _load(task, P, depth, seg)
    ones = torch.ones(n, dtype=torch.bool, device=DEV)
    zeros = torch.zeros(n, dtype=torch.bool, device=DEV)
    task.update_tactile(ones, ones)
    torch.manual_seed(11)
    task.update_external_cam(ones, ones, ones, zeros, zeros)
    # tactile: the reference's serial loop
    model = ot.SensorModel()
    handles = [[ot.OracleAllSight(model, int(P["mesh_id"][e]), int(P["bg_id"][e, k])) for k in range(3)] for e in range(n)]
    fp7 = np.concatenate([P["finger_pos"], P["finger_quat"]], axis=2)
    op7 = np.concatenate([P["plug_pos"], P["plug_quat"]], axis=1)
    want_t = ot.render_tactile_serial(model, handles, fp7, op7, 70)
    assert task.tactile_imgs.shape == (n, 3, 2048)
    assert np.abs(task.tactile_imgs.cpu().numpy() - want_t).max() <= 1.0 / 255 + 1e-6
    assert torch.equal(task.tactile_queue[:, 0], task.tactile_imgs)
    # point cloud: plug then socket from the same torch.randint stream
    torch.manual_seed(11)
    want_p = opcl.pcl_observation(opcl.build_cameras(gym), torch.from_numpy(depth), torch.from_numpy(seg))
    tol = 1e-5 * float(np.abs(gym.origins).max() + 1.0)
    np.testing.assert_allclose(task.pcl.cpu().numpy(), want_p.numpy(), rtol=1e-5, atol=tol)
    assert int(task.got_socket.min()) == 1 and torch.equal(task.pcl_queue[:, 0], task.pcl)
    od = task.obs_dict()
    assert od["tactile"].shape == (n, 1, 3, 2048) and od["pcl"].shape == (n, 1, 2400)
    # reset bookkeeping (factory_task_insertion.py:1753-1777)
    task.reset_idx(torch.tensor([1, 4], device=DEV))
    assert float(task.tactile_queue[1].abs().max()) == 0 and int(task.got_socket[4]) == 0 and int(task.got_socket[0]) == 1


def test_reset_then_masked_step_shows_zeros(built_lib):
    """factory_task_insertion.py:1753-1777 zeroes queues AND current buffers of the reset envs; an env whose
    update flags are off on the next step copies `tactile_queue[e, 0]` (= 0) (:578-579), keeps its zeroed pcl row
    (:1014-1027 writes only `update` rows) and its zeroed seg_buf (:934-940)."""
    n = 6
    gym, P, depth, seg = _inputs(n, seed=3)
    task = _task(n, gym, P, sampler="fps", falloff="none")
    _load(task, P, depth, seg)
    ones = torch.ones(n, dtype=torch.bool, device=DEV)
    zeros = torch.zeros(n, dtype=torch.bool, device=DEV)
    task.compute_observations(ones, ones, ones, ones, ones, zeros, zeros)
    assert float(task.tactile_imgs.abs().sum(-1).min()) > 0 and float(task.pcl.abs().sum(1).min()) > 0
    assert int(task.seg_buf.abs().sum(1).min()) > 0
    before = task.obs_packed.clone()
    ids = torch.tensor([1, 4], device=DEV)
    task.reset_idx(ids)
    for buf in (task.tactile_imgs, task.tactile_queue, task.pcl, task.pcl_queue, task.socket_pcl, task.seg_buf,
                task.got_socket):
        assert float(buf[ids].abs().sum()) == 0, "reset must zero the current buffers, not only the queues"
    upd = ones.clone()
    upd[ids] = False                     # the reset envs are not updated on the next step
    task.compute_observations(upd, ones, upd, ones, ones, zeros, zeros)
    assert float(task.tactile_imgs[ids].abs().sum()) == 0 and float(task.tactile_queue[ids].abs().sum()) == 0
    assert int(task.seg_buf[ids].abs().sum()) == 0
    # pcl: got_socket was cleared, so the socket cloud is recomputed and the row rewritten (`update | restarted`, :988-989)
    assert float(task.pcl[ids].abs().sum()) > 0 and int(task.got_socket.min()) == 1
    keep = torch.tensor([0, 2, 3, 5], device=DEV)
    assert torch.equal(task.obs_packed[keep], before[keep])
    task.compute_observations(ones, ones, ones, ones, ones, zeros, zeros)
    assert torch.equal(task.obs_packed, before)


def test_include_all_pcl_matches_oracle(built_lib):
    """include_all_pcl (FactoryTaskInsertionTactile.yaml:124): the unmasked scene cloud is drawn FIRST from the
    torch.randint stream (factory_task_insertion.py:946-949) and appended LAST to the pcl row (:1014-1027)."""
    from oracle import pcl as opcl
    n, tp = 5, 256
    gym, P, depth, seg = _inputs(n, seed=5)
    task = _task(n, gym, P, sampler="reference", tactile=False, include_all_pcl=True, total_points=tp)
    _load(task, P, depth, seg)
    ones = torch.ones(n, dtype=torch.bool, device=DEV)
    zeros = torch.zeros(n, dtype=torch.bool, device=DEV)
    torch.manual_seed(5)
    task.update_external_cam(ones, ones, ones, zeros, zeros)
    torch.manual_seed(5)
    want = opcl.pcl_observation(opcl.build_cameras(gym), torch.from_numpy(depth), torch.from_numpy(seg),
                                include_all_pcl=True, total_points=tp)
    assert task.pcl.shape == (n, (400 + 400 + tp) * 3)
    tol = 1e-5 * float(np.abs(gym.origins).max() + 1.0)
    np.testing.assert_allclose(task.pcl.cpu().numpy(), want.numpy(), rtol=1e-5, atol=tol)
    # FPS sampler on the same clouds: every all-scene sample is one of the env's box-filtered scene points
    task2 = _task(n, gym, P, sampler="fps", tactile=False, include_all_pcl=True, total_points=tp)
    _load(task2, P, depth, seg)
    task2.update_external_cam(ones, ones, ones, zeros, zeros)
    allc = task2.pcl.view(n, -1, 3)[:, 800:].cpu().numpy()
    cams = opcl.build_cameras(gym)
    scene = opcl.get_ptd(cams, torch.from_numpy(depth), opcl.filter_pts)
    for e in range(n):
        pool = scene[e].numpy()
        assert pool.shape[0] > tp
        d = np.abs(allc[e][:, None, :] - pool[None, :, :]).max(-1).min(-1)
        assert d.max() <= tol
        assert len(np.unique(allc[e], axis=0)) == tp      # FPS never repeats while unpicked points remain


def test_multi_region_path_is_identical(built_lib):
    """Small region budgets force the contact kernel to cut every window into several regions (obs then
    comes from the global-memory path); results must not depend on the cut."""
    from isaacgyminsertion_b200.allsight_render import BatchedAllSight
    n = 7
    packed = assets.load_packed()
    P = synthetic.tactile_poses(n, packed, seed=2)
    eng = BatchedAllSight(n, P["mesh_id"], P["bg_id"], device=DEV, falloff="none")
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(DEV)
    args = (t(P["finger_pos"]), t(P["finger_quat"]), t(P["plug_pos"]), t(P["plug_quat"]))
    eng.render(*args)
    ref = (eng.color.clone(), eng.gel_depth.clone(), eng.obs.clone())
    assert int((eng.contact_counts() > 0).sum()) >= 5
    for budget in (49 + 15, 700, 3000):
        eng.region_budget = budget          # IgiTactileFrames.region_budget (test hook, per call)
        eng.color.zero_(); eng.gel_depth.fill_(-1); eng.obs.zero_()
        eng.render(*args)
        assert torch.equal(eng.color, ref[0]) and torch.equal(eng.gel_depth, ref[1]), budget
        assert torch.equal(eng.obs, ref[2]), budget


def test_host_pipeline_matches_direct_calls(built_lib):
    from isaacgyminsertion_b200.pipeline import HostObsPipeline
    n = 12
    gym, P, depth, seg = _inputs(n, seed=4)
    task = _task(n, gym, P, sampler="fps")
    _load(task, P, depth, seg)
    ones = torch.ones(n, dtype=torch.bool, device=DEV)
    zeros = torch.zeros(n, dtype=torch.bool, device=DEV)
    task.update_tactile(ones, ones)
    task.update_external_cam(ones, ones, ones, zeros, zeros)
    want = task.obs_packed.cpu().clone()
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    host = [pin(P[k]) for k in ("finger_pos", "finger_quat", "plug_pos", "plug_quat")] + [pin(depth), pin(seg)]
    # second set of poses so that consecutive steps differ
    P2 = dict(P)
    P2["finger_pos"] = P["finger_pos"] + np.float32(0.0015)
    # (under the default light model the tactile image never leaves the background, so the cloud must differ too)
    host2 = [pin(P2[k]) for k in ("finger_pos", "finger_quat", "plug_pos", "plug_quat")] + [pin(depth * np.float32(1.01)), pin(seg)]
    pipe = HostObsPipeline(task, sampler_socket_every_step=True)
    pending = [pipe.step(*(host if i % 2 == 0 else host2)) for i in range(5)]
    outs = [p.wait().clone() for p in pending[-3:]]      # ring of 3 slots: the last three are still valid
    assert torch.equal(outs[0], want) and torch.equal(outs[2], want)      # steps 2 and 4 used `host`
    assert not torch.equal(outs[1], want)
    assert torch.equal(pipe.flush(), want)
    # the host may keep the segmentation ids as uint8: a quarter of the upload, same observations
    host8 = host[:5] + [torch.from_numpy(seg.astype(np.uint8)).pin_memory()]
    assert torch.equal(pipe.step(*host8).wait(), want) and pipe.last_h2d_bytes < sum(t.numel() * t.element_size() for t in host)


def test_graphed_step_equals_eager_step(built_lib):
    """The CUDA-graph replay of compute_observations (pipeline.GraphedObsStep) gives the eager step's rows bit for
    bit, follows new input values written into the captured tensors, and keeps the update-mask semantics."""
    from isaacgyminsertion_b200.pipeline import GraphedObsStep
    n = 10                                 # the reference's visuotactile operating point (scripts/train_s3.sh:5)
    gym, P, depth, seg = _inputs(n, seed=6)
    task = _task(n, gym, P, sampler="fps", falloff="none")
    _load(task, P, depth, seg)
    ones = torch.ones(n, dtype=torch.bool, device=DEV)
    zeros = torch.zeros(n, dtype=torch.bool, device=DEV)
    task.invalidate_socket_cache()
    task.compute_observations(ones, ones, ones, ones, ones, zeros, zeros)
    want = task.obs_packed.clone()
    upd = ones.clone()
    g = GraphedObsStep(task, masks=(upd, ones, upd, ones, ones, zeros, zeros), socket_every_step=True)
    task.obs_packed.zero_()
    assert torch.equal(g(), want)
    # new poses / depth written INTO the captured input tensors
    fp_all = task.left_finger_pos._base if task.left_finger_pos._base is not None else task.left_finger_pos
    fp_all += 0.0015
    task.cam_renders *= 1.01
    task.invalidate_socket_cache()
    task.compute_observations(ones, ones, ones, ones, ones, zeros, zeros)
    want2 = task.obs_packed.clone()
    assert not torch.equal(want2, want)
    task.obs_packed.zero_()
    assert torch.equal(g(), want2)
    # masks are read from the captured tensors too: envs switched off keep their rows
    upd[::2] = False
    fp_all -= 0.0015
    g()
    assert torch.equal(task.obs_packed[::2, :6144], want2[::2, :6144])          # tactile part untouched
    assert torch.equal(task.obs_packed[1::2, :6144], want[1::2, :6144])
    assert g.check_overflow() is False


# --------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n_envs,falloff", [(1024, "inverse_square"), (4096, "inverse_square"), (4096, "none")])
def test_full_size_tactile_properties(built_lib, n_envs, falloff):
    gym, P, depth, seg = _inputs(n_envs, seed=0)
    task = _task(n_envs, gym, P, sampler="fps", pcl_cam=False, falloff=falloff)
    _load(task, P, depth, seg)
    eng = task.tactile_engine
    ones = torch.ones(n_envs, dtype=torch.bool, device=DEV)
    task.update_tactile(ones, ones)
    eng.check_overflow()
    counts = eng.contact_counts().reshape(-1)
    F = 3 * n_envs
    obs = task.tactile_imgs.reshape(F, 2048)
    color = eng.color.reshape(F, 224, 224, 3)
    gd = eng.gel_depth.reshape(F, 224, 224)
    # contact mix of the synthetic workload: visible imprints in a good half of the frames
    touched = (gd != 0).flatten(1).any(1)
    assert 0.4 < float(touched.float().mean()) < 0.9
    # no-contact invariant: colour == real background frame, depth 0, obs == the empty observation
    idle = ~touched
    bg = eng.bg_real[eng.bg_index.long()]
    assert torch.equal(color[idle], bg[idle])
    assert torch.equal(obs[idle], eng.obs_empty[None].expand(int(idle.sum()), -1))
    assert bool((counts[touched] > 0).all())
    # gel depth: >= 0, zero outside the candidate window, never deeper than the gel itself
    assert float(gd.min()) >= 0.0 and float((gd - eng.depth0[None]).max()) <= 0.0
    # the fused observation equals the standalone K3 kernel applied to the colour image (bit-exact)
    lib = _lib.load()
    obs2 = torch.empty((F, 2048), dtype=torch.float32, device=DEV)
    import ctypes as c
    _lib.check(lib.igi_tactile_obs(_lib.dptr(color), _lib.dptr(eng.bg_real), _lib.dptr(eng.bg_index), c.c_int(F),
                                   _lib.dptr(obs2), c.c_int64(2048), _lib.stream_ptr(eng.device)), "igi_tactile_obs")
    assert torch.equal(obs2, obs)
    # determinism and the update mask
    snap = (obs.clone(), color[::97].clone())
    gd_snap = gd.reshape(n_envs, 3, 224, 224)[::16].clone()
    task.update_tactile(ones, ones)
    assert torch.equal(task.tactile_imgs.reshape(F, 2048), snap[0]) and torch.equal(color[::97], snap[1])
    task.plug_pos = task.plug_pos + 0.001
    half = torch.zeros(n_envs, dtype=torch.bool, device=DEV)
    half[: n_envs // 2] = True
    task.update_tactile(half, ones)
    now = task.tactile_imgs.reshape(n_envs, 3, 2048)
    assert torch.equal(now[n_envs // 2:], snap[0].reshape(n_envs, 3, 2048)[n_envs // 2:])
    gd_now = eng.gel_depth.reshape(n_envs, 3, 224, 224)[::16]
    k = gd_now.shape[0] // 2
    assert torch.equal(gd_now[k:], gd_snap[k:]) and not torch.equal(gd_now[:k], gd_snap[:k])
    if falloff == "none":      # under inverse_square every fragment clips like the gel behind it: the image stays the background
        assert not torch.equal(now[: n_envs // 2], snap[0].reshape(n_envs, 3, 2048)[: n_envs // 2])


@pytest.mark.parametrize("n_envs", [1024, 4096])
def test_full_size_pcl_properties(built_lib, n_envs):
    from isaacgyminsertion_b200.pcl_utils import filter_pts
    gym, P, depth, seg = _inputs(n_envs, seed=0)
    task = _task(n_envs, gym, P, sampler="fps", tactile=False)
    _load(task, P, depth, seg)
    eng = task.pcl_generator.engine
    d, s = task.cam_renders, task.seg_renders
    pts, cnt, any_ = eng.compact(d, s, (2, 3), filter_pts.box)
    # counts against a vectorised torch restatement of mask -> depth test -> unproject -> box
    for c, sid in enumerate((2, 3)):
        md = d.reshape(n_envs, -1) * (s.reshape(n_envs, -1) == sid)
        valid = md > -1.0
        H, W = gym.height, gym.width
        px = eng.uvx[:, None, :].expand(-1, H, -1).reshape(n_envs, -1) * md
        py = eng.uvy[:, :, None].expand(-1, -1, W).reshape(n_envs, -1) * md
        pz = eng.uvz[:, None] * md
        hom = torch.stack([px, py, pz, torch.ones_like(px)], -1)
        w = hom @ eng.ext.reshape(n_envs, 4, 4)
        o = (w @ eng.e2g_inv.reshape(n_envs, 4, 4).transpose(1, 2))[..., :3]
        b = filter_pts.box
        keep = valid & (o[..., 0] >= b[0]) & (o[..., 0] <= b[1]) & (o[..., 1] >= b[2]) & (o[..., 1] <= b[3]) & \
            (o[..., 2] >= b[4]) & (o[..., 2] <= b[5])
        # a point within an ulp of a box face may fall on either side in the batched matmul
        diff = (keep.sum(1).int() - cnt[:, c]).abs()
        assert int(diff.max()) <= 3 and float((diff != 0).float().mean()) < 0.05, (int(diff.max()), float((diff != 0).float().mean()))
    for c, m in ((0, 400), (1, 400)):
        out, idx = eng.sample_fps(pts, cnt, any_, c, m, return_idx=True)
        n = cnt[:, c].long()
        assert bool((idx >= 0).all()) and bool((idx < n[:, None].clamp(min=1)).all())
        # farthest-point property: while distinct points remain, no index repeats
        srt = idx.sort(1).values
        distinct = (srt[:, 1:] != srt[:, :-1]).sum(1) + 1
        assert torch.equal(distinct, torch.minimum(n, torch.full_like(n, m)).clamp(min=1))
        # gathered points are the compacted points at those indices, and lie inside the workspace box
        gathered = torch.gather(pts[:, c], 1, idx.long()[..., None].expand(-1, -1, 3))
        live = (any_[:, c] != 0) & (n > 0)
        assert torch.equal(out[live], gathered[live]) and float(out[~live].abs().sum()) == 0.0
        b = filter_pts.box
        ol = out[live]
        assert float(ol[..., 0].min()) >= b[0] and float(ol[..., 0].max()) <= b[1]
        assert float(ol[..., 2].min()) >= b[4] and float(ol[..., 2].max()) <= b[5]
    # task level: determinism, socket cache and update mask
    ones = torch.ones(n_envs, dtype=torch.bool, device=DEV)
    zeros = torch.zeros(n_envs, dtype=torch.bool, device=DEV)
    task.update_external_cam(ones, ones, ones, zeros, zeros)
    a = task.pcl.clone()
    assert not task._socket_pending and int(task.got_socket.min()) == 1
    task.update_external_cam(ones, ones, ones, zeros, zeros)      # socket cloud now cached
    assert torch.equal(task.pcl, a)
    task.cam_renders = task.cam_renders * 1.01
    task.update_external_cam(zeros, ones, ones, zeros, zeros)     # update_freq off: rows untouched
    assert torch.equal(task.pcl, a)
    task.update_external_cam(ones, ones, ones, zeros, zeros)
    assert not torch.equal(task.pcl[:, :1200], a[:, :1200]) and torch.equal(task.pcl[:, 1200:], a[:, 1200:])


def test_edge_cases(built_lib):
    """Empty batch, a single env, and envs whose cloud is empty (all-miss depth)."""
    from isaacgyminsertion_b200.pcl_utils import filter_pts
    lib = _lib.load()
    n = 3
    gym, P, depth, seg = _inputs(n, seed=7)
    depth = depth.copy()
    depth[1] = -np.inf                                   # camera sees nothing in env 1
    seg = seg.copy()
    seg[2][seg[2] == 3] = 0                              # no socket pixels in env 2
    task = _task(n, gym, P, sampler="fps")
    _load(task, P, depth, seg)
    ones = torch.ones(n, dtype=torch.bool, device=DEV)
    zeros = torch.zeros(n, dtype=torch.bool, device=DEV)
    task.update_external_cam(ones, ones, ones, zeros, zeros)
    pcl = task.pcl.reshape(n, 800, 3)
    assert float(pcl[1].abs().sum()) == 0.0              # pcl_utils.py:175-183: zeros when nothing is visible
    assert float(pcl[2, 400:].abs().sum()) == 0.0 and float(pcl[2, :400].abs().sum()) > 0
    assert float(pcl[0].abs().sum()) > 0
    # zero-sized batches are accepted by the C-ABI without a launch
    before = lib.igi_launch_count()
    import ctypes as c
    buf = torch.zeros((n, 400, 3), device=DEV)
    rc = lib.igi_fps(_lib.dptr(buf), c.c_int64(1200), None, None, c.c_int64(1), c.c_int(400), c.c_int(0),
                     c.c_int(16), _lib.dptr(buf), c.c_int64(1200), None, c.c_int(0), _lib.stream_ptr(task.device))
    assert rc == 0 and lib.igi_launch_count() == before
    rc = lib.igi_fps_balanced(_lib.dptr(buf), c.c_int64(1200), _lib.dptr(task.got_socket), None, c.c_int64(1), c.c_int(0),
                              c.c_int(16), _lib.dptr(buf), c.c_int64(1200), None, _lib.dptr(task.got_socket),
                              c.c_int(0), _lib.stream_ptr(task.device))
    assert rc == 0 and lib.igi_launch_count() == before
