"""GPU parity: the sm_100a (P) kernels through the C-ABI against the CPU oracle and the
golden vectors of the real reference.

Tolerances (north star): segmentation masks, counts, sample / FPS indices bit-exact;
point coordinates within 1e-5 relative.  The reference evaluates `pts @ inv(view)` in
GLOBAL coordinates and then subtracts the env origin (pcl_utils.py:83-85), so the relative
bound applies to the global-coordinate magnitude: atol = 1e-5 * max|global coordinate|.
"""
import os

import numpy as np
import pytest
import torch

from isaacgyminsertion_b200 import synthetic
from oracle import fps as ofps
from oracle import pcl as opcl

pytestmark = pytest.mark.gpu


def _tol(gym):
    return 1e-5 * float(np.abs(gym.origins).max() + 1.0)


@pytest.fixture(scope="module")
def golden(golden_dir):
    return np.load(os.path.join(golden_dir, "pcl_golden.npz"))


@pytest.fixture(scope="module")
def cam(golden, built_lib):
    from isaacgyminsertion_b200.pcl_utils import CameraPointCloud, filter_pts
    gym = synthetic.SyntheticGym(golden["depth"].shape[0], seed=int(golden["seed"]))
    gen = CameraPointCloud(None, gym, gym.envs, gym.camera_handles, gym.camera_props, sample_num=400,
                           filter_func=filter_pts, pt_in_local=True, graphics_device="cuda:0",
                           compute_device="cuda:0")
    return gym, gen


def test_golden_reference_sampler(golden, cam):
    gym, gen = cam
    depth = torch.from_numpy(golden["depth"]).cuda()
    seg = torch.from_numpy(golden["seg"]).cuda()
    N = depth.shape[0]
    torch.manual_seed(42)
    plug_depth = (depth.flatten(1) * (seg.flatten(1) == 2)).reshape(N, gym.height, gym.width)
    plug = gen.get_point_cloud(plug_depth, sample_num=400)
    # fused masking variant for the socket
    socket = gen.get_point_cloud(depth, sample_num=400, seg=seg, seg_id=3)
    probe = torch.randint(0, 1 << 20, (8,))
    tol = _tol(gym)
    np.testing.assert_allclose(plug.cpu().numpy(), golden["plug"], rtol=1e-5, atol=tol)
    np.testing.assert_allclose(socket.cpu().numpy(), golden["socket"], rtol=1e-5, atol=tol)
    assert np.array_equal(probe.numpy(), golden["rng_probe"]), "CPU generator not advanced like the reference"
    assert not plug[5].any() and not plug[6].any() and not socket[7].any()


def test_golden_compaction_counts_and_order(golden, cam):
    gym, gen = cam
    depth = torch.from_numpy(golden["depth"]).cuda()
    seg = torch.from_numpy(golden["seg"]).cuda()
    from isaacgyminsertion_b200.pcl_utils import filter_pts
    pts, cnt, any_ = gen.engine.compact(depth, seg, (2, 3), filter_pts.box)
    cnt = cnt.cpu().numpy()
    assert np.array_equal(cnt[:, 0], golden["plug_counts"])
    assert np.array_equal(cnt[:, 1], golden["socket_counts"])
    assert np.array_equal(any_.cpu().numpy(), (cnt > 0).astype(np.int32))
    got = torch.cat([pts[e, 0, :cnt[e, 0]] for e in range(len(cnt))]).cpu().numpy()
    np.testing.assert_allclose(got, golden["plug_all"], rtol=1e-5, atol=_tol(gym))
    got = torch.cat([pts[e, 1, :cnt[e, 1]] for e in range(len(cnt))]).cpu().numpy()
    np.testing.assert_allclose(got, golden["socket_all"], rtol=1e-5, atol=_tol(gym))


def test_get_ptd_cuda_and_convert(golden, cam):
    gym, gen = cam
    from isaacgyminsertion_b200.pcl_utils import PointCloudGenerator
    depth = torch.from_numpy(golden["depth"]).cuda()
    from isaacgyminsertion_b200.pcl_utils import CameraPointCloud
    raw_gen = CameraPointCloud(None, gym, gym.envs, gym.camera_handles, gym.camera_props, sample_num=400,
                               filter_func=None, pt_in_local=True, graphics_device="cuda:0",
                               compute_device="cuda:0")
    lst = raw_gen.get_ptd_cuda(depth[:3], env_ids=[0, 1, 2])
    np.testing.assert_allclose(lst[0].cpu().numpy(), golden["unfiltered0"], rtol=1e-5, atol=_tol(gym))
    o = gym.get_env_origin(0)
    e2g = np.identity(4)
    e2g[:3, 3] = [o.x, o.y, o.z]
    g1 = PointCloudGenerator(gym.get_camera_proj_matrix(None, 0, 0), gym.get_camera_view_matrix(None, 0, 0),
                             e2g, camera_props=gym.camera_props[0], depth_max=1.0, device="cuda:0")
    np.testing.assert_allclose(g1.convert(depth[0]).cpu().numpy(), golden["unfiltered0"], rtol=1e-5,
                               atol=_tol(gym))


def test_proc_pts_uses_the_envs_own_camera(golden, cam):
    """`_proc_pts(env_id, depth_image)` (pcl_utils.py:186-193) takes ONE image and that env's tables; any
    env_id > 0 must work and equal the env's entry of the batched `get_ptd_cuda`."""
    gym, gen = cam
    depth = torch.from_numpy(golden["depth"]).cuda()
    lst = gen.get_ptd_cuda(depth)
    for e in (0, 1, depth.shape[0] - 1):
        one = gen._proc_pts(e, depth[e])
        assert torch.equal(one, lst[e]) and one.shape[0] > 0
    sub = gen.get_ptd_cuda(depth, env_ids=[2, 0])
    assert torch.equal(sub[0], lst[2]) and torch.equal(sub[1], lst[0])


@pytest.mark.parametrize("n_envs,seed", [(1, 0), (67, 1), (300, 2)])
def test_seeded_scenes_vs_oracle(built_lib, n_envs, seed):
    from isaacgyminsertion_b200.pcl_utils import CameraPointCloud, filter_pts
    gym = synthetic.SyntheticGym(n_envs, seed=seed)
    pp, pq, sp = synthetic.scene_poses(n_envs, seed=seed)
    depth, seg = synthetic.external_camera_frames(gym, pp, pq, sp, seed=seed)
    cams = opcl.build_cameras(gym)
    d_t, s_t = torch.from_numpy(depth), torch.from_numpy(seg)
    torch.manual_seed(5)
    want = opcl.pcl_observation(cams, d_t, s_t)
    gen = CameraPointCloud(None, gym, gym.envs, gym.camera_handles, gym.camera_props, sample_num=400,
                           filter_func=filter_pts, pt_in_local=True, graphics_device="cuda:0",
                           compute_device="cuda:0")
    torch.manual_seed(5)
    plug = gen.get_point_cloud(d_t.cuda(), seg=s_t.cuda(), seg_id=2)
    socket = gen.get_point_cloud(d_t.cuda(), seg=s_t.cuda(), seg_id=3)
    got = torch.cat([plug, socket], dim=1).flatten(1).cpu()
    np.testing.assert_allclose(got.numpy(), want.numpy(), rtol=1e-5, atol=_tol(gym))


def test_fps_pipeline_indices_bit_exact(golden, cam):
    gym, gen = cam
    from isaacgyminsertion_b200.pcl_utils import filter_pts
    depth = torch.from_numpy(golden["depth"]).cuda()
    seg = torch.from_numpy(golden["seg"]).cuda()
    pts, cnt, any_ = gen.engine.compact(depth, seg, (2, 3), filter_pts.box)
    for cls in (0, 1):
        out, idx = gen.engine.sample_fps(pts, cnt, any_, cls, 400, return_idx=True)
        c = cnt[:, cls].cpu().numpy()
        lst = [pts[e, cls, :c[e]].cpu().numpy() for e in range(len(c))]
        want_pts, want_idx = ofps.fps_batch(lst, 400)
        assert np.array_equal(idx.cpu().numpy(), want_idx)
        assert np.array_equal(out.cpu().numpy(), want_pts)
        # the static-assignment entry point (igi_fps) gives the same result
        out_s, idx_s = gen.engine.sample_fps_static(pts, cnt, any_, cls, 400, return_idx=True)
        assert torch.equal(idx_s, idx) and torch.equal(out_s, out)
    # every class in one size-ordered launch: (n, C, m, 3), the packed [plug | socket] layout
    both, idx_b = gen.engine.sample_fps(pts, cnt, any_, None, 400, return_idx=True)
    for cls in (0, 1):
        out, idx = gen.engine.sample_fps(pts, cnt, any_, cls, 400, return_idx=True)
        assert torch.equal(both[:, cls], out) and torch.equal(idx_b[:, cls], idx)


def test_fps_balanced_mixed_sizes(built_lib):
    """Size-ordered schedule over tasks of every size class (dead, warp-resident, cooperative, block
    kernel), interleaved so that the ordered list differs from the task order."""
    import ctypes as c
    from isaacgyminsertion_b200 import _lib
    lib = _lib.load()
    sizes = [0, 1, 5, 100, 128, 129, 160, 161, 224, 256, 257, 384, 385, 512, 513, 640, 700, 768, 769, 1024, 1025, 3000, 17, 400, 64, 383, 1023, 2, 96]
    rng = np.random.default_rng(11)
    T, cap, m = 3 * len(sizes), 3000, 400
    counts = np.array([sizes[(7 * t) % len(sizes)] for t in range(T)], dtype=np.int32)
    anyf = np.ones(T, dtype=np.int32)
    anyf[5] = 0                                          # a dead task with points
    pts = (rng.random((T, cap, 3)) * 0.5 + 0.1).astype(np.float32)
    pts[3, 50:100] = pts[3, 0:50]                        # exact ties
    d_pts = torch.from_numpy(pts).cuda()
    d_cnt, d_any = torch.from_numpy(counts).cuda(), torch.from_numpy(anyf).cuda()
    out = torch.full((T, m, 3), -7.0, device="cuda")
    idx = torch.full((T, m), -7, dtype=torch.int32, device="cuda")
    scratch = torch.empty(T + 8, dtype=torch.int32, device="cuda")
    rc = lib.igi_fps_balanced(_lib.dptr(d_pts), c.c_int64(cap * 3), _lib.dptr(d_cnt), _lib.dptr(d_any), c.c_int64(1),
                              c.c_int(T), c.c_int(m), _lib.dptr(out), c.c_int64(m * 3), _lib.dptr(idx),
                              _lib.dptr(scratch), c.c_int(0), _lib.stream_ptr(d_pts.device))
    _lib.check(rc, "igi_fps_balanced")
    sched = scratch[:5].cpu().numpy()
    live = (counts > 0) & (anyf != 0)
    assert sched[4] == int(((counts > 1024) & live).sum())
    assert sched[1] == T - sched[4] and sched[0] == int(((counts > 256) & (counts <= 1024) & live).sum())   # FW_MAXN = 256
    order = scratch[8:8 + sched[1]].cpu().numpy()
    assert len(set(order.tolist())) == sched[1]          # every listed task exactly once
    key = np.where(live, counts, 0)[order]
    assert np.all(((key[:-1] - 1).clip(0) >> 4) >= ((key[1:] - 1).clip(0) >> 4))   # longest first (16-point buckets)
    got_idx, got = idx.cpu().numpy(), out.cpu().numpy()
    for t in range(T):
        if not live[t]:
            assert not got_idx[t].any() and not got[t].any(), t
            continue
        want = ofps.furthest_point_sample(pts[t, :counts[t]], m)
        assert np.array_equal(got_idx[t], want), (t, counts[t])
        assert np.array_equal(got[t], pts[t][want]), t


@pytest.mark.parametrize("n,m", [(1, 4), (2, 5), (37, 16), (37, 64), (128, 400), (129, 300), (384, 400), (385, 400),
                                 (400, 400), (513, 64), (1024, 400), (1025, 50), (2048, 33), (3000, 400), (5184, 400),
                                 (8192, 40), (8193, 24)])
def test_fps_standalone(built_lib, n, m):
    from isaacgyminsertion_b200.pcl_utils import furthest_point_sample
    rng = np.random.default_rng(n)
    B = 5
    pts = (rng.random((B, n, 3)) * 0.5 + 0.1).astype(np.float32)
    if n >= 37:
        pts[1, n // 2:] = pts[1, : n - n // 2]      # duplicated points -> exact distance ties
        pts[2, :] = pts[2, 0]                       # all identical
        pts[3, ::3] = 0.0                           # |p|^2 <= 1e-3: never candidates
    idx = furthest_point_sample(torch.from_numpy(pts).cuda(), m).cpu().numpy()
    for b in range(B):
        assert np.array_equal(idx[b], ofps.furthest_point_sample(pts[b], m)), f"batch {b}"


def test_fps_matches_golden_vectors(built_lib, golden_dir):
    """tests/golden/fps_golden.npz (indices from the literal emulation of the published kernel, tools/make_golden_fps.py):
    one warp (37), the cooperative CTA (400), the cluster and the one-CTA kernel (5184) all reproduce them bit for bit."""
    from isaacgyminsertion_b200.pcl_utils import furthest_point_sample
    g = np.load(os.path.join(golden_dir, "fps_golden.npz"))
    for name in sorted(k[:-4] for k in g.files if k.endswith("_idx")):
        want = g[name + "_idx"]
        d = torch.from_numpy(g[name + "_pts"]).cuda()[None]
        for flags in (0, 1):      # 1 = IGI_FPS_NO_CLUSTER
            got = furthest_point_sample(d, len(want), flags=flags).cpu().numpy()[0]
            assert np.array_equal(got, want), (name, flags)


def test_fps_cluster_equals_block_kernel(built_lib):
    """1025..8192 points: the thread-block-cluster kernel and the one-CTA kernel pick the same indices."""
    from isaacgyminsertion_b200 import _lib
    from isaacgyminsertion_b200.pcl_utils import furthest_point_sample
    lib = _lib.load()
    rng = np.random.default_rng(3)
    for n, m in ((1300, 200), (4097, 128), (6500, 64)):
        pts = (rng.random((7, n, 3)) * 0.5 + 0.1).astype(np.float32)
        pts[4, 100:200] = pts[4, 0:100]
        d = torch.from_numpy(pts).cuda()
        a = furthest_point_sample(d, m).cpu().numpy()
        b = furthest_point_sample(d, m, flags=1).cpu().numpy()      # IGI_FPS_NO_CLUSTER
        assert np.array_equal(a, b), (n, m)
        assert np.array_equal(a[4], ofps.furthest_point_sample(pts[4], m))


def test_fps_balanced_long_list_of_big_tasks(built_lib):
    """More than 512 tasks above 1024 points: the cluster kernel steps aside (device-side switch on the schedule's
    big-task count) and the one-CTA kernel serves them; 512 or fewer: the clusters do.  Same indices either way."""
    import ctypes as c
    from isaacgyminsertion_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(5)
    cap, m = 1100, 12
    for T in (520, 40):
        pts = (rng.random((T, cap, 3)) * 0.5 + 0.1).astype(np.float32)
        counts = np.full(T, 1030, dtype=np.int32)
        counts[::7] = 300                                   # some resident-size tasks in between
        d_pts, d_cnt = torch.from_numpy(pts).cuda(), torch.from_numpy(counts).cuda()
        d_any = torch.ones(T, dtype=torch.int32, device="cuda")
        idx = torch.full((T, m), -7, dtype=torch.int32, device="cuda")
        scratch = torch.empty(T + 8, dtype=torch.int32, device="cuda")
        rc = lib.igi_fps_balanced(_lib.dptr(d_pts), c.c_int64(cap * 3), _lib.dptr(d_cnt), _lib.dptr(d_any), c.c_int64(1),
                                  c.c_int(T), c.c_int(m), None, c.c_int64(m * 3), _lib.dptr(idx), _lib.dptr(scratch),
                                  c.c_int(0), _lib.stream_ptr(d_pts.device))
        _lib.check(rc, "igi_fps_balanced")
        assert int(scratch[4]) == int((counts > 1024).sum())
        got = idx.cpu().numpy()
        for t in list(range(0, T, 37)) + [T - 1]:
            assert np.array_equal(got[t], ofps.furthest_point_sample(pts[t, :counts[t]], m)), (T, t)
        assert (got >= 0).all()


def test_non_tma_path_matches(built_lib):
    """H*W*4 not a multiple of 16 -> plain loads instead of the bulk-copy path."""
    from isaacgyminsertion_b200.pcl_utils import BatchedPointCloud
    gym = synthetic.SyntheticGym(9, seed=4, width=45, height=31)
    pp, pq, sp = synthetic.scene_poses(9, seed=4)
    depth, seg = synthetic.external_camera_frames(gym, pp, pq, sp, seed=4)
    cams = opcl.build_cameras(gym)
    eng = BatchedPointCloud(gym._proj, gym._view, [np.block([[np.eye(3), gym.origins[e][:, None]],
                                                             [np.zeros((1, 3)), np.ones((1, 1))]])
                                                   for e in range(9)], 45, 31, device="cuda:0")
    pts, cnt, _ = eng.compact(torch.from_numpy(depth).cuda(), torch.from_numpy(seg).cuda(), (2,),
                              opcl_box())
    want = opcl.get_ptd(cams, opcl.masked_depth(torch.from_numpy(depth), torch.from_numpy(seg), 2))
    assert [len(w) for w in want] == cnt[:, 0].tolist()
    for e in range(9):
        np.testing.assert_allclose(pts[e, 0, :len(want[e])].cpu().numpy(), want[e].numpy(), rtol=1e-5,
                                   atol=_tol(gym))


def opcl_box():
    from isaacgyminsertion_b200.pcl_utils import filter_pts
    return filter_pts.box
