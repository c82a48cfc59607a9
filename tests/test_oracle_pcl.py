"""CPU: the (P) oracle against golden vectors produced by the REAL reference classes
(tools/make_golden_pcl.py), plus the host-side RNG mirror and the FPS oracle."""
import os

import numpy as np
import pytest
import torch

from isaacgyminsertion_b200 import synthetic
from isaacgyminsertion_b200.pcl_utils import TorchCpuRandintStream, uv_table, filter_pts as box_filter
from oracle import fps as ofps
from oracle import pcl as opcl


@pytest.fixture(scope="module")
def golden(golden_dir):
    return np.load(os.path.join(golden_dir, "pcl_golden.npz"))


@pytest.fixture(scope="module")
def scene(golden):
    gym = synthetic.SyntheticGym(golden["depth"].shape[0], seed=int(golden["seed"]))
    return gym, opcl.build_cameras(gym)


def test_synthetic_frames_are_reproducible(golden):
    N = golden["depth"].shape[0]
    gym = synthetic.SyntheticGym(N, seed=3)
    pp, pq, sp = synthetic.scene_poses(N, seed=3)
    depth, seg = synthetic.external_camera_frames(gym, pp, pq, sp, seed=3)
    keep = [e for e in range(N) if e not in (5, 6, 7, 8)]
    assert np.array_equal(depth[keep], golden["depth"][keep])
    assert np.array_equal(seg[keep], golden["seg"][keep])


def test_camera_tables_match_reference(golden, scene):
    gym, cams = scene
    assert np.array_equal(cams[0].uv_one_in_cam.numpy(), golden["uv_table0"])
    assert np.array_equal(cams[0].ext_mat.numpy(), golden["ext0"])
    assert np.array_equal(uv_table(gym.get_camera_proj_matrix(None, 0, 0), gym.width, gym.height).numpy(),
                          golden["uv_table0"])


def test_convert_matches_reference(golden, scene):
    _, cams = scene
    pts = cams[0].convert(torch.from_numpy(golden["depth"][0]))
    assert np.array_equal(pts.numpy(), golden["unfiltered0"])


def test_point_cloud_matches_reference_bit_exact(golden, scene):
    _, cams = scene
    depth = torch.from_numpy(golden["depth"])
    seg = torch.from_numpy(golden["seg"])
    torch.manual_seed(42)
    plug, _, plug_all = opcl.get_point_cloud(cams, opcl.masked_depth(depth, seg, 2), 400, return_idx=True)
    socket, _, socket_all = opcl.get_point_cloud(cams, opcl.masked_depth(depth, seg, 3), 400, return_idx=True)
    probe = torch.randint(0, 1 << 20, (8,))
    assert np.array_equal(np.array([len(p) for p in plug_all]), golden["plug_counts"])
    assert np.array_equal(np.array([len(p) for p in socket_all]), golden["socket_counts"])
    assert np.array_equal(torch.cat(plug_all).numpy(), golden["plug_all"])
    assert np.array_equal(torch.cat(socket_all).numpy(), golden["socket_all"])
    assert np.array_equal(plug.numpy(), golden["plug"])
    assert np.array_equal(socket.numpy(), golden["socket"])
    assert np.array_equal(probe.numpy(), golden["rng_probe"])
    # degenerate envs: empty cloud -> all zeros (pcl_utils.py:175,179-183)
    assert not plug[5].any() and not plug[6].any() and not socket[7].any()


def test_box_filter_equals_reference_filter():
    g = torch.Generator().manual_seed(0)
    pts = torch.rand((5000, 3), generator=g) * torch.tensor([1.0, 1.2, 0.8]) - torch.tensor([0.1, 0.6, 0.1])
    pts[:5] = torch.tensor([[0.1, -0.4, 0.001], [0.7, 0.4, 0.6], [0.7001, 0, 0.1], [0.3, 0, 0.0009], [0.3, 0.41, 0.2]])
    assert torch.equal(box_filter(pts), opcl.filter_pts(pts))


def test_rng_stream_mirror_matches_torch_randint():
    """`torch.randint(0, n, (m,))` == raw MT19937 words % n, and commit() advances torch identically."""
    stream = TorchCpuRandintStream()
    torch.manual_seed(1234)
    torch.rand(7)  # arbitrary position inside the state block
    counts = [1, 5, 399, 400, 5184, 77, 1000, 3]
    raw = stream.peek(len(counts) * 400)
    mine = [raw[i * 400:(i + 1) * 400].astype(np.int64) % c for i, c in enumerate(counts)]
    ref = [torch.randint(0, c, (400,)).numpy() for c in counts]
    for a, b in zip(mine, ref):
        assert np.array_equal(a, b)
    after_ref = torch.randint(0, 1 << 20, (16,))
    torch.manual_seed(1234)
    torch.rand(7)
    stream.commit(len(counts) * 400)
    after_mine = torch.randint(0, 1 << 20, (16,))
    assert torch.equal(after_ref, after_mine)


def test_rng_stream_crosses_state_regeneration():
    stream = TorchCpuRandintStream()
    torch.manual_seed(7)
    raw = stream.peek(2000)  # > 624 words: crosses three twists
    ref = torch.randint(0, 1000, (2000,)).numpy()
    assert np.array_equal(raw.astype(np.int64) % 1000, ref)


def test_fps_oracle_properties():
    rng = np.random.default_rng(0)
    pts = (rng.random((300, 3)) + 0.2).astype(np.float32)
    idx = ofps.furthest_point_sample(pts, 64)
    assert idx[0] == 0 and len(set(idx.tolist())) == 64
    # second pick is the farthest point from point 0
    d = ((pts - pts[0]) ** 2).sum(1)
    assert idx[1] == int(np.argmax(d))
    # min pairwise distance of the picks is non-increasing in pick order (FPS property)
    cover = []
    for j in range(1, 64):
        dj = ((pts[idx[:j]] - pts[idx[j]]) ** 2).sum(1).min()
        cover.append(dj)
    assert all(cover[i] >= cover[i + 1] - 1e-7 for i in range(len(cover) - 1))


def test_fps_oracle_degenerate():
    pts = np.full((5, 3), 0.5, dtype=np.float32)
    pts[3] = [0.9, 0.5, 0.5]
    idx = ofps.furthest_point_sample(pts, 6)
    assert idx[0] == 0 and idx[1] == 3
    # all remaining distances are 0: tie rule (B=4): bit-reversed (k mod 4) then k -> k=0 (rev 0)
    assert idx[2] == 0
    # points near the origin never become candidates -> index 0 forever
    z = np.zeros((4, 3), dtype=np.float32)
    assert ofps.furthest_point_sample(z, 3).tolist() == [0, 0, 0]
    out, idx = ofps.fps_batch([np.zeros((0, 3)), pts], 4)
    assert not out[0].any() and idx[1][1] == 3


def _fmaf():
    import ctypes, ctypes.util
    libm = ctypes.CDLL(ctypes.util.find_library("m") or "libm.so.6")
    libm.fmaf.restype = ctypes.c_float
    libm.fmaf.argtypes = [ctypes.c_float] * 3
    return lambda a, b, c: np.float32(libm.fmaf(float(a), float(b), float(c)))


def _fps_literal(pts, m):
    """Literal emulation of the published pointnet2_ops kernel (furthest_point_sampling_kernel): B threads stride over
    the points keeping the first maximum (strict >), then the shared-memory tree `__update(t, t + s)` for
    s = B/2 .. 1, where a tie keeps the lower position.  Sums of squares in nvcc's contraction of the source
    expression (mul, fma, fma — oracle/fps.c header).  Slow; small cases only."""
    fma = _fmaf()
    pts = np.ascontiguousarray(pts, dtype=np.float32)
    n = len(pts)
    B = 1 << min(int(np.floor(np.log2(n))), 9)
    temp = np.full(n, np.float32(1e10), dtype=np.float32)
    idx = np.zeros(m, dtype=np.int32)
    old = 0
    f = np.float32
    for j in range(1, m):
        best = np.full(B, f(-1), dtype=np.float32)
        besti = np.zeros(B, dtype=np.int64)
        x1, y1, z1 = pts[old]
        for tid in range(B):
            for k in range(tid, n, B):
                x2, y2, z2 = pts[k]
                mag = fma(z2, z2, fma(x2, x2, f(y2 * y2)))
                if mag <= f(1e-3):
                    continue
                dx, dy, dz = f(x2 - x1), f(y2 - y1), f(z2 - z1)
                d = fma(dz, dz, fma(dx, dx, f(dy * dy)))
                d2 = min(d, temp[k])
                temp[k] = d2
                if d2 > best[tid]:
                    best[tid], besti[tid] = d2, k
        s = B // 2
        while s >= 1:
            for t in range(s):
                if best[t + s] > best[t]:
                    best[t], besti[t] = best[t + s], besti[t + s]
            s //= 2
        old = int(besti[0])
        idx[j] = old
    return idx


def test_fps_tie_rule_equals_literal_block_reduction():
    """The oracle's tie key (bit-reversed k mod B, then k) == the outcome of the literal strided scan +
    tree reduction, including exact distance ties, identical points and never-candidate points."""
    rng = np.random.default_rng(4)
    for n, m in ((1, 3), (2, 4), (5, 8), (37, 20), (70, 40), (130, 24)):
        pts = (rng.random((n, 3)) * 0.5 + 0.1).astype(np.float32)
        cases = [pts]
        if n >= 5:
            dup = pts.copy(); dup[n // 2:] = dup[: n - n // 2]; cases.append(dup)          # exact ties
            same = pts.copy(); same[:] = same[0]; cases.append(same)                         # all identical
            zer = pts.copy(); zer[::3] = 0.0; cases.append(zer)                              # |p|^2 <= 1e-3
            grid = (np.stack(np.meshgrid(np.arange(4), np.arange(4), np.arange(4)), -1).reshape(-1, 3)[:n] * 0.25 + 0.25)
            cases.append(grid.astype(np.float32))                                             # lattice: many equal distances
        for c in cases:
            assert np.array_equal(ofps.furthest_point_sample(c, m), _fps_literal(c, m)), (n, m)


def test_fps_oracle_matches_golden_vectors(golden_dir):
    """tests/golden/fps_golden.npz (tools/make_golden_fps.py: indices from the literal emulation of the published
    kernel at N = 1, 37, 400, 5184 with duplicate / identical / never-candidate / lattice inputs) == oracle/fps.c."""
    g = np.load(os.path.join(golden_dir, "fps_golden.npz"))
    names = sorted(k[:-4] for k in g.files if k.endswith("_idx"))
    assert len(names) == 12
    for name in names:
        want = g[name + "_idx"]
        assert np.array_equal(ofps.furthest_point_sample(g[name + "_pts"], len(want)), want), name


def test_fps_oracle_degenerate_inputs():
    """Empty task, a single pick, fewer points than picks, and a cloud with no candidate at all (every |p|^2 <= 1e-3):
    the last one returns index 0 for every pick, like the published kernel whose reduction starts from (best=-1, i=0)."""
    assert ofps.furthest_point_sample(np.zeros((0, 3), np.float32), 5).tolist() == [0] * 5
    pts = (np.random.default_rng(0).random((9, 3)) * 0.5 + 0.1).astype(np.float32)
    assert ofps.furthest_point_sample(pts, 1).tolist() == [0]
    many = ofps.furthest_point_sample(pts, 30)
    assert np.array_equal(many, _fps_literal(pts, 30))
    assert sorted(set(many[:9].tolist())) == list(range(9))          # the first n picks visit every point once
    assert len(set(many[9:].tolist())) == 1                          # then the arg-max of all-zero distances repeats
    tiny = (pts * 0.01).astype(np.float32)                           # |p|^2 < 1e-3 everywhere
    assert ofps.furthest_point_sample(tiny, 6).tolist() == [0] * 6
    assert _fps_literal(tiny, 6).tolist() == [0] * 6
    out, idx = ofps.fps_batch([np.zeros((0, 3), np.float32), np.zeros((4, 3), np.float32), pts], 4)
    assert not out[0].any() and not out[1].any() and out[2].any() and idx[:2].sum() == 0   # `pts.any()` rule
