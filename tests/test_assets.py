"""CPU: mesh loading semantics of the tactile path (reference allsight_render.py:101-109:
`trimesh.load`, in-place x,y scale, smooth vertex normals recomputed from the faces).

trimesh is absent here, so `assets.load_obj / merge_vertices / angle_weighted_normals` restate its
published behaviour (DESIGN.md "mesh loading"); these tests pin that restatement to hand-computed answers
on a small OBJ and to the packed arrays the renderer ships with."""
import os

import numpy as np
import pytest

from isaacgyminsertion_b200 import assets

REF_MESH = "/root/reference/assets/factory/mesh/factory_insertion"

# three faces around the origin on the coordinate planes (unequal corner angles at (2,0,0)), a face given with
# NEGATIVE (relative) indices, a v/vt/vn corner format, a degenerate face, a crease duplicate (same position, other
# normal) and a smooth duplicate (same position, same normal)
SMALL_OBJ = """# hand-made
v 0 0 0
v 2 0 0
v 0 1 0
v 0 0 3
v 2 0 0
v 2 0 0
v 5 5 5
vn 0 0 1
vn 0 0 1
vn 0 0 1
vn 0 1 0
vn 0 1 0
vn 0 0 1
vn 1 0 0
vt 0.5 0.5
f 1//1 2//2 3//3
f 1//1 4//4 5//5
f -7//-7 -5//-5 -4//-4
f 1/1/1 6/1/6 3/1/3
f 2//2 2//2 3//3
"""


@pytest.fixture()
def small_obj(tmp_path):
    p = tmp_path / "small.obj"
    p.write_text(SMALL_OBJ)
    return str(p)


def test_load_obj_indices_and_normals(small_obj):
    V, F = assets.load_obj(small_obj)
    assert V.shape == (7, 3) and F.shape == (5, 3)
    assert F.tolist() == [[0, 1, 2], [0, 3, 4], [0, 2, 3], [0, 5, 2], [1, 1, 2]]      # negative = relative to the end
    V2, F2, VN = assets.load_obj(small_obj, with_normals=True)
    assert np.array_equal(V2, V) and np.array_equal(F2, F) and VN.shape == (7, 3)       # vn index == v index everywhere
    assert VN[3].tolist() == [0, 1, 0] and VN[5].tolist() == [0, 0, 1]


def test_merge_keeps_creases_and_merges_smooth_duplicates(small_obj):
    V, F, VN = assets.load_obj(small_obj, with_normals=True)
    Vm, Fm = assets.merge_vertices(V, F, VN)
    # vertex 5 = (2,0,0) with normal +z duplicates vertex 1 -> merged; vertex 4 = (2,0,0) with normal +y is a
    # crease copy -> kept; vertex 6 is unreferenced -> dropped
    assert len(Vm) == 5
    assert Fm.tolist() == [[0, 1, 2], [0, 3, 4], [0, 2, 3], [0, 1, 2], [1, 1, 2]]
    assert np.array_equal(Vm[Fm], V[F]), "merging must not move any triangle"
    # without file normals: position only (trimesh merges everything that coincides)
    Vp, Fp = assets.merge_vertices(V, F)
    assert len(Vp) == 4 and np.array_equal(Vp[Fp], V[F])
    # normals that agree to 2 decimals merge (digits_norm = 2), a third-decimal difference does not split
    VN2 = VN.copy()
    VN2[5] = [0.0, 0.004, 1.0]
    assert len(assets.merge_vertices(V, F, VN2)[0]) == 5
    VN2[5] = [0.0, 0.02, 1.0]
    assert len(assets.merge_vertices(V, F, VN2)[0]) == 6


def test_angle_weighted_normals_by_hand():
    V = np.array([[0, 0, 0], [2, 0, 0], [0, 1, 0], [0, 0, 3.0]])
    F = np.array([[0, 1, 2],      # xy plane, normal +z; corner angle at (2,0,0): atan(1/2)
                  [0, 3, 1],      # zx plane, normal +y; corner angle at (2,0,0): atan(3/2)
                  [0, 2, 3],      # yz plane, normal +x
                  [1, 1, 2]])     # degenerate: contributes nothing
    vn = assets.angle_weighted_normals(V, F)
    assert np.allclose(vn[0], np.ones(3) / np.sqrt(3))                      # three right angles
    a, b = np.arctan2(1, 2), np.arctan2(3, 2)
    want = np.array([0, b, a]) / np.hypot(a, b)
    assert np.allclose(vn[1], want, atol=1e-12)
    assert not np.allclose(vn[1], np.array([0, 3, 1]) / np.sqrt(10), atol=1e-3), "area weighting would give this"
    assert not np.allclose(vn[1], np.array([0, 1, 1]) / np.sqrt(2), atol=1e-3), "uniform weighting would give this"
    assert np.allclose(np.linalg.norm(vn, axis=1), 1.0)


def test_load_peg_scales_xy_then_recomputes_normals(small_obj):
    s = 1.1
    V, VN, F = assets.load_peg_from_obj(small_obj, s)
    assert V.dtype == np.float32 and VN.dtype == np.float32 and F.dtype == np.int32
    assert len(V) == 5 and np.allclose(V[1], [2 * s, 0, 0]) and np.allclose(V[3], [0, 0, 3])     # z is not scaled
    # vertex 1 = (2s,0,0) now belongs to the xy faces only (its crease copy, vertex 4, took the zx face)
    assert np.allclose(VN[1], [0, 0, 1], atol=1e-6) and np.allclose(VN[4], [0, 1, 0], atol=1e-6)
    # the origin: right angles in every face, the xy plane counted twice (faces 0 and 3 coincide after the merge)
    assert np.allclose(VN[0], np.array([1, 1, 2]) / np.sqrt(6), atol=1e-6)
    # a corner whose angle changes with the scale: vertex 2 = (0,s,0): xy face angle atan(2s/s), yz face angle atan(3/s)
    a_xy, a_yz = np.arctan2(2 * s, s), np.arctan2(3, s)
    want = np.array([a_yz, 0, 2 * a_xy])      # the xy plane is counted twice (faces 0 and 3 coincide after the merge)
    assert np.allclose(VN[2], want / np.linalg.norm(want), atol=1e-6)
    unscaled = np.array([np.arctan2(3, 1), 0, 2 * np.arctan2(2, 1)])
    assert not np.allclose(VN[2], unscaled / np.linalg.norm(unscaled), atol=1e-4), "normals must follow the scaled mesh"


def test_packed_pegs_are_consistent():
    a = assets.load_packed()
    names = [str(n) for n in a["peg_names"]]
    assert len(names) == 7
    for i in range(7):
        V, VN, F = a[f"peg_{i}_v"], a[f"peg_{i}_vn"], a[f"peg_{i}_f"]
        assert F.min() == 0 and F.max() == len(V) - 1, "every vertex referenced (unreferenced ones are dropped)"
        assert np.allclose(np.linalg.norm(VN, axis=1), 1.0, atol=1e-5)
        assert np.allclose(assets.angle_weighted_normals(V.astype(np.float64), F), VN, atol=2e-3)
        assert abs(V[:, 2].max() - 0.0762) < 1e-4 and abs(V[:, 2].min()) < 2e-6      # 76.2 mm pegs
    # an independent geometric check: on the wall of the round peg (uniform x,y scale keeps it a cylinder) the smooth
    # normal is radial
    i = names.index("yellow_round_peg_2in")
    V, VN = a[f"peg_{i}_v"].astype(np.float64), a[f"peg_{i}_vn"].astype(np.float64)
    r = np.hypot(V[:, 0], V[:, 1])
    wall = (V[:, 2] > 0.02) & (V[:, 2] < 0.06) & (r > 0.98 * r.max())
    radial = np.stack([V[wall, 0] / r[wall], V[wall, 1] / r[wall], np.zeros(wall.sum())], axis=1)
    cosang = np.einsum("ij,ij->i", radial, VN[wall])
    assert wall.sum() > 100 and np.degrees(np.arccos(np.clip(cosang, -1, 1))).max() < 3.0


@pytest.mark.skipif(not os.path.isdir(REF_MESH), reason="reference meshes are only mounted in the build container")
@pytest.mark.parametrize("idx,fname", [(0, "hexagon_peg.obj"), (3, "small_triangle_peg.obj"),
                                       (5, "yellow_round_peg_2in.obj"),
                                       (6, "factory_square_peg_32mm_loose_subdiv_3x.obj")])
def test_load_peg_from_obj_reproduces_packed_arrays(idx, fname):
    a = assets.load_packed()
    V, VN, F = assets.load_peg_from_obj(os.path.join(REF_MESH, fname), float(a["peg_scales"][idx]))
    assert np.array_equal(V, a[f"peg_{idx}_v"]) and np.array_equal(F, a[f"peg_{idx}_f"])
    assert np.array_equal(VN, a[f"peg_{idx}_vn"])
    # crease copies survive the merge (trimesh merge_norm=False): more vertices than distinct positions, except
    # for the file without `vn` records, which merges on position alone
    n_pos = len(np.unique(np.round(V.astype(np.float64), 8), axis=0))
    assert (len(V) > n_pos) == (fname != "yellow_round_peg_2in.obj")
