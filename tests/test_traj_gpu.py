"""GPU parity: the trajectory logger (csrc/traj.cu behind DataLoggerSim) against the golden run of the REAL
reference DataLoggerSim: episode buffers, done log, counters and every saved trajectory bit-for-bit, plus the
on-disk round trip of the .npz files."""
import glob
import os

import numpy as np
import pytest
import torch

from test_oracle_traj import SHAPES, replay

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _logger(g, tmp_path, save=True):
    from isaacgyminsertion_b200.experience import DataLoggerSim
    kw = {k + "_shape": (torch.Size(v) if isinstance(v, tuple) else v) for k, v in SHAPES.items()}
    return DataLoggerSim(int(g["N"]), int(g["T"]), DEV, str(tmp_path), 10 ** 9, save, **kw)


def test_logger_matches_reference(built_lib, golden_dir, tmp_path):
    g = np.load(os.path.join(golden_dir, "traj_golden.npz"))
    lg = _logger(g, tmp_path)
    saved = []
    lg._save_batch_trajectories = lambda d: saved.append({k: np.array(v) for k, v in d.items()})
    replay(g, lg, conv=lambda a: torch.from_numpy(a).to(DEV))
    for k in SHAPES:
        np.testing.assert_array_equal(lg.log_data[k].cpu().numpy(), g["final_" + k])
    np.testing.assert_array_equal(lg.done.cpu().numpy(), g["final_done"])
    np.testing.assert_array_equal(lg.env_step_counter.cpu().numpy(), g["final_counter"])
    assert len(saved) == int(g["n_saved"])
    for i, item in enumerate(saved):
        assert set(item) == set(SHAPES) | {"done"}
        for k, v in item.items():
            assert v.dtype == g[f"saved{i}_{k}"].dtype
            np.testing.assert_array_equal(v, g[f"saved{i}_{k}"])


def test_files_round_trip_and_no_save_mode(built_lib, golden_dir, tmp_path):
    g = np.load(os.path.join(golden_dir, "traj_golden.npz"))
    lg = _logger(g, tmp_path / "a")
    replay(g, lg, conv=lambda a: torch.from_numpy(a).to(DEV))
    paths = lg.close()
    assert len(paths) == int(g["n_saved"]) == len(glob.glob(str(tmp_path / "a" / "*" / "*.npz")))
    order = sorted(paths, key=lambda p: int(os.path.basename(p).rsplit("_", 1)[1][:-4]))
    for i, p in enumerate(order):
        with np.load(p) as z:
            assert set(z.files) == set(SHAPES) | {"done"}
            for k in z.files:
                np.testing.assert_array_equal(z[k], g[f"saved{i}_{k}"])
    # save_trajectory=False: same buffers, nothing written
    lg2 = _logger(g, tmp_path / "b", save=False)
    for i in range(int(g["n_steps"])):
        ln, dn = g[f"in{i}_flags"]
        t = lambda a: torch.from_numpy(a).to(DEV)
        lg2.update(save_trajectory=False, tactile=t(g[f"in{i}_tactile"]), seg=t(g[f"in{i}_seg"]), action=t(g[f"in{i}_action"]),
                   latent=None if ln else t(g[f"in{i}_latent"]), done=None if dn else t(g[f"in{i}_done"]))
    for k in SHAPES:
        np.testing.assert_array_equal(lg2.log_data[k].cpu().numpy(), g["final_" + k])
    assert not glob.glob(str(tmp_path / "b" / "*" / "*.npz"))
    lg2.reset()
    assert float(lg2.log_data["tactile"].abs().max()) == 0 and int(lg2.env_step_counter.max()) == 0


def test_logger_takes_strided_task_rows_and_flags_overrun(built_lib, tmp_path):
    """tactile rows are views into the packed [tactile | pcl] observation buffer; an env that outruns its
    episode buffer raises like the reference's index error."""
    from isaacgyminsertion_b200.experience import DataLoggerSim
    N, T = 5, 3
    packed = torch.rand((N, 3 * 2048 + 2400), device=DEV)
    tact = packed[:, :3 * 2048].view(N, 3, 2048)
    lg = DataLoggerSim(N, T, DEV, str(tmp_path), 10, True, tactile_shape=tact.shape[1:], seg_shape=5184)
    seg = torch.randint(0, 4, (N, 5184), dtype=torch.int32, device=DEV)
    done = torch.zeros(N, dtype=torch.bool, device=DEV)
    lg.update(tactile=tact, seg=seg, done=done)
    assert torch.equal(lg.log_data["tactile"][:, 0], tact) and torch.equal(lg.log_data["seg"][:, 0], seg.float())
    done[2] = True
    lg.update(tactile=tact * 2, seg=seg, done=done)
    (p,) = lg.close()
    with np.load(p) as z:
        np.testing.assert_array_equal(z["tactile"][1], (tact[2] * 2).cpu().numpy())
        assert z["done"].tolist() == [False, True, False] and z["seg"].shape == (T, 5184)
    assert int(lg.env_step_counter[2]) == 0 and int(lg.env_step_counter[0]) == 2
    lg.update(tactile=tact, seg=seg, done=None)
    with pytest.raises(IndexError):
        lg.update(tactile=tact, seg=seg, done=None)
