"""Oracle of the trajectory logger against the golden run of the REAL DataLoggerSim (tools/make_golden_traj.py)."""
import os

import numpy as np
import torch

from oracle import experience as oexp

SHAPES = {"tactile": (3, 16), "seg": (24,), "action": 6, "latent": 8}


def replay(g, logger, conv=lambda a: torch.from_numpy(a)):
    for i in range(int(g["n_steps"])):
        latent_none, done_none = g[f"in{i}_flags"]
        logger.update(save_trajectory=True, tactile=conv(g[f"in{i}_tactile"]), seg=conv(g[f"in{i}_seg"]),
                      action=conv(g[f"in{i}_action"]), latent=None if latent_none else conv(g[f"in{i}_latent"]),
                      done=None if done_none else conv(g[f"in{i}_done"]))


def test_oracle_matches_reference_logger(golden_dir):
    g = np.load(os.path.join(golden_dir, "traj_golden.npz"))
    lg = oexp.DataLoggerSim(int(g["N"]), int(g["T"]), **SHAPES)
    replay(g, lg)
    for k in SHAPES:
        np.testing.assert_array_equal(lg.log_data[k].numpy(), g["final_" + k])
    np.testing.assert_array_equal(lg.done.numpy(), g["final_done"])
    np.testing.assert_array_equal(lg.env_step_counter.numpy(), g["final_counter"])
    assert len(lg.saved) == int(g["n_saved"]) > 10
    for i, item in enumerate(lg.saved):
        assert set(item) == set(SHAPES) | {"done"}
        for k, v in item.items():
            np.testing.assert_array_equal(v.numpy(), g[f"saved{i}_{k}"])
