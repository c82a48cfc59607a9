#!/usr/bin/env python
"""Generate tests/golden/traj_golden.npz by running the REAL DataLoggerSim (algo/ppo/experience.py:352-490).

Build-container only (needs /root/reference).  The module imports gym / deepdish at the top (unused by
DataLoggerSim); they are stubbed.  The logger is built with save_trajectory=False (no worker processes),
then `update(save_trajectory=True)` runs with `_save_batch_trajectories` replaced by a list append, so the
reference's own lines build every saved trajectory; the final exit() of the collection run is not reached
(total_trajectories is set out of reach).
"""
import importlib.util
import os
import sys
from unittest import mock

import numpy as np
import torch

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
REF = "/root/reference"
for s in ("gym", "deepdish"):
    sys.modules.setdefault(s, mock.MagicMock())
spec = importlib.util.spec_from_file_location("ref_experience", f"{REF}/algo/ppo/experience.py")
ref = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref)


def script(seed=3, N=6, T=9, steps=23):
    """Scripted episode: per step the logged rows and the done flags (some steps log None / int32 rows)."""
    rng = np.random.default_rng(seed)
    shapes = {"tactile": (3, 16), "seg": (24,), "action": 6, "latent": 8}
    rows = []
    age = np.zeros(N, dtype=np.int64)
    for t in range(steps):
        r = {"tactile": rng.random((N, 3, 16), dtype=np.float32),
             "seg": rng.integers(0, 4, (N, 24)).astype(np.int32),
             "action": rng.standard_normal((N, 6)).astype(np.float32),
             "latent_none": bool(t % 5 == 2)}
        r["latent"] = rng.standard_normal((N, 8)).astype(np.float32)
        age += 1
        done = (rng.random(N) < 0.18) | (age >= T)          # an env never outruns its episode buffer
        r["done_none"] = bool(t == 4) and not (age >= T).any()
        if r["done_none"]:
            done[:] = False
        age[done] = 0
        r["done"] = done
        rows.append(r)
    return shapes, rows, N, T


def main():
    shapes, rows, N, T = script()
    kw = {k + "_shape": (torch.Size(v) if isinstance(v, tuple) else v) for k, v in shapes.items()}
    lg = ref.DataLoggerSim(N, T, "cpu", "/tmp/igi_traj_golden", 10 ** 9, False, **kw)
    lg.total_trajectories, lg.pbar = 10 ** 9, mock.MagicMock()
    saved = []
    lg._save_batch_trajectories = lambda d: saved.append({k: np.asarray(v) for k, v in d.items()})
    for r in rows:
        lg.update(save_trajectory=True, tactile=torch.from_numpy(r["tactile"]), seg=torch.from_numpy(r["seg"]),
                  action=torch.from_numpy(r["action"]), latent=None if r["latent_none"] else torch.from_numpy(r["latent"]),
                  done=None if r["done_none"] else torch.from_numpy(r["done"]))
    out = {"N": N, "T": T, "n_steps": len(rows), "n_saved": len(saved)}
    for i, r in enumerate(rows):
        for k in ("tactile", "seg", "action", "latent", "done"):
            out[f"in{i}_{k}"] = r[k]
        out[f"in{i}_flags"] = np.array([r["latent_none"], r["done_none"]])
    for k, v in lg.log_data.items():
        out["final_" + k] = v.numpy()
    out["final_done"] = lg.done.numpy()
    out["final_counter"] = lg.env_step_counter.numpy()
    for i, d in enumerate(saved):
        for k, v in d.items():
            out[f"saved{i}_{k}"] = v
    path = os.path.join(ROOT, "tests", "golden", "traj_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, "saved trajectories:", len(saved), os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
