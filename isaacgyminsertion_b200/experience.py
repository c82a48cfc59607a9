"""On-disk trajectory format of the student's offline data, same names and signatures as
the reference's:

  DataLoggerSim    algo/ppo/experience.py:352-490   per-env episode buffers + one .npz per finished trajectory
  SimLogger        algo/ppo/experience.py:634-746   which env tensors are logged under which key

The episode buffers live on the device and are filled straight from the observation buffers this library
produces (`tactile_imgs`, `image_buf`, `seg_buf`, ... rows may be strided views of the packed observation
buffer): csrc/traj.cu scatters every key's rows at the envs' own step counters, marks `done`, compacts the ids
of the envs that finished and gathers their whole trajectories into ONE staging block per key, so a step
with k finished envs costs one device->host copy per key instead of the reference's k * keys `.clone().cpu()`
calls.  Files are written by a small thread pool (zlib releases the GIL) in the reference's layout:
`<dir>/<writer idx>/<timestamp>_<n>.npz` holding {key: (T, *shape) float32, 'done': (T,) bool}; the `_<n>`
suffix keeps two trajectories finished within the same second apart (the reference's names collide there).

CPU tensors raise: there is no CPU fallback.
"""
import ctypes as _c
import os
import threading
from concurrent.futures import ThreadPoolExecutor
from datetime import datetime

import numpy as np
import torch

from . import _lib

ROT_MAT_SIZE = 9


def _stream(dev):
    return _lib.stream_ptr(dev)


class DataLoggerSim:
    def __init__(self, num_envs, episode_length, device, dir_path, total_trajectories, save_trajectory, **kwargs):
        self.lib = _lib.load()
        self.num_envs = num_envs
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("DataLoggerSim: expected a CUDA device (no CPU fallback)")
        self.transitions_per_env = episode_length
        self.data_shapes = {}
        os.makedirs(dir_path, exist_ok=True)
        self.dir = dir_path
        for key, value in kwargs.items():
            if key.endswith("_shape"):
                self.data_shapes[key.replace("_shape", "")] = value
        self.trajectory_ctr = 0
        self.total_trajectories = total_trajectories if save_trajectory else None
        self.finished = False            # the reference calls exit() once total_trajectories are on disk
        self.num_workers = 8
        self._pool = ThreadPoolExecutor(max_workers=self.num_workers) if save_trajectory else None
        self._pending = []
        self._file_ctr = 0
        self._lock = threading.Lock()
        self._init_buffers()

    # ------------------------------------------------------------------ buffers
    def _init_buffers(self):
        N, T, dev = self.num_envs, self.transitions_per_env, self.device
        self.log_data = {}
        for key, shape in self.data_shapes.items():
            if shape is not None:
                tail = tuple(shape) if isinstance(shape, (tuple, list, torch.Size)) else (shape,)
                self.log_data[key] = torch.zeros((N, T) + tail, dtype=torch.float32, device=dev)
        self._pitch = (T + 15) // 16 * 16           # done rows padded to 16 bytes for the gather
        self._done_u8 = torch.zeros((N, self._pitch), dtype=torch.uint8, device=dev)
        self.env_step_counter = torch.zeros((N, 1), dtype=torch.long, device=dev)
        self.env_ids = torch.arange(N, dtype=torch.long, device=dev).unsqueeze(-1)
        self._done_ids = torch.zeros(N, dtype=torch.int32, device=dev)
        self._n_done = torch.zeros(1, dtype=torch.int32, device=dev)
        self._overflow = torch.zeros(1, dtype=torch.int32, device=dev)
        self._n_done_host = torch.zeros(2, dtype=torch.int32).pin_memory()
        self._all_ids = torch.arange(N, dtype=torch.int32, device=dev)
        self._n_all = torch.full((1,), N, dtype=torch.int32, device=dev)

    @property
    def done(self):
        """(num_envs, episode_length) bool view of the done log, as the reference's attribute."""
        return self._done_u8[:, :self.transitions_per_env].bool()

    def _gather(self, buf, row_bytes, ids, n_ids, max_ids, staging, zero_after):
        rc = self.lib.igi_traj_gather(_c.c_void_p(buf.data_ptr()), _lib.dptr(ids, torch.int32), _lib.dptr(n_ids, torch.int32),
                                      _c.c_int(max_ids), _c.c_longlong(row_bytes),
                                      _c.c_void_p(staging.data_ptr() if staging is not None else 0),
                                      _c.c_int(1 if zero_after else 0), _stream(self.device))
        _lib.check(rc, "igi_traj_gather")

    def _reset_buffers(self, env_ids):
        """Zero the listed envs' episode buffers, done rows and step counters (experience.py:417-420)."""
        ids = torch.as_tensor(env_ids, device=self.device).reshape(-1).to(torch.int32).contiguous()
        n = torch.full((1,), ids.numel(), dtype=torch.int32, device=self.device)
        self._reset_listed(ids, n, ids.numel())

    def _reset_listed(self, ids, n_ids, max_ids):
        for buf in self.log_data.values():
            self._gather(buf, buf[0].numel() * 4, ids, n_ids, max_ids, None, True)
        self._gather(self._done_u8, self._pitch, ids, n_ids, max_ids, None, True)
        rc = self.lib.igi_traj_reset_counters(_lib.dptr(self.env_step_counter, torch.int64), _lib.dptr(ids, torch.int32),
                                              _lib.dptr(n_ids, torch.int32), _c.c_int(max_ids), _stream(self.device))
        _lib.check(rc, "igi_traj_reset_counters")

    # ------------------------------------------------------------------ one env step
    @torch.no_grad()
    def update(self, save_trajectory=True, **kwargs):
        N, T, dev = self.num_envs, self.transitions_per_env, self.device
        keep = []                                    # converted temporaries stay referenced until enqueued
        for key, value in kwargs.items():
            if key == "done":
                continue
            log = self.log_data[key]
            L = log[0, 0].numel()
            x = None
            if value is not None:
                if not value.is_cuda:
                    raise RuntimeError(f"DataLoggerSim.update: '{key}' must be a CUDA tensor (no CPU fallback)")
                x = value.reshape(N, -1) if value.dim() != 2 else value
                if x.shape[1] != L:
                    raise RuntimeError(f"DataLoggerSim.update: '{key}' has {x.shape[1]} values per env, expected {L}")
                if x.dtype not in (torch.float32, torch.int32) or x.stride(1) != 1:
                    x = x.to(torch.float32).contiguous()
                keep.append(x)
            rc = self.lib.igi_traj_append(_lib.dptr(log, torch.float32), _c.c_void_p(x.data_ptr() if x is not None else 0),
                                          _c.c_int(1 if x is not None and x.dtype == torch.int32 else 0),
                                          _c.c_int64(x.stride(0) if x is not None else L),
                                          _lib.dptr(self.env_step_counter, torch.int64), _c.c_int(N), _c.c_int(T),
                                          _c.c_longlong(L), _lib.dptr(self._overflow, torch.int32), _stream(dev))
            _lib.check(rc, "igi_traj_append")
        done = kwargs.get("done", None)
        d8 = None
        if done is not None:
            d8 = done.to(device=dev, dtype=torch.bool).to(torch.uint8).contiguous()
            keep.append(d8)
        rc = self.lib.igi_traj_step(_lib.dptr(self._done_u8, torch.uint8), _c.c_int(self._pitch), _lib.dptr(d8, torch.uint8),
                                    _lib.dptr(self.env_step_counter, torch.int64), _c.c_int(N), _c.c_int(T),
                                    _lib.dptr(self._done_ids, torch.int32), _lib.dptr(self._n_done, torch.int32),
                                    _lib.dptr(self._overflow, torch.int32), _stream(dev))
        _lib.check(rc, "igi_traj_step")
        # the one host sync of the step (the reference syncs in `done.nonzero()`): how many envs finished
        self._n_done_host[0:1].copy_(self._n_done, non_blocking=True)
        self._n_done_host[1:2].copy_(self._overflow, non_blocking=True)
        torch.cuda.current_stream(dev).synchronize()
        k, ovf = int(self._n_done_host[0]), int(self._n_done_host[1])
        if ovf:
            self._overflow.zero_()
            raise IndexError(f"DataLoggerSim.update: an env logged more than episode_length={T} steps without done")
        if k == 0:
            return
        if save_trajectory:
            staged = {}
            for key, buf in self.log_data.items():
                st = torch.empty((k,) + tuple(buf.shape[1:]), dtype=torch.float32, device=dev)
                self._gather(buf, buf[0].numel() * 4, self._done_ids, self._n_done, k, st, True)
                staged[key] = st
            st = torch.empty((k, self._pitch), dtype=torch.uint8, device=dev)
            self._gather(self._done_u8, self._pitch, self._done_ids, self._n_done, k, st, True)
            staged["done"] = st
            host = {key: v.cpu() for key, v in staged.items()}     # one D2H copy per key
            rc = self.lib.igi_traj_reset_counters(_lib.dptr(self.env_step_counter, torch.int64),
                                                  _lib.dptr(self._done_ids, torch.int32), _lib.dptr(self._n_done, torch.int32),
                                                  _c.c_int(k), _stream(dev))
            _lib.check(rc, "igi_traj_reset_counters")
            self.trajectory_ctr += k
            for j in range(k):
                item = {key: host[key][j].numpy() for key in self.log_data}
                item["done"] = host["done"][j, :T].numpy().astype(bool)
                self._save_batch_trajectories(item)
            if self.total_trajectories is not None and self.trajectory_ctr >= self.total_trajectories:
                self.close()
                self.finished = True
                print("Data collection finished!")
        else:
            self._reset_listed(self._done_ids, self._n_done, k)

    # ------------------------------------------------------------------ writer
    def _save_batch_trajectories(self, data):
        with self._lock:
            n = self._file_ctr
            self._file_ctr += 1
        q_id = n % self.num_workers
        if self._pool is None:
            self._pool = ThreadPoolExecutor(max_workers=self.num_workers)
        self._pending.append(self._pool.submit(self._write, data, q_id, n))

    def _write(self, item, q_idx, n):
        data_path = os.path.join(self.dir, f"{q_idx}")
        os.makedirs(data_path, exist_ok=True)
        name = f'{datetime.now().strftime("%Y-%m-%d_%H-%M-%S")}_{n:06d}.npz'
        np.savez_compressed(os.path.join(data_path, name), **item)
        return os.path.join(data_path, name)

    def close(self):
        """Wait for every queued trajectory to be on disk; returns the file paths written since the last call."""
        paths = [f.result() for f in self._pending]
        self._pending = []
        return paths

    def get_data(self):
        return self.log_data

    def reset(self):
        self._reset_listed(self._all_ids, self._n_all, self.num_envs)


class SimLogger:
    """experience.py:634-746 for an env-like object: the observation keys this library owns (tactile, img, seg)
    are logged straight from the task's device buffers; every other key is whatever tensor the env exposes
    under the reference's attribute name (keys whose attribute is missing are left out)."""

    # log key -> env attribute (experience.py:698-730); poses in the robot base frame are passed by the caller
    ATTRS = {"arm_joints": "arm_dof_pos", "target": "targets", "rigid_physics_params": "rigid_physics_params",
             "plug_hand_pos": "plug_hand_pos", "plug_hand_quat": "plug_hand_quat", "plug_pos_error": "plug_pos_error",
             "plug_quat_error": "plug_quat_error", "finger_normalized_forces": "finger_normalized_forces",
             "plug_heights": "plug_heights", "obs_hist": "obs_buf", "obs_hist_stud": "obs_student_buf",
             "priv_obs": "states_buf", "hand_joints": "hand_joints", "img": "image_buf", "seg": "seg_buf",
             "tactile": "tactile_imgs"}

    def __init__(self, env, log_folder, total_trajectories=1, collect_data=True, action_dim=None, latent_dim=None,
                 extra_shapes=None):
        self.env = env
        items = {}
        for key, attr in self.ATTRS.items():
            t = getattr(env, attr, None)
            if t is not None:
                items[key + "_shape"] = t.shape[1:] if t.dim() > 2 else t.shape[-1]
        if action_dim:
            items["action_shape"] = action_dim
        if latent_dim:
            items["latent_shape"] = latent_dim
        items.update(extra_shapes or {})
        self._keys = [k[:-6] for k in items]
        self.data_logger_init = lambda x: DataLoggerSim(env.num_envs, env.max_episode_length, env.device, log_folder,
                                                        total_trajectories, save_trajectory=collect_data, **items)
        self.data_logger = None

    def log_trajectory_data(self, action, latent, done, save_trajectory=True, **extra):
        if self.data_logger is None:
            self.data_logger = self.data_logger_init(None)
        log_data = {}
        for key in self._keys:
            if key in self.ATTRS and getattr(self.env, self.ATTRS[key], None) is not None:
                log_data[key] = getattr(self.env, self.ATTRS[key])
        if "action" in self._keys:
            log_data["action"] = action
        if "latent" in self._keys and latent is not None:
            log_data["latent"] = latent
        log_data.update(extra)
        log_data["done"] = done
        self.data_logger.update(save_trajectory=save_trajectory, **log_data)
