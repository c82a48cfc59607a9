"""CPU oracle for the on-disk trajectory logger — TEST INFRASTRUCTURE ONLY.

Restates DataLoggerSim (algo/ppo/experience.py:352-490) with CPU torch in the reference's op order:
per-env episode buffers `log_data[key]` of shape (num_envs, episode_length, *shape) f32, a (num_envs,
episode_length) bool `done` buffer and a per-env step counter; `update()` scatters the step's rows at the
counters (:426-444), and for every env whose `done` is set hands its whole trajectory
{key: (T, *shape) f32, 'done': (T,) bool} to the writer (:445-455), then zeroes that env's buffers (:417-420).
The writer stores one `np.savez_compressed` file per trajectory (:472-490).

Parity pin: tests/golden/traj_golden.npz holds the buffers and the saved trajectories produced by the REAL
DataLoggerSim run on a scripted episode (tools/make_golden_traj.py); tests/test_oracle_traj.py compares
this restatement against them bit-for-bit.
"""
import torch


class DataLoggerSim:
    def __init__(self, num_envs, episode_length, **shapes):
        """shapes: key -> int | torch.Size | tuple (the reference's `<key>_shape` kwargs without the suffix)."""
        self.num_envs, self.T = num_envs, episode_length
        self.data_shapes = dict(shapes)
        self.saved = []          # what the reference puts on its writer queues, in order
        self._init_buffers()

    def _init_buffers(self):     # :400-415
        self.log_data = {}
        for key, shape in self.data_shapes.items():
            if shape is None:
                continue
            tail = tuple(shape) if isinstance(shape, (tuple, list, torch.Size)) else (shape,)
            self.log_data[key] = torch.zeros((self.num_envs, self.T) + tail, dtype=torch.float32)
        self.done = torch.zeros((self.num_envs, self.T), dtype=torch.bool)
        self.env_step_counter = torch.zeros((self.num_envs, 1), dtype=torch.long)
        self.env_ids = torch.arange(self.num_envs, dtype=torch.long).unsqueeze(-1)

    def update(self, save_trajectory=True, **kwargs):   # :426-468
        for key, value in kwargs.items():
            if key == "done":
                continue
            if value is None:
                value = torch.zeros((self.num_envs, self.data_shapes[key]), dtype=torch.float32)
            self.log_data[key][self.env_ids, self.env_step_counter, ...] = value.clone().unsqueeze(1).to(torch.float32)
        done = kwargs.get("done", None)
        if done is None:
            done = torch.zeros(self.num_envs, dtype=torch.bool)
        done = done.clone().to(torch.bool)
        self.done[self.env_ids, self.env_step_counter, ...] = done.unsqueeze(1)
        self.env_step_counter += 1
        ids = done.to(torch.long).nonzero()
        if len(ids) > 0:
            ids = ids.squeeze(1)
            if save_trajectory:
                for e in ids:
                    item = {k: self.log_data[k][e, ...].clone() for k in self.log_data}
                    item["done"] = self.done[e, ...].clone()
                    self.saved.append(item)
            for buf in self.log_data.values():          # _reset_buffers :417-420
                buf[ids, ...] = 0.0
            self.done[ids, ...] = False
            self.env_step_counter[ids, ...] = 0
