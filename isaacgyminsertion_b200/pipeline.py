"""Host-buffer front end of the observation path.

`HostObsPipeline` is the call a user with HOST data makes: every `step()` takes the step's
inputs as (pinned) host tensors — fingertip / plug poses and the external camera's depth +
segmentation images, i.e. what the reference task reads from IsaacGym each step
(factory_task_insertion.py:481-484, 918-919) — enqueues upload, kernels and download, and
returns a `PendingObs` whose `.wait()` yields the packed observation rows
`[tactile 3*2048 | pcl 2400]` of that step in pinned host memory.

Copies and kernels run on three streams (upload, compute, download) over a ring of SLOTS device
input / device output / host output buffers, so in steady state the H2D copy of step i+1 and the
D2H copy of step i-1 overlap the kernels of step i.  A result stays valid until SLOTS more steps
have been submitted; host input buffers must not be modified until the step's `.wait()` returned
(or `uploaded.synchronize()`).
"""
import os

import torch


SLOTS = int(os.environ.get("IGI_PIPE_SLOTS", "3"))   # ring depth (>= 2); 3 = one step uploading, one computing, one downloading


class PendingObs:
    def __init__(self, tensor, event, uploaded):
        self.tensor, self.event, self.uploaded = tensor, event, uploaded

    def wait(self):
        self.event.synchronize()
        return self.tensor


class HostObsPipeline:
    def __init__(self, task, sampler_socket_every_step=False):
        self.task = task
        dev = task.device
        self.dev = dev
        self.s_up = torch.cuda.Stream(device=dev)
        self.s_down = torch.cuda.Stream(device=dev)
        N = task.num_envs
        H, W = task.res[1], task.res[0]
        f32, i32 = torch.float32, torch.int32

        def mk():
            return dict(fpos=torch.empty((N, 3, 3), dtype=f32, device=dev), fquat=torch.empty((N, 3, 4), dtype=f32, device=dev),
                        ppos=torch.empty((N, 3), dtype=f32, device=dev), pquat=torch.empty((N, 4), dtype=f32, device=dev),
                        depth=torch.empty((N, H, W), dtype=f32, device=dev), seg=torch.empty((N, H, W), dtype=i32, device=dev))
        self.d_in = [mk() for _ in range(SLOTS)]
        # segmentation ids fit a byte (table 0, arm 1, plug 2, socket 3, factory_env_insertion.py:814-848): a host that
        # keeps them as uint8 uploads a quarter of the bytes; they are widened on the device into the int32 image the
        # task reads (the reference's camera tensor type)
        self.d_seg8 = None
        self.d_out = [torch.empty_like(task.obs_packed) for _ in range(SLOTS)]
        self.h_out = [torch.empty(task.obs_packed.shape, dtype=f32).pin_memory() for _ in range(SLOTS)]
        self.ev_up = [torch.cuda.Event() for _ in range(SLOTS)]      # inputs of slot ready on the device
        self.ev_done = [torch.cuda.Event() for _ in range(SLOTS)]    # kernels of slot finished (inputs free, output snapshot ready)
        self.ev_down = [torch.cuda.Event() for _ in range(SLOTS)]    # host copy of slot complete
        self.i = 0
        self.socket_every_step = sampler_socket_every_step
        N_ = N
        self._ones = torch.ones(N_, dtype=torch.bool, device=dev)
        self._zeros = torch.zeros(N_, dtype=torch.bool, device=dev)
        self.h2d_bytes = sum(t.numel() * t.element_size() for t in self.d_in[0].values())
        self.d2h_bytes = self.h_out[0].numel() * 4

    @torch.no_grad()
    def step(self, fpos, fquat, ppos, pquat, depth, seg, update=None):
        """Submit one step (host tensors, ideally pinned); never blocks the host.  Returns a
        PendingObs for this step's observations.  `seg` may be int32 (the reference's tensor type) or uint8
        (the ids are 0..3): the uint8 form uploads a quarter of the bytes and is widened on the device."""
        self.last_h2d_bytes = sum(t.numel() * t.element_size() for t in (fpos, fquat, ppos, pquat, depth, seg))
        task, i = self.task, self.i
        slot = i % SLOTS
        cur = torch.cuda.current_stream(self.dev)
        # upload into the slot once the kernels that last read it (step i-SLOTS) are done
        with torch.cuda.stream(self.s_up):
            if i >= SLOTS:
                self.s_up.wait_event(self.ev_done[slot])
            d = self.d_in[slot]
            for k, src in (("fpos", fpos), ("fquat", fquat), ("ppos", ppos), ("pquat", pquat), ("depth", depth)):
                d[k].copy_(src.reshape(d[k].shape), non_blocking=True)
            if seg.dtype == torch.uint8:
                if self.d_seg8 is None:
                    self.d_seg8 = [torch.empty(d["seg"].shape, dtype=torch.uint8, device=self.dev) for _ in range(SLOTS)]
                self.d_seg8[slot].copy_(seg.reshape(d["seg"].shape), non_blocking=True)
                d["seg"].copy_(self.d_seg8[slot])            # widen on the device (upload stream)
            else:
                d["seg"].copy_(seg.reshape(d["seg"].shape), non_blocking=True)
            self.ev_up[slot].record(self.s_up)
        # kernels
        cur.wait_event(self.ev_up[slot])
        if i >= SLOTS:
            cur.wait_event(self.ev_down[slot])       # the download that read d_out[slot] is done
        task.left_finger_pos, task.right_finger_pos, task.middle_finger_pos = d["fpos"][:, 0], d["fpos"][:, 1], d["fpos"][:, 2]
        task.left_finger_quat, task.right_finger_quat, task.middle_finger_quat = d["fquat"][:, 0], d["fquat"][:, 1], d["fquat"][:, 2]
        task.plug_pos, task.plug_quat = d["ppos"], d["pquat"]
        task.cam_renders, task.seg_renders = d["depth"], d["seg"]
        up = self._ones if update is None else update
        if task.pcl_cam and self.socket_every_step:
            task.invalidate_socket_cache()
        task.compute_observations(up, up, up, up, up, self._zeros, self._zeros)
        self.d_out[slot].copy_(task.obs_packed)
        self.ev_done[slot].record(cur)
        # download
        with torch.cuda.stream(self.s_down):
            self.s_down.wait_event(self.ev_done[slot])
            self.h_out[slot].copy_(self.d_out[slot], non_blocking=True)
            self.ev_down[slot].record(self.s_down)
        self.i += 1
        self._last = PendingObs(self.h_out[slot], self.ev_down[slot], self.ev_up[slot])
        return self._last

    def flush(self):
        """Wait for the last submitted step and return its host observations."""
        if self.i == 0:
            return None
        return self._last.wait()


class GraphedObsStep:
    """`compute_observations` of a task captured ONCE into a CUDA graph and replayed per step: at the env
    counts the reference itself runs with these observations on (10 and 256, scripts/train_s3.sh:5,
    train_s2.sh:5) a step is a dozen launches whose host cost exceeds their device time; the replay is one
    launch.  Inputs are read from the tensors the task attributes point at when the graph is captured
    (`left_finger_pos`, ..., `cam_renders`, `seg_renders`): write new values INTO them between replays.
    Needs the FPS sampler (the reference sampler reads the host RNG every step) and update masks that live
    in fixed tensors (default: all ones / no noise)."""

    def __init__(self, task, masks=None, socket_every_step=False, warmup=3):
        if task.pcl_cam and task.sampler != "fps":
            raise RuntimeError("GraphedObsStep needs sampler='fps' (the reference sampler uploads host RNG words per step)")
        self.task = task
        dev = task.device
        N = task.num_envs
        ones = torch.ones(N, dtype=torch.bool, device=dev)
        zeros = torch.zeros(N, dtype=torch.bool, device=dev)
        self.masks = masks if masks is not None else (ones, ones, ones, ones, ones, zeros, zeros)
        self.socket_every_step = socket_every_step
        eng = task.tactile_engine if task.tactile else None
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(max(warmup, 1)):       # first-use work (function attributes, scratch, side stream) outside the capture
                self._body()
            if eng is not None:
                eng.check_overflow()
                eng.capturing = True               # no host-visible telemetry inside the graph
            try:
                self.graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(self.graph, stream=side):
                    self._body()
            finally:
                if eng is not None:
                    eng.capturing = False
        torch.cuda.current_stream(dev).wait_stream(side)

    def _body(self):
        t = self.task
        if t.pcl_cam and self.socket_every_step:
            t.invalidate_socket_cache()
        t.compute_observations(*self.masks)

    def __call__(self):
        self.graph.replay()
        return self.task.obs_packed

    def check_overflow(self):
        """Triangle-list overflow of the replays so far (host sync; the graph itself carries no telemetry)."""
        if self.task.tactile:
            eng = self.task.tactile_engine
            eng._ovf_host.copy_(eng._counters[2:4])
            eng._ovf_event = torch.cuda.Event()
            eng._ovf_event.record(torch.cuda.current_stream(eng.device))
            return eng.check_overflow()
        return False
