"""Multi-GPU plumbing: one process per GPU, each owning a contiguous env slice; the only
collective on the observation path is ONE all-gather per step of the packed observation
rows onto every rank (the learner reads rank 0's copy).  SURVEY.md 8e.

The reference exchanges no observations (each rank trains on its own envs,
isaacgyminsertion/train.py:58-64; ext_adapt.py:172-178 only all-reduces gradients), so
this is the glue the north star adds, not a replacement of reference code.
"""
import ctypes
import os

import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """Rendezvous from torchrun's env (RANK / LOCAL_RANK / WORLD_SIZE / MASTER_*)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, local_rank, world


def bind_to_gpu_numa(local_rank):
    """Bind this process to the CPUs NVML reports as local to GPU `local_rank` and prefer that NUMA node for its
    memory, BEFORE pinned host buffers are allocated (their pages are placed by the allocating thread's policy).
    With 8 ranks on a two-socket host, unbound ranks stage half of their host<->device traffic through the other
    socket.  Returns a small dict for the bench line (or {"error": ...}); never raises."""
    try:
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        idx = (int(vis.split(",")[local_rank]) if vis and all(v.strip().isdigit() for v in vis.split(","))
               else local_rank)
        h = pynvml.nvmlDeviceGetHandleByIndex(idx)
        ncpu = os.cpu_count() or 1
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (int(m) >> b) & 1}
        allowed = os.sched_getaffinity(0)
        use = cpus & allowed
        if not use:
            return {"bound": False, "why": "no overlap between the GPU-local CPUs and the allowed CPUs"}
        os.sched_setaffinity(0, use)
        node = None
        base = "/sys/devices/system/node"
        if os.path.isdir(base):
            for name in sorted(os.listdir(base)):
                if name.startswith("node") and name[4:].isdigit():
                    try:
                        txt = open(os.path.join(base, name, "cpulist")).read().strip()
                    except OSError:
                        continue
                    members = set()
                    for part in txt.split(","):
                        if part:
                            a, _, b = part.partition("-")
                            members.update(range(int(a), int(b or a) + 1))
                    if min(use) in members:
                        node = int(name[4:])
                        break
        mem = False
        if node is not None and node < 64:
            try:     # set_mempolicy(MPOL_PREFERRED, {node}): x86_64 syscall 238, aarch64 237
                import ctypes
                import platform
                nr = 238 if platform.machine() == "x86_64" else 237
                m = ctypes.c_ulong(1 << node)
                mem = ctypes.CDLL(None, use_errno=True).syscall(nr, 1, ctypes.byref(m), 65) == 0
            except Exception:
                mem = False
        return {"bound": True, "gpu": idx, "cpus": len(use), "cpu_range": [min(use), max(use)], "node": node,
                "mem_preferred": bool(mem)}
    except Exception as e:      # no NVML, no permission: run unbound
        return {"bound": False, "why": f"{type(e).__name__}: {e}"}


def shard_word_window(nonempty_per_rank, rank, m):
    """Reference sampler on sharded envs (pcl_utils.BatchedPointCloud.sample_reference): the reference draws m
    `torch.randint` words per NON-EMPTY env in global env order, so rank `rank` starts `skip` words into the stream
    and the whole job consumes `total` words.  -> (skip, total)."""
    counts = [int(c) for c in nonempty_per_rank]
    return m * sum(counts[:rank]), m * sum(counts)


def env_slice(total_envs, rank, world):
    """Contiguous slice [lo, hi) of global env ids owned by `rank` (remainder to low ranks)."""
    base, rem = divmod(total_envs, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_observations(obs_packed, total_envs=None, out=None):
    """All-gather the packed [n_local, row] observation rows of every rank into
    [total_envs, row] (global env order).  Equal slices use one all_gather_into_tensor;
    ragged slices (total_envs % world != 0) pad to the largest slice."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return obs_packed
    world = dist.get_world_size()
    n_local, row = obs_packed.shape
    if total_envs is None:
        total_envs = n_local * world
    if total_envs % world == 0:
        if out is None:
            out = torch.empty((total_envs, row), dtype=obs_packed.dtype, device=obs_packed.device)
        dist.all_gather_into_tensor(out, obs_packed.contiguous())
        return out
    n_max = -(-total_envs // world)
    padded = torch.zeros((n_max, row), dtype=obs_packed.dtype, device=obs_packed.device)
    padded[:n_local] = obs_packed
    buf = torch.empty((world * n_max, row), dtype=obs_packed.dtype, device=obs_packed.device)
    dist.all_gather_into_tensor(buf, padded)
    parts = []
    for r in range(world):
        lo, hi = env_slice(total_envs, r, world)
        parts.append(buf[r * n_max: r * n_max + (hi - lo)])
    return torch.cat(parts, dim=0)


# ---------------------------------------------------------------------------------------------------
# per-step gather object: NCCL all-gather, or copy-engine peer copies into the learner's buffer
# ---------------------------------------------------------------------------------------------------
class _DevMem:
    """Raw device block as a __cuda_array_interface__ object (torch.as_tensor aliases it)."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(ptr), False),
                                         "version": 3, "strides": None}


def _view(ptr, nbytes, device, dtype, shape):
    return torch.as_tensor(_DevMem(ptr, nbytes), device=device).view(dtype).view(shape)


class ObsGather:
    """One gather of the packed observation rows per env step.

    transport="nccl": `all_gather_into_tensor` on a side stream (every rank receives every row; the
        learner reads rank `learner`'s copy).  NCCL's copy kernels need SMs, which tac_contact's persistent
        CTAs own for most of a step.
    transport="p2p": every rank copies its rows straight into the LEARNER's buffer (one cudaMalloc block
        exported over CUDA IPC) with a copy-engine peer copy over NVLink and publishes the step number with
        a second 4-byte copy; the learner's stream waits for the step number of every rank with a stream
        memory operation (`cuStreamWaitValue32`), peers wait for the learner's "consumed" word the same way
        before they overwrite a slot.  No kernel, no SM, 1/world of the all-gather's traffic.

    gather(obs) enqueues this step's transfer behind the work already on the current stream and returns a
    ticket; wait(ticket) makes the current stream wait for it and returns the (total_envs, row) tensor on the
    learner (None on the other ranks with "p2p").  The tensor stays valid until the next gather() call.
    """

    def __init__(self, obs_like, total_envs=None, transport="p2p", learner=0, slots=2):
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        self.device = obs_like.device
        self.rows, self.row = obs_like.shape
        self.total = int(total_envs) if total_envs is not None else self.rows * self.world
        self.transport = transport if self.world > 1 else "none"
        self.learner, self.slots = int(learner), max(int(slots), 2)
        self.lo, self.hi = env_slice(self.total, self.rank, self.world)
        if self.hi - self.lo != self.rows:
            raise RuntimeError(f"rank {self.rank} owns envs [{self.lo},{self.hi}) but obs has {self.rows} rows")
        self.step = 0
        self.bytes_per_step = self.rows * self.row * 4
        self.comm_ms = None
        if self.transport == "none":
            return
        if self.device.type != "cuda":
            # host tensors (gloo, the CPU tests): one synchronous collective per step, same slot rotation
            if transport != "nccl":
                raise RuntimeError("ObsGather on host tensors supports only the collective transport ('nccl')")
            self.transport = "host"
            self._out = [torch.empty((self.total, self.row), dtype=torch.float32) for _ in range(self.slots)]
            return
        self.stream = torch.cuda.Stream(device=self.device)
        self._ev_ready = torch.cuda.Event()
        if self.transport == "nccl":
            self._out = [torch.empty((self.total, self.row), dtype=torch.float32, device=self.device)
                         for _ in range(self.slots)]
            self._send = [torch.empty((self.rows, self.row), dtype=torch.float32, device=self.device)
                          for _ in range(self.slots)] if self.total % self.world == 0 else None
            self._work = None
        elif self.transport == "p2p":
            self._init_p2p()
        else:
            raise ValueError(f"unknown transport {transport!r}")

    # ---- p2p setup --------------------------------------------------------------------------------
    def _init_p2p(self):
        """Collective.  Either every rank ends up with the learner's block mapped and working stream memory
        operations, or EVERY rank raises RuntimeError (so a caller can fall back to the NCCL transport on all ranks
        at once): each step that can fail locally is followed by an exchange of the outcome."""
        from . import _lib
        self.lib = lib = _lib.load()
        self._check = _lib.check
        c = ctypes
        is_learner = self.rank == self.learner
        slot_bytes = self.total * self.row * 4
        # learner: data slots + one "arrived" word per rank (padded to 256 B);  everyone: a "consumed" word + staging
        my_bytes = (self.slots * slot_bytes + 256 * self.world) if is_learner else 0
        my_bytes += 1024
        ptr, handle = c.c_void_p(), (c.c_ubyte * 64)()
        err = None
        self._own, self._mapped = None, {}
        try:
            with torch.cuda.device(self.device):
                self._check(lib.igi_peer_alloc(c.c_ulonglong(my_bytes), c.byref(ptr), handle), "igi_peer_alloc")
                self._own = ptr.value
                # stream memory operations must work on this device (probe on a spare word of the own block)
                probe = self._own + my_bytes - 1024 + 16
                st = c.c_void_p(self.stream.cuda_stream)
                self._check(lib.igi_stream_write_value32(st, c.c_void_p(probe), c.c_uint(1)), "igi_stream_write_value32")
                self._check(lib.igi_stream_wait_value32_geq(st, c.c_void_p(probe), c.c_uint(1)), "igi_stream_wait_value32_geq")
                self.stream.synchronize()
        except RuntimeError as e:
            err = str(e)
        outcomes = [None] * self.world
        dist.all_gather_object(outcomes, (bytes(handle), err))
        self._raise_if_any(outcomes, "allocating the peer block / stream memory operations")
        handles = [o[0] for o in outcomes]
        # control words at the END of every rank's block: [consumed | staging_a | staging_b | probe]
        self._ctl = self._own + my_bytes - 1024
        err = None
        try:
            if is_learner:
                self._data = self._own
                self._arrived = self._own + self.slots * slot_bytes
                self._peer_ctl = {}
                for r in range(self.world):
                    if r == self.rank:
                        continue
                    p = c.c_void_p()
                    h = (c.c_ubyte * 64).from_buffer_copy(handles[r])
                    with torch.cuda.device(self.device):
                        self._check(lib.igi_peer_open(h, c.byref(p)), "igi_peer_open")
                    self._mapped[r] = p.value
                    self._peer_ctl[r] = p.value            # a non-learner block is just its control words
                self._slot_views = [_view(self._data + i * slot_bytes, slot_bytes, self.device, torch.float32,
                                          (self.total, self.row)) for i in range(self.slots)]
                self._ev_done = torch.cuda.Event()
                self._ev_t0 = torch.cuda.Event(enable_timing=True)
                self._ev_t1 = torch.cuda.Event(enable_timing=True)
            else:
                p = c.c_void_p()
                h = (c.c_ubyte * 64).from_buffer_copy(handles[self.learner])
                with torch.cuda.device(self.device):
                    self._check(lib.igi_peer_open(h, c.byref(p)), "igi_peer_open")
                self._mapped[self.learner] = p.value
                self._data = p.value
                self._arrived = p.value + self.slots * slot_bytes
        except RuntimeError as e:
            err = str(e)
        outcomes = [None] * self.world
        dist.all_gather_object(outcomes, (None, err))
        self._raise_if_any(outcomes, "mapping the peer blocks over CUDA IPC")
        self._slot_bytes = slot_bytes
        # local snapshots of this rank's rows (two, alternating): the copy over NVLink reads the snapshot, so the next
        # step's kernels can rewrite the rows while the transfer is still in flight
        self._ev_sent = [torch.cuda.Event(), torch.cuda.Event()]
        self._snap = None if is_learner else [torch.empty((self.rows, self.row), dtype=torch.float32, device=self.device)
                                              for _ in range(2)]
        dist.barrier()

    def _raise_if_any(self, outcomes, what):
        bad = [(r, o[1]) for r, o in enumerate(outcomes) if o[1]]
        if not bad:
            return
        with torch.cuda.device(self.device):       # release what this rank holds, then fail on every rank alike
            for p in self._mapped.values():
                self.lib.igi_peer_close(ctypes.c_void_p(p))
            self._mapped = {}
            dist.barrier()
            if self._own:
                self.lib.igi_peer_free(ctypes.c_void_p(self._own))
                self._own = None
        self.transport = "none"
        raise RuntimeError(f"ObsGather p2p transport unavailable ({what}): " +
                           "; ".join(f"rank {r}: {m}" for r, m in bad))

    # ---- per step ---------------------------------------------------------------------------------
    @torch.no_grad()
    def gather(self, obs):
        if self.transport == "none":
            return obs
        self.step += 1
        s, slot = self.step, self.step % self.slots
        if self.transport == "host":
            out = gather_observations(obs, self.total, out=self._out[slot] if self.total % self.world == 0 else None)
            return ("host", out, None)
        cur = torch.cuda.current_stream(self.device)
        if self.transport == "nccl":
            if self._work is not None:
                self._work.wait()
            if self._send is None:          # ragged slices: padded gather
                snap = obs.clone()
                self._ev_ready.record(cur)
                with torch.cuda.stream(self.stream):
                    self.stream.wait_event(self._ev_ready)
                    out = gather_observations(snap, self.total)
                    ev = torch.cuda.Event()
                    ev.record(self.stream)
                return ("nccl", out, ev)
            buf = self._send[slot]
            buf.copy_(obs)                  # snapshot: the kernels of the next step rewrite obs while NCCL reads
            self._ev_ready.record(cur)
            with torch.cuda.stream(self.stream):
                self.stream.wait_event(self._ev_ready)
                self._work = dist.all_gather_into_tensor(self._out[slot], buf, async_op=True)
            return ("nccl", self._out[slot], self._work)
        # ---- p2p
        if obs.shape != (self.rows, self.row) or not obs.is_contiguous() or obs.dtype != torch.float32:
            raise RuntimeError("ObsGather: obs must be the contiguous (rows, row) f32 tensor given at construction")
        c, lib = ctypes, self.lib
        st = c.c_void_p(self.stream.cuda_stream)
        is_learner = self.rank == self.learner
        with torch.cuda.device(self.device):
            # Snapshot of this rank's rows by a small kernel ON THE CALLER'S STREAM (0.05 ms for 140 MB): into the
            # learner's slot directly on the learner, into a local buffer elsewhere.  The next step's kernels may then
            # rewrite obs at once, and no local copy competes with host<->device traffic for a copy engine.
            dst = self._data + slot * self._slot_bytes + self.lo * self.row * 4
            snap = dst if is_learner else self._snap[s & 1].data_ptr()
            if not is_learner and s > 2:
                cur.wait_event(self._ev_sent[s & 1])    # the transfer of step s - 2 has read this snapshot buffer
            self._check(lib.igi_queue_push(c.c_void_p(snap), c.c_void_p(obs.data_ptr()), c.c_int(0),
                                           c.c_int64(self.row), c.c_int(self.rows), c.c_int(1),
                                           c.c_longlong(self.row), c.c_void_p(cur.cuda_stream)), "igi_queue_push")
            self._ev_ready.record(cur)                  # rows of this step are in place behind this point
            self.stream.wait_event(self._ev_ready)
            if is_learner:
                # everything the caller enqueued so far may still read the slot of step s - 1: that result is declared
                # consumed now (contract: valid until the next gather call)
                if s > 1:
                    self._check(lib.igi_stream_write_value32(st, c.c_void_p(self._ctl + 4), c.c_uint(s - 1)),
                                "igi_stream_write_value32")
                    for r, pctl in self._peer_ctl.items():
                        self._check(lib.igi_peer_copy_async(c.c_void_p(pctl), c.c_void_p(self._ctl + 4),
                                                            c.c_ulonglong(4), st), "igi_peer_copy_async")
                self._ev_t0.record(self.stream)
                self._check(lib.igi_stream_write_value32(st, c.c_void_p(self._arrived + 256 * self.rank), c.c_uint(s)),
                            "igi_stream_write_value32")
                for r in range(self.world):
                    if r != self.rank:
                        self._check(lib.igi_stream_wait_value32_geq(st, c.c_void_p(self._arrived + 256 * r),
                                                                    c.c_uint(s)), "igi_stream_wait_value32_geq")
                self._ev_t1.record(self.stream)
            else:
                if s - self.slots >= 1:
                    # the slot of step s was last used by step s - slots: wait until the learner consumed that step
                    self._check(lib.igi_stream_wait_value32_geq(st, c.c_void_p(self._ctl), c.c_uint(s - self.slots)),
                                "igi_stream_wait_value32_geq")
                # copy-engine peer copy over NVLink, then the step number into the learner's arrived[rank] word
                self._check(lib.igi_peer_copy_async(c.c_void_p(dst), c.c_void_p(snap),
                                                    c.c_ulonglong(self.bytes_per_step), st), "igi_peer_copy_async")
                self._check(lib.igi_stream_write_value32(st, c.c_void_p(self._ctl + 8), c.c_uint(s)),
                            "igi_stream_write_value32")
                self._check(lib.igi_peer_copy_async(c.c_void_p(self._arrived + 256 * self.rank), c.c_void_p(self._ctl + 8),
                                                    c.c_ulonglong(4), st), "igi_peer_copy_async")
                self._ev_sent[s & 1].record(self.stream)    # the snapshot buffer is free again behind this point
        ev = torch.cuda.Event()
        ev.record(self.stream)
        return ("p2p", self._slot_views[slot] if is_learner else None, ev)

    def wait(self, ticket):
        """Current stream waits for the ticket's transfer; returns the gathered tensor (learner) or None."""
        if self.transport == "none":
            return ticket
        kind, out, h = ticket
        if kind == "host":
            return out
        cur = torch.cuda.current_stream(self.device)
        if kind == "nccl" and not isinstance(h, torch.cuda.Event):
            h.wait()
            cur.wait_stream(self.stream)
        else:
            cur.wait_event(h)
        return out

    def protect_source(self):
        """Kept for callers written against the first version: every transport now snapshots the rows on the caller's
        stream inside gather(), so obs may be rewritten as soon as gather() has returned."""
        return None

    def last_comm_ms(self):
        """Learner, p2p: device time from 'own rows ready' to 'every rank's rows arrived' of the last step."""
        if self.transport == "p2p" and self.rank == self.learner and self.step > 0:
            self._ev_t1.synchronize()
            return self._ev_t0.elapsed_time(self._ev_t1)
        return None

    def close(self):
        if self.transport == "nccl" and self._work is not None:
            self._work.wait()
        if self.transport != "p2p":
            self.transport = "none"
            return
        torch.cuda.synchronize(self.device)
        dist.barrier()
        with torch.cuda.device(self.device):
            for p in self._mapped.values():
                self.lib.igi_peer_close(ctypes.c_void_p(p))
            self._mapped = {}
            dist.barrier()
            self._slot_views = None
            self.lib.igi_peer_free(ctypes.c_void_p(self._own))
        self.transport = "none"
