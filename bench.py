#!/usr/bin/env python
"""Benchmark of the observation hot path (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--config obs4096|tactile1024|pcl1024|sweep|small]
                  [--envs E] [--scaling weak|strong] [--falloff inverse_square|none] [--gather p2p|nccl]
  python bench.py --impl reference ...      # the CPU oracle on the host cores

A "step" = one pass of the visuotactile observation path over E envs per GPU:
E x 3 tactile frames (224x224 render -> 2048-float obs) + E point-cloud observations
(96x54 depth+seg -> 400 plug + 400 socket points), followed for N > 1 by the gather of the packed
observation rows onto the learner rank.  Unit: obs/s, 1 obs = one env's 3 tactile frames + 1 cloud.
Inputs are synthetic (IsaacGym is a closed dependency).

Configurations (BASELINE.json `configs`):
  obs4096     (default) config 4/5 headline: 4096 envs per GPU (weak scaling) or 4096 envs in total sharded
              over the ranks (`--scaling strong`, config 4 as written: 512 envs/rank at 8 GPUs)
  tactile1024 config 2: tactile path alone, 1024 envs
  pcl1024     config 3: point-cloud path alone, 1024 envs
  sweep       config 5: resident step at 256 ... 16384 envs per GPU, one JSON line with a `sweep` array
  small       the reference's own operating points (10 and 256 envs, scripts/train_s3.sh:5, train_s2.sh:5):
              eager launches and the CUDA-graph replay of the step
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

_REAL_STDOUT = None
METRIC = "visuotactile_obs_per_s"
UNIT = "obs/s"
TACTILE_BYTES_PER_FRAME = 861312      # SURVEY.md 8d / BASELINE.md 4
PCL_BYTES_PER_ENV_FPS = 51200
PCL_BYTES_PER_ENV_REF = 57600
FILL_BYTES_PER_FRAME = 150528 + 200704 + 8192   # color, gel_depth, obs written (bg_real / obs_empty sources are L2-resident)
SWEEP_ENVS = (256, 512, 1024, 2048, 4096, 8192, 16384)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--config", default="obs4096", choices=["obs4096", "tactile1024", "pcl1024", "sweep", "small"])
    ap.add_argument("--envs", type=int, default=None, help="envs per GPU (weak) / in total (strong); default from --config")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--sampler", default="fps", choices=["fps", "reference"])
    ap.add_argument("--falloff", default=None, choices=["inverse_square", "none"],
                    help="light model (DESIGN.md 'light model'); default = the shipped yaml (inverse_square)")
    ap.add_argument("--gather", default="p2p", choices=["p2p", "nccl"],
                    help="N > 1: copy-engine peer copies into the learner's buffer (default) or NCCL all-gather")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-components", action="store_true", help="skip the per-kernel timings / roofline block")
    ap.add_argument("--no-alt-falloff", action="store_true", help="skip the second resident leg under the other light model")
    ap.add_argument("--sync-gather", action="store_true", help="wait for every step's gather before the next step")
    ap.add_argument("--cpu-sample-envs", type=int, default=256)   # ~12 s of one host core
    ap.add_argument("--no-overlap", action="store_true", help="point-cloud path on the same stream as the tactile path")
    ap.add_argument("--no-numa", action="store_true", help="do not bind the process to the GPU's NUMA node")
    ap.add_argument("--one-input-set", action="store_true",
                    help="every timed step sees the same poses / camera frames (default: two seeded sets, alternating)")
    ap.add_argument("--seg-int32-host", action="store_true",
                    help="e2e leg: keep the host copy of the segmentation image as int32 (default: uint8, widened on the device)")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------
# synthetic workload
# ------------------------------------------------------------------------------------------
def make_inputs(n_envs, global_offset, total, seed=0):
    from isaacgyminsertion_b200 import assets, synthetic
    packed = assets.load_packed()
    gym = synthetic.SyntheticGym(n_envs, seed=seed, global_env_offset=global_offset, total_envs=total)
    P = synthetic.tactile_poses(n_envs, packed, seed=seed, global_env_offset=global_offset)
    _, _, socket_pos = synthetic.scene_poses(n_envs, seed=seed, assets=packed, global_env_offset=global_offset)
    depth, seg = synthetic.external_camera_frames(gym, P["plug_pos"].astype(np.float64),
                                                  P["plug_quat"].astype(np.float64), socket_pos, seed=seed)
    return gym, P, depth, seg


class ClockSampler(threading.Thread):
    """SM clock + throttle reasons during the timed region (B200_PROFILING.md clocks line).  NVML through
    pynvml (one sample per ~2 ms, so even an 80 ms timed region gets tens of samples); falls back to polling
    nvidia-smi when pynvml is missing."""

    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.stop_flag = False
        self.samples = []      # (sm_mhz, reasons bitmask over NAMES)
        self.sm_max = None
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[index]) if vis and all(v.strip().isdigit() for v in vis.split(",")) else index
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.sm_max = int(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def _sample_nvml(self):
        n = self.nvml
        sm = int(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM))
        r = int(n.nvmlDeviceGetCurrentClocksEventReasons(self.handle)) if hasattr(
            n, "nvmlDeviceGetCurrentClocksEventReasons") else int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
        bits = 0
        for i, mask in enumerate((0x8, 0x40, 0x20, 0x4)):   # HwSlowdown, HwThermalSlowdown, SwThermalSlowdown, SwPowerCap
            if r & mask:
                bits |= 1 << i
        self.samples.append((sm, bits))

    def _sample_smi(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i",
                              str(self.index)], capture_output=True, text=True, timeout=5).stdout.strip()
        if out:
            f = [x.strip() for x in out.split(",")]
            self.sm_max = int(f[1])
            bits = sum(1 << i for i in range(4) if f[2 + i].lower().startswith("active"))
            self.samples.append((int(f[0]), bits))

    def run(self):
        while not self.stop_flag:
            try:
                if self.nvml is not None:
                    self._sample_nvml()
                    time.sleep(0.002)
                else:
                    self._sample_smi()
                    time.sleep(0.05)
            except Exception:
                time.sleep(0.05)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.sm_max, "reasons": ["unavailable"], "samples": 0}
        sm = sorted(s[0] for s in self.samples)
        reasons = [n for i, n in enumerate(self.NAMES) if any(s[1] >> i & 1 for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_min_mhz": sm[0], "sm_max_mhz": self.sm_max, "reasons": reasons,
                "samples": len(self.samples), "source": "nvml" if self.nvml is not None else "nvidia-smi"}


# ------------------------------------------------------------------------------------------
# CPU oracle legs
# ------------------------------------------------------------------------------------------
_W = {}


def _cpu_worker_init():
    import cv2
    import torch
    cv2.setNumThreads(1)
    torch.set_num_threads(1)
    from oracle import tactile as ot
    _W["model"] = ot.SensorModel()


def _cpu_env(args):
    """One env of the reference's serial loop: 3 tactile frames + plug & socket cloud."""
    import torch
    from oracle import pcl as opcl
    from oracle import tactile as ot
    (mesh_id, bg_ids, fpos, fquat, ppos, pquat, proj, view, origin, depth, seg) = args
    model = _W["model"]
    obj_tf = ot.xyzquat_to_tf_numpy(np.concatenate([ppos, pquat])[None])[0]
    for n in range(3):
        h = ot.OracleAllSight(model, int(mesh_id), int(bg_ids[n]))
        ftf = ot.xyzquat_to_tf_numpy(np.concatenate([fpos[n], fquat[n]])[None])[0]
        h.update_pose_given_sim_pose(ftf, obj_tf)
        color, _ = h.render(obj_tf, 70)
        ot.tactile_obs(color, h.bg_img, h.mask)
    e2g = np.identity(4)
    e2g[:3, 3] = origin
    cam = opcl.CameraOracle(proj, view, e2g, depth.shape[1], depth.shape[0])
    d, s = torch.from_numpy(depth[None]), torch.from_numpy(seg[None])
    opcl.pcl_observation([cam], d, s)
    return 1


def cpu_jobs(gym, P, depth, seg, n):
    return [(P["mesh_id"][e], P["bg_id"][e], P["finger_pos"][e], P["finger_quat"][e], P["plug_pos"][e],
             P["plug_quat"][e], gym._proj[e], gym._view[e], gym.origins[e], depth[e], seg[e]) for e in range(n)]


def build_oracle():
    subprocess.run(["make", "-C", os.path.join(ROOT, "oracle")], check=True, capture_output=True)


CPU_SCENE_NOTE = ("oracle scene per frame = posed peg + the 21 824 gel triangles that can reach the in-gel camera "
                  "(of the reference's 231 146-triangle gel mesh, which pyrender submits whole): the port does LESS work "
                  "than the reference would, so GPU / CPU ratios built on it are conservative")


def cpu_baseline_serial(gym, P, depth, seg, n):
    """The oracle's serial per-env loop on ONE host core (reference structure)."""
    build_oracle()
    _cpu_worker_init()
    jobs = cpu_jobs(gym, P, depth, seg, n)
    _cpu_env(jobs[0])
    t0 = time.perf_counter()
    for j in jobs:
        _cpu_env(j)
    dt = time.perf_counter() - t0
    return {"value": n / dt, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": f"{n} envs (={3 * n} tactile frames + {n} plug+socket clouds), serial oracle loop, "
                      f"{dt:.2f} s; " + CPU_SCENE_NOTE}


def run_reference(args):
    """--impl reference: the CPU oracle with every host core (one process per core)."""
    import multiprocessing as mp
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    build_oracle()
    try:
        cores = len(os.sched_getaffinity(0))     # the cores this process may actually use (cgroup / affinity aware)
    except AttributeError:
        cores = os.cpu_count() or 1
    cores = max(cores, 1)
    n = max(cores * 8, 32)      # envs per step: ~0.6 s of every core per step
    gym, P, depth, seg = make_inputs(n, 0, n)
    jobs = cpu_jobs(gym, P, depth, seg, n)
    ctx = mp.get_context("fork")
    with ctx.Pool(cores, initializer=_cpu_worker_init) as pool:
        for _ in range(max(args.warmup, 1)):
            pool.map(_cpu_env, jobs[:cores])
        t0 = time.perf_counter()
        for _ in range(args.steps):
            pool.map(_cpu_env, jobs, chunksize=max(n // cores, 1))
        dt = time.perf_counter() - t0
    value = n * args.steps / dt
    envs = args.envs or 4096
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"visuotactile obs, bounded sample of {n} envs per step (same generators as the "
                               f"{envs}-env GPU workload)", "envs_per_step": n, "scene": CPU_SCENE_NOTE},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{n} envs/step x {args.steps} steps, {cores} worker processes"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------
class Workload:
    """One rank's task object + device-resident inputs + pinned host copies of them."""

    def __init__(self, E, offset, total, dev, args, tactile=True, pcl=True, falloff=None, pin=True, two_sets=False):
        import torch
        from isaacgyminsertion_b200.task_obs import FactoryTaskInsertionTactileObs
        self.E, self.dev = E, dev
        self.gym, self.P, self.depth_np, self.seg_np = make_inputs(E, offset, total)
        P = self.P
        self.task = FactoryTaskInsertionTactileObs(
            E, self.gym, P["mesh_id"], P["bg_id"], device=dev, sampler=args.sampler, strict_rng=False,
            overlap_streams=not args.no_overlap, tactile=tactile, pcl_cam=pcl, falloff=falloff,
            global_env_offset=offset, total_envs=total)

        def host(a):
            t = torch.from_numpy(np.ascontiguousarray(a))
            return t.pin_memory() if pin else t
        self.h = dict(fpos=host(P["finger_pos"]), fquat=host(P["finger_quat"]), ppos=host(P["plug_pos"]),
                      pquat=host(P["plug_quat"]), depth=host(self.depth_np), seg=host(self.seg_np))
        self.d = {k: v.to(dev) for k, v in self.h.items()}
        if not args.seg_int32_host:
            # host copy of the segmentation image as uint8 (ids 0..3): a quarter of the upload; widened on the device
            assert self.seg_np.min() >= 0 and self.seg_np.max() <= 255
            self.h["seg"] = host(self.seg_np.astype(np.uint8))
        self.sets_h, self.sets_d = [self.h], [self.d]
        if two_sets:
            # a second set of poses and camera frames (seed 1; same meshes and backgrounds, they are static per env):
            # the timed steps alternate between the two, so the contact mix under the work stealing is not frozen
            _, P2, depth2, seg2 = make_inputs(E, offset, total, seed=1)
            h2 = dict(fpos=host(P2["finger_pos"]), fquat=host(P2["finger_quat"]), ppos=host(P2["plug_pos"]),
                      pquat=host(P2["plug_quat"]), depth=host(depth2), seg=host(seg2))
            self.sets_d.append({k: v.to(dev) for k, v in h2.items()})
            if not args.seg_int32_host:
                h2["seg"] = host(seg2.astype(np.uint8))
            self.sets_h.append(h2)
        self.calls = 0
        self.ones = torch.ones(E, dtype=torch.bool, device=dev)
        self.zeros = torch.zeros(E, dtype=torch.bool, device=dev)
        self.load_state()

    def load_state(self, k=0):
        t, d = self.task, self.sets_d[k % len(self.sets_d)]
        t.left_finger_pos, t.right_finger_pos, t.middle_finger_pos = d["fpos"][:, 0], d["fpos"][:, 1], d["fpos"][:, 2]
        t.left_finger_quat, t.right_finger_quat, t.middle_finger_quat = d["fquat"][:, 0], d["fquat"][:, 1], d["fquat"][:, 2]
        t.plug_pos, t.plug_quat = d["ppos"], d["pquat"]
        t.cam_renders, t.seg_renders = d["depth"], d["seg"]

    def step(self):
        t = self.task
        if len(self.sets_d) > 1:
            self.load_state(self.calls)      # attribute swaps only: no copy, no launch
        self.calls += 1
        if t.pcl_cam:
            t.invalidate_socket_cache()   # worst case: every env restarted -> socket cloud recomputed each step
        # update_tactile + update_external_cam with the reference's mask arguments (task :862-887)
        o, z = self.ones, self.zeros
        t.compute_observations(o, o, o, o, o, z, z)


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
        return
    # rank 0 prints exactly ONE line on stdout.  Libraries write there too (NCCL's "NCCL version ..." banner comes
    # from C code), so fd 1 is pointed at stderr for the run and the JSON line goes to the saved real stdout.
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    import torch
    import torch.distributed as dist
    from isaacgyminsertion_b200 import _lib
    from isaacgyminsertion_b200 import dist as igdist

    rank, local_rank, world = igdist.init_from_env()
    if world != args.gpus and world > 1:
        args.gpus = world
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa = None if args.no_numa else igdist.bind_to_gpu_numa(local_rank)   # before any pinned allocation
    lib = _lib.load()

    cfg = args.config
    tactile, pcl = cfg != "pcl1024", cfg != "tactile1024"
    envs = args.envs or (1024 if cfg in ("tactile1024", "pcl1024") else 4096)
    if args.scaling == "strong":
        lo, hi = igdist.env_slice(envs, rank, world)
        E, offset, total = hi - lo, lo, envs
    else:
        E, offset, total = envs, rank * envs, envs * world

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup, after=None):
        for i in range(warmup):
            fn(i)
        if after:
            after()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(i)
        if after:
            after()      # drains the copy / gather streams: every step's result has landed
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    W = max(args.warmup, 3)

    # ---- sweep / small: resident step at several sizes, one line -------------------------------------------
    if cfg in ("sweep", "small"):
        from isaacgyminsertion_b200.pipeline import GraphedObsStep
        sizes = SWEEP_ENVS if cfg == "sweep" else (10, 256)
        rows = []
        sampler = ClockSampler(local_rank)
        sampler.start()
        launches = 0
        for n in sizes:
            wl = Workload(n, rank * n, n * world, dev, args, falloff=args.falloff, pin=False)
            l0 = lib.igi_launch_count()
            ms = timed(lambda i: wl.step(), args.steps, W) / args.steps
            per_step = (lib.igi_launch_count() - l0) // (args.steps + W)
            launches += per_step * args.steps
            row = {"envs_per_gpu": n, "ms_per_step": ms, "obs_per_s": n * world / (ms * 1e-3), "launches_per_step": int(per_step)}
            if n <= 1024:
                g = GraphedObsStep(wl.task, socket_every_step=True)
                row["graph_ms_per_step"] = timed(lambda i: g(), args.steps, W) / args.steps
                row["graph_obs_per_s"] = n * world / (row["graph_ms_per_step"] * 1e-3)
                g.check_overflow()
            wl.task.tactile_engine.check_overflow()
            rows.append(row)
            del wl
            torch.cuda.empty_cache()
        sampler.stop_flag = True
        head = next((r for r in rows if r["envs_per_gpu"] == 4096), rows[-1])
        if rank == 0:
            line = {"metric": METRIC, "value": head["obs_per_s"], "unit": UNIT, "n_gpus": world, "steps": args.steps,
                    "warmup": W, "ms_per_step": head["ms_per_step"], "higher_is_better": True, "scaling": "weak",
                    "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                    "config": {"workload": f"FactoryTaskInsertionTactile visuotactile obs, resident step at "
                                           f"{list(sizes)} envs per GPU (value = the {head['envs_per_gpu']}-env point)",
                               "sweep": cfg, "sampler": args.sampler, "falloff": args.falloff or "inverse_square",
                               "gather": "none (resident step of each rank; the gather is timed by the default config)"},
                    "sweep": rows, "clocks": sampler.summary(), "gpu_launches": int(launches), "impl": "b200"}
            _REAL_STDOUT.write(json.dumps(line) + "\n")
            _REAL_STDOUT.flush()
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- one size: the headline (obs4096) or a single-path config ----------------------------------------------
    wl = Workload(E, offset, total, dev, args, tactile=tactile, pcl=pcl, falloff=args.falloff, two_sets=not args.one_input_set)
    task = wl.task
    gather, gather_note = None, None
    if world > 1:
        try:
            gather = igdist.ObsGather(task.obs_packed, total_envs=total, transport=args.gather)
        except RuntimeError as e:      # raised on every rank alike (dist.ObsGather._init_p2p): fall back together
            if args.gather != "p2p":
                raise
            gather_note = f"p2p transport unavailable on this box, NCCL all-gather used instead: {e}"[:400]
            args.gather = "nccl"
            gather = igdist.ObsGather(task.obs_packed, total_envs=total, transport="nccl")
    pending = [None]

    def obs_step(i):
        if gather is not None:
            gather.protect_source()        # last step's transfer has read obs_packed
        wl.step()
        if gather is not None:
            t = gather.gather(task.obs_packed)
            if args.sync_gather:
                gather.wait(t)
            pending[0] = t

    def finish():
        if gather is not None and pending[0] is not None:
            gather.wait(pending[0])
            pending[0] = None

    for i in range(W):
        obs_step(i)
    finish()
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = lib.igi_launch_count()
    ms = timed(obs_step, args.steps, 0, after=finish)
    launches = int(lib.igi_launch_count() - launches0)   # kernels of libigi_b200.so inside the timed region
    sampler.stop_flag = True
    if tactile:
        task.tactile_engine.check_overflow()
    value = total * args.steps / (ms * 1e-3)
    comm = None
    if gather is not None:
        comm = {"transport": args.gather, "bytes_per_rank_per_step": gather.bytes_per_step,
                "bytes_into_learner_per_step": gather.bytes_per_step * (world - 1),
                "bytes_received_per_rank_per_step": gather.bytes_per_step * (world - 1) if args.gather == "nccl" else 0,
                "learner_wait_ms_last_step": gather.last_comm_ms(),
                "overlap": "none (sync)" if args.sync_gather else "transfer of step i under the kernels of step i+1",
                "sm_use": "none: copy-engine peer copies + stream memory ops" if args.gather == "p2p" else "NCCL all-gather kernels"}
        if gather_note:
            comm["note"] = gather_note

    # ---- the same resident leg under the other light model (single GPU): the two exact early-outs of tac_contact
    # only fire when fragments clip, so the default number is reported next to the one that cannot use them
    extra = {}
    if tactile and world == 1 and not args.no_alt_falloff:
        other = "none" if (args.falloff or "inverse_square") == "inverse_square" else "inverse_square"
        wl2 = Workload(E, offset, total, dev, args, tactile=tactile, pcl=pcl, falloff=other, pin=False,
                       two_sets=not args.one_input_set)
        ms2 = timed(lambda i: wl2.step(), args.steps, W) / args.steps
        t2 = timed(lambda i: wl2.task.update_tactile(wl2.ones, wl2.ones), max(args.steps // 2, 5), 3) / max(args.steps // 2, 5)
        extra["alt_falloff"] = {"falloff": other, "ms_per_step": ms2, "value": total / (ms2 * 1e-3), "unit": UNIT,
                                "tactile_pipeline_ms": t2,
                                "tactile_roofline_frac": TACTILE_BYTES_PER_FRAME * 3 * E / (t2 * 1e-3) / 1e9 / _peak()[0]}
        del wl2
        torch.cuda.empty_cache()

    # ---- component timings + roofline (rank 0's GPU, same buffers, CUDA events) -----------------
    if not args.no_components:
        K = max(args.steps // 2, 5)
        frames = 3 * E
        peak, peak_src = _peak()

        def gbs(nbytes, t_ms):
            return nbytes / (t_ms * 1e-3) / 1e9
        kernels = {}
        ones, zeros = wl.ones, wl.zeros
        if tactile:
            eng = task.tactile_engine
            fp = (task.left_finger_pos, task.right_finger_pos, task.middle_finger_pos)
            fq = (task.left_finger_quat, task.right_finger_quat, task.middle_finger_quat)

            def stage(mask):
                return lambda i: eng.render(fp, fq, task.plug_pos, task.plug_quat, obs_out=task.tactile_imgs,
                                            stage_mask=mask)
            t_tac = timed(lambda i: task.update_tactile(ones, ones), K, 3) / K
            t_geom = timed(stage(1), K, 3) / K
            t_geomfill = timed(stage(8), K, 3) / K
            t_fill = timed(stage(2), K, 3) / K
            t_contact = timed(stage(4), K, 3) / K
            kernels.update({
                "tac_fill": {"ms": t_fill, "bytes": FILL_BYTES_PER_FRAME * frames},
                "tac_geom": {"ms": t_geom, "bytes": None},
                "tac_geom_fused_fill": {"ms": t_geomfill, "bytes": FILL_BYTES_PER_FRAME * frames},
                "tac_contact": {"ms": t_contact, "bytes": None},
                "tactile_pipeline": {"ms": t_tac, "bytes": TACTILE_BYTES_PER_FRAME * frames}})
        if pcl:
            def pcl_only(i):
                task.invalidate_socket_cache()
                task.update_external_cam(ones, ones, ones, zeros, zeros)
            t_pcl = timed(pcl_only, K, 3) / K
            gen = task.pcl_generator.engine
            from isaacgyminsertion_b200.pcl_utils import filter_pts
            d_depth, d_seg = wl.d["depth"], wl.d["seg"]
            t_compact = timed(lambda i: gen.compact(d_depth, d_seg, (2, 3), filter_pts.box), K, 3) / K
            pts, cnt, any_ = gen.compact(d_depth, d_seg, (2, 3), filter_pts.box)
            t_fps = timed(lambda i: gen.sample_fps(pts, cnt, any_, 0, 400, out=task._plug_pts), K, 3) / K
            t_fps_both = timed(lambda i: gen.sample_fps(pts, cnt, any_, None, 400, out=task._both_pts), K, 3) / K
            kernels.update({
                "pcl_compact": {"ms": t_compact, "bytes": (20736 * 2 + 128 + 96 * 4 + 54 * 4) * E},
                "pcl_fps_plug": {"ms": t_fps, "bytes": None},
                "pcl_fps_plug_socket": {"ms": t_fps_both, "bytes": None},
                "pcl_pipeline": {"ms": t_pcl, "bytes": (PCL_BYTES_PER_ENV_FPS if args.sampler == "fps"
                                                        else PCL_BYTES_PER_ENV_REF) * E}})
        for k, v in kernels.items():
            if v["bytes"]:
                v["gbs"] = gbs(v["bytes"], v["ms"])
                v["frac"] = v["gbs"] / peak
        if tactile:
            # Dominant work = the tactile path.  Its two launches per step (tac_geom with the fused no-contact
            # fill, then tac_contact over the frames with surviving triangles) share ONE algorithmic byte budget
            # per frame (861 312 B), so they are reported together: achieved = frames x 861 312 B /
            # (t_geom_fill + t_contact), both measured alone with CUDA events on the launching stream.  Charging
            # the whole budget to either launch alone would overstate it.
            t_dom = t_geomfill + t_contact
            dom_bytes = TACTILE_BYTES_PER_FRAME * frames
            ach = gbs(dom_bytes, t_dom)
            traffic = None     # measured DRAM traffic of those launches (ncu --set full capture of this command, profiles/)
            try:
                tr = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
                if tr.get("envs_per_gpu") == E:
                    ks = tr["kernels"]
                    traffic = sum(ks[k]["dram_bytes_read"] + ks[k]["dram_bytes_write"] for k in ("tac_geom_fused_fill", "tac_contact"))
            except Exception:
                pass
            counts = eng.contact_counts()
            extra["roofline"] = {"bound": "hbm", "kernel": "tac_geom(+fill) + tac_contact = the tactile path, 2 launches/step",
                                 "achieved": ach, "peak": peak, "unit": "GB/s",
                                 "frac": ach / peak, "traffic": traffic, "peak_source": peak_src,
                                 "algorithmic_bytes_per_launch": dom_bytes, "ms_per_launch": t_dom,
                                 "share_of_step": t_dom / (ms / args.steps),
                                 "hbm_floor_ms": FILL_BYTES_PER_FRAME * frames / (peak * 1e9) * 1e3,
                                 "note": "mandatory DRAM traffic is the 359 424 B/frame of outputs (static sources are L2-resident); "
                                         "tac_contact is issue-bound (rasterise + shade), not DRAM-bound",
                                 "pipeline": {"tactile": {"ms": t_tac, "GBps": kernels["tactile_pipeline"]["gbs"],
                                                          "frac": kernels["tactile_pipeline"]["frac"]}}}
            extra["contact"] = {"frames": frames, "frames_with_candidates": int((counts > 0).sum().item()),
                                "mean_candidate_tris": float(counts.clamp(min=0).float().mean().item()),
                                "max_candidate_tris": int(counts.max().item()), "kmax": eng.kmax}
            extra["tactile_frames_per_s"] = frames / (t_tac * 1e-3) * world
        if pcl:
            pr = {"ms": t_pcl, "GBps": kernels["pcl_pipeline"]["gbs"], "frac": kernels["pcl_pipeline"]["frac"]}
            if tactile:
                extra["roofline"]["pipeline"]["pcl"] = pr
            else:
                extra["roofline"] = {"bound": "hbm", "kernel": "pcl_compact + FPS = the point-cloud path", "achieved": pr["GBps"],
                                     "peak": peak, "unit": "GB/s", "frac": pr["frac"], "traffic": None, "peak_source": peak_src,
                                     "algorithmic_bytes_per_launch": kernels["pcl_pipeline"]["bytes"], "ms_per_launch": t_pcl,
                                     "note": "FPS is a chain of dependent picks (latency / issue bound), not DRAM-bound"}
            extra["pcl_obs_per_s"] = E / (t_pcl * 1e-3) * world
        extra["kernels"] = kernels

    # ---- end to end: host buffers in, host result out, every step ------------------------------
    e2e = None
    if not args.no_e2e:
        from isaacgyminsertion_b200.pipeline import HostObsPipeline
        from isaacgyminsertion_b200 import pipeline as _pl
        E2E_DEPTH = _pl.SLOTS - 1
        pipe = HostObsPipeline(task, sampler_socket_every_step=True)
        handles = []

        def e2e_step(i):
            h = wl.sets_h[i % len(wl.sets_h)]
            # host buffers in, host result out, every step: upload / kernels / download of
            # neighbouring steps overlap on three streams (isaacgyminsertion_b200.pipeline)
            if gather is not None:
                gather.protect_source()
            handles.append(pipe.step(h["fpos"], h["fquat"], h["ppos"], h["pquat"], h["depth"], h["seg"]))
            if len(handles) > E2E_DEPTH:
                handles.pop(0).wait()      # the learner reads the observations of step i - (SLOTS - 1)
            if gather is not None:
                pending[0] = gather.gather(task.obs_packed)
        # the same K steps as the resident leg; the timed region includes filling and draining the 3-stage
        # pipeline (first upload before any kernel, last download after the last kernel: ~5.6 ms at 4096 envs)
        K = max(args.steps, 5)

        def e2e_finish():
            while handles:
                handles.pop(0).wait()
            finish()
        ms_e2e = timed(e2e_step, K, 3, after=e2e_finish)
        e2e = {"value": total * K / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": pipe.last_h2d_bytes,
               "seg_host_dtype": str(wl.sets_h[0]["seg"].dtype).replace("torch.", ""),
               "d2h_bytes_per_step": pipe.d2h_bytes, "ms_per_step": ms_e2e / K, "steps": K,
               "overlap": "upload / kernels / download of neighbouring steps on 3 streams, 3-slot ring",
               "numa": numa}
        wl.load_state()

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline_serial(wl.gym, wl.P, wl.depth_np, wl.seg_np, min(args.cpu_sample_envs, E))

    if gather is not None:
        gather.close()
    if rank == 0:
        what = ("3 allsight 224x224 tactile frames" if tactile else "") + (" + " if tactile and pcl else "") + \
               ("96x54 depth+seg -> 400+400-pt cloud" if pcl else "")
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": W,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"FactoryTaskInsertionTactile visuotactile obs ({cfg}): {E} envs/GPU x ({what}), "
                                   f"sampler={args.sampler}, socket cloud recomputed every step",
                       "config": cfg, "envs_per_gpu": E, "total_envs": total, "sensors_per_env": 3, "sampler": args.sampler,
                       "falloff": args.falloff or "inverse_square (shipped yaml)",
                       "inputs": ("one seeded set of poses / camera frames" if args.one_input_set else
                                  "two seeded sets of poses / camera frames, alternating every step"),
                       "l2": "per-step outputs (4.4 GB at 4096 envs) and depth/seg inputs (170 MB) exceed the 126 MB L2",
                       "gather": ("none" if world == 1 else f"{args.gather}, " + ("sync" if args.sync_gather else "overlapped"))},
            "clocks": sampler.summary(),
            "gpu_launches": launches,
            "impl": "b200",
        }
        line.update(extra)
        if comm:
            line["comm"] = comm
        if e2e:
            line["e2e"] = e2e
        if cpu:
            line["cpu_baseline"] = cpu
        _REAL_STDOUT.write(json.dumps(line) + "\n")
        _REAL_STDOUT.flush()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def _peak():
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    return (float(peaks.get("hbm_gbs", 6650.0)),
            "measured (MEASURED_PEAKS.json hbm_gbs)" if peaks else "fallback 6650 GB/s (B200_PROFILING.md)")


if __name__ == "__main__":
    main()
