#!/usr/bin/env python
"""bench.py — benchmark of rcs_b200 (contract: task description / DESIGN.md §6).

The default run prints ONE JSON line. Its top-level keys are the headline workload `fkm`:
BASELINE.json's metric "FK+Jacobian+M(q) states/sec FP64" on the arm and batch of configs[1]
("7-DoF arm batched FK + Jacobians, 1M uniformly sampled joint states"), the configuration
north_star's target is quoted on:
  per state  in : q (7 doubles)
             out: A_BI of all 12 bodies (12 x 96 B) + JT, JR of BallTip (2 x 3 x 7 doubles)
                  + M(q) (7 x 7 doubles)
             = 1936 algorithmic bytes; one fkChainKernel<MASS = 2, NC = 7> launch per step and rank.
Under "secondary" the same line carries one complete record (value, roofline, e2e, cpu_baseline,
clocks, launches) for each other configuration of BASELINE.json, measured in the same process
right after the headline:
  ik       configs[3], the second half of the metric ("RMR IK solves/sec"): 7-DoF arm, XYZ + ABC
           tasks on BallTip, lambda 1e-8, alpha 0.05, 10 RMR iterations, 4,194,304 targets in
           total (strong scaling: sharded over the ranks).
  kinetic  configs[2]: humanoid (dof 65, nJ 34, 63 bodies), 262,144 states per GPU,
           M + h + F_gravity + E; 10,840 algorithmic bytes per state.
  wholebody configs[4]: humanoid, FK of all 63 bodies + the 20 task-Jacobian rows of the
           reference's cAction.xml + M(q), 1,048,576 states per GPU; N > 1 adds the all-gather
           of both hand-tip frames of every state (192 B per state) to the summary statistics.
`--workload X` measures X alone (fk = configs[1] literally, without M(q)).

One step = one pass of the hot path over one batch + (N > 1) one NCCL all-gather of per-rank
summary statistics, the only collective of the path.
  value     units/s with inputs resident in HBM, CUDA-event timed on the launching stream,
            barrier + synchronize on both sides, max over ranks
  e2e       the same through the host-pointer C ABI (RCSB_HOST_PTRS): pinned host buffers,
            H2D of the inputs and D2H of every result inside the timed region
  roofline  algorithmic bytes (or FP64 flops) per launch / mean kernel time of the timed steps
  cpu_baseline / --impl reference: the unmodified reference library (oracle/_ref) on the
            host cores, one RcsGraph_clone (+ controller and solver) per thread, cloned before
            the clock starts, one RcsGraph_setState per state.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

GRAPHS = os.path.join(ROOT, "tests", "golden", "graphs")
UNIT = {"fk": "states/s", "fkm": "states/s", "kinetic": "states/s", "ik": "solves/s", "wholebody": "states/s"}
SECONDARY = ("ik", "kinetic", "wholebody")
# active task set of the reference's GenericHumanoid/cAction.xml (nx = 20): hand positions, COG x / y,
# both feet as 6-D poses relative to the Heel
WB_TASKS = [("XYZ", "RightHandTip"), ("XYZ", "LeftHandTip"), ("COGX", "Heel"), ("COGY", "Heel"),
            ("XYZABC", "FootL", "Heel"), ("XYZABC", "FootR", "Heel")]
METRIC = {
    "fk": "FK+Jacobian states/sec FP64 (7-DoF arm, 1M states per GPU)",
    "fkm": "FK+Jacobian+M(q) states/sec FP64 (7-DoF arm, 1M states per GPU)",
    "kinetic": "kinetic terms M,h,g states/sec FP64 (humanoid, 256K states per GPU)",
    "ik": "RMR IK solves/sec FP64 (7-DoF arm, XYZ+ABC, 10 iterations, 4M targets)",
    "wholebody": "whole-body FK + all-task Jacobians + M(q) states/sec FP64 (humanoid, 1M states per GPU)",
}
IK_ITERS, IK_LAMBDA, IK_ALPHA = 10, 1.0e-8, 0.05


def host_threads() -> int:
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


class Workload:
    """Sizes, byte counts and descriptions of one BASELINE.json configuration."""

    def __init__(self, name: str, world: int):
        self.name = name
        self.world = world
        if name == "fk":
            self.graph, self.eff = "wam_arm", "BallTip"
            self.per_gpu = 1 << 20
            self.scaling = "weak"
            self.bytes_per_unit = 7 * 8 + 12 * 96 + 2 * 3 * 7 * 8            # 1544
            self.kernel = "rcsb::fkChainKernel<MASS = 0, NC = 7> (serial chain of rotations: straight-line walk)"
            self.desc = ("configs[1] without M(q): WAM 7-DoF arm (dof 7, nJ 7, 12 bodies), FK all body frames "
                         "+ JT/JR of BallTip, uniformly sampled joint states")
            self.l2 = "per-step working set 1.6 GB per GPU >> 126 MB L2 (no flush needed)"
        elif name == "fkm":
            self.graph, self.eff = "wam_arm", "BallTip"
            self.per_gpu = 1 << 20
            self.scaling = "weak"
            self.bytes_per_unit = 7 * 8 + 12 * 96 + 2 * 3 * 7 * 8 + 7 * 7 * 8  # 1936
            self.kernel = "rcsb::fkChainKernel<MASS = 2, NC = 7> (serial chain of rotations: straight-line walk, M tip to root by inverse transforms)"
            self.desc = ("configs[1] + M(q), the configuration the metric and north_star's target are quoted "
                         "on: WAM 7-DoF arm (dof 7, nJ 7, 12 bodies), FK all 12 body frames + JT/JR of "
                         "BallTip + mass matrix M(q) from one launch (rcsb_fk_jacobians_mass), uniformly "
                         "sampled joint states")
            self.l2 = "per-step working set 2.0 GB per GPU >> 126 MB L2 (no flush needed)"
        elif name == "kinetic":
            self.graph, self.eff = "humanoid", None
            self.per_gpu = 1 << 18
            self.scaling = "weak"
            self.bytes_per_unit = 2 * 65 * 8 + (34 * 34 + 2 * 34 + 1) * 8     # 10840
            self.kernel = "rcsb::ktWalkKernel<VEL> + rcsb::ktAssembleKernel<VEL>"
            self.desc = ("configs[2]: GenericHumanoid (dof 65, nJ 34, 63 bodies), "
                         "RcsGraph_computeKineticTerms M, h, F_gravity, E; q and q_dot uniformly sampled")
            self.l2 = "per-step working set 2.8 GB per GPU >> 126 MB L2 (no flush needed)"
        elif name == "ik":
            self.graph, self.eff = "wam_arm", "BallTip"
            total = 1 << 22
            self.per_gpu = total // world
            self.scaling = "strong"
            self.bytes_per_unit = 56 + 48 + 56 + 8                            # q in, x_des, q out, det
            self.kernel = "rcsb::ikChainKernel<7, 6, LEFT = 0, MODE = 1>"
            self.desc = ("configs[3]: WAM 7-DoF arm, IkSolverRMR right inverse, tasks XYZ + ABC on BallTip, "
                         f"lambda {IK_LAMBDA}, alpha {IK_ALPHA}, {IK_ITERS} iterations per target; targets = "
                         "poses of uniformly sampled states, start = goal state + U(-0.1, 0.1) rad")
            self.l2 = "per-step working set 0.7 GB in total >> 126 MB L2 (no flush needed)"
        elif name == "wholebody":
            self.graph, self.eff = "humanoid", None
            self.per_gpu = 1 << 20
            self.scaling = "weak"
            # q + 63 frames + 20 task rows x 34 + M
            self.bytes_per_unit = 65 * 8 + 63 * 96 + 20 * 34 * 8 + 34 * 34 * 8   # 21256
            self.kernel = "rcsb::ktWalkKernel<WB> (frames + records) + rcsb::ktAssembleKernel (M, COG and task Jacobians)"
            self.desc = ("configs[4]: GenericHumanoid (dof 65, nJ 34, 63 bodies), FK of all 63 bodies + the task "
                         "Jacobians of cAction.xml's active set (hands XYZ, COG x/y, both feet XYZABC relative to "
                         "the Heel: nx 20) + mass matrix; N > 1: all-gather of both hand-tip frames (192 B per state)")
            self.l2 = "per-step working set 22 GB per GPU >> 126 MB L2 (no flush needed)"
        else:
            raise SystemExit(f"unknown workload {name}")
        self.graph_file = os.path.join(GRAPHS, self.graph + ".xml")

    def config(self):
        return {
            "workload": self.desc,
            "units_per_gpu": self.per_gpu,
            "global_units": self.per_gpu * self.world,
            "bytes_per_unit": self.bytes_per_unit,
            "parallelism": f"dp{self.world} (independent state slices; all-gather of summary statistics)",
            "l2": self.l2,
        }


# ---------------------------------------------------------------------------------------------
# inputs (identical arrays for the GPU path and the CPU reference)
# ---------------------------------------------------------------------------------------------

def euler_xyz(A9: np.ndarray) -> np.ndarray:
    """x-y-z Euler angles of row-major rotations (input preparation, Rcs_Mat3d.c:956)."""
    A = A9.reshape(-1, 3, 3)
    cy = np.sqrt(A[:, 0, 0] ** 2 + A[:, 1, 0] ** 2)
    return np.stack([-np.arctan2(A[:, 2, 1], A[:, 2, 2]), -np.arctan2(-A[:, 2, 0], cy),
                     -np.arctan2(A[:, 1, 0], A[:, 0, 0])], axis=1)


def ik_inputs(layout, model_fk, first: int, count: int, dof: int, tip: int):
    """(q_start, x_des) for `count` targets starting at global index `first`. model_fk maps a
    q array to the effector frames (GPU model or reference graph)."""
    from rcs_b200 import workload

    q_goal = workload.sample_states(layout, count, first=(1 << 24) + first, margin=0.2)
    frames = model_fk(q_goal)
    x_des = np.ascontiguousarray(np.concatenate([frames[:, :3], euler_xyz(frames[:, 3:])], axis=1))
    u = workload.uniform01(workload.SEED, first, count, dof, stream=7)
    q_start = np.ascontiguousarray(q_goal + 0.1 * (2.0 * u - 1.0))
    return q_start, x_des


# ---------------------------------------------------------------------------------------------
# reference arm / CPU baseline
# ---------------------------------------------------------------------------------------------

class ReferenceRunner:
    """The reference's own functions (oracle/_ref: unmodified Rcs sources) over a sample: per
    state ONE RcsGraph_setState followed by the reference functions of the workload; every thread
    clones its graph / controller before the clock starts (oracle/ref_build/ref_api.cpp)."""

    def __init__(self, wl: Workload, threads: int):
        from oracle import ref

        self.ref, self.wl, self.threads = ref, wl, threads
        self.g = ref.RefGraph(wl.graph_file)
        self.layout = self.g.layout()
        if wl.name == "ik":
            self.ik = ref.RefIk(wl.graph_file, [("XYZ", wl.eff), ("ABC", wl.eff)])
        if wl.name == "wholebody":
            self.ik = ref.RefIk(wl.graph_file, WB_TASKS)
        self.sample = 0

    def prepare(self, sample: int):
        from rcs_b200 import workload

        wl, g = self.wl, self.g
        self.sample = sample
        if wl.name in ("fk", "fkm"):
            self.q = workload.sample_states(self.layout, sample)
            self.eff = [g.body_index(wl.eff)]
        elif wl.name == "wholebody":
            self.q = workload.sample_states(self.layout, sample)
        elif wl.name == "kinetic":
            self.q, self.qd = workload.sample_states(self.layout, sample, with_velocity=True)
        else:
            tip = g.body_index(wl.eff)
            self.ref.set_threads(self.threads)
            self.q, self.x_des = ik_inputs(self.layout, lambda q: g.fk(q)["A_BI"][:, tip], 0, sample,
                                           g.dof, tip)

    def run(self) -> float:
        """One pass over the sample on self.threads threads; returns seconds inside the library."""
        self.ref.set_threads(self.threads)
        wl, g = self.wl, self.g
        if wl.name == "fk":
            g.fk_jacobians(self.q, self.eff)
            return g.last_seconds
        if wl.name == "fkm":
            # RcsGraph_setState -> bodyPointJacobian / rotationJacobian -> RcsGraph_computeMassMatrix
            g.fk_jacobians(self.q, self.eff, want=("A_BI", "JT", "JR", "M"))
            return g.last_seconds
        if wl.name == "kinetic":
            g.kinetic_terms(self.q, self.qd)
            return g.last_seconds
        if wl.name == "wholebody":
            # RcsGraph_setState (frames of every body), ControllerBase::computeJ, RcsGraph_computeMassMatrix
            self.ik.wholebody(self.q)
            return self.ik.last_seconds
        self.ik.rmr(self.q, self.x_des, lam=IK_LAMBDA, alpha=IK_ALPHA, iters=IK_ITERS)
        return self.ik.last_seconds

    def calibrate(self, target_seconds: float, cap: int) -> int:
        """Sample size (power of two) whose pass takes about target_seconds."""
        n = {"fk": 1 << 14, "fkm": 1 << 13, "kinetic": 1 << 8, "ik": 1 << 12, "wholebody": 1 << 8}[self.wl.name]
        self.prepare(n)
        rate = n / max(self.run(), 1e-9)
        want = int(min(cap, max(n, rate * target_seconds)))
        return 1 << (want.bit_length() - 1)


def reference_record(name: str, args, step_seconds: float):
    """`steps` timed passes of the reference over a bounded sample of workload `name`."""
    wl = Workload(name, args.gpus)
    threads = host_threads()
    runner = ReferenceRunner(wl, threads)
    sample = runner.calibrate(step_seconds, wl.per_gpu * wl.world)
    runner.prepare(sample)
    for _ in range(args.warmup):
        runner.run()
    t0 = time.perf_counter()
    secs = [runner.run() for _ in range(args.steps)]
    wall = time.perf_counter() - t0
    total = sum(secs)
    value = sample * args.steps / total
    return {
        "impl": "reference",
        "metric": METRIC[wl.name], "value": value, "unit": UNIT[wl.name], "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": wl.scaling, "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": wl.config(),
        "cpu_baseline": {"value": value, "unit": UNIT[wl.name], "cores": threads, "kind": "reference",
                         "sample": f"{sample} units per step x {args.steps} steps, one RcsGraph_clone per "
                                   f"thread made before the clock starts, one RcsGraph_setState per state, "
                                   f"{threads} threads (wall incl. thread start and cloning {wall:.2f} s)"},
        "e2e": {"value": value, "unit": UNIT[wl.name], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import ref

    if not ref.available():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/librcsref.so missing"}))
        return
    line = reference_record(args.workload, args, 1.0)
    if args.workload == "fkm" and not args.no_secondary:
        line["secondary"] = {name: reference_record(name, args, 0.4) for name in SECONDARY}
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------

class ClockSampler:
    """SM clock and throttle reasons of one GPU, sampled on a thread while the GPU works: through
    NVML every 2 ms where pynvml is importable (several samples even inside a 10 ms timed region),
    else by polling nvidia-smi (one sample per ~60 ms)."""

    def __init__(self, device_index: int):
        self.samples = []          # (perf_counter, sm MHz, max MHz, [reasons])
        self.stop = threading.Event()
        self.idx = device_index
        self.source = "nvidia-smi"
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.marks = {}

    def mark(self, name: str):
        self.marks[name] = time.perf_counter()

    def _run_nvml(self) -> bool:
        try:
            import pynvml

            pynvml.nvmlInit()
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = self.idx
            if visible:
                ids = [v for v in visible.split(",") if v.strip() != ""]
                if idx < len(ids) and ids[idx].strip().isdigit():
                    idx = int(ids[idx])
            h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            smax = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            bits = {
                "hw_slowdown": getattr(pynvml, "nvmlClocksEventReasonHwSlowdown", 0x8),
                "hw_thermal_slowdown": getattr(pynvml, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                "sw_thermal_slowdown": getattr(pynvml, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                "sw_power_cap": getattr(pynvml, "nvmlClocksEventReasonSwPowerCap", 0x4),
            }
            pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
        except Exception:
            return False
        self.source = "nvml"
        while not self.stop.is_set():
            try:
                sm = float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
                r = int(pynvml.nvmlDeviceGetCurrentClocksEventReasons(h))
                self.samples.append((time.perf_counter(), sm, smax, [n for n, b in bits.items() if r & b]))
            except Exception:
                pass
            self.stop.wait(0.002)
        return True

    def _run(self):
        if self._run_nvml():
            return
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self.stop.is_set():
            try:
                out = subprocess.run(
                    ["nvidia-smi", f"--id={self.idx}", f"--query-gpu={q}",
                     "--format=csv,noheader,nounits"],
                    capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    s = [x.strip() for x in out.split(",")]
                    self.samples.append((time.perf_counter(), float(s[0]), float(s[1]),
                                         [n for n, v in zip(names, s[2:6]) if v.lower().startswith("active")]))
            except Exception:
                pass
            self.stop.wait(0.05)

    def __enter__(self):
        self.thread.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        self.thread.join(timeout=10)

    def summary(self):
        """Median SM clock and the union of the throttle reasons over ALL samples (settle loop,
        warm-up, timed steps and the hold loop: identical launches back to back), plus the same
        for the samples that fell between the two barriers of the timed region."""
        def agg(rows):
            sm = [r[1] for r in rows]
            reasons = sorted({n for r in rows for n in r[3]})
            return (float(np.median(sm)) if sm else None), reasons

        sm, reasons = agg(self.samples)
        smax = max([r[2] for r in self.samples], default=0.0)
        t0, t1 = self.marks.get("timed_begin"), self.marks.get("timed_end")
        timed = [r for r in self.samples if t0 is not None and t1 is not None and t0 <= r[0] <= t1]
        sm_t, reasons_t = agg(timed)
        return {"sm_mhz": sm, "sm_max_mhz": smax or None, "reasons": reasons, "samples": len(self.samples),
                "source": self.source,
                "covers": "settle loop + warm-up + timed steps + hold loop (the same launches back to back)",
                "timed_region": {"samples": len(timed), "sm_mhz": sm_t, "reasons": reasons_t,
                                 "seconds": (t1 - t0) if timed or (t0 and t1) else None}}


# ---------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------

def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (burst copy)"
    except Exception:
        return 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"


def profile_record(name: str):
    """Per-launch DRAM traffic / FP64 instruction counts taken from the committed ncu captures."""
    try:
        with open(os.path.join(ROOT, "profiles", "kernel_counters.json")) as f:
            return json.load(f).get(name, {})
    except Exception:
        return {}


_FP64_PEAK = None


def fp64_peak_tflops():
    """DFMA microbenchmark (tools/fp64_peak.cu) run live; falls back to the committed figure."""
    global _FP64_PEAK
    if _FP64_PEAK is None:
        exe = os.path.join(ROOT, "build", "fp64_peak")
        try:
            out = subprocess.run([exe], capture_output=True, text=True, timeout=60).stdout
            _FP64_PEAK = (float(json.loads(out)["fp64_tflops"]), "build/fp64_peak DFMA microbenchmark, this run")
        except Exception:
            _FP64_PEAK = (34.2, "tools/fp64_peak.cu measured on B200 in round 1 (34.2 TFLOP/s)")
    return _FP64_PEAK


class Ctx:
    """Process-wide state of our arm: rank, device, process group."""

    def __init__(self):
        import torch
        import torch.distributed as dist

        self.torch, self.dist = torch, dist
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device: rcs_b200 has no CPU fallback")
        # host side of the end-to-end path: run this rank's threads and place its pinned buffers
        # next to its GPU where the container allows it (rcsb_bind_host_to_device)
        import rcs_b200

        self.numa = rcs_b200.bind_host_to_device(self.local)
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, x: float) -> float:
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())


def measure_ours(name: str, args, ctx: Ctx, cpu_seconds: float = 4.0):
    """One workload on this rank's GPU; returns the record on rank 0, None elsewhere."""
    import rcs_b200
    from rcs_b200 import TASK_ABC, TASK_XYZ, shard, workload

    torch = ctx.torch
    rank, world, local, dev = ctx.rank, ctx.world, ctx.local, ctx.dev
    wl = Workload(name, world)
    g = rcs_b200.Graph(wl.graph_file)
    m = rcs_b200.Model(g, local)
    lay = g.layout()
    n = wl.per_gpu
    first = rank * n                     # every rank owns its slice of the global sequence
    f64 = torch.float64

    def pinned(shape):
        return rcs_b200.pinned_empty(shape)

    def dev_empty(shape):
        return torch.empty(shape, dtype=f64, device=dev)

    e2e_units = n
    # ---- workload-specific buffers and the two call forms (device pointers / host pointers) ----
    if wl.name in ("fk", "fkm"):
        eff = [g.body_index(wl.eff)]
        q_host = workload.sample_states(lay, n, first=first)
        q = torch.from_numpy(q_host).to(dev)
        shapes = {"A_BI": (n, g.n_bodies, 12), "JT": (n, 1, 3, g.nJ), "JR": (n, 1, 3, g.nJ)}
        if wl.name == "fkm":
            shapes["M"] = (n, g.nJ, g.nJ)
        out = {k: dev_empty(s) for k, s in shapes.items()}
        hq = pinned((n, g.dof))
        hq.copy_(torch.from_numpy(q_host))
        hout = {k: pinned(s) for k, s in shapes.items()}
        h2d, d2h = n * g.dof * 8, sum(int(np.prod(sh[1:])) for sh in shapes.values()) * n * 8

        def device_call():
            m.fk_jacobians(q, eff, out=out)

        def host_call():
            m.fk_jacobians(hq, eff, out=hout)

        def stat_source():
            return out["A_BI"][:: 1024, eff[0], 2]          # effector height, strided sample

        def same():
            return bool(torch.equal(hout["JT"][:1000], out["JT"][:1000].cpu()))
    elif wl.name == "wholebody":
        rh, lh, heel, fl, fr = [g.body_index(x) for x in ("RightHandTip", "LeftHandTip", "Heel", "FootL", "FootR")]
        spec = [("POS", rh, -1, -1), ("POS", lh, -1, -1), ("POS", fl, heel, heel), ("OMEGA", fl, heel, heel),
                ("POS", fr, heel, heel), ("OMEGA", fr, heel, heel)]
        q_host = workload.sample_states(lay, n, first=first)
        q = torch.from_numpy(q_host).to(dev)
        shapes = {"A_BI": (n, g.n_bodies, 12), "Jt": (n, len(spec), 3, g.nJ), "M": (n, g.nJ, g.nJ),
                  "Jcog": (n, 3, g.nJ)}
        out = {k: dev_empty(sh) for k, sh in shapes.items()}
        # end to end on a bounded slice (pinning 22 GB of host memory is not part of the path)
        ne = min(n, 1 << 17)
        hq = pinned((ne, g.dof))
        hq.copy_(torch.from_numpy(q_host[:ne]))
        hout = {k: pinned((ne,) + sh[1:]) for k, sh in shapes.items()}
        h2d, d2h = ne * g.dof * 8, sum(int(np.prod(sh[1:])) for sh in shapes.values()) * ne * 8
        e2e_units = ne
        hands = torch.tensor([rh, lh], device=dev)
        compact = dev_empty((n, 2, 12))
        gathered_frames = dev_empty((world * n, 2, 12)) if world > 1 else None

        def device_call():
            m.wholebody(q, spec, out=out)

        def host_call():
            m.wholebody(hq, spec, out=hout)

        def stat_source():
            return out["A_BI"][:: 1024, rh, 2]

        def same():
            return bool(torch.equal(hout["Jt"][:500], out["Jt"][:500].cpu()) and
                        torch.equal(hout["M"][:500], out["M"][:500].cpu()))
    elif wl.name == "kinetic":
        q_host, qd_host = workload.sample_states(lay, n, first=first, with_velocity=True)
        q, qd = torch.from_numpy(q_host).to(dev), torch.from_numpy(qd_host).to(dev)
        shapes = {"M": (n, g.nJ, g.nJ), "h": (n, g.nJ), "Fg": (n, g.nJ), "E": (n,)}
        out = {k: dev_empty(s) for k, s in shapes.items()}
        hq, hqd = pinned((n, g.dof)), pinned((n, g.dof))
        hq.copy_(torch.from_numpy(q_host))
        hqd.copy_(torch.from_numpy(qd_host))
        hout = {k: pinned(s) for k, s in shapes.items()}
        h2d, d2h = 2 * n * g.dof * 8, n * (g.nJ * g.nJ + 2 * g.nJ + 1) * 8

        def device_call():
            m.kinetic_terms(q, qd, out=out)

        def host_call():
            m.kinetic_terms(hq, hqd, out=hout)

        def stat_source():
            return out["E"][:: 256]

        def same():
            return bool(torch.equal(hout["h"][:1000], out["h"][:1000].cpu()))
    else:
        tip = g.body_index(wl.eff)
        tasks = [(TASK_XYZ, tip), (TASK_ABC, tip)]

        def gpu_frames(qa):
            return m.fk(torch.from_numpy(qa).to(dev), bodies=[tip])["A_BI"][:, 0].cpu().numpy()

        q_start_host, x_host = ik_inputs(lay, gpu_frames, first, n, g.dof, tip)
        q_start, x_des = torch.from_numpy(q_start_host).to(dev), torch.from_numpy(x_host).to(dev)
        q = q_start.clone()
        out = {"det": dev_empty((n,))}
        hq0, hq, hx = pinned((n, g.dof)), pinned((n, g.dof)), pinned((n, 6))
        hq0.copy_(torch.from_numpy(q_start_host))
        hx.copy_(torch.from_numpy(x_host))
        hout = {"det": pinned((n,))}
        h2d, d2h = n * (g.dof + 6) * 8, n * (g.dof + 1) * 8

        def device_call():
            q.copy_(q_start)                                 # restore the start states (untimed kernel-wise)
            m.ik_rmr(tasks, q, x_des, lam=IK_LAMBDA, alpha=IK_ALPHA, iters=IK_ITERS, out=out)

        # the solver works in place on q: every end-to-end step gets its own pinned copy of the start
        # states, made before the clock starts (a 235 MB host-to-host copy per step is the caller's
        # business, not the library's; it cost 6 of 16 ms when it sat inside the timed region)
        e2e_calls = max(2, min(args.steps, 5)) + 1
        hq_fresh = [hq] + [pinned((n, g.dof)) for _ in range(e2e_calls - 1)]
        for t in hq_fresh:
            t.copy_(hq0)
        hq_turn = [0]

        def host_call():
            cur = hq_fresh[hq_turn[0] % len(hq_fresh)]
            hq_turn[0] += 1
            m.ik_rmr(tasks, cur, hx, lam=IK_LAMBDA, alpha=IK_ALPHA, iters=IK_ITERS, out=hout)

        def stat_source():
            return out["det"][:: 1024]

        def same():
            return bool(torch.equal(hq[:1000], q[:1000].cpu()))

    # per-step slots, so that the all-gather of step i can overlap the kernel of step i + 1
    n_slots = args.steps + args.warmup + 1
    stats = torch.zeros((n_slots, 5), dtype=f64, device=dev)
    gathered = torch.zeros((n_slots, 5 * world), dtype=f64, device=dev)
    pending = []

    def exchange(slot):
        """One librcsb reduction kernel on the compute stream, then the NCCL all-gather of its 5
        doubles, asynchronous: it waits for the reduction only."""
        if world > 1:
            rcs_b200.summary_stats(stat_source(), out=stats[slot])
            _, work = shard.all_gather_stats(stats[slot], out=gathered[slot], async_op=True)
            pending.append(work)
            if wl.name == "wholebody":
                # configs[4]: all-gather of per-state results, the two hand-tip frames of every state
                torch.index_select(out["A_BI"], 1, hands, out=compact)
                pending.append(shard.all_gather_results(compact, out=gathered_frames, async_op=True)[1])

    def drain():
        while pending:
            pending.pop().wait()

    barrier = ctx.barrier
    with ClockSampler(local) as clocks:
        # bring the SM clocks up before the warm-up steps: a step is 0.5 ms, so a short --warmup
        # alone would be timed on a GPU that is still ramping from idle (measured: 1.87 instead of
        # 2.05 G states/s with --warmup 3 --steps 5)
        t_settle = time.perf_counter()
        while time.perf_counter() - t_settle < args.settle:
            device_call()
            torch.cuda.synchronize()
        for i in range(args.warmup):
            device_call()
            exchange(args.steps + i)
        drain()
        barrier()

        # the IK solver works in place: every timed step gets its own copy of the start states, made
        # before the clock starts (inputs are resident when the timed region begins; up to 64 steps,
        # 235 MB each for the 4M-target batch - beyond that the states are restored inside the loop)
        q_steps = None
        if wl.name == "ik" and args.steps <= 64:
            q_steps = [q_start.clone() for _ in range(args.steps)]
            torch.cuda.synchronize()
        launches0 = rcs_b200.kernel_launch_count()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
        kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
               for _ in range(args.steps)]
        barrier()
        clocks.mark("timed_begin")
        ev[0].record()
        for i in range(args.steps):
            if wl.name == "ik" and q_steps is None:
                q.copy_(q_start)
            kev[i][0].record()
            if wl.name == "ik":
                m.ik_rmr(tasks, q if q_steps is None else q_steps[i], x_des, lam=IK_LAMBDA, alpha=IK_ALPHA,
                         iters=IK_ITERS, out=out)
            else:
                device_call()
            kev[i][1].record()
            exchange(i)
            ev[i + 1].record()
        drain()                          # every all-gather has completed before the clock stops
        end = torch.cuda.Event(enable_timing=True)
        end.record()
        barrier()
        clocks.mark("timed_end")
        launches = rcs_b200.kernel_launch_count() - launches0
        q_steps = None                   # release the per-step copies
        # keep the GPU busy a little longer so that short runs still get clock samples under load
        t_hold = time.perf_counter()
        while time.perf_counter() - t_hold < args.hold:
            device_call()
            torch.cuda.synchronize()
    total_ms = ev[0].elapsed_time(end)
    kern_ms = [a.elapsed_time(b) for a, b in kev]
    total_ms = ctx.max_over_ranks(total_ms)
    value = n * world * args.steps / (total_ms * 1e-3)

    # ---- end-to-end through the host-pointer ABI (pinned buffers) ---------------------------
    e2e_steps = max(2, min(args.steps, 5))
    host_call()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        host_call()
    t1 = time.perf_counter()
    e2e_seconds = ctx.max_over_ranks(t1 - t0)
    e2e_value = e2e_units * world * e2e_steps / e2e_seconds
    device_call()
    torch.cuda.synchronize()
    results_match = same()               # pinned results equal the device results of the same states

    line = None
    if rank == 0:
        k_ms = float(np.mean(kern_ms))
        counters = profile_record(wl.name)
        # the committed capture may cover fewer units than one step (whole-body: one pass of the
        # kernel pair over 414,336 states): scale its DRAM bytes to the units of a step
        if counters.get("dram_bytes_per_launch") and counters.get("units_per_launch"):
            counters = dict(counters)
            counters["dram_bytes_per_launch"] = counters["dram_bytes_per_launch"] * n / counters["units_per_launch"]
        traffic_note = ("dram__bytes_read.sum + dram__bytes_write.sum of the committed `ncu --set full` captures "
                        "of this workload's kernels (profiles/kernel_counters.json, profiles/r02_*), per step; "
                        "not measured in this run")
        if wl.name == "ik":
            peak, peak_src = fp64_peak_tflops()
            flops = counters.get("fp64_flops_per_unit")
            achieved = (flops * n / (k_ms * 1e-3) / 1e12) if flops else None
            roof = {"bound": "fp64", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                    "frac": (achieved / peak) if achieved else None,
                    "traffic": counters.get("dram_bytes_per_launch"),
                    "flops_per_unit": flops,
                    "flops_source": "ncu FP64 instruction counts of one launch (DFMA x 2 + DMUL + DADD), "
                                    "profiles/kernel_counters.json",
                    "hbm_frac": wl.bytes_per_unit * n / (k_ms * 1e-3) / 1e9 / measured_peaks()[0]}
        else:
            peak, peak_src = measured_peaks()
            achieved = wl.bytes_per_unit * n / (k_ms * 1e-3) / 1e9
            roof = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": achieved / peak, "traffic": counters.get("dram_bytes_per_launch")}
        roof.update({"kernel": counters.get("kernel_names", wl.kernel), "kernel_ms": k_ms, "peak_source": peak_src,
                     "algorithmic_per_launch": wl.bytes_per_unit * n, "traffic_source": traffic_note})
        line = {
            "metric": METRIC[wl.name], "value": value, "unit": UNIT[wl.name], "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps,
            "higher_is_better": True, "scaling": wl.scaling, "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": wl.config(),
            "roofline": roof,
            "e2e": {"value": e2e_value, "unit": UNIT[wl.name], "h2d_bytes_per_step": h2d * world,
                    "d2h_bytes_per_step": d2h * world, "steps": e2e_steps, "units_per_step": e2e_units * world,
                    "results_match_device_path": results_match,
                    "host_gbs_per_gpu": (h2d + d2h) * e2e_steps / e2e_seconds / 1e9,
                    "host_binding": ctx.numa},
            "gpu_launches": int(launches),
            "clocks": clocks.summary(),
        }
        if world == 1 and not args.no_cpu_baseline:
            from oracle import ref

            if ref.available():
                threads = host_threads()
                runner = ReferenceRunner(wl, threads)
                sample = runner.calibrate(cpu_seconds, n)
                runner.prepare(sample)
                secs = [runner.run() for _ in range(3)]
                one = ReferenceRunner(wl, 1)
                s1 = one.calibrate(1.0, n)
                one.prepare(s1)
                t_one = one.run()
                line["cpu_baseline"] = {
                    "value": sample / min(secs), "unit": UNIT[wl.name], "cores": threads,
                    "kind": "reference", "single_core_value": s1 / t_one,
                    "sample": f"reference library (unmodified Rcs sources), {sample} units x 3 passes on "
                              f"{threads} threads (best pass), one RcsGraph_clone per thread made before the "
                              f"clock starts, one RcsGraph_setState per state; single core: {s1} units x 1"}
            else:
                line["cpu_baseline"] = {"value": None, "unit": UNIT[wl.name], "cores": 0,
                                        "kind": "reference", "sample": "oracle/_ref/librcsref.so missing"}
    # free this workload's buffers before the next one is set up
    del out, hout
    torch.cuda.empty_cache()
    return line


def measure_group_abi(args, ctx: Ctx):
    """The same workloads through the device-group entry points of the C ABI (rcsb_group_*): ONE
    process (rank 0) drives all N GPUs with host arrays of the whole batch; the library slices,
    runs one host thread + stream + model per device and all-gathers the per-rank statistics over
    NCCL itself. End to end by construction (host buffers in, results back). The other ranks wait
    on the CPU meanwhile."""
    import rcs_b200
    from rcs_b200 import TASK_ABC, TASK_XYZ, workload

    torch = ctx.torch
    world = ctx.world
    out = {"devices": world, "api": "rcsb_group_fk_jacobians_mass / rcsb_group_ik_rmr (include/rcsb.h)"}
    steps = max(2, min(args.steps, 4))

    # fkm: 256K states per device (host arrays of the whole batch: 0.5 GB of pinned memory per device)
    g = rcs_b200.Graph(Workload("fkm", world).graph_file)
    grp = rcs_b200.Group(g, list(range(world)))
    n = (1 << 18) * world
    tip = g.body_index("BallTip")
    q = rcs_b200.pinned_empty((n, g.dof))
    q.copy_(torch.from_numpy(workload.sample_states(g.layout(), n)))
    res = {"A_BI": rcs_b200.pinned_empty((n, g.n_bodies, 12)), "JT": rcs_b200.pinned_empty((n, 1, 3, g.nJ)),
           "JR": rcs_b200.pinned_empty((n, 1, 3, g.nJ)), "M": rcs_b200.pinned_empty((n, g.nJ, g.nJ))}
    grp.fk_jacobians(q, [tip], res)
    t0 = time.perf_counter()
    for _ in range(steps):
        st = grp.fk_jacobians(q, [tip], res)
    dt = time.perf_counter() - t0
    bytes_per_state = 8 * (g.dof + g.n_bodies * 12 + 6 * g.nJ + g.nJ * g.nJ)
    out["fkm"] = {"value": n * steps / dt, "unit": "states/s", "units_per_step": n, "steps": steps,
                  "host_gbs_total": n * steps * bytes_per_state / dt / 1e9,
                  "gathered_stats_counts": [float(x) for x in st[:, 0]]}
    del res, q

    # ik: 4M targets in total (strong scaling)
    n = 1 << 22
    m0 = rcs_b200.Model(g, ctx.local)
    lay = g.layout()

    def gpu_frames(qa):
        return m0.fk(torch.from_numpy(qa).to(ctx.dev), bodies=[tip])["A_BI"][:, 0].cpu().numpy()

    q_start, x_des = ik_inputs(lay, gpu_frames, 0, n, g.dof, tip)
    hq0, hq, hx = rcs_b200.pinned_empty((n, g.dof)), rcs_b200.pinned_empty((n, g.dof)), rcs_b200.pinned_empty((n, 6))
    hq0.copy_(torch.from_numpy(q_start))
    hx.copy_(torch.from_numpy(x_des))
    o = {"det": rcs_b200.pinned_empty((n,))}
    tasks = [(TASK_XYZ, tip), (TASK_ABC, tip)]
    hq.copy_(hq0)
    grp.ik_rmr(tasks, hq, hx, o, lam=IK_LAMBDA, alpha=IK_ALPHA, iters=IK_ITERS)
    secs = 0.0
    for _ in range(steps):
        hq.copy_(hq0)
        t0 = time.perf_counter()
        st = grp.ik_rmr(tasks, hq, hx, o, lam=IK_LAMBDA, alpha=IK_ALPHA, iters=IK_ITERS)
        secs += time.perf_counter() - t0
    out["ik"] = {"value": n * steps / secs, "unit": "solves/s", "units_per_step": n, "steps": steps,
                 "gathered_stats_counts": [float(x) for x in st[:, 0]]}
    return out


def run_ours(args):
    ctx = Ctx()
    line = measure_ours(args.workload, args, ctx)
    secondary = {}
    if args.workload == "fkm" and not args.no_secondary:
        for name in SECONDARY:
            secondary[name] = measure_ours(name, args, ctx, cpu_seconds=2.0)
        # device groups of the C ABI: rank 0 alone drives every GPU; the other ranks wait on the
        # CPU (a gloo barrier: no kernel of theirs spins on the GPUs rank 0 is using)
        cpu_group = ctx.dist.new_group(backend="gloo") if ctx.world > 1 else None
        ctx.torch.cuda.synchronize()
        group_rec = None
        if ctx.rank == 0:
            try:
                group_rec = measure_group_abi(args, ctx)
            except Exception as e:      # the headline must not depend on this record
                group_rec = {"error": f"{type(e).__name__}: {e}"}
        if cpu_group is not None:
            ctx.dist.barrier(group=cpu_group)
        if group_rec is not None:
            secondary["group_abi"] = group_rec
    if ctx.rank == 0:
        if secondary:
            line["secondary"] = secondary
        print(json.dumps(line))
    if ctx.world > 1:
        ctx.dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="fkm", choices=["fkm", "fk", "kinetic", "ik", "wholebody"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true",
                    help="default workload only: skip the ik / kinetic / wholebody records")
    ap.add_argument("--settle", type=float, default=1.0,
                    help="seconds of untimed work before the warm-up steps (clock ramp from idle)")
    ap.add_argument("--hold", type=float, default=1.0,
                    help="seconds of extra (untimed) load for the clock sampler")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
