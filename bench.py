#!/usr/bin/env python
"""bench.py — headline benchmark of rcs_b200 (contract: see the task description / DESIGN.md).

Workload at every N (weak scaling, one process per GPU): BASELINE.json configs[1],
"7-DoF arm batched FK + Jacobians, 1M uniformly sampled joint states, FP64":
  per state  in : q (7 doubles)
             out: A_BI of all 12 bodies (12 x 96 B) + JT, JR of BallTip (2 x 3 x 7 doubles)
             = 1544 algorithmic bytes.
One step = one pass of the hot path over one batch of 1,048,576 states per GPU
(one fkKernel launch per rank) + one NCCL allgather of per-rank summary statistics (N > 1).

  value  states/s, inputs resident in HBM, CUDA-event timed, max over ranks
  e2e    states/s through the host-pointer C ABI (pinned host buffers; H2D of q and D2H of
         every result inside the timed region)
  roofline.achieved = 1544 B x states per launch / mean launch duration
  cpu_baseline / --impl reference: the reference library (oracle/_ref, unmodified Rcs
         sources) on the host cores, one RcsGraph_clone per thread.
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

N_STATES = 1 << 20
GRAPH = os.path.join(ROOT, "tests", "golden", "graphs", "wam_arm.xml")
EFFECTOR = "BallTip"
BYTES_PER_STATE = 7 * 8 + 12 * 96 + 2 * 3 * 7 * 8   # 1544
METRIC = "FK+Jacobian states/sec FP64 (7-DoF arm, 1M states per GPU)"
UNIT = "states/s"


def host_threads() -> int:
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def workload_config(n_gpus: int) -> dict:
    return {
        "workload": "configs[1]: WAM 7-DoF arm (dof 7, nJ 7, 12 bodies), FK all body frames + "
                    "JT/JR of BallTip, uniformly sampled joint states",
        "states_per_gpu": N_STATES,
        "global_states": N_STATES * n_gpus,
        "bytes_per_state": BYTES_PER_STATE,
        "parallelism": f"dp{n_gpus} (independent state slices; allgather of summary statistics)",
        "l2": "per-step working set 1.6 GB per GPU >> 126 MB L2 (no flush needed)",
    }


# ---------------------------------------------------------------------------------------------
# reference arm / CPU baseline
# ---------------------------------------------------------------------------------------------

def reference_rate(sample_states: int, repeats: int, threads: int):
    """FK + Jacobians of `sample_states` states with the reference library, `repeats` times.
    Returns (states/s of the best repeat, seconds list)."""
    from oracle import ref
    from rcs_b200 import workload

    g = ref.RefGraph(GRAPH)
    eff = [g.body_index(EFFECTOR)]
    q = workload.sample_states(g.layout(), sample_states)
    ref.set_threads(threads)
    A = np.empty((sample_states, g.n_bodies, 12))
    JT = np.empty((sample_states, 1, 3, g.nJ))
    JR = np.empty((sample_states, 1, 3, g.nJ))
    import ctypes

    L = ref.lib()
    effarr = (ctypes.c_int * 1)(*eff)

    def once():
        return L.ref_fk_jacobians(g._h, sample_states, g.dof, q.ctypes.data, 1, effarr, None,
                                  A.ctypes.data, JT.ctypes.data, JR.ctypes.data)

    secs = [once() for _ in range(repeats)]
    return sample_states / min(secs), secs


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import ref

    if not ref.available():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/librcsref.so missing"}))
        return
    threads = host_threads()
    # calibrate so that one step stays around a second of wall time
    rate, _ = reference_rate(1 << 14, 1, threads)
    sample = int(min(N_STATES, max(1 << 14, rate * 1.0)))
    sample = 1 << (sample.bit_length() - 1)
    for _ in range(args.warmup):
        reference_rate(sample, 1, threads)
    t0 = time.perf_counter()
    _, secs = reference_rate(sample, args.steps, threads)
    wall = time.perf_counter() - t0
    total = sum(secs)
    value = sample * args.steps / total
    line = {
        "impl": "reference",
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": workload_config(args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "reference",
                         "sample": f"{sample} states per step x {args.steps} steps, one "
                                   f"RcsGraph_clone per thread, {threads} threads "
                                   f"(wall incl. thread start {wall:.2f} s)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------

class ClockSampler:
    def __init__(self, device_index: int):
        self.samples = []
        self.stop = threading.Event()
        self.idx = device_index
        self.thread = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        while not self.stop.is_set():
            try:
                out = subprocess.run(
                    ["nvidia-smi", f"--id={self.idx}", f"--query-gpu={q}",
                     "--format=csv,noheader,nounits"],
                    capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
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
        sm, smax, reasons = [], 0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            try:
                sm.append(float(s[0]))
                smax = max(smax, float(s[1]))
                for n, v in zip(names, s[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except (ValueError, IndexError):
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax or None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------

def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (burst copy)"
    except Exception:
        return 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"


def ncu_traffic():
    p = os.path.join(ROOT, "profiles", "fk_traffic.json")
    try:
        with open(p) as f:
            return json.load(f).get("dram_bytes_per_launch")
    except Exception:
        return None


def run_ours(args):
    import torch
    import torch.distributed as dist

    import rcs_b200
    from rcs_b200 import workload

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: rcs_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    g = rcs_b200.Graph(GRAPH)
    m = rcs_b200.Model(g, local)
    eff = [g.body_index(EFFECTOR)]
    n = N_STATES
    # every rank owns its own slice of the global, counter-based state sequence
    q_host = workload.sample_states(g.layout(), n, first=rank * n)
    q = torch.from_numpy(q_host).to(dev)
    out = {
        "A_BI": torch.empty((n, g.n_bodies, 12), dtype=torch.float64, device=dev),
        "JT": torch.empty((n, 1, 3, g.nJ), dtype=torch.float64, device=dev),
        "JR": torch.empty((n, 1, 3, g.nJ), dtype=torch.float64, device=dev),
    }
    stats = torch.zeros(4, dtype=torch.float64, device=dev)
    gathered = torch.zeros(4 * world, dtype=torch.float64, device=dev)
    tip = eff[0]

    def step():
        m.fk_jacobians(q, eff, out=out)
        if world > 1:
            # summary statistics of this rank's slice: mean effector position + |JT| checksum
            stats[:3] = out["A_BI"][:: 4096, tip, :3].mean(dim=0)
            stats[3] = out["JT"][:: 4096].abs().sum()
            dist.all_gather_into_tensor(gathered, stats)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()

    launches0 = rcs_b200.kernel_launch_count()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
           for _ in range(args.steps)]
    with ClockSampler(local) as clocks:
        barrier()
        ev[0].record()
        for i in range(args.steps):
            kev[i][0].record()
            m.fk_jacobians(q, eff, out=out)
            kev[i][1].record()
            if world > 1:
                stats[:3] = out["A_BI"][:: 4096, tip, :3].mean(dim=0)
                stats[3] = out["JT"][:: 4096].abs().sum()
                dist.all_gather_into_tensor(gathered, stats)
            ev[i + 1].record()
        barrier()
        launches = rcs_b200.kernel_launch_count() - launches0
        # keep the GPU busy a little longer so that short runs still get clock samples under load
        t_hold = time.perf_counter()
        while time.perf_counter() - t_hold < args.hold:
            m.fk_jacobians(q, eff, out=out)
            torch.cuda.synchronize()
    total_ms = ev[0].elapsed_time(ev[-1])
    kern_ms = [a.elapsed_time(b) for a, b in kev]

    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    value = n * world * args.steps / (total_ms * 1e-3)

    # ---- end-to-end through the host-pointer ABI (pinned buffers) ---------------------------
    hq = torch.empty((n, g.dof), dtype=torch.float64).pin_memory()
    hq.copy_(torch.from_numpy(q_host))
    hout = {
        "A_BI": torch.empty((n, g.n_bodies, 12), dtype=torch.float64).pin_memory(),
        "JT": torch.empty((n, 1, 3, g.nJ), dtype=torch.float64).pin_memory(),
        "JR": torch.empty((n, 1, 3, g.nJ), dtype=torch.float64).pin_memory(),
    }
    e2e_steps = max(2, min(args.steps, 5))
    m.fk_jacobians(hq, eff, out=hout)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        m.fk_jacobians(hq, eff, out=hout)
    t1 = time.perf_counter()
    te = torch.tensor([t1 - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = n * world * e2e_steps / float(te.item())
    # the pinned results must equal the device results of the same states
    same = bool(torch.equal(hout["JT"][:1000], out["JT"][:1000].cpu()))

    if rank == 0:
        peak, peak_src = measured_peak_gbs()
        k_ms = float(np.mean(kern_ms))
        achieved = BYTES_PER_STATE * n / (k_ms * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": workload_config(world),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": ncu_traffic(),
                         "kernel": "rcsb::fkKernel", "kernel_ms": k_ms, "peak_source": peak_src,
                         "bytes_per_launch": BYTES_PER_STATE * n},
            "e2e": {"value": e2e_value, "unit": UNIT,
                    "h2d_bytes_per_step": n * g.dof * 8 * world,
                    "d2h_bytes_per_step": n * (g.n_bodies * 12 + 2 * 3 * g.nJ) * 8 * world,
                    "steps": e2e_steps, "results_match_device_path": same},
            "gpu_launches": int(launches),
            "clocks": clocks.summary(),
        }
        if world == 1 and not args.no_cpu_baseline:
            from oracle import ref

            if ref.available():
                threads = host_threads()
                rate1, _ = reference_rate(1 << 16, 2, 1)
                rateN, secs = reference_rate(N_STATES, 3, threads)
                line["cpu_baseline"] = {
                    "value": rateN, "unit": UNIT, "cores": threads, "kind": "reference",
                    "single_core_value": rate1,
                    "sample": f"reference library (unmodified Rcs sources), {N_STATES} states x 3 "
                              f"passes on {threads} threads (best pass), one RcsGraph_clone per "
                              f"thread; single core: 65536 states x 2"}
            else:
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "reference",
                                        "sample": "oracle/_ref/librcsref.so missing"}
        print(json.dumps(line))

    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
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
