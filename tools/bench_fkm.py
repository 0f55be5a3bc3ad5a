#!/usr/bin/env python
"""Times the fused "FK + Jacobians + M(q)" launch (rcsb_fk_jacobians_mass) on the 7-DoF arm:
BASELINE.json's target configuration (1M states, all 12 body frames, JT/JR of BallTip, M).
Prints one JSON line with the fraction of the measured HBM peak.

    python tools/bench_fkm.py [--states 1048576] [--steps 20]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import rcs_b200  # noqa: E402
from rcs_b200 import workload  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--states", type=int, default=1 << 20)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    args = ap.parse_args()
    g = rcs_b200.Graph(os.path.join(ROOT, "tests", "golden", "graphs", "wam_arm.xml"))
    m = rcs_b200.Model(g, 0)
    eff = [g.body_index("BallTip")]
    n = args.states
    q = torch.from_numpy(workload.sample_states(g.layout(), n)).cuda()
    shapes = {"A_BI": (n, g.n_bodies, 12), "JT": (n, 1, 3, g.nJ), "JR": (n, 1, 3, g.nJ), "M": (n, g.nJ, g.nJ)}
    out = {k: torch.empty(s, dtype=torch.float64, device="cuda") for k, s in shapes.items()}
    for _ in range(args.warmup):
        m.fk_jacobians(q, eff, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        m.fk_jacobians(q, eff, out=out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    bytes_per_state = 7 * 8 + 12 * 96 + 2 * 3 * 7 * 8 + 49 * 8
    peak = 6544.3
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    gbs = bytes_per_state * n / (ms * 1e-3) / 1e9
    print(json.dumps({"workload": "wam_arm FK (12 frames) + JT/JR(BallTip) + M(q), one launch", "states": n,
                      "ms": ms, "states_per_s": n / (ms * 1e-3), "bytes_per_state": bytes_per_state,
                      "GBps": gbs, "frac_of_measured_hbm_peak": gbs / peak}))


if __name__ == "__main__":
    main()
