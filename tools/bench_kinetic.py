#!/usr/bin/env python
"""Times rcsb_kinetic_terms on BASELINE.json configs[2]: humanoid (dof 65, nJ 34, 63 bodies),
256K states, M + h + F_gravity + E, FP64, device-resident inputs. Prints one JSON line.

    python tools/bench_kinetic.py [--graph humanoid] [--states 262144] [--steps 5]
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
    ap.add_argument("--graph", default="humanoid")
    ap.add_argument("--states", type=int, default=1 << 18)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--m-only", action="store_true", help="mass matrix only, no velocities")
    ap.add_argument("--dynamics", action="store_true",
                    help="time rcsb_direct_dynamics (qdd, E; M stays on the SM) instead")
    args = ap.parse_args()

    g = rcs_b200.Graph(os.path.join(ROOT, "tests", "golden", "graphs", args.graph + ".xml"))
    m = rcs_b200.Model(g, 0)
    n = args.states
    q, qd = workload.sample_states(g.layout(), n, with_velocity=True)
    dq, dqd = torch.from_numpy(q).cuda(), torch.from_numpy(qd).cuda()
    out = {
        "M": torch.empty((n, g.nJ, g.nJ), dtype=torch.float64, device="cuda"),
        "h": torch.empty((n, g.nJ), dtype=torch.float64, device="cuda"),
        "Fg": torch.empty((n, g.nJ), dtype=torch.float64, device="cuda"),
        "E": torch.empty((n,), dtype=torch.float64, device="cuda"),
    }
    if args.m_only:
        out = {"M": out["M"]}

        def call():
            m.kinetic_terms(dq, out=out)
    elif args.dynamics:
        out = {"qdd": torch.empty((n, g.nJ), dtype=torch.float64, device="cuda"), "E": out["E"]}

        def call():
            m.direct_dynamics(dq, dqd, out=out)
    else:
        def call():
            m.kinetic_terms(dq, dqd, out=out)
    for _ in range(args.warmup):
        call()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        call()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    bytes_per_state = 2 * g.dof * 8 + (g.nJ * g.nJ + 2 * g.nJ + 1) * 8
    if args.m_only:
        bytes_per_state = g.dof * 8 + g.nJ * g.nJ * 8
    if args.dynamics:
        bytes_per_state = 2 * g.dof * 8 + (g.nJ + 1) * 8
    print(json.dumps({
        "workload": f"{args.graph} " + ("direct dynamics qdd,E" if args.dynamics else "mass matrix only" if args.m_only else "kinetic terms M,h,Fg,E"), "states": n, "ms": ms,
        "states_per_s": n / (ms * 1e-3), "bytes_per_state": bytes_per_state,
        "GBps": bytes_per_state * n / (ms * 1e-3) / 1e9,
        "blocks_per_sm": os.environ.get("RCSB_KT_BLOCKS_PER_SM", "auto")}))


if __name__ == "__main__":
    main()
