#!/usr/bin/env python
"""Times fkKernel variants on the 7-DoF arm, 1M states: all frames + Jacobians (configs[1]),
the minimal "effector frame + Jacobians" record (488 B), frames only, and FK with
velocities. Prints one JSON line per variant with the fraction of the measured HBM peak."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import rcs_b200  # noqa: E402
from rcs_b200 import workload  # noqa: E402


def timed(fn, steps=20, warmup=5):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def main():
    graph = sys.argv[1] if len(sys.argv) > 1 else "wam_arm"
    effn = sys.argv[2] if len(sys.argv) > 2 else "BallTip"
    n = int(sys.argv[3]) if len(sys.argv) > 3 else 1 << 20
    g = rcs_b200.Graph(os.path.join(ROOT, "tests", "golden", "graphs", graph + ".xml"))
    m = rcs_b200.Model(g, 0)
    tip = g.body_index(effn)
    q_h, qd_h = workload.sample_states(g.layout(), n, with_velocity=True)
    q, qd = torch.from_numpy(q_h).cuda(), torch.from_numpy(qd_h).cuda()
    peak = 6544.3
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    nb, nJ, dof = g.n_bodies, g.nJ, g.dof

    def buf(*shape):
        return torch.empty(shape, dtype=torch.float64, device="cuda")

    variants = []
    o1 = {"A_BI": buf(n, nb, 12), "JT": buf(n, 1, 3, nJ), "JR": buf(n, 1, 3, nJ)}
    variants.append(("all frames + JT/JR", dof * 8 + nb * 96 + 6 * nJ * 8, lambda: m.fk_jacobians(q, [tip], out=o1)))
    o2 = {"A_BI": buf(n, 1, 12), "JT": buf(n, 1, 3, nJ), "JR": buf(n, 1, 3, nJ)}
    variants.append(("effector frame + JT/JR", dof * 8 + 96 + 6 * nJ * 8,
                     lambda: m.fk_jacobians(q, [tip], bodies=[tip], out=o2)))
    o3 = {"A_BI": buf(n, nb, 12)}
    variants.append(("all frames", dof * 8 + nb * 96, lambda: m.fk(q, out=o3)))
    o4 = {"A_BI": buf(n, nb, 12), "xdot": buf(n, nb, 3), "omega": buf(n, nb, 3)}
    variants.append(("all frames + velocities", 2 * dof * 8 + nb * (96 + 48),
                     lambda: m.fk(q, qd, want=("A_BI", "xdot", "omega"), out=o4)))
    for name, bytes_per_state, fn in variants:
        ms = timed(fn)
        gbs = bytes_per_state * n / (ms * 1e-3) / 1e9
        print(json.dumps({"graph": graph, "variant": name, "states": n, "ms": ms,
                          "states_per_s": n / (ms * 1e-3), "bytes_per_state": bytes_per_state,
                          "GBps": gbs, "frac_of_measured_hbm_peak": gbs / peak}))


if __name__ == "__main__":
    main()
