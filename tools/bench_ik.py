#!/usr/bin/env python
"""Times rcsb_ik_rmr on BASELINE.json configs[3]: WAM 7-DoF arm, XYZ + ABC tasks on BallTip,
lambda 1e-8, alpha 0.05, 10 iterations per target, device-resident inputs. Prints one JSON line.

    python tools/bench_ik.py [--states 4194304] [--iters 10] [--steps 5]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

import rcs_b200  # noqa: E402
from rcs_b200 import TASK_ABC, TASK_XYZ, workload  # noqa: E402


def euler_xyz(R9: torch.Tensor) -> torch.Tensor:
    """Input preparation only (x-y-z Euler angles of the goal frames, Rcs_Mat3d.c:956)."""
    A = R9.reshape(-1, 3, 3)
    cy = torch.sqrt(A[:, 0, 0] ** 2 + A[:, 1, 0] ** 2)
    return torch.stack([-torch.atan2(A[:, 2, 1], A[:, 2, 2]), -torch.atan2(-A[:, 2, 0], cy),
                        -torch.atan2(A[:, 1, 0], A[:, 0, 0])], dim=1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--states", type=int, default=1 << 22)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--graph", default="wam_arm")
    ap.add_argument("--effector", default="BallTip")
    ap.add_argument("--effector2", default=None, help="second effector (XYZ+ABC as well): nx = 12")
    ap.add_argument("--left", action="store_true", help="IkSolverRMR::solveLeftInverse, lambda 1e-3")
    args = ap.parse_args()

    g = rcs_b200.Graph(os.path.join(ROOT, "tests", "golden", "graphs", args.graph + ".xml"))
    m = rcs_b200.Model(g, 0)
    tip = g.body_index(args.effector)
    tasks = [(TASK_XYZ, tip), (TASK_ABC, tip)]
    effs = [tip]
    if args.effector2:
        tip2 = g.body_index(args.effector2)
        tasks += [(TASK_XYZ, tip2), (TASK_ABC, tip2)]
        effs.append(tip2)
    n = args.states
    lay = g.layout()
    q_goal = torch.from_numpy(workload.sample_states(lay, n, first=1 << 24, margin=0.2)).cuda()
    frames = m.fk(q_goal, bodies=effs)["A_BI"]
    x_des = torch.cat([torch.cat([frames[:, i, :3], euler_xyz(frames[:, i, 3:])], dim=1)
                       for i in range(len(effs))], dim=1).contiguous()
    u = torch.from_numpy(workload.uniform01(workload.SEED, 0, n, g.dof, stream=7)).cuda()
    q_start = (q_goal + 0.1 * (2.0 * u - 1.0)).contiguous()
    q = q_start.clone()
    det = torch.empty(n, dtype=torch.float64, device="cuda")
    for _ in range(args.warmup):
        q.copy_(q_start)
        m.ik_rmr(tasks, q, x_des, iters=args.iters, out={"det": det}, left_inverse=args.left,
                 lam=1.0e-3 if args.left else 1.0e-8)
    torch.cuda.synchronize()
    ms = []
    for _ in range(args.steps):
        q.copy_(q_start)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        m.ik_rmr(tasks, q, x_des, iters=args.iters, out={"det": det}, left_inverse=args.left,
                 lam=1.0e-3 if args.left else 1.0e-8)
        e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    t = float(np.mean(ms))
    print(json.dumps({"workload": f"{args.graph}/{args.effector}{'+' + args.effector2 if args.effector2 else ''} RMR IK XYZ+ABC, {args.iters} iterations", "targets": n,
                      "ms": t, "solves_per_s": n / (t * 1e-3),
                      "iterations_per_s": n * args.iters / (t * 1e-3),
                      "bytes_per_solve": 56 + 48 + 56 + 8,
                      "singular_fraction": float((det == 0).double().mean())}))


if __name__ == "__main__":
    main()
