#!/usr/bin/env python
"""How much of an IK batch meets the FLAT 1e-12 bar, per kernel (run on the GPU box).

For every IK kernel of rcsb_ik.cu this script compares rcsb_ik_rmr with the reference's
ControllerBase + IkSolverRMR (oracle/_ref) on the target distribution of bench.py's `ik`
workload (BASELINE.json configs[3]: goal poses of uniformly sampled states, start = goal state
+ U(-0.1, 0.1) rad) and reports, after 1 and after 10 iterations,
    the fraction of states with |q_ours - q_ref|_inf <= 1e-12 * max(1, |q_ref - q_start|_inf)
overall and by decile of the condition number of J W J^T + lambda 1 (right inverse) or
J_r^T J_r + lambda 1 (left inverse) — the matrix whose inverse every step goes through.
Writes profiles/r02_ik_error_report.json and .md.

    python tools/ik_error_report.py [--n 16384]
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import rcs_b200  # noqa: E402
from rcs_b200 import TASK_ABC, TASK_XYZ, workload  # noqa: E402
from oracle import ref  # noqa: E402
from conftest import graph_file  # noqa: E402
import bench  # noqa: E402

CODE = {"XYZ": TASK_XYZ, "ABC": TASK_ABC}
BAR = 1.0e-12

# (label, kernel, graph, tasks, left inverse, lambda, force general kernel, states)
CASES = [
    ("arm XYZ+ABC, right inverse", "ikChainKernel<7, 6, LEFT = 0, MODE = 1> (folded chain)", "wam_arm", [("XYZ", "BallTip"), ("ABC", "BallTip")],
     False, 1.0e-8, False, None),
    ("arm XYZ+ABC, right inverse, general kernel", "ikKernel", "wam_arm", [("XYZ", "BallTip"), ("ABC", "BallTip")],
     False, 1.0e-8, True, None),
    ("arm XYZ+ABC, left inverse (lambda 1e-3)", "ikChainKernel<7, 6, LEFT = 1, MODE = 1> (folded chain)", "wam_arm",
     [("XYZ", "BallTip"), ("ABC", "BallTip")], True, 1.0e-3, False, None),
    ("arm XYZ+ABC, left inverse, general kernel", "ikKernel<LEFT>", "wam_arm",
     [("XYZ", "BallTip"), ("ABC", "BallTip")], True, 1.0e-3, True, None),
    ("humanoid right hand XYZ+ABC (13 chain columns)", "ikChainLoopKernel", "humanoid",
     [("XYZ", "RightHandTip"), ("ABC", "RightHandTip")], False, 1.0e-8, False, 2048),
    ("humanoid both hands XYZ (two effectors)", "ikMultiKernel", "humanoid",
     [("XYZ", "RightHandTip"), ("XYZ", "LeftHandTip")], False, 1.0e-8, False, 2048),
]


def condition(rik, q, lam, left):
    _, J = rik.task_kinematics(q)
    w = rik.graph.inv_wq()
    if left:
        Jr = J * w[None, None, :]
        A = np.einsum("nki,nkj->nij", Jr, Jr) + lam * np.eye(J.shape[2])
    else:
        A = np.einsum("nij,j,nkj->nik", J, w, J) + lam * np.eye(J.shape[1])
    return np.linalg.cond(A)


def deciles(cond, within):
    order = np.argsort(cond)
    rows = []
    for d in range(10):
        idx = order[d * len(order) // 10:(d + 1) * len(order) // 10]
        rows.append({"cond_lo": float(cond[idx[0]]), "cond_hi": float(cond[idx[-1]]),
                     "fraction_within_flat_bar": float(within[idx].mean())})
    return rows


def run_case(label, kernel, graph, task_names, left, lam, general, n_states, n_default):
    n = n_states or n_default
    g = rcs_b200.Graph(graph_file(graph))
    m = rcs_b200.Model(g, 0)
    rik = ref.RefIk(graph_file(graph), task_names)
    tasks = [(CODE[k], g.body_index(e)) for k, e in task_names]
    lay = g.layout()
    if graph == "wam_arm":
        tip = g.body_index("BallTip")
        q0, x_des = bench.ik_inputs(lay, lambda qa: rik.graph.fk(qa)["A_BI"][:, tip], 0, n, g.dof, tip)
    else:
        q_goal = workload.sample_states(lay, n, first=(1 << 24), margin=0.2)
        x_des, _ = rik.task_kinematics(q_goal)
        u = workload.uniform01(workload.SEED, 0, n, g.dof, stream=7)
        q0 = np.ascontiguousarray(q_goal + 0.1 * (2.0 * u - 1.0))
        x_des = np.ascontiguousarray(x_des)
    if general:
        os.environ["RCSB_IK_GENERAL"] = "1"
    else:
        os.environ.pop("RCSB_IK_GENERAL", None)
    out = {"case": label, "kernel": kernel, "graph": graph, "states": n, "lambda": lam,
           "left_inverse": left}
    # conditioning along the reference trajectory
    q_ref = q0.copy()
    cond_max = np.zeros(n)
    for it in range(10):
        cond_max = np.maximum(cond_max, condition(rik, q_ref, lam, left))
        if it == 0:
            cond_first = cond_max.copy()
        q_ref, _, _, det = rik.rmr(q_ref, x_des, lam=lam, iters=1, left=left)
        if it == 0:
            q_ref1 = q_ref.copy()
    for iters, qr, cond in ((1, q_ref1, cond_first), (10, q_ref, cond_max)):
        q = q0.copy()
        m.ik_rmr(tasks, q, x_des, lam=lam, iters=iters, left_inverse=left)
        err = np.abs(q - qr).max(axis=1)
        scale = np.maximum(1.0, np.abs(qr - q0).max(axis=1))
        within = err <= BAR * scale
        rel = err / scale
        out[f"iters_{iters}"] = {
            "fraction_within_flat_1e-12": float(within.mean()),
            "rel_err_quantiles_50_90_99_100": [float(x) for x in np.quantile(rel, [0.5, 0.9, 0.99, 1.0])],
            "cond_quantiles_50_90_99_100": [float(x) for x in np.quantile(cond, [0.5, 0.9, 0.99, 1.0])],
            "fraction_within_conditioned_bar": float(
                (err <= BAR * (10.0 if iters == 10 else 1.0) * np.maximum(1.0, cond / 100.0) * scale).mean()),
            "by_cond_decile": deciles(cond, within),
        }
    os.environ.pop("RCSB_IK_GENERAL", None)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=16384)
    args = ap.parse_args()
    ref.set_threads(bench.host_threads())
    results = [run_case(*c, args.n) for c in CASES]
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    with open(os.path.join(ROOT, "profiles", "r02_ik_error_report.json"), "w") as f:
        json.dump(results, f, indent=1)
    lines = ["# IK parity against the reference under the FLAT 1e-12 bar (tools/ik_error_report.py, B200)", "",
             "Fraction of states with `|q - q_ref|_inf <= 1e-12 * max(1, |q_ref - q_start|_inf)` after 1 and 10",
             "RMR iterations; targets and start states as in bench.py's `ik` workload. `cond` is the 2-norm",
             "condition number of the matrix each step inverts (max along the reference trajectory for 10",
             "iterations).", "",
             "| case | kernel | states | 1 it.: within 1e-12 | 1 it.: median / max rel. err | 10 it.: within 1e-12 | "
             "10 it.: median / max rel. err | median / max cond |", "|---|---|---|---|---|---|---|---|"]
    for r in results:
        a, b = r["iters_1"], r["iters_10"]
        lines.append(
            f"| {r['case']} | `{r['kernel']}` | {r['states']} | {a['fraction_within_flat_1e-12']:.4f} | "
            f"{a['rel_err_quantiles_50_90_99_100'][0]:.1e} / {a['rel_err_quantiles_50_90_99_100'][3]:.1e} | "
            f"{b['fraction_within_flat_1e-12']:.4f} | {b['rel_err_quantiles_50_90_99_100'][0]:.1e} / "
            f"{b['rel_err_quantiles_50_90_99_100'][3]:.1e} | {b['cond_quantiles_50_90_99_100'][0]:.1e} / "
            f"{b['cond_quantiles_50_90_99_100'][3]:.1e} |")
    lines += ["", "By decile of cond (10 iterations, fraction within the flat bar):", ""]
    for r in results:
        row = ", ".join(f"[{d['cond_lo']:.0e}..{d['cond_hi']:.0e}] {d['fraction_within_flat_bar']:.3f}"
                        for d in r["iters_10"]["by_cond_decile"])
        lines.append(f"- {r['case']}: {row}")
    with open(os.path.join(ROOT, "profiles", "r02_ik_error_report.md"), "w") as f:
        f.write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main()
