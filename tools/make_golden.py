#!/usr/bin/env python
"""Records golden input/output vectors from the UNMODIFIED reference (oracle/_ref/librcsref.so,
built from /root/reference by oracle/ref_build/Makefile) into tests/golden/*.npz.

The reference itself has no stored vectors for this path, and /root/reference
does not exist on the GPU box; these small fixtures pin the numpy oracle and the CUDA path
there. Re-run in the CPU container after changing graphs or cases:

    python tools/make_golden.py [case ...]
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref  # noqa: E402
from rcs_b200 import workload  # noqa: E402

GRAPHS = os.path.join(ROOT, "tests", "golden", "graphs")
CASES = {
    "wam_arm": dict(n=24, effectors=["BallTip", "Link4"]),
    "wam_scenario": dict(n=16, effectors=["Hand", "PowerGrip"]),
    "humanoid": dict(n=12, effectors=["RightHandTip", "LeftHandTip", "FootL"]),
    "husky": dict(n=8, effectors=None),
    "gantry": dict(n=16, effectors=["Tool", "Camera"]),
    # 5- and 9-coefficient polynomial couplings (Rcs_joint.c:318-462); a third of the states have
    # the master joints below, a third above their limits (tangent extension of the polynomial)
    "polycoupled": dict(n=24, effectors=["Tip", "Probe"], beyond=True),
    # several joint axes per body behind translation-only transforms: the folded chain walks of
    # fkChainKernel / ikChainKernel with row shifts
    "wrist6": dict(n=24, effectors=["Tip", "Fore"]),
}


def push_masters_beyond_limits(lay, q):
    """States 1, 4, 7, ... get every master joint below q_min, states 2, 5, 8, ... above q_max."""
    masters = sorted({j["coupledTo"] for j in lay["joints"] if j["coupledTo"] >= 0})
    q = q.copy()
    for mj in masters:
        lo, hi = lay["joints"][mj]["q_min"], lay["joints"][mj]["q_max"]
        span = hi - lo
        q[1::3, mj] = lo - 0.05 * span - 0.3 * span * np.linspace(0.0, 1.0, len(q[1::3]))
        q[2::3, mj] = hi + 0.05 * span + 0.3 * span * np.linspace(0.0, 1.0, len(q[2::3]))
    return np.ascontiguousarray(q)


def main():
    ref.set_threads(1)
    only = sys.argv[1:]          # optional: names of the cases to (re)write
    for name, case in CASES.items():
        if only and name not in only:
            continue
        f = os.path.join(GRAPHS, name + ".xml")
        g = ref.RefGraph(f)
        lay = g.layout()
        n = case["n"]
        q, qd = workload.sample_states(lay, n, first=1000, with_velocity=True)
        if case.get("beyond"):
            q = push_masters_beyond_limits(lay, q)
        names = [b["name"] for b in lay["bodies"]]
        effs = case["effectors"] or [names[-1], names[len(names) // 2]]
        eff = [names.index(e) for e in effs]
        pts = np.random.default_rng(3).uniform(-0.1, 0.1, size=(len(eff), 3))
        fk = g.fk(q, qd, want=("A_BI", "A_JI", "xdot", "omega", "q"))
        jac = g.fk_jacobians(q, eff, points=pts, want=("JT", "JR"))
        kt = g.kinetic_terms(q, qd)
        cog, jcog = g.cog(q)
        # relative task Jacobians (RcsGraph_3dPosJacobian / 3dOmegaJacobian) and direct dynamics
        a, b, c = eff[0], eff[-1], len(names) // 3
        rel_spec = [("POS", a, -1, -1), ("POS", a, -1, c), ("POS", a, b, b), ("POS", a, b, c),
                    ("POS", -1, b, c), ("OMEGA", a, -1, -1), ("OMEGA", a, b, b), ("OMEGA", a, b, c),
                    ("OMEGA", a, -1, c), ("OMEGA", a, b, -1)]
        rel = g.task_jacobians(q, rel_spec)
        rng = np.random.default_rng(9)
        f_ext, f_jnt = rng.normal(size=(n, g.nJ)), rng.normal(size=(n, g.nJ))
        qdd, _ = g.direct_dynamics(q, qd, f_ext, f_jnt)
        out = dict(q=q, qd=qd, eff=np.array(eff), pts=pts, cog=cog, Jcog=jcog,
                   rel_spec=np.array([[1 if k == "POS" else 2, e, rb, rf] for k, e, rb, rf in rel_spec]),
                   rel_J=rel, dd_Fext=f_ext, dd_Fjnt=f_jnt, dd_qdd=qdd,
                   **{"fk_" + k: v for k, v in fk.items()},
                   **{"jac_" + k: v for k, v in jac.items()},
                   **{"kt_" + k: v for k, v in kt.items()})
        if g.nJ <= 8:
            out["mass_M"] = g.mass_matrix(q)          # RcsGraph_computeMassMatrix (fused FK + J + M launch)
        if name in ("wam_arm", "humanoid", "gantry", "wrist6"):
            tip = {"wam_arm": "BallTip", "humanoid": "RightHandTip", "gantry": "Tool", "wrist6": "Tip"}[name]
            ik = ref.RefIk(f, [("XYZ", tip), ("ABC", tip)])
            q_goal = workload.sample_states(lay, n, first=2000, margin=0.2)
            x_des, J = ik.task_kinematics(q_goal)
            q_start = q_goal + 0.1 * np.random.default_rng(5).uniform(-1, 1, size=q_goal.shape)
            q1, dx1, dq1, det1 = ik.rmr(q_start, x_des, iters=1)
            q10, dx10, dq10, det10 = ik.rmr(q_start, x_des, iters=10)
            if name != "humanoid":
                # IkSolverRMR::solveLeftInverse, lambda 1e-3 (keeps the nq x nq matrix of the redundant
                # arm well conditioned)
                ql, dxl, dql, detl = ik.rmr(q_start, x_des, lam=1.0e-3, iters=1, left=True)
                ql5, _, _, _ = ik.rmr(q_start, x_des, lam=1.0e-3, iters=5, left=True)
                out.update(ikl_q1=ql, ikl_dx1=dxl, ikl_dq1=dql, ikl_det1=detl, ikl_q5=ql5)
            out.update(ik_tip=np.array(names.index(tip)), ik_q_start=q_start, ik_x_des=x_des,
                       ik_J_goal=J, ik_q1=q1, ik_dx1=dx1, ik_dq1=dq1, ik_det1=det1, ik_q10=q10,
                       ik_dx10=dx10, ik_det10=det10)
        path = os.path.join(ROOT, "tests", "golden", name + ".npz")
        np.savez_compressed(path, **out)
        print(f"{name}: {len(out)} arrays, {os.path.getsize(path) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
