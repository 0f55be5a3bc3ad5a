"""GPU parity of rcsb_ik_constraint_rmr against the reference's IkSolverConstraintRMR
(src/RcsCore/IkSolverConstraintRMR.cpp:156-262, joint-limit active set): states sampled over the whole
joint range, so a good part of the steps ends within 2 degrees of a limit and is solved a second time
with joint constraints appended."""
import numpy as np
import pytest

import rcs_b200
from rcs_b200 import workload
from conftest import assert_close, graph_file
from test_ik_gpu import assert_step_close
from test_ik_tasks_gpu import _spec

pytestmark = pytest.mark.gpu


def _expected_violations(g, names, q0, dq_plain):
    """joints the unconstrained step leaves inside the 2 degree margin (IkSolverConstraintRMR.cpp:183-208)"""
    lay = g.layout()
    own = {t[1] for t in names if t[0] == "Joint"}
    margin = (np.pi / 180.0) * 2.0
    n = np.zeros(q0.shape[0], dtype=int)
    for j, jd in enumerate(lay["joints"]):
        if jd["constrained"] or jd["coupledTo"] >= 0 or jd["name"] in own:
            continue
        x = q0[:, j] + dq_plain[:, j]
        n += ((x < jd["q_min"] + margin) | (x > jd["q_max"] - margin)).astype(int)
    return n


@pytest.mark.parametrize("name,names,lam", [
    ("wam_arm", [("XYZ", "BallTip")], 1.0e-6),
    ("wam_arm", [("XYZ", "BallTip"), ("ABC", "BallTip")], 1.0e-4),
    ("wam_arm", [("Joint", "ArmJoint2"), ("XYZ", "BallTip", "Link2")], 1.0e-6),
    ("gantry", [("XYZ", "Tool"), ("POLARZ", "Tool")], 1.0e-6),
    ("humanoid", [("XYZ", "RightHandTip"), ("XYZ", "LeftHandTip"), ("COGX", "Heel"), ("COGY", "Heel")], 1.0e-6),
], ids=["arm-xyz", "arm-pose-overdetermined", "arm-own-joint-task", "gantry", "humanoid"])
def test_constraint_solver_step(name, names, lam, refmod, cuda_device):
    g = rcs_b200.Graph(graph_file(name))
    m = rcs_b200.Model(g, cuda_device)
    rik = refmod.RefIk(graph_file(name), names)
    n = 400 if name != "humanoid" else 128
    q_goal = workload.sample_states(g.layout(), n, first=5, margin=0.0)
    x_des, _ = rik.task_kinematics(q_goal)
    rng = np.random.default_rng(11)
    q0 = np.ascontiguousarray(q_goal + 0.05 * rng.uniform(-1, 1, size=q_goal.shape))
    x_des = np.ascontiguousarray(x_des)
    _, _, dq_plain, _ = rik.rmr(q0, x_des, lam=lam)
    q_ref, dx_ref, dq_ref, det_ref = rik.rmr(q0, x_des, lam=lam, constraint=True)
    q = q0.copy()
    out = m.ik_tasks_rmr(_spec(g, names), q, x_des, lam=lam, want=("dx", "dq", "det", "n_violations"),
                         constraint=True)
    nv = _expected_violations(g, names, q0, dq_plain)
    # a joint whose step ends within rounding of the margin may be classified either way
    lay = g.layout()
    margin = (np.pi / 180.0) * 2.0
    edge = np.zeros(n, dtype=bool)
    for j, jd in enumerate(lay["joints"]):
        x = q0[:, j] + dq_plain[:, j]
        edge |= (np.abs(x - (jd["q_min"] + margin)) < 1e-9) | (np.abs(x - (jd["q_max"] - margin)) < 1e-9)
    ok = ~edge & (out["n_violations"] >= 0)
    assert ok.mean() > 0.95
    assert np.array_equal(out["n_violations"][ok].astype(int), nv[ok])
    assert (nv > 0).mean() > 0.05, "the batch must exercise the second solve"
    assert_close(out["dx"], dx_ref, what="dx")
    # conditioning of the system that was solved last: task rows plus the constraint rows
    _, J = rik.task_kinematics(q0)
    w = rik.graph.inv_wq()
    cond = np.empty(n)
    col = {j: jd["jacobiIndex"] for j, jd in enumerate(lay["joints"])}
    own = {t[1] for t in names if t[0] == "Joint"}
    for i in range(n):
        rows = [J[i]]
        for j, jd in enumerate(lay["joints"]):
            if jd["constrained"] or jd["coupledTo"] >= 0 or jd["name"] in own:
                continue
            x = q0[i, j] + dq_plain[i, j]
            if x < jd["q_min"] + margin or x > jd["q_max"] - margin:
                e = np.zeros((1, J.shape[2]))
                e[0, col[j]] = 1.0
                rows.append(e)
        Ji = np.concatenate(rows, axis=0)
        cond[i] = np.linalg.cond((Ji * w[None, :]) @ Ji.T + lam * np.eye(Ji.shape[0]))
    assert_step_close(out["dq"][ok], dq_ref[ok], cond[ok], "dq")
    assert_step_close((q - q0)[ok], (q_ref - q0)[ok], cond[ok], "q")
    rel = np.abs(out["det"][ok] / det_ref[ok] - 1.0)
    assert np.all(rel <= 1e-11 * np.maximum(1.0, cond[ok] / 100.0)), "det"


def test_constraint_solver_iterates_and_keeps_joints_inside(refmod, cuda_device):
    """five iterations from device pointers; without violations the result is rcsb_ik_tasks_rmr's"""
    import torch

    names = [("XYZ", "BallTip")]
    g = rcs_b200.Graph(graph_file("wam_arm"))
    m = rcs_b200.Model(g, cuda_device)
    rik = refmod.RefIk(graph_file("wam_arm"), names)
    n, lam = 256, 1.0e-6
    q_goal = workload.sample_states(g.layout(), n, first=77, margin=0.0)
    x_des, _ = rik.task_kinematics(q_goal)
    q0 = np.ascontiguousarray(q_goal + 0.03)
    q_ref = q0.copy()
    for _ in range(5):
        q_ref, _, _, _ = rik.rmr(q_ref, np.ascontiguousarray(x_des), lam=lam, constraint=True)
    qd = torch.from_numpy(q0.copy()).cuda(cuda_device)
    xd = torch.from_numpy(np.ascontiguousarray(x_des)).cuda(cuda_device)
    m.ik_tasks_rmr(_spec(g, names), qd, xd, lam=lam, iters=5, want=(), constraint=True)
    torch.cuda.synchronize()
    q = qd.cpu().numpy()
    good = np.abs(q - q_ref).max(axis=1) <= 1e-9
    assert good.mean() > 0.9      # states that flip a violation test by rounding take another branch
    # states well inside their limits: identical to the unconstrained solver
    inner = workload.sample_states(g.layout(), 64, first=3, margin=0.3)
    xi, _ = rik.task_kinematics(inner)
    qa = np.ascontiguousarray(inner + 0.01)
    qb = qa.copy()
    oa = m.ik_tasks_rmr(_spec(g, names), qa, np.ascontiguousarray(xi), lam=lam, want=("dq",))
    ob = m.ik_tasks_rmr(_spec(g, names), qb, np.ascontiguousarray(xi), lam=lam, want=("dq", "n_violations"),
                        constraint=True)
    assert np.all(ob["n_violations"] == 0)
    assert np.array_equal(oa["dq"], ob["dq"]) and np.array_equal(qa, qb)
