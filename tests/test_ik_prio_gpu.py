"""GPU parity of rcsb_ik_prio_rmr against the reference's IkSolverPrioRMR::solve
(src/RcsCore/IkSolverPrioRMR.cpp:178-420): two priority levels and the joint-limit gradient in
the null space of both, looped like examples/ExampleIK.cpp."""
import numpy as np
import pytest

import rcs_b200
from rcs_b200 import TASK_ABC, TASK_XYZ, workload
from conftest import graph_file

pytestmark = pytest.mark.gpu
CODE = {"XYZ": TASK_XYZ, "ABC": TASK_ABC}


def _setup(name, task_names, refmod, dev):
    g = rcs_b200.Graph(graph_file(name))
    m = rcs_b200.Model(g, dev)
    rik = refmod.RefIk(graph_file(name), task_names)
    return g, m, rik, [(CODE[k], g.body_index(e)) for k, e in task_names]


def _targets(g, rik, n, first):
    q_goal = workload.sample_states(g.layout(), n, first=first, margin=0.2)
    x_des, _ = rik.task_kinematics(q_goal)
    rng = np.random.default_rng(first)
    return (np.ascontiguousarray(q_goal + 0.1 * rng.uniform(-1, 1, size=q_goal.shape)),
            np.ascontiguousarray(x_des))


def _conds(rik, q, rows1, rows2, lam):
    """Condition numbers of the two systems the solver inverts: J1 W J1^T + lambda 1 and
    (J2 N1) W (J2 N1)^T + lambda 1 (reference Jacobians, numpy algebra)."""
    _, J = rik.task_kinematics(q)
    w = rik.graph.inv_wq()
    J1, J2 = J[:, rows1], J[:, rows2]
    A1 = np.einsum("nij,j,nkj->nik", J1, w, J1) + lam * np.eye(len(rows1))
    P1 = np.einsum("j,nij,nik->njk", w, J1, np.linalg.inv(A1))
    N1 = np.eye(J.shape[2])[None] - np.einsum("njk,nki->nji", P1, J1)
    J2N1 = np.einsum("nij,njk->nik", J2, N1)
    A2 = np.einsum("nij,j,nkj->nik", J2N1, w, J2N1) + lam * np.eye(len(rows2))
    c2 = np.linalg.cond(A2) if len(rows2) else np.ones(len(q))
    return np.maximum(np.linalg.cond(A1), 1.0) * np.maximum(c2, 1.0)


def _close(ours, ref, cond, what, mult=1.0):
    err = np.abs(ours - ref).max(axis=1)
    scale = np.maximum(1.0, np.abs(ref).max(axis=1))
    tol = 1e-12 * mult * np.maximum(1.0, cond / 100.0) * scale
    assert np.all(np.isfinite(ours)), what
    assert np.all(err <= tol), f"{what}: worst err/tol {float((err / tol).max()):.2f}"


@pytest.mark.parametrize("name,task_names,prio,lam", [
    ("wam_arm", [("XYZ", "BallTip"), ("ABC", "BallTip")], [0, 1], 1.0e-4),
    ("wam_arm", [("XYZ", "BallTip"), ("ABC", "BallTip")], [1, 0], 1.0e-4),
    ("wam_arm", [("XYZ", "BallTip"), ("ABC", "BallTip")], [0, 1], 1.0e-8),
    ("humanoid", [("XYZ", "RightHandTip"), ("XYZ", "LeftHandTip"), ("ABC", "RightHandTip")], [0, 1, 1], 1.0e-6),
    ("gantry", [("XYZ", "Tool"), ("ABC", "Tool")], [0, 1], 1.0e-5),
])
def test_one_prioritised_step(name, task_names, prio, lam, refmod, cuda_device):
    g, m, rik, tasks = _setup(name, task_names, refmod, cuda_device)
    n = 400 if name != "humanoid" else 96
    q0, x_des = _targets(g, rik, n, first=17)
    ref = rik.prio_rmr(q0, x_des, prio, lam=lam)
    q = q0.copy()
    out = m.ik_prio_rmr(tasks, prio, q, x_des, lam=lam)
    rows = np.repeat(np.array(prio), 3)
    cond = _conds(rik, q0, np.where(rows == 0)[0], np.where(rows == 1)[0], lam)
    assert np.array_equal(out["ok"], ref["ok"]) and ref["ok"].all()
    _close(out["dq1"], ref["dq1"], cond, "dq1")
    _close(out["dq2"], ref["dq2"], cond, "dq2")
    _close(out["dq_ns"], ref["dq_ns"], cond, "dq_ns")
    _close(q - q0, ref["q"] - q0, cond, "q")
    assert np.max(np.abs(ref["dq2"])) > 1e-6 and np.max(np.abs(ref["dq_ns"])) > 1e-9


def test_single_level_equals_the_plain_solver(refmod, cuda_device):
    """All tasks on level 0: dq1 is the task-space step of solveRightInverse, dq2 = 0."""
    g, m, rik, tasks = _setup("wam_arm", [("XYZ", "BallTip"), ("ABC", "BallTip")], refmod, cuda_device)
    q0, x_des = _targets(g, rik, 300, first=5)
    qa, qb = q0.copy(), q0.copy()
    a = m.ik_prio_rmr(tasks, [0, 0], qa, x_des)
    b = m.ik_rmr_ex(tasks, qb, x_des, want=("dq_ts", "dq_ns"))
    assert np.max(np.abs(a["dq2"])) == 0.0
    np.testing.assert_allclose(a["dq1"], b["dq_ts"], rtol=0, atol=1e-10)
    np.testing.assert_allclose(a["dq_ns"], b["dq_ns"], rtol=0, atol=1e-10)


def test_activations_ignored_levels_and_iterations(refmod, cuda_device):
    import torch

    names = [("XYZ", "RightHandTip"), ("XYZ", "LeftHandTip"), ("ABC", "RightHandTip")]
    g, m, rik, tasks = _setup("humanoid", names, refmod, cuda_device)
    q0, x_des = _targets(g, rik, 64, first=3)
    prio, act, lam = [0, 1, 2], [1.0, 0.5, 1.0], 1.0e-6      # level 2 is ignored by the solver
    ref = rik.prio_rmr(q0, x_des, prio, activation=act, lam=lam, iters=3)
    dq_, dx_ = torch.from_numpy(q0.copy()).cuda(cuda_device), torch.from_numpy(x_des).cuda(cuda_device)
    out = m.ik_prio_rmr(tasks, prio, dq_, dx_, activation=act, lam=lam, iters=3)
    torch.cuda.synchronize()
    q = dq_.cpu().numpy()
    cond = np.zeros(len(q0))
    qr = q0.copy()
    for _ in range(3):
        cond = np.maximum(cond, _conds(rik, qr, np.arange(0, 3), np.arange(3, 6), lam))
        qr = rik.prio_rmr(qr, x_des, prio, activation=act, lam=lam)["q"]
    _close(q - q0, ref["q"] - q0, cond, "q after 3 iterations", mult=3.0)
    _close(out["dq1"].cpu().numpy(), ref["dq1"], cond, "dq1", mult=3.0)
    # a masked task changes the result: same call with the second task switched off
    off = rik.prio_rmr(q0, x_des, prio, activation=[1.0, 0.0, 1.0], lam=lam)
    q2 = q0.copy()
    o2 = m.ik_prio_rmr(tasks, prio, q2, x_des, activation=[1.0, 0.0, 1.0], lam=lam)
    assert np.max(np.abs(o2["dq2"])) == 0.0 and np.max(np.abs(off["dq2"])) == 0.0
    _close(o2["dq1"], off["dq1"], _conds(rik, q0, np.arange(0, 3), np.arange(0, 0), lam), "dq1, level 1 empty")
