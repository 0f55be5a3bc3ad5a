"""GPU parity of rcsb_ik_rmr_ex — the remaining arguments of the reference's solver interface:
task activations (ControllerBase::computeJ / compressToActive / computeDX with a_des,
src/RcsCore/ControllerBase.cpp:776, :1204, :896), the m x 1 lambda of solveRightInverse
(IkSolverRMR.cpp:385-415, MatNd_rwPinv Rcs_MatNd.c:1833), the dq_ts / dq_ns outputs, and the
Binary / Linear / Approximate task-space blending of the left inverse (TaskSpaceBlender.cpp) —
against the reference's ControllerBase + IkSolverRMR driven with the same arguments."""
import numpy as np
import pytest

import rcs_b200
from rcs_b200 import BLEND_APPROXIMATE, BLEND_BINARY, BLEND_LINEAR, TASK_ABC, TASK_XYZ, TASK_XYZABC, workload
from conftest import assert_close, graph_file
from test_ik_gpu import assert_step_close

pytestmark = pytest.mark.gpu

CODE = {"XYZ": TASK_XYZ, "ABC": TASK_ABC, "XYZABC": TASK_XYZABC}


def _setup(name, task_names, refmod, dev):
    g = rcs_b200.Graph(graph_file(name))
    m = rcs_b200.Model(g, dev)
    rik = refmod.RefIk(graph_file(name), task_names)
    tasks = [(CODE[k], g.body_index(e)) for k, e in task_names]
    return g, m, rik, tasks


def _targets(g, rik, n, first=0, perturb=0.1):
    q_goal = workload.sample_states(g.layout(), n, first=first, margin=0.2)
    x_des, _ = rik.task_kinematics(q_goal)
    rng = np.random.default_rng(3 + first)
    return (np.ascontiguousarray(q_goal + perturb * rng.uniform(-1, 1, size=q_goal.shape)),
            np.ascontiguousarray(x_des))


def _active_condition(rik_active, q, lam_rows):
    """cond of J W J^T + diag(lambda) over the ACTIVE rows (reference Jacobian of the active set)."""
    _, J = rik_active.task_kinematics(q)
    w = rik_active.graph.inv_wq()
    A = np.einsum("nij,j,nkj->nik", J, w, J) + np.diag(lam_rows)
    return np.linalg.cond(A)


ARM = [("XYZ", "BallTip"), ("ABC", "BallTip")]
HUM = [("XYZ", "RightHandTip"), ("ABC", "RightHandTip"), ("XYZ", "LeftHandTip")]


def test_activation_mask_keeps_the_fast_kernel_and_equals_the_reduced_problem(refmod, cuda_device):
    """activation = (1, 0): the ABC task drops out of J and dx; x_des keeps its full layout."""
    g, m, rik, tasks = _setup("wam_arm", ARM, refmod, cuda_device)
    q0, x_des = _targets(g, rik, 1500)
    ref = rik.rmr_ex(q0, x_des, activation=[1.0, 0.0])
    before = rcs_b200.kernel_launch_count()
    q = q0.copy()
    out = m.ik_rmr_ex(tasks, q, x_des, activation=[1.0, 0.0], want=("dx", "dq", "det"))
    assert rcs_b200.kernel_launch_count() - before == 2          # gather + solver
    assert out["dx"].shape == (1500, 3)
    assert_close(out["dx"], ref["dx"][:, :3], what="dx (active rows)")
    rik_xyz = refmod.RefIk(graph_file("wam_arm"), ARM[:1])
    cond = _active_condition(rik_xyz, q0, 1e-8 * np.ones(3))
    assert_step_close(out["dq"], ref["dq_ts"] + ref["dq_ns"], cond, "dq")
    assert_step_close(q - q0, ref["q"] - q0, cond, "q")
    # the same numbers as solving the one-task problem directly
    q2 = q0.copy()
    m.ik_rmr(tasks[:1], q2, np.ascontiguousarray(x_des[:, :3]))
    assert np.array_equal(q, q2)


@pytest.mark.parametrize("scale_dx", [False, True])
@pytest.mark.parametrize("name,task_names,act", [
    ("wam_arm", ARM, [0.5, 1.0]), ("wam_arm", ARM, [1.0, 0.25]), ("humanoid", HUM, [1.0, 0.0, 0.5]),
    ("humanoid", HUM, [0.3, 0.7, 0.0])])
def test_activations_and_split_steps(name, task_names, act, scale_dx, refmod, cuda_device):
    g, m, rik, tasks = _setup(name, task_names, refmod, cuda_device)
    n = 600 if name == "wam_arm" else 200
    q0, x_des = _targets(g, rik, n, first=40)
    ref = rik.rmr_ex(q0, x_des, activation=act, scale_dx=scale_dx)
    q = q0.copy()
    out = m.ik_rmr_ex(tasks, q, x_des, activation=act, scale_dx=scale_dx,
                      want=("dx", "dq", "det", "dq_ts", "dq_ns"))
    nxa = out["dx"].shape[1]
    assert nxa == 3 * sum(a > 0 for a in act)
    assert_close(out["dx"], ref["dx"][:, :nxa], what="dx")
    rik_act = refmod.RefIk(graph_file(name), [t for t, a in zip(task_names, act) if a > 0])
    cond = _active_condition(rik_act, q0, 1e-8 * np.ones(nxa))
    assert_step_close(out["dq_ts"], ref["dq_ts"], cond, "dq_ts")
    assert_step_close(out["dq_ns"], ref["dq_ns"], cond, "dq_ns")
    assert_step_close(out["dq"], ref["dq_ts"] + ref["dq_ns"], cond, "dq")
    assert_step_close(q - q0, ref["q"] - q0, cond, "q")
    rel = np.abs(out["det"] / ref["det"] - 1.0)
    assert np.all(rel <= 1e-11 * np.maximum(1.0, cond / 100.0)), "det"


def test_per_row_lambda(refmod, cuda_device):
    g, m, rik, tasks = _setup("wam_arm", ARM, refmod, cuda_device)
    q0, x_des = _targets(g, rik, 800, first=7)
    lam = np.array([1e-8, 1e-6, 1e-4, 1e-3, 1e-5, 1e-7])
    ref = rik.rmr_ex(q0, x_des, lambda_rows=lam)
    q = q0.copy()
    out = m.ik_rmr_ex(tasks, q, x_des, lambda_rows=lam, want=("dx", "det", "dq_ts", "dq_ns"))
    cond = _active_condition(rik, q0, lam)
    assert_close(out["dx"], ref["dx"], what="dx")
    assert_step_close(out["dq_ts"], ref["dq_ts"], cond, "dq_ts")
    assert_step_close(out["dq_ns"], ref["dq_ns"], cond, "dq_ns")
    assert np.all(np.abs(out["det"] / ref["det"] - 1.0) <= 1e-11 * np.maximum(1.0, cond / 100.0))
    # and it is not the scalar-lambda result
    plain = rik.rmr_ex(q0, x_des)
    assert np.max(np.abs(plain["dq_ts"] - ref["dq_ts"])) > 1e-6


def test_ten_iterations_with_activations_host_and_device_pointers(refmod, cuda_device):
    import torch

    g, m, rik, tasks = _setup("wam_arm", ARM, refmod, cuda_device)
    q0, x_des = _targets(g, rik, 1024, first=99)
    act = [1.0, 0.5]
    ref = rik.rmr_ex(q0, x_des, activation=act, scale_dx=True, iters=10)
    q = q0.copy()
    m.ik_rmr_ex(tasks, q, x_des, activation=act, scale_dx=True, iters=10, want=("det",))
    dq_, dx_ = torch.from_numpy(q0.copy()).cuda(cuda_device), torch.from_numpy(x_des).cuda(cuda_device)
    m.ik_rmr_ex(tasks, dq_, dx_, activation=act, scale_dx=True, iters=10, want=("det",))
    torch.cuda.synchronize()
    assert np.array_equal(dq_.cpu().numpy(), q)
    cond = np.zeros(len(q0))
    qr = q0.copy()
    for _ in range(10):
        cond = np.maximum(cond, _active_condition(rik, qr, 1e-8 * np.ones(6)))
        qr = rik.rmr_ex(qr, x_des, activation=act, scale_dx=True)["q"]
    assert np.array_equal(qr, ref["q"])
    ok = cond < 1e4
    assert ok.mean() > 0.6
    assert_step_close(q[ok] - q0[ok], ref["q"][ok] - q0[ok], 10.0 * cond[ok], "q after 10 iterations")


@pytest.mark.parametrize("blending", [BLEND_BINARY, BLEND_LINEAR, BLEND_APPROXIMATE])
@pytest.mark.parametrize("name,task_names,act", [("wam_arm", ARM, [1.0, 0.3]), ("wam_arm", ARM, [0.6, 1.0]),
                                                 ("gantry", [("XYZ", "Tool"), ("ABC", "Tool")], [0.8, 0.2])])
def test_left_inverse_task_space_blending(name, task_names, act, blending, refmod, cuda_device):
    """solveLeftInverse with TaskSpaceBlender weights: Wx = 1 / a / the closed-form approximate
    weights computed from the rows of the weighted Jacobian (TaskSpaceBlender.cpp:1109-1190)."""
    g, m, rik, tasks = _setup(name, task_names, refmod, cuda_device)
    q0, x_des = _targets(g, rik, 500, first=13)
    lam = 1.0e-3
    ref = rik.rmr_ex(q0, x_des, activation=act, lam=lam, left=True, blending=blending)
    q = q0.copy()
    out = m.ik_rmr_ex(tasks, q, x_des, activation=act, lam=lam, left_inverse=True, blending=blending,
                      want=("dx", "dq", "det", "dq_ts", "dq_ns"))
    assert_close(out["dx"], ref["dx"], what="dx")
    # the nq x nq system J_r^T Wx J_r + lambda 1: conditioning bounded by (|J_r|^2 |Wx| + lambda) / lambda
    _, J = rik.task_kinematics(q0)
    w = rik.graph.inv_wq()
    Jr = J * w[None, None, :]
    cond = (np.linalg.norm(Jr, axis=(1, 2)) ** 2 * max(1.0, max(act)) + lam) / lam
    for k in ("dq_ts", "dq_ns"):
        assert_step_close(out[k], ref[k], cond, k)
    assert_step_close(q - q0, ref["q"] - q0, cond, "q")
    assert np.all(np.abs(out["det"] / ref["det"] - 1.0) <= 1e-11 * np.maximum(1.0, cond / 100.0))
    if blending != BLEND_BINARY:
        binary = rik.rmr_ex(q0, x_des, activation=act, lam=lam, left=True, blending=0)
        assert np.max(np.abs(binary["dq_ts"] - ref["dq_ts"])) > 1e-6      # the modes really differ


def test_unsupported_requests_fail_loudly(cuda_device):
    g = rcs_b200.Graph(graph_file("wam_arm"))
    m = rcs_b200.Model(g, cuda_device)
    tip = g.body_index("BallTip")
    tasks = [(TASK_XYZ, tip), (TASK_ABC, tip)]
    q, x = np.zeros((4, 7)), np.zeros((4, 6))
    with pytest.raises(rcs_b200.RcsbError, match="not implemented"):
        m.ik_rmr_ex(tasks, q, x, blending=3, left_inverse=True)          # IndependentTasks
    with pytest.raises(rcs_b200.RcsbError, match="no task is active"):
        m.ik_rmr_ex(tasks, q, x, activation=[0.0, 0.0])
    with pytest.raises(rcs_b200.RcsbError, match="scalar lambda"):
        m.ik_rmr_ex(tasks, q, x, lambda_rows=np.ones(6), left_inverse=True)
