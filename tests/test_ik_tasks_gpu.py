"""GPU parity of rcsb_ik_tasks_rmr — RMR steps over mixed task stacks with reference bodies and
frames — against the reference's ControllerBase (TaskPosition3D, TaskEuler3D, TaskPose6D, TaskCOM1D,
TaskJoint through its TaskFactory) + IkSolverRMR::solveRightInverse. Includes the active task set
of GenericHumanoid/cAction.xml (nx = 20), BASELINE.json configs[4]."""
import numpy as np
import pytest

import rcs_b200
from rcs_b200 import (TASK_ABC, TASK_COGX, TASK_COGY, TASK_COGZ, TASK_JOINT, TASK_POLARX, TASK_POLARY, TASK_POLARZ,
                      TASK_XYZ, TASK_XYZABC, workload)
from conftest import assert_close, graph_file
from test_ik_gpu import assert_step_close

pytestmark = pytest.mark.gpu
CODE = {"XYZ": TASK_XYZ, "ABC": TASK_ABC, "XYZABC": TASK_XYZABC, "COGX": TASK_COGX, "COGY": TASK_COGY,
        "COGZ": TASK_COGZ, "Joint": TASK_JOINT, "POLARX": TASK_POLARX, "POLARY": TASK_POLARY, "POLARZ": TASK_POLARZ}


def _spec(g, names):
    """reference task tuples (controlVariable, effector[, refBdy[, refFrame]]) -> RcsbTaskSpec tuples"""
    lay = g.layout()
    jidx = {j["name"]: i for i, j in enumerate(lay["joints"])}
    out = []
    for t in names:
        if t[0] == "Joint":
            if len(t) > 2:      # reference joint and gain (TaskJoint's refJnt / refGain)
                out.append((CODE[t[0]], jidx[t[1]], jidx[t[2]], -1, float(t[3]) if len(t) > 3 else 1.0))
            else:
                out.append((CODE[t[0]], jidx[t[1]]))
        else:
            out.append(tuple([CODE[t[0]]] + [g.body_index(b) if b else -1 for b in t[1:]]))
    return out


def _run(name, names, refmod, dev, n, lam=1.0e-6, iters=1, first=21):
    g = rcs_b200.Graph(graph_file(name))
    m = rcs_b200.Model(g, dev)
    rik = refmod.RefIk(graph_file(name), names)
    q_goal = workload.sample_states(g.layout(), n, first=first, margin=0.25)
    x_des, _ = rik.task_kinematics(q_goal)
    rng = np.random.default_rng(first)
    q0 = np.ascontiguousarray(q_goal + 0.08 * rng.uniform(-1, 1, size=q_goal.shape))
    x_des = np.ascontiguousarray(x_des)
    q_ref, dx_ref, dq_ref, det_ref = rik.rmr(q0, x_des, lam=lam, iters=iters)
    q = q0.copy()
    out = m.ik_tasks_rmr(_spec(g, names), q, x_des, lam=lam, iters=iters, want=("dx", "dq", "det"))
    _, J = rik.task_kinematics(q0)
    w = rik.graph.inv_wq()
    cond = np.linalg.cond(np.einsum("nij,j,nkj->nik", J, w, J) + lam * np.eye(J.shape[1]))
    return dict(q0=q0, q=q, out=out, q_ref=q_ref, dx_ref=dx_ref, dq_ref=dq_ref, det_ref=det_ref, cond=cond)


def _check(r, iters=1):
    assert_close(r["out"]["dx"], r["dx_ref"], what="dx") if iters == 1 else None
    mult = float(iters)
    assert_step_close(r["q"] - r["q0"], r["q_ref"] - r["q0"], mult * r["cond"], "q")
    if iters == 1:
        assert_step_close(r["out"]["dq"], r["dq_ref"], r["cond"], "dq")
        rel = np.abs(r["out"]["det"] / r["det_ref"] - 1.0)
        assert np.all(rel <= 1e-11 * np.maximum(1.0, r["cond"] / 100.0)), "det"


def test_cAction_task_set_of_the_humanoid(refmod, cuda_device):
    """hands XYZ, COG x / y of the whole robot (subtree of the Heel), both feet XYZABC relative to the
    Heel: the nx = 20 stack of configs[4], solved."""
    names = [("XYZ", "RightHandTip"), ("XYZ", "LeftHandTip"), ("COGX", "Heel"), ("COGY", "Heel"),
             ("XYZABC", "FootL", "Heel"), ("XYZABC", "FootR", "Heel")]
    r = _run("humanoid", names, refmod, cuda_device, 128)
    assert r["out"]["dx"].shape[1] == 20
    _check(r)


@pytest.mark.parametrize("name,names", [
    ("wam_arm", [("XYZ", "BallTip", "Link2", "Link4"), ("ABC", "BallTip", "Link2")]),
    ("wam_arm", [("XYZ", "BallTip", "", "Link3"), ("ABC", "BallTip", "Link4", "Link2")]),
    ("wam_arm", [("XYZABC", "BallTip", "Link1"), ("COGZ", "")]),
    ("wam_arm", [("Joint", "ArmJoint2"), ("XYZ", "BallTip"), ("COGX", "Link3")]),
    ("wam_arm", [("ABC", "BallTip", "Base", "Base"), ("XYZ", "Link4", "Base")]),
    ("gantry", [("XYZ", "Tool", "Trolley"), ("ABC", "Tool", "Quill", "Trolley"), ("COGY", "Quill")]),
    ("gantry", [("Joint", "WristYaw"), ("Joint", "GantryX"), ("XYZ", "Camera", "", "WristLink")]),
    ("wam_arm", [("COGX", "", "Link3"), ("COGZ", "", "Link3"), ("XYZ", "BallTip")]),
    ("wam_arm", [("COGY", "Link2", "Base"), ("COGX", "Link2", "Base"), ("ABC", "BallTip")]),
    ("humanoid", [("COGX", "", "FootL"), ("COGY", "", "FootL"), ("XYZ", "RightHandTip", "FootL")]),
], ids=["rel-frames", "frame-only", "pose6d-cogz", "joint-xyz-subtree-cog", "fixed-reference", "gantry-rel",
        "gantry-joints", "cog-rel-articulated", "cog-subtree-rel-fixed", "humanoid-cog-rel-foot"])
def test_reference_bodies_frames_cog_and_joint_tasks(name, names, refmod, cuda_device):
    """Every branch of RcsGraph_3dPosJacobian / RcsGraph_3dOmegaJacobian (world, refFrame only,
    articulated refBdy, refFrame != refBdy, fixed reference body), relative Euler angles, COG of the
    whole graph and of a subtree, joint tasks."""
    r = _run(name, names, refmod, cuda_device, 300)
    _check(r)


@pytest.mark.parametrize("name,names", [
    ("wam_arm", [("Joint", "ArmJoint2", "ArmJoint4", "-0.5"), ("XYZ", "BallTip"), ("Joint", "ArmJoint6", "ArmJoint7")]),
    ("gantry", [("Joint", "GantryX", "GantryY", "2.0"), ("XYZ", "Tool")]),
], ids=["arm-gain-and-default-gain", "gantry-translational-joints"])
def test_joint_tasks_with_a_reference_joint(name, names, refmod, cuda_device):
    """TaskJoint with refJnt / refGain (TaskJoint.cpp:70-82, :227-265): x = q + refGain q_ref, the
    Jacobian row carries refGain in the reference joint's column; refGain defaults to 1."""
    r = _run(name, names, refmod, cuda_device, 300)
    _check(r)
    r3 = _run(name, names, refmod, cuda_device, 200, iters=3, first=5)
    _check(r3, iters=3)


@pytest.mark.parametrize("name,names", [
    ("wam_arm", [("POLARZ", "BallTip"), ("XYZ", "BallTip")]),
    ("wam_arm", [("POLARX", "BallTip", "Link2"), ("Joint", "ArmJoint1")]),
    ("wam_arm", [("XYZ", "Link4"), ("POLARY", "BallTip", "Base")]),
    ("gantry", [("POLARZ", "Tool", "Trolley"), ("XYZ", "Camera")]),
    ("humanoid", [("POLARZ", "RightHandTip", "Heel"), ("POLARX", "LeftHandTip"), ("COGX", "Heel"), ("COGY", "Heel")]),
], ids=["polar-z-world", "polar-x-relative", "polar-y-fixed-reference", "gantry-polar", "humanoid-polar"])
def test_polar_tasks(name, names, refmod, cuda_device):
    """TaskPolar2D for every axisDirection, world / articulated / fixed reference body: polar-angle
    targets, Vec3d_getPolarErrorFromDirection, rows of the effector-frame angular Jacobian. The start
    states are 0.08 rad away from the goal states, so the angle between current and desired axis stays
    away from 0 where its acos() is ill-conditioned in the reference itself."""
    r = _run(name, names, refmod, cuda_device, 300)
    _check(r)


def test_polar_task_on_target_and_opposite_axis(refmod, cuda_device):
    """Ends of the range of Vec3d_getPolarErrorFromDirection: on target the error is exactly 0; with the
    desired axis (numerically almost) opposite to the current one the rotation axis is decided by
    rounding noise in the reference itself, the angle is pi."""
    names = [("POLARZ", "BallTip")]
    g = rcs_b200.Graph(graph_file("wam_arm"))
    m = rcs_b200.Model(g, cuda_device)
    rik = refmod.RefIk(graph_file("wam_arm"), names)
    q0 = np.zeros((2, g.dof))      # BallTip's z axis is the world's at q = 0: polar angles (0, 0)
    x_cur, _ = rik.task_kinematics(q0)
    assert np.all(x_cur == 0.0)
    x_des = np.ascontiguousarray(x_cur)
    x_des[1] = [np.pi, 0.0]
    _, dx_ref, dq_ref, _ = rik.rmr(q0, x_des, lam=1e-6, iters=1)
    q = q0.copy()
    out = m.ik_tasks_rmr(_spec(g, names), q, x_des, lam=1e-6, iters=1, want=("dx", "dq"))
    assert np.all(out["dx"][0] == 0.0) and np.all(dx_ref[0] == 0.0)
    assert_close(out["dq"][0], dq_ref[0], what="dq on target")
    assert abs(np.linalg.norm(out["dx"][1]) - np.pi) < 1e-12 and abs(np.linalg.norm(dx_ref[1]) - np.pi) < 1e-12


def test_ten_iterations_of_a_mixed_stack_device_pointers(refmod, cuda_device):
    import torch

    names = [("XYZ", "RightHandTip"), ("COGX", "Heel"), ("COGY", "Heel"), ("XYZABC", "FootL", "Heel")]
    g = rcs_b200.Graph(graph_file("humanoid"))
    m = rcs_b200.Model(g, cuda_device)
    rik = refmod.RefIk(graph_file("humanoid"), names)
    n, lam = 64, 1.0e-6
    q_goal = workload.sample_states(g.layout(), n, first=8, margin=0.25)
    x_des, _ = rik.task_kinematics(q_goal)
    q0 = np.ascontiguousarray(q_goal + 0.05)
    q_ref = q0.copy()
    cond = np.zeros(n)
    w = rik.graph.inv_wq()
    for _ in range(10):
        _, J = rik.task_kinematics(q_ref)
        cond = np.maximum(cond, np.linalg.cond(np.einsum("nij,j,nkj->nik", J, w, J) + lam * np.eye(J.shape[1])))
        q_ref, dx_ref, _, _ = rik.rmr(q_ref, np.ascontiguousarray(x_des), lam=lam, iters=1)
    dq_, dx_ = torch.from_numpy(q0.copy()).cuda(cuda_device), torch.from_numpy(np.ascontiguousarray(x_des)).cuda(cuda_device)
    out = m.ik_tasks_rmr(_spec(g, names), dq_, dx_, lam=lam, iters=10, want=("dx",))
    torch.cuda.synchronize()
    q = dq_.cpu().numpy()
    assert_step_close(q - q0, q_ref - q0, 10.0 * cond, "q after 10 iterations")
    assert np.median(np.abs(out["dx"].cpu().numpy())) < 1e-3       # converges


def test_task_stack_errors(cuda_device):
    g = rcs_b200.Graph(graph_file("wam_scenario"))
    m = rcs_b200.Model(g, cuda_device)
    with pytest.raises(rcs_b200.RcsbError, match="coupled"):
        m.ik_tasks_rmr([(TASK_XYZ, g.body_index("BallTip"))], np.zeros((2, g.dof)), np.zeros((2, 3)))
    g2 = rcs_b200.Graph(graph_file("wam_arm"))
    m2 = rcs_b200.Model(g2, cuda_device)
    with pytest.raises(rcs_b200.RcsbError, match="same subtree"):
        m2.ik_tasks_rmr([(TASK_COGX, -1), (TASK_COGY, 3)], np.zeros((2, 7)), np.zeros((2, 2)))
    with pytest.raises(rcs_b200.RcsbError, match="reference frame"):
        m2.ik_tasks_rmr([(TASK_COGX, -1, 2, 3)], np.zeros((2, 7)), np.zeros((2, 1)))
