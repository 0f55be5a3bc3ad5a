"""GPU parity on graphs that come from URDF files (tests/golden/urdf/): forward kinematics with
velocities, Jacobians, kinetic terms and an IK step, ours (URDF loader -> flattener -> kernels) against
the reference (its URDF loader -> RcsGraph_setState / Jacobians / RcsGraph_computeKineticTerms /
IkSolverRMR). Covers asymmetric inertia tensors (the reference's URDF loader mirrors products of
inertia with the opposite sign), mimic joints, skew joint axes and the truncated backward chains of
URDF sub-trees attached with prev=."""
import os

import numpy as np
import pytest

import rcs_b200
from rcs_b200 import TASK_ABC, TASK_XYZ, workload
from conftest import ROOT, assert_close, assert_structural_zeros
from test_ik_gpu import assert_step_close

pytestmark = pytest.mark.gpu
URDF_DIR = os.path.join(ROOT, "tests", "golden", "urdf")
CASES = [("urdfbot.urdf", ["finger", "camera", "forearm"]), ("urdfarm.urdf", ["tip", "boom"])]


def _both(name, refmod, dev):
    rcs_b200.lib().Rcs_addResourcePath(URDF_DIR.encode())
    refmod.lib().ref_add_resource_path(URDF_DIR.encode())
    path = os.path.join(URDF_DIR, name)
    g = rcs_b200.Graph(path)
    return g, rcs_b200.Model(g, dev), refmod.RefGraph(path)


@pytest.mark.parametrize("name,effectors", CASES)
def test_urdf_fk_jacobians_kinetic_terms(name, effectors, refmod, cuda_device):
    g, m, r = _both(name, refmod, cuda_device)
    n = 400
    q, qd = workload.sample_states(g.layout(), n, first=17, with_velocity=True)
    ours = m.fk(q, qd, want=("A_BI", "A_JI", "xdot", "omega", "q"))
    ref = r.fk(q, qd, want=("A_BI", "A_JI", "xdot", "omega", "q"))
    for k in ("A_BI", "A_JI", "xdot", "omega", "q"):
        assert_close(ours[k], ref[k], what=f"{name} {k}")
    eff = [g.body_index(e) for e in effectors]
    pts = np.array([[0.01 * i, -0.02, 0.03] for i in range(len(eff))])
    oj = m.fk_jacobians(q, eff, points=pts, want=("JT", "JR"))
    rj = r.fk_jacobians(q, eff, points=pts, want=("JT", "JR"))
    for k in ("JT", "JR"):
        assert_close(oj[k], rj[k], what=f"{name} {k}")
        assert_structural_zeros(oj[k], rj[k], what=f"{name} {k}")
    ok = m.kinetic_terms(q, qd)
    rk = r.kinetic_terms(q, qd)
    for k in ("M", "h", "Fg", "E"):
        assert_close(ok[k], rk[k], tol=1e-11 if k == "h" else 1e-12, what=f"{name} {k}")
    assert_close(m.kinetic_terms(q, want=("M",))["M"], r.mass_matrix(q), what=f"{name} M (mass matrix only)")


def test_urdf_ik_step(refmod, cuda_device):
    g, m, _ = _both("urdfbot.urdf", refmod, cuda_device)
    names = [("XYZ", "wrist"), ("ABC", "wrist")]
    refmod.lib().ref_add_resource_path(URDF_DIR.encode())
    rik = refmod.RefIk(os.path.join(URDF_DIR, "urdfbot.urdf"), names)
    n, lam = 300, 1.0e-4
    q_goal = workload.sample_states(g.layout(), n, first=4, margin=0.2)
    x_des, _ = rik.task_kinematics(q_goal)
    rng = np.random.default_rng(3)
    q0 = np.ascontiguousarray(q_goal + 0.05 * rng.uniform(-1, 1, size=q_goal.shape))
    q_ref, dx_ref, dq_ref, det_ref = rik.rmr(q0, np.ascontiguousarray(x_des), lam=lam)
    q = q0.copy()
    wrist = g.body_index("wrist")
    out = m.ik_rmr([(TASK_XYZ, wrist), (TASK_ABC, wrist)], q, np.ascontiguousarray(x_des), lam=lam,
                   want=("dx", "dq", "det"))
    _, J = rik.task_kinematics(q0)
    w = rik.graph.inv_wq()
    cond = np.linalg.cond(np.einsum("nij,j,nkj->nik", J, w, J) + lam * np.eye(J.shape[1]))
    assert_close(out["dx"], dx_ref, what="dx")
    assert_step_close(out["dq"], dq_ref, cond, "dq")
    assert_step_close(q - q0, q_ref - q0, cond, "q")


def test_urdf_nodes_in_a_graph_file(refmod, cuda_device):
    """<URDF> nodes of a Graph file: an arm below an articulated body, a two-root robot on rigid-body
    joints. The reference's loader wires the joints of a URDF tree before the tree is attached and
    leaves their backward chains ending at the URDF root; RcsGraph_clone re-inserts every joint and
    with that completes them. The oracle evaluates on clones (one per thread), our loader wires the
    chains like the clone (rcsb_parser.c, urdfFromXML) — so everything must agree, including the
    Turntable's column in the Jacobians of the attached arm; x_dot = JT q_dot and omega = JR q_dot
    (the reference's own invariant, bin/TestGraph.cpp:545-566) are checked on top."""
    g, m, r = _both("urdf_in_graph.xml", refmod, cuda_device)
    lay = g.layout()
    n = 300
    q, qd = workload.sample_states(lay, n, first=23, with_velocity=True)
    ours = m.fk(q, qd, want=("A_BI", "A_JI", "xdot", "omega", "q"))
    ref = r.fk(q, qd, want=("A_BI", "A_JI", "xdot", "omega", "q"))
    for k in ("A_BI", "A_JI", "xdot", "omega", "q"):
        assert_close(ours[k], ref[k], what=k)
    names = ["tip_A", "boom_A", "finger_B", "camera_B", "Turntable"]
    eff = [g.body_index(e) for e in names]
    oj = m.fk_jacobians(q, eff, want=("JT", "JR"))
    rj = r.fk_jacobians(q, eff, want=("JT", "JR"))
    for k in ("JT", "JR"):
        assert_close(oj[k], rj[k], what=k)
        assert_structural_zeros(oj[k], rj[k], what=k)
    assert np.abs(oj["JT"][:, 0, :, 0]).max() > 0.1      # the Turntable moves the arm's tip
    cols = [j["jointIndex"] for j in lay["joints"] if j["jacobiIndex"] >= 0]
    qd_ik = qd[:, cols]
    for i, e in enumerate(eff):
        if names[i] == "finger_B":
            continue    # moved by a URDF mimic joint: constrained (no column) but driven by its master
        assert_close(np.einsum("nrc,nc->nr", oj["JT"][:, i], qd_ik), ours["xdot"][:, e], what=f"x_dot = JT q_dot, {names[i]}")
        assert_close(np.einsum("nrc,nc->nr", oj["JR"][:, i], qd_ik), ours["omega"][:, e], what=f"omega = JR q_dot, {names[i]}")
    ok = m.kinetic_terms(q, qd)
    rk = r.kinetic_terms(q, qd)
    for k in ("M", "h", "Fg", "E"):
        assert_close(ok[k], rk[k], tol=1e-11 if k == "h" else 1e-12, what=k)
    assert np.abs(ok["M"][:, 0, 1:4]).max() > 1e-3
