"""Pins the numpy oracle (oracle/rcs_oracle.py): golden vectors recorded from the unmodified
reference (tools/make_golden.py), the known answers of SURVEY.md §8c, the invariants the
reference's own tests check (bin/TestGraph.cpp:314-352, :545-566; Rcs_mathTests.cpp:590-726),
and — where oracle/_ref/librcsref.so exists — the reference itself on fresh samples.

No GPU. Tolerance 1e-12 relative to the array's inf-norm (floor 1) unless stated.
"""
import json
import os

import numpy as np
import pytest

from conftest import GRAPH_DIR, GRAPHS, ROOT, assert_close, assert_structural_zeros, graph_file
from oracle import rcs_oracle
from rcs_b200 import workload

GOLDEN = ["wam_arm", "wam_scenario", "humanoid", "husky", "gantry", "polycoupled", "wrist6"]


def layout(name):
    with open(os.path.join(GRAPH_DIR, name + ".layout.json")) as f:
        return json.load(f)


def golden(name):
    return np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))


@pytest.mark.parametrize("name", GOLDEN)
def test_forward_kinematics_matches_golden(name):
    z, o = golden(name), rcs_oracle.Oracle(layout(name))
    out = o.fk(z["q"], z["qd"])
    for k in ("A_BI", "A_JI", "xdot", "omega", "q"):
        assert_close(out[k], z["fk_" + k], what=f"{name}:{k}")


@pytest.mark.parametrize("name", GOLDEN)
def test_jacobians_match_golden(name):
    z, o = golden(name), rcs_oracle.Oracle(layout(name))
    out = o.fk_jacobians(z["q"], [int(e) for e in z["eff"]], points=z["pts"])
    for k in ("JT", "JR"):
        assert_close(out[k], z["jac_" + k], what=f"{name}:{k}")
        assert_structural_zeros(out[k], z["jac_" + k], what=f"{name}:{k}")


@pytest.mark.parametrize("name", GOLDEN)
def test_kinetic_terms_match_golden(name):
    z, o = golden(name), rcs_oracle.Oracle(layout(name))
    out = o.kinetic_terms(z["q"], z["qd"])
    for k in ("M", "h", "Fg", "E"):
        assert_close(out[k], z["kt_" + k], what=f"{name}:{k}")


@pytest.mark.parametrize("name", GOLDEN)
def test_cog_matches_golden(name):
    z, o = golden(name), rcs_oracle.Oracle(layout(name))
    cog, jcog = o.cog(z["q"])
    assert_close(cog, z["cog"], what=f"{name}:cog")
    assert_close(jcog, z["Jcog"], what=f"{name}:Jcog")


@pytest.mark.parametrize("name", GOLDEN)
def test_relative_task_jacobians_match_golden(name):
    z, o = golden(name), rcs_oracle.Oracle(layout(name))
    spec = [("POS" if k == 1 else "OMEGA", int(e), int(rb), int(rf)) for k, e, rb, rf in z["rel_spec"]]
    assert_close(o.task_jacobians(z["q"], spec), z["rel_J"], what=f"{name}: 3dPos / 3dOmega Jacobians")


@pytest.mark.parametrize("name", GOLDEN)
def test_direct_dynamics_matches_golden(name):
    """The solve goes through M^-1: bar scaled by cond(M) per state (see tests/test_kinetic_gpu.py)."""
    z, o = golden(name), rcs_oracle.Oracle(layout(name))
    qdd, _ = o.direct_dynamics(z["q"], z["qd"], z["dd_Fext"], z["dd_Fjnt"])
    cond = np.linalg.cond(z["kt_M"])
    err = np.abs(qdd - z["dd_qdd"]).max(axis=1)
    tol = 1e-12 * np.maximum(1.0, cond / 100.0) * np.maximum(1.0, np.abs(z["dd_qdd"]).max(axis=1))
    assert np.all(err <= tol), float((err / tol).max())


@pytest.mark.parametrize("name", ["wam_arm", "gantry", "wrist6"])
def test_left_inverse_and_mass_matrix_match_golden(name):
    """Recorded from the reference: IkSolverRMR::solveLeftInverse steps (lambda 1e-3) and
    RcsGraph_computeMassMatrix."""
    z, o = golden(name), rcs_oracle.Oracle(layout(name))
    tip = int(z["ik_tip"])
    tasks = [("XYZ", tip), ("ABC", tip)]
    q1, dx1, dq1, det1 = o.ik_rmr(z["ik_q_start"], tasks, z["ik_x_des"], lam=1.0e-3, iters=1, left=True)
    assert_close(dx1, z["ikl_dx1"], what="dx")
    ok = z["ikl_det1"] > 0.0
    assert ok.all()
    assert np.abs(dq1 - z["ikl_dq1"]).max() <= 1e-9 and np.abs(q1 - z["ikl_q1"]).max() <= 1e-9
    assert np.abs(det1 / z["ikl_det1"] - 1.0).max() <= 1e-9
    q5 = o.ik_rmr(z["ik_q_start"], tasks, z["ik_x_des"], lam=1.0e-3, iters=5, left=True)[0]
    assert np.abs(q5 - z["ikl_q5"]).max() <= 1e-8
    assert_close(o.kinetic_terms(z["q"])["M"], z["mass_M"], what="M")


@pytest.mark.parametrize("name", ["wam_arm", "humanoid", "gantry", "wrist6"])
def test_ik_matches_golden(name):
    z, o = golden(name), rcs_oracle.Oracle(layout(name))
    tip = int(z["ik_tip"])
    tasks = [("XYZ", tip), ("ABC", tip)]
    x, J = o.task_kinematics(z["ik_q_start"], tasks)
    q1, dx1, dq1, det1 = o.ik_rmr(z["ik_q_start"], tasks, z["ik_x_des"], iters=1)
    assert_close(dx1, z["ik_dx1"], what="dx after 0 steps")
    # the step goes through (J W J^T + lambda 1)^-1: bar scaled by its condition number
    w = o.inv_wq()
    A = np.einsum("nij,j,nkj->nik", J, w, J) + 1e-8 * np.eye(J.shape[1])
    cond = np.linalg.cond(A)
    err = np.abs(dq1 - z["ik_dq1"]).max(axis=1)
    assert np.all(err <= 1e-12 * np.maximum(1.0, cond / 100.0) * np.maximum(1.0, np.abs(z["ik_dq1"]).max(axis=1)))
    assert np.all(np.abs(det1 / z["ik_det1"] - 1.0) <= 1e-11 * np.maximum(1.0, cond / 100.0))
    # ten iterations on the well-conditioned half
    q10, dx10, _, det10 = o.ik_rmr(z["ik_q_start"], tasks, z["ik_x_des"], iters=10)
    ok = cond < np.median(cond) * 1.0001
    assert_close(q10[ok], z["ik_q10"][ok], tol=1e-9, what="q after 10 steps")
    # goal Jacobian recorded by ControllerBase::computeJ
    _, Jg = o.task_kinematics(_goal_states(name, len(z["ik_q_start"])), tasks)
    assert_close(Jg, z["ik_J_goal"], what="J at the goal states")


def _goal_states(name, n):
    return workload.sample_states(layout(name), n, first=2000, margin=0.2)


def test_known_answers_of_the_survey():
    """WAM arm at its default state, qd_i = 0.1 (i + 1); values printed by the
    reference library during the survey."""
    lay = layout("wam_arm")
    o = rcs_oracle.Oracle(lay)
    q = np.array([lay["q"]])
    qd = 0.1 * (np.arange(7)[None, :] + 1.0)
    tip = [b["name"] for b in lay["bodies"]].index("BallTip")
    out = o.fk_jacobians(q, [tip])
    np.testing.assert_allclose(out["A_BI"][0, tip, :3],
                               [0.83632952411278927, 0.30514499825747998, 0.83766965256912063],
                               rtol=0, atol=1e-15)
    np.testing.assert_allclose(out["JT"][0, 0, 0],
                               [-0.16514499825747997, 0.47491641541908702, -0.14053874149066872,
                                -0.040723861313878298, 6.9e-18, 0.033210789380311002, 1.7e-17],
                               rtol=0, atol=2e-17 + 1e-15)
    np.testing.assert_allclose(out["JR"][0, 0, 0],
                               [0, -0.25881904510252074, 0.16773125949652062, -0.25881904510252074,
                                0.95125124256419769, 0.25881904510252085, 0.95125124256419769],
                               rtol=0, atol=1e-15)
    kt = o.kinetic_terms(q, qd)
    assert abs(kt["M"][0, 0, 0] - 0.4777633998038765) < 1e-14
    assert abs(kt["h"][0, 0] - (-0.027243766372026575)) < 1e-14
    assert abs(kt["Fg"][0, 1] - 10.532807395973947) < 1e-13
    assert abs(kt["E"][0] - 62.105160163860795) < 1e-12

    lay = layout("humanoid")
    o = rcs_oracle.Oracle(lay)
    q = np.array([lay["q"]])
    tip = [b["name"] for b in lay["bodies"]].index("RightHandTip")
    np.testing.assert_allclose(o.fk(q)["A_BI"][0, tip, :3],
                               [0.29073479823858578, -0.34759266218116014, 0.56461616414897053],
                               rtol=0, atol=1e-15)
    assert abs(o.kinetic_terms(q)["M"][0, 0, 0] - 54.080000000000005) < 1e-12


@pytest.mark.parametrize("name", GRAPHS)
def test_velocity_identities(name):
    """x_dot = JT q_dot and omega = JR q_dot (bin/TestGraph.cpp:545-566)."""
    lay = layout(name)
    o = rcs_oracle.Oracle(lay)
    q, qd = workload.sample_states(lay, 9, first=40, with_velocity=True)
    fk = o.fk(q, qd)
    cols = o.ik_columns()
    # A constrained joint that is also a coupled slave (dexbot knuck3-base) gets a velocity from
    # its master; the reference adds it to BODY->omega although the joint has no Jacobian
    # column, so the identity only holds for bodies not behind such a joint.
    moving_constrained = {j for j in range(o.dof) if o.jac[j] < 0 and np.any(fk["qd"][:, j] != 0.0)}
    bodies = [b for b in range(0, o.nb, max(1, o.nb // 6))
              if not moving_constrained.intersection(o.chain(b))]
    jac = o.fk_jacobians(q, bodies)
    qd_ik = fk["qd"][:, cols]
    assert_close(np.einsum("nbij,nj->nbi", jac["JT"], qd_ik), fk["xdot"][:, bodies], what="xdot")
    assert_close(np.einsum("nbij,nj->nbi", jac["JR"], qd_ik), fk["omega"][:, bodies], what="omega")


@pytest.mark.parametrize("name", ["wam_arm", "humanoid", "husky"])
def test_gravity_and_energy_identities(name):
    """F_gravity = J_cog^T (0, 0, -m g) (bin/TestGraph.cpp:314-352); M symmetric; and
    T = 1/2 sum_b (m |v_com|^2 + w^T I w) from the body velocities of the FK pass."""
    lay = layout(name)
    o = rcs_oracle.Oracle(lay)
    q, qd = workload.sample_states(lay, 6, first=7, with_velocity=True)
    kt = o.kinetic_terms(q, qd)
    cog, jcog = o.cog(q)
    mass = sum(b["m"] for b in lay["bodies"])
    assert_close(kt["Fg"], jcog[:, 2, :] * (-9.81 * mass), tol=1e-11, what="gravity identity")
    if name != "husky":   # husky.xml carries asymmetric inertia tensors; the reference keeps them
        assert_close(kt["M"], np.transpose(kt["M"], (0, 2, 1)), what="M symmetric")
    static = o.kinetic_terms(q)
    assert_close(static["E"], mass * 9.81 * cog[:, 2], tol=1e-11, what="potential energy")
    assert np.all(static["h"] == 0.0)


def test_cholesky_inverse_properties():
    """L L^T = A, A A^-1 = 1 at 1e-12 (Rcs_mathTests.cpp:590-726); failure -> det 0, zeros."""
    rng = np.random.default_rng(1)
    B = rng.normal(size=(50, 6, 6))
    A = np.einsum("nij,nkj->nik", B, B) + 0.5 * np.eye(6)
    o = rcs_oracle.Oracle(layout("wam_arm"))
    inv, det = o._cholesky_inverse(A)
    assert_close(np.einsum("nij,njk->nik", A, inv), np.tile(np.eye(6), (50, 1, 1)), tol=1e-10, what="A A^-1")
    np.testing.assert_allclose(det, np.linalg.det(A), rtol=1e-10)
    A[3] = -A[3]
    inv, det = o._cholesky_inverse(A)
    assert det[3] == 0.0 and np.all(inv[3] == 0.0) and det[2] > 0.0


def test_euler_conversions_round_trip():
    rng = np.random.default_rng(2)
    ea = rng.uniform(-1.5, 1.5, size=(200, 3))
    A = rcs_oracle.Oracle.from_euler(ea)
    assert_close(np.einsum("nij,nkj->nik", A, A), np.tile(np.eye(3), (200, 1, 1)), what="orthonormal")
    assert_close(rcs_oracle.Oracle.euler_angles(A.reshape(-1, 9)), ea, what="toEuler(fromEuler)")
    # gimbal branch (cy <= 16 FLT_EPSILON): b = +-90 deg, c reported as 0 (Rcs_Mat3d.c:956-972)
    g = rcs_oracle.Oracle.euler_angles(rcs_oracle.Oracle.from_euler(np.array([[0.3, np.pi / 2, 0.0]])).reshape(-1, 9))
    assert g[0, 2] == 0.0 and abs(g[0, 1] - np.pi / 2) < 1e-7


def test_state_vector_expansion_and_coupling():
    """RcsGraph_stateVectorFromIK (Rcs_graph.c:628) + updateSlaveJoints (Rcs_graph.c:2756)."""
    lay = layout("wam_scenario")
    o = rcs_oracle.Oracle(lay)
    slaves = [j for j in lay["joints"] if j["coupledTo"] >= 0]
    assert slaves, "wam_scenario has coupled finger joints"
    q = workload.sample_states(lay, 5)
    qf, _ = o.expand(q)
    for s in slaves:
        m = lay["joints"][s["coupledTo"]]
        if len(s["couplingFactors"]) == 1:
            expect = s["q_init"] + s["couplingFactors"][0] * (qf[:, m["jointIndex"]] - m["q_init"])
            assert np.array_equal(qf[:, s["jointIndex"]], expect)
    cols = o.ik_columns()
    q_ik = np.ascontiguousarray(q[:, cols])
    qf2, _ = o.expand(q_ik)
    constrained = [j["jointIndex"] for j in lay["joints"] if j["constrained"]]
    assert np.array_equal(qf2[:, constrained], np.tile(np.array(lay["q"])[constrained], (5, 1)))
    assert np.array_equal(qf2[:, cols], qf[:, cols])


def test_empty_batch():
    o = rcs_oracle.Oracle(layout("wam_arm"))
    out = o.fk_jacobians(np.empty((0, 7)), [3])
    assert out["A_BI"].shape == (0, 12, 12) and out["JT"].shape == (0, 1, 3, 7)
    assert o.kinetic_terms(np.empty((0, 7)))["M"].shape == (0, 7, 7)


# ---------------------------------------------------------------------------------------------
# against the reference itself (CPU container only: needs oracle/_ref built from /root/reference)
# ---------------------------------------------------------------------------------------------

@pytest.mark.parametrize("name", GRAPHS)
def test_oracle_vs_reference_library(name, refmod):
    lay = layout(name)
    o = rcs_oracle.Oracle(lay)
    r = refmod.RefGraph(graph_file(name))
    q, qd = workload.sample_states(lay, 33, first=500, with_velocity=True)
    a, b = o.fk(q, qd), r.fk(q, qd, want=("A_BI", "A_JI", "xdot", "omega", "q"))
    for k in b:
        assert_close(a[k], b[k], what=f"{name}:{k}")
    effs = [o.nb - 1, o.nb // 2]
    pts = np.random.default_rng(0).uniform(-0.3, 0.3, size=(2, 3))
    a, b = o.fk_jacobians(q, effs, points=pts), r.fk_jacobians(q, effs, points=pts)
    for k in ("JT", "JR"):
        assert_close(a[k], b[k], what=f"{name}:{k}")
    a, b = o.kinetic_terms(q, qd), r.kinetic_terms(q, qd)
    for k in ("M", "h", "Fg", "E"):
        assert_close(a[k], b[k], what=f"{name}:{k}")
    cog, jcog = o.cog(q)
    rcog, rjcog = r.cog(q)
    assert_close(cog, rcog, what="cog")
    assert_close(jcog, rjcog, what="Jcog")
    assert_close(o.inv_wq(), r.inv_wq(), what="invWq")
    grad, cost = o.joint_limit_gradient(q)
    rgrad, rcost = r.joint_limit_gradient(q)
    assert_close(grad, rgrad, what="joint limit gradient")
    assert_close(cost, rcost, what="joint limit cost")


def test_oracle_ik_vs_reference_library(refmod):
    lay = layout("wam_arm")
    o = rcs_oracle.Oracle(lay)
    tip = [b["name"] for b in lay["bodies"]].index("BallTip")
    rik = refmod.RefIk(graph_file("wam_arm"), [("XYZ", "BallTip"), ("ABC", "BallTip")])
    tasks = [("XYZ", tip), ("ABC", tip)]
    q0 = workload.sample_states(lay, 400, first=5000, margin=0.1)
    x_des, J = rik.task_kinematics(workload.sample_states(lay, 400, first=9000, margin=0.1))
    # far targets: axis-angle branch of the orientation error
    dx = o.ik_task_error(q0, tasks, x_des)
    _, dx_ref, dq_ref, det_ref = rik.rmr(q0, x_des, iters=1)
    assert_close(dx, dx_ref, what="dx (axis-angle branch)")
    # near targets: Euler-error branch (< 1 deg)
    x_here, _ = rik.task_kinematics(q0)
    q_near = q0 + 1e-3
    _, dx_ref, _, _ = rik.rmr(q_near, x_here, iters=1)
    assert_close(o.ik_task_error(q_near, tasks, x_here), dx_ref, what="dx (Euler-error branch)")


def test_oracle_ik_with_coupled_joints_vs_reference_library(refmod):
    """Reduced joint space of RcsGraph_coupledJointMatrix (Rcs_graph.c:2823, IkSolverRMR.cpp:
    424-443) on the WAM scenario (coupled finger joints)."""
    lay = layout("wam_scenario")
    o = rcs_oracle.Oracle(lay)
    names = [b["name"] for b in lay["bodies"]]
    tip = names.index("FingerTipTip2")
    rik = refmod.RefIk(graph_file("wam_scenario"), [("XYZ", "FingerTipTip2"), ("ABC", "FingerTipTip2")])
    tasks = [("XYZ", tip), ("ABC", tip)]
    q_goal = workload.sample_states(lay, 60, first=41, margin=0.2)
    x_des, _ = rik.task_kinematics(q_goal)
    q0 = q_goal + 0.05 * np.random.default_rng(3).uniform(-1, 1, size=q_goal.shape)
    q_ref, dx_ref, dq_ref, det_ref = rik.rmr(q0, x_des, iters=1)
    q1, dx1, dq1, det1 = o.ik_rmr(q0, tasks, x_des, iters=1)
    assert_close(dx1, dx_ref, what="dx")
    ok = det_ref > 1e-12
    assert ok.mean() > 0.8
    scale = np.maximum(1.0, np.abs(dq_ref).max(axis=1))
    assert np.all(np.abs(dq1 - dq_ref).max(axis=1)[ok] <= 1e-7 * scale[ok])   # ill-conditioned fingers
    well = ok & (det_ref > np.median(det_ref))
    assert_close(q1[well], q_ref[well], tol=1e-9, what="q after one step (coupled)")


@pytest.mark.parametrize("name,eff", [("gantry", "Tool"), ("wam_arm", "BallTip"), ("wam_scenario", "FingerTipTip2")])
def test_oracle_left_inverse_vs_reference_library(name, eff, refmod):
    """The numpy restatement of IkSolverRMR::solveLeftInverse (IkSolverRMR.cpp:205-272) against the
    reference library. lambda = 1e-3 keeps the nq x nq matrix of a redundant arm well conditioned."""
    lay = layout(name)
    o = rcs_oracle.Oracle(lay)
    tip = [b["name"] for b in lay["bodies"]].index(eff)
    rik = refmod.RefIk(graph_file(name), [("XYZ", eff), ("ABC", eff)])
    q_goal = workload.sample_states(lay, 40, first=9, margin=0.2)
    x_des, _ = rik.task_kinematics(q_goal)
    q0 = q_goal + 0.05
    for iters in (1, 4):
        qa, dxa, dqa, deta = rik.rmr(q0, x_des, lam=1.0e-3, iters=iters, left=True)
        qb, dxb, dqb, detb = o.ik_rmr(q0, [("XYZ", tip), ("ABC", tip)], x_des, lam=1.0e-3, iters=iters, left=True)
        ok = deta > 0.0
        assert ok.mean() > 0.9
        assert np.abs(dqa - dqb)[ok].max() < 1e-9 and np.abs(qa - qb)[ok].max() < 1e-9
        assert np.abs(deta[ok] / detb[ok] - 1.0).max() < 1e-9
        assert np.abs(dxa - dxb)[ok].max() < 1e-9
