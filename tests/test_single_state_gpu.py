"""GPU parity of the single-state reference API (include/rcsb.h, bottom): RcsGraph_setState,
RcsGraph_computeForwardKinematics, RcsGraph_bodyPointJacobian, RcsGraph_rotationJacobian and
RcsGraph_computeKineticTerms called the way reference code calls them (MatNd in, graph
structures / MatNd out), against the golden vectors recorded from the reference.

This is BASELINE.json configs[0]: "7-DoF WAM arm: single-state FK + end-effector world-point /
rotation Jacobian" (ExampleForwardKinematics / TestGraph plumbing).
"""
import ctypes

import numpy as np
import pytest

import rcs_b200
from conftest import ROOT, assert_close, graph_file

pytestmark = pytest.mark.gpu


class HTr(ctypes.Structure):
    _fields_ = [("org", ctypes.c_double * 3), ("rot", (ctypes.c_double * 3) * 3)]


class MatNd(ctypes.Structure):
    _fields_ = [("ele", ctypes.POINTER(ctypes.c_double)), ("m", ctypes.c_uint), ("n", ctypes.c_uint),
                ("size", ctypes.c_uint), ("stackMem", ctypes.c_bool)]


class RcsBody(ctypes.Structure):
    _fields_ = [("m", ctypes.c_double), ("rigid_body_joints", ctypes.c_bool),
                ("physicsSim", ctypes.c_int), ("x_dot", ctypes.c_double * 3),
                ("omega", ctypes.c_double * 3), ("confidence", ctypes.c_double), ("name", ctypes.c_char_p),
                ("xmlName", ctypes.c_char_p), ("suffix", ctypes.c_char_p),
                ("A_BP", ctypes.POINTER(HTr)), ("A_BI", ctypes.POINTER(HTr))]


def _lib():
    L = rcs_b200.lib()
    L.MatNd_create.restype = ctypes.POINTER(MatNd)
    L.RcsGraph_getBodyByIndex.restype = ctypes.POINTER(RcsBody)
    L.RcsGraph_getBodyByIndex.argtypes = [ctypes.c_void_p, ctypes.c_uint]
    L.RcsGraph_setState.argtypes = [ctypes.c_void_p, ctypes.POINTER(MatNd), ctypes.POINTER(MatNd)]
    L.RcsGraph_computeForwardKinematics.argtypes = [ctypes.c_void_p, ctypes.POINTER(MatNd),
                                                    ctypes.POINTER(MatNd)]
    L.RcsGraph_computeForwardKinematics.restype = None
    L.RcsGraph_bodyPointJacobian.argtypes = [ctypes.c_void_p, ctypes.POINTER(RcsBody),
                                             ctypes.POINTER(ctypes.c_double), ctypes.c_void_p,
                                             ctypes.POINTER(MatNd)]
    L.RcsGraph_bodyPointJacobian.restype = None
    L.RcsGraph_rotationJacobian.argtypes = [ctypes.c_void_p, ctypes.POINTER(RcsBody), ctypes.c_void_p,
                                            ctypes.POINTER(MatNd)]
    L.RcsGraph_rotationJacobian.restype = None
    L.RcsGraph_computeKineticTerms.argtypes = [ctypes.c_void_p] + [ctypes.POINTER(MatNd)] * 3
    L.RcsGraph_computeKineticTerms.restype = ctypes.c_double
    L.MatNd_destroy.argtypes = [ctypes.POINTER(MatNd)]
    L.RcsGraph_computeMassMatrix.argtypes = [ctypes.c_void_p, ctypes.POINTER(MatNd)]
    L.RcsGraph_computeMassMatrix.restype = None
    L.RcsGraph_computeGravityTorque.argtypes = [ctypes.c_void_p, ctypes.POINTER(MatNd)]
    L.RcsGraph_computeGravityTorque.restype = None
    L.RcsGraph_COG.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_double)]
    L.RcsGraph_COG.restype = ctypes.c_double
    L.RcsGraph_COGJacobian.argtypes = [ctypes.c_void_p, ctypes.POINTER(MatNd)]
    L.RcsGraph_COGJacobian.restype = None
    return L


def _mat(L, values=None, m=1, n=1):
    if values is not None:
        values = np.asarray(values, dtype=np.float64).reshape(-1)
        m, n = len(values), 1
    M = L.MatNd_create(m, n)
    if values is not None:
        for i, v in enumerate(values):
            M.contents.ele[i] = v
    return M


def _arr(M):
    return np.array([M.contents.ele[i] for i in range(M.contents.m * M.contents.n)]).reshape(
        M.contents.m, M.contents.n)


def _frames(L, g):
    out = []
    for b in range(g.n_bodies):
        body = L.RcsGraph_getBodyByIndex(g.handle, b).contents
        A = body.A_BI.contents
        out.append(list(A.org) + [A.rot[i][j] for i in range(3) for j in range(3)])
    return np.array(out)


@pytest.mark.parametrize("name", ["wam_arm", "wam_scenario", "humanoid"])
def test_reference_call_sequence_matches_golden(name, cuda_device):
    L = _lib()
    z = np.load(f"{ROOT}/tests/golden/{name}.npz")
    g = rcs_b200.Graph(graph_file(name))
    nJ = g.nJ
    J = _mat(L, m=3, n=nJ)
    Mm, h, Fg = _mat(L, m=nJ, n=nJ), _mat(L, m=nJ, n=1), _mat(L, m=nJ, n=1)
    for s in range(4):
        q, qd = _mat(L, z["q"][s]), _mat(L, z["qd"][s])
        L.RcsGraph_setState(g.handle, q, qd)
        assert_close(_frames(L, g), z["fk_A_BI"][s], what=f"{name}: A_BI of state {s}")
        for e, b in enumerate(z["eff"]):
            body = L.RcsGraph_getBodyByIndex(g.handle, int(b))
            vel = np.array(list(body.contents.x_dot) + list(body.contents.omega))
            assert_close(vel, np.concatenate([z["fk_xdot"][s, b], z["fk_omega"][s, b]]), what="x_dot, omega")
            pt = (ctypes.c_double * 3)(*z["pts"][e])
            L.RcsGraph_bodyPointJacobian(g.handle, body, pt, None, J)
            assert (J.contents.m, J.contents.n) == (3, nJ)
            assert_close(_arr(J), z["jac_JT"][s, e], what=f"{name}: JT")
            L.RcsGraph_rotationJacobian(g.handle, body, None, J)
            assert_close(_arr(J), z["jac_JR"][s, e], what=f"{name}: JR")
        E = L.RcsGraph_computeKineticTerms(g.handle, Mm, h, Fg)
        assert_close(_arr(Mm), z["kt_M"][s], what=f"{name}: M")
        assert_close(_arr(h)[:, 0], z["kt_h"][s], what=f"{name}: h")
        assert_close(_arr(Fg)[:, 0], z["kt_Fg"][s], what=f"{name}: F_gravity")
        assert abs(E - z["kt_E"][s]) <= 1e-12 * max(1.0, abs(z["kt_E"][s]))
        # reference invariants (bin/TestGraph.cpp:314-352): the separate entry points agree
        L.RcsGraph_computeMassMatrix(g.handle, Mm)
        assert_close(_arr(Mm), z["kt_M"][s], what=f"{name}: computeMassMatrix")
        L.RcsGraph_computeGravityTorque(g.handle, Fg)
        assert_close(_arr(Fg)[:, 0], z["kt_Fg"][s], tol=1e-11, what=f"{name}: computeGravityTorque")
        r_cog = (ctypes.c_double * 3)()
        mass = L.RcsGraph_COG(g.handle, r_cog)
        assert abs(mass - g.mass()) < 1e-12 * max(1.0, mass)
        assert_close(np.array(list(r_cog)), z["cog"][s], what=f"{name}: COG")
        L.RcsGraph_COGJacobian(g.handle, J)
        assert_close(_arr(J), z["Jcog"][s], what=f"{name}: COG Jacobian")
        L.MatNd_destroy(q)
        L.MatNd_destroy(qd)
    # NULL body = world frame: zero Jacobian (Rcs_kinematics.c:188-191)
    L.RcsGraph_bodyPointJacobian(g.handle, None, None, None, J)
    assert np.all(_arr(J) == 0.0)


def test_set_state_from_ik_vector_and_coupling_flag(cuda_device):
    """q of nJ rows is expanded (Rcs_graph.c:628); the return value reports whether a slave joint
    had to be moved (Rcs_graph.c:2756)."""
    L = _lib()
    L.RcsGraph_setState.restype = ctypes.c_bool
    g = rcs_b200.Graph(graph_file("wam_scenario"))
    lay = g.layout()
    z = np.load(f"{ROOT}/tests/golden/wam_scenario.npz")
    cols = [j["jointIndex"] for j in lay["joints"] if j["jacobiIndex"] >= 0]
    q_ik = _mat(L, z["fk_q"][1][cols])
    changed = L.RcsGraph_setState(g.handle, q_ik, None)
    assert not changed                     # fk_q is already consistent with the coupling rules
    assert_close(_frames(L, g), z["fk_A_BI"][1], what="A_BI from an IK state vector")
    slave = next(j for j in lay["joints"] if j["coupledTo"] >= 0 and not j["constrained"])
    q = z["fk_q"][1].copy()
    q[slave["jointIndex"]] += 0.1
    assert L.RcsGraph_setState(g.handle, _mat(L, q), None)
    assert_close(_frames(L, g), z["fk_A_BI"][1], what="slave joint overwritten from its master")


def test_compute_forward_kinematics_takes_slaves_as_given(cuda_device, refmod):
    """RcsGraph_computeForwardKinematics does not apply the coupling rules (Rcs_graph.c:301)."""
    L = _lib()
    g = rcs_b200.Graph(graph_file("wam_arm"))
    z = np.load(f"{ROOT}/tests/golden/wam_arm.npz")
    L.RcsGraph_computeForwardKinematics(g.handle, _mat(L, z["q"][2]), _mat(L, z["qd"][2]))
    assert_close(_frames(L, g), z["fk_A_BI"][2], what="A_BI")


def test_wrong_dimensions_are_reported(cuda_device):
    L = _lib()
    g = rcs_b200.Graph(graph_file("wam_arm"))
    bad = _mat(L, np.zeros(5))
    L.RcsGraph_setState.restype = ctypes.c_bool
    assert L.RcsGraph_setState(g.handle, bad, None) is False
    assert b"wrong dimensions" in L.rcsb_last_error()


def test_relative_jacobian_wrappers(cuda_device, refmod):
    """RcsGraph_3dPosJacobian / RcsGraph_3dOmegaJacobian with body pointers, as reference code
    calls them (Rcs_kinematics.c:566, :789)."""
    L = _lib()
    for fn in (L.RcsGraph_3dPosJacobian, L.RcsGraph_3dOmegaJacobian):
        fn.argtypes = [ctypes.c_void_p] + [ctypes.POINTER(RcsBody)] * 3 + [ctypes.POINTER(MatNd)]
        fn.restype = None
    g = rcs_b200.Graph(graph_file("humanoid"))
    r = refmod.RefGraph(graph_file("humanoid"))
    z = np.load(f"{ROOT}/tests/golden/humanoid.npz")
    names = g.body_names()
    a, b, c = names.index("RightHandTip"), names.index("LeftHandTip"), names.index("UpperBody")
    ref = r.task_jacobians(z["q"][:1], [("POS", a, b, c), ("OMEGA", a, b, c), ("POS", a, -1, -1)])
    L.RcsGraph_setState(g.handle, _mat(L, z["q"][0]), None)
    body = lambda i: L.RcsGraph_getBodyByIndex(g.handle, i)   # noqa: E731
    J = _mat(L, m=3, n=g.nJ)
    L.RcsGraph_3dPosJacobian(g.handle, body(a), body(b), body(c), J)
    assert_close(_arr(J), ref[0, 0], what="3dPosJacobian")
    L.RcsGraph_3dOmegaJacobian(g.handle, body(a), body(b), body(c), J)
    assert_close(_arr(J), ref[0, 1], what="3dOmegaJacobian")
    L.RcsGraph_3dPosJacobian(g.handle, body(a), None, None, J)
    assert_close(_arr(J), ref[0, 2], what="3dPosJacobian (world)")


def test_jacobians_in_a_rotated_frame_world_point_and_body_point(cuda_device):
    """A_BI != NULL rotates the columns into that frame (MatNd_rotateSelf, Rcs_kinematics.c:160);
    RcsGraph_worldPointJacobian (:178) takes the point in world coordinates;
    RcsGraph_bodyPoint (:47) maps a body-fixed point with the frames of the last FK."""
    L = _lib()
    Rot = (ctypes.c_double * 3) * 3
    L.RcsGraph_bodyPointJacobian.argtypes = [ctypes.c_void_p, ctypes.POINTER(RcsBody),
                                             ctypes.POINTER(ctypes.c_double), ctypes.POINTER(Rot),
                                             ctypes.POINTER(MatNd)]
    L.RcsGraph_rotationJacobian.argtypes = [ctypes.c_void_p, ctypes.POINTER(RcsBody), ctypes.POINTER(Rot),
                                            ctypes.POINTER(MatNd)]
    L.RcsGraph_worldPointJacobian.argtypes = L.RcsGraph_bodyPointJacobian.argtypes
    L.RcsGraph_worldPointJacobian.restype = None
    L.RcsGraph_bodyPoint.argtypes = [ctypes.POINTER(RcsBody), ctypes.POINTER(ctypes.c_double),
                                     ctypes.POINTER(ctypes.c_double)]
    L.RcsGraph_bodyPoint.restype = None
    z = np.load(f"{ROOT}/tests/golden/humanoid.npz")
    g = rcs_b200.Graph(graph_file("humanoid"))
    J = _mat(L, m=3, n=g.nJ)
    L.RcsGraph_setState(g.handle, _mat(L, z["q"][1]), None)
    e = 0
    b = int(z["eff"][e])
    body = L.RcsGraph_getBodyByIndex(g.handle, b)
    frame = z["fk_A_BI"][1, b]
    A = frame[3:].reshape(3, 3)
    rot = Rot(*[(ctypes.c_double * 3)(*row) for row in A])
    pt = (ctypes.c_double * 3)(*z["pts"][e])
    L.RcsGraph_bodyPointJacobian(g.handle, body, pt, ctypes.byref(rot), J)
    assert_close(_arr(J), A @ z["jac_JT"][1, e], what="JT in the body frame")
    L.RcsGraph_rotationJacobian(g.handle, body, ctypes.byref(rot), J)
    assert_close(_arr(J), A @ z["jac_JR"][1, e], what="JR in the body frame")
    I_pt = (ctypes.c_double * 3)()
    L.RcsGraph_bodyPoint(body, pt, I_pt)
    want = frame[:3] + A.T @ z["pts"][e]
    assert_close(np.array(list(I_pt)), want, what="RcsGraph_bodyPoint")
    L.RcsGraph_worldPointJacobian(g.handle, body, I_pt, None, J)
    assert_close(_arr(J), z["jac_JT"][1, e], what="worldPointJacobian of the same point")
    L.RcsGraph_worldPointJacobian(g.handle, body, None, None, J)
    L.RcsGraph_bodyPointJacobian(g.handle, body, None, None, _mat(L, m=3, n=g.nJ))
    origin = (ctypes.c_double * 3)()
    L.RcsGraph_bodyPoint(body, None, origin)
    assert_close(np.array(list(origin)), frame[:3], what="body origin")
    L.RcsGraph_bodyPoint(None, pt, origin)
    assert list(origin) == list(pt)


def test_compute_body_kinematics_updates_only_the_subtree(cuda_device):
    """RcsGraph_computeBodyKinematics (Rcs_graph.c:327): the body (or its subtree) takes the
    frames of the explicit state, every other body keeps those of the last full pass."""
    L = _lib()
    L.RcsGraph_computeBodyKinematics.argtypes = [ctypes.c_void_p, ctypes.POINTER(RcsBody),
                                                 ctypes.POINTER(MatNd), ctypes.POINTER(MatNd), ctypes.c_bool]
    L.RcsGraph_computeBodyKinematics.restype = None
    z = np.load(f"{ROOT}/tests/golden/humanoid.npz")
    g = rcs_b200.Graph(graph_file("humanoid"))
    lay = g.layout()
    names = g.body_names()
    L.RcsGraph_computeForwardKinematics(g.handle, _mat(L, z["fk_q"][0]), _mat(L, z["qd"][0]))
    old = _frames(L, g)
    assert_close(old, z["fk_A_BI"][0], what="state 0")
    root = names.index("RightSh0")
    parents = [b["parent"] for b in lay["bodies"]]

    def in_subtree(i):
        while i >= 0:
            if i == root:
                return True
            i = parents[i]
        return False

    sub = np.array([in_subtree(i) for i in range(len(names))])
    assert sub.sum() > 1 and (~sub).sum() > 1
    L.RcsGraph_computeBodyKinematics(g.handle, L.RcsGraph_getBodyByIndex(g.handle, root),
                                     _mat(L, z["fk_q"][2]), _mat(L, z["qd"][2]), True)
    new = _frames(L, g)
    assert_close(new[sub], z["fk_A_BI"][2][sub], what="subtree at state 2")
    assert np.array_equal(new[~sub], old[~sub])
    L.RcsGraph_computeBodyKinematics(g.handle, L.RcsGraph_getBodyByIndex(g.handle, root),
                                     _mat(L, z["fk_q"][3]), None, False)
    one = _frames(L, g)
    assert_close(one[root], z["fk_A_BI"][3][root], what="single body at state 3")
    keep = np.arange(len(names)) != root
    assert np.array_equal(one[keep], new[keep])
