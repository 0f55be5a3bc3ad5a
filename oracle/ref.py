"""ctypes wrapper of oracle/_ref/librcsref.so — the UNMODIFIED reference (HRI-EU/Rcs RcsCore)
compiled by oracle/ref_build/Makefile, driven one state at a time.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the cpu_baseline /
--impl reference legs of bench.py. Nothing under rcs_b200/ may import this module.

Every function returns exactly what the reference functions produce (see the citations in
oracle/ref_build/ref_api.cpp); this is the anchor the numpy restatement (oracle/rcs_oracle.py)
and the CUDA path are pinned against.
"""
from __future__ import annotations

import ctypes
import json
import os
from typing import Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def library_path() -> str:
    return os.path.join(_HERE, "_ref", "librcsref.so")


def available() -> bool:
    return os.path.exists(library_path())


def lib() -> ctypes.CDLL:
    global _LIB
    if _LIB is None:
        L = ctypes.CDLL(library_path())
        vp, ci, cl, cd = ctypes.c_void_p, ctypes.c_int, ctypes.c_long, ctypes.c_double
        L.ref_graph_create.restype = vp
        L.ref_graph_create.argtypes = [ctypes.c_char_p]
        L.ref_graph_destroy.argtypes = [vp]
        L.ref_graph_layout_json.restype = vp
        L.ref_graph_layout_json.argtypes = [vp]
        L.ref_free.argtypes = [vp]
        for n in ("ref_graph_dof", "ref_graph_nj", "ref_graph_nbodies"):
            getattr(L, n).restype = ci
            getattr(L, n).argtypes = [vp]
        L.ref_body_index.restype = ci
        L.ref_body_index.argtypes = [vp, ctypes.c_char_p]
        L.ref_graph_check.restype = ci
        L.ref_graph_check.argtypes = [vp, ctypes.POINTER(ci), ctypes.POINTER(ci)]
        L.ref_fk.restype = cd
        L.ref_fk.argtypes = [vp, cl, ci] + [vp] * 7
        L.ref_fk_jacobians.restype = cd
        L.ref_fk_jacobians.argtypes = [vp, cl, ci, vp, ci, vp, vp, vp, vp, vp]
        L.ref_fk_jacobians_mass.restype = cd
        L.ref_fk_jacobians_mass.argtypes = [vp, cl, ci, vp, ci, vp, vp, vp, vp, vp, vp]
        L.ref_ik_wholebody.restype = cd
        L.ref_ik_wholebody.argtypes = [vp, cl, ci, vp, vp, vp, vp]
        L.ref_cog.restype = cd
        L.ref_cog.argtypes = [vp, cl, ci, vp, vp, vp]
        L.ref_task_jacobians.restype = cd
        L.ref_task_jacobians.argtypes = [vp, cl, ci, vp, ci, vp, vp]
        L.ref_direct_dynamics.restype = cd
        L.ref_direct_dynamics.argtypes = [vp, cl, ci] + [vp] * 6
        L.ref_kinetic_terms.restype = cd
        L.ref_kinetic_terms.argtypes = [vp, cl, ci] + [vp] * 6
        L.ref_mass_matrix.restype = cd
        L.ref_mass_matrix.argtypes = [vp, cl, ci, vp, vp]
        L.ref_inv_wq.argtypes = [vp, vp]
        L.ref_joint_limit_gradient.restype = cd
        L.ref_joint_limit_gradient.argtypes = [vp, cl, ci, vp, vp, vp]
        L.ref_ik_create.restype = vp
        L.ref_ik_create.argtypes = [ctypes.c_char_p, ctypes.c_char_p]
        L.ref_ik_destroy.argtypes = [vp]
        L.ref_ik_task_dim.restype = ci
        L.ref_ik_task_dim.argtypes = [vp]
        L.ref_ik_graph.restype = vp
        L.ref_ik_graph.argtypes = [vp]
        L.ref_ik_task_kinematics.restype = cd
        L.ref_ik_task_kinematics.argtypes = [vp, cl, ci, vp, vp, vp]
        L.ref_ik_rmr.restype = cd
        L.ref_ik_rmr.argtypes = [vp, cl, vp, vp, cd, cd, ci, vp, vp, vp]
        L.ref_ik_rmr_left.restype = cd
        L.ref_ik_rmr_left.argtypes = [vp, cl, vp, vp, cd, cd, ci, vp, vp, vp]
        L.ref_ik_constraint_rmr.restype = cd
        L.ref_ik_constraint_rmr.argtypes = [vp, cl, vp, vp, cd, cd, ci, vp, vp, vp]
        L.ref_ik_rmr_ex.restype = cd
        L.ref_ik_rmr_ex.argtypes = [vp, cl, vp, vp, vp, ci, vp, ci, cd, cd, ci, ci, ci, vp, vp, vp, vp]
        L.ref_ik_prio_rmr.restype = cd
        L.ref_ik_prio_rmr.argtypes = [vp, cl, vp, vp, vp, vp, cd, cd, ci, vp, vp, vp, vp]
        L.ref_set_threads.argtypes = [ci]
        L.ref_set_log_level.argtypes = [ci]
        _LIB = L
    return _LIB


def set_threads(n: int) -> None:
    lib().ref_set_threads(int(n))


def _p(a: Optional[np.ndarray]):
    if a is None:
        return None
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
    return ctypes.c_void_p(a.ctypes.data)


def _c(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float64)


class RefGraph:
    """RcsGraph created by the reference's own loader (src/RcsCore/Rcs_graph.c:398)."""

    def __init__(self, xml_file: str, _handle=None, _owner=None):
        L = lib()
        self._owner = _owner
        self._h = _handle if _handle is not None else L.ref_graph_create(xml_file.encode())
        if not self._h:
            raise RuntimeError(f"reference failed to load {xml_file}")
        self.dof = L.ref_graph_dof(self._h)
        self.nJ = L.ref_graph_nj(self._h)
        self.n_bodies = L.ref_graph_nbodies(self._h)
        self.last_seconds = 0.0

    def layout(self) -> dict:
        L = lib()
        p = L.ref_graph_layout_json(self._h)
        s = ctypes.string_at(p).decode()
        L.ref_free(p)
        return json.loads(s)

    def clone(self) -> "RefGraph":
        """RcsGraph_clone (Rcs_graph.c:2193)"""
        L = lib()
        L.ref_graph_clone.restype = ctypes.c_void_p
        L.ref_graph_clone.argtypes = [ctypes.c_void_p]
        return RefGraph("", _handle=L.ref_graph_clone(self._h))

    def body_index(self, name: str) -> int:
        return lib().ref_body_index(self._h, name.encode())

    def check(self):
        ne, nw = ctypes.c_int(0), ctypes.c_int(0)
        ok = lib().ref_graph_check(self._h, ctypes.byref(ne), ctypes.byref(nw))
        return bool(ok), ne.value, nw.value

    # RcsGraph_setState (Rcs_graph.c:231)
    def fk(self, q, qd=None, want=("A_BI",)):
        q = _c(q)
        qd = None if qd is None else _c(qd)
        N, qdim = q.shape
        out = {}
        if "A_BI" in want:
            out["A_BI"] = np.empty((N, self.n_bodies, 12))
        if "A_JI" in want:
            out["A_JI"] = np.empty((N, self.dof, 12))
        if "xdot" in want:
            out["xdot"] = np.empty((N, self.n_bodies, 3))
        if "omega" in want:
            out["omega"] = np.empty((N, self.n_bodies, 3))
        if "q" in want:
            out["q"] = np.empty((N, self.dof))
        self.last_seconds = lib().ref_fk(
            self._h, N, qdim, _p(q), _p(qd), _p(out.get("A_BI")), _p(out.get("A_JI")),
            _p(out.get("xdot")), _p(out.get("omega")), _p(out.get("q")))
        return out

    # + RcsGraph_bodyPointJacobian / RcsGraph_rotationJacobian (Rcs_kinematics.c:200, :416)
    def fk_jacobians(self, q, effectors: Sequence[int], points=None, want=("A_BI", "JT", "JR")):
        q = _c(q)
        N, qdim = q.shape
        ne = len(effectors)
        eff = (ctypes.c_int * max(ne, 1))(*effectors)
        pts = None if points is None else _c(np.asarray(points).reshape(ne, 3))
        out = {}
        if "A_BI" in want:
            out["A_BI"] = np.empty((N, self.n_bodies, 12))
        if "JT" in want:
            out["JT"] = np.empty((N, ne, 3, self.nJ))
        if "JR" in want:
            out["JR"] = np.empty((N, ne, 3, self.nJ))
        if "M" in want:
            # + RcsGraph_computeMassMatrix on the same graph state: one RcsGraph_setState per state
            out["M"] = np.empty((N, self.nJ, self.nJ))
            self.last_seconds = lib().ref_fk_jacobians_mass(
                self._h, N, qdim, _p(q), ne, eff, _p(pts), _p(out.get("A_BI")), _p(out.get("JT")),
                _p(out.get("JR")), _p(out["M"]))
            return out
        self.last_seconds = lib().ref_fk_jacobians(
            self._h, N, qdim, _p(q), ne, eff, _p(pts), _p(out.get("A_BI")), _p(out.get("JT")),
            _p(out.get("JR")))
        return out

    def cog(self, q):
        q = _c(q)
        N, qdim = q.shape
        cog = np.empty((N, 3))
        J = np.empty((N, 3, self.nJ))
        self.last_seconds = lib().ref_cog(self._h, N, qdim, _p(q), _p(cog), _p(J))
        return cog, J

    # RcsGraph_3dPosJacobian / RcsGraph_3dOmegaJacobian (Rcs_kinematics.c:566, :789)
    def task_jacobians(self, q, spec):
        """spec: sequence of (kind "POS" | "OMEGA", effector, refBdy, refFrame) body indices,
        -1 for NULL. Returns J: N x len(spec) x 3 x nJ."""
        q = _c(q)
        N, qdim = q.shape
        flat = []
        for kind, eff, rb, rf in spec:
            flat += [1 if kind == "POS" else 2, int(eff), int(rb), int(rf)]
        arr = (ctypes.c_int * max(len(flat), 1))(*flat)
        J = np.empty((N, len(spec), 3, self.nJ))
        self.last_seconds = lib().ref_task_jacobians(self._h, N, qdim, _p(q), len(spec), arr, _p(J))
        return J

    # RcsGraph_computeKineticTerms (Rcs_dynamics.c:439)
    def kinetic_terms(self, q, qd=None, want=("M", "h", "Fg", "E")):
        q = _c(q)
        qd = None if qd is None else _c(qd)
        N, qdim = q.shape
        out = {}
        if "M" in want:
            out["M"] = np.empty((N, self.nJ, self.nJ))
        if "h" in want:
            out["h"] = np.empty((N, self.nJ))
        if "Fg" in want:
            out["Fg"] = np.empty((N, self.nJ))
        if "E" in want:
            out["E"] = np.empty((N,))
        self.last_seconds = lib().ref_kinetic_terms(
            self._h, N, qdim, _p(q), _p(qd), _p(out.get("M")), _p(out.get("h")),
            _p(out.get("Fg")), _p(out.get("E")))
        return out

    # Rcs_directDynamics (Rcs_dynamics.c:673)
    def direct_dynamics(self, q, qd, F_ext=None, F_jnt=None):
        q, qd = _c(q), _c(qd)
        N, qdim = q.shape
        fe = None if F_ext is None else _c(F_ext)
        fj = None if F_jnt is None else _c(F_jnt)
        qdd, E = np.empty((N, self.nJ)), np.empty((N,))
        self.last_seconds = lib().ref_direct_dynamics(self._h, N, qdim, _p(q), _p(qd), _p(fe), _p(fj),
                                                      _p(qdd), _p(E))
        return qdd, E

    def mass_matrix(self, q):
        q = _c(q)
        N, qdim = q.shape
        M = np.empty((N, self.nJ, self.nJ))
        self.last_seconds = lib().ref_mass_matrix(self._h, N, qdim, _p(q), _p(M))
        return M

    def inv_wq(self):
        w = np.empty((self.nJ,))
        lib().ref_inv_wq(self._h, _p(w))
        return w

    def joint_limit_gradient(self, q):
        q = _c(q)
        N, qdim = q.shape
        dH = np.empty((N, self.nJ))
        cost = np.empty((N,))
        lib().ref_joint_limit_gradient(self._h, N, qdim, _p(q), _p(dH), _p(cost))
        return dH, cost


class RefIk:
    """ControllerBase + IkSolverRMR of the reference with XYZ / ABC tasks, all active."""

    def __init__(self, graph_file: str, tasks: Sequence):
        """tasks: sequence of (controlVariable, effector body name[, refBdy name[, refFrame name]]);
        any controlVariable of the reference's TaskFactory (XYZ, ABC, XYZABC, COGX, ...)."""
        L = lib()
        spec = ";".join(":".join(str(x) for x in t) for t in tasks)
        self._h = L.ref_ik_create(graph_file.encode(), spec.encode())
        if not self._h:
            raise RuntimeError("reference controller creation failed")
        self.nx = L.ref_ik_task_dim(self._h)
        self.graph = RefGraph("", _handle=L.ref_ik_graph(self._h), _owner=self)
        self.last_seconds = 0.0

    def task_kinematics(self, q, want_x=True):
        q = _c(q)
        N, qdim = q.shape
        x = np.empty((N, self.nx)) if want_x else None
        J = np.empty((N, self.nx, self.graph.nJ))
        self.last_seconds = lib().ref_ik_task_kinematics(self._h, N, qdim, _p(q), _p(x) if want_x else None,
                                                         _p(J))
        return x, J

    def wholebody(self, q, want=("A_BI", "J", "M")):
        """RcsGraph_setState + ControllerBase::computeJ + RcsGraph_computeMassMatrix, one pass per
        state on the controller's graph."""
        q = _c(q)
        N, qdim = q.shape
        g = self.graph
        out = {}
        if "A_BI" in want:
            out["A_BI"] = np.empty((N, g.n_bodies, 12))
        if "J" in want:
            out["J"] = np.empty((N, self.nx, g.nJ))
        if "M" in want:
            out["M"] = np.empty((N, g.nJ, g.nJ))
        self.last_seconds = lib().ref_ik_wholebody(self._h, N, qdim, _p(q), _p(out.get("A_BI")),
                                                   _p(out.get("J")), _p(out.get("M")))
        return out

    def rmr_ex(self, q, x_des, activation=None, scale_dx=False, lambda_rows=None, lam=1.0e-8, alpha=0.05,
               iters=1, left=False, blending=0):
        """The reference's full solver interface: IkSolverRMR::solveRightInverse / solveLeftInverse
        (dq_ts, dq_ns, dx, dH, activation, lambda | lambda vector) with ControllerBase::computeDX
        (dx, x_des[, activation]). Returns dict(q, dx [active rows], dq_ts, dq_ns, det)."""
        q = _c(q).copy()
        x_des = _c(x_des)
        N = q.shape[0]
        act = None if activation is None else _c(np.asarray(activation, dtype=np.float64))
        lr = None if lambda_rows is None else _c(np.asarray(lambda_rows, dtype=np.float64))
        dx = np.zeros((N, self.nx))
        ts, ns = np.empty((N, self.graph.dof)), np.empty((N, self.graph.dof))
        det = np.empty((N,))
        self.last_seconds = lib().ref_ik_rmr_ex(
            self._h, N, _p(q), _p(x_des), _p(act), int(bool(scale_dx)), _p(lr), 0 if lr is None else len(lr),
            float(lam), float(alpha), int(iters), int(bool(left)), int(blending), _p(dx), _p(ts), _p(ns), _p(det))
        return dict(q=q, dx=dx, dq_ts=ts, dq_ns=ns, det=det)

    def prio_rmr(self, q, x_des, priority, activation=None, lam=1.0e-8, alpha=0.05, iters=1):
        """IkSolverPrioRMR::solve (IkSolverPrioRMR.cpp:178) looped; returns dict(q, dq1, dq2, dq_ns, ok)."""
        q = _c(q).copy()
        x_des = _c(x_des)
        N = q.shape[0]
        prio = (ctypes.c_int * len(priority))(*[int(p) for p in priority])
        act = None if activation is None else _c(np.asarray(activation, dtype=np.float64))
        d1, d2, dn = (np.empty((N, self.graph.dof)) for _ in range(3))
        ok = np.empty((N,))
        self.last_seconds = lib().ref_ik_prio_rmr(self._h, N, _p(q), _p(x_des), prio, _p(act), float(lam),
                                                  float(alpha), int(iters), _p(d1), _p(d2), _p(dn), _p(ok))
        return dict(q=q, dq1=d1, dq2=d2, dq_ns=dn, ok=ok)

    def rmr(self, q, x_des, lam=1.0e-8, alpha=0.05, iters=1, left=False, constraint=False):
        """Returns (q_after, dx_last, dq_last, det_last); loop of examples/ExampleIK.cpp with
        solveRightInverse, or solveLeftInverse (IkSolverRMR.cpp:205) if `left`, or
        IkSolverConstraintRMR::solveRightInverse (IkSolverConstraintRMR.cpp:156) if `constraint`."""
        q = _c(q).copy()
        x_des = _c(x_des)
        N = q.shape[0]
        dx = np.empty((N, self.nx))
        dq = np.empty((N, self.graph.dof))
        det = np.empty((N,))
        fn = lib().ref_ik_rmr_left if left else lib().ref_ik_rmr
        if constraint:
            fn = lib().ref_ik_constraint_rmr
        self.last_seconds = fn(
            self._h, N, _p(q), _p(x_des), float(lam), float(alpha), int(iters), _p(dx), _p(dq),
            _p(det))
        return q, dx, dq, det
