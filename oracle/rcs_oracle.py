"""CPU restatement (numpy) of the reference algorithms on the hot path of HRI-EU/Rcs.

TEST INFRASTRUCTURE ONLY. Imported by tests/, __graft_entry__.smoke() and nothing under
rcs_b200/. It restates, state by state but vectorised over the batch, WHAT THE REFERENCE DOES
(not what the CUDA kernels do): pointer-order tree walk, backward joint-chain Jacobians,
sum-over-bodies mass matrix, Hessian-based Coriolis vector, explicit Cholesky inverse in the
IK step. Each function cites the reference file:line it follows.

Pinned: tests/test_oracle_cpu.py checks this module against
  - the reference library itself (oracle/_ref/librcsref.so, unmodified Rcs sources) where it
    is available, and
  - the golden vectors tests/golden/*.npz that tools/make_golden.py recorded from that
    library (they travel to machines without /root/reference),
  - the known answers of SURVEY.md §8c.

Input: the layout dict of a graph (RcsGraph_layoutJSON of rcs_b200 or ref_graph_layout_json
of the reference library — both are compared bit for bit in the tests).
"""
from __future__ import annotations

import numpy as np

GRAVITY = 9.81          # RCS_GRAVITY, src/RcsCore/Rcs_typedef.h:46
FLT_EPSILON = 1.1920928955078125e-07


def _cross(a, b):
    return np.cross(a, b)


def _htr(v):
    v = np.asarray(v, dtype=np.float64)
    return v[:3].copy(), v[3:].reshape(3, 3).copy()


class Oracle:
    def __init__(self, layout: dict):
        self.lay = layout
        self.bodies = layout["bodies"]
        self.joints = layout["joints"]
        self.nb = layout["nBodies"]
        self.dof = layout["dof"]
        self.nJ = layout["nJ"]
        self.q_ref = np.array(layout["q"], dtype=np.float64)
        self.jac = np.array([j["jacobiIndex"] for j in self.joints], dtype=int)
        self.is_rot = np.array([j["type"] <= 2 for j in self.joints], dtype=bool)
        self.dir = np.array([j["dirIdx"] for j in self.joints], dtype=int)
        # RcsBody_lastJointBeforeBody (src/RcsCore/Rcs_body.c:1706)
        self.last_jnt = []
        for b in range(self.nb):
            bb = b
            while bb >= 0 and not self.bodies[bb]["joints"]:
                bb = self.bodies[bb]["parent"]
            self.last_jnt.append(self.bodies[bb]["joints"][-1] if bb >= 0 else -1)

    # ------------------------------------------------------------------------------------
    # state vector handling
    # ------------------------------------------------------------------------------------
    def chain(self, body: int):
        """Backward joint chain of a body: lastJointBeforeBody, then jnt->prev links
        (src/RcsCore/Rcs_kinematics.c:163-171)."""
        j = self.last_jnt[body]
        while j >= 0:
            yield j
            j = self.joints[j]["prev"]

    def _poly(self, c, x):
        # RcsJoint_calcCouplingPolynomial (src/RcsCore/Rcs_joint.c:330)
        order = len(c) - 1
        return sum(c[i] * x ** (order - i) for i in range(order + 1))

    def _poly_d(self, c, x):
        # RcsJoint_calcCouplingPolynomialDerivative (src/RcsCore/Rcs_joint.c:350)
        order = len(c) - 1
        return sum((order - i) * c[i] * x ** (order - 1 - i) for i in range(order))

    def expand(self, q, qd=None):
        """RcsGraph_setState steps 1-2: IK -> full state (Rcs_graph.c:628), then
        RcsGraph_updateSlaveJoints (Rcs_graph.c:2756, Rcs_joint.c:375, :432)."""
        q = np.asarray(q, dtype=np.float64)
        n = q.shape[0]

        def full(x, fill):
            if x.shape[1] == self.dof:
                return x.copy()
            assert x.shape[1] == self.nJ, "q has wrong dimensions"
            out = np.tile(fill, (n, 1))
            for j in range(self.dof):
                if self.jac[j] >= 0:
                    out[:, j] = x[:, self.jac[j]]
            return out

        qf = full(q, self.q_ref)
        qdf = None if qd is None else full(np.asarray(qd, dtype=np.float64), np.zeros(self.dof))
        for j, jn in enumerate(self.joints):
            mj = jn["coupledTo"]
            if mj < 0:
                continue
            ms = self.joints[mj]
            c = jn["couplingFactors"]
            qm = qf[:, mj].copy()
            if len(c) == 1:
                qf[:, j] = jn["q_init"] + c[0] * (qm - ms["q_init"])
                if qdf is not None:
                    qdf[:, j] = c[0] * qdf[:, mj]
            else:
                lo, hi = ms["q_min"], ms["q_max"]
                inside = jn["q_init"] + self._poly(c, qm - ms["q_init"])
                below = jn["q_min"] - self._poly_d(c, lo - ms["q_init"]) * (lo - qm)
                above = jn["q_max"] + self._poly_d(c, hi - ms["q_init"]) * (qm - hi)
                qf[:, j] = np.where(qm < lo, below, np.where(qm > hi, above, inside))
                if qdf is not None:
                    qdf[:, j] = self._poly_d(c, np.clip(qm, lo, hi)) * qdf[:, mj]
        return qf, qdf

    # ------------------------------------------------------------------------------------
    # forward kinematics: RcsGraph_internalBodyKinematics (src/RcsCore/Rcs_graph.c:61-226)
    # ------------------------------------------------------------------------------------
    def fk(self, q, qd=None):
        qf, qdf = self.expand(q, qd)
        n = qf.shape[0]
        org = np.zeros((n, self.nb, 3))
        rot = np.zeros((n, self.nb, 3, 3))
        jorg = np.zeros((n, self.dof, 3))
        jrot = np.zeros((n, self.dof, 3, 3))
        xdot = np.zeros((n, self.nb, 3))
        omega = np.zeros((n, self.nb, 3))

        for b, body in enumerate(self.bodies):
            p = body["parent"]
            if p >= 0:
                o, R = org[:, p].copy(), rot[:, p].copy()
            else:
                o, R = np.zeros((n, 3)), np.tile(np.eye(3), (n, 1, 1))

            for j in body["joints"]:
                jn = self.joints[j]
                if jn["A_JP"] is not None:
                    a_org, a_rot = _htr(jn["A_JP"])
                    # org' = org + rot^T r ; rot' = A_JP.rot rot   (Rcs_graph.c:83-84)
                    o = o + np.einsum("nki,k->ni", R, a_org)
                    R = np.einsum("ik,nkj->nij", a_rot, R)
                qi = qf[:, j]
                d = self.dir[j]
                if not self.is_rot[j]:
                    o = o + qi[:, None] * R[:, d, :]                      # Rcs_graph.c:101
                else:
                    # Mat3d_rotateSelfAboutXYZAxis (src/RcsCore/Rcs_Mat3d.c:808-848)
                    c, s = np.cos(qi)[:, None], np.sin(qi)[:, None]
                    u, v = [(1, 2), (2, 0), (0, 1)][d]
                    ru, rv = R[:, u, :].copy(), R[:, v, :].copy()
                    R = R.copy()
                    R[:, u, :] = c * ru + s * rv
                    R[:, v, :] = -s * ru + c * rv
                jorg[:, j], jrot[:, j] = o, R

            if body["A_BP"] is not None:
                a_org, a_rot = _htr(body["A_BP"])
                o = o + np.einsum("nki,k->ni", R, a_org)                  # Rcs_graph.c:156-157
                R = np.einsum("ik,nkj->nij", a_rot, R)
            org[:, b], rot[:, b] = o, R

            if qdf is not None and p >= 0:                                # Rcs_graph.c:163-224
                xdot[:, b] = xdot[:, p] + _cross(omega[:, p], o - org[:, p])
                omega[:, b] = omega[:, p]
            if qdf is not None:
                for j in body["joints"]:
                    axis = jrot[:, j, self.dir[j], :]
                    if not self.is_rot[j]:
                        xdot[:, b] += axis * qdf[:, j, None]
                    else:
                        om_i = axis * qdf[:, j, None]
                        omega[:, b] += om_i
                        xdot[:, b] += _cross(om_i, o - jorg[:, j])

        out = {
            "A_BI": np.concatenate([org, rot.reshape(n, self.nb, 9)], axis=2),
            "A_JI": np.concatenate([jorg, jrot.reshape(n, self.dof, 9)], axis=2),
            "q": qf,
        }
        if qdf is not None:
            out["xdot"], out["omega"], out["qd"] = xdot, omega, qdf
        self._last = (org, rot, jorg, jrot)
        return out

    # ------------------------------------------------------------------------------------
    # Jacobians: src/RcsCore/Rcs_kinematics.c:80-215 (point), :416-458 (rotation)
    # ------------------------------------------------------------------------------------
    def _point_jacobian(self, body, I_p, jorg, jrot):
        n = I_p.shape[0]
        J = np.zeros((n, 3, self.nJ))
        for j in self.chain(body):
            if self.jac[j] < 0:
                continue
            axis = jrot[:, j, self.dir[j], :]
            J[:, :, self.jac[j]] = _cross(axis, I_p - jorg[:, j]) if self.is_rot[j] else axis
        return J

    def _rotation_jacobian(self, body, jrot, n):
        J = np.zeros((n, 3, self.nJ))
        for j in self.chain(body):
            if self.jac[j] >= 0 and self.is_rot[j]:
                J[:, :, self.jac[j]] = jrot[:, j, self.dir[j], :]
        return J

    def fk_jacobians(self, q, effectors, points=None):
        out = self.fk(q)
        org, rot, jorg, jrot = self._last
        n = org.shape[0]
        JT = np.zeros((n, len(effectors), 3, self.nJ))
        JR = np.zeros((n, len(effectors), 3, self.nJ))
        for e, b in enumerate(effectors):
            I_p = org[:, b].copy()
            if points is not None:                                        # Rcs_kinematics.c:47
                I_p = I_p + np.einsum("nki,k->ni", rot[:, b], np.asarray(points)[e])
            JT[:, e] = self._point_jacobian(b, I_p, jorg, jrot)
            JR[:, e] = self._rotation_jacobian(b, jrot, n)
        out["JT"], out["JR"] = JT, JR
        return out

    # ------------------------------------------------------------------------------------
    # kinetic terms: RcsGraph_computeKineticTerms (src/RcsCore/Rcs_dynamics.c:439-573)
    # ------------------------------------------------------------------------------------
    def _dot_jacobians_times_qd(self, body, I_p, jorg, jrot, qd_ik):
        """(Jdot_T qd, Jdot_R qd) through the analytic Hessians, as the reference does:
        RcsGraph_bodyPointHessian (Rcs_kinematics.c:311-390) and RcsGraph_rotationHessian
        (Rcs_kinematics.c:472-526), Jdot = H qd, then Jdot qd."""
        n = I_p.shape[0]
        JTd = np.zeros((n, 3, self.nJ))
        JRd = np.zeros((n, 3, self.nJ))
        chain = [j for j in self.chain(body) if self.jac[j] >= 0]
        for a, j in enumerate(chain):
            aj = jrot[:, j, self.dir[j], :]
            cj = self.jac[j]
            for k in chain[a:]:
                ak = jrot[:, k, self.dir[k], :]
                ck = self.jac[k]
                # point Hessian: symmetric in (j, k)
                if self.is_rot[j] and self.is_rot[k]:
                    sj = _cross(ak, _cross(aj, I_p - jorg[:, j]))
                elif (not self.is_rot[j]) and self.is_rot[k]:
                    sj = _cross(ak, aj)
                else:
                    sj = None
                if sj is not None:
                    JTd[:, :, cj] += sj * qd_ik[:, ck, None]
                    if ck != cj:
                        JTd[:, :, ck] += sj * qd_ik[:, cj, None]
                # rotation Hessian: only dJ_j/dq_k for k before j in the chain
                if self.is_rot[j] and self.is_rot[k] and ck != cj:
                    JRd[:, :, cj] += _cross(ak, aj) * qd_ik[:, ck, None]
        return (np.einsum("nij,nj->ni", JTd, qd_ik), np.einsum("nij,nj->ni", JRd, qd_ik))

    def kinetic_terms(self, q, qd=None):
        fk = self.fk(q, qd if qd is not None else np.zeros_like(np.asarray(q)))
        org, rot, jorg, jrot = self._last
        n = org.shape[0]
        nJ = self.nJ
        qd_full = fk["qd"]
        qd_ik = np.zeros((n, nJ))                                         # Rcs_dynamics.c:449-451
        for j in range(self.dof):
            if self.jac[j] >= 0:
                qd_ik[:, self.jac[j]] = qd_full[:, j]
        M = np.zeros((n, nJ, nJ))
        h = np.zeros((n, nJ))
        Fg = np.zeros((n, nJ))
        V = np.zeros(n)

        for b, body in enumerate(self.bodies):
            m = body["m"]
            if m == 0.0:
                continue
            com, inertia = _htr(body["Inertia"])
            R = rot[:, b]
            I_p = org[:, b] + np.einsum("nki,k->ni", R, com)
            JT = self._point_jacobian(b, I_p, jorg, jrot)
            JR = self._rotation_jacobian(b, jrot, n)
            # I_I = A_BI^T I_body A_BI  (Mat3d_similarityTransform, Rcs_Mat3d.c:364)
            I_I = np.einsum("nki,kl,nlj->nij", R, inertia, R)
            M += m * np.einsum("nki,nkj->nij", JT, JT) + np.einsum("nki,nkl,nlj->nij", JR, I_I, JR)

            aT, aR = self._dot_jacobians_times_qd(b, I_p, jorg, jrot, qd_ik)
            om = fk["omega"][:, b]                                        # BODY->omega
            h_rot = _cross(om, np.einsum("nij,nj->ni", I_I, om))         # Rcs_dynamics.c:346-377
            wrench_t = m * aT
            wrench_r = np.einsum("nij,nj->ni", I_I, aR) + h_rot
            h -= np.einsum("nki,nk->ni", JT, wrench_t) + np.einsum("nki,nk->ni", JR, wrench_r)

            Fg += JT[:, 2, :] * (-GRAVITY * m)                            # Rcs_dynamics.c:540-546
            V += m * GRAVITY * I_p[:, 2]

        T = 0.5 * np.einsum("ni,nij,nj->n", qd_ik, M, qd_ik)
        return {"M": M, "h": h, "Fg": Fg, "E": V + T}

    # ------------------------------------------------------------------------------------
    # task Jacobians relative to a reference body / frame:
    # RcsGraph_3dPosJacobian / RcsGraph_3dOmegaJacobian (Rcs_kinematics.c:566-615, :789-822)
    # ------------------------------------------------------------------------------------
    def is_articulated(self, body: int) -> bool:
        """RcsBody_isArticulated (src/RcsCore/Rcs_body.c:688)."""
        return any(self.jac[j] >= 0 for j in self.chain(body))

    def task_jacobians(self, q, spec):
        """spec: [(kind "POS" | "OMEGA", effector, refBdy, refFrame)], body indices, -1 = NULL."""
        self.fk(q)
        org, rot, jorg, jrot = self._last
        n = org.shape[0]
        zero = np.zeros((n, 3, self.nJ))

        def JT(b):
            return zero.copy() if b < 0 else self._point_jacobian(b, org[:, b], jorg, jrot)

        def JR(b):
            return zero.copy() if b < 0 else self._rotation_jacobian(b, jrot, n)

        out = np.zeros((n, len(spec), 3, self.nJ))
        for t, (kind, eff, rb, rf) in enumerate(spec):
            if kind == "POS":
                if rb < 0:
                    J = JT(eff)
                    if rf >= 0:
                        J = np.einsum("nij,njk->nik", rot[:, rf], J)
                else:
                    r2 = org[:, eff] if eff >= 0 else np.zeros((n, 3))
                    r12 = r2 - org[:, rb]
                    J = JT(eff) - JT(rb) + np.cross(r12[:, :, None], JR(rf), axis=1)
                    A = rot[:, rf] if (rf >= 0 and rf != rb) else rot[:, rb]
                    J = np.einsum("nij,njk->nik", A, J)
            else:
                if rb < 0 and rf < 0:
                    J = JR(eff)
                elif rb >= 0 and rf == rb and not self.is_articulated(rb):
                    J = np.einsum("nij,njk->nik", rot[:, rb], JR(eff))
                else:
                    J = JR(eff) - JR(rb)
                    if rf >= 0:
                        J = np.einsum("nij,njk->nik", rot[:, rf], J)
            out[:, t] = J
        return out

    # ------------------------------------------------------------------------------------
    # centre of gravity: RcsGraph_COG / RcsGraph_COGJacobian (Rcs_kinematics.c:904, :973)
    # ------------------------------------------------------------------------------------
    def cog(self, q):
        self.fk(q)
        org, rot, jorg, jrot = self._last
        n = org.shape[0]
        r = np.zeros((n, 3))
        J = np.zeros((n, 3, self.nJ))
        mass = 0.0
        for b, body in enumerate(self.bodies):
            m = body["m"]
            if not m > 0.0:
                continue
            com, _ = _htr(body["Inertia"])
            I_p = org[:, b] + np.einsum("nki,k->ni", rot[:, b], com)
            r += m * I_p
            J += m * self._point_jacobian(b, I_p, jorg, jrot)
            mass += m
        if mass > 0.0:
            r *= 1.0 / mass
            J *= 1.0 / mass
        return r, J

    # ------------------------------------------------------------------------------------
    # joint-space metric and joint-limit cost
    # ------------------------------------------------------------------------------------
    def ik_columns(self):
        """jointIndex of every Jacobian column (inverse of jacobiIndex)."""
        cols = [None] * self.nJ
        for j in range(self.dof):
            if self.jac[j] >= 0:
                cols[self.jac[j]] = j
        return cols

    def inv_wq(self):
        """RcsGraph_getInvWq, RcsStateIK (src/RcsCore/Rcs_graph.c:1087)."""
        return np.array([self.joints[j]["weightMetric"] * (self.joints[j]["q_max"] - self.joints[j]["q_min"])
                         for j in self.ik_columns()])

    def joint_limit_gradient(self, q):
        """RcsGraph_jointLimitGradient, RcsStateIK (src/RcsCore/Rcs_kinematics.c:1496):
        returns (dH [n x nJ], cost [n])."""
        qf, _ = self.expand(q)
        cols = self.ik_columns()
        rng = np.array([self.joints[j]["q_max"] - self.joints[j]["q_min"] for j in cols])
        wjl = np.array([self.joints[j]["weightJL"] for j in cols])
        q0 = np.array([self.joints[j]["q0"] for j in cols])
        delta = qf[:, cols] - q0
        ok = rng > 0.0
        safe = np.where(ok, rng, 1.0)
        dH = np.where(ok, wjl * delta / (safe * safe), 0.0)
        cost = 0.5 * np.sum(np.where(ok, wjl * (delta / safe) ** 2, 0.0), axis=1)
        return dH, cost

    # ------------------------------------------------------------------------------------
    # direct dynamics: Rcs_directDynamics (src/RcsCore/Rcs_dynamics.c:673-804) with
    # MatNd_choleskySolve (src/RcsCore/Rcs_linAlg.c:133-187, factor :66-101)
    # ------------------------------------------------------------------------------------
    def direct_dynamics(self, q, qd, F_ext=None, F_jnt=None):
        kt = self.kinetic_terms(q, qd)
        b = kt["Fg"].copy()
        if F_ext is not None:
            b = b + F_ext
        if F_jnt is not None:
            b = b - F_jnt
        b = b + kt["h"]
        L = kt["M"].copy()
        ns, n, _ = L.shape
        fail = np.zeros(ns, dtype=bool)
        for i in range(n):
            for k in range(i + 1):
                s = L[:, i, k] - np.einsum("nt,nt->n", L[:, i, :k], L[:, k, :k])
                if i > k:
                    with np.errstate(divide="ignore", invalid="ignore"):
                        L[:, i, k] = s / L[:, k, k]
                else:
                    fail |= ~(s > 0.0)
                    L[:, i, i] = np.sqrt(np.where(s > 0.0, s, 1.0))
        x = np.zeros_like(b)
        for i in range(n):
            x[:, i] = b[:, i]
            for j in range(i):
                x[:, i] -= L[:, i, j] * x[:, j]
            x[:, i] /= L[:, i, i]
        for i in range(n - 1, -1, -1):
            for j in range(i + 1, n):
                x[:, i] -= L[:, j, i] * x[:, j]
            x[:, i] /= L[:, i, i]
        x[fail] = 0.0
        return x, kt["E"]

    # ------------------------------------------------------------------------------------
    # orientation helpers
    # ------------------------------------------------------------------------------------
    @staticmethod
    def euler_angles(rot9):
        """Mat3d_toEulerAngles (src/RcsCore/Rcs_Mat3d.c:956-972); rot9: (n, 9) row-major."""
        A = np.asarray(rot9).reshape(-1, 3, 3)
        cy = np.sqrt(A[:, 0, 0] ** 2 + A[:, 1, 0] ** 2)
        reg = cy > 16.0 * FLT_EPSILON
        ea = np.zeros((A.shape[0], 3))
        ea[:, 0] = np.where(reg, -np.arctan2(A[:, 2, 1], A[:, 2, 2]), -np.arctan2(-A[:, 1, 2], A[:, 1, 1]))
        ea[:, 1] = -np.arctan2(-A[:, 2, 0], cy)
        ea[:, 2] = np.where(reg, -np.arctan2(A[:, 1, 0], A[:, 0, 0]), 0.0)
        return ea

    @staticmethod
    def from_euler(ea):
        """Mat3d_fromEulerAngles (src/RcsCore/Rcs_Mat3d.c:1152-1175)."""
        ea = np.asarray(ea)
        sa, ca = np.sin(ea[:, 0]), np.cos(ea[:, 0])
        sb, cb = np.sin(ea[:, 1]), np.cos(ea[:, 1])
        sg, cg = np.sin(ea[:, 2]), np.cos(ea[:, 2])
        A = np.empty((ea.shape[0], 3, 3))
        A[:, 0, 0] = cb * cg
        A[:, 0, 1] = ca * sg + sa * sb * cg
        A[:, 0, 2] = sa * sg - ca * sb * cg
        A[:, 1, 0] = -cb * sg
        A[:, 1, 1] = ca * cg - sa * sb * sg
        A[:, 1, 2] = sa * cg + ca * sb * sg
        A[:, 2, 0] = sb
        A[:, 2, 1] = -sa * cb
        A[:, 2, 2] = ca * cb
        return A

    def euler_error(self, x_des, x_curr):
        """TaskEuler3D::computeDX (src/RcsCore/TaskEuler3D.cpp:243-262)."""
        Ac, Ad = self.from_euler(x_curr), self.from_euler(x_des)
        cs = 0.5 * (np.einsum("nij,nij->n", Ad, Ac) - 1.0)                # Mat3d_diffAngle :1613
        angle = np.arccos(np.clip(cs, -1.0, 1.0))                         # Math_acos
        # < 1 deg: 0.5 (n x nd + s x sd + a x ad)   (Mat3d_getEulerError :1044)
        small = 0.5 * sum(_cross(Ac[:, r], Ad[:, r]) for r in range(3))
        # else: axis of A_des A_curr^T, rotated back by A_curr^T, times the angle
        A21 = np.einsum("nik,njk->nij", Ad, Ac)                           # Mat3d_getAxisAngle :1227
        ax = np.stack([A21[:, 1, 2] - A21[:, 2, 1], A21[:, 2, 0] - A21[:, 0, 2],
                       A21[:, 0, 1] - A21[:, 1, 0]], axis=1)
        ln = np.linalg.norm(ax, axis=1, keepdims=True)
        ax = np.where(ln > 0.0, ax / np.where(ln > 0.0, ln, 1.0), ax)
        large = np.einsum("nki,nk->ni", Ac, ax) * angle[:, None]
        # (the angle == pi branch of external/Rcs_thirdPartyMath.c:107-150 is not restated:
        #  it has measure zero for the sampled test inputs)
        return np.where((angle < np.pi / 180.0)[:, None], small, large)

    # ------------------------------------------------------------------------------------
    # IK: ControllerBase + IkSolverRMR::solveRightInverse
    # ------------------------------------------------------------------------------------
    def task_kinematics(self, q, tasks):
        """x and stacked J (ControllerBase::computeX / computeJ, ControllerBase.cpp:776, :867)
        for tasks = [("XYZ" | "ABC", body index), ...] with world reference."""
        self.fk(q)
        org, rot, jorg, jrot = self._last
        n = org.shape[0]
        xs, Js = [], []
        for kind, b in tasks:
            if kind == "XYZ":                                             # TaskPosition3D.cpp:127
                xs.append(org[:, b])
                Js.append(self._point_jacobian(b, org[:, b], jorg, jrot))
            else:                                                         # TaskEuler3D.cpp:110
                xs.append(self.euler_angles(rot[:, b].reshape(n, 9)))
                Js.append(self._rotation_jacobian(b, jrot, n))
        return np.concatenate(xs, axis=1), np.concatenate(Js, axis=1)

    def ik_task_error(self, q, tasks, x_des):
        """ControllerBase::computeDX (ControllerBase.cpp:896) with all activations 1."""
        x, _ = self.task_kinematics(q, tasks)
        dx = np.zeros_like(x)
        for t, (kind, _) in enumerate(tasks):
            sl = slice(3 * t, 3 * t + 3)
            if kind == "XYZ":
                dx[:, sl] = x_des[:, sl] - x[:, sl]                       # TaskGenericIK.cpp:80
            else:
                dx[:, sl] = self.euler_error(x_des[:, sl], x[:, sl])
        return dx

    def _cholesky_inverse(self, A):
        """MatNd_choleskyInverse (src/RcsCore/Rcs_linAlg.c:66-101, :210-227, :270-307):
        Banachiewicz L, in-place inverse of L, L^-T L^-1; det = prod L_ii^2, failure -> 0."""
        n_states, n, _ = A.shape
        L = A.copy()
        det = np.ones(n_states)
        fail = np.zeros(n_states, dtype=bool)
        for i in range(n):
            for k in range(i + 1):
                s = L[:, i, k] - np.einsum("nt,nt->n", L[:, i, :k], L[:, k, :k])
                if i > k:
                    with np.errstate(divide="ignore", invalid="ignore"):
                        L[:, i, k] = s / L[:, k, k]
                else:
                    fail |= ~(s > 0.0)
                    sq = np.sqrt(np.where(s > 0.0, s, 1.0))
                    L[:, i, i] = sq
                    det *= sq * sq
        for k in range(n):
            L[:, k, k] = 1.0 / L[:, k, k]
            for i in range(k + 1, n):
                L[:, i, k] = -np.einsum("nt,nt->n", L[:, i, k:i], L[:, k:i, k]) / L[:, i, i]
        inv = np.zeros_like(A)
        for r in range(n):
            for c in range(r, n):
                s = np.einsum("nt,nt->n", L[:, c:, r], L[:, c:, c])
                inv[:, r, c] = s
                inv[:, c, r] = s
        det = np.where(fail, 0.0, det)
        inv[fail] = 0.0
        return inv, det

    def coupled_joint_matrix(self, qf):
        """RcsGraph_coupledJointMatrix (src/RcsCore/Rcs_graph.c:2823): A [n x nJ x nqr] and
        A+ = (A^T A)^-1 A^T [n x nqr x nJ] at the full states qf."""
        n = qf.shape[0]
        nJ = self.nJ
        H = np.zeros((n, nJ, nJ))
        col_sum = np.zeros((n, nJ))
        for j, jn in enumerate(self.joints):
            c = self.jac[j]
            if c < 0:
                continue
            mj = jn["coupledTo"]
            if mj < 0:
                H[:, c, c] = 1.0
                col_sum[:, c] += 1.0
            else:
                ms = self.joints[mj]
                coef = jn["couplingFactors"]
                if len(coef) == 1:
                    sens = np.full(n, coef[0])
                else:                                    # RcsJoint_computeSlaveJointVelocity, qd = 1
                    sens = self._poly_d(coef, np.clip(qf[:, mj], ms["q_min"], ms["q_max"])) * 1.0
                H[:, c, self.jac[mj]] = sens
                col_sum[:, self.jac[mj]] += sens
        keep = [c for c in range(nJ) if np.all(np.abs(col_sum[:, c]) > 0.0)]
        A = H[:, :, keep] / np.abs(col_sum[:, None, keep])
        inv_diag = 1.0 / np.einsum("nij,nij->nj", A, A)
        invA = inv_diag[:, :, None] * np.transpose(A, (0, 2, 1))
        return A, invA

    def ik_rmr(self, q, tasks, x_des, lam=1.0e-8, alpha=0.05, iters=1, left=False):
        """examples/ExampleIK.cpp:154-177 with IkSolverRMR::solveRightInverse
        (src/RcsCore/IkSolverRMR.cpp:415-596, kinematically coupled joints in the reduced space
        of RcsGraph_coupledJointMatrix) and MatNd_rwPinv (src/RcsCore/Rcs_MatNd.c:1833); with
        `left` IkSolverRMR::solveLeftInverse (:205-272; computeKinematics :176-200 scales the
        Jacobian by the joint metric, binary task blending = unit task weights) and MatNd_rwPinv2
        (Rcs_MatNd.c:1920). Returns (q, dx_last, dq_last, det_last)."""
        q = self.expand(np.asarray(q, dtype=np.float64))[0]
        x_des = np.asarray(x_des, dtype=np.float64)
        nJ = self.nJ
        cols = self.ik_columns()
        inv_wq = self.inv_wq()
        coupled = any(jn["coupledTo"] >= 0 and self.jac[j] >= 0 for j, jn in enumerate(self.joints))
        dx = dq_full = det = None
        for _ in range(iters):
            dx = self.ik_task_error(q, tasks, x_des)
            _, J = self.task_kinematics(q, tasks)
            # RcsGraph_jointLimitGradient (Rcs_kinematics.c:1496), then dH *= alpha
            dHr = -(self.joint_limit_gradient(q)[0] * alpha)
            w = np.tile(inv_wq, (q.shape[0], 1))
            if coupled:
                A, invA = self.coupled_joint_matrix(q)
                J = np.einsum("nij,njk->nik", J, A)
                w = np.einsum("nij,j->ni", invA, inv_wq)
                dHr = np.einsum("nij,nj->ni", invA, dHr)
            nq = J.shape[2]
            if left:
                Jr = J * w[:, None, :]                                    # J A invWq_r
                B = np.einsum("nki,nkj->nij", Jr, Jr) + lam * np.eye(nq)
                inv, det = self._cholesky_inverse(B)
                pinv = np.einsum("nij,nkj->nik", inv, Jr)                 # (J^T J + lam)^-1 J^T
                N = np.eye(nq) - np.einsum("nij,njk->nik", pinv, Jr)
                dqr = np.einsum("nij,nj->ni", pinv, dx) + np.einsum("nij,nj->ni", N * w[:, None, :], dHr)
                dqr = dqr * w                                             # back to the unweighted space
                dq = np.einsum("nij,nj->ni", A, dqr) if coupled else dqr
                dq[det == 0.0] = 0.0
                dq_full = np.zeros_like(q)
                dq_full[:, cols] = dq
                q = self.expand(q + dq_full)[0]
                continue
            WJt = w[:, :, None] * np.transpose(J, (0, 2, 1))
            M = np.einsum("nij,njk->nik", J, WJt) + lam * np.eye(J.shape[1])
            inv, det = self._cholesky_inverse(M)
            pinv = np.einsum("nij,njk->nik", WJt, inv)
            N = np.eye(nq) - np.einsum("nij,njk->nik", pinv, J)           # MatNd_nullspace :3758
            NinvW = N * w[:, None, :]
            ts = np.einsum("nij,nj->ni", pinv, dx)
            ns = np.einsum("nij,nj->ni", NinvW, dHr)
            if coupled:
                ts = np.einsum("nij,nj->ni", A, ts)
                ns = np.einsum("nij,nj->ni", A, ns)
            dq = ts + ns
            dq[det == 0.0] = 0.0
            dq_full = np.zeros_like(q)
            dq_full[:, cols] = dq
            q = self.expand(q + dq_full)[0]                               # RcsGraph_setState
        return q, dx, dq_full, det
