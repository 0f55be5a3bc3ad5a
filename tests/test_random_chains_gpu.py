"""GPU parity on randomly generated serial arms: the kernels that fold a chain at plan time
(fkChainKernel, ikChainKernel<..., MODE 1 / 2>: one transform per stretch between two columns, row
shifts for the joint axes, per-level composite mass constants, rcsb_plan.h) against the kernels that
interpret the flat model (RCSB_FK_GENERAL=1 / RCSB_IK_GENERAL=1) and against the reference library.

Every graph is written as an Rcs XML file: 1 to 8 rotational joints about random axes, one or two
joints per body, joint and body transforms that are absent, translation-only or full rotations, fixed
side bodies with mass hanging off random links, random symmetric positive definite inertia tensors."""
import os

import numpy as np
import pytest

import rcs_b200
from rcs_b200 import TASK_ABC, TASK_XYZ, workload
from conftest import assert_close, assert_structural_zeros

pytestmark = pytest.mark.gpu


def _rot(rng):
    """random rotation matrix (row-major 9 numbers), from a unit quaternion"""
    q = rng.normal(size=4)
    q /= np.linalg.norm(q)
    w, x, y, z = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def _transform(rng, kind):
    """kind 0: none, 1: translation only, 2: rotation only, 3: both -> XML attribute text or None"""
    if kind == 0:
        return None
    t = rng.uniform(-0.3, 0.3, size=3) if kind in (1, 3) else np.zeros(3)
    R = _rot(rng) if kind in (2, 3) else np.eye(3)
    return " ".join(repr(float(v)) for v in list(t) + list(R.reshape(-1)))


def _mass_attrs(rng):
    A = rng.normal(size=(3, 3))
    I = (A @ A.T + 0.2 * np.eye(3)) * rng.uniform(0.001, 0.05)
    I = 0.5 * (I + I.T)
    m = rng.uniform(0.1, 5.0)
    c = rng.uniform(-0.1, 0.1, size=3)
    return (f'mass="{m!r}" cogVector="{c[0]!r} {c[1]!r} {c[2]!r}" inertia="' +
            " ".join(repr(float(v)) for v in I.reshape(-1)) + '"')


def random_arm_xml(seed):
    rng = np.random.default_rng(seed)
    n_joints = int(rng.integers(1, 9))
    lines = [f'<Graph name="random_arm_{seed}" >', f'  <Body name="Base" physics="kinematic" {_mass_attrs(rng)} />']
    prev, made, link = "Base", 0, 0
    side = 0
    while made < n_joints:
        link += 1
        per_body = int(min(n_joints - made, rng.integers(1, 3)))
        name = f"Link{link}"
        tr = _transform(rng, int(rng.integers(0, 4)))
        attrs = f'prev="{prev}" ' + ('physics="dynamic" ' + _mass_attrs(rng) if rng.random() < 0.85 else 'physics="none" mass="0"')
        if tr:
            attrs += f' transform="{tr}"'
        lines.append(f'  <Body name="{name}" {attrs} >')
        for _ in range(per_body):
            made += 1
            lo = float(rng.uniform(-170, -30))
            hi = float(rng.uniform(30, 170))
            q0 = float(rng.uniform(lo * 0.5, hi * 0.5))
            jt = _transform(rng, int(rng.choice([0, 0, 1, 3])))
            ja = (f'name="J{made}" type="Rot{"XYZ"[int(rng.integers(0, 3))]}" range="{lo!r} {q0!r} {hi!r}" '
                  f'weightJL="{float(rng.uniform(0.5, 2.0))!r}" weightMetric="{float(rng.uniform(0.5, 1.5))!r}"')
            if jt:
                ja += f' transform="{jt}"'
            lines.append(f'    <Joint {ja} />')
        lines.append('  </Body>')
        # fixed side bodies hanging off this link (some with mass, one nested)
        for _ in range(int(rng.integers(0, 3))):
            side += 1
            st = _transform(rng, int(rng.integers(0, 4)))
            sa = f'prev="{name}" ' + ('physics="dynamic" ' + _mass_attrs(rng) if rng.random() < 0.6 else 'physics="none" mass="0"')
            if st:
                sa += f' transform="{st}"'
            lines.append(f'  <Body name="Side{side}" {sa} />')
            if rng.random() < 0.3:
                nt = _transform(rng, 3)
                lines.append(f'  <Body name="Side{side}b" prev="Side{side}" physics="dynamic" {_mass_attrs(rng)} '
                             f'transform="{nt}" />')
        prev = name
    tt = _transform(rng, int(rng.choice([1, 3])))
    lines.append(f'  <Body name="Tip" prev="{prev}" physics="none" mass="0" transform="{tt}" />')
    lines.append('</Graph>')
    return "\n".join(lines) + "\n"


@pytest.mark.parametrize("seed", list(range(10)))
def test_folded_chain_kernels_on_random_arms(seed, refmod, cuda_device, monkeypatch, tmp_path):
    path = os.path.join(tmp_path, f"arm{seed}.xml")
    with open(path, "w") as f:
        f.write(random_arm_xml(seed))
    g = rcs_b200.Graph(path)
    r = refmod.RefGraph(path)
    assert g.layout()["nJ"] == r.layout()["nJ"] == g.dof
    m = rcs_b200.Model(g, cuda_device)
    names = g.body_names()
    tip = names.index("Tip")
    eff = [tip if seed % 3 else names.index("Link1")]
    pts = np.array([[0.02, -0.03, 0.04]])
    q = workload.sample_states(g.layout(), 600 + seed, first=seed)

    # ---- frames, Jacobians, mass matrix: folded walk, interpreting walk, reference ----
    ref = r.fk_jacobians(q, eff, points=pts)
    M_ref = r.mass_matrix(q)
    out = {}
    for mode in ("chain", "general"):
        if mode == "general":
            monkeypatch.setenv("RCSB_FK_GENERAL", "1")
        out[mode] = m.fk_jacobians(q, eff, points=pts, want=("A_BI", "JT", "JR", "M"))
        for k in ("A_BI", "JT", "JR"):
            assert_close(out[mode][k], ref[k], what=f"seed {seed}/{mode}:{k}")
        assert_structural_zeros(out[mode]["JT"], ref["JT"], what=f"seed {seed}/{mode}:JT")
        assert_close(out[mode]["M"], M_ref, what=f"seed {seed}/{mode}:M")
    monkeypatch.delenv("RCSB_FK_GENERAL")
    differs = False
    for k in ("A_BI", "JT", "JR", "M"):
        d = np.abs(out["chain"][k] - out["general"][k]).max()
        assert d <= 2e-14 * max(1.0, np.abs(out["general"][k]).max()), (seed, k, d)
        differs = differs or d > 0.0
    assert differs, "two different kernels were expected to run (folded constants round differently)"

    # ---- one RMR step, right and left inverse: folded chain, general kernel, reference ----
    rik = refmod.RefIk(path, [("XYZ", "Tip"), ("ABC", "Tip")])
    tasks = [(TASK_XYZ, tip), (TASK_ABC, tip)]
    q_goal = workload.sample_states(g.layout(), 300, first=100 + seed, margin=0.2)
    x_des, J = rik.task_kinematics(q_goal)
    q0 = np.ascontiguousarray(q_goal + 0.05 * np.random.default_rng(seed).uniform(-1, 1, size=q_goal.shape))
    for left, lam in ((False, 1.0e-4), (True, 1.0e-3)):
        q_ref, dx_ref, dq_ref, det_ref = rik.rmr(q0, x_des, lam=lam, iters=1, left=left)
        w = rik.graph.inv_wq()
        _, J0 = rik.task_kinematics(q0)
        if left:
            Jr = J0 * w[None, None, :]
            cond = np.linalg.cond(np.einsum("nki,nkj->nij", Jr, Jr) + lam * np.eye(J0.shape[2]))
        else:
            cond = np.linalg.cond(np.einsum("nij,j,nkj->nik", J0, w, J0) + lam * np.eye(J0.shape[1]))
        res = {}
        for mode in ("chain", "general"):
            if mode == "general":
                monkeypatch.setenv("RCSB_IK_GENERAL", "1")
            qq = q0.copy()
            o = m.ik_rmr(tasks, qq, x_des, lam=lam, iters=1, left_inverse=left, want=("dx", "dq", "det"))
            res[mode] = (qq, o)
            assert_close(o["dx"], dx_ref, what=f"seed {seed}/{mode}: dx")
            ok = det_ref > 0.0
            err = np.abs(o["dq"][ok] - dq_ref[ok]).max(axis=1)
            tol = 1e-12 * np.maximum(1.0, cond[ok] / 100.0) * np.maximum(1.0, np.abs(dq_ref[ok]).max(axis=1))
            assert np.all(err <= tol), (seed, mode, left, float((err / tol).max()))
        monkeypatch.delenv("RCSB_IK_GENERAL")
