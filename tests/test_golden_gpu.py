"""GPU parity against the committed golden vectors (tests/golden/*.npz, recorded from the
unmodified reference) and the numpy oracle, plus size-independent properties at
BASELINE.json's full sizes (1M states on the 7-DoF arm, 256K humanoid states).

Tolerance 1e-12 relative to the array's inf-norm (floor 1), as BASELINE.json states; integer
layout (body order, DoF indexing) is compared bit for bit in tests/test_host_cpu.py.
"""
import json
import os

import numpy as np
import pytest

import rcs_b200
from rcs_b200 import TASK_ABC, TASK_XYZ, workload
from conftest import GRAPH_DIR, ROOT, assert_close, assert_structural_zeros, graph_file
from oracle import rcs_oracle

pytestmark = pytest.mark.gpu
GOLDEN = ["wam_arm", "wam_scenario", "humanoid", "husky", "gantry", "polycoupled", "wrist6"]


def _golden(name):
    return np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))


def _layout(name):
    with open(os.path.join(GRAPH_DIR, name + ".layout.json")) as f:
        return json.load(f)


@pytest.mark.parametrize("name", GOLDEN)
def test_fk_jacobians_kinetic_terms_match_golden(name, cuda_device):
    z = _golden(name)
    g = rcs_b200.Graph(graph_file(name))
    m = rcs_b200.Model(g, cuda_device)
    q, qd = np.ascontiguousarray(z["q"]), np.ascontiguousarray(z["qd"])
    fk = m.fk(q, qd, want=("A_BI", "A_JI", "xdot", "omega", "q"))
    for k in ("A_BI", "A_JI", "xdot", "omega", "q"):
        assert_close(fk[k], z["fk_" + k], what=f"{name}:{k}")
    jac = m.fk_jacobians(q, [int(e) for e in z["eff"]], points=z["pts"], want=("JT", "JR"))
    for k in ("JT", "JR"):
        assert_close(jac[k], z["jac_" + k], what=f"{name}:{k}")
        assert_structural_zeros(jac[k], z["jac_" + k], what=f"{name}:{k}")
    kt = m.kinetic_terms(q, qd)
    for k in ("M", "h", "Fg", "E"):
        assert_close(kt[k], z["kt_" + k], what=f"{name}:{k}")
    cg = m.cog(q)
    assert_close(cg["cog"], z["cog"], what=f"{name}:cog")
    assert_close(cg["Jcog"], z["Jcog"], what=f"{name}:Jcog")
    spec = [("POS" if k == 1 else "OMEGA", int(e), int(rb), int(rf)) for k, e, rb, rf in z["rel_spec"]]
    assert_close(m.task_jacobians(q, spec), z["rel_J"], what=f"{name}: relative task Jacobians")
    dd = m.direct_dynamics(q, qd, np.ascontiguousarray(z["dd_Fext"]), np.ascontiguousarray(z["dd_Fjnt"]))
    cond = np.linalg.cond(z["kt_M"])
    err = np.abs(dd["qdd"] - z["dd_qdd"]).max(axis=1)
    tol = 1e-12 * np.maximum(1.0, cond / 100.0) * np.maximum(1.0, np.abs(z["dd_qdd"]).max(axis=1))
    assert np.all(err <= tol), f"{name}: direct dynamics, worst err/tol {float((err / tol).max()):.2f}"


@pytest.mark.parametrize("name", ["wam_arm", "gantry", "wrist6"])
def test_fused_mass_matrix_and_left_inverse_match_golden(name, cuda_device):
    """The fused FK + Jacobians + M(q) launch (serial-chain variant on both graphs, the gantry with
    translational joints) and RCSB_IK_LEFT_INVERSE steps against vectors recorded from the
    reference (RcsGraph_computeMassMatrix; IkSolverRMR::solveLeftInverse, lambda 1e-3)."""
    z = _golden(name)
    g = rcs_b200.Graph(graph_file(name))
    m = rcs_b200.Model(g, cuda_device)
    q = np.ascontiguousarray(z["q"])
    out = m.fk_jacobians(q, [int(e) for e in z["eff"]], points=z["pts"], want=("A_BI", "JT", "JR", "M"))
    assert_close(out["M"], z["mass_M"], what=f"{name}:M (fused)")
    assert_structural_zeros(out["M"], z["mass_M"], what=f"{name}:M (fused)")
    assert_close(out["JT"], z["jac_JT"], what=f"{name}:JT (fused)")
    assert_close(out["JR"], z["jac_JR"], what=f"{name}:JR (fused)")
    assert_close(out["A_BI"], z["fk_A_BI"], what=f"{name}:A_BI (fused)")
    tip = int(z["ik_tip"])
    qs = np.ascontiguousarray(z["ik_q_start"]).copy()
    ik = m.ik_rmr([(TASK_XYZ, tip), (TASK_ABC, tip)], qs, np.ascontiguousarray(z["ik_x_des"]), lam=1.0e-3,
                  iters=1, want=("dx", "dq", "det"), left_inverse=True)
    assert_close(ik["dx"], z["ikl_dx1"], what="dx")
    # lambda = 1e-3 bounds the condition number of J_r^T J_r + lambda 1 by (s_max^2 + lambda) / lambda
    scale = np.maximum(1.0, np.abs(z["ikl_dq1"]).max(axis=1))
    assert np.all(np.abs(ik["dq"] - z["ikl_dq1"]).max(axis=1) <= 1e-9 * scale)
    assert np.all(np.abs(qs - z["ikl_q1"]).max(axis=1) <= 1e-9 * scale)
    assert np.all(np.abs(ik["det"] / z["ikl_det1"] - 1.0) <= 1e-9)
    q5 = np.ascontiguousarray(z["ik_q_start"]).copy()
    m.ik_rmr([(TASK_XYZ, tip), (TASK_ABC, tip)], q5, np.ascontiguousarray(z["ik_x_des"]), lam=1.0e-3, iters=5,
             left_inverse=True)
    assert np.abs(q5 - z["ikl_q5"]).max() <= 1e-8


@pytest.mark.parametrize("name", ["wam_arm", "humanoid", "gantry", "wrist6"])
def test_ik_matches_golden(name, cuda_device):
    z = _golden(name)
    g = rcs_b200.Graph(graph_file(name))
    m = rcs_b200.Model(g, cuda_device)
    o = rcs_oracle.Oracle(_layout(name))
    tip = int(z["ik_tip"])
    q = np.ascontiguousarray(z["ik_q_start"]).copy()
    out = m.ik_rmr([(TASK_XYZ, tip), (TASK_ABC, tip)], q, np.ascontiguousarray(z["ik_x_des"]), iters=1,
                   want=("dx", "dq", "det"))
    assert_close(out["dx"], z["ik_dx1"], what="dx")
    _, J = o.task_kinematics(z["ik_q_start"], [("XYZ", tip), ("ABC", tip)])
    cond = np.linalg.cond(np.einsum("nij,j,nkj->nik", J, o.inv_wq(), J) + 1e-8 * np.eye(6))
    err = np.abs(out["dq"] - z["ik_dq1"]).max(axis=1)
    tol = 1e-12 * np.maximum(1.0, cond / 100.0) * np.maximum(1.0, np.abs(z["ik_dq1"]).max(axis=1))
    assert np.all(err <= tol), float((err / tol).max())
    assert np.all(np.abs(out["det"] / z["ik_det1"] - 1.0) <= 1e-11 * np.maximum(1.0, cond / 100.0))


@pytest.mark.parametrize("name,n", [("wam_arm", 4096), ("wam_scenario", 512), ("humanoid", 256),
                                    ("dexbot", 128), ("schunk_sdh", 512)])
def test_against_numpy_oracle_on_seeded_inputs(name, n, cuda_device):
    lay = _layout(name)
    o = rcs_oracle.Oracle(lay)
    g = rcs_b200.Graph(graph_file(name))
    m = rcs_b200.Model(g, cuda_device)
    q, qd = workload.sample_states(lay, n, first=31337, with_velocity=True)
    effs = [o.nb - 1, o.nb // 3]
    ours = m.fk_jacobians(q, effs)
    ref = o.fk_jacobians(q, effs)
    for k in ("A_BI", "JT", "JR"):
        assert_close(ours[k], ref[k], what=f"{name}:{k}")
    kt, rkt = m.kinetic_terms(q, qd), o.kinetic_terms(q, qd)
    for k in ("M", "h", "Fg", "E"):
        assert_close(kt[k], rkt[k], what=f"{name}:{k}")


def test_full_size_arm_batch_properties(cuda_device):
    """configs[1] at full size: 1,048,576 states. Properties that need no oracle pass over all
    states (orthonormal frames, x_dot = JT q_dot, omega = JR q_dot, bin/TestGraph.cpp:545-566);
    a strided sample is compared with the oracle."""
    import torch

    n = 1 << 20
    lay = _layout("wam_arm")
    g = rcs_b200.Graph(graph_file("wam_arm"))
    m = rcs_b200.Model(g, cuda_device)
    tip = g.body_index("BallTip")
    q, qd = workload.sample_states(lay, n, with_velocity=True)
    dq, dqd = torch.from_numpy(q).cuda(cuda_device), torch.from_numpy(qd).cuda(cuda_device)
    out = m.fk_jacobians(dq, [tip])
    vel = m.fk(dq, dqd, bodies=[tip], want=("xdot", "omega"))
    torch.cuda.synchronize()
    R = out["A_BI"][:, :, 3:].reshape(n, g.n_bodies, 3, 3)
    eye = torch.eye(3, dtype=torch.float64, device=R.device)
    assert float((R @ R.transpose(-1, -2) - eye).abs().max()) < 1e-13
    xd = torch.einsum("nij,nj->ni", out["JT"][:, 0], dqd)
    om = torch.einsum("nij,nj->ni", out["JR"][:, 0], dqd)
    assert float((xd - vel["xdot"][:, 0]).abs().max()) < 1e-12 * max(1.0, float(vel["xdot"].abs().max()))
    assert float((om - vel["omega"][:, 0]).abs().max()) < 1e-12 * max(1.0, float(vel["omega"].abs().max()))
    idx = np.arange(0, n, 4099)
    ref = rcs_oracle.Oracle(lay).fk_jacobians(q[idx], [tip])
    for k in ("A_BI", "JT", "JR"):
        assert_close(out[k][idx].cpu().numpy(), ref[k], what=k)
    # the headline launch (frames + Jacobians + M from fkChainKernel) on all states: the same frames and
    # rotation Jacobians bit for bit, M symmetric bit for bit and positive definite everywhere, and equal to the
    # matrix of the kinetic-terms kernels (another algorithm: walk + assembly) on a strided sample
    fused = m.fk_jacobians(dq, [tip], want=("A_BI", "JT", "JR", "M"))
    torch.cuda.synchronize()
    assert torch.equal(fused["A_BI"], out["A_BI"]) and torch.equal(fused["JR"], out["JR"])
    # with M the columns are kept as spatial axes (w, v = o x w), JT = w x p + v instead of w x (p - o)
    assert float((fused["JT"] - out["JT"]).abs().max()) < 1e-14
    M = fused["M"]
    assert torch.equal(M, M.transpose(1, 2))
    _, info = torch.linalg.cholesky_ex(M)
    assert int(info.abs().max()) == 0
    didx = torch.from_numpy(idx).cuda(cuda_device)
    Mk = m.kinetic_terms(dq[didx].contiguous(), want=("M",))["M"]
    assert float((M[didx] - Mk).abs().max()) < 1e-12 * max(1.0, float(Mk.abs().max()))
    assert_close(M[didx].cpu().numpy(), rcs_oracle.Oracle(lay).kinetic_terms(q[idx], np.zeros_like(q[idx]))["M"], what="M")


def test_full_size_humanoid_kinetic_terms_properties(cuda_device):
    """configs[2] at full size: 262,144 humanoid states. M symmetric positive definite on every
    state (the humanoid's inertia tensors are symmetric), E = V + 1/2 qd^T M qd, and a strided
    sample against the oracle."""
    import torch

    n = 1 << 18
    lay = _layout("humanoid")
    o = rcs_oracle.Oracle(lay)
    g = rcs_b200.Graph(graph_file("humanoid"))
    m = rcs_b200.Model(g, cuda_device)
    q, qd = workload.sample_states(lay, n, with_velocity=True)
    dq, dqd = torch.from_numpy(q).cuda(cuda_device), torch.from_numpy(qd).cuda(cuda_device)
    kt = m.kinetic_terms(dq, dqd)
    static = m.kinetic_terms(dq, want=("E",))
    torch.cuda.synchronize()
    M = kt["M"]
    assert float((M - M.transpose(1, 2)).abs().max()) <= 1e-12 * float(M.abs().max())
    qd_ik = dqd[:, o.ik_columns()]
    T = 0.5 * torch.einsum("ni,nij,nj->n", qd_ik, M, qd_ik)
    assert float((kt["E"] - static["E"] - T).abs().max()) <= 1e-11 * float(kt["E"].abs().max())
    sub = M[:: 257]
    assert bool((torch.linalg.eigvalsh(0.5 * (sub + sub.transpose(1, 2))) > 0).all())
    idx = np.arange(0, n, 8209)
    ref = o.kinetic_terms(q[idx], qd[idx])
    for k in ("M", "h", "Fg", "E"):
        assert_close(kt[k][idx].cpu().numpy(), ref[k], what=k)


def test_ik_converges_on_reachable_targets_large_batch(cuda_device):
    """configs[3] shape (reduced to 262,144 targets): 10 RMR iterations towards reachable poses
    shrink the task error on every well-conditioned state; q stays finite; singular steps are
    flagged by det = 0 and leave q unchanged (in-band failure of the reference)."""
    import torch

    n = 1 << 18
    lay = _layout("wam_arm")
    o = rcs_oracle.Oracle(lay)
    g = rcs_b200.Graph(graph_file("wam_arm"))
    m = rcs_b200.Model(g, cuda_device)
    tip = g.body_index("BallTip")
    tasks = [(TASK_XYZ, tip), (TASK_ABC, tip)]
    q_goal = workload.sample_states(lay, n, first=1 << 22, margin=0.2)
    fk = m.fk(torch.from_numpy(q_goal).cuda(cuda_device), bodies=[tip])["A_BI"][:, 0].cpu().numpy()
    x_des = np.ascontiguousarray(np.concatenate([fk[:, :3], o.euler_angles(fk[:, 3:])], axis=1))
    rng = np.random.default_rng(4)
    q0 = np.ascontiguousarray(q_goal + 0.05 * rng.uniform(-1, 1, size=q_goal.shape))
    dq = torch.from_numpy(q0.copy()).cuda(cuda_device)
    dx_t = torch.from_numpy(x_des).cuda(cuda_device)
    first = m.ik_rmr(tasks, dq.clone(), dx_t, iters=1, want=("dx",))["dx"]
    out = m.ik_rmr(tasks, dq, dx_t, iters=10, want=("dx", "det"))
    torch.cuda.synchronize()
    assert bool(torch.isfinite(dq).all())
    e0 = first.abs().max(dim=1).values
    e9 = out["dx"].abs().max(dim=1).values
    ok = out["det"] > 1e-6
    assert float(ok.double().mean()) > 0.9
    assert float((e9[ok] < 0.5 * e0[ok] + 1e-9).double().mean()) > 0.99
    # strided sample against the oracle, one iteration
    idx = np.arange(0, n, 16411)
    _, dx_ref, _, _ = o.ik_rmr(q0[idx], [("XYZ", tip), ("ABC", tip)], x_des[idx], iters=1)
    assert_close(first[idx].cpu().numpy(), dx_ref, what="dx")
