"""GPU parity: batched kinetic terms through the C ABI vs the reference library.

Reference: RcsGraph_computeKineticTerms (Rcs_dynamics.c:439) after RcsGraph_setState(q, qd);
invariants from bin/TestGraph.cpp:288-374 (mass matrix twice, gravity through the COG Jacobian).
Tolerance 1e-12 relative to the array's inf-norm (floor 1).
"""
import numpy as np
import pytest

import rcs_b200
from rcs_b200 import workload
from conftest import assert_close, assert_structural_zeros, graph_file

pytestmark = pytest.mark.gpu

CASES = [("wam_arm", 500), ("wam_scenario", 200), ("schunk_sdh", 200), ("dexbot", 64),
         ("husky", 64), ("humanoid", 48), ("polycoupled", 300)]


def _models(name, refmod, dev):
    g = rcs_b200.Graph(graph_file(name))
    return g, rcs_b200.Model(g, dev), refmod.RefGraph(graph_file(name))


@pytest.mark.parametrize("name,n", CASES)
def test_kinetic_terms_host_pointers(name, n, refmod, cuda_device):
    g, m, r = _models(name, refmod, cuda_device)
    q, qd = workload.sample_states(g.layout(), n, with_velocity=True)
    ours = m.kinetic_terms(q, qd)
    ref = r.kinetic_terms(q, qd)
    for k in ("M", "h", "Fg", "E"):
        assert_close(ours[k], ref[k], what=f"{name}:{k}")
    assert_structural_zeros(ours["M"], ref["M"], what=f"{name}:M")
    # reference invariant: computeMassMatrix == computeKineticTerms.M
    assert_close(ours["M"], r.mass_matrix(q), what=f"{name}:M vs computeMassMatrix")


@pytest.mark.parametrize("name,n", [("wam_arm", 3000), ("humanoid", 40)])
def test_kinetic_terms_device_pointers_and_subsets(name, n, refmod, cuda_device):
    import torch

    g, m, r = _models(name, refmod, cuda_device)
    q, qd = workload.sample_states(g.layout(), n, with_velocity=True)
    dq = torch.from_numpy(q).cuda(cuda_device)
    dqd = torch.from_numpy(qd).cuda(cuda_device)
    ref = r.kinetic_terms(q, qd)
    full = m.kinetic_terms(dq, dqd)
    torch.cuda.synchronize()
    for k in ("M", "h", "Fg", "E"):
        assert_close(full[k].cpu().numpy(), ref[k], what=f"{name}:{k}")
    only_m = m.kinetic_terms(dq, want=("M",))
    assert_close(only_m["M"].cpu().numpy(), ref["M"], what=f"{name}:M only")
    # without velocities h is zero and E is the potential energy
    static = m.kinetic_terms(dq, want=("h", "E", "Fg"))
    ref_static = r.kinetic_terms(q, np.zeros_like(q), want=("h", "E", "Fg"))
    assert np.all(static["h"].cpu().numpy() == 0.0)
    assert_close(static["E"].cpu().numpy(), ref_static["E"], what=f"{name}:E static")
    assert_close(static["Fg"].cpu().numpy(), ref_static["Fg"], what=f"{name}:Fg static")


def test_gravity_identity(refmod, cuda_device):
    """F_gravity == J_cog^T (0, 0, -m g)  (bin/TestGraph.cpp:314-352)."""
    g, m, r = _models("humanoid", refmod, cuda_device)
    q = workload.sample_states(g.layout(), 32)
    fg = m.kinetic_terms(q, want=("Fg",))["Fg"]
    _, jcog = r.cog(q)
    expect = jcog[:, 2, :] * (-9.81 * g.mass())
    assert_close(fg, expect, tol=1e-11, what="gravity identity")


def test_mass_matrix_is_symmetric_positive_definite(cuda_device):
    g = rcs_b200.Graph(graph_file("humanoid"))
    m = rcs_b200.Model(g, cuda_device)
    q = workload.sample_states(g.layout(), 65)
    M = m.kinetic_terms(q, want=("M",))["M"]
    assert_close(M, np.transpose(M, (0, 2, 1)), what="M symmetric (symmetric inertia tensors)")
    assert np.all(np.linalg.eigvalsh(M) > 0.0)


def test_ragged_batches(refmod, cuda_device):
    g, m, r = _models("wam_arm", refmod, cuda_device)
    for n in (1, 31, 33, 65):
        q, qd = workload.sample_states(g.layout(), n, first=100 + n, with_velocity=True)
        ours = m.kinetic_terms(q, qd)
        ref = r.kinetic_terms(q, qd)
        for k in ("M", "h", "Fg", "E"):
            assert_close(ours[k], ref[k], what=f"N={n}:{k}")
    assert m.kinetic_terms(np.empty((0, g.dof)))["M"].shape == (0, g.nJ, g.nJ)


@pytest.mark.parametrize("name,n", [("wam_arm", 300), ("wam_scenario", 100), ("humanoid", 64), ("husky", 64)])
def test_cog_and_cog_jacobian(name, n, refmod, cuda_device):
    """RcsGraph_COG / RcsGraph_COGJacobian (Rcs_kinematics.c:904, :973) over a batch."""
    import torch

    g, m, r = _models(name, refmod, cuda_device)
    q = workload.sample_states(g.layout(), n, first=77)
    cog_ref, jcog_ref = r.cog(q)
    ours = m.cog(q)
    assert_close(ours["cog"], cog_ref, what=f"{name}:cog")
    assert_close(ours["Jcog"], jcog_ref, what=f"{name}:Jcog")
    assert_structural_zeros(ours["Jcog"], jcog_ref, what=f"{name}:Jcog")
    dev = m.cog(torch.from_numpy(q).cuda(cuda_device), want=("Jcog",))
    assert_close(dev["Jcog"].cpu().numpy(), jcog_ref, what=f"{name}:Jcog (device pointers)")
    # rcsb_kinetic_terms_cog: the same from the pass that also builds M
    launches0 = rcs_b200.kernel_launch_count()
    both = m.kinetic_terms(q, want=("M", "Jcog", "cog"))
    assert rcs_b200.kernel_launch_count() - launches0 == 2
    assert np.array_equal(both["Jcog"], ours["Jcog"]) and np.array_equal(both["cog"], ours["cog"])
    assert np.array_equal(both["M"], m.kinetic_terms(q, want=("M",))["M"])


@pytest.mark.parametrize("name,n", [("wam_arm", 2000), ("wam_scenario", 300), ("humanoid", 200), ("husky", 100)])
def test_direct_dynamics(name, n, refmod, cuda_device):
    """Rcs_directDynamics (Rcs_dynamics.c:673): M qdd = F_gravity + F_ext - F_jnt + h. The solve
    goes through M^-1, so the bar is scaled by the condition number of M per state:
    |ours - ref|_inf <= 1e-12 max(1, cond(M) / 100) max(1, |ref|_inf)."""
    import torch

    g, m, r = _models(name, refmod, cuda_device)
    q, qd = workload.sample_states(g.layout(), n, first=31, with_velocity=True)
    rng = np.random.default_rng(8)
    fe, fj = rng.normal(size=(n, g.nJ)), rng.normal(size=(n, g.nJ))
    ref_qdd, ref_E = r.direct_dynamics(q, qd, fe, fj)
    ours = m.direct_dynamics(q, qd, fe, fj)
    cond = np.linalg.cond(r.kinetic_terms(q, qd, want=("M",))["M"])
    err = np.abs(ours["qdd"] - ref_qdd).max(axis=1)
    tol = 1e-12 * np.maximum(1.0, cond / 100.0) * np.maximum(1.0, np.abs(ref_qdd).max(axis=1))
    assert np.all(np.isfinite(ours["qdd"]))
    assert np.all(err <= tol), f"{name}: worst err/tol {float((err / tol).max()):.2f}"
    assert_close(ours["E"], ref_E, what=f"{name}: E")
    # residual check that needs no oracle: M qdd = b
    kt = m.kinetic_terms(q, qd)
    b = kt["Fg"] + fe - fj + kt["h"]
    res = np.einsum("nij,nj->ni", np.tril(kt["M"]) + np.tril(kt["M"], -1).transpose(0, 2, 1), ours["qdd"]) - b
    assert np.all(np.abs(res).max(axis=1) <= 1e-9 * np.maximum(1.0, np.abs(b).max(axis=1)))
    # device pointers, no external forces
    dev = m.direct_dynamics(torch.from_numpy(q).cuda(cuda_device), torch.from_numpy(qd).cuda(cuda_device))
    ref0, _ = r.direct_dynamics(q, qd)
    err0 = np.abs(dev["qdd"].cpu().numpy() - ref0).max(axis=1)
    tol0 = 1e-12 * np.maximum(1.0, cond / 100.0) * np.maximum(1.0, np.abs(ref0).max(axis=1))
    assert np.all(err0 <= tol0)


@pytest.mark.parametrize("name,effs,n", [("wam_arm", ["BallTip"], 5000), ("wam_arm", ["BallTip", "Link4"], 333),
                                         ("schunk_sdh", ["DistalTouch3"], 500), ("humanoid", ["RightHandTip"], 64),
                                         ("husky", None, 64), ("gantry", ["Tool"], 2000),
                                         ("gantry", ["Tool", "Camera", "Trolley"], 333)])
def test_fused_fk_jacobians_mass_matrix(name, effs, n, refmod, cuda_device):
    """rcsb_fk_jacobians_mass: FK + Jacobians + M(q) from one launch (nJ <= 8, symmetric inertia:
    wam_arm and gantry = serial chains, composite variant, the latter with translational joints;
    schunk_sdh = branching hand, accumulator variant) or from the two general paths (humanoid: nJ 34; husky: asymmetric
    tensors) — same outputs either way, against the reference."""
    import torch

    g, m, r = _models(name, refmod, cuda_device)
    eff = [g.body_index(e) for e in effs] if effs else [g.n_bodies - 1]
    q = workload.sample_states(g.layout(), n, first=17)
    launches0 = rcs_b200.kernel_launch_count()
    ours = m.fk_jacobians(torch.from_numpy(q).cuda(cuda_device), eff, want=("A_BI", "JT", "JR", "M"))
    torch.cuda.synchronize()
    launches = rcs_b200.kernel_launch_count() - launches0
    fused = name in ("wam_arm", "schunk_sdh", "gantry")
    assert launches == (1 if fused else 3)
    ref = r.fk_jacobians(q, eff)
    for k in ("A_BI", "JT", "JR"):
        assert_close(ours[k].cpu().numpy(), ref[k], what=f"{name}:{k}")
    M_ref = r.mass_matrix(q)
    assert_close(ours["M"].cpu().numpy(), M_ref, what=f"{name}:M")
    assert_structural_zeros(ours["M"].cpu().numpy(), M_ref, what=f"{name}:M")
    # without the frames among the outputs a serial chain cannot read them back: the
    # per-column accumulator variant runs instead (same result up to rounding)
    host = m.fk_jacobians(q, eff, want=("JT", "M"))
    assert_close(host["M"], M_ref, what=f"{name}:M (frames not requested)")
    assert_close(host["JT"], ref["JT"], what=f"{name}:JT (frames not requested)")
    # a body selection that leaves out massive bodies: same fallback
    sel = m.fk_jacobians(q, eff, bodies=[eff[0]], want=("A_BI", "JR", "M"))
    assert_close(sel["M"], M_ref, what=f"{name}:M (one frame requested)")
    assert_close(sel["JR"], ref["JR"], what=f"{name}:JR (one frame requested)")
