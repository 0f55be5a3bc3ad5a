"""GPU parity: batched FK / Jacobians through the C ABI vs the reference library.

Reference calls: RcsGraph_setState (Rcs_graph.c:231), RcsGraph_bodyPointJacobian
(Rcs_kinematics.c:200), RcsGraph_rotationJacobian (Rcs_kinematics.c:416).
Tolerance: 1e-12 relative to the array's inf-norm (floor 1), as BASELINE.json states.
"""
import numpy as np
import pytest

import rcs_b200
from rcs_b200 import workload
from conftest import GRAPHS, assert_close, assert_structural_zeros, graph_file

pytestmark = pytest.mark.gpu


def _models(name, refmod, cuda_device):
    g = rcs_b200.Graph(graph_file(name))
    return g, rcs_b200.Model(g, cuda_device), refmod.RefGraph(graph_file(name))


@pytest.mark.parametrize("name", GRAPHS)
def test_fk_all_bodies_host_pointers(name, refmod, cuda_device):
    g, m, r = _models(name, refmod, cuda_device)
    q, qd = workload.sample_states(g.layout(), 777, with_velocity=True)
    ours = m.fk(q, qd, want=("A_BI", "A_JI", "xdot", "omega", "q"))
    ref = r.fk(q, qd, want=("A_BI", "A_JI", "xdot", "omega", "q"))
    for k in ("A_BI", "A_JI", "xdot", "omega", "q"):
        assert_close(ours[k], ref[k], what=f"{name}:{k}")


@pytest.mark.parametrize("name", ["wam_arm", "wam_scenario", "humanoid"])
def test_fk_ik_state_vector(name, refmod, cuda_device):
    """q of size nJ is expanded like RcsGraph_stateVectorFromIK (Rcs_graph.c:628)."""
    g, m, r = _models(name, refmod, cuda_device)
    lay = g.layout()
    q = workload.sample_states(lay, 300)
    cols = [j["jointIndex"] for j in lay["joints"] if j["jacobiIndex"] >= 0]
    q_ik = np.ascontiguousarray(q[:, cols])
    ours = m.fk(q_ik, want=("A_BI", "q"))
    ref = r.fk(q_ik, want=("A_BI", "q"))
    assert_close(ours["A_BI"], ref["A_BI"], what=f"{name}:A_BI")
    assert_close(ours["q"], ref["q"], what=f"{name}:q")


@pytest.mark.parametrize("name,effs", [
    ("wam_arm", ["BallTip"]),
    ("wam_scenario", ["Hand", "PowerGrip", "Base"]),
    ("humanoid", ["RightHandTip", "LeftHandTip", "Helmet", "FootL"]),
    ("dexbot", None),
    ("husky", None),
])
def test_fk_jacobians_device_pointers(name, effs, refmod, cuda_device):
    import torch

    g, m, r = _models(name, refmod, cuda_device)
    if effs is None:
        effs = [g.body_names()[-1], g.body_names()[len(g.body_names()) // 2]]
    eff = [g.body_index(e) for e in effs]
    rng = np.random.default_rng(5)
    pts = rng.uniform(-0.2, 0.2, size=(len(eff), 3))
    q = workload.sample_states(g.layout(), 1000 + 17)
    qd_dev = torch.from_numpy(q).to(f"cuda:{cuda_device}")
    ours = m.fk_jacobians(qd_dev, eff, points=pts)
    torch.cuda.synchronize()
    ref = r.fk_jacobians(q, eff, points=pts)
    for k in ("A_BI", "JT", "JR"):
        o = ours[k].cpu().numpy()
        assert_close(o, ref[k], what=f"{name}:{k}")
        if k != "A_BI":
            assert_structural_zeros(o, ref[k], what=f"{name}:{k}")


def test_body_selection_and_origin_points(refmod, cuda_device):
    g, m, r = _models("humanoid", refmod, cuda_device)
    sel = [g.body_index(n) for n in ("RightHandTip", "UpperBody", "FootR")]
    q = workload.sample_states(g.layout(), 129)
    ours = m.fk_jacobians(q, [sel[0]], bodies=sel)
    ref = r.fk_jacobians(q, [sel[0]])
    assert_close(ours["A_BI"], ref["A_BI"][:, sel, :], what="selected frames")
    assert_close(ours["JT"], ref["JT"], what="JT at body origin")
    assert_close(ours["JR"], ref["JR"], what="JR")


def test_empty_and_ragged_batches(refmod, cuda_device):
    g, m, r = _models("wam_arm", refmod, cuda_device)
    eff = [g.body_index("BallTip")]
    assert m.fk(np.empty((0, g.dof)))["A_BI"].shape == (0, g.n_bodies, 12)
    for n in (1, 31, 33, 127, 129):
        q = workload.sample_states(g.layout(), n, first=n)
        ours = m.fk_jacobians(q, eff)
        ref = r.fk_jacobians(q, eff)
        for k in ("A_BI", "JT", "JR"):
            assert_close(ours[k], ref[k], what=f"N={n}:{k}")


def test_wrong_dimensions_is_an_error(cuda_device):
    """The reference aborts with 'q has wrong dimensions' (Rcs_graph.c:250); we return it."""
    g = rcs_b200.Graph(graph_file("wam_arm"))
    m = rcs_b200.Model(g, cuda_device)
    with pytest.raises(rcs_b200.RcsbError, match="wrong dimensions"):
        m.fk(np.zeros((4, g.dof + 1)))


def test_known_answers_wam_default_state(cuda_device):
    """Values recorded from the reference at survey time."""
    g = rcs_b200.Graph(graph_file("wam_arm"))
    m = rcs_b200.Model(g, cuda_device)
    q = np.array([g.layout()["q"]])
    tip = g.body_index("BallTip")
    out = m.fk_jacobians(q, [tip])
    org = out["A_BI"][0, tip, :3]
    np.testing.assert_allclose(
        org, [0.83632952411278927, 0.30514499825747998, 0.83766965256912063], rtol=0, atol=1e-14)
    np.testing.assert_allclose(
        out["JT"][0, 0, 0],
        [-0.16514499825747997, 0.47491641541908702, -0.14053874149066872,
         -0.040723861313878298, 6.9e-18, 0.033210789380311002, 1.7e-17], rtol=0, atol=1e-14)
    np.testing.assert_allclose(
        out["JR"][0, 0, 0],
        [0, -0.25881904510252074, 0.16773125949652062, -0.25881904510252074,
         0.95125124256419769, 0.25881904510252085, 0.95125124256419769], rtol=0, atol=1e-14)


def _rel_spec(g, name):
    n = g.body_names()
    if name == "humanoid":
        a, b, c = n.index("RightHandTip"), n.index("LeftHandTip"), n.index("UpperBody")
        fixed = n.index("Heel")
    else:
        a, b, c = n.index("Hand"), n.index("Base"), n.index("PowerGrip")
        fixed = 0
    return [("POS", a, -1, -1), ("POS", a, -1, c), ("POS", a, b, b), ("POS", a, b, c), ("POS", a, c, -1),
            ("POS", -1, b, c), ("OMEGA", a, -1, -1), ("OMEGA", a, b, b), ("OMEGA", a, b, c),
            ("OMEGA", a, fixed, fixed), ("OMEGA", a, -1, c), ("OMEGA", a, b, -1)]


@pytest.mark.parametrize("name", ["humanoid", "wam_scenario"])
def test_relative_task_jacobians(name, refmod, cuda_device):
    """RcsGraph_3dPosJacobian / RcsGraph_3dOmegaJacobian (Rcs_kinematics.c:566, :789): world,
    rotated, relative to an articulated refBdy, with a separate refFrame, NULL effector."""
    import torch

    g, m, r = _models(name, refmod, cuda_device)
    spec = _rel_spec(g, name)
    q = workload.sample_states(g.layout(), 40000 if name == "wam_scenario" else 500, first=12)
    ref = r.task_jacobians(q[:500], spec)
    ours = m.task_jacobians(q, spec)
    assert_close(ours[:500], ref, what=f"{name}: task Jacobians (host pointers)")
    assert_structural_zeros(ours[:500], ref, what=f"{name}: task Jacobians")
    dev = m.task_jacobians(torch.from_numpy(q).cuda(cuda_device), spec)
    torch.cuda.synchronize()
    assert np.array_equal(dev.cpu().numpy(), ours)


def test_whole_body_task_stack_matches_controller(refmod, cuda_device):
    """BASELINE.json configs[4] shape: the active task set of the reference's
    GenericHumanoid/cAction.xml (hand positions, COG x/y, both feet as 6-D poses relative to the
    Heel: nx = 20) stacked from rcsb_task_jacobians + rcsb_cog, against the reference's own
    ControllerBase::computeJ (ControllerBase.cpp:776) with its task classes."""
    g = rcs_b200.Graph(graph_file("humanoid"))
    m = rcs_b200.Model(g, cuda_device)
    tasks = [("XYZ", "RightHandTip"), ("XYZ", "LeftHandTip"), ("COGX", "Heel"), ("COGY", "Heel"),
             ("XYZABC", "FootL", "Heel"), ("XYZABC", "FootR", "Heel")]
    ctrl = refmod.RefIk(graph_file("humanoid"), tasks)
    assert ctrl.nx == 20
    q = workload.sample_states(g.layout(), 700, first=5)
    _, J_ref = ctrl.task_kinematics(q)
    rh, lh, heel, fl, fr = [g.body_index(n) for n in ("RightHandTip", "LeftHandTip", "Heel", "FootL", "FootR")]
    Jt = m.task_jacobians(q, [("POS", rh, -1, -1), ("POS", lh, -1, -1), ("POS", fl, heel, heel),
                              ("OMEGA", fl, heel, heel), ("POS", fr, heel, heel), ("OMEGA", fr, heel, heel)])
    Jc = m.cog(q, want=("Jcog",))["Jcog"]
    J = np.concatenate([Jt[:, 0], Jt[:, 1], Jc[:, 0:1], Jc[:, 1:2], Jt[:, 2], Jt[:, 3], Jt[:, 4], Jt[:, 5]],
                       axis=1)
    assert_close(J, J_ref, what="stacked task Jacobian (20 x 34)")
    assert_structural_zeros(J, J_ref, what="stacked task Jacobian")


def test_too_many_effector_chains_is_a_clear_error(cuda_device):
    """The Jacobian scratch of fkKernel lives in shared memory: effectors whose chains need more
    than 227 KB are refused with a message (rcsb_task_jacobians splits them into groups itself)."""
    g = rcs_b200.Graph(graph_file("humanoid"))
    m = rcs_b200.Model(g, cuda_device)
    effs = [g.body_index(n) for n in ("RightHandTip", "LeftHandTip", "FootL", "FootR", "Helmet")]
    with pytest.raises(rcs_b200.RcsbError, match="shared memory"):
        m.fk_jacobians(np.zeros((4, g.dof)), effs)


@pytest.mark.parametrize("name,n", [("wam_arm", 1), ("wam_arm", 33), ("wam_arm", 1000), ("gantry", 257),
                                    ("humanoid", 95)])
def test_unaligned_outputs_ragged_batches_and_body_selections(name, n, refmod, cuda_device):
    """Output arrays that are only 8-byte aligned take the scalar store / load paths (frames
    written and, for the fused mass matrix of a serial chain, read back); batch sizes that are
    not a multiple of the warp or block size leave idle lanes; a body selection walks the
    pruned sub-tree. Same results in every combination."""
    import torch

    g, m, r = _models(name, refmod, cuda_device)
    q = workload.sample_states(g.layout(), n, first=3)
    dq = torch.from_numpy(q).cuda(cuda_device)
    eff = [g.n_bodies - 1]
    ref = r.fk_jacobians(q, eff)
    want = ("A_BI", "JT", "JR", "M") if g.nJ <= 8 else ("A_BI", "JT", "JR")

    def odd(shape):
        """torch tensor whose data pointer is 8 bytes past a 32-byte boundary"""
        flat = torch.zeros(int(np.prod(shape)) + 4, dtype=torch.float64, device=dq.device)
        assert flat.data_ptr() % 32 == 0
        return flat[1:1 + int(np.prod(shape))].view(shape)

    shapes = {"A_BI": (n, g.n_bodies, 12), "JT": (n, 1, 3, g.nJ), "JR": (n, 1, 3, g.nJ), "M": (n, g.nJ, g.nJ)}
    out = {k: odd(shapes[k]) for k in want}
    assert all(v.data_ptr() % 32 == 8 for v in out.values())
    m.fk_jacobians(dq, eff, want=want, out=out)
    torch.cuda.synchronize()
    for k in ("A_BI", "JT", "JR"):
        assert_close(out[k].cpu().numpy(), ref[k], what=f"{name}:{k} (unaligned)")
    if "M" in want:
        assert_close(out["M"].cpu().numpy(), r.mass_matrix(q), what=f"{name}:M (unaligned)")
    # pruned walk: the effector and one more body, unaligned as well
    sel = [eff[0], g.n_bodies // 2]
    frames = odd((n, 2, 12))
    pr = m.fk_jacobians(dq, eff, bodies=sel, want=("A_BI", "JT"), out={"A_BI": frames})
    torch.cuda.synchronize()
    assert_close(pr["A_BI"].cpu().numpy(), ref["A_BI"][:, sel], what=f"{name}: selected frames")
    assert_close(pr["JT"].cpu().numpy(), ref["JT"], what=f"{name}: JT from the pruned walk")
    only = m.fk(dq, bodies=[g.n_bodies // 3])["A_BI"]
    assert_close(only.cpu().numpy(), ref["A_BI"][:, [g.n_bodies // 3]], what=f"{name}: one frame")


@pytest.mark.parametrize("name,eff_name", [("wam_arm", "BallTip"), ("wam_arm", "Link4"), ("wrist6", "Tip"),
                                           ("wrist6", "Fore")])
def test_chain_walk_matches_the_interpreting_kernel_and_the_reference(name, eff_name, refmod, cuda_device,
                                                                      monkeypatch):
    """Serial chains of rotations take fkChainKernel (straight-line walk over folded transforms, the
    mass matrix tip to root by inverse transforms, rcsb_plan.h: chainNc); RCSB_FK_GENERAL=1 forces
    fkKernel, which interprets the flat model. Both against the reference: frames, Jacobians of a
    body-fixed point, M; all bodies and a body selection; wrist6 has several joint axes per body, so
    its bases carry row shifts."""
    g, m, r = _models(name, refmod, cuda_device)
    eff = [g.body_index(eff_name)]
    pts = np.array([[0.03, -0.02, 0.05]])
    q = workload.sample_states(g.layout(), 1500 + 7, first=13)
    ref = r.fk_jacobians(q, eff, points=pts)
    M_ref = r.mass_matrix(q)
    sel = [g.n_bodies - 1, 1, eff[0]] if eff[0] not in (1, g.n_bodies - 1) else [g.n_bodies - 1, 1]
    results = {}
    for mode in ("chain", "general"):
        if mode == "general":
            monkeypatch.setenv("RCSB_FK_GENERAL", "1")
        launches = rcs_b200.kernel_launch_count()
        full = m.fk_jacobians(q, eff, points=pts, want=("A_BI", "JT", "JR", "M"))
        assert rcs_b200.kernel_launch_count() - launches == 1
        part = m.fk_jacobians(q, eff, points=pts, bodies=sel, want=("A_BI", "JT", "JR"))
        only = m.fk(q, bodies=[g.n_bodies // 2])
        results[mode] = (full, part, only)
        for k in ("A_BI", "JT", "JR"):
            assert_close(full[k], ref[k], what=f"{name}/{mode}:{k}")
        assert_structural_zeros(full["JT"], ref["JT"], what=f"{name}/{mode}:JT")
        assert_structural_zeros(full["JR"], ref["JR"], what=f"{name}/{mode}:JR")
        assert_close(full["M"], M_ref, what=f"{name}/{mode}:M")
        assert np.array_equal(full["M"], np.swapaxes(full["M"], 1, 2)), "M symmetric bit for bit"
        assert_close(part["A_BI"], ref["A_BI"][:, sel], what=f"{name}/{mode}: selected frames")
        assert_close(part["JT"], ref["JT"], what=f"{name}/{mode}: JT with a selection")
        assert_close(only["A_BI"], ref["A_BI"][:, [g.n_bodies // 2]], what=f"{name}/{mode}: one frame")
    # the two kernels agree far below the parity bar
    for k in ("A_BI", "JT", "JR", "M"):
        d = np.abs(results["chain"][0][k] - results["general"][0][k]).max()
        assert d <= 1e-14 * max(1.0, np.abs(results["general"][0][k]).max()), (k, d)


@pytest.mark.parametrize("name,n", [("wam_arm", 1), ("wam_arm", 1000 + 13), ("wrist6", 777), ("wrist6", 64)])
def test_chain_walk_with_velocities(name, n, refmod, cuda_device, monkeypatch):
    """rcsb_fk with q_dot on a serial chain of rotations: fkChainVelKernel (the base carries its twist
    through the folded transforms; x_dot / omega of full warps leave as bulk stores, the last partial
    warp and unaligned outputs store directly) against fkKernel<VEL> (RCSB_FK_GENERAL=1) and the
    reference; all bodies, a body selection in another order, device pointers 8 bytes off alignment."""
    import torch

    g, m, r = _models(name, refmod, cuda_device)
    q, qd = workload.sample_states(g.layout(), n, first=7, with_velocity=True)
    ref = r.fk(q, qd, want=("A_BI", "xdot", "omega"))
    sel = [g.n_bodies - 1, 2, 1]
    res = {}
    for mode in ("chain", "general"):
        if mode == "general":
            monkeypatch.setenv("RCSB_FK_GENERAL", "1")
        launches = rcs_b200.kernel_launch_count()
        full = m.fk(q, qd, want=("A_BI", "xdot", "omega"))
        assert rcs_b200.kernel_launch_count() - launches == 1
        part = m.fk(q, qd, bodies=sel, want=("xdot", "omega", "A_BI"))
        only_om = m.fk(q, qd, want=("omega",))
        res[mode] = full
        for k in ("A_BI", "xdot", "omega"):
            assert_close(full[k], ref[k], what=f"{name}/{mode}:{k}")
            assert_close(part[k], ref[k][:, sel], what=f"{name}/{mode}: selected {k}")
        assert_close(only_om["omega"], ref["omega"], what=f"{name}/{mode}: omega alone")
        # outputs that are only 8-byte aligned
        dq_, dqd_ = torch.from_numpy(q).cuda(cuda_device), torch.from_numpy(qd).cuda(cuda_device)
        flat = torch.zeros(n * g.n_bodies * 3 + 4, dtype=torch.float64, device=dq_.device)
        xo = flat[1:1 + n * g.n_bodies * 3].view(n, g.n_bodies, 3)
        m.fk(dq_, dqd_, want=("xdot",), out={"xdot": xo})
        torch.cuda.synchronize()
        assert_close(xo.cpu().numpy(), ref["xdot"], what=f"{name}/{mode}: xdot (unaligned)")
    monkeypatch.delenv("RCSB_FK_GENERAL")
    for k in ("xdot", "omega"):
        d = np.abs(res["chain"][k] - res["general"][k]).max()
        assert d <= 1e-13 * max(1.0, np.abs(res["general"][k]).max()), (k, d)


def test_folded_kernels_with_angles_beyond_the_fast_sincos_range(refmod, cuda_device, monkeypatch):
    """sincosBatch serves |q| < 2^31; beyond that (and for non-finite values) the folding kernels call the
    library's sincos like the interpreting kernels do: same frames, Jacobians, M and IK step for joint
    values of 3e9 and 1e15 rad as from the reference (libm), NaN in -> NaN out without a fault."""
    from rcs_b200 import TASK_ABC, TASK_XYZ

    g, m, r = _models("wam_arm", refmod, cuda_device)
    tip = g.body_index("BallTip")
    q = workload.sample_states(g.layout(), 96, first=3)
    q[0::4, 2] = 3.0e9
    q[1::4, 5] = -1.0e15
    q[2::4, 0] = 2147483648.0
    ref = r.fk_jacobians(q, [tip])
    M_ref = r.mass_matrix(q)
    res = {}
    for mode in ("chain", "general"):
        if mode == "general":
            monkeypatch.setenv("RCSB_FK_GENERAL", "1")
        res[mode] = m.fk_jacobians(q, [tip], want=("A_BI", "JT", "JR", "M"))
        for k in ("A_BI", "JT", "JR"):
            assert_close(res[mode][k], ref[k], what=f"{mode}:{k}")
        assert_close(res[mode]["M"], M_ref, what=f"{mode}:M")
    monkeypatch.delenv("RCSB_FK_GENERAL")
    qn = q.copy()
    qn[5, 3] = np.nan
    out = m.fk_jacobians(qn, [tip], want=("A_BI", "JT", "JR", "M"))
    assert np.isnan(out["A_BI"][5]).any() and np.isfinite(out["A_BI"][6]).all()
    # one IK step from such states (targets: the effector poses of moderate states)
    rik = refmod.RefIk(graph_file("wam_arm"), [("XYZ", "BallTip"), ("ABC", "BallTip")])
    x_des, _ = rik.task_kinematics(workload.sample_states(g.layout(), 96, first=50, margin=0.2))
    x_des = np.ascontiguousarray(x_des)
    q_ref, dx_ref, dq_ref, det_ref = rik.rmr(q, x_des, lam=1.0e-3, iters=1)
    qq = q.copy()
    o = m.ik_rmr([(TASK_XYZ, tip), (TASK_ABC, tip)], qq, x_des, lam=1.0e-3, iters=1, want=("dx", "dq", "det"))
    assert_close(o["dx"], dx_ref, what="dx from huge angles")
    assert_close(o["dq"], dq_ref, tol=1e-9, what="dq from huge angles")


def test_polynomial_couplings_inside_and_beyond_the_master_limits(refmod, cuda_device):
    """5- and 9-coefficient coupled joints (RcsJoint_computeSlaveJointAngle / Velocity,
    src/RcsCore/Rcs_joint.c:375-462): inside the master's range the polynomial, beyond it the
    tangent at the limit; velocities with the reference's clipped-master slope. The kernels
    evaluate powers by repeated multiplication where the reference calls pow(): the results
    must still agree within the 1e-12 bar. Masters sweep from 40 % below to 40 % above their
    range, limits included exactly."""
    name = "polycoupled"
    g = rcs_b200.Graph(graph_file(name))
    m = rcs_b200.Model(g, cuda_device)
    r = refmod.RefGraph(graph_file(name))
    lay = g.layout()
    n = 1200
    q, qd = workload.sample_states(lay, n, first=77, with_velocity=True)
    masters = sorted({j["coupledTo"] for j in lay["joints"] if j["coupledTo"] >= 0})
    assert len(masters) == 2
    assert sorted(len(j["couplingFactors"]) for j in lay["joints"] if j["coupledTo"] >= 0) == [5, 5, 9]
    for mj in masters:
        lo, hi = lay["joints"][mj]["q_min"], lay["joints"][mj]["q_max"]
        q[:, mj] = np.linspace(lo - 0.4 * (hi - lo), hi + 0.4 * (hi - lo), n)
        q[0, mj], q[1, mj] = lo, hi
    q = np.ascontiguousarray(q)
    ours = m.fk(q, qd, want=("A_BI", "A_JI", "xdot", "omega", "q"))
    ref = r.fk(q, qd, want=("A_BI", "A_JI", "xdot", "omega", "q"))
    for k in ("q", "A_BI", "A_JI", "xdot", "omega"):
        assert_close(ours[k], ref[k], what=f"{name} {k}")
    slaves = [i for i, j in enumerate(lay["joints"]) if j["coupledTo"] >= 0]
    assert np.max(np.abs(ours["q"][:, slaves] - q[:, slaves])) > 0.1      # the slaves really moved
    kt, rk = m.kinetic_terms(q, qd), r.kinetic_terms(q, qd)
    for k in ("M", "h", "Fg", "E"):
        assert_close(kt[k], rk[k], what=f"{name} {k}")
    eff = [g.body_index("Tip"), g.body_index("Probe")]
    jo, jr = m.fk_jacobians(q, eff), r.fk_jacobians(q, eff)
    for k in ("JT", "JR"):
        assert_close(jo[k], jr[k], what=f"{name} {k}")


def test_host_pointer_pipeline_many_stages(refmod, cuda_device, monkeypatch):
    """The host-buffer path cuts a batch into stages that rotate over three streams / staging buffers /
    workspaces (H2D, kernels, D2H of different stages overlap). With 300-state stages a 5,000-state
    batch takes 17 of them: results must equal the device-pointer call bit for bit, for a path with
    per-stage workspaces (kinetic terms), one with an in-out array (IK) and plain FK."""
    import torch

    monkeypatch.setenv("RCSB_STAGE_CHUNK", "300")
    g, m, r = _models("wam_arm", refmod, cuda_device)
    n = 5000
    q, qd = workload.sample_states(g.layout(), n, first=3, with_velocity=True)
    tq, tqd = torch.from_numpy(q).cuda(cuda_device), torch.from_numpy(qd).cuda(cuda_device)
    host = m.fk(q, qd, want=("A_BI", "xdot"))
    dev = m.fk(tq, tqd, want=("A_BI", "xdot"))
    hk = m.kinetic_terms(q, qd)
    dk = m.kinetic_terms(tq, tqd)
    tip = g.body_index("BallTip")
    x_des = np.ascontiguousarray(np.concatenate([host["A_BI"][:, tip, :3] + 0.01, np.zeros((n, 3))], axis=1))
    tasks = [(rcs_b200.TASK_XYZ, tip), (rcs_b200.TASK_ABC, tip)]
    qh = q.copy()
    oh = m.ik_rmr(tasks, qh, x_des, lam=1e-6, iters=3, want=("dq", "det"))
    qdv = tq.clone()
    od = m.ik_rmr(tasks, qdv, torch.from_numpy(x_des).cuda(cuda_device), lam=1e-6, iters=3, want=("dq", "det"))
    torch.cuda.synchronize()
    for k in host:
        assert np.array_equal(host[k], dev[k].cpu().numpy()), k
    for k in hk:
        assert np.array_equal(hk[k], dk[k].cpu().numpy()), k
    assert np.array_equal(qh, qdv.cpu().numpy())
    for k in oh:
        assert np.array_equal(oh[k], od[k].cpu().numpy()), k
