"""GPU parity of rcsb_wholebody (BASELINE.json configs[4]): frames of all bodies + the task
Jacobians of the reference's GenericHumanoid/cAction.xml active set + M(q) + the COG Jacobian from
ONE pass of the kinetic-terms kernels, against the reference evaluated on one graph state
(RcsGraph_setState, ControllerBase::computeJ, RcsGraph_computeMassMatrix) and against the separate
entry points of this library."""
import numpy as np
import pytest

import rcs_b200
from rcs_b200 import workload
from conftest import assert_close, assert_structural_zeros, graph_file

pytestmark = pytest.mark.gpu

# cAction.xml's active set (nx = 20), as in bench.py
WB_TASKS = [("XYZ", "RightHandTip"), ("XYZ", "LeftHandTip"), ("COGX", "Heel"), ("COGY", "Heel"),
            ("XYZABC", "FootL", "Heel"), ("XYZABC", "FootR", "Heel")]


def _spec(g):
    rh, lh, heel, fl, fr = [g.body_index(x) for x in ("RightHandTip", "LeftHandTip", "Heel", "FootL", "FootR")]
    return [("POS", rh, -1, -1), ("POS", lh, -1, -1), ("POS", fl, heel, heel), ("OMEGA", fl, heel, heel),
            ("POS", fr, heel, heel), ("OMEGA", fr, heel, heel)]


def test_wholebody_matches_the_reference_controller(refmod, cuda_device):
    g = rcs_b200.Graph(graph_file("humanoid"))
    m = rcs_b200.Model(g, cuda_device)
    rik = refmod.RefIk(graph_file("humanoid"), WB_TASKS)
    n = 300
    q = workload.sample_states(g.layout(), n, first=11)
    before = rcs_b200.kernel_launch_count()
    ours = m.wholebody(q, _spec(g))
    assert rcs_b200.kernel_launch_count() - before == 2          # walk + assembly
    ref = rik.wholebody(q)
    assert_close(ours["A_BI"], ref["A_BI"], what="A_BI")
    assert_close(ours["M"], ref["M"], what="M")
    assert_structural_zeros(ours["M"], ref["M"], what="M")
    # stacked task Jacobian of the controller: hands (3 + 3), COG x / y rows, feet (6 + 6)
    J = ref["J"]
    assert J.shape == (n, 20, g.nJ)
    assert_close(ours["Jt"][:, 0], J[:, 0:3], what="J XYZ right hand")
    assert_close(ours["Jt"][:, 1], J[:, 3:6], what="J XYZ left hand")
    assert_close(ours["Jcog"][:, 0], J[:, 6], what="J COGX")
    assert_close(ours["Jcog"][:, 1], J[:, 7], what="J COGY")
    assert_close(ours["Jt"][:, 2], J[:, 8:11], what="J XYZ left foot rel. heel")
    assert_close(ours["Jt"][:, 3], J[:, 11:14], what="J ABC left foot rel. heel")
    assert_close(ours["Jt"][:, 4], J[:, 14:17], what="J XYZ right foot rel. heel")
    assert_close(ours["Jt"][:, 5], J[:, 17:20], what="J ABC right foot rel. heel")
    assert_structural_zeros(ours["Jt"].reshape(n, -1, g.nJ), np.concatenate([J[:, 0:6], J[:, 8:20]], axis=1),
                            what="task Jacobians")


@pytest.mark.parametrize("name", ["humanoid", "wam_scenario", "gantry"])
def test_wholebody_equals_the_separate_entry_points(name, refmod, cuda_device):
    """All twelve reference-body / reference-frame combinations of RcsGraph_3dPosJacobian /
    RcsGraph_3dOmegaJacobian, a body selection, device pointers."""
    import torch

    g = rcs_b200.Graph(graph_file(name))
    m = rcs_b200.Model(g, cuda_device)
    r = refmod.RefGraph(graph_file(name))
    nb = g.n_bodies
    a, b, c = nb - 1, nb // 2, nb // 3
    spec = [("POS", a, -1, -1), ("POS", a, -1, c), ("POS", a, b, b), ("POS", a, b, c), ("POS", -1, b, c),
            ("OMEGA", a, -1, -1), ("OMEGA", a, b, b), ("OMEGA", a, b, c), ("OMEGA", a, -1, c),
            ("OMEGA", a, b, -1), ("POS", b, a, a), ("OMEGA", c, a, b)]
    n = 257
    q = workload.sample_states(g.layout(), n, first=5)
    sel = [a, 0, c]
    dq = torch.from_numpy(q).cuda(cuda_device)
    ours = m.wholebody(dq, spec, bodies=sel)
    torch.cuda.synchronize()
    ours = {k: v.cpu().numpy() for k, v in ours.items()}
    assert_close(ours["Jt"], r.task_jacobians(q, spec), what=f"{name} task Jacobians")
    assert_close(ours["A_BI"], r.fk(q)["A_BI"][:, sel], what=f"{name} frames")
    assert_close(ours["M"], r.mass_matrix(q), what=f"{name} M")
    cog, jcog = r.cog(q)
    assert_close(ours["Jcog"], jcog, what=f"{name} Jcog")
    # host pointers, chunked staging, only some outputs
    part = m.wholebody(q, spec[:3], want=("Jt", "M"))
    assert np.array_equal(part["Jt"], ours["Jt"][:, :3]) and np.array_equal(part["M"], ours["M"])
    only_frames = m.wholebody(q, [], want=("A_BI",))
    assert_close(only_frames["A_BI"], r.fk(q)["A_BI"], what=f"{name} all frames")
