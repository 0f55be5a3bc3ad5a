"""Device groups of the C ABI (rcsb_group_*, include/rcsb.h): one process, the GPUs of the node,
host arrays of the whole batch. Results must be those of the single-device entry points on the
same states; the gathered table holds every rank's {count, sum, sum of squares, min, max} of the
per-state value named in the header. Runs with 1 device everywhere and with every device the box
has (NCCL all-gather between them) where there are several."""
import numpy as np
import pytest

import rcs_b200
from rcs_b200 import TASK_ABC, TASK_XYZ, workload
from conftest import graph_file

pytestmark = pytest.mark.gpu


def _device_sets():
    n = rcs_b200.lib().rcsb_device_count()
    sets = [[0]]
    if n >= 2:
        sets.append(list(range(n)))
    return sets


def _check_stats(st, values, group, n):
    for r in range(len(group.devices)):
        first, count = group.slice_of(r, n)
        v = values[first:first + count]
        assert st[r, 0] == count
        np.testing.assert_allclose(st[r, 1], v.sum(), rtol=1e-12)
        np.testing.assert_allclose(st[r, 2], (v * v).sum(), rtol=1e-12)
        assert st[r, 3] == v.min() and st[r, 4] == v.max()


@pytest.mark.parametrize("devices", _device_sets(), ids=lambda d: f"{len(d)}gpu")
def test_group_fk_jacobians_mass_and_ik(devices, cuda_device):
    g = rcs_b200.Graph(graph_file("wam_arm"))
    grp = rcs_b200.Group(g, devices)
    m = rcs_b200.Model(g, 0)
    tip = g.body_index("BallTip")
    n = 10001                                  # not divisible: slices differ by one state
    covered = [grp.slice_of(r, n) for r in range(len(devices))]
    assert covered[0][0] == 0 and sum(c for _, c in covered) == n
    assert all(covered[i][0] + covered[i][1] == covered[i + 1][0] for i in range(len(devices) - 1))

    q = workload.sample_states(g.layout(), n)
    out = {"A_BI": np.empty((n, g.n_bodies, 12)), "JT": np.empty((n, 1, 3, g.nJ)),
           "JR": np.empty((n, 1, 3, g.nJ)), "M": np.empty((n, g.nJ, g.nJ))}
    st = grp.fk_jacobians(q, [tip], out)
    one = m.fk_jacobians(q, [tip], want=("A_BI", "JT", "JR", "M"))
    for k in out:
        assert np.array_equal(out[k], one[k]), k
    _check_stats(st, out["A_BI"][:, 0, 2], grp, n)

    tasks = [(TASK_XYZ, tip), (TASK_ABC, tip)]
    q_goal = workload.sample_states(g.layout(), n, first=50, margin=0.2)
    frames = m.fk(q_goal, bodies=[tip])["A_BI"][:, 0]
    import bench

    x_des = np.ascontiguousarray(np.concatenate([frames[:, :3], bench.euler_xyz(frames[:, 3:])], axis=1))
    q0 = np.ascontiguousarray(q_goal + 0.05)
    qa, qb = q0.copy(), q0.copy()
    o = {"det": np.empty(n), "dq": np.empty((n, g.dof))}
    st = grp.ik_rmr(tasks, qa, x_des, o, iters=10)
    ref = m.ik_rmr(tasks, qb, x_des, iters=10, want=("det", "dq"))
    assert np.array_equal(qa, qb) and np.array_equal(o["det"], ref["det"]) and np.array_equal(o["dq"], ref["dq"])
    _check_stats(st, o["det"], grp, n)


@pytest.mark.parametrize("devices", _device_sets(), ids=lambda d: f"{len(d)}gpu")
def test_group_kinetic_terms_and_wholebody(devices, cuda_device):
    g = rcs_b200.Graph(graph_file("humanoid"))
    grp = rcs_b200.Group(g, devices)
    m = rcs_b200.Model(g, 0)
    n = 777
    q, qd = workload.sample_states(g.layout(), n, with_velocity=True)
    out = {"M": np.empty((n, g.nJ, g.nJ)), "h": np.empty((n, g.nJ)), "Fg": np.empty((n, g.nJ)), "E": np.empty(n)}
    st = grp.kinetic_terms(q, qd, out)
    one = m.kinetic_terms(q, qd)
    for k in out:
        assert np.array_equal(out[k], one[k]), k
    _check_stats(st, out["E"], grp, n)

    rh, heel, fl = g.body_index("RightHandTip"), g.body_index("Heel"), g.body_index("FootL")
    spec = [("POS", rh, -1, -1), ("POS", fl, heel, heel), ("OMEGA", fl, heel, heel)]
    wb = {"A_BI": np.empty((n, g.n_bodies, 12)), "Jt": np.empty((n, 3, 3, g.nJ)),
          "M": np.empty((n, g.nJ, g.nJ)), "Jcog": np.empty((n, 3, g.nJ))}
    st = grp.wholebody(q, spec, wb)
    one = m.wholebody(q, spec)
    for k in wb:
        assert np.array_equal(wb[k], one[k]), k
    _check_stats(st, wb["M"][:, 0, 0], grp, n)


def test_group_rejects_bad_device_lists(cuda_device):
    g = rcs_b200.Graph(graph_file("wam_arm"))
    with pytest.raises(rcs_b200.RcsbError):
        rcs_b200.Group(g, [0, 0])
    with pytest.raises(rcs_b200.RcsbError):
        rcs_b200.Group(g, [10 ** 6])
