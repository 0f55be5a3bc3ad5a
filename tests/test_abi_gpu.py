"""A graph created by the REFERENCE library (libRcsCore's own RcsGraph_create, exported by
oracle/_ref/librcsref.so) crosses the C boundary unchanged: the same pointer goes to
rcsb_model_create and to this library's single-state RcsGraph_* functions, and the results equal
those of the file route (our loader) and of the reference itself. include/rcsb_graph.h is binary
compatible with src/RcsCore/Rcs_typedef.h:66-259 (tests/test_abi_cpu.py pins the offsets)."""
import ctypes

import numpy as np
import pytest

import rcs_b200
from rcs_b200 import workload
from conftest import assert_close, graph_file

pytestmark = pytest.mark.gpu


class _MatNd(ctypes.Structure):
    _fields_ = [("ele", ctypes.POINTER(ctypes.c_double)), ("m", ctypes.c_uint), ("n", ctypes.c_uint),
                ("size", ctypes.c_uint), ("stackMem", ctypes.c_bool)]


def _reference_graph(refmod, name):
    """RcsGraph* made by the reference's loader, and the ctypes handle of the reference library."""
    R = refmod.lib()
    R.RcsGraph_create.restype = ctypes.c_void_p
    R.RcsGraph_create.argtypes = [ctypes.c_char_p]
    R.RcsGraph_destroy.argtypes = [ctypes.c_void_p]
    g = R.RcsGraph_create(graph_file(name).encode())
    assert g, "the reference failed to load the graph"
    return R, g


@pytest.mark.parametrize("name", ["wam_arm", "humanoid", "wam_scenario"])
def test_reference_created_graph_through_rcsb_model_create(refmod, cuda_device, name):
    L = rcs_b200.lib()
    R, g = _reference_graph(refmod, name)
    try:
        h = ctypes.c_void_p()
        rc = L.rcsb_model_create(ctypes.c_void_p(g), cuda_device, ctypes.byref(h))
        assert rc == 0, L.rcsb_last_error()
        ours_graph = rcs_b200.Graph(graph_file(name))
        m_file = rcs_b200.Model(ours_graph, cuda_device)
        m_ref = rcs_b200.Model.__new__(rcs_b200.Model)      # wrap the handle made from the reference's graph
        m_ref._h, m_ref.device = h, cuda_device
        m_ref.dof, m_ref.nJ, m_ref.n_bodies = L.rcsb_model_dof(h), L.rcsb_model_nj(h), L.rcsb_model_nbodies(h)
        assert (m_ref.dof, m_ref.nJ, m_ref.n_bodies) == (m_file.dof, m_file.nJ, m_file.n_bodies)

        lay = ours_graph.layout()
        q, qd = workload.sample_states(lay, 64, with_velocity=True)
        eff = [m_file.n_bodies - 1]
        a = m_ref.fk_jacobians(q, eff)
        b = m_file.fk_jacobians(q, eff)
        for k in ("A_BI", "JT", "JR"):
            assert np.array_equal(a[k], b[k]), k           # same flat model -> same bits
        ka, kb = m_ref.kinetic_terms(q, qd), m_file.kinetic_terms(q, qd)
        for k in ("M", "h", "Fg", "E"):
            assert np.array_equal(ka[k], kb[k]), k
        # and both equal the reference evaluated on ITS graph
        rg = refmod.RefGraph(graph_file(name))
        rr = rg.fk_jacobians(q, eff)
        for k in ("A_BI", "JT", "JR"):
            assert_close(a[k], rr[k], what=f"{name} {k}")
        rk = rg.kinetic_terms(q, qd)
        for k in ("M", "h", "Fg", "E"):
            assert_close(ka[k], rk[k], what=f"{name} {k}")
    finally:
        R.RcsGraph_destroy(g)


def test_single_state_functions_on_a_reference_created_graph(refmod, cuda_device):
    """This library's RcsGraph_setState / RcsGraph_computeKineticTerms write into the structs the
    reference allocated (A_BI of every body, q), and read back parameters edited through them."""
    L = rcs_b200.lib()
    R, g = _reference_graph(refmod, "wam_arm")
    try:
        L.MatNd_create.restype = ctypes.POINTER(_MatNd)
        L.RcsGraph_setState.argtypes = [ctypes.c_void_p, ctypes.POINTER(_MatNd), ctypes.POINTER(_MatNd)]
        L.RcsGraph_computeMassMatrix.argtypes = [ctypes.c_void_p, ctypes.POINTER(_MatNd)]
        L.RcsGraph_computeMassMatrix.restype = None
        L.rcsb_graph_release_device_cache.argtypes = [ctypes.c_void_p]
        L.rcsb_graph_release_device_cache.restype = None
        R.RcsGraph_getBodyByName.restype = ctypes.c_void_p
        R.RcsGraph_getBodyByName.argtypes = [ctypes.c_void_p, ctypes.c_char_p]

        rg = refmod.RefGraph(graph_file("wam_arm"))
        lay = rg.layout()
        q = workload.sample_states(lay, 1, first=5)
        qm = L.MatNd_create(7, 1)
        for i in range(7):
            qm.contents.ele[i] = q[0, i]
        L.RcsGraph_setState(ctypes.c_void_p(g), qm, None)          # OUR kernels, THEIR graph
        # the reference reads the frame our library stored in its RcsBody
        tip = R.RcsGraph_getBodyByName(g, b"BallTip")
        off = 8 + 4 + 4 + 24 + 24 + 8 + 3 * 8 + 8                    # RcsBody.A_BI (tests/golden/abi_layout.json)
        import json, os
        with open(os.path.join(os.path.dirname(__file__), "golden", "abi_layout.json")) as f:
            assert json.load(f)["RcsBody.A_BI"][0] == off
        a_bi = ctypes.cast(ctypes.c_void_p.from_address(tip + off).value, ctypes.POINTER(ctypes.c_double * 12))
        want = rg.fk(q)["A_BI"][0, rg.body_index("BallTip")]
        assert_close(np.array(a_bi.contents[:]), want, what="A_BI written into the reference's RcsBody")

        M = L.MatNd_create(7, 7)
        L.RcsGraph_computeMassMatrix(ctypes.c_void_p(g), M)
        M0 = np.array([M.contents.ele[i] for i in range(49)]).reshape(7, 7)
        assert_close(M0, rg.mass_matrix(q)[0], what="M on the reference's graph")

        # edit a mass through the reference's struct: the next call must see it (no stale model)
        ctypes.c_double.from_address(tip).value = 2.5                # RcsBody.m is the first field
        L.RcsGraph_computeMassMatrix(ctypes.c_void_p(g), M)
        M1 = np.array([M.contents.ele[i] for i in range(49)]).reshape(7, 7)
        R.ref_mass_matrix.restype = ctypes.c_double
        Mr = np.empty((1, 7, 7))
        qq = np.ascontiguousarray(q)
        R.ref_mass_matrix(ctypes.c_void_p(g), 1, 7, ctypes.c_void_p(qq.ctypes.data), ctypes.c_void_p(Mr.ctypes.data))
        assert np.max(np.abs(M1 - M0)) > 1e-3, "the edited mass did not reach the device model"
        assert_close(M1, Mr[0], what="M after editing a body mass")
        L.rcsb_graph_release_device_cache(ctypes.c_void_p(g))
    finally:
        R.RcsGraph_destroy(g)
