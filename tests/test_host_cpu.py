"""Host-side tests (no GPU): the C-ABI library loads and exports every symbol include/*.h
declares, the C graph loader reproduces the reference loader's integer layout bit for bit,
clone / check / XML round trip work, and every compute entry point fails loudly — instead of
falling back to the CPU — when no CUDA device is present.

Reference: RcsGraph_create (src/RcsCore/Rcs_graph.c:398), RcsGraph_clone (:2193),
RcsGraph_check (:1887), loader rules of SURVEY.md Appendix A.
"""
import ctypes
import glob
import json
import os
import re

import numpy as np
import pytest

import rcs_b200
from conftest import GRAPH_DIR, GRAPHS, REFERENCE_XML, ROOT, graph_file

INT_KEYS_BODY = ("parent", "joints", "name")
INT_KEYS_JOINT = ("jointIndex", "jacobiIndex", "type", "dirIdx", "constrained", "coupledTo",
                  "prev", "name")


def declared_functions():
    """Function names declared in include/*.h (extern "C" prototypes)."""
    names = set()
    for h in glob.glob(os.path.join(ROOT, "include", "*.h")):
        text = open(h).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        text = re.sub(r"^\s*#.*$", "", text, flags=re.M)
        for m in re.finditer(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\([^;{}()]*\)\s*;", text):
            names.add(m.group(1))
    return sorted(names)


def test_library_exports_every_declared_symbol():
    lib = rcs_b200.lib()
    names = declared_functions()
    assert {"rcsb_fk", "rcsb_fk_jacobians", "rcsb_kinetic_terms", "rcsb_ik_rmr", "RcsGraph_create",
            "RcsGraph_setState", "RcsGraph_computeKineticTerms"} <= set(names)
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, f"declared in include/*.h but not exported by librcsb.so: {missing}"


def _strip(d, ignore=("nShapes",)):
    if isinstance(d, dict):
        return {k: _strip(v) for k, v in d.items() if k not in ignore}
    if isinstance(d, list):
        return [_strip(v) for v in d]
    return d


@pytest.mark.parametrize("name", GRAPHS)
def test_loader_layout_is_bit_exact_with_the_recorded_reference_layout(name):
    """tests/golden/graphs/<name>.layout.json was written by the REFERENCE loader on the
    original file (tools/export_graphs.py); every index, link and parameter must be equal —
    doubles included, since json round-trips repr()."""
    ours = _strip(rcs_b200.Graph(graph_file(name)).layout())
    with open(os.path.join(GRAPH_DIR, name + ".layout.json")) as f:
        ref = json.load(f)
    assert ours["dof"] == ref["dof"] and ours["nJ"] == ref["nJ"] and ours["nBodies"] == ref["nBodies"]
    for a, b in zip(ours["bodies"], ref["bodies"]):
        for k in INT_KEYS_BODY:
            assert a[k] == b[k], (name, a["name"], k)
    for a, b in zip(ours["joints"], ref["joints"]):
        for k in INT_KEYS_JOINT:
            assert a[k] == b[k], (name, a["name"], k)
    assert ours == ref


def _reference_graph_files():
    if not os.path.isdir(REFERENCE_XML):
        return []
    return sorted(glob.glob(os.path.join(REFERENCE_XML, "*", "g*.xml")) +
                  [os.path.join(REFERENCE_XML, p) for p in
                   ("WAM/WAM-arm-primitives.xml", "Husky/husky.xml", "Schunk/sdh.xml")])


@pytest.mark.parametrize("path", _reference_graph_files(),
                         ids=lambda p: os.path.relpath(p, REFERENCE_XML))
def test_loader_matches_reference_loader_on_original_files(path, refmod):
    """Every graph file of the reference's config/xml tree that the reference itself loads:
    same layout from our expat loader (groups, prev=, xi:include, default inertia from
    shapes, coupled joints, rigid_body_joints)."""
    try:
        r = refmod.RefGraph(path)
    except RuntimeError:
        pytest.skip("the reference does not load this file in this build (meshes / URDF / bvh)")
    lib = rcs_b200.lib()
    lib.Rcs_clearResourcePath()
    lib.Rcs_addResourcePath(REFERENCE_XML.encode())
    lib.Rcs_addResourcePath(os.path.dirname(path).encode())
    try:
        g = rcs_b200.Graph(path)
    except rcs_b200.RcsbError as e:
        if "mesh" in str(e).lower() or "urdf" in str(e).lower():
            pytest.skip(f"unsupported by the host loader: {e}")
        raise
    ours, ref = g.layout(), r.layout()
    assert ours["dof"] == ref["dof"] and ours["nJ"] == ref["nJ"] and ours["nBodies"] == ref["nBodies"]
    assert [b["name"] for b in ours["bodies"]] == [b["name"] for b in ref["bodies"]]
    assert [j["name"] for j in ours["joints"]] == [j["name"] for j in ref["joints"]]
    for a, b in zip(ours["joints"], ref["joints"]):
        for k in INT_KEYS_JOINT:
            assert a[k] == b[k], (a["name"], k)
    assert ours == ref


def test_graph_check_clone_and_mass():
    g = rcs_b200.Graph(graph_file("humanoid"))
    ok, n_err, _ = g.check()
    assert ok and n_err == 0
    c = g.clone()
    assert c.layout() == g.layout()
    assert abs(g.mass() - 54.08) < 1e-9          # M[0][0] = total mass
    assert g.body_index("RightHandTip") == [b["name"] for b in g.layout()["bodies"]].index("RightHandTip")
    with pytest.raises(KeyError):
        g.body_index("no such body")


def test_depth_first_order_and_joint_indexing_rules():
    """RCSGRAPH_TRAVERSE_BODIES order = DFS; jointIndex follows body order then in-body order;
    jacobiIndex = running count of unconstrained joints (Rcs_graph.c:2932-2997)."""
    for name in GRAPHS:
        lay = rcs_b200.Graph(graph_file(name)).layout()
        for i, b in enumerate(lay["bodies"]):
            assert b["parent"] < i
        running, jidx = 0, 0
        for b in lay["bodies"]:
            for j in b["joints"]:
                jn = lay["joints"][j]
                assert jn["jointIndex"] == jidx == j
                jidx += 1
                if jn["constrained"]:
                    assert jn["jacobiIndex"] == -1
                else:
                    assert jn["jacobiIndex"] == running
                    running += 1
        assert jidx == lay["dof"] and running == lay["nJ"]


def test_missing_file_returns_null_like_the_reference():
    with pytest.raises(rcs_b200.RcsbError):
        rcs_b200.Graph("/nonexistent/graph.xml")


def test_xml_round_trip(tmp_path):
    g = rcs_b200.Graph(graph_file("wam_scenario"))
    out = tmp_path / "again.xml"
    g.write_xml(str(out))
    assert _strip(rcs_b200.Graph(str(out)).layout()) == _strip(g.layout())


def test_create_from_buffer_and_programmatic_rules():
    xml = b"""<Graph name="two">
      <Body name="base" mass="1" inertia="1 0 0 0 1 0 0 0 1"/>
      <Body name="slider" prev="base" mass="2" transform="0 0 0.5 0 0 90">
        <Joint name="j0" type="TransZ" range="-1 0.25 1"/>
        <Joint name="j1" type="RotX" range="90" constraint="true"/>
        <Joint name="j2" type="RotY" range="-10 20"/>
      </Body>
    </Graph>"""
    lib = rcs_b200.lib()
    h = lib.RcsGraph_createFromBuffer(xml, len(xml))
    assert h
    g = rcs_b200.Graph("", _handle=h)
    lay = g.layout()
    assert (lay["dof"], lay["nJ"], lay["nBodies"]) == (3, 2, 2)
    j0, j1, j2 = lay["joints"]
    assert (j0["type"], j0["dirIdx"], j0["q0"], j0["q_min"], j0["q_max"]) == (5, 2, 0.25, -1.0, 1.0)
    assert j1["constrained"] and j1["jacobiIndex"] == -1 and j1["q0"] == 0.0
    assert abs(j1["q_max"] - np.pi / 2) < 1e-15 and abs(j1["q_min"] + np.pi / 2) < 1e-15
    assert j2["jacobiIndex"] == 1 and abs(j2["q0"] - np.deg2rad(5.0)) < 1e-15   # 2 values: q0 = mean
    assert lay["bodies"][1]["A_BP"] is not None and lay["bodies"][0]["A_BP"] is None


def test_compute_calls_fail_loudly_without_a_device():
    """No CPU fallback: without a CUDA device the model cannot be created and the error says so."""
    lib = rcs_b200.lib()
    if lib.rcsb_device_count() > 0:
        pytest.skip("a CUDA device is present")
    g = rcs_b200.Graph(graph_file("wam_arm"))
    with pytest.raises(rcs_b200.RcsbError, match="(?i)cuda|device"):
        rcs_b200.Model(g, 0)
    # single-state reference API: N=1 through the same kernels -> error in non-abort mode
    q = lib.MatNd_create(7, 1)
    assert lib.RcsGraph_setState(g.handle, q, None) is False
    assert b"device" in lib.rcsb_last_error().lower() or b"cuda" in lib.rcsb_last_error().lower()
    lib.MatNd_destroy(q)


def test_product_package_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under rcs_b200/ may reference it."""
    for path in glob.glob(os.path.join(ROOT, "rcs_b200", "**", "*"), recursive=True):
        if os.path.isfile(path) and path.endswith((".py", ".c", ".cu", ".h", ".cuh", "Makefile")):
            text = open(path, errors="ignore").read()
            assert "oracle" not in text.replace("oracle/", "oracle/").lower() or \
                all("import" not in ln and "#include" not in ln and "dlopen" not in ln
                    for ln in text.splitlines() if "oracle" in ln.lower()), path


# ---------------------------------------------------------------------------------------------
# host helpers of the single-state API (O(dof) loops over the joint list; no device needed)
# ---------------------------------------------------------------------------------------------

class _MatNd(ctypes.Structure):
    _fields_ = [("ele", ctypes.POINTER(ctypes.c_double)), ("m", ctypes.c_uint), ("n", ctypes.c_uint),
                ("size", ctypes.c_uint), ("stackMem", ctypes.c_bool)]


class _RcsGraph(ctypes.Structure):
    # root, gBody[10] (10 x 184 bytes, tests/golden/abi_layout.json), dof, nJ, q, q_dot
    _fields_ = [("root", ctypes.c_void_p), ("gBody", ctypes.c_char * 1840), ("dof", ctypes.c_uint),
                ("nJ", ctypes.c_uint), ("q", ctypes.POINTER(_MatNd)), ("q_dot", ctypes.POINTER(_MatNd))]


def _host_lib():
    L = rcs_b200.lib()
    L.MatNd_create.restype = ctypes.POINTER(_MatNd)
    L.MatNd_create.argtypes = [ctypes.c_uint, ctypes.c_uint]
    L.RcsGraph_updateSlaveJoints.restype = ctypes.c_bool
    L.RcsGraph_updateSlaveJoints.argtypes = [ctypes.c_void_p, ctypes.POINTER(_MatNd), ctypes.POINTER(_MatNd)]
    L.RcsGraph_coupledJointMatrix.restype = ctypes.c_int
    L.RcsGraph_coupledJointMatrix.argtypes = [ctypes.c_void_p, ctypes.POINTER(_MatNd), ctypes.POINTER(_MatNd)]
    L.RcsGraph_jointLimitGradient.restype = ctypes.c_double
    L.RcsGraph_jointLimitGradient.argtypes = [ctypes.c_void_p, ctypes.POINTER(_MatNd), ctypes.c_int]
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
    c = M.contents
    return np.array([c.ele[i] for i in range(c.m * c.n)]).reshape(c.m, c.n)


@pytest.mark.parametrize("name", ["wam_scenario", "dexbot", "humanoid", "gantry"])
def test_slave_joint_update_coupling_matrix_and_limit_gradient_match_the_oracle(name):
    """RcsGraph_updateSlaveJoints (Rcs_graph.c:2756), RcsGraph_coupledJointMatrix (:2823) and
    RcsGraph_jointLimitGradient (Rcs_kinematics.c:1496) against the numpy restatement, which
    tests/test_oracle_cpu.py pins to the reference library."""
    from oracle import rcs_oracle
    from rcs_b200 import workload

    L = _host_lib()
    g = rcs_b200.Graph(graph_file(name))
    lay = g.layout()
    o = rcs_oracle.Oracle(lay)
    q, qd = workload.sample_states(lay, 3, first=99, with_velocity=True)
    rng = np.random.default_rng(5)
    for s in range(3):
        # perturb the slaves so that the update has something to do
        qs, qds = q[s].copy(), qd[s].copy()
        slaves = [j["jointIndex"] for j in lay["joints"] if j["coupledTo"] >= 0]
        qs[slaves] += 0.05 * rng.uniform(-1, 1, len(slaves))
        Q, QD = _mat(L, qs), _mat(L, qds)
        moved = L.RcsGraph_updateSlaveJoints(g.handle, Q, QD)
        want_q, want_qd = o.expand(qs[None, :], qds[None, :])
        assert moved == (len(slaves) > 0 and bool(np.any(np.abs(want_q[0] - qs) > 1e-8)))
        np.testing.assert_allclose(_arr(Q)[:, 0], want_q[0], rtol=0, atol=1e-15)
        np.testing.assert_allclose(_arr(QD)[:, 0], want_qd[0], rtol=0, atol=1e-15)

        # the remaining two read the state held by the graph
        G = ctypes.cast(g.handle, ctypes.POINTER(_RcsGraph)).contents
        for i in range(g.dof):
            G.q.contents.ele[i] = want_q[0, i]
        nJ = g.nJ
        nqr = nJ - sum(1 for j in lay["joints"] if j["coupledTo"] >= 0 and not j["constrained"])
        A, invA = _mat(L, m=nJ, n=max(nqr, 1)), _mat(L, m=max(nqr, 1), n=nJ)
        assert L.RcsGraph_coupledJointMatrix(g.handle, A, invA) == nJ - nqr
        A_ref, invA_ref = o.coupled_joint_matrix(want_q)
        assert _arr(A).shape == (nJ, nqr) and _arr(invA).shape == (nqr, nJ)
        np.testing.assert_allclose(_arr(A), A_ref[0], rtol=0, atol=1e-15)
        np.testing.assert_allclose(_arr(invA), invA_ref[0], rtol=0, atol=1e-15)
        np.testing.assert_allclose(_arr(invA) @ _arr(A), np.eye(nqr), rtol=0, atol=1e-14)

        dH = _mat(L, m=1, n=g.dof)
        cost = L.RcsGraph_jointLimitGradient(g.handle, dH, 1)            # RcsStateIK
        dH_ref, cost_ref = o.joint_limit_gradient(want_q)
        assert _arr(dH).shape == (1, nJ)
        np.testing.assert_allclose(_arr(dH)[0], dH_ref[0], rtol=0, atol=1e-15)
        assert abs(cost - cost_ref[0]) <= 1e-14 * max(1.0, abs(cost_ref[0]))
        cost_full = L.RcsGraph_jointLimitGradient(g.handle, dH, 0)        # RcsStateFull
        cols = [j["jointIndex"] for j in lay["joints"] if j["jacobiIndex"] >= 0]
        assert _arr(dH).shape == (1, g.dof) and cost_full == cost
        np.testing.assert_allclose(_arr(dH)[0, cols], dH_ref[0], rtol=0, atol=1e-15)


def test_bench_reference_arm_prints_the_contract_line(refmod):
    """`bench.py --impl reference` (the CPU reference arm the driver runs next to ours): one JSON
    line with the contract's keys, on the default workload, without touching a GPU."""
    import subprocess
    import sys

    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0"], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stderr[-2000:]
    line = json.loads(res.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["n_gpus"] == 1 and line["higher_is_better"] is True
    assert line["metric"].startswith("FK+Jacobian+M(q) states/sec") and line["unit"] == "states/s"
    assert line["value"] > 0 and line["dtype"] == "f64" and line["data"] == "synthetic"
    assert line["cpu_baseline"]["kind"] == "reference" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"] == {"value": line["value"], "unit": "states/s", "h2d_bytes_per_step": 0,
                           "d2h_bytes_per_step": 0}
    assert "workload" in line["config"] and line["vs_baseline"] is None
