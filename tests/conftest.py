"""pytest configuration: the `gpu` marker, paths of the graph fixtures, shared helpers."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GRAPH_DIR = os.path.join(ROOT, "tests", "golden", "graphs")
GRAPHS = ["wam_arm", "wam_scenario", "humanoid", "dexbot", "husky", "schunk_sdh", "gantry", "polycoupled", "wrist6"]
REFERENCE_XML = "/root/reference/config/xml"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def graph_file(name: str) -> str:
    return os.path.join(GRAPH_DIR, name + ".xml")


def assert_close(ours, ref, tol=1.0e-12, what=""):
    """|a - b| <= tol * max(1, ||ref||_inf); structural zeros of ref must be exactly 0."""
    ours = np.asarray(ours)
    ref = np.asarray(ref)
    assert ours.shape == ref.shape, f"{what}: shape {ours.shape} != {ref.shape}"
    assert np.all(np.isfinite(ours)), f"{what}: non-finite values"
    scale = max(1.0, float(np.max(np.abs(ref))) if ref.size else 1.0)
    err = float(np.max(np.abs(ours - ref))) if ref.size else 0.0
    assert err <= tol * scale, f"{what}: max abs err {err:.3e} > {tol:.0e} * {scale:.3g}"
    return err / scale


def assert_structural_zeros(ours, ref, what=""):
    """Entries that are exactly 0 for every state in the reference must be exactly 0 in ours."""
    ours = np.asarray(ours)
    ref = np.asarray(ref)
    always_zero = np.all(ref == 0.0, axis=0)
    bad = np.any(ours != 0.0, axis=0) & always_zero
    assert not np.any(bad), f"{what}: {int(bad.sum())} structural zeros are non-zero"


@pytest.fixture(scope="session")
def refmod():
    from oracle import ref

    if not ref.available():
        pytest.skip("oracle/_ref/librcsref.so not built (needs /root/reference)")
    return ref


@pytest.fixture(scope="session")
def cuda_device():
    import rcs_b200

    if rcs_b200.lib().rcsb_device_count() < 1:
        pytest.skip("no CUDA device")
    return 0
