"""Binary compatibility of include/rcsb_graph.h with the reference's structs
(src/RcsCore/Rcs_typedef.h:66-259, Rcs_HTr.h:62-66, Rcs_MatNd.h:92-99): size and offset of every
field, compiled with gcc from our header, against tests/golden/abi_layout.json (recorded from the
reference's own headers by tools/make_abi_golden.py) and, where /root/reference exists, against
those headers directly."""
import json
import os
import sys

import pytest

from conftest import ROOT

sys.path.insert(0, os.path.join(ROOT, "tools"))
import make_abi_golden  # noqa: E402


def test_struct_layout_matches_the_recorded_reference_layout():
    with open(os.path.join(ROOT, "tests", "golden", "abi_layout.json")) as f:
        golden = {k: list(v) for k, v in json.load(f).items()}
    ours = make_abi_golden.our_layout()
    assert set(ours) == set(golden)
    bad = {k: (ours[k], golden[k]) for k in golden if ours[k] != golden[k]}
    assert not bad, f"(ours, reference) size/offset mismatches: {bad}"
    # every field of the reference's structs is covered
    assert len(golden) == sum(len(v) + 1 for v in make_abi_golden.FIELDS.values())


@pytest.mark.skipif(not os.path.isdir("/root/reference/src/RcsCore"), reason="needs /root/reference")
def test_recorded_layout_is_the_reference_headers_layout():
    with open(os.path.join(ROOT, "tests", "golden", "abi_layout.json")) as f:
        golden = {k: list(v) for k, v in json.load(f).items()}
    assert make_abi_golden.reference_layout() == golden
