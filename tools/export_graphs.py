#!/usr/bin/env python
"""Exports the benchmark graphs as self-contained XML fixtures under tests/golden/graphs/.

/root/reference does not exist on the GPU box, so the graphs the tests and bench.py use are
re-serialised here (in the CPU container, where the reference tree is mounted) by
rcs_b200's own loader + writer: RcsGraph_create(reference file) -> RcsGraph_writeXML().
The exported files carry explicit mass / cogVector / inertia instead of shapes and 12-number
transforms instead of groups / includes; both rcs_b200 and the reference library
(oracle/_ref) load them to the same integer layout and parameters as the original files,
which this script verifies before writing.

    python tools/export_graphs.py
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import rcs_b200  # noqa: E402
from oracle import ref  # noqa: E402

REF_XML = "/root/reference/config/xml"
GRAPHS = {
    "wam_arm": "WAM/WAM-arm-primitives.xml",
    "wam_scenario": "WAM/gScenario.xml",
    "humanoid": "GenericHumanoid/gScenario.xml",
    "dexbot": "DexBot/gScenario.xml",
    "husky": "Husky/husky.xml",
    "schunk_sdh": "Schunk/sdh.xml",
}
# fixtures written by hand for this repository (not re-serialised): only their layout is recorded,
# by the reference loader
LOCAL = ["gantry", "polycoupled", "wrist6"]
# layout keys that legitimately differ after export (shapes are dropped)
IGNORE = {"nShapes"}


def strip(d):
    if isinstance(d, dict):
        return {k: strip(v) for k, v in d.items() if k not in IGNORE}
    if isinstance(d, list):
        return [strip(v) for v in d]
    return d


def main():
    out_dir = os.path.join(ROOT, "tests", "golden", "graphs")
    os.makedirs(out_dir, exist_ok=True)
    for name, rel in GRAPHS.items():
        src = os.path.join(REF_XML, rel)
        g = rcs_b200.Graph(src)
        dst = os.path.join(out_dir, name + ".xml")
        g.write_xml(dst)
        orig = strip(ref.RefGraph(src).layout())
        again_ours = strip(rcs_b200.Graph(dst).layout())
        again_ref = strip(ref.RefGraph(dst).layout())
        ok = orig == again_ours == again_ref
        print(f"{name:14s} dof={g.dof:3d} nJ={g.nJ:3d} nb={g.n_bodies:3d} "
              f"round-trip identical to reference load of original: {ok}")
        if not ok:
            for k in ("bodies", "joints"):
                for a, b in zip(orig[k], again_ref[k]):
                    if a != b:
                        diff = {x: (a[x], b[x]) for x in a if a[x] != b[x]}
                        print("   ", k, a["name"], diff)
            raise SystemExit(1)
        with open(os.path.join(out_dir, name + ".layout.json"), "w") as f:
            json.dump(orig, f, indent=0, sort_keys=True)
    for name in LOCAL:
        src = os.path.join(out_dir, name + ".xml")
        lay_ref = strip(ref.RefGraph(src).layout())
        ok = lay_ref == strip(rcs_b200.Graph(src).layout())
        print(f"{name:14s} dof={lay_ref['dof']:3d} nJ={lay_ref['nJ']:3d} nb={lay_ref['nBodies']:3d} "
              f"own fixture, both loaders agree: {ok}")
        if not ok:
            raise SystemExit(1)
        with open(os.path.join(out_dir, name + ".layout.json"), "w") as f:
            json.dump(lay_ref, f, indent=0, sort_keys=True)


if __name__ == "__main__":
    main()
