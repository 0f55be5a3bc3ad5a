#!/usr/bin/env python
"""Writes tests/golden/urdf/<name>.layout.json: the layout the REFERENCE's URDF loader
(RcsGraph_fromURDFFile, and the <URDF> node of RcsGraph_create) builds from the hand-written fixtures
in tests/golden/urdf/. Needs oracle/_ref/librcsref.so (built from /root/reference).

    python tools/make_urdf_golden.py
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref  # noqa: E402

D = os.path.join(ROOT, "tests", "golden", "urdf")


def main():
    ref.lib().ref_add_resource_path(D.encode())
    for name in ("urdfbot.urdf", "urdfarm.urdf", "urdf_in_graph.xml"):
        lay = ref.RefGraph(os.path.join(D, name)).layout()
        out = os.path.join(D, os.path.splitext(name)[0] + ".layout.json")
        with open(out, "w") as f:
            json.dump(lay, f, indent=0)
        print(out, lay["nBodies"], "bodies", lay["dof"], "dof", lay["nJ"], "columns")


if __name__ == "__main__":
    main()
