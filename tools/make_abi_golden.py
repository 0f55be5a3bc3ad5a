#!/usr/bin/env python
"""Records sizeof / offsetof of the reference's graph structs (src/RcsCore/Rcs_typedef.h,
Rcs_HTr.h, Rcs_MatNd.h), compiled here with gcc from the headers where they lie under
/root/reference, into tests/golden/abi_layout.json. tests/test_abi_cpu.py compiles the same
probe against include/rcsb_graph.h and compares: the structs of rcs_b200 are binary compatible
with libRcsCore's, so a graph built by either library can be handed to the other."""
import json
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

FIELDS = {
    "HTr": ["org", "rot"],
    "MatNd": ["ele", "m", "n", "size", "stackMem"],
    "RcsJoint": ["name", "q0", "q_init", "q_min", "q_max", "weightJL", "weightCA", "weightMetric",
                 "constrained", "type", "dirIdx", "jointIndex", "jacobiIndex", "A_JP", "A_JI", "maxTorque",
                 "speedLimit", "accLimit", "decLimit", "ctrlType", "coupledJointName", "couplingFactors",
                 "prev", "next", "coupledTo"],
    "RcsShape": ["type", "A_CB", "extents", "scale", "resizeable", "computeType", "meshFile", "textureFile",
                 "color", "material", "userData"],
    "RcsBody": ["m", "rigid_body_joints", "physicsSim", "x_dot", "omega", "confidence", "name", "xmlName",
                "suffix", "A_BP", "A_BI", "Inertia", "parent", "firstChild", "lastChild", "next", "prev",
                "shape", "jnt", "extraInfo"],
    "RcsGraph": ["root", "gBody", "dof", "nJ", "q", "q_dot", "xmlFile", "sensor", "userData"],
}


def probe_source(includes):
    lines = ["#include <stdio.h>", "#include <stddef.h>"] + [f"#include <{h}>" for h in includes]
    lines.append("int main(void) {")
    for st, fields in FIELDS.items():
        lines.append(f'  printf("{st} %zu\\n", sizeof({st}));')
        for f in fields:
            lines.append(f'  printf("{st}.{f} %zu %zu\\n", offsetof({st}, {f}), sizeof((({st}*)0)->{f}));')
    lines.append("  return 0; }")
    return "\n".join(lines)


def layout(includes, include_dirs):
    with tempfile.TemporaryDirectory() as tmp:
        src, exe = os.path.join(tmp, "probe.c"), os.path.join(tmp, "probe")
        with open(src, "w") as f:
            f.write(probe_source(includes))
        cmd = ["gcc", "-std=gnu99", "-o", exe, src] + [f"-I{d}" for d in include_dirs]
        subprocess.run(cmd, check=True)
        out = subprocess.run([exe], check=True, capture_output=True, text=True).stdout
    res = {}
    for ln in out.splitlines():
        parts = ln.split()
        res[parts[0]] = [int(x) for x in parts[1:]]
    return res


def reference_layout(ref="/root/reference"):
    return layout(["Rcs_typedef.h"], [os.path.join(ref, "src", "RcsCore")])


def our_layout():
    return layout(["rcsb_graph.h"], [os.path.join(ROOT, "include")])


if __name__ == "__main__":
    lay = reference_layout()
    path = os.path.join(ROOT, "tests", "golden", "abi_layout.json")
    with open(path, "w") as f:
        json.dump(lay, f, indent=1, sort_keys=True)
    ours = our_layout()
    bad = {k: (v, ours.get(k)) for k, v in lay.items() if ours.get(k) != v}
    print(f"wrote {path}: {len(lay)} entries; mismatches against include/rcsb_graph.h: {bad}")
    sys.exit(1 if bad else 0)
