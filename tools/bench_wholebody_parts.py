#!/usr/bin/env python
"""Times the three calls of the whole-body evaluation (BASELINE.json configs[4]) separately:
FK of all frames, the task Jacobians of cAction.xml's active set, M + COG Jacobian."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import rcs_b200  # noqa: E402
from rcs_b200 import workload  # noqa: E402


def timed(fn, steps=5, warmup=2):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = rcs_b200.kernel_launch_count()
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps, (rcs_b200.kernel_launch_count() - l0) // steps


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
    g = rcs_b200.Graph(os.path.join(ROOT, "tests", "golden", "graphs", "humanoid.xml"))
    m = rcs_b200.Model(g, 0)
    rh, lh, heel, fl, fr = [g.body_index(x) for x in ("RightHandTip", "LeftHandTip", "Heel", "FootL", "FootR")]
    spec = [("POS", rh, -1, -1), ("POS", lh, -1, -1), ("POS", fl, heel, heel), ("OMEGA", fl, heel, heel),
            ("POS", fr, heel, heel), ("OMEGA", fr, heel, heel)]
    q = torch.from_numpy(workload.sample_states(g.layout(), n)).cuda()

    def buf(*shape):
        return torch.empty(shape, dtype=torch.float64, device="cuda")

    A, Jt, M, Jc = buf(n, g.n_bodies, 12), buf(n, len(spec), 3, g.nJ), buf(n, g.nJ, g.nJ), buf(n, 3, g.nJ)
    parts = {
        "fk all frames": lambda: m.fk(q, out={"A_BI": A}),
        "task jacobians (6 x 3 rows)": lambda: m.task_jacobians(q, spec, out=Jt),
        "M + Jcog": lambda: m.kinetic_terms(q, want=("M", "Jcog"), out={"M": M, "Jcog": Jc}),
    }
    for name, fn in parts.items():
        ms, launches = timed(fn)
        print(json.dumps({"part": name, "states": n, "ms": ms, "launches": launches,
                          "states_per_s": n / (ms * 1e-3)}))


if __name__ == "__main__":
    main()
