"""GPU: rcsb_summary_stats (the payload of the multi-GPU all-gather) against numpy."""
import numpy as np
import pytest

import rcs_b200
from rcs_b200 import shard

pytestmark = pytest.mark.gpu


def test_summary_stats_contiguous_and_strided(cuda_device):
    import torch

    rng = np.random.default_rng(0)
    x = rng.normal(size=(100003, 7))
    t = torch.from_numpy(x).cuda(cuda_device)
    for view, ref in ((t[:, 3], x[:, 3]), (t[::17, 0], x[::17, 0]), (t.reshape(-1)[:5], x.reshape(-1)[:5])):
        s = rcs_b200.summary_stats(view).cpu().numpy()
        assert s[0] == ref.size
        np.testing.assert_allclose(s[1], ref.sum(), rtol=1e-12, atol=1e-9)
        np.testing.assert_allclose(s[2], (ref * ref).sum(), rtol=1e-12)
        assert s[3] == ref.min() and s[4] == ref.max()
    merged = shard.merge_stats(shard.all_gather_stats(shard.local_stats(t[:, 1])))
    assert abs(merged["mean"] - x[:, 1].mean()) < 1e-12 and merged["count"] == x.shape[0]


def test_bench_prints_the_contract_line(cuda_device):
    """`python bench.py` on one GPU: one JSON line with the contract's keys (metric of
    BASELINE.json, roofline, e2e through host pointers, launch count, clocks)."""
    import json
    import os
    import subprocess
    import sys

    from conftest import ROOT

    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "3", "--warmup", "3",
                          "--no-cpu-baseline", "--settle", "0.2", "--hold", "0.2"],
                         capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stderr[-2000:]
    line = json.loads(res.stdout.strip().splitlines()[-1])
    assert line["metric"].startswith("FK+Jacobian+M(q) states/sec") and line["unit"] == "states/s"
    assert line["n_gpus"] == 1 and line["steps"] == 3 and line["warmup"] == 3 and line["dtype"] == "f64"
    assert line["value"] > 1e8 and line["scaling"] == "weak" and line["vs_baseline"] is None
    roof = line["roofline"]
    assert roof["bound"] == "hbm" and 0.0 < roof["frac"] < 1.0
    assert abs(roof["frac"] - roof["achieved"] / roof["peak"]) < 1e-12
    assert roof["algorithmic_per_launch"] == 1936 * (1 << 20)
    e2e = line["e2e"]
    assert 0 < e2e["value"] < line["value"] and e2e["results_match_device_path"] is True
    assert e2e["h2d_bytes_per_step"] == 56 * (1 << 20) and e2e["d2h_bytes_per_step"] == 1880 * (1 << 20)
    assert line["gpu_launches"] == 3                      # one fused launch per timed step
    assert set(line["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}
