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
