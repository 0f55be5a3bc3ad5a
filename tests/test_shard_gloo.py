"""Multi-rank host logic on CPU: world_size 2 (and 3) over gloo. Each rank generates its own
slice of the global counter-based state sequence, evaluates it (here with the numpy oracle —
no GPU in this container), and the ranks exchange summary statistics with one all-gather, as
bench.py does over NCCL.
"""
import json
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import GRAPH_DIR
from rcs_b200 import shard, workload


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _layout(name):
    with open(os.path.join(GRAPH_DIR, name + ".layout.json")) as f:
        return json.load(f)


def test_slices_partition_the_batch():
    for n in (0, 1, 7, 1 << 20, (1 << 20) + 5):
        for world in (1, 2, 3, 8):
            pos = 0
            for r in range(world):
                first, count = shard.slice_of(r, world, n)
                assert first == pos and count >= 0
                pos += count
            assert pos == n
            sizes = [shard.slice_of(r, world, n)[1] for r in range(world)]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard.slice_of(2, 2, 10)


def test_rank_slices_equal_the_global_sequence():
    lay = _layout("wam_arm")
    n = 1000
    q_all, qd_all = workload.sample_states(lay, n, with_velocity=True)
    for world in (2, 3):
        parts = [workload.sample_states(lay, c, first=f, with_velocity=True)
                 for f, c in (shard.slice_of(r, world, n) for r in range(world))]
        assert np.array_equal(np.concatenate([p[0] for p in parts]), q_all)
        assert np.array_equal(np.concatenate([p[1] for p in parts]), qd_all)


def _worker(rank, world, port, n_global, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import rcs_oracle

        lay = _layout("wam_arm")
        o = rcs_oracle.Oracle(lay)
        tip = [b["name"] for b in lay["bodies"]].index("BallTip")
        first, count = shard.slice_of(rank, world, n_global)
        q = workload.sample_states(lay, count, first=first)
        height = o.fk(q)["A_BI"][:, tip, 2]          # per-state result: effector height
        gathered = shard.all_gather_stats(shard.local_stats(torch.from_numpy(height)))
        assert gathered.shape == (world, 5)
        merged = shard.merge_stats(gathered)
        # timing reduction used by bench.py: max over ranks
        t = torch.tensor([float(rank + 1)], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        # configs[4]: all-gather of compact per-state results (here the effector frame)
        frames = heights = None
        if n_global % world == 0:                            # the collective needs equal slices
            frames = shard.all_gather_results(
                torch.from_numpy(np.ascontiguousarray(o.fk(q)["A_BI"][:, tip]))).numpy().copy()
            work_out, work = shard.all_gather_results(torch.from_numpy(height.copy()), async_op=True)
            work.wait()
            heights = work_out.numpy().copy()
        ret[rank] = (merged, float(t.item()), gathered.numpy().copy(), frames, heights)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n_global", [(2, 1000), (3, 1001), (2, 1001)])
def test_all_gather_of_summary_statistics(world, n_global):
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), n_global, ret), nprocs=world, join=True)
    assert len(ret) == world
    from oracle import rcs_oracle

    lay = _layout("wam_arm")
    tip = [b["name"] for b in lay["bodies"]].index("BallTip")
    height = rcs_oracle.Oracle(lay).fk(workload.sample_states(lay, n_global))["A_BI"][:, tip, 2]
    frames_all = rcs_oracle.Oracle(lay).fk(workload.sample_states(lay, n_global))["A_BI"][:, tip]
    for r in range(world):
        merged, tmax, gathered, frames, heights = ret[r]
        if n_global % world == 0:                            # equal slices: rows in global order
            assert np.array_equal(frames, frames_all) and np.array_equal(heights, height)
        assert tmax == float(world)
        assert np.array_equal(gathered, ret[0][2])          # every rank holds the same table
        assert merged["count"] == n_global
        assert abs(merged["mean"] - height.mean()) < 1e-12
        assert abs(merged["std"] - height.std()) < 1e-10
        assert merged["min"] == height.min() and merged["max"] == height.max()


def test_stats_without_process_group_and_empty_slices():
    s = shard.local_stats(np.array([]))
    assert s[0] == 0 and np.isinf(s[3])
    g = shard.all_gather_stats(shard.local_stats(np.array([1.0, 3.0])))
    m = shard.merge_stats(g)
    assert (m["count"], m["mean"], m["min"], m["max"]) == (2.0, 2.0, 1.0, 3.0)
    both = np.stack([shard.local_stats(np.array([])), shard.local_stats(np.array([2.0]))])
    assert shard.merge_stats(both)["mean"] == 2.0
