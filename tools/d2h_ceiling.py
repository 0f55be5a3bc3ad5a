#!/usr/bin/env python
"""Host ceiling of the end-to-end path: concurrent device-to-host copies into pinned memory from
every GPU of the box, one process per GPU (torchrun), nothing else running. Plain cudaMemcpyAsync
(torch's copy_ of a pinned tensor), 1 GiB per copy.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29511 tools/d2h_ceiling.py [--bind]

--bind: rcsb_bind_host_to_device(local rank) before the pinned buffer is allocated (thread
affinity and preferred memory node of the GPU's NUMA node)."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--bind", action="store_true")
    ap.add_argument("--gib", type=float, default=1.0)
    ap.add_argument("--reps", type=int, default=8)
    args = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    info = ""
    if args.bind:
        import rcs_b200

        info = rcs_b200.bind_host_to_device(local)
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("gloo")
    n = int(args.gib * (1 << 30)) // 8
    dev = torch.empty(n, dtype=torch.float64, device=f"cuda:{local}").normal_()
    host = torch.empty(n, dtype=torch.float64, pin_memory=True)
    host.copy_(dev)
    torch.cuda.synchronize()
    res = {}
    for name, (dst, src) in (("d2h", (host, dev)), ("h2d", (dev, host))):
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(args.reps):
            dst.copy_(src, non_blocking=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        res[name] = args.reps * n * 8 / dt / 1e9
    # both directions at once, each on its own stream with its own buffers (PCIe is full duplex; what
    # the host side sustains is the question)
    dev2 = torch.empty(n, dtype=torch.float64, device=f"cuda:{local}")
    host2 = torch.empty(n, dtype=torch.float64, pin_memory=True).zero_()
    s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(args.reps):
        with torch.cuda.stream(s_out):
            host.copy_(dev, non_blocking=True)
        with torch.cuda.stream(s_in):
            dev2.copy_(host2, non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    res["duplex_each"] = args.reps * n * 8 / dt / 1e9
    gathered = [None] * world
    if world > 1:
        dist.all_gather_object(gathered, (local, res, info))
    else:
        gathered = [(local, res, info)]
    if rank == 0:
        out = {"gpus": world, "bind": args.bind,
               "d2h_gbs_per_gpu": [round(g[1]["d2h"], 2) for g in gathered],
               "h2d_gbs_per_gpu": [round(g[1]["h2d"], 2) for g in gathered],
               "d2h_gbs_total": round(sum(g[1]["d2h"] for g in gathered), 2),
               "h2d_gbs_total": round(sum(g[1]["h2d"] for g in gathered), 2),
               "duplex_gbs_each_direction_per_gpu": [round(g[1]["duplex_each"], 2) for g in gathered],
               "binding": [g[2] for g in gathered]}
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
