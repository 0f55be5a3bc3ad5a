#!/usr/bin/env python
"""Measures write-only, read-only and copy bandwidth of the device with torch (CUDA events, best
of 10 over 2 GiB buffers): context for kernels whose traffic is almost entirely stores (fkKernel
writes 96 % of its bytes) next to the copy figure of MEASURED_PEAKS.json."""
import json

import torch


def best(fn, n=10):
    t = []
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        t.append(e0.elapsed_time(e1))
    return min(t)


def main():
    n = 1 << 28                                   # doubles: 2 GiB
    a = torch.empty(n, dtype=torch.float64, device="cuda")
    b = torch.empty(n, dtype=torch.float64, device="cuda")
    a.fill_(1.0)
    torch.cuda.synchronize()
    w = best(lambda: a.fill_(2.0))
    c = best(lambda: b.copy_(a))
    r = best(lambda: a.sum())
    gb = n * 8 / 1e9
    print(json.dumps({"write_only_GBps": gb / (w * 1e-3), "copy_GBps": 2 * gb / (c * 1e-3),
                      "read_only_GBps": gb / (r * 1e-3), "buffer_GiB": 2}))


if __name__ == "__main__":
    main()
