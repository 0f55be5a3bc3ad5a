#!/usr/bin/env python
"""Turns an ncu report (.ncu-rep, brought back in gpurun_out/) into the tracked evidence under
profiles/: a text summary per kernel and an entry of profiles/kernel_counters.json that bench.py
reads for roofline.traffic / FP64 flop counts. Needs `ncu` on PATH (no GPU).

    python tools/profile_extract.py <workload> <report.ncu-rep> <units per launch> [tag]
        workload: fk | kinetic | ik (kinetic: pass the walk report, then the assembly report
        with workload "kinetic+" to add to the same entry)
"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
from ncu_summary import KEYS  # noqa: E402


def raw_page(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    return rows[0], rows[1], rows[2:]


def fp64_thread_ops(path):
    """(dfma, dmul, dadd, all thread instructions, warp instructions) from the SASS page."""
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[1]
    isrc, ith, iw = hdr.index("Source"), hdr.index("Predicated-On Thread Instructions Executed"), \
        hdr.index("Instructions Executed")
    c = {"DFMA": 0, "DMUL": 0, "DADD": 0}
    total_t = total_w = 0
    for r in rows[2:]:
        if len(r) <= max(isrc, ith, iw):
            continue
        t = r[isrc].split()
        if not t:
            continue
        op = (t[1] if t[0].startswith("@") and len(t) > 1 else t[0]).split(".")[0]
        n = int(r[ith] or 0)
        total_t += n
        total_w += int(r[iw] or 0)
        if op in c:
            c[op] += n
    return c["DFMA"], c["DMUL"], c["DADD"], total_t, total_w


def main():
    workload, path, units = sys.argv[1], sys.argv[2], int(sys.argv[3])
    tag = sys.argv[4] if len(sys.argv) > 4 else "r01"
    add = workload.endswith("+")
    workload = workload.rstrip("+")
    hdr, unit_row, rows = raw_page(path)
    r = rows[0]
    get = lambda k: r[hdr.index(k)] if k in hdr else None   # noqa: E731
    unit = lambda k: unit_row[hdr.index(k)] if k in hdr else ""   # noqa: E731

    def in_bytes(k):
        v, u = float(get(k)), unit(k).lower()
        return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}[u]

    kernel = get("Kernel Name")
    dram = in_bytes("dram__bytes_read.sum") + in_bytes("dram__bytes_write.sum")
    dur_us = float(get("gpu__time_duration.sum")) * {"us": 1, "ms": 1e3, "ns": 1e-3}[unit("gpu__time_duration.sum")]
    dfma, dmul, dadd, tinst, winst = fp64_thread_ops(path)
    flops = 2 * dfma + dmul + dadd

    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    short = kernel.split("(")[0].replace("void ", "").replace("rcsb::", "").replace("<", "_").replace(
        ">", "").replace(", ", "_").replace(" ", "")
    txt = os.path.join(ROOT, "profiles", f"{tag}_{workload}_{short}.txt")
    with open(txt, "w") as f:
        f.write(f"# ncu --set full --clock-control none, one launch of {kernel}\n")
        f.write(f"# report: {os.path.basename(path)}; units per launch: {units}\n")
        f.write("# per-launch time under ncu is cold-cache and serialised; bench.py times with CUDA events\n")
        for k in KEYS:
            if k in hdr:
                f.write(f"{k:88s} {get(k):>18s} {unit(k)}\n")
        f.write(f"{'dram bytes per launch (read + write)':88s} {dram:18.0f} byte\n")
        f.write(f"{'dram bytes per unit':88s} {dram / units:18.1f} byte\n")
        f.write(f"{'FP64 thread instructions: DFMA / DMUL / DADD':88s} {dfma} / {dmul} / {dadd}\n")
        f.write(f"{'FP64 flops per unit (2 DFMA + DMUL + DADD)':88s} {flops / units:18.1f}\n")
        f.write(f"{'warp instructions per unit':88s} {winst / units:18.1f}\n")
        f.write(f"{'achieved FP64 TFLOP/s under ncu':88s} {flops / (dur_us * 1e-6) / 1e12:18.2f}\n")
        f.write(f"{'achieved DRAM GB/s under ncu':88s} {dram / (dur_us * 1e-6) / 1e9:18.1f}\n")
    cj = os.path.join(ROOT, "profiles", "kernel_counters.json")
    data = json.load(open(cj)) if os.path.exists(cj) else {}
    e = data.get(workload, {}) if add else {}
    e["dram_bytes_per_launch"] = e.get("dram_bytes_per_launch", 0) + dram
    e["fp64_flops_per_unit"] = e.get("fp64_flops_per_unit", 0) + flops / units
    e["units_per_launch"] = units
    e.setdefault("kernels", []).append({"kernel": kernel, "ncu_us": dur_us, "dram_bytes": dram,
                                        "fp64_flops_per_unit": flops / units, "report": os.path.basename(path),
                                        "summary": os.path.relpath(txt, ROOT)})
    data[workload] = e
    json.dump(data, open(cj, "w"), indent=1, sort_keys=True)
    print(open(txt).read())


if __name__ == "__main__":
    main()
