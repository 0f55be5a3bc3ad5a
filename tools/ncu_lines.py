#!/usr/bin/env python
"""Stall samples and executed instructions of one profiled kernel, aggregated by CUDA source line.

ncu's CSV export of the source page carries per-SASS-instruction counters but not the line each
instruction belongs to; nvdisasm -g of the same cubin does. Both list the kernel's instructions in
address order, so the i-th instruction of one is the i-th of the other.

    python tools/ncu_lines.py <report.ncu-rep> <file.cubin> <mangled-name substring> [top N]
"""
import collections
import csv
import re
import subprocess
import sys


def sass_rows(report):
    out = subprocess.run(["ncu", "-i", report, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    res = []
    for r in rows[2:]:
        if len(r) < len(hdr) or not r[ix["Source"]].strip():
            continue
        res.append((r[ix["Source"]].strip(), int(r[ix["# Samples"]] or 0), int(r[ix["Instructions Executed"]] or 0)))
    return res


def line_table(cubin, name):
    out = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
    lines, cur, active = [], None, False
    for ln in out.splitlines():
        m = re.match(r"\s*\.text\.(\S+):", ln)
        if m:
            active = name in m.group(1)
            continue
        if not active:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", ln):
            lines.append(cur)
    return lines


def main():
    report, cubin, name = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    sass, lines = sass_rows(report), line_table(cubin, name)
    print(f"# {len(sass)} profiled instructions, {len(lines)} in the cubin")
    n = min(len(sass), len(lines))
    samples, instr = collections.Counter(), collections.Counter()
    for (src, s, i), ln in zip(sass[:n], lines[:n]):
        samples[ln] += s
        instr[ln] += i
    ts, ti = sum(samples.values()) or 1, sum(instr.values()) or 1
    print("# line: samples %, warp instructions %")
    for ln, s in samples.most_common(top):
        print(f"{ln[0] if ln else '?':22s} {ln[1] if ln else 0:5d}  {100.0 * s / ts:5.1f}  {100.0 * instr[ln] / ti:5.1f}")


if __name__ == "__main__":
    main()
