#!/bin/bash
# Runs on the GPU box (gpurun): one `ncu --set full` capture per hot kernel of the four workloads, each
# only after the same bench command has exited 0 without ncu; reports land in gpurun_out/<dir>/ and are
# turned into profiles/<tag>_*.txt + profiles/kernel_counters.json by tools/profile_extract.py here.
#   tools/capture_profiles.sh gpurun_out/r2x
set -u
OUT=${1:-gpurun_out/prof}
mkdir -p "$OUT"
B="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-secondary"
cap() {   # name workload kernel-regex skip
  local name=$1 wl=$2 regex=$3 skip=$4
  if $B --workload "$wl" > "$OUT/$name.plain.json" 2> "$OUT/$name.plain.err"; then
    ncu --set full --clock-control none --import-source on -k "regex:$regex" -s "$skip" -c 1 \
        -o "$OUT/$name" -f $B --workload "$wl" > "$OUT/$name.ncu.log" 2>&1
    echo "$name: plain rc 0, ncu rc $?"
  else
    echo "$name: plain run FAILED"
  fi
}
cap fkm fkm "fk(Chain)?Kernel" 2
cap fk fk "fk(Chain)?Kernel" 2
cap ik ik ikChainKernel 2
cap ktwalk kinetic ktWalkKernel 2
cap ktasm kinetic ktAssembleKernel 2
cap wbwalk wholebody ktWalkKernel 3
cap wbasm wholebody ktAssembleKernel 3
ls -la "$OUT"/*.ncu-rep
