#!/bin/bash
# ncu --set full of ONE whole-grid launch of a kernel (VC_WORKERS=0: one chunk, one stream), source counters kept
# usage (under gpurun): bash tools/ncu_one.sh <tag> <kernel-regex> [workload:n]
TAG=$1; KRE=$2; WL=${3:-twist:512}
mkdir -p gpurun_out
VC_WORKERS=0 timeout 900 ncu --set full --clock-control none --import-source on -k "regex:${KRE}" -s 1 -c 1 -f -o gpurun_out/${TAG} \
    python tools/quick_bench.py ${WL} > gpurun_out/${TAG}.log 2>&1
echo "ncu exit $? $(ls -la gpurun_out/${TAG}.ncu-rep 2>/dev/null)"
