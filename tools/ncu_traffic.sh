#!/bin/bash
# DRAM traffic of the whole-grid launches of the hot kernels (bench.py's profiled leg: one launch per kernel
# per step).  bench.py --steps 2 --warmup 3 runs 5 pipelined steps first (33 matching launches each): skip those.
# The report itself stays on the box (tens of MB); only the per-launch CSV comes back.
# usage (under gpurun): bash tools/ncu_traffic.sh <tag>
TAG=${1:-r01}
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none -k "regex:k_pass_xy|k_cell_measures|k_pass_z|k_classify_f32" \
    -s 165 -c 10 -f -o /tmp/${TAG}_traffic python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_traffic.log 2>&1
echo "ncu exit $?"
ncu -i /tmp/${TAG}_traffic.ncu-rep --page raw --csv --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,launch__grid_size > gpurun_out/${TAG}_traffic.csv 2>&1
cut -c1-300 gpurun_out/${TAG}_traffic.csv | tail -12
