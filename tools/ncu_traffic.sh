#!/bin/bash
# DRAM traffic of the whole-grid launches of the hot kernels (bench.py's profiled leg: one launch per kernel
# per step).  bench.py --steps 2 --warmup 3 runs 5 pipelined steps (8 z chunks each) first: skip those launches.
# usage (under gpurun): bash tools/ncu_traffic.sh <tag>
TAG=${1:-r01}
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:k_pass_xy|k_cell_measures|k_pass_z|k_classify_f32" \
    -s 160 -c 12 -f -o gpurun_out/${TAG}_traffic python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_traffic.log 2>&1
echo "ncu exit $?"
ncu -i gpurun_out/${TAG}_traffic.ncu-rep --page raw --csv --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,launch__grid_size > gpurun_out/${TAG}_traffic.csv 2>&1
cat gpurun_out/${TAG}_traffic.csv | cut -c1-400
