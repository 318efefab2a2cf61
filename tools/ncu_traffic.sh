#!/bin/bash
# DRAM traffic of ONE whole-grid / whole-slab launch of each hot kernel (ncu --set full --clock-control none).
# VC_WORKERS=0 runs every stage as one launch on one stream, so a launch = the whole grid (or slab).
# The reports stay on the box; the per-launch CSVs come back and tools/ncu_traffic_json.py turns them into profiles/traffic.json.
# usage (under gpurun): bash tools/ncu_traffic.sh <tag>
TAG=${1:-r02}
KRE='regex:k_classify_f32|k_pass_x|k_pass_y|k_pass_z|k_cell_measures'
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,launch__grid_size,smsp__inst_executed.sum
mkdir -p gpurun_out
cap() { # name skip script args...
  local name=$1 skip=$2; shift 2
  VC_WORKERS=0 timeout 900 ncu --set full --clock-control none -k "$KRE" -s $skip -c 5 -f -o /tmp/${TAG}_${name} python "$@" > gpurun_out/${TAG}_traffic_${name}.log 2>&1
  echo "ncu $name exit $?"
  ncu -i /tmp/${TAG}_${name}.ncu-rep --page raw --csv --metrics $M > gpurun_out/${TAG}_traffic_${name}.csv 2>&1
}
cap twist512 5 tools/quick_bench.py twist:512
cap assembly1024 5 tools/quick_bench.py assembly:1024
cap assembly1024_slab8 13 tools/slab_bench.py assembly:1024 8 2
cut -c1-200 gpurun_out/${TAG}_traffic_twist512.csv | tail -6
