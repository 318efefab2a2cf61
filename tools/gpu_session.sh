#!/bin/bash
# One GPU-box session: parity tests, bench line, ncu launch list, ncu full capture of the hot kernels.
# usage (under gpurun): bash tools/gpu_session.sh <tag> [kernel-regex]
TAG=${1:-r01}
KRE=${2:-k_pass_xy|k_cell_measures|k_detect_sites|k_pass_z}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
tail -3 gpurun_out/${TAG}_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench exit $?"
cat gpurun_out/${TAG}_bench.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_bench.log 2>&1; echo "ncu list exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:${KRE}" -s 8 -c 8 -f -o gpurun_out/${TAG}_prof \
    python tools/quick_bench.py twist:512 > gpurun_out/${TAG}_ncu_full.log 2>&1; echo "ncu full exit $?"
ls -la gpurun_out
