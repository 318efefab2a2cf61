timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for cfg in "1 512" "4 128" "8 64" "8 32" "4 64"; do
  set -- $cfg
  VC_WORKERS=$1 VC_ZCHUNK=$2 python tools/quick_bench.py twist:512 2>&1 | grep -E "==|edt|measures"
done
VC_WORKERS=8 VC_ZCHUNK=64 python tools/quick_bench.py sphere:512 torus:256 2>&1 | grep "=="
