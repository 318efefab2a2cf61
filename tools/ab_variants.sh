# A/B of kernel variants on the GPU box: serialized per-kernel times (VC_WORKERS=1, one chunk)
for v in "$@"; do
  echo "#### variant $v"
  VOXCORE_LIB=$PWD/voxel_ma_b200/lib/variants/libvoxcore_gpu_$v.so VC_WORKERS=1 VC_ZCHUNK=4096 python tools/quick_bench.py twist:512 2>&1 | grep -E "==|edt_pass|consistent"
done
