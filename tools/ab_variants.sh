# A/B of kernel variants on the GPU box: serialized per-kernel times (VC_WORKERS=0: one chunk, one stream)
# usage (under gpurun): bash tools/ab_variants.sh [workload:n] variant...   ("base" = the in-tree library)
WL=twist:512
case "$1" in *:*) WL=$1; shift;; esac
for v in "$@"; do
  echo "#### variant $v"
  LIB=$PWD/voxel_ma_b200/lib/variants/libvoxcore_gpu_$v.so
  [ "$v" = base ] && LIB=$PWD/voxel_ma_b200/lib/libvoxcore_gpu.so
  VOXCORE_LIB=$LIB VC_WORKERS=0 python tools/quick_bench.py $WL 2>&1 | grep -E "==|edt_pass|cell_meas|consistent"
done
