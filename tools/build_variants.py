"""Development aid: build kernel variants (extra -D flags) side by side for one A/B run on the GPU box."""
import sys
sys.path.insert(0, ".")
from voxel_ma_b200 import build as vb
VARIANTS = {
    "base": [],
    "ringy16": ["-DSR_RY=16"],
    "wi_nospill": ["-DWHATIF_NOSPILL"],
    "wi_norefill": ["-DWHATIF_NOREFILL"],
    "wi_nospill_norefill": ["-DWHATIF_NOSPILL", "-DWHATIF_NOREFILL"],
    "wi_nodiv": ["-DWHATIF_NODIV"],
    "wi_noback": ["-DWHATIF_NOBACK"],
    "wi_all": ["-DWHATIF_NOSPILL", "-DWHATIF_NOREFILL", "-DWHATIF_NODIV"],
    "minb12": ["-DXY_MINB=12"],
}
names = sys.argv[1:] or list(VARIANTS)
for n in names:
    print(n, vb.build_variant(n, VARIANTS[n]))
