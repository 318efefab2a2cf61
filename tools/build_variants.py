"""Development aid: build kernel variants (extra -D flags) side by side for one A/B run on the GPU box."""
import sys
sys.path.insert(0, ".")
from voxel_ma_b200 import build as vb
VARIANTS = {
    "base": [],
    "d2": ["-DED_DEPTH=2"],
    "d3": ["-DED_DEPTH=3"],
    "d6": ["-DED_DEPTH=6"],
    "d4_pf0": ["-DED_PFD=0"],
    "d4_pf24": ["-DED_PFD=24"],
    "d6_pf0": ["-DED_DEPTH=6", "-DED_PFD=0"],
    "wi_nospill": ["-DWHATIF_NOSPILL"],
}
names = sys.argv[1:] or list(VARIANTS)
for n in names:
    print(n, vb.build_variant(n, VARIANTS[n]))
