"""Development aid: build kernel variants (extra -D flags) side by side for one A/B run on the GPU box."""
import sys
sys.path.insert(0, ".")
from voxel_ma_b200 import build as vb
VARIANTS = {
    "base": [],
    "nostkpf": ["-DVC_PF_STK=0"],
    "stk12": ["-DVC_PF_STK=12"],
    "pf8": ["-DVC_PF=8"],
    "pf2": ["-DVC_PF=2"],
    "nol2": ["-DVC_PF_L2=0"],
    "l2_96": ["-DVC_PF_L2=96"],
    "occ40": ["-DXY_MINB_T=10", "-DXY_MINB_D=5"],
    "occ48": ["-DXY_MINB_T=12", "-DXY_MINB_D=6"],
    "occ24": ["-DXY_MINB_T=6", "-DXY_MINB_D=3"],
    "ring16": ["-DSR_R=16", "-DXY_MINB_T=6"],
    "ring4": ["-DSR_R=4"],
}
names = sys.argv[1:] or list(VARIANTS)
for n in names:
    print(n, vb.build_variant(n, VARIANTS[n]))
