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
    "mr2": ["-DMR_SUB=2"],
    "mr8": ["-DMR_SUB=8"],
    "mr16": ["-DMR_SUB=16"],
    "mr1": ["-DMR_SUB=1"],
    "pz16": ["-DPZ_BLOCKS_PER_SM=16"],
    "pz12": ["-DPZ_BLOCKS_PER_SM=12"],
    "d8_pf16": ["-DED_DEPTH=8", "-DED_PFD=16"],
    "minb6": ["-DXY_MINB=6"],
    "minb4": ["-DXY_MINB=4"],
    "ring16_minb4": ["-DXY_MINB=4", "-DSR_RX=16", "-DSR_RY=16"],
    "t64": ["-DXY_THREADS=64", "-DXY_MINB=16"],
    "t32": ["-DXY_THREADS=32", "-DXY_MINB=32"],
    "t256": ["-DXY_THREADS=256", "-DXY_MINB=4"],
    "mr_noorder": ["-DMR_ORDER=0"],
    "mr8o": ["-DMR_SUB=8"],
    "mr4o_s1024": ["-DMR_STAGE_MAX=1024"],
    "mr4o_s64": ["-DMR_STAGE_MAX=64"],
    "mr8o_s1024": ["-DMR_SUB=8", "-DMR_STAGE_MAX=1024"],
    "mr4_order": ["-DMR_SUB=4"],
    "mr4_s256": ["-DMR_STAGE_MAX=256"],
    "mr4_s128": ["-DMR_STAGE_MAX=128"],
    "mr4_s64": ["-DMR_STAGE_MAX=64"],
    "mr2_s256": ["-DMR_SUB=2", "-DMR_STAGE_MAX=256"],
    "mr2_s128": ["-DMR_SUB=2", "-DMR_STAGE_MAX=128"],
    "mr2_s64": ["-DMR_SUB=2", "-DMR_STAGE_MAX=64"],
    "mr4_s4096": ["-DMR_STAGE_MAX=4096"],
}
names = sys.argv[1:] or list(VARIANTS)
for n in names:
    print(n, vb.build_variant(n, VARIANTS[n]))
