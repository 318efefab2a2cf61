"""Deterministic input files for the CLI drop-in cases beyond vol2ma.
Used by tests/golden/make_golden.py (reference run, this container) and tests/test_gpu_zcli_dropin.py
(GPU run): both must feed the two binaries the same bytes."""
import os

import numpy as np

from voxel_ma_b200 import synth


def _f(x) -> str:
    return format(float(np.float32(x)), ".9g")


def write_radii_case(d: str, sites: np.ndarray) -> list[str]:
    """`-md=r ma.off bndry.node radii.txt`: a medial-axis stand-in (random points inside torus(48), some on the
    integer lattice so that several boundary points are exactly equidistant) and the boundary points as .node."""
    vol = synth.torus(48)
    rng = np.random.default_rng(8242)
    p = np.concatenate([rng.uniform(4, 43, (6000, 3)), np.floor(rng.uniform(4, 43, (2000, 3)))]).astype(np.float32)
    r = np.rint(p).astype(int)
    keep = vol[r[:, 2], r[:, 1], r[:, 0]] > 0
    p = p[keep]
    with open(os.path.join(d, "ma.off"), "w") as f:
        nf = len(p) // 3
        f.write(f"OFF\n{len(p)} {nf} 0\n")
        for v in p:
            f.write(f"{_f(v[0])} {_f(v[1])} {_f(v[2])}\n")
        for t in range(nf):
            f.write(f"3 {3*t} {3*t+1} {3*t+2}\n")
    with open(os.path.join(d, "bndry.node"), "w") as f:
        f.write("# boundary vertices of torus(48)\n")
        f.write(f"{len(sites)} 3 0 0\n")
        for i, s in enumerate(sites):
            f.write(f"{i} {_f(s[0])} {_f(s[1])} {_f(s[2])}\n")
    return ["-md=r", "ma.off", "bndry.node", "radii.txt"]


def write_funcmap_case(d: str) -> list[str]:
    """`-md=vol2ma -dofuncmap=bt3 -mcBase=mc vol.mrc out.ply`: a synthetic medial curve (points inside
    sphere(32), a forest order with 40 stationary roots, random measure)."""
    vol = synth.sphere(32)
    synth.write_mrc(os.path.join(d, "vol.mrc"), vol)
    rng = np.random.default_rng(977)
    p = rng.uniform(6, 25, (3000, 3)).astype(np.float32)
    r = np.rint(p).astype(int)
    p = p[vol[r[:, 2], r[:, 1], r[:, 0]] > 0][:600]
    n = len(p)
    order = np.arange(n)
    for i in range(40, n):
        order[i] = rng.integers(0, i)
    ms = rng.uniform(1, 5, n).astype(np.float32)
    with open(os.path.join(d, "mc.mc"), "w") as f:
        f.write(f"{n}\n")
        for v in p:
            f.write(f"{_f(v[0])} {_f(v[1])} {_f(v[2])}\n")
    with open(os.path.join(d, "mc.bt3.msure"), "w") as f:
        f.write(f"{n}\n")
        for s in ms:
            f.write(f"{_f(s)}\n")
    with open(os.path.join(d, "mc.mcorder"), "w") as f:
        f.write(f"{n}\n")
        for i in range(n):
            f.write(f"{i} {order[i]}\n")
    return ["-md=vol2ma", "-fullOrPruned=2", "-tt=0.04", "-dofuncmap=bt3", "-mcBase=mc", "vol.mrc", "out.ply"]
