"""All ranks of a sharded run as contexts of ONE process on ONE GPU (development aid: an 8-GPU box costs 8x the budget).

    python tools/peer_bench_1gpu.py [family:n] [nslabs] [rank to report] [reps]

Every slab context classifies its planes and detects / sorts / stores its run of site records into every other context's
receive buffer (vc_peer_open_ptrs: plain device pointers instead of IPC handles), then the reported rank collects and
runs its closest-site transform and measures.  The contexts share the GPU, so only the reported rank's PER-KERNEL times
mean anything (profiled: events around every launch) -- what the site pipeline of one rank costs per step (merge_rank,
line_sort, tables, pass Z) without paying for eight GPUs.  The slab bounds are bench.py's (slabs.aligned_bounds)."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, ".")
from voxel_ma_b200 import api, slabs, synth  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "assembly:1024"
world = int(sys.argv[2]) if len(sys.argv) > 2 else 8
who = int(sys.argv[3]) if len(sys.argv) > 3 else 2
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 5
fam, n = wl.split(":")
n = int(n)
nx = ny = nz = n
bounds = slabs.aligned_bounds(np.ones(nz), world, 32)
ndev = int(os.environ.get("VC_PEER_DEVICES", "1"))  # > 1: context k on device k % ndev (stores cross NVLink: ncu nvltx / nvlrx)
parts = []
for k in range(world):
    z0, z1 = bounds[k]
    lo, hi = max(z0 - 1, 0), min(z1 + 1, nz)
    c = api.Context(k % ndev)
    c.set_grid(nx, ny, nz, z0, z1)
    c.upload_volume(synth.make(fam, n, z0=lo, z1=hi), zlo=lo)
    parts.append(c)
cap = (6 << 20) // world + (1 << 18)
for k, c in enumerate(parts):
    c.peer_create(world, k, cap)
bases = [c.peer_buffer() for c in parts]
for c in parts:
    c.peer_open_ptrs(bases)


def step(profiled):
    for c in parts:
        c.classify_grid(fetch=False)
        c.sites_detect_local()
        c.sites_post_peers()
    me = parts[who]
    if profiled:
        me.profile(True)
        me.profile_reset()
    ns = me.sites_collect_peers()
    me.closest_and_measures()
    me.synchronize()
    rep = None
    if profiled:
        rep = me.profile_report()
        me.profile(False)
    for k, c in enumerate(parts):  # everybody collects: the double buffer needs every rank to follow
        if k != who:
            c.sites_collect_peers()
    return ns, rep


for _ in range(2):
    ns, _ = step(False)
acc = {}
for _ in range(reps):
    ns, rep = step(True)
    for k, v in rep.items():
        acc[k] = acc.get(k, 0.0) + v["ms"] / reps
# the pipelined (unprofiled) transform + measures of the reported rank, wall clock around a synchronised call
import time  # noqa: E402
me = parts[who]
wall = []
for _ in range(reps + 2):
    me.synchronize()
    t0 = time.perf_counter()
    me.closest_and_measures()
    me.synchronize()
    wall.append((time.perf_counter() - t0) * 1e3)
out = {"workload": f"{fam}{n}", "world": world, "rank": who, "planes": [int(v) for v in bounds[who]], "sites": int(ns),
       "kernels_ms_after_collect": {k: round(v, 4) for k, v in sorted(acc.items(), key=lambda kv: -kv[1])},
       "sum_ms": round(sum(acc.values()), 4),
       "closest_and_measures_pipelined_ms": round(float(np.median(wall[2:])), 4)}
print(json.dumps(out))
ids = parts[who].get_sites()
print("sites checksum", int(np.asarray(ids).view(np.uint32).sum(dtype=np.uint64)))
for c in parts:
    c.peer_close()
    c.close()
