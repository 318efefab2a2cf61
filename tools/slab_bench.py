"""One rank's share of a sharded run, timed on ONE GPU (development aid: an 8-GPU box costs 8x the budget).

    python tools/slab_bench.py [family:n] [nslabs] [slab]

Builds the global site records of the family:n grid slab by slab (vc_sites_detect_local / export, as every rank of
`bench.py --gpus nslabs` does), then times what rank `slab` runs per step after the exchange: classify, numbering
of the union (vc_sites_import_global) and vc_closest_and_measures on its planes, pipelined and per kernel."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from voxel_ma_b200 import api, synth  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "assembly:1024"
nslabs = int(sys.argv[2]) if len(sys.argv) > 2 else 8
k = int(sys.argv[3]) if len(sys.argv) > 3 else 2
fam, n = wl.split(":")
n = int(n)
nx = ny = nz = n
h = nz // nslabs
scratch = api.Context(0)
keys, corners = [], []
for s in range(nslabs):
    z0, z1 = s * h, (s + 1) * h
    lo, hi = max(z0 - 1, 0), min(z1 + 1, nz)
    scratch.set_grid(nx, ny, nz, z0, z1)
    scratch.upload_volume(synth.make(fam, n, z0=lo, z1=hi), zlo=lo)
    scratch.classify_grid(fetch=False)
    m = scratch.sites_detect_local()
    kk, cc = np.empty(m, np.uint64), np.empty(m, np.uint64)
    scratch.sites_export_local(kk, cc)
    keys.append(kk)
    corners.append(cc)
scratch.close()
keys, corners = np.concatenate(keys), np.concatenate(corners)
z0, z1 = k * h, (k + 1) * h
if os.environ.get("VC_SLAB"):  # explicit planes "z0,z1" instead of slab k
    z0, z1 = (int(v) for v in os.environ["VC_SLAB"].split(","))
lo, hi = max(z0 - 1, 0), min(z1 + 1, nz)
c = api.Context(0)
c.set_grid(nx, ny, nz, z0, z1)
c.upload_volume(synth.make(fam, n, z0=lo, z1=hi), zlo=lo)
import torch  # noqa: E402
dk, dc = torch.from_numpy(keys.view(np.int64)).cuda(), torch.from_numpy(corners.view(np.int64)).cuda()


def step():
    c.classify_grid(fetch=False)
    c.sites_detect_local()
    c.sites_import_global(dk.data_ptr(), dc.data_ptr(), len(keys))
    c.closest_and_measures()


for _ in range(3):
    step()
c.synchronize()
reps = 5
t = time.time()
for _ in range(reps):
    step()
c.synchronize()
wall = (time.time() - t) / reps
c.profile(True)
c.profile_reset()
for _ in range(reps):
    step()
rep = c.profile_report()
tot = sum(v["ms"] for v in rep.values()) / reps
print(f"== {fam}{n} slab {k}/{nslabs} planes [{z0},{z1}) workers={os.environ.get('VC_WORKERS')} zchunk={os.environ.get('VC_ZCHUNK')}: "
      f"{len(keys)} sites, wall/step={wall*1e3:.3f} ms, kernels/step (serialized)={tot:.3f} ms")
for kname, v in sorted(rep.items(), key=lambda kv: -kv[1]["ms"])[:12]:
    print(f"   {kname:28s} {v['ms']/reps:9.3f} ms  x{v['launches']//reps}")
c.close()
