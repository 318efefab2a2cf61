"""Ad-hoc per-kernel timing on the GPU box (development aid; bench.py is the contract)."""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from voxel_ma_b200 import api, synth  # noqa: E402


def run(name, n, reps=3):
    t = time.time()
    vol = synth.make(name, n)
    gen = time.time() - t
    nz, ny, nx = vol.shape
    c = api.Context(0)
    c.set_grid(nx, ny, nz)
    c.upload_volume(vol)
    ns = c.run_dense()  # warm-up (allocations, local-memory resize)
    c.run_dense()
    c.synchronize()
    t = time.time()
    for _ in range(reps):
        c.run_dense()
    wall = (time.time() - t) / reps  # run_dense returns after a stream sync; pipelined over worker streams
    c.profile(True)  # profiling serialises the z-chunk pipeline onto one stream
    c.profile_reset()
    for _ in range(reps):
        c.run_dense()
    rep = c.profile_report()
    tot = sum(v["ms"] for v in rep.values()) / reps
    import os
    print(f"== {name}{n} workers={os.environ.get('VC_WORKERS')} zchunk={os.environ.get('VC_ZCHUNK')}: sites={ns} gen={gen:.1f}s wall/step={wall*1e3:.2f} ms kernels/step={tot:.2f} ms "
          f"-> {nx*ny*nz/wall:.3e} v/s (wall) {nx*ny*nz/(tot*1e-3):.3e} v/s (kernel sum)")
    for k, v in sorted(rep.items(), key=lambda kv: -kv[1]["ms"]):
        print(f"   {k:28s} {v['ms']/reps:9.3f} ms  x{v['launches']//reps}")
    # full-size property: d2x4 is consistent with the id
    ids = c.download(api.ARR_ID)
    d2 = c.download(api.ARR_D2X4)
    sites = c.get_sites()
    zz, yy, xx = np.meshgrid(np.arange(nz, dtype=np.float32), np.arange(ny, dtype=np.float32), np.arange(nx, dtype=np.float32), indexing="ij", sparse=True)
    s = sites[ids]
    chk = 4 * ((s[..., 0] - xx) ** 2 + (s[..., 1] - yy) ** 2 + (s[..., 2] - zz) ** 2)
    print("   d2 consistent with id:", bool(np.array_equal(chk.astype(np.uint32), d2)), " max d:", float(np.sqrt(d2.max() / 4)))
    c.close()
    return {"name": name, "n": n, "sites": ns, "ms": tot, "kernels": rep}


if __name__ == "__main__":
    out = []
    for arg in sys.argv[1:]:
        name, n = arg.split(":")
        out.append(run(name, int(n)))
    json.dump(out, open("gpurun_out/quick_bench.json", "w"), indent=1)
