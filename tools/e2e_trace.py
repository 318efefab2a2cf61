"""Development aid: stage timeline of vc_run_dense_host_compact (VC_TRACE=1) on twist512."""
import os, sys, time
import numpy as np
sys.path.insert(0, ".")
os.environ["VC_TRACE"] = "1"
from voxel_ma_b200 import api, synth
vol = synth.make(sys.argv[1] if len(sys.argv) > 1 else "twist", int(sys.argv[2]) if len(sys.argv) > 2 else 512)
nz, ny, nx = vol.shape
pv = api.PinnedArray(vol.shape, np.float32); pv.array[...] = vol
c = api.Context(0); c.set_grid(nx, ny, nz)
cap = int((vol > 0).sum()) + 16
bits = api.PinnedArray((nz * ny, nx // 32 + 1), np.uint32)
vert = api.PinnedArray((cap,), np.uint32); ids = api.PinnedArray((cap,), np.int32); d2 = api.PinnedArray((cap,), np.uint32)
lam = api.PinnedArray((7, cap), np.float32); rad = api.PinnedArray((cap,), np.float32)
for i in range(5):
    t0 = time.perf_counter()
    c.run_dense_host_compact(pv.array, cap, bits.array, vert.array, ids.array, d2.array, lam.array, rad.array)
    print(f"wall {1e3*(time.perf_counter()-t0):.3f} ms", file=sys.stderr)
