"""Timestamped stdout of the reference CLI and of the GPU drop-in CLI on the same volume, side by side: where the two
processes' wall clocks part.  (Development aid; executes oracle/_ref.)   python tools/cli_timeline.py sphere:256"""
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from voxel_ma_b200 import synth  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref", "main_voroUtility")
GPU = os.path.join(ROOT, "voxel_ma_b200", "host", "_build", "main_voroUtility_gpu")
ARGS = ["-md=vol2ma", "-fullOrPruned=2", "-tt=0.04", "vol.mrc", "out.ply"]


def run(exe, vol):
    with tempfile.TemporaryDirectory() as d:
        synth.write_mrc(os.path.join(d, "vol.mrc"), vol)
        subprocess.check_call(["cp", os.path.join(os.path.dirname(exe), "cycle8.txt"), d])
        t0 = time.perf_counter()
        p = subprocess.Popen([exe, *ARGS], cwd=d, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, bufsize=1,
                             env=dict(os.environ, VC_DROPIN_TRACE="1"))
        out = []
        for line in p.stdout:
            out.append((time.perf_counter() - t0, line.rstrip()))
        p.wait()
        out.append((time.perf_counter() - t0, "<exit>"))
    return out


kind, n = (sys.argv[1] if len(sys.argv) > 1 else "sphere:256").split(":")
vol = getattr(synth, kind)(int(n))
run(GPU, vol)  # warm the box (page cache, driver)
a, b = run(REF, vol), run(GPU, vol)
ia = ib = 0
print(f"{'ref s':>8} {'gpu s':>8}  line")
while ia < len(a) or ib < len(b):
    la = a[ia] if ia < len(a) else None
    lb = b[ib] if ib < len(b) else None
    if la and lb and la[1] == lb[1]:
        print(f"{la[0]:8.3f} {lb[0]:8.3f}  {la[1][:100]}")
        ia += 1
        ib += 1
    elif lb and (not la or lb[1] not in [x[1] for x in a[ia:ia + 30]]):
        print(f"{'':8} {lb[0]:8.3f}  {lb[1][:100]}")
        ib += 1
    else:
        print(f"{la[0]:8.3f} {'':8}  {la[1][:100]}")
        ia += 1
