"""The GPU drop-in CLI alone on one synthetic volume, with its trace (VC_DROPIN_TRACE=1) and stage lines.   python tools/cli_trace_only.py sphere:512"""
import os, subprocess, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from voxel_ma_b200 import synth  # noqa: E402
GPU = os.path.join(ROOT, "voxel_ma_b200", "host", "_build", "main_voroUtility_gpu")
kind, n = (sys.argv[1] if len(sys.argv) > 1 else "sphere:256").split(":")
with tempfile.TemporaryDirectory() as d:
    synth.write_mrc(os.path.join(d, "vol.mrc"), getattr(synth, kind)(int(n)))
    subprocess.check_call(["cp", os.path.join(os.path.dirname(GPU), "cycle8.txt"), d])
    t0 = time.perf_counter()
    r = subprocess.run([GPU, "-md=vol2ma", "-fullOrPruned=2", "-tt=0.04", "vol.mrc", "out.ply"], cwd=d, capture_output=True, text=True,
                       env=dict(os.environ, VC_DROPIN_TRACE="1"))
    print("wall", round(time.perf_counter() - t0, 3), "rc", r.returncode)
    print("\n".join(l for l in r.stdout.splitlines() if l.startswith("time")))
    print("\n".join(l for l in r.stderr.splitlines() if l.startswith("[vc dropin]")))
