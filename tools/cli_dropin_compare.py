"""Run the unmodified reference CLI (oracle/_ref/main_voroUtility) and the GPU drop-in CLI
(voxel_ma_b200/host/_build/main_voroUtility_gpu) on the same synthetic volume and compare every
output file byte for byte; prints wall times of both.  Test infrastructure (executes oracle/_ref).

    python tools/cli_dropin_compare.py sphere:64 torus:96 twist:64
"""
import hashlib
import json
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
        r = subprocess.run([exe, *ARGS], cwd=d, capture_output=True, text=True, env=dict(os.environ, VC_DROPIN_TRACE="1"))
        dt = time.perf_counter() - t0
        files = {f: hashlib.sha256(open(os.path.join(d, f), "rb").read()).hexdigest()
                 for f in sorted(os.listdir(d)) if f.startswith("out")}
        stages = [l.strip() for l in r.stdout.splitlines() if l.startswith("time")]
        trace = [l.strip()[len("[vc dropin] "):] for l in r.stderr.splitlines() if l.startswith("[vc dropin]")]
    return r.returncode, dt, files, stages, trace


def main():
    ok = True
    for spec in sys.argv[1:] or ["sphere:64"]:
        kind, n = spec.split(":")
        vol = getattr(synth, kind)(int(n))
        rc_r, t_r, f_r, s_r, _ = run(REF, vol)
        rc_g, t_g, f_g, s_g, tr_g = run(GPU, vol)
        same = rc_r == 0 and rc_g == 0 and f_r == f_g and len(f_r) > 0
        ok &= same
        print(json.dumps({"volume": spec, "identical_outputs": same, "files": sorted(f_r), "ref_wall_s": round(t_r, 3),
                          "gpu_cli_wall_s": round(t_g, 3), "rc": [rc_r, rc_g], "ref_stages": s_r, "gpu_stages": s_g, "gpu_trace": tr_g}))
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
