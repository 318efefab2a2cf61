"""bench.py's contract, as far as it can be checked without a GPU: the reference arm prints ONE JSON line with the keys
the driver reads (and really runs the reference's operators), and the GPU arm fails loudly -- no JSON line, non-zero exit
-- when there is no device (no CPU fallback behind the product path)."""
import json
import os
import subprocess
import sys

import pytest

from oracle import bindings as ob

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args, env=None):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, cwd=ROOT,
                          env=dict(os.environ, **(env or {})), timeout=600)


@pytest.mark.skipif(not ob.have_ref(), reason="oracle/_ref (the compiled reference) is not built")
def test_reference_arm_line():
    r = _run("--impl", "reference", "--workload", "sphere64", "--steps", "1", "--warmup", "0")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, r.stdout[-2000:]
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "vertices/s" and d["higher_is_better"] is True
    assert d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 0 and d["scaling"] == "strong"
    assert d["config"]["workload"] == "sphere64x64x64"
    assert d["value"] > 0 and abs(d["ms_per_step"] - 64 ** 3 / d["value"] * 1e3) < 1e-6 * d["ms_per_step"]
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["sample"] and cb["value"] == d["value"]
    assert d["extrapolated"] is True and set(cb["sampled_fraction"]) == {"A_extract", "B_closest", "C_measures"}
    assert d["e2e"] == {"value": d["value"], "unit": "vertices/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    m = d["cpu_reference_measured"]  # the un-extrapolated point: all three stages over the whole torus256 grid
    assert m["extrapolated"] is False and m["value"] > 0 and m["seconds"]["B_ann_every_vertex"] > 0


def test_gpu_arm_fails_loudly_without_a_device():
    r = _run("--steps", "1", "--warmup", "0", "--workload", "sphere64", "--no-cpu-baseline", env={"CUDA_VISIBLE_DEVICES": ""})
    assert r.returncode != 0
    assert not [l for l in r.stdout.splitlines() if l.startswith("{")], "no bench line without a GPU"


def test_reference_arm_other_ranks_exit_quietly():
    r = _run("--impl", "reference", "--workload", "sphere64", "--steps", "1", "--warmup", "0", env={"RANK": "1", "WORLD_SIZE": "2"})
    assert r.returncode == 0 and r.stdout.strip() == ""
