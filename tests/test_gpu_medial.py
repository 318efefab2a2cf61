"""the medial complex of the dense product (csrc/vc_medial.cu, host/tools/densecore_gpu.cpp).

The dual quads are compared bit for bit with the oracle's restatement of the builder's rule (PARITY UNPINNED for
the complex itself: the reference's complex is TetGen's Voronoi diagram; lambda is the pinned lambdaForFace), and
the complex is judged on statistics: closed, lambda range, Euler characteristic through the reference's own
cellcomplex / thinning code (densecore_gpu links it)."""
import os
import re
import subprocess

import numpy as np
import pytest

from oracle import bindings as ob
from tests.cases import small_cases
from voxel_ma_b200 import api, medial, synth
from voxel_ma_b200 import build as vb

pytestmark = pytest.mark.gpu
CASES = small_cases()


@pytest.mark.parametrize("name", sorted(CASES))
def test_dual_quads_match_the_oracle(ctx_factory, name):
    vol = CASES[name]
    nz, ny, nx = vol.shape
    c = ctx_factory()
    c.set_grid(nx, ny, nz)
    c.upload_volume(vol)
    ns = c.run_dense()
    if ns == 0:
        pytest.skip("no boundary")
    inside, ids, sites = c.download(api.ARR_INSIDE), c.download(api.ARR_ID), c.get_sites()
    got = c.medial_quads()
    want = ob.medial_quads(sites, ids, inside, nx, ny, nz)
    for g, w, what in zip(got, want, ("anchor", "axis", "site_a", "site_b", "lambda")):
        assert g.shape == w.shape and np.array_equal(g.view(np.uint8), w.view(np.uint8)), what
    # lambda is the value the edge3 plane holds for that grid edge
    e3 = c.download(api.ARR_EDGE3).reshape(3, -1)
    assert np.array_equal(got[4], e3[got[1], got[0]])


def test_dual_quads_on_slabs(ctx_factory):
    """a slab context emits the quads anchored in its own planes: the union over slabs is the whole grid's list"""
    vol = synth.make("assembly", 48)
    nz, ny, nx = vol.shape
    o_inside = ob.classify_grid(vol)
    o_sites = ob.extract_sites(o_inside)
    o_ids, _ = ob.closest_grid(o_sites, nx, ny, nz)
    want = ob.medial_quads(o_sites, o_ids, o_inside, nx, ny, nz)
    cuts = [0, 13, 14, 33, 48]
    parts, recs = [], []
    for k in range(len(cuts) - 1):
        p = ctx_factory()
        z0, z1 = cuts[k], cuts[k + 1]
        lo, hi = max(z0 - 1, 0), min(z1 + 1, nz)
        p.set_grid(nx, ny, nz, z0, z1)
        p.upload_volume(vol[lo:hi], zlo=lo)
        p.classify_grid(fetch=False)
        n = p.sites_detect_local()
        kk, cc = np.empty(n, np.uint64), np.empty(n, np.uint64)
        p.sites_export_local(kk, cc)
        parts.append(p)
        recs.append((kk, cc))
    keys, corners = np.concatenate([r[0] for r in recs]), np.concatenate([r[1] for r in recs])
    got = [[], [], [], [], []]
    for k, p in enumerate(parts):
        p.sites_import_global(keys, corners, len(keys))
        p.closest_and_measures()
        q = p.medial_quads()
        got[0].append(q[0].astype(np.int64) + cuts[k] * ny * nx)  # anchors are relative to the slab's first plane
        for j in range(1, 5):
            got[j].append(q[j])
    assert np.array_equal(np.concatenate(got[0]), want[0].astype(np.int64))
    for j in range(1, 5):
        assert np.array_equal(np.concatenate(got[j]), want[j])


def test_complex_statistics_sphere64(ctx_factory):
    vol = synth.sphere(64)
    nz, ny, nx = vol.shape
    c = ctx_factory()
    c.set_grid(nx, ny, nz)
    c.upload_volume(vol)
    c.run_dense()
    anchor, axis, a, b, lam = c.medial_quads()
    cx = medial.build_complex(anchor, axis, nx, ny)
    st = medial.statistics(cx, lam)
    # closed: every quad's 4 sides are edges of the complex, every vertex is a cube centre strictly inside the grid
    assert st["F"] == len(anchor) > 10000 and st["E"] >= st["F"] and st["V"] > 0
    v = cx["vertices"]
    assert (v > 0).all() and (v[:, 0] < nx - 1).all() and (v[:, 1] < ny - 1).all() and (v[:, 2] < nz - 1).all()
    # lambda: distances between distinct lattice sites of a ball of radius ~22 voxels (the reference's own range on the
    # same volume, SURVEY 8c, is [1.41421, 15] after its vertex merge; this complex keeps the unmerged near-surface faces)
    assert st["lambda_min"] >= 1.0 and st["lambda_max"] <= 2 * 0.36 * 64
    # sites of a quad are the closest sites of its two grid vertices: never equal
    assert (a != b).all()
    # the part of the complex deep inside (lambda >= 1/4 of the diameter) is one connected sheet system around the centre
    core, lcore = medial.thin_by_threshold(cx, lam, 0.25 * st["lambda_max"])
    centre = core["vertices"].mean(0)
    assert np.abs(centre - 31.0).max() < 1.5


def _tool():
    vb.build_dropin_cli()
    exe = os.path.join(os.path.dirname(vb.DROPIN_CLI), "densecore_gpu")
    if not os.path.exists(exe):
        pytest.skip("densecore_gpu was not built (needs the reference tree at build time)")
    return exe


def test_densecore_tool_runs_the_reference_back_end(tmp_path):
    """GPU dual quads -> the reference's cellcomplex + CellComplexThinning + PLY writer (all linked from the reference
    tree): files written, thinning keeps the Euler characteristic of what it collapses (simple-pair collapses are
    homotopy equivalences, src/ccthin.cpp:201-424) and shrinks the complex"""
    exe = _tool()
    d = str(tmp_path)
    synth.write_mrc(os.path.join(d, "sphere64.mrc"), synth.sphere(64))
    subprocess.check_call(["cp", os.path.join(os.path.dirname(exe), "cycle8.txt"), d])
    r = subprocess.run([exe, "sphere64.mrc", "core", "0.04,0.1"], cwd=d, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    out = r.stdout
    eulers = [int(m) for m in re.findall(r"euler = V - E \+ F = (-?\d+)", out)]
    vefs = [tuple(int(x) for x in m) for m in re.findall(r"V / E / F / T / C = (\d+) / (\d+) / (\d+) / \d+ / \d+", out)]
    assert len(eulers) == 3 and len(vefs) == 3, out[-2000:]
    assert vefs[0][2] > vefs[1][2] > vefs[2][2] > 0, vefs  # thinning removes faces, more with a larger threshold
    assert os.path.getsize(os.path.join(d, "core.ply")) > 100000
    assert os.path.exists(os.path.join(d, "core_thinned0.04.ply")) and os.path.exists(os.path.join(d, "core_thinned0.1.ply"))
    m = re.search(r"face measure range: \[([0-9.]+),([0-9.]+)\]", out)
    assert m and float(m.group(1)) >= 1.0 and float(m.group(2)) <= 46.1
    print(out[-1500:])
