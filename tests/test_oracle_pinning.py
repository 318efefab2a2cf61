"""The CPU oracle (oracle/oracle.c) pinned against
  (1) golden vectors produced by the UNMODIFIED reference (tests/golden/, generator committed), incl.
      the only known-answer NN fixture of the reference tree, 3rdparty/ann/sample/sample.save;
  (2) the reference itself, live, when oracle/_ref is present (this container / the GPU box)."""
import os

import numpy as np
import pytest

from oracle import bindings as ob
from tests.cases import small_cases
from voxel_ma_b200 import synth

G = os.path.join(os.path.dirname(__file__), "golden")


def test_ann_sample_known_answers():
    g = np.load(os.path.join(G, "ann_sample.npz"))
    idx, d2 = ob.closest_points(g["data"], g["query"])
    assert np.array_equal(idx, g["nn_idx"])
    # sample.save prints sqrt(d2) with 6 significant digits (ann_sample.cpp takes the root)
    assert np.allclose(np.sqrt(d2), g["nn_dist"], rtol=5e-6)


@pytest.mark.parametrize("t", ["test1", "test2"])
def test_ann_test_files(t):
    g = np.load(os.path.join(G, f"ann_{t}.npz"))
    idx, d2 = ob.closest_points(g["data"], g["query"])  # 2-D and 8-D
    assert np.array_equal(idx, g["brute_idx"]) and np.array_equal(d2, g["brute_d2"])
    assert np.array_equal(d2, g["kd_d2"])  # average_error = 0 in test{1,2}.save


@pytest.mark.parametrize("k", ["sphere16", "twist20", "noise_9x14x11"])
def test_dense_small_golden(k):
    g = np.load(os.path.join(G, "dense_small.npz"))
    vol = g[k + "_vol"]
    nz, ny, nx = vol.shape
    inside = ob.classify_grid(vol)
    assert np.array_equal(inside, g[k + "_inside"])
    assert np.array_equal(ob.classify_grid_f64_zfast(synth.to_zfast_f64(vol), nx, ny, nz), g[k + "_inside"])
    sites = ob.extract_sites(inside)
    assert np.array_equal(sites, g[k + "_sites"])
    ids, d2x4, d2 = ob.closest_grid(sites, nx, ny, nz, want_d2=True)
    assert np.array_equal(ids, g[k + "_brute_idx"])
    assert np.array_equal(d2, g[k + "_d2"])
    assert np.array_equal(d2x4, (4 * g[k + "_d2"]).astype(np.uint32))
    # the kd-tree agrees on the distance everywhere and on the id wherever the minimum is unique
    kd = g[k + "_kd_idx"]
    s64 = sites.astype(np.float64)
    zz, yy, xx = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    q = np.stack([xx, yy, zz], -1).astype(np.float64)
    dk = ((s64[kd] - q) ** 2).sum(-1)
    assert np.array_equal(dk, d2)


def test_site_counts_and_hashes():
    g = np.load(os.path.join(G, "sites_hashes.npz"))
    import hashlib
    for n in (32, 64, 128):
        v = synth.sphere(n)
        inside = ob.classify_grid(v)
        sites = ob.extract_sites(inside)
        assert len(sites) == int(g[f"sphere{n}_nsites"])
        assert hashlib.sha256(sites.tobytes()).hexdigest() == str(g[f"sphere{n}_sites_sha256"])
        assert hashlib.sha256(inside.tobytes()).hexdigest() == str(g[f"sphere{n}_inside_sha256"])
    assert int(g["sphere64_nsites"]) == 9200 and int(g["sphere128_nsites"]) == 37328


def test_pipeline_golden_tagvert_lambda_radii():
    g = np.load(os.path.join(G, "pipeline_sphere24.npz"))
    inside = ob.classify_grid(g["vol"])
    sites = ob.extract_sites(inside)
    assert np.array_equal(sites, g["sites"])
    # a4: VoroInfo::tagVert on TetGen's raw Voronoi vertices (most sit exactly on half-integers)
    assert np.array_equal(ob.classify_points(inside, g["tet_vpts"]), g["tet_vtag"])
    # a7: computeFacesMeasure;  a8: computeInfoRelatedtoSites
    assert np.array_equal(ob.face_lambda(sites, g["face_sites"]), g["face_lambda"])
    assert np.array_equal(ob.vertex_radii(sites, g["vts"], g["site_of_v"]), g["radii"])
    # measure range after merge (extractInsideWithMeasure): every value is a lambda of some face
    lam = set(np.unique(g["face_lambda"]).tolist())
    assert set(np.unique(g["f_msure"]).tolist()) <= lam | {0.0}


@pytest.mark.skipif(not ob.have_ref(), reason="oracle/_ref not built")
@pytest.mark.parametrize("name", ["sphere32", "twist48", "noise_ragged", "sparse_ragged", "special_values", "full_box"])
def test_oracle_vs_live_reference(name):
    vol = small_cases()[name]
    nz, ny, nx = vol.shape
    inside = ob.classify_grid(vol)
    assert np.array_equal(inside, ob.ref_classify_grid(vol))
    sites = ob.extract_sites(inside)
    assert np.array_equal(sites, ob.ref_extract_sites(vol))
    if len(sites) == 0:
        return
    rng = np.random.default_rng(1)
    q = rng.uniform(-2, max(nx, ny, nz) + 2, size=(2000, 3))
    q[:800] = np.round(q[:800] * 2) / 2
    oi, od = ob.closest_points(sites, q)
    bi, bd = ob.ref_ann(sites, q, brute=True)
    ki, kd = ob.ref_ann(sites, q, brute=False)
    assert np.array_equal(oi, bi) and np.array_equal(od, bd) and np.array_equal(od, kd)
    p = rng.uniform(-1, max(nx, ny, nz), size=(3000, 3)).astype(np.float32)
    p[:1500] = np.round(p[:1500] * 2) / 2
    assert np.array_equal(ob.classify_points(inside, p), ob.ref_tag_points(vol, p))
    a, b = sites[rng.integers(0, len(sites), 4000)], sites[rng.integers(0, len(sites), 4000)]
    pairs = np.stack([np.arange(4000), np.arange(4000)], -1).astype(np.int32)
    both = np.concatenate([a, b])
    pairs[:, 1] += 4000
    assert np.array_equal(ob.face_lambda(both, pairs), ob.ref_lambda(a, b))


def test_fixed_radius_search_golden():
    """annkFRSearch (src/voxelapps.cpp:346-353): counts and the full in-range rows of the REAL ANNkd_tree
    (tests/golden/ann_fr_search.npz) -- inclusive radius, rows compared in (distance, id) order."""
    g = np.load(os.path.join(G, "ann_fr_search.npz"))
    cnt, off, idx, d2 = ob.radius_search(g["sites"], g["q"], g["sq_rad"])
    assert np.array_equal(cnt, g["count"])
    assert np.array_equal(idx, g["idx"]) and np.array_equal(d2, g["d2"])
    assert (cnt[100:200] >= 1).all()  # radius == nearest distance exactly: the nearest site is in range


@pytest.mark.skipif(not ob.have_ref(), reason="oracle/_ref not built")
def test_fixed_radius_search_vs_live_reference():
    rng = np.random.default_rng(99)
    sites = rng.uniform(0, 10, (400, 3))
    q = rng.uniform(-1, 11, (200, 3))
    sq = rng.uniform(0, 9, 200)
    rc, ri, rd = ob.ref_ann_fr(sites, q, sq)
    cnt, off, idx, d2 = ob.radius_search(sites, q, sq)
    assert np.array_equal(cnt, rc)
    for i in range(len(q)):
        o = np.lexsort((ri[i, : rc[i]], rd[i, : rc[i]]))
        assert np.array_equal(idx[off[i]:off[i + 1]], ri[i, : rc[i]][o])
        assert np.array_equal(d2[off[i]:off[i + 1]], rd[i, : rc[i]][o])


def test_kdtree_f32_golden():
    """8f-2: the float32 nearest-point restatement against the real trimesh::KDtree (golden from the reference)"""
    g = np.load(os.path.join(G, "kdtree_f32.npz"))
    idx, d2 = ob.closest_points_f32(g["pts"], g["q"], float(g["max_d2"]))
    assert np.array_equal(idx >= 0, g["found"]) and 0 < g["found"].sum() < len(g["found"])
    dist = np.where(idx >= 0, np.sqrt(np.maximum(d2, 0), dtype=np.float32), np.float32(-1))
    assert np.array_equal(dist, g["dist"])


@pytest.mark.skipif(not ob.have_ref(), reason="oracle/_ref not built")
def test_kdtree_f32_vs_live_reference():
    rng = np.random.default_rng(99)
    pts = rng.uniform(0, 30, (4000, 3)).astype(np.float32)
    q = rng.uniform(-3, 33, (3000, 3)).astype(np.float32)
    ridx, rdist = ob.ref_kdtree_closest(pts, q, 40.0)
    idx, d2 = ob.closest_points_f32(pts, q, 40.0)
    assert np.array_equal(ridx >= 0, idx >= 0)
    assert np.array_equal(rdist, np.where(idx >= 0, np.sqrt(np.maximum(d2, 0), dtype=np.float32), np.float32(-1)))
