import os

import numpy as np
import pytest

from oracle import bindings as ob
from tests import hostcore as hc
from voxel_ma_b200 import slabs, synth


def test_generators_are_deterministic_and_slabwise():
    for name, n in (("sphere", 24), ("torus", 32), ("twist", 40), ("assembly", 40)):
        a = synth.make(name, n)
        assert a.dtype == np.float32 and a.shape == (n, n, n)
        assert np.array_equal(a, synth.make(name, n))
        assert np.array_equal(synth.make(name, n, z0=5, z1=17), a[5:17])
        assert 0 < (a > 0).mean() < 0.6
    b = synth.make("assembly", (24, 32, 48))
    assert b.shape == (48, 32, 24)


def test_mrc_roundtrip(tmp_path):
    v = synth.sphere(16)
    p = os.path.join(tmp_path, "v.mrc")
    synth.write_mrc(p, v)
    assert os.path.getsize(p) == 1024 + v.size * 4
    assert np.array_equal(synth.read_mrc(p), v)


def test_slab_bounds_cover_grid():
    for nz, w in ((512, 8), (513, 4), (10, 3), (7, 7)):
        b = [slabs.slab_bounds(nz, w, r) for r in range(w)]
        assert b[0][0] == 0 and b[-1][1] == nz
        assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
        assert max(z1 - z0 for z0, z1 in b) - min(z1 - z0 for z0, z1 in b) <= 1
    with pytest.raises(ValueError):
        slabs.slab_bounds(4, 5, 0)


def test_slab_site_records_union_equals_whole_grid_sites():
    vol = synth.assembly(36, count=9)
    nz, ny, nx = vol.shape
    inside = ob.classify_grid(vol)
    want = ob.extract_sites(inside)
    keys, corners = [], []
    for r in range(3):
        z0, z1 = slabs.slab_bounds(nz, 3, r)
        lo, hi = slabs.resident_planes(z0, z1, nz)
        czb, cze = slabs.owned_corner_planes(z0, z1, nz)
        k, c = hc.site_records(inside[lo:hi], nx, ny, nz, lo, czb, cze)
        keys.append(k)
        corners.append(c)
    k = np.concatenate(keys)
    c = np.concatenate(corners)
    assert len(np.unique(k)) == len(k) == len(want)
    got = slabs.unpack_corners(c[np.argsort(k)]).astype(np.float32) - 0.5
    assert np.array_equal(got, want)


def test_balanced_bounds():
    from voxel_ma_b200 import slabs
    rng = np.random.default_rng(3)
    for nz, world in ((128, 8), (1024, 8), (37, 5), (8, 8), (9, 2)):
        w = 1.0 + 5.0 * rng.random(nz) * (np.arange(nz) > nz // 3)
        b = slabs.balanced_bounds(w, world)
        assert b[0][0] == 0 and b[-1][1] == nz and all(b[r][1] == b[r + 1][0] for r in range(world - 1))
        assert all(z1 > z0 for z0, z1 in b)
        loads = np.array([w[z0:z1].sum() for z0, z1 in b])
        if nz >= 16 * world:  # fine enough to balance: no slab more than one heavy plane over the mean
            assert loads.max() - w.sum() / world <= w.max() + 1e-9
            uniform = np.array([w[z0:z1].sum() for z0, z1 in (slabs.slab_bounds(nz, world, r) for r in range(world))])
            assert loads.max() <= uniform.max() + 1e-9
    # uniform weights: equal heights (within one plane)
    b = slabs.balanced_bounds(np.ones(1000), 8)
    assert {z1 - z0 for z0, z1 in b} == {125}
    with pytest.raises(ValueError):
        slabs.balanced_bounds(np.ones(3), 4)


def test_aligned_bounds():
    """slab cuts snapped so that planes + halo plane of every slab below the top fill whole 32-plane warps of pass X"""
    from voxel_ma_b200 import slabs
    rng = np.random.default_rng(5)
    for nz, world in ((1024, 8), (1024, 4), (1024, 2), (2048, 8), (512, 8), (700, 3)):
        w = 1.0 + 3.0 * rng.random(nz)
        b = slabs.aligned_bounds(w, world, 32)
        assert b[0][0] == 0 and b[-1][1] == nz and all(b[r][1] == b[r + 1][0] for r in range(world - 1))
        assert all(z1 > z0 for z0, z1 in b)
        assert all((z1 - z0 + 1) % 32 == 0 for z0, z1 in b[:-1])
    assert [z1 - z0 for z0, z1 in slabs.aligned_bounds(np.ones(1024), 8, 32)] == [127] * 7 + [135]
    # too thin for aligned slabs: the balanced cuts are kept
    assert slabs.aligned_bounds(np.ones(40), 3, 32) == slabs.balanced_bounds(np.ones(40), 3)
    assert slabs.aligned_bounds(np.ones(64), 1, 32) == [(0, 64)]
