"""The kernels' exact-integer core (voxel_ma_b200/csrc/vc_core.h), driven line by line on the CPU
through tests/host_harness.cpp and compared with the oracle -- runs without a GPU."""
import numpy as np
import pytest

from oracle import bindings as ob
from tests import hostcore as hc
from tests.cases import small_cases

CASES = small_cases()


@pytest.mark.parametrize("name", sorted(CASES))
def test_site_numbering_matches_reference_order(name):
    inside = ob.classify_grid(CASES[name])
    assert np.array_equal(hc.sites(inside), ob.extract_sites(inside))


@pytest.mark.parametrize("name", sorted(CASES))
def test_separable_transform_is_exact(name):
    vol = CASES[name]
    nz, ny, nx = vol.shape
    sites = ob.extract_sites(ob.classify_grid(vol))
    if len(sites) == 0:
        pytest.skip("no boundary")
    ids, d2 = hc.closest_grid(sites, nx, ny, nz)
    o_ids, o_d2 = ob.closest_grid(sites, nx, ny, nz)
    assert np.array_equal(d2, o_d2) and np.array_equal(ids, o_ids)
    if nz > 6:
        ids2, d22 = hc.closest_grid(sites, nx, ny, nz, 2, nz - 3)
        assert np.array_equal(ids2, o_ids[2:nz - 3]) and np.array_equal(d22, o_d2[2:nz - 3])


def _brute_line(H, ntgt):
    p = np.arange(len(H), dtype=np.int64)[None, :]
    t = np.arange(ntgt, dtype=np.int64)[:, None]
    d = 2 * (t - p) + 1
    val = (H[None, :] >> np.uint64(32)).astype(np.int64) + d * d
    key = (val.astype(np.uint64) << np.uint64(32)) | (H[None, :] & np.uint64(0xFFFFFFFF))
    key = np.where(H[None, :] == np.uint64(0xFFFFFFFFFFFFFFFF), np.uint64(0xFFFFFFFFFFFFFFFF), key)
    return key.min(axis=1)


@pytest.mark.parametrize("n,dmax,density", [(7, 10, 1.0), (64, 50, 0.5), (513, 3_000_000, 0.3), (2049, 33_000_000, 0.7),
                                            (2049, 8, 1.0), (1025, 12_000_000, 0.02), (300, 0, 1.0)])
def test_envelope_line_random(n, dmax, density):
    """single lines with distances up to the 2048^3 range, dense ties (equal D, distinct ids)"""
    rng = np.random.default_rng(n * 7 + dmax % 97)
    for rep in range(6):
        D = (rng.integers(0, dmax + 1, size=n).astype(np.uint64) // np.uint64(8)) * np.uint64(8) + np.uint64(1)
        ids = rng.permutation(33_000_000)[:n].astype(np.uint64)
        H = (D << np.uint64(32)) | ids
        H[rng.random(n) > density] = np.uint64(0xFFFFFFFFFFFFFFFF)
        out = hc.envelope(H, n - 1)
        assert np.array_equal(out, _brute_line(H, n - 1))
        assert np.array_equal(hc.envelope_pruned(H, n - 1), out)
        # pass X's form: a bitmap says which candidates exist, the others are never read
        assert np.array_equal(hc.envelope_masked(H, n - 1), out)
        # the round-2 scan (the one the kernels run): integer-pruned stack over the live candidates
        assert np.array_equal(hc.envelope_pruned(H, n - 1), out)
        assert np.array_equal(hc.envelope_pruned(H, max(n // 3, 1)), _brute_line(H, max(n // 3, 1)))  # targets end early
    empty = np.full(n, 0xFFFFFFFFFFFFFFFF, np.uint64)
    assert (hc.envelope(empty, n - 1) == np.uint64(0xFFFFFFFFFFFFFFFF)).all()
    assert (hc.envelope_masked(empty, n - 1) == np.uint64(0xFFFFFFFFFFFFFFFF)).all()
    assert (hc.envelope_pruned(empty, n - 1) == np.uint64(0xFFFFFFFFFFFFFFFF)).all()


@pytest.mark.parametrize("n,levels", [(33, 2), (257, 3), (1025, 2), (2049, 4)])
def test_envelope_line_touching_parabolas(n, levels):
    """candidates built so that three and more parabolas meet in one point exactly at a target: the entry
    that only touches the envelope there must survive the forward scan when it holds the lowest id"""
    rng = np.random.default_rng(n + levels)
    for rep in range(3):
        t0 = int(rng.integers(1, n - 2))
        p = np.arange(n, dtype=np.int64)
        base = int((2 * (max(t0, n - t0)) + 3) ** 2 + 64)
        # all parabolas through (t0, base): D_p = base - (2(t0-p)+1)^2, some pushed up by multiples of 8
        D = base - (2 * (t0 - p) + 1) ** 2 + 8 * rng.integers(0, levels, n)
        ids = rng.permutation(33_000_000)[:n].astype(np.uint64)
        H = (D.astype(np.uint64) << np.uint64(32)) | ids
        H[rng.random(n) > 0.8] = np.uint64(0xFFFFFFFFFFFFFFFF)
        out = hc.envelope(H, n - 1)
        assert np.array_equal(out, _brute_line(H, n - 1))
        assert np.array_equal(hc.envelope_pruned(H, n - 1), out)


def test_sep_division_is_exact():
    """vc_sep's double-precision floor division against integer arithmetic: every divisor 8w, w in [1, 2048]"""
    assert hc.sep_sweep() == 0


@pytest.mark.parametrize("name", sorted(CASES))
def test_round2_data_flow_is_exact(name):
    """compact columns -> pass Z -> pass X over the row's columns -> pass Y over the live rows (vc_edt.cu's data flow)"""
    vol = CASES[name]
    nz, ny, nx = vol.shape
    sites = ob.extract_sites(ob.classify_grid(vol))
    if len(sites) == 0:
        pytest.skip("no boundary")
    o_ids, o_d2 = ob.closest_grid(sites, nx, ny, nz)
    ids, d2 = hc.closest_grid2(sites, nx, ny, nz)
    assert np.array_equal(d2, o_d2) and np.array_equal(ids, o_ids)
    if nz > 6:
        ids2, d22 = hc.closest_grid2(sites, nx, ny, nz, 2, nz - 3)
        assert np.array_equal(ids2, o_ids[2:nz - 3]) and np.array_equal(d22, o_d2[2:nz - 3])
