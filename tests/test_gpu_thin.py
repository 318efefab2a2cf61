"""GPU parity of K6 (vc_ref_counts, vc_simple_pairs) against the oracle: same pairs, same order."""
import numpy as np
import pytest

from oracle import bindings as ob
from tests.thin_cases import thin_cases
from voxel_ma_b200 import api

pytestmark = pytest.mark.gpu
CASES = thin_cases()


@pytest.fixture(scope="module")
def ctx(ctx_factory):
    return ctx_factory()


@pytest.mark.parametrize("name", sorted(CASES))
def test_simple_pairs_same_queue(ctx, name):
    c = CASES[name]
    want = ob.simple_pairs(**c)
    got = ctx.simple_pairs(**c)
    assert got.shape == want.shape and np.array_equal(got, want)
    nomark = dict(c, face_to_remove=None)
    assert np.array_equal(ctx.simple_pairs(**nomark), ob.simple_pairs(**nomark))


def test_ref_counts(ctx):
    rng = np.random.default_rng(3)
    for n, bins in ((0, 5), (17, 3), (1_000_003, 65_537)):
        idx = rng.integers(0, bins, n).astype(np.int32)
        assert np.array_equal(ctx.ref_counts(idx, bins), ob.ref_counts(idx, bins))
    with pytest.raises(api.VoxcoreError, match="outside"):
        ctx.ref_counts(np.array([0, 7], np.int32), 5)


def test_simple_pairs_rejects_bad_neighbour(ctx):
    c = dict(CASES["small"])
    c["edge_ref"] = np.ones_like(c["edge_ref"])
    c["edge_face0"] = c["edge_face0"].copy()
    c["edge_face0"][5] = 10_000
    with pytest.raises(api.VoxcoreError, match="first-neighbour"):
        ctx.simple_pairs(**c)


def test_seeding_matches_the_reference_queue(ctx):
    """the queue the unmodified compiled CellComplexThinning::prune seeded (tests/golden/thin_seed_sphere24.npz, made by
    tests/golden/make_thin_golden.py through oracle/ref_thinspy.cpp): same pairs in the same order from vc_simple_pairs,
    the same reference counts from vc_ref_counts on the complex's own incidence lists"""
    import os
    d = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "thin_seed_sphere24.npz"))
    for tag in ("lo", "mid", "hi"):
        c = {k: d[f"{tag}_{k}"] for k in ("edge_ref", "edge_face0", "face_measure", "vert_ref", "vert_edge0", "edge_measure",
                                          "face_to_remove")}
        got = ctx.simple_pairs(f_t=float(d[f"{tag}_f_t"]), l_t=float(d[f"{tag}_l_t"]), **c)
        assert np.array_equal(got, d[f"{tag}_pairs"]), tag
        assert np.array_equal(ctx.ref_counts(d[f"{tag}_edge_ends"], len(c["vert_ref"])), c["vert_ref"])
        assert np.array_equal(ctx.ref_counts(d[f"{tag}_face_edges"], len(c["edge_ref"])), c["edge_ref"])
