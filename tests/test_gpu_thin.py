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
