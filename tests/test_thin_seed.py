"""K6 on the CPU: the oracle's restatement of the prune seeding (src/ccthin.cpp:246-270) against a
direct numpy statement of the same rule, and its pin to the queue size the reference CLI prints."""
import re
import os

import numpy as np

from oracle import bindings as ob
from tests.thin_cases import thin_cases

HERE = os.path.dirname(os.path.abspath(__file__))


def _numpy_pairs(c):
    e = np.flatnonzero(c["edge_ref"] == 1)
    f = c["edge_face0"][e]
    keep = (c["face_to_remove"][f] != 0) | (c["face_measure"][f] < np.float32(c["f_t"]))
    fe = np.stack([np.ones(keep.sum(), np.int32), f[keep], e[keep].astype(np.int32)], 1)
    v = np.flatnonzero(c["vert_ref"] == 1)
    ed = c["vert_edge0"][v]
    keep = c["edge_measure"][ed] < np.float32(c["l_t"])
    ev = np.stack([np.zeros(keep.sum(), np.int32), ed[keep], v[keep].astype(np.int32)], 1)
    return np.concatenate([fe, ev]).astype(np.int32)


def test_oracle_seeding_matches_the_rule():
    for name, c in thin_cases().items():
        with np.errstate(invalid="ignore"):
            want = _numpy_pairs(c)
        got = ob.simple_pairs(**c)
        assert np.array_equal(got, want), name


def test_oracle_ref_counts():
    rng = np.random.default_rng(0)
    idx = rng.integers(0, 1000, 50_000)
    assert np.array_equal(ob.ref_counts(idx, 1000), np.bincount(idx, minlength=1000))


def test_reference_cli_queue_size_is_recorded():
    """the pin used by the GPU CLI test: `after init, q size: 5472` on sphere64 (unmodified reference)"""
    txt = open(os.path.join(HERE, "golden", "cli_sphere64.txt")).read()
    assert re.search(r"after init, q size: 5472", txt)
