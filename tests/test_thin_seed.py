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


def test_oracle_seeding_matches_the_reference_queue():
    """tests/golden/thin_seed_sphere24.npz holds the queue the UNMODIFIED compiled CellComplexThinning::prune seeded on the
    reference's own inside complex of sphere24 (captured through the interposed prune_while_iteration,
    oracle/ref_thinspy.cpp), in push order, for three thresholds -- and the state it read.  The oracle's K6 must
    reproduce it exactly, and its reference counts must be the histograms of the complex's incidence lists."""
    d = np.load(os.path.join(HERE, "golden", "thin_seed_sphere24.npz"))
    for tag in ("lo", "mid", "hi"):
        c = {k: d[f"{tag}_{k}"] for k in ("edge_ref", "edge_face0", "face_measure", "vert_ref", "vert_edge0", "edge_measure",
                                          "face_to_remove")}
        got = ob.simple_pairs(f_t=float(d[f"{tag}_f_t"]), l_t=float(d[f"{tag}_l_t"]), **c)
        assert np.array_equal(got, d[f"{tag}_pairs"]), tag
        assert np.array_equal(ob.ref_counts(d[f"{tag}_edge_ends"], len(c["vert_ref"])), c["vert_ref"])
        assert np.array_equal(ob.ref_counts(d[f"{tag}_face_edges"], len(c["edge_ref"])), c["edge_ref"])
    assert len(d["mid_pairs"]) > 100 and len(d["lo_pairs"]) < len(d["mid_pairs"])
