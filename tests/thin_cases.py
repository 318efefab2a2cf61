"""Seeded inputs for the K6 tests (seeding of CellComplexThinning::prune)."""
import numpy as np


def thin_case(seed, ne, nf, nv):
    rng = np.random.default_rng(seed)
    edge_ref = rng.integers(0, 4, ne).astype(np.int32)
    vert_ref = rng.integers(0, 4, nv).astype(np.int32)
    edge_face0 = rng.integers(0, max(nf, 1), ne).astype(np.int32)
    vert_edge0 = rng.integers(0, max(ne, 1), nv).astype(np.int32)
    # measures on a coarse lattice so many sit exactly on the threshold (the test is strict <); some NaN
    face_measure = (rng.integers(0, 9, nf) * 0.25).astype(np.float32)
    edge_measure = (rng.integers(0, 9, ne) * 0.25).astype(np.float32)
    if nf > 3:
        face_measure[rng.integers(0, nf, 3)] = np.nan
    if ne > 3:
        edge_measure[rng.integers(0, ne, 3)] = np.nan
    to_remove = (rng.random(nf) < 0.05).astype(np.uint8)
    return dict(edge_ref=edge_ref, edge_face0=edge_face0, face_measure=face_measure, f_t=1.0, vert_ref=vert_ref,
                vert_edge0=vert_edge0, edge_measure=edge_measure, l_t=0.75, face_to_remove=to_remove)


def thin_cases():
    return {
        "small": thin_case(1, 37, 21, 29),
        "one_block": thin_case(2, 200, 90, 56),
        "ragged": thin_case(3, 5003, 3001, 2999),
        "many_blocks": thin_case(4, 300_000, 150_000, 120_011),  # > 1024 blocks: the block scan loops
        "no_vertices": thin_case(5, 100, 40, 0),
        "no_edges": thin_case(6, 0, 10, 0),
    }
