"""Seeded input volumes shared by the parity tests (sizes the CPU oracle finishes in seconds)."""
import numpy as np

from voxel_ma_b200 import synth


def small_cases():
    rng = np.random.default_rng(20181)
    cases = {
        "sphere32": synth.sphere(32),
        "torus40": synth.torus(40),
        "twist48": synth.twist(48),
        "assembly48": synth.assembly(48, count=12),
        "noise_ragged": rng.standard_normal((9, 14, 11)).astype(np.float32),   # nz=9, ny=14, nx=11
        "sparse_ragged": ((rng.random((20, 17, 23)) > 0.97).astype(np.float32) * 2 - 1),
        "full_box": np.ones((5, 6, 7), np.float32),                                # every voxel inside
        "single_voxel": np.pad(np.ones((1, 1, 1), np.float32), 3, constant_values=-1.0),
        "one_plane": (rng.random((1, 8, 9)) > 0.5).astype(np.float32) - 0.5,       # nz = 1
        "line": (rng.random((1, 1, 37)) > 0.5).astype(np.float32) - 0.5,           # ny = nz = 1
    }
    # zeros / -0.0 / NaN / inf are outside except +inf (include/spaceinfo.h:125-127: value > 0.0)
    special = rng.standard_normal((6, 7, 8)).astype(np.float32)
    special.flat[::5] = 0.0
    special.flat[1::7] = -0.0
    special.flat[2::11] = np.nan
    special.flat[3::13] = np.inf
    special.flat[4::17] = -np.inf
    cases["special_values"] = special
    return cases
