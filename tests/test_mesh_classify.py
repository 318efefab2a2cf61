"""Stage 1' (mesh -> inside/outside flags) on the CPU: the oracle's brute-force even-odd count
against analytic solids, and the product's integer rules (voxel_ma_b200/csrc/vc_mesh_core.h, driven
through tests/host_harness.cpp in the kernels' toggle-row / suffix-parity data flow) against the
oracle.  PARITY UNPINNED w.r.t. the reference: it has no mesh voxeliser."""
import numpy as np
import pytest

from oracle import bindings as ob
from tests import hostcore
from tests.mesh_cases import mesh_cases
from voxel_ma_b200 import synth

CASES = mesh_cases()


def test_oracle_matches_analytic_sphere_and_torus():
    n = 48
    v, t = synth.sphere_mesh(n, nu=192, nv=96)
    rc, ins = ob.classify_mesh(v, t, n, n, n)
    assert rc == 0
    z, y, x = np.meshgrid(*(np.arange(n),) * 3, indexing="ij")
    c = np.array([(n - 1) / 2.0 + 0.137, (n - 1) / 2.0 - 0.211, (n - 1) / 2.0 + 0.319])
    d = np.sqrt((x - c[0]) ** 2 + (y - c[1]) ** 2 + (z - c[2]) ** 2)
    far = np.abs(d - 0.35 * n) > 0.25  # chord error of the tessellation + 1/256 snapping
    assert np.array_equal(ins.astype(bool)[far], (d < 0.35 * n)[far])
    v, t = synth.torus_mesh(n, nu=256, nv=128)
    rc, ins = ob.classify_mesh(v, t, n, n, n)
    R, r = 88.0 * n / 256, 38.0 * n / 256
    cc = np.array([(n - 1) / 2.0 + 0.25, (n - 1) / 2.0 - 0.125, (n - 1) / 2.0 + 0.0625])
    q = np.sqrt((x - cc[0]) ** 2 + (y - cc[1]) ** 2) - R
    f = r - np.sqrt(q * q + (z - cc[2]) ** 2)
    far = np.abs(f) > 0.25
    assert np.array_equal(ins.astype(bool)[far], (f > 0)[far])


def test_oracle_tie_rules_on_lattice_boxes():
    """Centres exactly on the surface are decided by the tie rules: a crossing at x_c flips 256 i < x_c
    (x: lo <= i < hi) and a column point on an edge is shifted by (+eps^2, -eps) in (y, z)
    (y: lo <= j < hi, z: lo < k <= hi) -- every lattice box owns a full hi - lo voxels per axis."""
    v, t = synth.box_mesh((4, 4, 4), (20, 10, 12))
    rc, ins = ob.classify_mesh(v, t, 32, 24, 16)
    want = np.zeros((16, 24, 32), np.uint8)
    want[5:13, 4:10, 4:20] = 1
    assert rc == 0 and np.array_equal(ins, want)
    # two boxes sharing the face x = 12: the union has no seam and no double count
    v1, t1 = synth.box_mesh((4, 4, 4), (12, 10, 12))
    v2, t2 = synth.box_mesh((12, 4, 4), (20, 10, 12))
    rc, ins2 = ob.classify_mesh(np.concatenate([v1, v2]), np.concatenate([t1, t2 + 8]), 32, 24, 16)
    assert np.array_equal(ins2, want)


@pytest.mark.parametrize("name", sorted(CASES))
def test_kernel_rules_match_oracle(name):
    v, t, (nx, ny, nz), M = CASES[name]
    if M is not None:
        pytest.skip("the host harness takes voxel-space vertices (the transform is covered on the GPU)")
    rc_o, want = ob.classify_mesh(v, t, nx, ny, nz)
    rc_h, got = hostcore.classify_mesh(v, t, nx, ny, nz)
    assert rc_o == rc_h == 0
    assert np.array_equal(got, want)


def test_out_of_range_vertex_is_an_error():
    v, t = synth.box_mesh((0, 0, 0), (5000, 4, 4))
    assert ob.classify_mesh(v, t, 8, 8, 8)[0] == 1
    assert hostcore.classify_mesh(v, t, 8, 8, 8)[0] == 1
