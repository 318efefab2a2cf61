"""north_star stage 3, optional outputs: circumradius and object angle of each cell's closest-point set
(csrc/vc_circum.cu).  PARITY UNPINNED -- the reference has neither; the GPU planes are compared with the
oracle's independent restatement of the builder's definition (Welzl's recursion there, enumeration of support sets on the
GPU) to 1e-12, and with properties that hold by construction."""
import numpy as np
import pytest

from oracle import bindings as ob
from tests.cases import small_cases
from voxel_ma_b200 import api, synth

pytestmark = pytest.mark.gpu
CASES = small_cases()
TOL = 1e-12  # float64, same formulas in another order: absolute and relative


@pytest.mark.parametrize("name", sorted(CASES))
def test_circumradius_and_angle_planes(ctx_factory, name):
    vol = CASES[name]
    nz, ny, nx = vol.shape
    c = ctx_factory()
    c.set_grid(nx, ny, nz)
    c.upload_volume(vol)
    if c.run_dense() == 0:
        pytest.skip("no boundary")
    inside, ids, sites = c.download(api.ARR_INSIDE), c.download(api.ARR_ID), c.get_sites()
    circ, ang = c.cell_circum_angle_grid()
    o_circ, o_ang = ob.cell_circum_angle_grid(sites, ids, inside, nx, ny, nz)
    assert np.allclose(circ, o_circ, rtol=TOL, atol=TOL)
    assert np.allclose(ang, o_ang, rtol=TOL, atol=TOL)
    # by construction: a 2-set's ball is the diametral one (lambda / 2); both measures vanish with lambda; angle <= pi/2
    e3 = c.download(api.ARR_EDGE3).astype(np.float64)
    assert np.allclose(circ[:3], e3 / 2, rtol=1e-6, atol=1e-6)
    cube = c.download(api.ARR_CUBE)
    assert not circ[6][cube == 0].any() and not ang[6][cube == 0].any()
    assert (ang >= 0).all() and (ang <= np.pi / 2 + 1e-15).all()
    # a face's / the cube's ball encloses the balls of its edges' 2-sets
    assert (circ[3] >= np.maximum(circ[0], circ[1]) * (circ[3] > 0) - 1e-9).all()
    assert (circ[6] >= circ[3:6].max(0) * (circ[6] > 0) - 1e-9).all()


def test_circumradius_on_a_slab_and_a_plane_range(ctx_factory):
    vol = synth.make("torus", 40)
    nz, ny, nx = vol.shape
    whole = ctx_factory()
    whole.set_grid(nx, ny, nz)
    whole.upload_volume(vol)
    whole.run_dense()
    circ, ang = whole.cell_circum_angle_grid()
    c2, a2 = whole.cell_circum_angle_grid(11, 19)
    assert np.array_equal(c2, circ[:, 11:19]) and np.array_equal(a2, ang[:, 11:19])
    sites_k = whole.get_sites()
    # the same planes from a slab context (halo plane recomputed): identical
    p = ctx_factory()
    z0, z1 = 13, 27
    p.set_grid(nx, ny, nz, z0, z1)
    p.upload_volume(vol[z0 - 1:z1 + 1], zlo=z0 - 1)
    p.classify_grid(fetch=False)
    p.set_sites(sites_k)
    p.closest_and_measures()
    c3, a3 = p.cell_circum_angle_grid()
    assert np.array_equal(c3, circ[:, z0:z1]) and np.array_equal(a3, ang[:, z0:z1])
