"""GPU parity of stage 1' (vc_classify_mesh): the warp-ballot toggle/parity kernels against the
oracle's brute-force even-odd count, bit-exact flags, then the rest of the hot path on those flags."""
import numpy as np
import pytest

from oracle import bindings as ob
from tests.mesh_cases import mesh_cases
from voxel_ma_b200 import api, synth

pytestmark = pytest.mark.gpu
CASES = mesh_cases()


@pytest.fixture(scope="module")
def ctx(ctx_factory):
    return ctx_factory()


@pytest.mark.parametrize("name", sorted(CASES))
def test_mesh_flags_bit_exact(ctx, name):
    v, t, (nx, ny, nz), M = CASES[name]
    rc, want = ob.classify_mesh(v, t, nx, ny, nz, M)
    assert rc == 0
    ctx.set_grid(nx, ny, nz)
    got = ctx.classify_mesh(v, t, M)
    assert np.array_equal(got, want)
    assert np.array_equal(ctx.classify_grid(), want)  # the flags are the resident classification


def test_mesh_then_sites_closest_measures(ctx):
    """The flags of a mesh feed the unchanged downstream stages: same results as uploading a volume
    with the same occupancy."""
    n = 40
    v, t = synth.torus_mesh(n)
    ctx.set_grid(n, n, n)
    inside = ctx.classify_mesh(v, t)
    ns = ctx.run_dense()
    got = [ctx.download(a) for a in (api.ARR_ID, api.ARR_D2X4, api.ARR_EDGE3, api.ARR_FACE3, api.ARR_CUBE, api.ARR_RADIUS)]
    sites = ctx.get_sites()
    o_sites = ob.extract_sites(inside)
    assert ns == len(o_sites) and np.array_equal(sites, o_sites)
    o_ids, o_d2 = ob.closest_grid(o_sites, n, n, n)
    oe, of, oc, orad = ob.cell_measures_grid(o_sites, o_ids, inside, n, n, n)
    for g, w in zip(got, (o_ids, o_d2, oe, of, oc, orad)):
        assert np.array_equal(g, w)


def test_mesh_slab_contexts_agree_with_the_whole_grid(ctx, ctx_factory):
    n = 36
    v, t = synth.sphere_mesh(n)
    ctx.set_grid(n, n, n)
    whole = ctx.classify_mesh(v, t)
    c2 = ctx_factory()
    for z0, z1 in ((0, 13), (13, 29), (29, n)):
        c2.set_grid(n, n, n, z0, z1)
        part = c2.classify_mesh(v, t)
        assert np.array_equal(part, whole[z0:z1])


def test_mesh_errors(ctx):
    ctx.set_grid(8, 8, 8)
    v, t = synth.box_mesh((0, 0, 0), (5000, 4, 4))
    with pytest.raises(api.VoxcoreError, match="outside the supported range"):
        ctx.classify_mesh(v, t)
    v, t = synth.box_mesh((1, 1, 1), (5, 4, 4))
    t = t.copy()
    t[3, 1] = 99
    with pytest.raises(api.VoxcoreError, match="vertex index"):
        ctx.classify_mesh(v, t)


def test_mesh_large_properties(ctx):
    """256^3 torus, fine mesh (about 1 triangle per column): flags agree with the implicit solid away
    from the surface, and device-resident inputs give the same flags as host inputs."""
    n = 256
    v, t = synth.torus_mesh(n, nu=1024, nv=512)
    ctx.set_grid(n, n, n)
    ins = ctx.classify_mesh(v, t).astype(bool)
    z, y, x = np.meshgrid(*(np.arange(n, dtype=np.float32),) * 3, indexing="ij")
    cc = np.array([(n - 1) / 2.0 + 0.25, (n - 1) / 2.0 - 0.125, (n - 1) / 2.0 + 0.0625], np.float32)
    q = np.sqrt((x - cc[0]) ** 2 + (y - cc[1]) ** 2) - np.float32(88.0)
    f = np.float32(38.0) - np.sqrt(q * q + (z - cc[2]) ** 2)
    far = np.abs(f) > 0.25
    assert np.array_equal(ins[far], (f > 0)[far])
