"""GPU parity tests: the CUDA path, called through the C ABI (include/voxcore_gpu.h), against the
CPU oracle (oracle/oracle.c, pinned to the reference in test_oracle_pinning.py) on the same seeded
inputs.  Integer / index / flag results must be bit-exact; float32 measures must be bit-exact too
(BASELINE: "within 1e-9 relative" -- the tolerance used below is 0)."""
import numpy as np
import pytest

from oracle import bindings as ob
from tests.cases import small_cases
from voxel_ma_b200 import api, synth

pytestmark = pytest.mark.gpu

CASES = small_cases()


@pytest.fixture(scope="module")
def ctx(ctx_factory):
    return ctx_factory()


def _run_all(ctx, vol):
    nz, ny, nx = vol.shape
    ctx.set_grid(nx, ny, nz)
    ctx.upload_volume(vol)
    inside = ctx.classify_grid()
    n = ctx.extract_sites()
    sites = ctx.get_sites()
    assert len(sites) == n
    return inside, sites


@pytest.mark.parametrize("name", sorted(CASES))
def test_classify_and_sites_bit_exact(ctx, name):
    vol = CASES[name]
    inside, sites = _run_all(ctx, vol)
    o_inside = ob.classify_grid(vol)
    assert np.array_equal(inside, o_inside)
    o_sites = ob.extract_sites(o_inside)
    assert sites.shape == o_sites.shape
    assert np.array_equal(sites, o_sites), "site order must be the reference's first-encounter order"


@pytest.mark.parametrize("name", sorted(CASES))
def test_closest_grid_bit_exact(ctx, name):
    vol = CASES[name]
    inside, sites = _run_all(ctx, vol)
    if len(sites) == 0:
        pytest.skip("no boundary")
    nz, ny, nx = vol.shape
    ids, d2x4 = ctx.closest_grid()
    o_ids, o_d2x4 = ob.closest_grid(sites, nx, ny, nz)
    assert np.array_equal(d2x4, o_d2x4)
    assert np.array_equal(ids, o_ids), "ties must go to the lowest site id (ANNbruteForce rule)"


@pytest.mark.parametrize("name", sorted(CASES))
def test_cell_measures_bit_exact(ctx, name):
    vol = CASES[name]
    inside, sites = _run_all(ctx, vol)
    if len(sites) == 0:
        pytest.skip("no boundary")
    nz, ny, nx = vol.shape
    ids, _ = ctx.closest_grid()
    e, f, c, r = ctx.cell_measures_grid()
    oe, of, oc, orad = ob.cell_measures_grid(sites, ids, inside, nx, ny, nz)
    for got, want, what in ((e, oe, "edge"), (f, of, "face"), (c, oc, "cube"), (r, orad, "radius")):
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), what  # tolerance: 0 ulp


def test_run_dense_matches_staged(ctx):
    vol = synth.sphere(48)
    inside, sites = _run_all(ctx, vol)
    ids, d2 = ctx.closest_grid()
    e, f, c, r = ctx.cell_measures_grid()
    n = ctx.run_dense()
    assert n == len(sites)
    assert np.array_equal(ctx.download(api.ARR_INSIDE), inside)
    assert np.array_equal(ctx.download(api.ARR_ID), ids)
    assert np.array_equal(ctx.download(api.ARR_D2X4), d2)
    assert np.array_equal(ctx.download(api.ARR_EDGE3), e)
    assert np.array_equal(ctx.download(api.ARR_FACE3), f)
    assert np.array_equal(ctx.download(api.ARR_CUBE), c)
    assert np.array_equal(ctx.download(api.ARR_RADIUS), r)


def test_run_dense_host_matches(ctx):
    vol = synth.torus(56)
    nz, ny, nx = vol.shape
    ctx.set_grid(nx, ny, nz)
    s = (nz, ny, nx)
    inside = np.empty(s, np.uint8)
    ids = np.empty(s, np.int32)
    d2 = np.empty(s, np.uint32)
    e, f = np.empty((3,) + s, np.float32), np.empty((3,) + s, np.float32)
    c, r = np.empty(s, np.float32), np.empty(s, np.float32)
    n = ctx.run_dense_host(vol, inside, ids, d2, e, f, c, r)
    o_inside = ob.classify_grid(vol)
    o_sites = ob.extract_sites(o_inside)
    assert n == len(o_sites)
    o_ids, o_d2 = ob.closest_grid(o_sites, nx, ny, nz)
    assert np.array_equal(inside, o_inside) and np.array_equal(ids, o_ids) and np.array_equal(d2, o_d2)
    oe, of, oc, orad = ob.cell_measures_grid(o_sites, o_ids, o_inside, nx, ny, nz)
    assert np.array_equal(e, oe) and np.array_equal(f, of) and np.array_equal(c, oc) and np.array_equal(r, orad)


def _check_compact(vol, n, nsites, bits, vert, ids, d2, lam, rad):
    """compact records == the oracle's dense planes at the inside vertices, in ascending linear order"""
    nz, ny, nx = vol.shape
    o_inside = ob.classify_grid(vol)
    o_sites = ob.extract_sites(o_inside)
    o_ids, o_d2 = ob.closest_grid(o_sites, nx, ny, nz) if len(o_sites) else (np.full(vol.shape, -1, np.int32), None)
    where = np.flatnonzero(o_inside.ravel()).astype(np.uint32)
    assert n == len(where) and nsites == len(o_sites)
    assert np.array_equal(vert[:n], where)
    if bits is not None:
        wr = nx // 32 + 1
        unpacked = np.unpackbits(bits.reshape(nz, ny, wr).view(np.uint8), axis=-1, bitorder="little")[:, :, :nx]
        assert np.array_equal(unpacked, o_inside)
    if n == 0 or len(o_sites) == 0:
        return
    oe, of, oc, orad = ob.cell_measures_grid(o_sites, o_ids, o_inside, nx, ny, nz)
    assert np.array_equal(ids[:n], o_ids.ravel()[where]) and np.array_equal(d2[:n], o_d2.ravel()[where])
    planes = [oe[0], oe[1], oe[2], of[0], of[1], of[2], oc]
    for k, pl in enumerate(planes):
        assert np.array_equal(lam[k, :n], pl.ravel()[where]), f"lambda plane {k}"
    assert np.array_equal(rad[:n], orad.ravel()[where])


@pytest.mark.parametrize("name", sorted(CASES))
def test_compact_records_bit_exact(ctx, name):
    vol = CASES[name]
    nz, ny, nx = vol.shape
    ctx.set_grid(nx, ny, nz)
    ctx.upload_volume(vol)
    ns = ctx.run_dense()
    vert, ids, d2, lam, rad = ctx.compact_records()
    _check_compact(vol, len(vert), ns, None, vert, ids, d2, lam, rad)


@pytest.mark.parametrize("mode", [1, 2])
@pytest.mark.parametrize("fam,n,workers,zchunk", [("torus", 56, 8, 0), ("twist", 40, 3, 5), ("assembly", 33, 0, 0), ("sphere", 64, 4, 7)])
def test_run_dense_host_compact_matches(ctx, fam, n, workers, zchunk, mode):
    """host volume in, compact product out (upload classified chunk by chunk, records copied back per z chunk);
    mode 1 = dense measure planes + gather, mode 2 = records computed directly per inside vertex"""
    vol = synth.make(fam, n)
    nz, ny, nx = vol.shape
    ctx.set_grid(nx, ny, nz)
    ctx.set_pipeline(workers, zchunk)
    ctx.set_compact_mode(mode)
    with pytest.raises(api.VoxcoreError, match="capacity"):
        ctx.run_dense_host_compact(vol, 1, vert=np.empty(1, np.uint32))
    cap = int((vol > 0).sum()) + 17
    bits = np.empty((nz * ny, nx // 32 + 1), np.uint32)
    vert, ids, d2 = np.empty(cap, np.uint32), np.empty(cap, np.int32), np.empty(cap, np.uint32)
    lam, rad = np.empty((7, cap), np.float32), np.empty(cap, np.float32)
    idd, d2d = np.empty(vol.shape, np.int32), np.empty(vol.shape, np.uint32)
    for _ in range(2):  # twice: buffers reused
        n_in, ns = ctx.run_dense_host_compact(vol, cap, bits, vert, ids, d2, lam, rad, idd, d2d)
        _check_compact(vol, n_in, ns, bits, vert, ids, d2, lam, rad)
        o_sites = ob.extract_sites(ob.classify_grid(vol))
        o_ids, o_d2 = ob.closest_grid(o_sites, nx, ny, nz)
        assert np.array_equal(idd, o_ids) and np.array_equal(d2d, o_d2)
    # record arrays that are the rows of one 11 x cap block take the single-copy path
    blk = np.empty((11, cap), np.uint32)
    n_in, ns = ctx.run_dense_host_compact(vol, cap, bits, blk[0], blk[1].view(np.int32), blk[2], blk[3:10].view(np.float32),
                                          blk[10].view(np.float32))
    _check_compact(vol, n_in, ns, bits, blk[0], blk[1].view(np.int32), blk[2], blk[3:10].view(np.float32), blk[10].view(np.float32))
    if mode == 2:  # the dense float planes were not produced by that call, and the ABI says so
        with pytest.raises(api.VoxcoreError, match="not computed"):
            ctx.download(api.ARR_CUBE)


@pytest.mark.parametrize("name", sorted(CASES))
def test_run_dense_host_compact_direct_records_on_edge_cases(ctx, name):
    vol = CASES[name]
    nz, ny, nx = vol.shape
    ctx.set_grid(nx, ny, nz)
    ctx.set_compact_mode(2)
    cap = int(ob.classify_grid(vol).sum()) + 1
    vert, ids, d2 = np.empty(cap, np.uint32), np.empty(cap, np.int32), np.empty(cap, np.uint32)
    lam, rad = np.empty((7, cap), np.float32), np.empty(cap, np.float32)
    n_in, ns = ctx.run_dense_host_compact(np.ascontiguousarray(vol, np.float32), cap, None, vert, ids, d2, lam, rad)
    _check_compact(vol, n_in, ns, None, vert, ids, d2, lam, rad)


def test_f64_zfast_upload(ctx):
    vol = synth.sphere(24)
    nz, ny, nx = vol.shape
    ctx.upload_volume_f64_zfast(synth.to_zfast_f64(vol), nx, ny, nz)
    inside = ctx.classify_grid()
    assert np.array_equal(inside, ob.classify_grid(vol))
    assert ctx.extract_sites() == len(ob.extract_sites(inside))


def test_slab_contexts_reproduce_whole_grid(ctx_factory):
    """z-slab sharding on one device: two contexts, halo recompute, site exchange through host
    buffers -- the same calls the multi-GPU path makes around its all-gather."""
    vol = synth.assembly(40, count=10)
    nz, ny, nx = vol.shape
    whole = ctx_factory()
    inside, sites = _run_all(whole, vol)
    ids, d2 = whole.closest_grid()
    e, f, c, r = whole.cell_measures_grid()
    cuts = [0, 13, 40]
    parts = [ctx_factory() for _ in range(2)]
    recs = []
    for k, p in enumerate(parts):
        z0, z1 = cuts[k], cuts[k + 1]
        p.set_grid(nx, ny, nz, z0, z1)
        lo, hi = max(z0 - 1, 0), min(z1 + 1, nz)
        p.upload_volume(vol[lo:hi], zlo=lo)
        assert np.array_equal(p.classify_grid(), inside[z0:z1])
        n = p.sites_detect_local()
        keys, corners = np.empty(n, np.uint64), np.empty(n, np.uint64)
        p.sites_export_local(keys, corners)
        recs.append((keys, corners))
    keys = np.concatenate([k for k, _ in recs])
    corners = np.concatenate([c for _, c in recs])
    assert len(keys) == len(sites)
    for k, p in enumerate(parts):
        z0, z1 = cuts[k], cuts[k + 1]
        p.sites_import_global(keys, corners, len(keys))
        assert np.array_equal(p.get_sites(), sites)
        pi, pd = p.closest_grid()
        assert np.array_equal(pi, ids[z0:z1]) and np.array_equal(pd, d2[z0:z1])
        pe, pf, pc, pr = p.cell_measures_grid()
        assert np.array_equal(pe, e[:, z0:z1]) and np.array_equal(pf, f[:, z0:z1])
        assert np.array_equal(pc, c[z0:z1]) and np.array_equal(pr, r[z0:z1])


@pytest.mark.parametrize("workers,zchunk", [(0, 0), (1, 1), (3, 5), (8, 0), (16, 2), (2, 1000)])
def test_pipeline_shape_does_not_change_results(ctx_factory, workers, zchunk):
    """z-chunk pipeline over worker streams (halo plane recomputed per chunk) == the staged single-stream calls"""
    vol = synth.assembly(44, count=14)
    c = ctx_factory()
    inside, sites = _run_all(c, vol)
    ids, d2 = c.closest_grid()
    e, f, cu, r = c.cell_measures_grid()
    c.set_pipeline(workers, zchunk)
    for _ in range(2):  # second pass: buffers reused, chunks race on the shared halo planes with equal values
        assert c.run_dense() == len(sites)
        assert np.array_equal(c.download(api.ARR_ID), ids) and np.array_equal(c.download(api.ARR_D2X4), d2)
        assert np.array_equal(c.download(api.ARR_EDGE3), e) and np.array_equal(c.download(api.ARR_FACE3), f)
        assert np.array_equal(c.download(api.ARR_CUBE), cu) and np.array_equal(c.download(api.ARR_RADIUS), r)


def test_slab_contexts_pipelined_second_half(ctx_factory):
    """the multi-GPU step as bench.py runs it: classify, detect, exchange, import, vc_closest_and_measures"""
    vol = synth.twist(40)
    nz, ny, nx = vol.shape
    o_inside = ob.classify_grid(vol)
    o_sites = ob.extract_sites(o_inside)
    o_ids, o_d2 = ob.closest_grid(o_sites, nx, ny, nz)
    oe, of, oc, orad = ob.cell_measures_grid(o_sites, o_ids, o_inside, nx, ny, nz)
    cuts = [0, 9, 10, 31, 40]  # includes a one-plane slab
    parts, recs = [], []
    for k in range(len(cuts) - 1):
        p = ctx_factory()
        z0, z1 = cuts[k], cuts[k + 1]
        p.set_grid(nx, ny, nz, z0, z1)
        lo, hi = max(z0 - 1, 0), min(z1 + 1, nz)
        p.upload_volume(vol[lo:hi], zlo=lo)
        p.classify_grid(fetch=False)
        n = p.sites_detect_local()
        keys, corners = np.empty(n, np.uint64), np.empty(n, np.uint64)
        p.sites_export_local(keys, corners)
        parts.append(p)
        recs.append((keys, corners))
    keys = np.concatenate([k for k, _ in recs][::-1])  # any order of the records must do
    corners = np.concatenate([c for _, c in recs][::-1])
    for k, p in enumerate(parts):
        z0, z1 = cuts[k], cuts[k + 1]
        p.set_pipeline(4, 3)
        p.sites_import_global(keys, corners, len(keys))
        p.closest_and_measures()
        assert np.array_equal(p.download(api.ARR_ID), o_ids[z0:z1]) and np.array_equal(p.download(api.ARR_D2X4), o_d2[z0:z1])
        assert np.array_equal(p.download(api.ARR_EDGE3), oe[:, z0:z1]) and np.array_equal(p.download(api.ARR_FACE3), of[:, z0:z1])
        assert np.array_equal(p.download(api.ARR_CUBE), oc[z0:z1]) and np.array_equal(p.download(api.ARR_RADIUS), orad[z0:z1])


def test_peer_exchange_matches_the_gathered_import(ctx_factory):
    """the exchange over peer memory (vc_peer.cu): records stored straight into every rank's receive
    buffer, posted with a release store, collected by waiting on the own header.  Three slab contexts
    of one process (vc_peer_open_ptrs), three exchanges in a row (both parities of the double
    buffer, and a changed volume in between), results == oracle."""
    vol = synth.twist(40)
    nz, ny, nx = vol.shape
    cuts = [0, 9, 10, 40]
    world = len(cuts) - 1
    parts = [ctx_factory() for _ in range(world)]
    for k, p in enumerate(parts):
        p.set_grid(nx, ny, nz, cuts[k], cuts[k + 1])
        p.peer_create(world, k, 20000)
    bases = [p.peer_buffer() for p in parts]
    for p in parts:
        p.peer_open_ptrs(bases)
    for it, v in enumerate([vol, synth.assembly(40, count=10), vol]):
        o_inside = ob.classify_grid(v)
        o_sites = ob.extract_sites(o_inside)
        o_ids, o_d2 = ob.closest_grid(o_sites, nx, ny, nz)
        for k, p in enumerate(parts):
            z0, z1 = cuts[k], cuts[k + 1]
            lo, hi = max(z0 - 1, 0), min(z1 + 1, nz)
            p.upload_volume(v[lo:hi], zlo=lo)
            p.classify_grid(fetch=False)
            p.sites_post_peers()  # asynchronous: nobody waits before everybody has posted
        for k, p in enumerate(parts):
            z0, z1 = cuts[k], cuts[k + 1]
            assert p.sites_collect_peers() == len(o_sites)
            assert np.array_equal(p.get_sites(), o_sites)
            p.closest_and_measures()
            assert np.array_equal(p.download(api.ARR_ID), o_ids[z0:z1]) and np.array_equal(p.download(api.ARR_D2X4), o_d2[z0:z1])
    for p in parts:
        p.peer_close()


def test_slab_group_host_compact_pipeline(ctx_factory):
    """vc_run_dense_host_compact on the slab contexts of a peer group (the N>1 end-to-end step): every rank uploads
    its own planes, the site records travel through the peer exchange, each rank gets the records of its owned
    planes.  The ranks of one process run in threads (the call blocks until every rank has posted)."""
    import threading
    vol = synth.make("twist", 48)
    nz, ny, nx = vol.shape
    o_inside = ob.classify_grid(vol)
    o_sites = ob.extract_sites(o_inside)
    o_ids, o_d2 = ob.closest_grid(o_sites, nx, ny, nz)
    oe, of, oc, orad = ob.cell_measures_grid(o_sites, o_ids, o_inside, nx, ny, nz)
    cuts = [0, 17, 30, 48]
    world = len(cuts) - 1
    parts = [ctx_factory() for _ in range(world)]
    for k, p in enumerate(parts):
        p.set_grid(nx, ny, nz, cuts[k], cuts[k + 1])
        p.peer_create(world, k, 40000)
    bases = [p.peer_buffer() for p in parts]
    for p in parts:
        p.peer_open_ptrs(bases)
    # One lock-step round first, from this thread, so that every device buffer exists: in ONE process a cudaMalloc /
    # cudaFree of one rank waits for the whole device, i.e. for another rank's wait kernel, which waits for this
    # rank's post (one process per GPU has no such coupling; here it would only trip the bounded wait).
    pins = []
    for k, p in enumerate(parts):
        z0, z1 = cuts[k], cuts[k + 1]
        lo, hi = max(z0 - 1, 0), min(z1 + 1, nz)
        pv = api.PinnedArray((hi - lo, ny, nx), np.float32)
        pv.array[...] = vol[lo:hi]
        pins.append(pv)
        p.upload_volume(pv.array, zlo=lo)
        p.classify_grid(fetch=False)
        p.sites_post_peers()
    for p in parts:
        p.sites_collect_peers()
        p.closest_and_measures()
        p.compact_records()
    res, errs = [None] * world, []

    def rank(k):
        try:
            z0, z1 = cuts[k], cuts[k + 1]
            cap = int(o_inside[z0:z1].sum()) + 5
            out = dict(bits=np.empty(((z1 - z0) * ny, nx // 32 + 1), np.uint32), vert=np.empty(cap, np.uint32), ids=np.empty(cap, np.int32),
                       d2=np.empty(cap, np.uint32), lam=np.empty((7, cap), np.float32), rad=np.empty(cap, np.float32))
            n_in, ns = parts[k].run_dense_host_compact(pins[k].array, cap, out["bits"], out["vert"], out["ids"], out["d2"],
                                                       out["lam"], out["rad"])
            res[k] = (n_in, ns, out)
        except Exception as ex:  # surfaced below
            errs.append(ex)

    for _ in range(2):  # both parities of the exchange buffers
        th = [threading.Thread(target=rank, args=(k,)) for k in range(world)]
        [t.start() for t in th]
        [t.join() for t in th]
        assert not errs, errs
        for k in range(world):
            z0, z1 = cuts[k], cuts[k + 1]
            n_in, ns, out = res[k]
            where = np.flatnonzero(o_inside[z0:z1].ravel()).astype(np.uint32)
            assert ns == len(o_sites) and n_in == len(where)
            assert np.array_equal(out["vert"][:n_in], where)
            assert np.array_equal(out["ids"][:n_in], o_ids[z0:z1].ravel()[where]) and np.array_equal(out["d2"][:n_in], o_d2[z0:z1].ravel()[where])
            planes = [oe[0], oe[1], oe[2], of[0], of[1], of[2], oc]
            for j, pl in enumerate(planes):
                assert np.array_equal(out["lam"][j, :n_in], pl[z0:z1].ravel()[where]), (k, j)
            assert np.array_equal(out["rad"][:n_in], orad[z0:z1].ravel()[where])
            unpacked = np.unpackbits(out["bits"].reshape(z1 - z0, ny, -1).view(np.uint8), axis=-1, bitorder="little")[:, :, :nx]
            assert np.array_equal(unpacked, o_inside[z0:z1])
    for p in parts:
        p.peer_close()


def test_peer_exchange_errors_do_not_hang(ctx_factory):
    """a rank that never posts -> VC_ERR_STATE after the bounded wait; too small a capacity -> VC_ERR_NOMEM"""
    vol = synth.twist(24)
    nz, ny, nx = vol.shape

    def pair(cap):
        a, b = ctx_factory(), ctx_factory()
        for k, p in enumerate((a, b)):
            z0, z1 = [0, 12][k], [12, 24][k]
            p.set_grid(nx, ny, nz, z0, z1)
            p.peer_create(2, k, cap)
            p.upload_volume(vol[max(z0 - 1, 0):min(z1 + 1, nz)], zlo=max(z0 - 1, 0))
            p.classify_grid(fetch=False)
        bases = [a.peer_buffer(), b.peer_buffer()]
        a.peer_open_ptrs(bases)
        b.peer_open_ptrs(bases)
        return a, b

    a, b = pair(20000)
    a.peer_set_timeout(300)
    a.sites_post_peers()
    with pytest.raises(api.VoxcoreError, match="timed out"):
        a.sites_collect_peers()  # b has not posted
    b.sites_post_peers()
    nref = len(ob.extract_sites(ob.classify_grid(vol)))
    assert b.sites_collect_peers() == nref  # a's post is still there
    assert a.sites_collect_peers() == nref  # ... and a's own exchange is still collectable after its time-out
    # ranks created with different capacities would overwrite each other's regions: detected, not silent
    a, b = pair(20000)
    b.peer_create(2, 1, 10000)
    bases = [a.peer_buffer(), b.peer_buffer()]
    with pytest.raises(api.VoxcoreError, match="another capacity"):
        a.peer_open_ptrs(bases)
    with pytest.raises(api.VoxcoreError, match="not mapped"):
        a.sites_post_peers()  # nothing is ever stored into a buffer laid out differently
    a, b = pair(16)
    with pytest.raises(api.VoxcoreError, match="capacity"):
        a.sites_post_peers()  # nothing is stored beyond a region: the rank that overflows says so before it posts
    with pytest.raises(api.VoxcoreError, match="no peer group"):
        ctx_factory().sites_post_peers()


@pytest.mark.parametrize("fam,n", [("twist", 256), ("torus", 256), ("assembly", 200)])
def test_full_size_properties(ctx_factory, fam, n):
    """sizes the CPU oracle cannot finish: size-independent properties instead.
    (1) d2x4 is exactly the distance to the reported site; (2) no site is closer at a sample of vertices
    (brute force over all sites); (3) ties report the lowest id; (4) radius == sqrt(d2) in float32;
    (5) measures are 0 wherever the anchor vertex is outside; (6) cube >= faces >= edges where valid."""
    vol = synth.make(fam, n)
    nz, ny, nx = vol.shape
    c = ctx_factory()
    c.set_grid(nx, ny, nz)
    c.upload_volume(vol)
    ns = c.run_dense()
    sites = c.get_sites()
    assert ns == len(sites) and ns > 0
    inside = c.download(api.ARR_INSIDE)
    assert np.array_equal(inside, (vol > 0).astype(np.uint8))
    ids, d2 = c.download(api.ARR_ID), c.download(api.ARR_D2X4)
    zz, yy, xx = np.meshgrid(np.arange(nz, dtype=np.float32), np.arange(ny, dtype=np.float32),
                             np.arange(nx, dtype=np.float32), indexing="ij", sparse=True)
    s = sites[ids]
    chk = 4.0 * ((s[..., 0] - xx) ** 2 + (s[..., 1] - yy) ** 2 + (s[..., 2] - zz) ** 2)
    assert np.array_equal(chk.astype(np.uint32), d2)                                     # (1)
    rng = np.random.default_rng(n)
    q = np.stack([rng.integers(0, nx, 600), rng.integers(0, ny, 600), rng.integers(0, nz, 600)], -1)
    dd = 4.0 * ((sites[None, :, :].astype(np.float64) - q[:, None, :]) ** 2).sum(-1)
    best = dd.min(1)
    assert np.array_equal(best.astype(np.uint32), d2[q[:, 2], q[:, 1], q[:, 0]])           # (2)
    first = np.array([np.flatnonzero(dd[i] == best[i])[0] for i in range(len(q))])
    assert np.array_equal(first.astype(np.int32), ids[q[:, 2], q[:, 1], q[:, 0]])          # (3)
    r = c.download(api.ARR_RADIUS)
    assert np.array_equal(r, np.sqrt(d2.astype(np.float32) * np.float32(0.25)))            # (4)
    e, f, cu = c.download(api.ARR_EDGE3), c.download(api.ARR_FACE3), c.download(api.ARR_CUBE)
    out = inside == 0
    assert not e[:, out].any() and not f[:, out].any() and not cu[out].any()               # (5)
    v = cu > 0
    assert (cu[v] >= f[:, v].max(0)).all() and (f.max(0)[v] >= 0).all()
    vf = f[0] > 0
    assert (f[0][vf] >= np.maximum(e[0], e[1])[vf]).all()                                   # (6)


def test_set_sites_lattice_and_points(ctx):
    vol = synth.sphere(32)
    inside, sites = _run_all(ctx, vol)
    nz, ny, nx = vol.shape
    # external sample set in a different order: ids follow the given order
    rng = np.random.default_rng(3)
    perm = rng.permutation(len(sites))
    ext = sites[perm]
    ctx.set_sites(ext)
    ids, d2 = ctx.closest_grid()
    o_ids, o_d2 = ob.closest_grid(ext, nx, ny, nz)
    assert np.array_equal(ids, o_ids) and np.array_equal(d2, o_d2)
    # arbitrary query points: drop-in for annkSearch(k=1, eps=0)
    q = rng.uniform(-3, 35, size=(4000, 3))
    q[:500] = np.round(q[:500] * 2) / 2  # many exact ties on the half-integer lattice
    pid, pd2 = ctx.closest_points(q)
    oid, od2 = ob.closest_points(ext, q)
    assert np.array_equal(pd2, od2)
    assert np.array_equal(pid, oid)


def test_general_sites_cell_list(ctx):
    rng = np.random.default_rng(11)
    nx, ny, nz = 24, 20, 16
    ctx.set_grid(nx, ny, nz)
    sites = rng.uniform(0, 20, size=(700, 3)).astype(np.float32)
    sites[100:110] = sites[90:100]  # duplicates: lowest id must win
    ctx.set_sites(sites)
    q = rng.uniform(-5, 30, size=(3000, 3))
    pid, pd2 = ctx.closest_points(q)
    oid, od2 = ob.closest_points(sites, q)
    assert np.array_equal(pd2, od2) and np.array_equal(pid, oid)
    ids, _ = ctx.closest_grid()
    o_ids, _, o_d2 = ob.closest_grid(sites, nx, ny, nz, want_d2=True)
    assert np.array_equal(ids, o_ids)


def test_complex_side_operators(ctx):
    vol = synth.sphere(32)
    inside, sites = _run_all(ctx, vol)
    rng = np.random.default_rng(5)
    pairs = rng.integers(0, len(sites), size=(5000, 2)).astype(np.int32)
    assert np.array_equal(ctx.face_lambda(pairs), ob.face_lambda(sites, pairs))
    v = rng.uniform(0, 31, size=(5000, 3)).astype(np.float32)
    sv = rng.integers(-1, len(sites), size=5000).astype(np.int32)
    assert np.array_equal(ctx.vertex_radii(v, sv), ob.vertex_radii(sites, v, sv))
    # tagVert on points that sit exactly on voxel boundaries (round half away from zero)
    p = rng.uniform(-2, 34, size=(6000, 3)).astype(np.float32)
    p[:3000] = np.round(p[:3000] * 2) / 2
    assert np.array_equal(ctx.classify_points(p), ob.classify_points(inside, p))
    M = np.array([0.5, 0, 0, 0, 0, 0.25, 0, 0, 0, 0, 2.0, 0, 1.0, -2.0, 3.0, 1.0])
    assert np.array_equal(ctx.classify_points(p, M), ob.classify_points(inside, p, M))
    # max aggregation over CSR adjacency
    lam = ctx.face_lambda(pairs)
    counts = rng.integers(0, 6, size=900)
    off = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
    items = rng.integers(0, len(lam), size=off[-1]).astype(np.int32)
    valid = (rng.random(len(lam)) > 0.3).astype(np.uint8)
    assert np.array_equal(ctx.segment_max(off, items, lam, valid), ob.segment_max(off, items, lam, valid))


def test_errors_are_reported(ctx_factory):
    c = ctx_factory()
    with pytest.raises(api.VoxcoreError):
        c.classify_grid()  # no grid / volume yet
    c.set_grid(8, 8, 8)
    with pytest.raises(api.VoxcoreError):
        c.closest_grid()  # no sites
    with pytest.raises(api.VoxcoreError):
        c.set_grid(4096, 8, 8)


# ---- fixed-radius query (annkFRSearch drop-in) ----------------------------------------------------
def test_radius_search_golden_inputs(ctx):
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "ann_fr_search.npz"))
    ctx.set_grid(16, 16, 16)
    ctx.set_sites(g["sites"].astype(np.float32))
    cnt, off, idx, d2 = ctx.radius_search(g["q"], g["sq_rad"])
    assert np.array_equal(cnt, g["count"])
    assert np.array_equal(idx, g["idx"]) and np.array_equal(d2, g["d2"])


def test_radius_search_general_sites_and_edge_radii(ctx):
    rng = np.random.default_rng(5)
    sites = rng.uniform(0, 40, (5000, 3)).astype(np.float32)
    q = rng.uniform(-5, 45, (3000, 3))
    sq = rng.uniform(0, 25, 3000)
    sq[:10] = 0.0          # only an exact hit is in range
    sq[10:20] = -1.0       # nothing is
    sq[20:24] = 1e9        # everything is
    q[:5] = sites[:5].astype(np.float64)  # exact hits (self matches are allowed, ANN_ALLOW_SELF_MATCH)
    ctx.set_grid(40, 40, 40)
    ctx.set_sites(sites)
    cnt, off, idx, d2 = ctx.radius_search(q, sq)
    ocnt, ooff, oidx, od2 = ob.radius_search(sites.astype(np.float64), q, sq)
    assert np.array_equal(cnt, ocnt) and (cnt[:5] >= 1).all() and (cnt[10:20] == 0).all() and (cnt[20:24] == 5000).all()
    assert np.array_equal(idx, oidx) and np.array_equal(d2, od2)
    assert np.array_equal(ctx.radius_search(q, sq, fetch=False), ocnt)


# ---- float32 nearest point (trimesh::KDtree::closest_to_pt drop-in) ---------------------
def test_closest_points_f32_golden(ctx):
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "kdtree_f32.npz"))
    ctx.set_grid(32, 32, 32)
    ctx.set_sites(g["pts"])  # lattice sites
    idx, d2 = ctx.closest_points_f32(g["q"], float(g["max_d2"]))
    assert np.array_equal(idx >= 0, g["found"])
    dist = np.where(idx >= 0, np.sqrt(np.maximum(d2, 0), dtype=np.float32), np.float32(-1))
    assert np.array_equal(dist, g["dist"]), "radii must equal the real KD-tree's bit for bit"
    oi, od2 = ob.closest_points_f32(g["pts"], g["q"], float(g["max_d2"]))
    assert np.array_equal(idx, oi) and np.array_equal(d2, od2)  # ties -> lowest id, as the restatement


def test_closest_points_f32_general_sites(ctx):
    rng = np.random.default_rng(31)
    pts = rng.uniform(0, 50, (20000, 3)).astype(np.float32)
    q = np.concatenate([rng.uniform(-10, 60, (6000, 3)), pts[:50]]).astype(np.float32)
    ctx.set_grid(50, 50, 50)
    ctx.set_sites(pts)
    for lim in (0.0, 3.0, float("inf")):
        idx, d2 = ctx.closest_points_f32(q, lim)
        oi, od2 = ob.closest_points_f32(pts, q, lim)
        assert np.array_equal(idx, oi) and np.array_equal(d2, od2)
    assert (d2[-50:] == 0).all()


# ---- MRC mode 0 volumes (signed bytes) ---------------------------------------------------------------
def _as_i8(vol):
    return np.clip(np.rint(vol * 16.0), -127, 127).astype(np.int8)


@pytest.mark.parametrize("fam,n", [("twist", 64), ("assembly", 33), ("sphere", 40), ("torus", 96)])
def test_int8_volume_classify_and_pipeline(ctx, fam, n):
    """vc_volume_upload_i8 / vc_run_dense_host_compact_i8: the byte volume classifies like the same values as floats"""
    v8 = _as_i8(synth.make(fam, n))
    vf = v8.astype(np.float32)
    nz, ny, nx = v8.shape
    o_inside = ob.classify_grid(vf)
    ctx.set_grid(nx, ny, nz)
    ctx.upload_volume(v8)
    assert np.array_equal(ctx.classify_grid(), o_inside)
    assert ctx.extract_sites() == len(ob.extract_sites(o_inside))
    cap = int(o_inside.sum()) + 3
    bits = np.empty((nz * ny, nx // 32 + 1), np.uint32)
    vert, ids, d2 = np.empty(cap, np.uint32), np.empty(cap, np.int32), np.empty(cap, np.uint32)
    lam, rad = np.empty((7, cap), np.float32), np.empty(cap, np.float32)
    n_in, ns = ctx.run_dense_host_compact(v8, cap, bits, vert, ids, d2, lam, rad)
    _check_compact(vf, n_in, ns, bits, vert, ids, d2, lam, rad)


def test_int8_volume_sign_edge_values(ctx):
    v8 = np.zeros((5, 7, 64), np.int8)
    v8[1:4, 2:5, 10:50] = np.array([-128, -1, 0, 1, 127], np.int8)[np.arange(40) % 5][None, None, :]
    ctx.set_grid(64, 7, 5)
    ctx.upload_volume(v8)
    assert np.array_equal(ctx.classify_grid(), (v8 > 0).astype(np.uint8))


def test_site_detection_capacity_regrows(ctx_factory):
    """single-pass site detection: a volume with far more sites than the context has seen before overflows the
    capacity carried over from the previous call and is detected again with the exact size"""
    c = ctx_factory()
    for vol in (synth.sphere(20), synth.torus(96), synth.sphere(20), synth.twist(72)):
        nz, ny, nx = vol.shape
        c.set_grid(nx, ny, nz)
        c.upload_volume(vol)
        inside = c.classify_grid()
        o_sites = ob.extract_sites(ob.classify_grid(vol))
        assert c.extract_sites() == len(o_sites)
        assert np.array_equal(c.get_sites(), o_sites)
