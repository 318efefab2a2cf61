"""world_size-2 gloo test of the multi-rank host logic (slab split, halo planes, variable-length
site-record all-gather, identical numbering on every rank).  The per-slab detection runs through the
kernels' integer core on the CPU (tests/host_harness.cpp); on the GPU box the same exchange carries
the records vc_sites_export_local produces (bench.py --gpus N)."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import bindings as ob
    from tests import hostcore as hc
    from voxel_ma_b200 import slabs, synth

    n = 30
    z0, z1 = slabs.slab_bounds(n, world, rank)
    lo, hi = slabs.resident_planes(z0, z1, n)
    vol = synth.assembly(n, count=8, z0=lo, z1=hi)  # each rank builds only its own planes
    inside = ob.classify_grid(vol)
    czb, cze = slabs.owned_corner_planes(z0, z1, n)
    k, c = hc.site_records(inside, n, n, n, lo, czb, cze)
    ak, ac = slabs.exchange_site_records(torch.from_numpy(k.view(np.int64)), torch.from_numpy(c.view(np.int64)))
    ak, ac = ak.numpy().view(np.uint64), ac.numpy().view(np.uint64)
    order = np.argsort(ak, kind="stable")
    sites = slabs.unpack_corners(ac[order]).astype(np.float32) - 0.5
    # every rank must hold the same numbered list, and it must be the reference's
    digest = torch.tensor([int(ak[order].sum() % (2 ** 62)), len(ak)], dtype=torch.int64)
    all_d = [torch.zeros_like(digest) for _ in range(world)]
    dist.all_gather(all_d, digest)
    same = all(torch.equal(all_d[0], d) for d in all_d)
    full = synth.assembly(n, count=8)
    want = ob.extract_sites(ob.classify_grid(full))
    # closest ids of the rank's own planes from the host core, against the whole-grid oracle
    ids, d2 = hc.closest_grid(sites, n, n, n, z0, z1)
    o_ids, o_d2 = ob.closest_grid(want, n, n, n, z0, z1)
    q.put((rank, same, bool(np.array_equal(sites, want)), bool(np.array_equal(ids, o_ids) and np.array_equal(d2, o_d2))))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_site_exchange_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(r[0] for r in res) == [0, 1]
    assert all(r[1] and r[2] and r[3] for r in res), res
