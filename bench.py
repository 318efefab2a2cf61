#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric: grid vertices/s through classify + closest-site + measures.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME] [--grid NX,NY,NZ] [--weak]

One "step" = one pass of the hot path (classify -> boundary samples -> closest site per grid vertex
-> cell measures) over the synthetic volume, which is already resident in HBM when the timed region
starts.  Workload, by default the SAME for every N so that the 1/2/4/8-GPU numbers form one STRONG-scaling
curve: BASELINE configs[3], assembly1024 (1024^3 vertices, 2.7e6 boundary samples; 75 GB resident on one
B200), cut into N z-slabs, one process per GPU, the site records exchanged over peer memory (NVLink).  At
N=1 the line also carries "at_512": the same measurements on configs[2], twist512 -- the 512^3
configuration the metric's single-GPU target is quoted on.  `--weak` keeps 512^3 vertices per GPU instead
(512^3 / 512x512x1024 / 512x1024x1024 / 1024^3), `--workload twist512` etc. picks one explicitly.
Rank 0 prints ONE JSON line.  `--impl reference` times the reference's own CPU operators instead
(oracle/_ref, bounded sample) and prints the same line with "impl": "reference".
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "grid vertices/sec (classify+closest-pt+measure)"
UNIT = "vertices/s"
# algorithmic bytes per grid vertex, SURVEY.md section 8(d) (stated again in DESIGN.md)
BYTES_CLASSIFY, BYTES_CLOSEST, BYTES_MEASURES = 5, 8, 32
BYTES_PIPELINE = BYTES_CLASSIFY + BYTES_CLOSEST + BYTES_MEASURES  # 45
STAGE_OF = {"classify_f32": "classify", "classify_i8": "classify", "edt_pass_z": "closest", "edt_pass_x": "closest",
            "edt_pass_y": "closest", "cell_measures": "measures"}
STAGE_BYTES = {"classify": BYTES_CLASSIFY, "closest": BYTES_CLOSEST, "measures": BYTES_MEASURES}
WEAK_GRIDS = {1: (512, 512, 512), 2: (512, 512, 1024), 4: (512, 1024, 1024), 8: (1024, 1024, 1024)}
STRONG = ("assembly", (1024, 1024, 1024))  # BASELINE configs[3]
AT_512 = ("twist", (512, 512, 512))        # BASELINE configs[2]


def workload_for(n_gpus, args):
    """-> (family, grid, scaling)"""
    if args.grid:
        g = tuple(int(v) for v in args.grid.split(","))
        return args.workload or "assembly", g, "strong"
    if args.workload:
        fam = args.workload.rstrip("0123456789")
        side = int(args.workload[len(fam):] or 512)
        return fam, (side, side, side), "strong"
    if args.weak:
        if n_gpus == 1:
            return "twist", WEAK_GRIDS[1], "weak"
        return "assembly", WEAK_GRIDS.get(n_gpus, (512, 512, 512 * n_gpus)), "weak"
    return STRONG[0], STRONG[1], "strong"


def make_planes(fam, grid, z0, z1):
    from voxel_ma_b200 import synth
    nx, ny, nz = grid
    if fam in ("assembly", "stress"):
        return synth.make(fam, grid if not (nx == ny == nz) else nx, z0=z0, z1=z1)
    assert nx == ny == nz, "only the assembly family has non-cubic grids"
    return synth.make(fam, nx, z0=z0, z1=z1)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(device), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "20"], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [l.split(",") for l in open(self.f.name).read().strip().splitlines() if l.count(",") >= 8]
        os.unlink(self.f.name)
        if not rows:
            return out
        sm = [float(r[1]) for r in rows if r[1].strip().replace(".", "").isdigit()]
        mx = [float(r[2]) for r in rows if r[2].strip().replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any("Active" in r[5 + k] and "Not" not in r[5 + k] for r in rows)]
        out.update(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None, reasons=reasons,
                   samples=len(rows))
        return out


class quiet_stdout:
    """The reference prints progress to stdout (cout/printf); bench.py must print ONE JSON line, so
    fd 1 is pointed at stderr while the CPU legs run."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)

    def __exit__(self, *a):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        os.close(self.saved)


# ------------------------------------------------------------------------------------------------
# the reference's CPU operators on a bounded sample of the same workload
# ------------------------------------------------------------------------------------------------
def _ann_worker(args):
    sites, q = args
    from oracle import bindings as ob
    ob.ref_ann(sites, q, brute=False)
    return ob.ref().ref_last_seconds()


def _sites_worker(block):
    from oracle import bindings as ob
    ob.ref_extract_sites(block)
    return ob.ref().ref_last_seconds()


_SITE_CACHE = {}


def full_site_set(fam, grid):
    """The boundary samples of the WHOLE grid, in the reference's order (oracle/oracle.c port of extractBoundaryVts,
    classified slab by slab); not timed -- the ANN leg needs the true site set at every size.  Cached per workload."""
    from oracle import bindings as ob
    key = (fam, grid)
    if key not in _SITE_CACHE:
        nx, ny, nz = grid
        inside = np.empty((nz, ny, nx), np.uint8)
        for z in range(0, nz, 128):
            inside[z:z + 128] = ob.classify_grid(make_planes(fam, grid, z, min(z + 128, nz)))
        _SITE_CACHE[key] = ob.extract_sites(inside)
    return _SITE_CACHE[key]


def cpu_reference_rate(fam, grid, budget_s=20.0, cores=None):
    """vertices/s of the reference's CPU operators for this path, all host cores, on a BOUNDED SAMPLE of the workload;
    the whole-grid figure is an extrapolation (the per-vertex costs of the three stages are measured on their samples
    and summed, exactly as they would add up in a whole-grid run) and is labelled as such.

    A  classify + boundary samples: the reference's Surfacer::extractBoundaryVts (it calls voxTaggedAsInside 7x per
       voxel) on one sub-block per core, through oracle/_ref/libvoxref.so;
    B  closest: the reference's ANNkd_tree build + annkSearch(k=1, eps=0) over the FULL site set of the grid, one forked
       process per core (ANN keeps search state in globals), uniformly random grid vertices as queries;
    C  measures: the lambdaForFace dictionary -- a PORT (oracle/oracle.c, OpenMP): the reference evaluates lambdaForFace
       on a Voronoi complex, the dense iteration space is the builder's; ids are synthetic but
       distinct across neighbours so that every lambda is really evaluated."""
    import multiprocessing as mp

    from oracle import bindings as ob
    kind = "reference" if ob.have_ref() else "port"
    cores = cores or os.cpu_count() or 1
    nx, ny, nz = grid
    nvert = nx * ny * nz
    t_total0 = time.perf_counter()
    sites = full_site_set(fam, grid)
    nsites = len(sites)
    # a slab through the middle of the object, thick enough to hold the sub-blocks
    zs = min(nz, 64)
    slab0 = max(0, nz // 2 - zs // 2)
    planes = make_planes(fam, grid, slab0, slab0 + zs)
    ctx = mp.get_context("fork")
    # --- stage A: site extraction rate (voxels/s), one sub-block per core
    bs = min(ny, 128)
    blocks = []
    for k in range(cores):
        y0 = (k * bs) % max(ny - bs + 1, 1)
        x0 = ((k * bs) // max(ny - bs + 1, 1) * bs) % max(nx - bs + 1, 1)
        blocks.append(np.ascontiguousarray(planes[:, y0:y0 + bs, x0:x0 + bs]))
    t0 = time.perf_counter()
    if kind == "reference":
        with ctx.Pool(cores) as pool:
            secs = pool.map(_sites_worker, blocks)
        wall_a = max(secs)
    else:
        for b in blocks:
            ob.extract_sites(ob.classify_grid(b))
        wall_a = (time.perf_counter() - t0) / cores
    nvox_a = sum(b.size for b in blocks)
    rate_a = nvox_a / max(wall_a, 1e-9)
    # --- stage B: ANN 1-NN per grid vertex; size the sample from a short probe
    rng = np.random.default_rng(7)

    def queries(m):
        return np.stack([rng.integers(0, nx, m), rng.integers(0, ny, m), rng.integers(0, nz, m)], -1).astype(np.float64)
    s64 = sites.astype(np.float64)
    if kind == "reference" and nsites > 0:
        probe = 4000
        t_probe = _ann_worker((s64, queries(probe)))
        per_q = max(t_probe / probe, 1e-7)
        m_per_core = int(min(max(budget_s * 0.6 / per_q, 2000), 3_000_000))
        with ctx.Pool(cores) as pool:
            secs = pool.map(_ann_worker, [(s64, queries(m_per_core)) for _ in range(cores)])
        rate_b = cores * m_per_core / max(max(secs), 1e-9)
        nq = cores * m_per_core
    else:
        nq = 2000
        t0 = time.perf_counter()
        ob.closest_points(s64, queries(nq)) if nsites else None
        rate_b = nq / max(time.perf_counter() - t0, 1e-9)
    # --- stage C: measures on the slab (C port of the dictionary; OpenMP over the cores)
    ins = ob.classify_grid(planes)
    ids = ((np.arange(planes.size, dtype=np.int64) * 7919) % max(nsites, 1)).astype(np.int32).reshape(planes.shape)
    t0 = time.perf_counter()
    if nsites:
        ob.cell_measures_grid(sites, ids, ins, nx, ny, planes.shape[0], 0, planes.shape[0] - 1)
    nv_c = planes[:-1].size
    rate_c = nv_c / max(time.perf_counter() - t0, 1e-9)
    value = 1.0 / (1.0 / rate_a + 1.0 / rate_b + 1.0 / rate_c)
    sample = (f"{fam} {nx}x{ny}x{nz}, {nsites} sites (the full site set); A: reference extractBoundaryVts on {cores} sub-blocks of "
              f"{zs}x{bs}x{bs} voxels ({rate_a:.3g} voxels/s); B: reference ANN kd-tree build + {nq} annkSearch(k=1,eps=0) queries at "
              f"random grid vertices, one forked process per core ({rate_b:.3g} q/s); C: lambda dictionary PORT on {planes.shape[0]-1} "
              f"planes ({rate_c:.3g} v/s); per-vertex costs summed and extrapolated to the whole grid; "
              f"{time.perf_counter()-t_total0:.1f}s of wall")
    return {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample, "extrapolated": True,
            "sampled_fraction": {"A_extract": nvox_a / nvert, "B_closest": nq / nvert, "C_measures": nv_c / nvert},
            "stage_kinds": {"A_extract": kind, "B_closest": kind, "C_measures": "port"}}


def cpu_reference_full(fam="torus", n=256, cores=None):
    """A MEASURED (not extrapolated) CPU run of the whole path at BASELINE configs[1], torus256: the reference's
    extractBoundaryVts over the whole volume (one thread: it is not parallel), the reference's ANN kd-tree at EVERY
    grid vertex (forked workers on all cores), the lambda dictionary port on the whole grid.  ~10 s on 16 cores."""
    from oracle import bindings as ob
    from voxel_ma_b200 import synth
    cores = cores or os.cpu_count() or 1
    vol = synth.make(fam, n)
    nz, ny, nx = vol.shape
    t0 = time.perf_counter()
    if ob.have_ref():
        sites = ob.ref_extract_sites(vol)
        t_a = ob.ref().ref_last_seconds()
        kind = "reference"
    else:
        sites = ob.extract_sites(ob.classify_grid(vol))
        t_a = time.perf_counter() - t0
        kind = "port"
    t1 = time.perf_counter()
    if ob.have_ref():
        ids, _, _ = ob.ref_ann_grid(sites, nx, ny, 0, nz, workers=cores)
    else:
        ids, _ = ob.closest_grid(sites, nx, ny, nz)
    t_b = time.perf_counter() - t1
    ins = ob.classify_grid(vol)
    t2 = time.perf_counter()
    ob.cell_measures_grid(sites, np.ascontiguousarray(ids), ins, nx, ny, nz)
    t_c = time.perf_counter() - t2
    tot = t_a + t_b + t_c
    return {"workload": f"{fam}{n}", "vertices": nx * ny * nz, "sites": int(len(sites)), "cores": cores, "kind": kind,
            "seconds": {"A_extract_1thread": t_a, "B_ann_every_vertex": t_b, "C_measures_port": t_c, "total": tot},
            "value": nx * ny * nz / tot, "unit": UNIT, "extrapolated": False}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    fam, grid, scaling = workload_for(args.gpus, args)
    vals = []
    t0 = time.perf_counter()
    cb = None
    with quiet_stdout():
        full_site_set(fam, grid)  # untimed set-up, shared by the repeats
        for _ in range(max(1, min(args.steps, 3))):  # each "step" = one bounded sample; the CPU path has no warm-up state
            cb = cpu_reference_rate(fam, grid, budget_s=12.0)
            vals.append(cb["value"])
        measured = None
        try:
            measured = cpu_reference_full()
        except Exception as ex:  # the second point must not cost the first
            measured = {"error": repr(ex)}
    v = float(np.median(vals))
    cb["value"] = v
    nvert = grid[0] * grid[1] * grid[2]
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": nvert / v * 1e3, "higher_is_better": True, "scaling": scaling,
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "extrapolated": True,
        "config": {"workload": f"{fam}{grid[0]}x{grid[1]}x{grid[2]}", "note": "reference CPU operators (Surfacer + ANN kd-tree; "
                   "lambda dictionary port) on a bounded sample of the workload; value and ms_per_step are the whole-grid "
                   "EXTRAPOLATION of the sampled per-vertex costs (cpu_baseline.sampled_fraction); cpu_reference_measured is a full, "
                   "un-extrapolated run of the same three stages at torus256"},
        "cpu_baseline": cb,
        "cpu_reference_measured": measured,
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "wall_s": time.perf_counter() - t0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
def bind_to_gpu_numa(local):
    """Pin this rank's threads (and, by first touch, its pinned staging buffers) to the NUMA node its GPU hangs off:
    with 8 ranks all left on node 0 every upload of the other socket's GPUs crosses the inter-socket link.
    Returns a short description for the JSON line; never fails the run."""
    try:
        import torch
        pr = torch.cuda.get_device_properties(local)
        bdf = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read())
        if node < 0:
            return {"gpu": bdf, "numa_node": node, "bound": False}
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = cpus & os.sched_getaffinity(0)
        if allowed:
            os.sched_setaffinity(0, allowed)
        return {"gpu": bdf, "numa_node": node, "bound": bool(allowed), "cpus": len(allowed)}
    except Exception as ex:
        return {"bound": False, "why": repr(ex)[:120]}


def load_traffic(workload, kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum of one whole-slab launch of `kernel`, from the committed ncu capture"""
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        for entry in tj["captures"]:
            if entry["workload"] == workload and kernel in entry["kernels"]:
                k = entry["kernels"][kernel]
                return (k["dram_read_gb"] + k["dram_write_gb"]) * 1e9, entry["source"], k.get("sm_throughput_pct")
    except Exception:
        pass
    return None, None, None


def points_path(ctx, grid, args, nq=4_000_000):
    """The arbitrary-point query path (vc_closest_points).  Two query sets over this workload's sites: (1) the shape the
    reference's own callers have -- points INSIDE the solid (Voronoi vertices kept by tagVert, medial-axis vertices,
    src/voxelapps.cpp:184-228, src/exporters.cpp:629-636): random inside grid vertices jittered by +-0.5 voxel; (2) the
    worst case for an expanding-shell search: uniformly random points of the whole box (most of them far from any site)."""
    nx, ny, nz = grid
    rng = np.random.default_rng(11)
    ctx.run_dense()  # the end-to-end legs before this one leave no dense planes behind
    vert = ctx.compact_records()[0].astype(np.int64)
    pick = vert[rng.integers(0, len(vert), nq)]
    q_in = np.stack([pick % nx, (pick // nx) % ny, pick // (nx * ny)], -1).astype(np.float64) + rng.uniform(-0.5, 0.5, (nq, 3))
    q_box = np.stack([rng.uniform(-0.5, nx - 0.5, nq), rng.uniform(-0.5, ny - 0.5, nq), rng.uniform(-0.5, nz - 0.5, nq)], -1)
    del vert, pick
    ctx.closest_points(q_in[:4096])  # builds the cell list of the resident sites (lazily, once per site set)
    out = {"sites": ctx.num_sites(), "note": "vc_closest_points from host arrays (copies inside the timed call): queries sorted by "
           "cell on the device, exact expanding-shell search over the cell list, double precision, ties to the lowest id"}
    for name, q in (("inside_queries", q_in), ("box_queries", q_box)):
        ctx.closest_points(q)
        ts = []
        for _ in range(3):
            t0 = time.perf_counter()
            ids, d2 = ctx.closest_points(q)
            ts.append(time.perf_counter() - t0)
        ctx.profile(True)
        ctx.profile_reset()
        ctx.closest_points(q)
        prof = {k: round(v["ms"], 4) for k, v in ctx.profile_report().items()}
        ctx.profile(False)
        rec = {"queries": nq, "gpu_q_per_s": nq / float(np.median(ts)), "gpu_ms_per_call": float(np.median(ts)) * 1e3,
               "gpu_kernels_ms": prof, "h2d_bytes": int(q.nbytes), "d2h_bytes": int(ids.nbytes + d2.nbytes),
               "mean_distance": float(np.sqrt(d2).mean())}
        if not args.no_cpu_baseline:
            import multiprocessing as mp
            from oracle import bindings as ob
            if ob.have_ref():
                cores = os.cpu_count() or 1
                sites = ctx.get_sites().astype(np.float64)
                m = 20000
                with quiet_stdout():
                    with mp.get_context("fork").Pool(cores) as pool:
                        secs = pool.map(_ann_worker, [(sites, q[k * m:(k + 1) * m]) for k in range(cores)])
                    a_id, a_d2 = ob.ref_ann(sites, q[:m])
                rec["cpu_ann"] = {"q_per_s": cores * m / max(secs), "cores": cores, "queries": cores * m, "kind": "reference",
                                  "note": "ANNkd_tree build + annkSearch(k=1, eps=0), one forked process per core"}
                # same answers: distance exactly, id up to ANN's traversal-dependent choice among equidistant sites
                rec["agrees_with_ann"] = bool(np.array_equal(a_d2, d2[:m]) and ((ids[:m] == a_id) | (ids[:m] < a_id)).all())
        out[name] = rec
    return out


def measure(args, fam, grid, scaling, dist, rank, world, local, sub=False):
    """One workload on this process group.  Returns the JSON line (rank 0) or None.  sub=True: the condensed record of
    the secondary N=1 workload ("at_512"): value, e2e and rooflines only."""
    import torch

    from voxel_ma_b200 import api, slabs

    nx, ny, nz = grid
    z0, z1 = slabs.slab_bounds(nz, world, rank)
    lo, hi = slabs.resident_planes(z0, z1, nz)
    planes = make_planes(fam, grid, lo, hi)
    heights = "equal"
    if world > 1 and (args.balance > 0 or args.align > 1):
        # slab heights from a per-plane work estimate 1 + c * inside fraction (the measures of a plane cost more where
        # it cuts solids), cuts snapped so that a slab's planes + its halo plane fill whole warps of pass X
        # (slabs.aligned_bounds); every rank counts its own planes, the counts are gathered
        frac = (planes[z0 - lo:z1 - lo] > 0).reshape(z1 - z0, -1).mean(axis=1)
        per = [None] * world
        dist.all_gather_object(per, (z0, frac.astype(np.float64)))
        full = np.zeros(nz)
        for zz, f in per:
            full[zz:zz + len(f)] = f
        w = 1.0 + max(args.balance, 0.0) * full
        z0, z1 = (slabs.aligned_bounds(w, world, args.align) if args.align > 1 else slabs.balanced_bounds(w, world))[rank]
        lo, hi = slabs.resident_planes(z0, z1, nz)
        planes = make_planes(fam, grid, lo, hi)
        heights = f"weights 1 + {args.balance} x inside fraction of the plane, cuts aligned to {args.align} planes incl. the halo plane"
        print(f"[rank {rank}] slab [{z0},{z1}) = {z1 - z0} planes", file=sys.stderr)
    ctx = api.Context(local)
    ctx.set_grid(nx, ny, nz, z0, z1)
    # pinned staging of the slab (the e2e leg copies from here every step)
    pin_vol = api.PinnedArray(planes.shape, np.float32)
    pin_vol.array[...] = planes
    del planes
    ctx.upload_volume(pin_vol.array, zlo=lo)
    nv_local = nx * ny * (z1 - z0)
    nv_total = nx * ny * nz
    stream = torch.cuda.ExternalStream(ctx.stream(), device=torch.device("cuda", local))
    state = {"nsites": 0}

    def step():
        if world == 1:
            state["nsites"] = ctx.run_dense()
            return
        ctx.classify_grid(fetch=False)
        if state.get("peers") is not None:  # records stored straight into every rank's memory over NVLink
            state["nsites"] = state["peers"].exchange()
            ctx.closest_and_measures()
            return
        n = ctx.sites_detect_local()
        state["nlocal"] = n
        keys = torch.empty(max(n, 1), dtype=torch.int64, device="cuda")
        corners = torch.empty(max(n, 1), dtype=torch.int64, device="cuda")
        ctx.sites_export_local(keys.data_ptr(), corners.data_ptr())
        ak, ac = slabs.exchange_site_records(keys[:n], corners[:n])
        torch.cuda.current_stream().synchronize()
        ctx.sites_import_global(ak.data_ptr(), ac.data_ptr(), ak.numel())
        state["nsites"] = ak.numel()
        ctx.closest_and_measures()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        ctx.synchronize()

    # ---- set-up steps (not warm-up: first-touch allocations; for N>1 the cross-check of the two exchange paths)
    if world > 1 and args.exchange == "peers":
        # one step over NCCL first: it sizes the receive regions and is the cross-check of the peer path
        step()
        ctx.synchronize()
        ref_sites = ctx.get_sites()
        t = torch.tensor([state["nlocal"]], dtype=torch.int64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        state["peers"] = slabs.PeerExchange(ctx, 2 * int(t[0]) + 4096)
        step()
        ctx.synchronize()
        if not np.array_equal(ctx.get_sites(), ref_sites):
            raise SystemExit("peer-memory exchange and NCCL all-gather disagree on the site numbering")
        del ref_sites
    else:
        step()
    # ---- W warm-up steps exactly as asked (the timing rules want W >= 3: the default; a smaller W is the caller's choice)
    for _ in range(args.warmup):
        step()
    # ---- timed region: exactly K steps, CUDA events on the stream the kernels are launched on
    barrier()
    sampler = ClockSampler(local) if rank == 0 and not sub else None
    launches0 = ctx.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    launches = (ctx.launch_count() - launches0) // max(args.steps, 1)
    # ---- same K steps again with every launch bracketed by events: per-kernel durations (roofline)
    ctx.profile(True)
    ctx.profile_reset()
    barrier()
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p0.record(stream)
    for _ in range(args.steps):
        step()
    p1.record(stream)
    barrier()
    ms_prof = p0.elapsed_time(p1)
    prof = ctx.profile_report()
    ctx.profile(False)
    if os.environ.get("VC_BENCH_RANKS"):  # development aid: every rank's own kernel times (load balance across slabs)
        print(f"[rank {rank}] ms/step {ms / args.steps:.3f} profiled {ms_prof / args.steps:.3f} " +
              " ".join(f"{k}={v['ms'] / args.steps:.3f}" for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])[:8]), file=sys.stderr)
    clocks = sampler.stop() if sampler else None
    if dist is not None:
        t = torch.tensor([ms, ms_prof], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_prof = float(t[0]), float(t[1])
    ms_per_step = ms / args.steps
    value = nv_total / (ms_per_step * 1e-3)

    def time_e2e(fn):
        """seconds per call: median of individually timed calls (each call ends with its own stream syncs), max over ranks"""
        reps = max(3, min(args.steps, 10))
        fn()
        barrier()
        ts = []
        for _ in range(reps):
            t0 = time.perf_counter()
            fn()
            ts.append(time.perf_counter() - t0)
        barrier()
        dt = float(np.median(ts))
        if dist is not None:
            t = torch.tensor([dt], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t[0])
        return dt

    def total(*vals):
        if dist is None:
            return [int(v) for v in vals]
        t = torch.tensor(list(vals), dtype=torch.int64, device="cuda")
        dist.all_reduce(t)
        return [int(v) for v in t]

    e2e, e2e_variants = None, {}
    if not args.no_e2e:
        (h2d,) = total(pin_vol.array.nbytes)
        s3 = (z1 - z0, ny, nx)
        small = nv_local <= 160_000_000  # the variants that return dense planes need 8..41 B of pinned memory per vertex
        if small:
            outs = {
                "inside": api.PinnedArray(s3, np.uint8), "id": api.PinnedArray(s3, np.int32), "d2": api.PinnedArray(s3, np.uint32),
                "edge3": api.PinnedArray((3,) + s3, np.float32), "face3": api.PinnedArray((3,) + s3, np.float32),
                "cube": api.PinnedArray(s3, np.float32), "radius": api.PinnedArray(s3, np.float32),
            }

            def step_e2e():
                if world == 1:
                    ctx.run_dense_host(pin_vol.array, outs["inside"].array, outs["id"].array, outs["d2"].array,
                                       outs["edge3"].array, outs["face3"].array, outs["cube"].array, outs["radius"].array)
                    return
                ctx.upload_volume(pin_vol.array, zlo=lo)
                step()
                for k, which in (("inside", api.ARR_INSIDE), ("id", api.ARR_ID), ("d2", api.ARR_D2X4), ("edge3", api.ARR_EDGE3),
                                 ("face3", api.ARR_FACE3), ("cube", api.ARR_CUBE), ("radius", api.ARR_RADIUS)):
                    ctx.lib.vc_download(ctx.h, which, outs[k].array.ctypes.data)

            dt = time_e2e(step_e2e)
            (d2h_planes,) = total(sum(o.array.nbytes for o in outs.values()))
            e2e_variants["all_planes"] = {"value": nv_total / dt, "ms_per_step": dt * 1e3, "h2d_bytes_per_step": h2d,
                                          "d2h_bytes_per_step": d2h_planes,
                                          "result": "inside u8, id i32, 4d2 u32, 7 lambda planes f32, radius f32 for every grid vertex"}
            del outs
        # ---- e2e, compact product (the headline): pinned host volume in; occupancy bit rows + one record
        # (vertex, id, 4d2, 7 lambda, radius) per INSIDE vertex out.  Measures anchored at outside vertices are 0
        # by definition, so nothing is lost; on a slab ctx the same call uploads the rank's
        # planes and exchanges the site records over peer memory.
        if world == 1 or state.get("peers") is not None:
            n_in = ctx.compact_count()
            cap = n_in + 1024
            wr = nx // 32 + 1
            # the record arrays are the rows of ONE pinned 11 x cap block (vert | id | 4d2 | 7 lambda | radius): a z chunk's
            # records then come back in a single 2-D copy
            rec = api.PinnedArray((11, cap), np.uint32)
            cb = {"bits": api.PinnedArray(((z1 - z0) * ny, wr), np.uint32), "vert": rec.array[0], "id": rec.array[1].view(np.int32),
                  "d2": rec.array[2], "lam": rec.array[3:10].view(np.float32), "rad": rec.array[10].view(np.float32)}

            def step_compact(dense=None):
                return ctx.run_dense_host_compact(pin_vol.array, cap, cb["bits"].array, cb["vert"], cb["id"], cb["d2"], cb["lam"], cb["rad"],
                                                  *(dense or (None, None)))

            e2e_s = time_e2e(step_compact)
            (d2h,) = total(int(cb["bits"].array.nbytes + n_in * 44))
            (n_in_total,) = total(n_in)
            # the e2e result really is the device result: spot-check the records against the resident planes
            got_n, _ = step_compact()
            step()  # the dense planes (the compact step computes the records of few inside vertices directly)
            ctx.synchronize()
            zc = min(z1 - z0, 128)  # a band of planes is enough for the cross-check (and bounded in host memory)
            ids_dev = ctx.download_planes(api.ARR_ID, z0, z0 + zc).ravel()
            cube_dev = ctx.download_planes(api.ARR_CUBE, z0, z0 + zc).ravel()
            v = cb["vert"][:got_n]
            m = v < ids_dev.size
            if got_n != n_in or not (np.array_equal(cb["id"][:got_n][m], ids_dev[v[m]]) and
                                     np.array_equal(cb["lam"][6, :got_n][m], cube_dev[v[m]])):
                raise SystemExit("compact e2e records disagree with the dense planes")
            del ids_dev, cube_dev
            result = "occupancy bit rows + (vertex u32, id i32, 4d2 u32, 7 lambda f32, radius f32) per inside vertex"
            e2e_variants["compact"] = {"value": nv_total / e2e_s, "ms_per_step": e2e_s * 1e3, "h2d_bytes_per_step": h2d,
                                       "d2h_bytes_per_step": d2h, "inside_vertices": n_in_total, "result": result}
            e2e = {"value": nv_total / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                   "ms_per_step": e2e_s * 1e3, "result": result}
            if world == 1 and small:
                # the same pass from an MRC mode 0 (signed byte) volume -- a format the reference's reader takes as well
                # (isosurface_tao/reader.h:235-239); the synthetic field is quantised to 1/16, which moves the surface a
                # little, so this is another volume, timed for what the narrower upload buys
                v8 = api.PinnedArray(pin_vol.array.shape, np.int8)
                v8.array[...] = np.clip(np.rint(pin_vol.array * 16.0), -127, 127).astype(np.int8)
                ctx.upload_volume(v8.array)
                ctx.classify_grid(fetch=False)
                cap8 = ctx.compact_count() + 1024
                rec8 = api.PinnedArray((11, cap8), np.uint32)
                dt8 = time_e2e(lambda: ctx.run_dense_host_compact(v8.array, cap8, cb["bits"].array, rec8.array[0], rec8.array[1].view(np.int32),
                                                                  rec8.array[2], rec8.array[3:10].view(np.float32),
                                                                  rec8.array[10].view(np.float32)))
                e2e_variants["compact_int8_volume"] = {"value": nv_total / dt8, "ms_per_step": dt8 * 1e3, "h2d_bytes_per_step": int(v8.array.nbytes),
                                                       "d2h_bytes_per_step": int(cb["bits"].array.nbytes + (cap8 - 1024) * 44),
                                                       "result": "compact product from an MRC mode 0 (int8) volume (field quantised to 1/16)"}
                del v8, rec8
                ctx.upload_volume(pin_vol.array, zlo=lo)
                ctx.classify_grid(fetch=False)
            if small:
                dense = (api.PinnedArray(s3, np.int32), api.PinnedArray(s3, np.uint32))
                dt = time_e2e(lambda: step_compact((dense[0].array, dense[1].array)))
                e2e_variants["compact_plus_dense_ids"] = {"value": nv_total / dt, "ms_per_step": dt * 1e3, "h2d_bytes_per_step": h2d,
                                                          "d2h_bytes_per_step": d2h + 8 * nv_total,
                                                          "result": "compact product + id i32 and 4d2 u32 planes for every grid vertex"}
                del dense
        elif "all_planes" in e2e_variants:
            ap = e2e_variants["all_planes"]
            e2e = {"value": ap["value"], "unit": UNIT, "h2d_bytes_per_step": ap["h2d_bytes_per_step"],
                   "d2h_bytes_per_step": ap["d2h_bytes_per_step"], "ms_per_step": ap["ms_per_step"], "result": ap["result"]}

    line = None
    if rank == 0:
        peak, peak_src = peaks()
        steps = args.steps
        wname = f"{fam}{nx}x{ny}x{nz}"
        kern = {k: {"ms_per_launch": v["ms"] / max(v["launches"], 1), "launches_per_step": v["launches"] / steps,
                    "ms_per_step": v["ms"] / steps} for k, v in prof.items()}
        # roofline on SURVEY 8(d)'s algorithmic bytes: 5 / 8 / 32 B per grid vertex for classify / closest (passes Z+X+Y
        # together) / measures, x the vertices this rank's launches process, / the event-timed duration
        stage_ms = {}
        for k, v in kern.items():
            if k in STAGE_OF:
                stage_ms[STAGE_OF[k]] = stage_ms.get(STAGE_OF[k], 0.0) + v["ms_per_step"]
        stages = {st: {"bytes_per_vertex": STAGE_BYTES[st], "ms_per_step": t, "achieved": STAGE_BYTES[st] * nv_local / (t * 1e-3) / 1e9,
                       "frac": STAGE_BYTES[st] * nv_local / (t * 1e-3) / 1e9 / peak} for st, t in stage_ms.items() if t > 0}
        roof = None
        hot = [k for k in kern if k in STAGE_OF]
        if hot:
            dom = max(hot, key=lambda k: kern[k]["ms_per_step"])
            st = STAGE_OF[dom]
            nbytes = STAGE_BYTES[st] * nv_local  # the whole stage's algorithmic bytes are charged to its dominant kernel
            ach = nbytes / (kern[dom]["ms_per_step"] * 1e-3) / 1e9
            traffic, traffic_src, sm_pct = load_traffic(wname if world == 1 else f"{wname}/slab{world}", dom)
            roof = {"kernel": dom, "stage": st, "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                    "traffic": traffic, "traffic_source": traffic_src, "sm_throughput_pct_ncu": sm_pct, "peak_source": peak_src,
                    "bytes_per_launch": nbytes, "ms_per_launch": kern[dom]["ms_per_step"],
                    "share_of_step": kern[dom]["ms_per_step"] / (ms_prof / steps),
                    "stage_frac": stages[st]["frac"],
                    "note": f"achieved = SURVEY 8(d) bytes of the {st} stage ({STAGE_BYTES[st]} B/vertex x {nv_local} vertices) / this kernel's "
                            "event-timed duration per step; stage_frac divides the same bytes by the whole stage's kernels"}
        pipe_ach = BYTES_PIPELINE * nv_local / (ms_per_step * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "u64/f32",
            "data": "synthetic",
            "config": {"workload": wname, "sites": state["nsites"], "z_slabs": world,
                       "slab_heights": heights,
                       "vertices_per_gpu": nv_local,
                       "exchange": ("none (one slab)" if world == 1 else
                                    "peer memory: every rank stores its sorted run of records into every rank over NVLink, ids by ranking across the runs (vc_peer.cu)"
                                    if state.get("peers") is not None else "NCCL all-gather"), "l2": "inputs larger than L2 (no flush needed)",
                       "outputs": "inside u8, id i32, 4d2 u32, 7 lambda planes f32, radius f32 (resident in HBM)",
                       "e2e_result": e2e["result"] if e2e else None},
            "e2e": e2e,
            "e2e_variants": e2e_variants,
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": roof,
            "roofline_stages": stages,
            "roofline_pipeline": {"bound": "hbm", "bytes_per_vertex": BYTES_PIPELINE, "achieved": pipe_ach, "peak": peak,
                                  "unit": "GB/s", "frac": pipe_ach / peak, "note": "45 B/vertex (SURVEY 8d) x vertices / step time, per GPU"},
            "kernels": {k: round(v["ms_per_step"], 4) for k, v in sorted(kern.items(), key=lambda kv: -kv[1]["ms_per_step"])},
            "ms_per_step_profiled": ms_prof / steps,
        }
        if sub:
            line = {k: line[k] for k in ("value", "unit", "ms_per_step", "config", "e2e", "e2e_variants", "gpu_launches", "roofline",
                                         "roofline_stages", "roofline_pipeline", "kernels", "ms_per_step_profiled")}
        if world == 1 and fam == "twist" and not args.no_e2e:
            # stage 1' (new functionality, no reference counterpart): the same pass starting from the closed triangle
            # mesh of the twisted plate instead of a voxel volume -- parity classification, then sites / closest / measures
            try:
                from voxel_ma_b200 import synth
                mv, mt = synth.twist_mesh(nx)
                ctx.classify_mesh(mv, mt, fetch=False)
                ctx.run_dense()
                ts = []
                for _ in range(5):
                    t0 = time.perf_counter()
                    ctx.classify_mesh(mv, mt, fetch=False)
                    ns_mesh = ctx.run_dense()
                    ts.append(time.perf_counter() - t0)
                ctx.profile(True)
                ctx.profile_reset()
                ctx.classify_mesh(mv, mt, fetch=False)
                mprof = {k: round(v["ms"], 4) for k, v in ctx.profile_report().items() if k.startswith("mesh_")}
                ctx.profile(False)
                line["mesh_path"] = {"ms_per_step": float(np.median(ts)) * 1e3, "value": nv_total / float(np.median(ts)), "unit": UNIT,
                                     "triangles": int(len(mt)), "sites": int(ns_mesh), "classify_kernels_ms": mprof,
                                     "note": "vc_classify_mesh (warp-ballot parity over the surface triangles, mesh uploaded from the host "
                                             "each step) + sites + closest + measures; wall clock per step"}
            except Exception as ex:
                line["mesh_path"] = {"error": repr(ex)}
    if rank == 0 and world == 1 and not args.no_points and not sub:
        # the arbitrary-point query path (vc_closest_points: the drop-in for ANNkd_tree::annkSearch(k=1, eps=0) behind
        # voxelapps / estimateRadiiField): uniformly random float64 queries in the grid's box over this workload's sites,
        # host arrays in and out; next to it the reference's kd-tree on all host cores over a sample of the same queries
        try:
            line["points_path"] = points_path(ctx, grid, args)
        except Exception as ex:
            line["points_path"] = {"error": repr(ex)}
    barrier()
    if state.get("peers") is not None:
        state["peers"].close()
    ctx.close()
    del pin_vol
    return line


def run_ours(args):
    import torch

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    dist = None
    torch.cuda.set_device(local)
    numa = bind_to_gpu_numa(local) if not args.no_numa else {"bound": False, "why": "--no-numa"}
    # stdout must carry ONE line (rank 0's JSON): anything a library prints there meanwhile (NCCL's version banner
    # under NCCL_DEBUG=VERSION, for one) is sent to stderr until the line is ready
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(obj):
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        print(json.dumps(obj), flush=True)
        os.dup2(2, 1)

    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    fam, grid, scaling = workload_for(world, args)
    line = measure(args, fam, grid, scaling, dist, rank, world, local)
    if rank == 0:
        line["numa"] = numa
    default_workload = not (args.grid or args.workload or args.weak)
    if world == 1 and default_workload and not args.no_at512:
        # BASELINE configs[2], the 512^3 configuration the metric's single-GPU target (>= 1e10 vertices/s end to end) is quoted on
        line["at_512"] = measure(args, AT_512[0], AT_512[1], "strong", None, 0, 1, local, sub=True)
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            try:
                with quiet_stdout():
                    line["cpu_baseline"] = cpu_reference_rate(fam, grid, budget_s=15.0)
            except Exception as ex:  # the checker missing must not hide the GPU number
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "unavailable", "sample": repr(ex)}
        emit(line)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=None, help="sphere256 | torus256 | twist512 | assembly1024 ...")
    ap.add_argument("--grid", default=None, help="NX,NY,NZ (assembly family)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--weak", action="store_true", help="weak scaling instead: 512^3 vertices per GPU (twist512 at N=1, assembly family "
                    "512x512x1024 / 512x1024x1024 / 1024^3 at N=2/4/8)")
    ap.add_argument("--no-at512", action="store_true", help="N=1: skip the secondary twist512 measurement")
    ap.add_argument("--no-points", action="store_true", help="N=1: skip the arbitrary-point query leg (points_path)")
    ap.add_argument("--no-numa", action="store_true", help="do not bind the rank to its GPU's NUMA node")
    ap.add_argument("--balance", type=float, default=0.0, help="N>1: weight c of the per-plane inside fraction in the slab-height "
                    "estimate 1 + c*fraction (0 = equal weights)")
    ap.add_argument("--align", type=int, default=32, help="N>1: snap the slab cuts so that planes + halo plane of a slab are a multiple "
                    "of this many planes (32 = the planes one warp of pass X holds; 1 = no snapping)")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer legs")
    ap.add_argument("--exchange", default="peers", choices=["peers", "nccl"],
                    help="N>1: how the site records travel between ranks (peer-memory kernel stores, or NCCL all-gather)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
