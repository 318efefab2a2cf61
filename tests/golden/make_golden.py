"""Generate the golden fixtures under tests/golden/ from the REAL reference.

Run in the build container only (needs /root/reference and oracle/_ref built by
`make -f oracle/Makefile.ref`):   python tests/golden/make_golden.py

Everything written here is an OUTPUT OF THE UNMODIFIED REFERENCE CODE (Surfacer, SpaceConverter,
VoroInfo, ANN, MeasureForMA, the whole vol2ma pipeline incl. TetGen) on seeded synthetic inputs, plus
the ANN sample fixture transcribed from 3rdparty/ann/sample/{data.pts,query.pts,sample.save}.
The fixtures are small (a few hundred KB) and travel to the GPU box, where /root/reference does
not exist.
"""
import hashlib
import os
import re
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import bindings as ob  # noqa: E402
from voxel_ma_b200 import synth  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def ann_sample():
    """3rdparty/ann/sample: 20 data points, 10 queries (2-D) and the golden (index, distance) pairs
    of sample.save:22-51 (distances there are NON-squared, printed with 6 significant digits)."""
    d = os.path.join(REF, "3rdparty/ann/sample")
    data = np.loadtxt(os.path.join(d, "data.pts"))
    query = np.loadtxt(os.path.join(d, "query.pts"))
    txt = open(os.path.join(d, "sample.save")).read()
    nn = re.findall(r"^\s*0\s+(\d+)\s+([0-9.eE+-]+)\s*$", txt, flags=re.M)
    idx = np.array([int(a) for a, _ in nn], np.int32)
    dist = np.array([float(b) for _, b in nn], np.float64)
    assert len(idx) == len(query) == 10 and len(data) == 20
    np.savez(os.path.join(OUT, "ann_sample.npz"), data=data, query=query, nn_idx=idx, nn_dist=dist)


def ann_tests():
    """3rdparty/ann/test/test{1,2}: the .save logs record average_error = rank_error = 0 against
    ANNbruteForce at eps = 0; here the reference's kd-tree and brute-force answers (k=1) for those very
    data/query files are stored (test2 is 8-D)."""
    d = os.path.join(REF, "3rdparty/ann/test")
    for t in ("test1", "test2"):
        data = np.loadtxt(os.path.join(d, f"{t}-data.pts"))
        query = np.loadtxt(os.path.join(d, f"{t}-query.pts"))
        ki, kd = ob.ref_ann(data, query, brute=False)
        bi, bd = ob.ref_ann(data, query, brute=True)
        assert np.array_equal(kd, bd)
        np.savez_compressed(os.path.join(OUT, f"ann_{t}.npz"), data=data, query=query, kd_idx=ki, kd_d2=kd,
                            brute_idx=bi, brute_d2=bd)


def sites_and_closest():
    """Surfacer::extractBoundaryVts + voxTaggedAsInside + ANN (kd-tree d2, brute-force ids) on small
    seeded volumes; larger ones are stored as counts + hashes."""
    rng = np.random.default_rng(20181)
    vols = {
        "sphere16": synth.sphere(16),
        "twist20": synth.twist(20),
        "noise_9x14x11": rng.standard_normal((9, 14, 11)).astype(np.float32),
    }
    out = {}
    for k, v in vols.items():
        nz, ny, nx = v.shape
        inside = ob.ref_classify_grid(v)
        sites = ob.ref_extract_sites(v)
        zz, yy, xx = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
        q = np.stack([xx, yy, zz], -1).reshape(-1, 3).astype(np.float64)
        bi, bd = ob.ref_ann(sites, q, brute=True)
        ki, kd = ob.ref_ann(sites, q, brute=False)
        assert np.array_equal(bd, kd)
        out[k + "_vol"] = v
        out[k + "_inside"] = inside
        out[k + "_sites"] = sites
        out[k + "_brute_idx"] = bi.reshape(nz, ny, nx)
        out[k + "_d2"] = bd.reshape(nz, ny, nx)
        out[k + "_kd_idx"] = ki.reshape(nz, ny, nx)
    np.savez_compressed(os.path.join(OUT, "dense_small.npz"), **out)
    big = {}
    for n in (32, 64, 128):
        v = synth.sphere(n)
        sites = ob.ref_extract_sites(v)
        big[f"sphere{n}_nsites"] = len(sites)
        big[f"sphere{n}_sites_sha256"] = sha(sites)
        big[f"sphere{n}_inside_sha256"] = sha(ob.ref_classify_grid(v))
        big[f"sphere{n}_first8"] = sites[:8]
    np.savez(os.path.join(OUT, "sites_hashes.npz"), **big)


def pipeline():
    """The reference's whole vol2ma pipeline (computeVD with TetGen, preprocessVoro,
    extractInsideWithMeasure) on sphere(24): TetGen's Voronoi vertices + tagVert verdicts (a4),
    per-face site pairs + lambda (a7), per-vertex radii (a8), and V/E/F measure statistics."""
    v = synth.sphere(24)
    p = ob.ref_pipeline(v, preprocess=False)
    p2 = ob.ref_pipeline(v, preprocess=True)
    np.savez_compressed(
        os.path.join(OUT, "pipeline_sphere24.npz"), vol=v, sites=p["sites"], tet_vpts=p["tet_vpts"],
        tet_vtag=p["tet_vtag"], vts=p["vts"], radii=p["radii"], site_of_v=p["site_of_v"],
        face_sites=p["face_sites"], face_lambda=p["face_lambda"],
        counts_after_load=np.array([p["counts"][k] for k in ("vts", "edges", "faces")]),
        counts_after_merge=np.array([p2["counts"][k] for k in ("vts", "edges", "faces")]),
        v_msure=p2["v_msure"], e_msure=p2["e_msure"], f_msure=p2["f_msure"])


def fr_search():
    """annkFRSearch (a5, fixed radius; src/voxelapps.cpp:346-353) from the REAL ANNkd_tree: lattice sites
    of a small sphere (many equal distances; radii that hit a distance exactly, the inclusive case) and
    random points.  Stored per query: count and the (id, d2) rows sorted by (d2, id)."""
    sites = ob.ref_extract_sites(synth.sphere(16)).astype(np.float64)
    rng = np.random.default_rng(4242)
    q = np.concatenate([rng.uniform(-1, 17, (150, 3)), np.floor(rng.uniform(0, 16, (150, 3)))])
    _, nn_d2 = ob.ref_ann(sites, q)
    # the reference's radius: nearest squared distance + eps (exactly-on-the-sphere ties), plus larger balls
    sq = np.concatenate([nn_d2[:100] + 1e-7, nn_d2[100:200], rng.uniform(0, 30, 100)])
    cnt, idx, d2 = ob.ref_ann_fr(sites, q, sq)
    rows_i, rows_d = [], []
    for i in range(len(q)):
        o = np.lexsort((idx[i, : cnt[i]], d2[i, : cnt[i]]))
        rows_i.append(idx[i, : cnt[i]][o])
        rows_d.append(d2[i, : cnt[i]][o])
    np.savez_compressed(os.path.join(OUT, "ann_fr_search.npz"), sites=sites, q=q, sq_rad=sq, count=cnt,
                        idx=np.concatenate(rows_i), d2=np.concatenate(rows_d))


def cli_sphere64():
    """BASELINE config 1: sphere64 through the unmodified main_voroUtility -md=vol2ma; output file
    hashes and the counts the CLI prints (26224/43250/17027 -> 8955/20138/11184,
    thinned 4605/0/9048)."""
    with tempfile.TemporaryDirectory() as d:
        synth.write_mrc(os.path.join(d, "sphere64.mrc"), synth.sphere(64))
        subprocess.check_call(["cp", os.path.join(ROOT, "oracle/_ref/cycle8.txt"), d])
        log = subprocess.run([ob.REF_CLI, "-md=vol2ma", "-fullOrPruned=2", "-tt=0.04", "sphere64.mrc", "out.ply"],
                             cwd=d, capture_output=True, text=True, check=True).stdout
        files = {}
        for f in ("out.ply", "out.r", "out_thinned0.04.ply", "out_thinned0.04.r"):
            files[f] = hashlib.sha256(open(os.path.join(d, f), "rb").read()).hexdigest()
    keep = [l for l in log.splitlines() if re.search(r"extracted sites|loaded voro size|after merging voro size|"
                                                     r"inside part extracted|# remaining|measure range|q size", l)]
    with open(os.path.join(OUT, "cli_sphere64.txt"), "w") as f:
        f.write("# main_voroUtility -md=vol2ma -fullOrPruned=2 -tt=0.04 sphere64.mrc out.ply  (unmodified reference)\n")
        for l in keep:
            f.write(l.strip() + "\n")
        for k, h in files.items():
            f.write(f"sha256 {k} {h}\n")


def cli_more():
    """SURVEY 8f-2 / 8f-3 through the unmodified CLI: `-md=r` (estimateRadiiField, trimesh KDtree) and
    `-dofuncmap=bt3` (assignScalarToSites -> match_voro_with_medialcurve, ANN) on the inputs of
    tests/cli_cases.py; output file hashes."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import cli_cases
    out = ["# unmodified reference CLI on the inputs of tests/cli_cases.py"]
    with tempfile.TemporaryDirectory() as d:
        args = cli_cases.write_radii_case(d, ob.ref_extract_sites(synth.torus(48)))
        subprocess.run([ob.REF_CLI, *args], cwd=d, capture_output=True, text=True, check=True)
        out.append("case radii " + " ".join(args))
        out.append("sha256 radii.txt " + hashlib.sha256(open(os.path.join(d, "radii.txt"), "rb").read()).hexdigest())
    with tempfile.TemporaryDirectory() as d:
        args = cli_cases.write_funcmap_case(d)
        subprocess.check_call(["cp", os.path.join(ROOT, "oracle/_ref/cycle8.txt"), d])
        subprocess.run([ob.REF_CLI, *args], cwd=d, capture_output=True, text=True, check=True)
        out.append("case funcmap " + " ".join(args))
        for f in ("out.bt3.msure", "dbg_mc_smoothed.msure", "out.ply", "out.r"):
            out.append(f"sha256 {f} " + hashlib.sha256(open(os.path.join(d, f), "rb").read()).hexdigest())
    open(os.path.join(OUT, "cli_more.txt"), "w").write("\n".join(out) + "\n")


def kdtree_f32():
    """trimesh::KDtree::closest_to_pt + trimesh::dist exactly as estimateRadiiField uses them (8f-2), from the
    REAL tree: lattice boundary points of torus(32), float queries, lattice queries (exact ties) and far
    queries, with a finite search limit (some queries find nothing)."""
    pts = ob.ref_extract_sites(synth.torus(32))
    rng = np.random.default_rng(515)
    q = np.concatenate([rng.uniform(0, 31, (1500, 3)), rng.integers(0, 32, (700, 3)), rng.uniform(-20, 50, (300, 3))]).astype(np.float32)
    lim = np.float32(150.0)
    idx, dist = ob.ref_kdtree_closest(pts, q, lim)
    np.savez_compressed(os.path.join(OUT, "kdtree_f32.npz"), pts=pts, q=q, max_d2=lim, found=(idx >= 0), dist=dist)


if __name__ == "__main__":
    if not (os.path.isdir(REF) and ob.have_ref()):
        sys.exit("needs /root/reference and oracle/_ref (make -f oracle/Makefile.ref)")
    ann_sample()
    ann_tests()
    sites_and_closest()
    pipeline()
    fr_search()
    cli_sphere64()
    cli_more()
    kdtree_f32()
    print("golden fixtures written to", OUT)
