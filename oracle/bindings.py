"""ctypes bindings for the CPU checker -- TEST INFRASTRUCTURE ONLY.

Two libraries:
  * ``oracle/_build/liboracle.so``  the plain-C restatement (oracle/oracle.c), built by ``build_oracle()``
  * ``oracle/_ref/libvoxref.so``    the UNMODIFIED reference compiled from /root/reference by
                                    oracle/Makefile.ref (present only where it was built; it travels
                                    to the GPU box as a prebuilt file)

Only tests/, ``__graft_entry__.smoke()`` and bench.py's cpu_baseline / ``--impl reference`` legs may
import this module.  Nothing under voxel_ma_b200/ does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
ORACLE_SO = os.path.join(HERE, "_build", "liboracle.so")
REF_SO = os.path.join(HERE, "_ref", "libvoxref.so")
REF_CLI = os.path.join(HERE, "_ref", "main_voroUtility")
REFERENCE_TREE = "/root/reference"

_f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")
_u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
_u32p = np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS")


def build_oracle(force: bool = False) -> str:
    """gcc-compile oracle/oracle.c (no -march, -ffp-contract=off)."""
    src = os.path.join(HERE, "oracle.c")
    if force or not os.path.exists(ORACLE_SO) or os.path.getmtime(ORACLE_SO) < os.path.getmtime(src):
        os.makedirs(os.path.dirname(ORACLE_SO), exist_ok=True)
        subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-fopenmp", "-shared", "-fPIC",
                               src, "-o", ORACLE_SO, "-lm"])
    return ORACLE_SO


def build_ref(jobs: int = 8) -> str | None:
    """Build oracle/_ref from /root/reference when that tree is present (this container only)."""
    if not os.path.isdir(REFERENCE_TREE):
        return REF_SO if os.path.exists(REF_SO) else None
    subprocess.check_call(["make", "-f", os.path.join("oracle", "Makefile.ref"), f"-j{jobs}"],
                          cwd=ROOT, stdout=subprocess.DEVNULL)
    return REF_SO


def have_ref() -> bool:
    return os.path.exists(REF_SO)


_oracle = None
_ref = None


def oracle():
    global _oracle
    if _oracle is None:
        lib = C.CDLL(build_oracle())
        lib.orc_classify_grid_f32.argtypes = [_f32p, C.c_int, C.c_int, C.c_int, _u8p]
        lib.orc_classify_grid_f64_zfast.argtypes = [_f64p, C.c_int, C.c_int, C.c_int, _u8p]
        lib.orc_classify_points.argtypes = [_u8p, C.c_int, C.c_int, C.c_int, _f32p, C.c_int64, C.c_void_p, _u8p]
        lib.orc_extract_sites.argtypes = [_u8p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int64]
        lib.orc_extract_sites.restype = C.c_int64
        lib.orc_closest_points.argtypes = [_f64p, C.c_int64, C.c_int, _f64p, C.c_int64, _i32p, C.c_void_p]
        lib.orc_closest_points_f32.argtypes = [_f32p, C.c_int64, _f32p, C.c_int64, C.c_float, _i32p, C.c_void_p]
        lib.orc_closest_grid.argtypes = [_f32p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                         _i32p, C.c_void_p, C.c_void_p]
        lib.orc_closest_grid_sample.argtypes = [_f32p, C.c_int64, _i32p, C.c_int64, _i32p, _u32p]
        lib.orc_cell_circum_angle_grid.argtypes = [_f32p, _i32p, _u8p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _f64p, _f64p]
        lib.orc_medial_quads.argtypes = [_f32p, _i32p, _u8p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int64,
                                         C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.orc_medial_quads.restype = C.c_int64
        lib.orc_face_lambda.argtypes = [_f32p, _i32p, C.c_int64, _f32p]
        lib.orc_vertex_radii.argtypes = [_f32p, _f32p, C.c_int64, _i32p, _f32p]
        lib.orc_segment_max.argtypes = [_i32p, _i32p, C.c_int64, _f32p, C.c_void_p, _f32p]
        lib.orc_cell_measures_grid.argtypes = [_f32p, _i32p, _u8p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                               _f32p, _f32p, _f32p, C.c_void_p]
        lib.orc_classify_mesh.argtypes = [_f32p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_int, C.c_int, C.c_int, _u8p]
        lib.orc_classify_mesh.restype = C.c_int
        lib.orc_radius_search.argtypes = [_f64p, C.c_int64, _f64p, C.c_int64, _f64p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.orc_ref_counts.argtypes = [_i32p, C.c_int64, C.c_int64, _i32p]
        lib.orc_simple_pairs.argtypes = [_i32p, _i32p, C.c_int64, _f32p, C.c_void_p, C.c_float, _i32p, _i32p, C.c_int64, _f32p,
                                         C.c_float, _i32p]
        lib.orc_simple_pairs.restype = C.c_int64
        _oracle = lib
    return _oracle


def ref():
    global _ref
    if _ref is None:
        if not have_ref():
            raise RuntimeError("oracle/_ref/libvoxref.so not built (make -f oracle/Makefile.ref)")
        lib = C.CDLL(REF_SO)
        lib.ref_set_data_dir.argtypes = [C.c_char_p]
        lib.ref_set_data_dir(os.path.join(HERE, "_ref").encode())
        lib.ref_last_seconds.restype = C.c_double
        lib.ref_classify_grid.argtypes = [_f64p, C.c_int, C.c_int, C.c_int, _u8p]
        lib.ref_tag_points.argtypes = [_f64p, C.c_int, C.c_int, C.c_int, _f32p, C.c_int64, _u8p]
        lib.ref_extract_sites.argtypes = [_f64p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int64]
        lib.ref_extract_sites.restype = C.c_int64
        for f in (lib.ref_ann_kd_search, lib.ref_ann_brute_search):
            f.argtypes = [_f64p, C.c_int, C.c_int, _f64p, C.c_int64, _i32p, _f64p]
        lib.ref_ann_kd_grid.argtypes = [_f64p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        lib.ref_ann_kd_fr_search.argtypes = [_f64p, C.c_int, C.c_int, _f64p, C.c_int64, _f64p, C.c_int,
                                             _i32p, C.c_void_p, C.c_void_p]
        lib.ref_kdtree_closest.argtypes = [_f32p, C.c_int64, _f32p, C.c_int64, C.c_float, _i32p, _f32p]
        lib.ref_lambda_for_face.argtypes = [_f32p, _f32p, C.c_int64, _f32p]
        lib.ref_pipeline_run.argtypes = [_f64p, C.c_int, C.c_int, C.c_int, C.c_int]
        lib.ref_pipeline_run.restype = C.c_void_p
        lib.ref_pipeline_free.argtypes = [C.c_void_p]
        lib.ref_pipeline_count.argtypes = [C.c_void_p, C.c_int]
        lib.ref_pipeline_count.restype = C.c_int64
        lib.ref_pipeline_sites.argtypes = [C.c_void_p, _f32p]
        lib.ref_pipeline_vts.argtypes = [C.c_void_p, _f32p, _f32p, _i32p]
        lib.ref_pipeline_faces.argtypes = [C.c_void_p, _i32p, _f32p]
        lib.ref_pipeline_tet_vpts.argtypes = [C.c_void_p, _f32p, _u8p]
        lib.ref_pipeline_measures.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        _ref = lib
    return _ref


# ---------------------------------------------------------------- numpy-level helpers (oracle.c)
def _dims(vol):
    nz, ny, nx = vol.shape
    return nx, ny, nz


def classify_grid(vol_xfast: np.ndarray) -> np.ndarray:
    v = np.ascontiguousarray(vol_xfast, np.float32)
    out = np.empty(v.shape, np.uint8)
    oracle().orc_classify_grid_f32(v, *_dims(v), out)
    return out


def classify_grid_f64_zfast(vol_zfast: np.ndarray, nx, ny, nz) -> np.ndarray:
    out = np.empty((nz, ny, nx), np.uint8)
    oracle().orc_classify_grid_f64_zfast(np.ascontiguousarray(vol_zfast, np.float64).ravel(), nx, ny, nz, out)
    return out


def classify_points(inside: np.ndarray, xyz: np.ndarray, M=None) -> np.ndarray:
    xyz = np.ascontiguousarray(xyz, np.float32).reshape(-1, 3)
    out = np.empty(len(xyz), np.uint8)
    m = None if M is None else np.ascontiguousarray(M, np.float64)
    oracle().orc_classify_points(np.ascontiguousarray(inside, np.uint8), *_dims(inside), xyz, len(xyz),
                                 None if m is None else m.ctypes.data, out)
    return out


def classify_mesh(verts, tris, nx, ny, nz, M=None):
    """1' (parity unpinned): even-odd classification of the grid from a closed triangle mesh.
    Returns (rc, inside[z][y][x])."""
    verts = np.ascontiguousarray(verts, np.float32).reshape(-1, 3)
    tris = np.ascontiguousarray(tris, np.uint32).reshape(-1, 3)
    m = None if M is None else np.ascontiguousarray(M, np.float64)
    out = np.empty((nz, ny, nx), np.uint8)
    rc = oracle().orc_classify_mesh(verts, len(verts), tris.ctypes.data, len(tris), None if m is None else m.ctypes.data,
                                    nx, ny, nz, out)
    return rc, out


def radius_search(sites, q, sq_rad, fetch=True):
    """orc_radius_search: counts, and (CSR) ids / squared distances of every site within sq_rad[i]."""
    s = np.ascontiguousarray(sites, np.float64).reshape(-1, 3)
    q = np.ascontiguousarray(q, np.float64).reshape(-1, 3)
    r = np.ascontiguousarray(np.broadcast_to(np.asarray(sq_rad, np.float64), (len(q),)))
    cnt = np.empty(len(q), np.int32)
    oracle().orc_radius_search(s, len(s), q, len(q), r, None, cnt.ctypes.data, None, None)
    if not fetch:
        return cnt
    off = np.zeros(len(q) + 1, np.int64)
    np.cumsum(cnt, out=off[1:])
    idx = np.empty(max(int(off[-1]), 1), np.int32)
    d2 = np.empty(max(int(off[-1]), 1), np.float64)
    oracle().orc_radius_search(s, len(s), q, len(q), r, off.ctypes.data, None, idx.ctypes.data, d2.ctypes.data)
    return cnt, off, idx[: off[-1]], d2[: off[-1]]


def ref_counts(idx, nbins):
    i = np.ascontiguousarray(idx, np.int32).ravel()
    out = np.empty(nbins, np.int32)
    oracle().orc_ref_counts(i, len(i), nbins, out)
    return out


def simple_pairs(edge_ref, edge_face0, face_measure, f_t, vert_ref, vert_edge0, edge_measure, l_t, face_to_remove=None):
    er, ef = (np.ascontiguousarray(a, np.int32) for a in (edge_ref, edge_face0))
    vr, ve = (np.ascontiguousarray(a, np.int32) for a in (vert_ref, vert_edge0))
    fm, em = (np.ascontiguousarray(a, np.float32) for a in (face_measure, edge_measure))
    tr = None if face_to_remove is None else np.ascontiguousarray(face_to_remove, np.uint8)
    out = np.empty((max(len(er) + len(vr), 1), 3), np.int32)
    n = oracle().orc_simple_pairs(er, ef, len(er), fm, None if tr is None else tr.ctypes.data, f_t, vr, ve, len(vr), em, l_t, out)
    return out[:n].copy()


def extract_sites(inside: np.ndarray) -> np.ndarray:
    ins = np.ascontiguousarray(inside, np.uint8)
    cap = 1 << 22  # one scan when the sites fit the first guess (the scan of a 1024^3 grid takes tens of seconds)
    out = np.empty((cap, 3), np.float32)
    n = oracle().orc_extract_sites(ins, *_dims(ins), out.ctypes.data, cap)
    if n > cap:
        out = np.empty((n, 3), np.float32)
        oracle().orc_extract_sites(ins, *_dims(ins), out.ctypes.data, n)
    return out[:n].copy()


def closest_points(sites: np.ndarray, q: np.ndarray):
    s = np.ascontiguousarray(sites, np.float64)
    q = np.ascontiguousarray(q, np.float64)
    dim = s.shape[1]
    idx = np.empty(len(q), np.int32)
    d2 = np.empty(len(q), np.float64)
    oracle().orc_closest_points(s, len(s), dim, q, len(q), idx, d2.ctypes.data)
    return idx, d2


def closest_points_f32(pts, q, max_d2=0.0):
    """float32 nearest point (trimesh KDtree semantics): (idx, d2), -1 where nothing within max_d2"""
    p = np.ascontiguousarray(pts, np.float32).reshape(-1, 3)
    q = np.ascontiguousarray(q, np.float32).reshape(-1, 3)
    idx = np.empty(len(q), np.int32)
    d2 = np.empty(len(q), np.float32)
    oracle().orc_closest_points_f32(p, len(p), q, len(q), float(max_d2), idx, d2.ctypes.data)
    return idx, d2


def closest_grid(sites_xyz: np.ndarray, nx, ny, nz, z0=0, z1=None, want_d2=False):
    z1 = nz if z1 is None else z1
    s = np.ascontiguousarray(sites_xyz, np.float32)
    ids = np.empty((z1 - z0, ny, nx), np.int32)
    d2x4 = np.empty((z1 - z0, ny, nx), np.uint32)
    d2 = np.empty((z1 - z0, ny, nx), np.float64) if want_d2 else None
    oracle().orc_closest_grid(s, len(s), nx, ny, nz, z0, z1, ids, d2x4.ctypes.data,
                              None if d2 is None else d2.ctypes.data)
    return (ids, d2x4, d2) if want_d2 else (ids, d2x4)


def closest_grid_sample(sites_xyz, q_xyz):
    """(lowest id at the minimum distance, 4*d2) at the integer vertices q_xyz [n,3] = (x, y, z): all sites scanned"""
    s = np.ascontiguousarray(sites_xyz, np.float32).reshape(-1, 3)
    q = np.ascontiguousarray(q_xyz, np.int32).reshape(-1, 3)
    ids = np.empty(len(q), np.int32)
    d2 = np.empty(len(q), np.uint32)
    oracle().orc_closest_grid_sample(s, len(s), q, len(q), ids, d2)
    return ids, d2


def cell_circum_angle_grid(sites_xyz, ids, inside, nx, ny, nz, z0=0, z1=None):
    """ids / inside hold planes [z0, min(z1+1, nz)); outputs [7][z1-z0][ny][nx] float64 (circumradius, object angle)"""
    z1 = nz if z1 is None else z1
    circ = np.empty((7, z1 - z0, ny, nx), np.float64)
    ang = np.empty_like(circ)
    oracle().orc_cell_circum_angle_grid(np.ascontiguousarray(sites_xyz, np.float32), np.ascontiguousarray(ids, np.int32),
                                        np.ascontiguousarray(inside, np.uint8), nx, ny, nz, z0, z1, circ, ang)
    return circ, ang


def medial_quads(sites_xyz, ids, inside, nx, ny, nz, z0=0, z1=None, zlo=0):
    """dual quads of the crossing grid edges anchored in planes [z0, z1); ids / inside hold planes [zlo, zlo + len)"""
    z1 = nz if z1 is None else z1
    s = np.ascontiguousarray(sites_xyz, np.float32)
    i = np.ascontiguousarray(ids, np.int32)
    f = np.ascontiguousarray(inside, np.uint8)
    zhi = zlo + i.shape[0]
    n = oracle().orc_medial_quads(s, i, f, nx, ny, nz, zlo, zhi, z0, z1, 0, None, None, None, None, None)
    m = max(int(n), 1)
    anchor, axis = np.empty(m, np.uint32), np.empty(m, np.uint8)
    a, b, lam = np.empty(m, np.int32), np.empty(m, np.int32), np.empty(m, np.float32)
    oracle().orc_medial_quads(s, i, f, nx, ny, nz, zlo, zhi, z0, z1, m, anchor.ctypes.data, axis.ctypes.data, a.ctypes.data,
                              b.ctypes.data, lam.ctypes.data)
    return anchor[:n], axis[:n], a[:n], b[:n], lam[:n]


def face_lambda(sites_xyz, site_pairs):
    p = np.ascontiguousarray(site_pairs, np.int32).reshape(-1, 2)
    out = np.empty(len(p), np.float32)
    oracle().orc_face_lambda(np.ascontiguousarray(sites_xyz, np.float32), p, len(p), out)
    return out


def vertex_radii(sites_xyz, v_xyz, site_of_v):
    v = np.ascontiguousarray(v_xyz, np.float32).reshape(-1, 3)
    out = np.empty(len(v), np.float32)
    oracle().orc_vertex_radii(np.ascontiguousarray(sites_xyz, np.float32), v, len(v),
                              np.ascontiguousarray(site_of_v, np.int32), out)
    return out


def segment_max(off, items, face_lambda_, face_valid=None):
    off = np.ascontiguousarray(off, np.int32)
    out = np.empty(len(off) - 1, np.float32)
    fv = None if face_valid is None else np.ascontiguousarray(face_valid, np.uint8)
    oracle().orc_segment_max(off, np.ascontiguousarray(items, np.int32), len(out),
                             np.ascontiguousarray(face_lambda_, np.float32),
                             None if fv is None else fv.ctypes.data, out)
    return out


def cell_measures_grid(sites_xyz, ids, inside, nx, ny, nz, z0=0, z1=None, want_radius=True):
    """ids / inside hold planes [z0, min(z1+1, nz)); outputs cover [z0, z1)."""
    z1 = nz if z1 is None else z1
    nzs = z1 - z0
    e = np.empty((3, nzs, ny, nx), np.float32)
    f = np.empty((3, nzs, ny, nx), np.float32)
    c = np.empty((nzs, ny, nx), np.float32)
    r = np.empty((nzs, ny, nx), np.float32) if want_radius else None
    oracle().orc_cell_measures_grid(np.ascontiguousarray(sites_xyz, np.float32),
                                    np.ascontiguousarray(ids, np.int32), np.ascontiguousarray(inside, np.uint8),
                                    nx, ny, nz, z0, z1, e, f, c, None if r is None else r.ctypes.data)
    return e, f, c, r


# ---------------------------------------------------------------- numpy-level helpers (real reference)
def ref_classify_grid(vol_xfast):
    from voxel_ma_b200.synth import to_zfast_f64
    nx, ny, nz = _dims(vol_xfast)
    out = np.empty((nz, ny, nx), np.uint8)
    ref().ref_classify_grid(to_zfast_f64(vol_xfast).ravel(), nx, ny, nz, out)
    return out


def ref_tag_points(vol_xfast, xyz):
    from voxel_ma_b200.synth import to_zfast_f64
    nx, ny, nz = _dims(vol_xfast)
    xyz = np.ascontiguousarray(xyz, np.float32).reshape(-1, 3)
    out = np.empty(len(xyz), np.uint8)
    ref().ref_tag_points(to_zfast_f64(vol_xfast).ravel(), nx, ny, nz, xyz, len(xyz), out)
    return out


def ref_extract_sites(vol_xfast):
    from voxel_ma_b200.synth import to_zfast_f64
    nx, ny, nz = _dims(vol_xfast)
    zf = to_zfast_f64(vol_xfast).ravel()
    n = ref().ref_extract_sites(zf, nx, ny, nz, None, 0)
    out = np.empty((n, 3), np.float32)
    ref().ref_extract_sites(zf, nx, ny, nz, out.ctypes.data, n)
    return out


def ref_ann(sites, q, brute=False):
    s = np.ascontiguousarray(sites, np.float64)
    q = np.ascontiguousarray(q, np.float64)
    idx = np.empty(len(q), np.int32)
    d2 = np.empty(len(q), np.float64)
    fn = ref().ref_ann_brute_search if brute else ref().ref_ann_kd_search
    fn(s, len(s), s.shape[1], q, len(q), idx, d2)
    return idx, d2


_ann_grid_shared = None  # (sites, nx, ny, z0, id mapping, d2 mapping): inherited by the forked workers


def _ann_grid_worker(span):
    za, zb = span
    s, nx, ny, z0, id_buf, d2_buf = _ann_grid_shared
    n = (zb - za) * ny * nx
    off = (za - z0) * ny * nx
    ids = np.frombuffer(id_buf, np.int32, n, off * 4)
    d2 = np.frombuffer(d2_buf, np.float64, n, off * 8)
    ref().ref_ann_kd_grid(s, len(s), nx, ny, za, zb, ids.ctypes.data, d2.ctypes.data)
    return ref().ref_last_seconds()


def ref_ann_grid(sites, nx, ny, z0, z1, workers=None):
    """The real ANNkd_tree::annkSearch(k=1, eps=0) at EVERY grid vertex of the planes [z0, z1): forked
    processes (ANN's search state is global), each with its own tree per group of planes, results written
    into shared anonymous mappings.  Returns (ids int32, d2 float64) shaped [z1-z0, ny, nx] and the CPU
    seconds spent inside the reference (tree builds + queries, summed over the workers)."""
    global _ann_grid_shared
    import mmap
    import multiprocessing as mp
    s = np.ascontiguousarray(sites, np.float64).reshape(-1, 3)
    workers = workers or os.cpu_count() or 1
    n = (z1 - z0) * ny * nx
    id_buf, d2_buf = mmap.mmap(-1, max(n * 4, 1)), mmap.mmap(-1, max(n * 8, 1))
    _ann_grid_shared = (s, nx, ny, z0, id_buf, d2_buf)
    # planes dealt in groups (a few per worker) so that slow, deep regions spread over the workers
    step = max(1, -(-(z1 - z0) // (workers * 3)))
    spans = [(za, min(za + step, z1)) for za in range(z0, z1, step)]
    try:
        with mp.get_context("fork").Pool(workers) as pool:
            secs = pool.map(_ann_grid_worker, spans, chunksize=1)
    finally:
        _ann_grid_shared = None
    ids = np.frombuffer(id_buf, np.int32, n).reshape(z1 - z0, ny, nx)
    d2 = np.frombuffer(d2_buf, np.float64, n).reshape(z1 - z0, ny, nx)
    return ids, d2, float(sum(secs))


def ref_ann_fr(sites, q, sq_rad):
    """The reference's two-call pattern (src/voxelapps.cpp:346-353): counts, then every point in range.
    Returns (counts, list of (ids, d2) per query) straight from ANNkd_tree::annkFRSearch."""
    s = np.ascontiguousarray(sites, np.float64)
    q = np.ascontiguousarray(q, np.float64).reshape(-1, 3)
    r = np.ascontiguousarray(np.broadcast_to(np.asarray(sq_rad, np.float64), (len(q),)))
    cnt = np.empty(len(q), np.int32)
    ref().ref_ann_kd_fr_search(s, len(s), 3, q, len(q), r, 0, cnt, None, None)
    kmax = int(cnt.max()) if len(cnt) else 0
    idx = np.full((len(q), max(kmax, 1)), -1, np.int32)
    d2 = np.full((len(q), max(kmax, 1)), -1.0, np.float64)
    if kmax:
        ref().ref_ann_kd_fr_search(s, len(s), 3, q, len(q), r, kmax, cnt, idx.ctypes.data, d2.ctypes.data)
    return cnt, idx, d2


def ref_kdtree_closest(pts, q, max_d2):
    """the real trimesh::KDtree::closest_to_pt + trimesh::dist, as estimateRadiiField uses them: (idx, dist)"""
    p = np.ascontiguousarray(pts, np.float32).reshape(-1, 3)
    q = np.ascontiguousarray(q, np.float32).reshape(-1, 3)
    idx = np.empty(len(q), np.int32)
    d = np.empty(len(q), np.float32)
    ref().ref_kdtree_closest(p, len(p), q, len(q), float(max_d2), idx, d)
    return idx, d


def ref_lambda(a, b):
    a = np.ascontiguousarray(a, np.float32).reshape(-1, 3)
    b = np.ascontiguousarray(b, np.float32).reshape(-1, 3)
    out = np.empty(len(a), np.float32)
    ref().ref_lambda_for_face(a, b, len(a), out)
    return out


def ref_pipeline(vol_xfast, preprocess=False) -> dict:
    """Run the reference's own computeVD [+ preprocessVoro + extractInsideWithMeasure] and return its
    state as numpy arrays."""
    from voxel_ma_b200.synth import to_zfast_f64
    nx, ny, nz = _dims(vol_xfast)
    L = ref()
    h = L.ref_pipeline_run(to_zfast_f64(vol_xfast).ravel(), nx, ny, nz, 1 if preprocess else 0)
    if not h:
        raise RuntimeError("reference pipeline failed")
    try:
        cnt = lambda k: int(L.ref_pipeline_count(h, k))
        out = {"counts": {k: cnt(i) for i, k in enumerate(
            ["sites", "vts", "edges", "faces", "tet_vpts", "out_vts", "out_edges", "out_tris", "face_sites"])}}
        sites = np.empty((cnt(0), 3), np.float32)
        L.ref_pipeline_sites(h, sites)
        out["sites"] = sites
        tv = np.empty((cnt(4), 3), np.float32)
        tt = np.empty(cnt(4), np.uint8)
        L.ref_pipeline_tet_vpts(h, tv, tt)
        out["tet_vpts"], out["tet_vtag"] = tv, tt
        if not preprocess:
            v = np.empty((cnt(1), 3), np.float32)
            r = np.zeros(cnt(1), np.float32)
            sv = np.full(cnt(1), -1, np.int32)
            L.ref_pipeline_vts(h, v, r, sv)
            out["vts"], out["radii"], out["site_of_v"] = v, r, sv
            fs = np.empty((cnt(8), 2), np.int32)
            fl = np.empty(cnt(8), np.float32)
            L.ref_pipeline_faces(h, fs, fl)
            out["face_sites"], out["face_lambda"] = fs, fl
        else:
            vm = np.empty(cnt(5), np.float32)
            em = np.empty(cnt(6), np.float32)
            fm = np.empty(cnt(7), np.float32)
            L.ref_pipeline_measures(h, vm.ctypes.data, em.ctypes.data, fm.ctypes.data, None)
            out["v_msure"], out["e_msure"], out["f_msure"] = vm, em, fm
        return out
    finally:
        L.ref_pipeline_free(h)
