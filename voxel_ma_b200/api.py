"""Python binding of the C ABI (include/voxcore_gpu.h) -- used by tests/ and bench.py.

Every call goes through ``libvoxcore_gpu.so`` (hand-written CUDA, sm_100a).  There is no CPU
fallback: if the library is missing or no CUDA device is visible, constructing a ``Context``
raises.  Arrays are numpy, x fastest: ``vol[z, y, x]``.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
# VOXCORE_LIB: development override (kernel variants built side by side by tools/build_variants.py)
LIB_PATH = os.environ.get("VOXCORE_LIB") or os.path.join(HERE, "lib", "libvoxcore_gpu.so")

VC_OK = 0
ARR_INSIDE, ARR_ID, ARR_D2X4, ARR_EDGE3, ARR_FACE3, ARR_CUBE, ARR_RADIUS = range(7)

# every symbol include/voxcore_gpu.h declares (tests check the library exports all of them)
SYMBOLS = [
    "vc_abi_version", "vc_warmup", "vc_ctx_create", "vc_ctx_destroy", "vc_last_error", "vc_stream", "vc_synchronize",
    "vc_host_alloc", "vc_host_free", "vc_set_grid", "vc_volume_upload_f32", "vc_volume_upload_i8", "vc_volume_upload_f64_zfast",
    "vc_classify_grid", "vc_classify_points", "vc_classify_mesh", "vc_extract_sites", "vc_get_sites", "vc_set_sites", "vc_num_sites",
    "vc_sites_detect_local", "vc_sites_export_local", "vc_sites_import_global", "vc_peer_create", "vc_peer_open", "vc_peer_open_ptrs",
    "vc_peer_buffer", "vc_peer_close", "vc_peer_set_timeout", "vc_sites_post_peers", "vc_sites_collect_peers", "vc_closest_grid",
    "vc_closest_points", "vc_closest_points_f32", "vc_radius_search", "vc_cell_measures_grid", "vc_face_lambda", "vc_vertex_radii", "vc_segment_max", "vc_ref_counts", "vc_simple_pairs",
    "vc_run_dense", "vc_closest_and_measures", "vc_set_pipeline", "vc_download", "vc_download_planes", "vc_device_ptr", "vc_run_dense_host", "vc_compact_count", "vc_compact_records", "vc_medial_quads_count", "vc_medial_quads", "vc_cell_circum_angle_grid",
    "vc_run_dense_host_compact", "vc_run_dense_host_compact_i8", "vc_set_compact_mode", "vc_profile_enable", "vc_profile_reset",
    "vc_profile_count", "vc_profile_get", "vc_launch_count",
]


class VoxcoreError(RuntimeError):
    pass


_lib = None


def load_library(path: str | None = None):
    """dlopen libvoxcore_gpu.so; raises if it has not been built (python -m voxel_ma_b200.build)."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise VoxcoreError(f"{p} is missing: build it with `python -m voxel_ma_b200.build` "
                           "(there is no CPU fallback for this path)")
    lib = C.CDLL(p)
    vp, i64, i32 = C.c_void_p, C.c_int64, C.c_int
    lib.vc_abi_version.restype = i32
    lib.vc_ctx_create.argtypes = [i32, C.POINTER(vp)]
    lib.vc_ctx_destroy.argtypes = [vp]
    lib.vc_ctx_destroy.restype = None
    lib.vc_last_error.argtypes = [vp]
    lib.vc_last_error.restype = C.c_char_p
    lib.vc_stream.argtypes = [vp]
    lib.vc_stream.restype = vp
    lib.vc_synchronize.argtypes = [vp]
    lib.vc_host_alloc.argtypes = [C.c_size_t]
    lib.vc_host_alloc.restype = vp
    lib.vc_host_free.argtypes = [vp]
    lib.vc_host_free.restype = None
    lib.vc_set_grid.argtypes = [vp, i32, i32, i32, i32, i32]
    lib.vc_volume_upload_f32.argtypes = [vp, vp, i32, i32]
    lib.vc_volume_upload_f64_zfast.argtypes = [vp, vp]
    lib.vc_classify_grid.argtypes = [vp, vp]
    lib.vc_classify_points.argtypes = [vp, vp, i64, vp, vp]
    lib.vc_classify_mesh.argtypes = [vp, vp, i64, vp, i64, vp, vp]
    lib.vc_extract_sites.argtypes = [vp, C.POINTER(i64)]
    lib.vc_get_sites.argtypes = [vp, vp]
    lib.vc_set_sites.argtypes = [vp, vp, i64]
    lib.vc_num_sites.argtypes = [vp]
    lib.vc_num_sites.restype = i64
    lib.vc_sites_detect_local.argtypes = [vp, C.POINTER(i64)]
    lib.vc_sites_export_local.argtypes = [vp, vp, vp]
    lib.vc_sites_import_global.argtypes = [vp, vp, vp, i64]
    lib.vc_peer_create.argtypes = [vp, i32, i32, i64, vp]
    lib.vc_peer_open.argtypes = [vp, vp]
    lib.vc_peer_open_ptrs.argtypes = [vp, vp]
    lib.vc_peer_buffer.argtypes = [vp]
    lib.vc_peer_buffer.restype = vp
    lib.vc_peer_close.argtypes = [vp]
    lib.vc_peer_set_timeout.argtypes = [vp, i64]
    lib.vc_sites_post_peers.argtypes = [vp]
    lib.vc_sites_collect_peers.argtypes = [vp, C.POINTER(i64)]
    lib.vc_set_compact_mode.argtypes = [vp, i32]
    lib.vc_compact_count.argtypes = [vp, C.POINTER(i64)]
    lib.vc_compact_records.argtypes = [vp, i64, vp, vp, vp, vp, vp]
    lib.vc_run_dense_host_compact.argtypes = [vp, vp, vp, i64, C.POINTER(i64), vp, vp, vp, vp, vp, vp, vp, C.POINTER(i64)]
    lib.vc_run_dense_host_compact_i8.argtypes = [vp, vp, vp, i64, C.POINTER(i64), vp, vp, vp, vp, vp, vp, vp, C.POINTER(i64)]
    lib.vc_volume_upload_i8.argtypes = [vp, vp, i32, i32]
    lib.vc_closest_grid.argtypes = [vp, vp, vp]
    lib.vc_closest_points.argtypes = [vp, vp, i64, vp, vp]
    lib.vc_closest_points_f32.argtypes = [vp, vp, i64, C.c_float, vp, vp]
    lib.vc_radius_search.argtypes = [vp, vp, vp, i64, vp, vp, vp, vp]
    lib.vc_cell_measures_grid.argtypes = [vp, vp, vp, vp, vp]
    lib.vc_face_lambda.argtypes = [vp, vp, i64, vp]
    lib.vc_vertex_radii.argtypes = [vp, vp, i64, vp, vp]
    lib.vc_segment_max.argtypes = [vp, vp, vp, i64, vp, i64, vp, vp]
    lib.vc_ref_counts.argtypes = [vp, vp, i64, i64, vp]
    lib.vc_simple_pairs.argtypes = [vp, vp, vp, i64, vp, vp, i64, C.c_float, vp, vp, i64, vp, C.c_float, vp, i64, C.POINTER(i64)]
    lib.vc_run_dense.argtypes = [vp, C.POINTER(i64)]
    lib.vc_closest_and_measures.argtypes = [vp]
    lib.vc_set_pipeline.argtypes = [vp, i32, i32]
    lib.vc_download.argtypes = [vp, i32, vp]
    lib.vc_download_planes.argtypes = [vp, i32, i32, i32, vp]
    lib.vc_device_ptr.argtypes = [vp, i32]
    lib.vc_device_ptr.restype = vp
    lib.vc_run_dense_host.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, vp, C.POINTER(i64)]
    lib.vc_cell_circum_angle_grid.argtypes = [vp, i32, i32, vp, vp]
    lib.vc_medial_quads_count.argtypes = [vp, C.POINTER(i64)]
    lib.vc_medial_quads.argtypes = [vp, i64, vp, vp, vp, vp, vp, C.POINTER(i64)]
    lib.vc_profile_enable.argtypes = [vp, i32]
    lib.vc_profile_reset.argtypes = [vp]
    lib.vc_profile_count.argtypes = [vp]
    lib.vc_profile_get.argtypes = [vp, i32, C.POINTER(C.c_char_p), C.POINTER(C.c_double), C.POINTER(i64)]
    lib.vc_launch_count.argtypes = [vp]
    lib.vc_launch_count.restype = i64
    if path is None:
        _lib = lib
    return lib


def _ptr(a):
    if a is None:
        return None
    if isinstance(a, int):
        return C.c_void_p(a)
    return C.c_void_p(a.ctypes.data)


class PinnedArray:
    """numpy view over page-locked memory from vc_host_alloc (freed with the object)."""

    def __init__(self, shape, dtype):
        self._lib = load_library()
        n = int(np.prod(shape)) * np.dtype(dtype).itemsize
        self._p = self._lib.vc_host_alloc(max(n, 1))
        if not self._p:
            raise VoxcoreError("vc_host_alloc failed")
        buf = (C.c_uint8 * max(n, 1)).from_address(self._p)
        self.array = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)

    def __del__(self):
        try:
            if getattr(self, "_p", None):
                self.array = None
                self._lib.vc_host_free(self._p)
                self._p = None
        except Exception:
            pass


class Context:
    """One GPU context = one z-slab of the grid (the whole grid on a single GPU)."""

    def __init__(self, device: int = 0):
        self.lib = load_library()
        h = C.c_void_p()
        st = self.lib.vc_ctx_create(device, C.byref(h))
        if st != VC_OK:
            raise VoxcoreError(f"vc_ctx_create(device={device}) failed with {st}: no usable CUDA device "
                               "(this path has no CPU fallback)")
        self.h = h
        self.nx = self.ny = self.nz = self.z0 = self.z1 = 0

    def close(self):
        if getattr(self, "h", None):
            self.lib.vc_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _ck(self, st):
        if st != VC_OK:
            raise VoxcoreError(f"status {st}: {self.lib.vc_last_error(self.h).decode()}")

    # ---- grid / volume
    def set_grid(self, nx, ny, nz, z0=0, z1=None):
        z1 = nz if z1 is None else z1
        self._ck(self.lib.vc_set_grid(self.h, nx, ny, nz, z0, z1))
        self.nx, self.ny, self.nz, self.z0, self.z1 = nx, ny, nz, z0, z1

    @property
    def slab_shape(self):
        return (self.z1 - self.z0, self.ny, self.nx)

    def upload_volume(self, vol: np.ndarray, zlo: int = 0):
        """vol[z,y,x] planes starting at global plane zlo (whole volume: zlo=0): float32 (MRC mode 2), or int8 as it is
        (MRC mode 0)."""
        i8 = vol.dtype == np.int8
        v = np.ascontiguousarray(vol, np.int8 if i8 else np.float32)
        if v.ndim != 3:
            raise ValueError(f"upload_volume wants [z][y][x] planes, got shape {v.shape}")
        if not self.nx:
            nz, ny, nx = v.shape
            self.set_grid(nx, ny, nz)
        if v.shape[1:] != (self.ny, self.nx):  # a stale grid would be read with the wrong row length: silently wrong results
            raise ValueError(f"volume planes are {v.shape[1:]} but the context's grid is {(self.ny, self.nx)} (y, x): call set_grid first")
        if zlo < 0 or zlo + v.shape[0] > self.nz:
            raise ValueError(f"planes [{zlo}, {zlo + v.shape[0]}) lie outside the grid's {self.nz} planes")
        fn = self.lib.vc_volume_upload_i8 if i8 else self.lib.vc_volume_upload_f32
        self._ck(fn(self.h, _ptr(v), zlo, zlo + v.shape[0]))

    def upload_volume_f64_zfast(self, vol_zfast: np.ndarray, nx, ny, nz):
        self.set_grid(nx, ny, nz)
        v = np.ascontiguousarray(vol_zfast, np.float64)
        self._ck(self.lib.vc_volume_upload_f64_zfast(self.h, _ptr(v)))

    # ---- stages
    def classify_grid(self, fetch=True):
        out = np.empty(self.slab_shape, np.uint8) if fetch else None
        self._ck(self.lib.vc_classify_grid(self.h, _ptr(out)))
        return out

    def classify_mesh(self, verts, tris, M=None, fetch=True):
        """Stage 1': parity classification of the grid from a closed triangle mesh (vc_classify_mesh)."""
        v = np.ascontiguousarray(verts, np.float32).reshape(-1, 3)
        t = np.ascontiguousarray(tris, np.uint32).reshape(-1, 3)
        m = None if M is None else np.ascontiguousarray(M, np.float64)
        out = np.empty(self.slab_shape, np.uint8) if fetch else None
        self._ck(self.lib.vc_classify_mesh(self.h, _ptr(v), len(v), _ptr(t), len(t), _ptr(m), _ptr(out)))
        return out

    def classify_points(self, xyz, M=None):
        p = np.ascontiguousarray(xyz, np.float32).reshape(-1, 3)
        out = np.empty(len(p), np.uint8)
        m = None if M is None else np.ascontiguousarray(M, np.float64)
        self._ck(self.lib.vc_classify_points(self.h, _ptr(p), len(p), _ptr(m), _ptr(out)))
        return out

    def extract_sites(self) -> int:
        n = C.c_int64()
        self._ck(self.lib.vc_extract_sites(self.h, C.byref(n)))
        return n.value

    def num_sites(self) -> int:
        return int(self.lib.vc_num_sites(self.h))

    def get_sites(self) -> np.ndarray:
        out = np.empty((max(self.num_sites(), 0), 3), np.float32)
        self._ck(self.lib.vc_get_sites(self.h, _ptr(out)))
        return out

    def set_sites(self, xyz):
        p = np.ascontiguousarray(xyz, np.float32).reshape(-1, 3)
        self._ck(self.lib.vc_set_sites(self.h, _ptr(p), len(p)))

    def sites_detect_local(self) -> int:
        n = C.c_int64()
        self._ck(self.lib.vc_sites_detect_local(self.h, C.byref(n)))
        return n.value

    def sites_export_local(self, keys, corners):
        """keys / corners: numpy uint64 arrays or raw device pointers (int)."""
        self._ck(self.lib.vc_sites_export_local(self.h, _ptr(keys), _ptr(corners)))

    def sites_import_global(self, keys, corners, n):
        self._ck(self.lib.vc_sites_import_global(self.h, _ptr(keys), _ptr(corners), n))

    # ---- the exchange over peer memory (csrc/vc_peer.cu)
    def peer_create(self, world: int, rank: int, cap: int) -> bytes:
        """Allocate this rank's receive buffer; returns its 64-byte CUDA IPC handle."""
        h = C.create_string_buffer(64)
        self._ck(self.lib.vc_peer_create(self.h, world, rank, cap, C.cast(h, C.c_void_p)))
        return h.raw

    def peer_open(self, handles) -> None:
        """handles: the IPC handles of all ranks in rank order (one process per GPU)."""
        blob = b"".join(bytes(h) for h in handles)
        buf = C.create_string_buffer(blob, len(blob))
        self._ck(self.lib.vc_peer_open(self.h, C.cast(buf, C.c_void_p)))

    def peer_open_ptrs(self, bases) -> None:
        """bases: receive-buffer pointers of all ranks (contexts of one process)."""
        arr = (C.c_void_p * len(bases))(*bases)
        self._ck(self.lib.vc_peer_open_ptrs(self.h, C.cast(arr, C.c_void_p)))

    def peer_buffer(self) -> int:
        return self.lib.vc_peer_buffer(self.h)

    def peer_set_timeout(self, milliseconds: int) -> None:
        """bound of the wait for the other ranks' posts (default 10 s / VC_PEER_TIMEOUT_MS)"""
        self._ck(self.lib.vc_peer_set_timeout(self.h, int(milliseconds)))

    def peer_close(self) -> None:
        self._ck(self.lib.vc_peer_close(self.h))

    def sites_post_peers(self) -> None:
        self._ck(self.lib.vc_sites_post_peers(self.h))

    def sites_collect_peers(self) -> int:
        n = C.c_int64()
        self._ck(self.lib.vc_sites_collect_peers(self.h, C.byref(n)))
        return n.value

    def closest_grid(self, fetch=True):
        ids = np.empty(self.slab_shape, np.int32) if fetch else None
        d2 = np.empty(self.slab_shape, np.uint32) if fetch else None
        self._ck(self.lib.vc_closest_grid(self.h, _ptr(ids), _ptr(d2)))
        return ids, d2

    def closest_points_f32(self, q, max_d2=0.0):
        """trimesh::KDtree::closest_to_pt for a batch: float32 distances, (id, d2); -1 where nothing within max_d2"""
        q = np.ascontiguousarray(q, np.float32).reshape(-1, 3)
        ids = np.empty(len(q), np.int32)
        d2 = np.empty(len(q), np.float32)
        self._ck(self.lib.vc_closest_points_f32(self.h, _ptr(q), len(q), float(max_d2), _ptr(ids), _ptr(d2)))
        return ids, d2

    def closest_points(self, q):
        q = np.ascontiguousarray(q, np.float64).reshape(-1, 3)
        ids = np.empty(len(q), np.int32)
        d2 = np.empty(len(q), np.float64)
        self._ck(self.lib.vc_closest_points(self.h, _ptr(q), len(q), _ptr(ids), _ptr(d2)))
        return ids, d2

    def radius_search(self, q, sq_rad, fetch=True):
        """annkFRSearch drop-in: counts, then (CSR) ids / squared distances of all sites within sq_rad[i]."""
        p = np.ascontiguousarray(q, np.float64).reshape(-1, 3)
        r = np.ascontiguousarray(np.broadcast_to(np.asarray(sq_rad, np.float64), (len(p),)))
        cnt = np.empty(len(p), np.int32)
        self._ck(self.lib.vc_radius_search(self.h, _ptr(p), _ptr(r), len(p), None, _ptr(cnt), None, None))
        if not fetch:
            return cnt
        off = np.zeros(len(p) + 1, np.int64)
        np.cumsum(cnt, out=off[1:])
        idx = np.empty(max(int(off[-1]), 1), np.int32)
        d2 = np.empty(max(int(off[-1]), 1), np.float64)
        self._ck(self.lib.vc_radius_search(self.h, _ptr(p), _ptr(r), len(p), _ptr(off), None, _ptr(idx), _ptr(d2)))
        return cnt, off, idx[: off[-1]], d2[: off[-1]]

    def cell_measures_grid(self, fetch=True):
        s = self.slab_shape
        if fetch:
            e, f = np.empty((3,) + s, np.float32), np.empty((3,) + s, np.float32)
            c, r = np.empty(s, np.float32), np.empty(s, np.float32)
        else:
            e = f = c = r = None
        self._ck(self.lib.vc_cell_measures_grid(self.h, _ptr(e), _ptr(f), _ptr(c), _ptr(r)))
        return e, f, c, r

    def face_lambda(self, site_pairs):
        p = np.ascontiguousarray(site_pairs, np.int32).reshape(-1, 2)
        out = np.empty(len(p), np.float32)
        self._ck(self.lib.vc_face_lambda(self.h, _ptr(p), len(p), _ptr(out)))
        return out

    def vertex_radii(self, v_xyz, site_of_v):
        v = np.ascontiguousarray(v_xyz, np.float32).reshape(-1, 3)
        s = np.ascontiguousarray(site_of_v, np.int32)
        out = np.empty(len(v), np.float32)
        self._ck(self.lib.vc_vertex_radii(self.h, _ptr(v), len(v), _ptr(s), _ptr(out)))
        return out

    def segment_max(self, off, items, value, valid=None):
        off = np.ascontiguousarray(off, np.int32)
        items = np.ascontiguousarray(items, np.int32)
        value = np.ascontiguousarray(value, np.float32)
        vv = None if valid is None else np.ascontiguousarray(valid, np.uint8)
        out = np.empty(len(off) - 1, np.float32)
        self._ck(self.lib.vc_segment_max(self.h, _ptr(off), _ptr(items), len(out), _ptr(value), len(value), _ptr(vv),
                                         _ptr(out)))
        return out

    # ---- whole path
    def ref_counts(self, idx, nbins):
        """K6: cellcomplex::refCntPerVert / refCntPerEdge as a histogram of incidence indices."""
        i = np.ascontiguousarray(idx, np.int32).ravel()
        out = np.empty(nbins, np.int32)
        self._ck(self.lib.vc_ref_counts(self.h, _ptr(i), len(i), nbins, _ptr(out)))
        return out

    def simple_pairs(self, edge_ref, edge_face0, face_measure, f_t, vert_ref, vert_edge0, edge_measure, l_t, face_to_remove=None):
        """K6: the queue seeding of CellComplexThinning::prune; returns int32 (n, 3) rows (type, idx0, idx1)."""
        er, ef = (np.ascontiguousarray(a, np.int32) for a in (edge_ref, edge_face0))
        vr, ve = (np.ascontiguousarray(a, np.int32) for a in (vert_ref, vert_edge0))
        fm, em = (np.ascontiguousarray(a, np.float32) for a in (face_measure, edge_measure))
        tr = None if face_to_remove is None else np.ascontiguousarray(face_to_remove, np.uint8)
        n = C.c_int64(0)
        cap = len(er) + len(vr)
        out = np.empty((max(cap, 1), 3), np.int32)
        self._ck(self.lib.vc_simple_pairs(self.h, _ptr(er), _ptr(ef), len(er), _ptr(fm), _ptr(tr), len(fm), C.c_float(f_t), _ptr(vr),
                                          _ptr(ve), len(vr), _ptr(em), C.c_float(l_t), _ptr(out), cap, C.byref(n)))
        return out[: n.value].copy()

    def run_dense(self) -> int:
        n = C.c_int64()
        self._ck(self.lib.vc_run_dense(self.h, C.byref(n)))
        return n.value

    def closest_and_measures(self):
        """stages 2+3 for this ctx's planes (sites already set), pipelined over z chunks; results stay on the device"""
        self._ck(self.lib.vc_closest_and_measures(self.h))

    def set_pipeline(self, workers=8, zchunk=0):
        self._ck(self.lib.vc_set_pipeline(self.h, workers, zchunk))

    def download(self, which) -> np.ndarray:
        s = self.slab_shape
        shape, dt = {
            ARR_INSIDE: (s, np.uint8), ARR_ID: (s, np.int32), ARR_D2X4: (s, np.uint32),
            ARR_EDGE3: ((3,) + s, np.float32), ARR_FACE3: ((3,) + s, np.float32),
            ARR_CUBE: (s, np.float32), ARR_RADIUS: (s, np.float32),
        }[which]
        out = np.empty(shape, dt)
        self._ck(self.lib.vc_download(self.h, which, _ptr(out)))
        return out

    def download_planes(self, which, za, zb) -> np.ndarray:
        """planes [za, zb) (global z) of a result array; ids / 4d2 may include the halo plane z1 of a slab"""
        s = (zb - za,) + self.slab_shape[1:]
        shape, dt = {
            ARR_INSIDE: (s, np.uint8), ARR_ID: (s, np.int32), ARR_D2X4: (s, np.uint32),
            ARR_EDGE3: ((3,) + s, np.float32), ARR_FACE3: ((3,) + s, np.float32),
            ARR_CUBE: (s, np.float32), ARR_RADIUS: (s, np.float32),
        }[which]
        out = np.empty(shape, dt)
        self._ck(self.lib.vc_download_planes(self.h, which, za, zb, _ptr(out)))
        return out

    def device_ptr(self, which) -> int:
        return int(self.lib.vc_device_ptr(self.h, which) or 0)

    @staticmethod
    def _host_buf(a, dtype, count, what):
        """the C ABI takes raw host pointers: a wrong dtype, a strided view or a short array would make the copies read or
        write past the buffer"""
        if a is None:
            return
        if not isinstance(a, np.ndarray) or a.dtype != np.dtype(dtype) or not a.flags.c_contiguous or a.size < count:
            raise ValueError(f"{what}: need a C-contiguous {np.dtype(dtype).name} array of >= {count} elements, got "
                             f"{getattr(a, 'dtype', type(a))} {getattr(a, 'shape', '')}")

    def run_dense_host(self, vol, inside=None, ids=None, d2x4=None, edge3=None, face3=None, cube=None, radius=None) -> int:
        nv = self.nx * self.ny * self.nz
        self._host_buf(vol, np.float32, nv, "vol")
        for a, dt, k, what in ((inside, np.uint8, 1, "inside"), (ids, np.int32, 1, "ids"), (d2x4, np.uint32, 1, "d2x4"),
                               (edge3, np.float32, 3, "edge3"), (face3, np.float32, 3, "face3"), (cube, np.float32, 1, "cube"),
                               (radius, np.float32, 1, "radius")):
            self._host_buf(a, dt, k * nv, what)
        n = C.c_int64()
        self._ck(self.lib.vc_run_dense_host(self.h, _ptr(vol), _ptr(inside), _ptr(ids), _ptr(d2x4), _ptr(edge3),
                                            _ptr(face3), _ptr(cube), _ptr(radius), C.byref(n)))
        return n.value

    # ---- compact product: records of the inside vertices (csrc/vc_compact.cu)
    def set_compact_mode(self, mode: int) -> None:
        """0 automatic, 1 dense planes + gather, 2 records computed directly (vc_run_dense_host_compact)"""
        self._ck(self.lib.vc_set_compact_mode(self.h, mode))

    def compact_count(self) -> int:
        n = C.c_int64()
        self._ck(self.lib.vc_compact_count(self.h, C.byref(n)))
        return n.value

    def compact_records(self):
        """(vert u32[n], id i32[n], d2x4 u32[n], lambda7 f32[7][n], radius f32[n]) of the inside vertices"""
        n = self.compact_count()
        vert, ids, d2 = np.empty(n, np.uint32), np.empty(n, np.int32), np.empty(n, np.uint32)
        lam, rad = np.empty((7, n), np.float32), np.empty(n, np.float32)
        self._ck(self.lib.vc_compact_records(self.h, n, _ptr(vert), _ptr(ids), _ptr(d2), _ptr(lam), _ptr(rad)))
        return vert, ids, d2, lam, rad

    def run_dense_host_compact(self, vol, cap, inside_bits=None, vert=None, ids=None, d2x4=None, lambda7=None, radius=None,
                               id_dense=None, d2x4_dense=None):
        """-> (n_inside, n_sites).  lambda7 must be laid out [7][cap]."""
        planes = self.z1 - self.z0
        nres = (min(self.z1 + 1, self.nz) - max(self.z0 - 1, 0)) * self.ny * self.nx  # the slab's resident voxel planes
        self._host_buf(vol, np.int8 if vol.dtype == np.int8 else np.float32, nres, "vol")
        self._host_buf(inside_bits, np.uint32, planes * self.ny * (self.nx // 32 + 1), "inside_bits")
        for a, dt, k, what in ((vert, np.uint32, cap, "vert"), (ids, np.int32, cap, "ids"), (d2x4, np.uint32, cap, "d2x4"),
                               (radius, np.float32, cap, "radius"), (id_dense, np.int32, planes * self.ny * self.nx, "id_dense"),
                               (d2x4_dense, np.uint32, planes * self.ny * self.nx, "d2x4_dense")):
            self._host_buf(a, dt, k, what)
        if lambda7 is not None:
            self._host_buf(lambda7, np.float32, 7 * cap, "lambda7")
            if lambda7.shape != (7, cap):
                raise ValueError(f"lambda7 must be laid out [7][cap] = (7, {cap}), got {lambda7.shape}")
        n, ns = C.c_int64(), C.c_int64()
        fn = self.lib.vc_run_dense_host_compact_i8 if vol.dtype == np.int8 else self.lib.vc_run_dense_host_compact
        self._ck(fn(self.h, _ptr(vol), _ptr(inside_bits), cap, C.byref(n), _ptr(vert), _ptr(ids),
                                                    _ptr(d2x4), _ptr(lambda7), _ptr(radius), _ptr(id_dense), _ptr(d2x4_dense),
                                                    C.byref(ns)))
        return n.value, ns.value

    def cell_circum_angle_grid(self, za=None, zb=None):
        """(circumradius f64 [7][z][y][x], object angle f64 [7][z][y][x]) of the planes [za, zb) (default: all owned planes)"""
        za = self.z0 if za is None else za
        zb = self.z1 if zb is None else zb
        circ = np.empty((7, zb - za, self.ny, self.nx), np.float64)
        ang = np.empty_like(circ)
        self._ck(self.lib.vc_cell_circum_angle_grid(self.h, za, zb, _ptr(circ), _ptr(ang)))
        return circ, ang

    # ---- the medial complex of the dense product (csrc/vc_medial.cu)
    def medial_quads(self):
        """(anchor u32[n], axis u8[n], site_a i32[n], site_b i32[n], lambda f32[n]) of the dual quads, ascending (z, y, x, axis)"""
        n = C.c_int64()
        self._ck(self.lib.vc_medial_quads_count(self.h, C.byref(n)))
        m = max(n.value, 1)
        anchor, axis = np.empty(m, np.uint32), np.empty(m, np.uint8)
        a, b, lam = np.empty(m, np.int32), np.empty(m, np.int32), np.empty(m, np.float32)
        self._ck(self.lib.vc_medial_quads(self.h, m, _ptr(anchor), _ptr(axis), _ptr(a), _ptr(b), _ptr(lam), C.byref(n)))
        k = n.value
        return anchor[:k], axis[:k], a[:k], b[:k], lam[:k]

    def synchronize(self):
        self._ck(self.lib.vc_synchronize(self.h))

    def stream(self) -> int:
        return int(self.lib.vc_stream(self.h) or 0)

    # ---- instrumentation
    def profile(self, on=True):
        self.lib.vc_profile_enable(self.h, 1 if on else 0)

    def profile_reset(self):
        self.lib.vc_profile_reset(self.h)

    def profile_report(self) -> dict:
        out = {}
        for i in range(self.lib.vc_profile_count(self.h)):
            name, ms, n = C.c_char_p(), C.c_double(), C.c_int64()
            self.lib.vc_profile_get(self.h, i, C.byref(name), C.byref(ms), C.byref(n))
            out[name.value.decode()] = {"ms": ms.value, "launches": n.value}
        return out

    def launch_count(self) -> int:
        return int(self.lib.vc_launch_count(self.h))
