"""Golden vectors for K6 (the seeding of CellComplexThinning::prune) from the UNMODIFIED compiled reference.

    python tests/golden/make_thin_golden.py          (in the build container: needs /root/reference built by oracle/Makefile.ref)

The reference pipeline (-md=vol2ma up to the measures, oracle/ref_shim.cpp) runs on the sphere24 volume of
pipeline_sphere24.npz; its inside complex and measures go through the reference's own cellcomplex /
CellComplexThinning::setup / assignElementValues / preprocess / prune, and the queue prune() seeds is captured by
oracle/ref_thinspy.cpp (the one member it interposes, prune_while_iteration, receives it).  Stored per threshold: the
queue in push order, and the state the seeding read -- the inputs of orc_simple_pairs / vc_simple_pairs -- plus the flat
incidence lists whose histograms are the reference counts (orc_ref_counts / vc_ref_counts).
Runs in a process of its own: the spy library has to be loaded, RTLD_GLOBAL, before libvoxref.so."""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = os.path.join(ROOT, "oracle", "_ref")


def main():
    spy = C.CDLL(os.path.join(REF, "libvoxref_thinspy.so"), mode=C.RTLD_GLOBAL)
    ref = C.CDLL(os.path.join(REF, "libvoxref.so"), mode=C.RTLD_GLOBAL)
    ref.ref_set_data_dir(REF.encode())
    vol = np.load(os.path.join(HERE, "pipeline_sphere24.npz"))["vol"]  # [z][y][x] float32
    nz, ny, nx = vol.shape
    zfast = np.ascontiguousarray(vol.astype(np.float64).transpose(2, 1, 0))  # Tao's order: [x][y][z]
    ref.ref_pipeline_run.restype = C.c_void_p
    ref.ref_pipeline_run.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]
    ref.ref_pipeline_count.restype = C.c_int64
    ref.ref_pipeline_count.argtypes = [C.c_void_p, C.c_int]
    h = ref.ref_pipeline_run(zfast.ctypes.data, nx, ny, nz, 1)
    assert h
    nv, ne, nf = (int(ref.ref_pipeline_count(h, k)) for k in (5, 6, 7))
    vts, edges, tris = np.empty((nv, 3), np.float32), np.empty((ne, 2), np.int32), np.empty((nf, 3), np.int32)
    ref.ref_pipeline_inside.argtypes = [C.c_void_p] * 4
    ref.ref_pipeline_inside(h, vts.ctypes.data, edges.ctypes.data, tris.ctypes.data)
    vm, em, fm = np.empty(nv, np.float32), np.empty(ne, np.float32), np.empty(nf, np.float32)
    ref.ref_pipeline_measures.argtypes = [C.c_void_p] * 5
    ref.ref_pipeline_measures(h, vm.ctypes.data, em.ctypes.data, fm.ctypes.data, None)
    print("inside complex", nv, ne, nf, "face measure range", float(fm.min()), float(fm.max()))
    spy.ref_thin_seed.restype = C.c_int64
    spy.ref_thin_seed.argtypes = ([C.c_void_p, C.c_int64] * 3 + [C.c_void_p] * 3 + [C.c_float, C.c_float, C.c_void_p, C.c_int64] +
                                  [C.c_void_p] * 3 + [C.c_int64] + [C.c_void_p] * 7)
    out = {}
    for tag, t in (("lo", float(np.quantile(fm, 0.25))), ("mid", 1.6), ("hi", float(fm.max()) + 1.0)):
        pairs = np.empty((ne + nv, 3), np.int32)
        edge_ref, edge_face0, edge_m = np.empty(ne, np.int32), np.empty(ne, np.int32), np.empty(ne, np.float32)
        vert_ref, vert_edge0 = np.empty(nv, np.int32), np.empty(nv, np.int32)
        face_m, face_rm = np.empty(nf, np.float32), np.empty(nf, np.uint8)
        ends, face_edges = np.empty(2 * ne, np.int32), np.empty(3 * nf, np.int32)
        sizes = np.zeros(3, np.int64)
        n = spy.ref_thin_seed(vts.ctypes.data, nv, edges.ctypes.data, ne, tris.ctypes.data, nf, vm.ctypes.data, em.ctypes.data,
                              fm.ctypes.data, t, t, pairs.ctypes.data, len(pairs), edge_ref.ctypes.data, edge_face0.ctypes.data,
                              edge_m.ctypes.data, ne, vert_ref.ctypes.data, vert_edge0.ctypes.data, face_m.ctypes.data,
                              face_rm.ctypes.data, ends.ctypes.data, face_edges.ctypes.data, sizes.ctypes.data)
        assert n >= 0, (n, sizes)
        print(tag, "threshold", t, "seeded pairs", n, "(face-edge", int((pairs[:n, 0] == 1).sum()), ")")
        for k, v in dict(pairs=pairs[:n].copy(), edge_ref=edge_ref, edge_face0=edge_face0, edge_measure=edge_m, vert_ref=vert_ref,
                         vert_edge0=vert_edge0, face_measure=face_m, face_to_remove=face_rm, edge_ends=ends, face_edges=face_edges,
                         f_t=np.float32(t), l_t=np.float32(t)).items():
            out[f"{tag}_{k}"] = v
    ref.ref_pipeline_free.argtypes = [C.c_void_p]
    ref.ref_pipeline_free(h)
    np.savez_compressed(os.path.join(HERE, "thin_seed_sphere24.npz"), **out)
    print("wrote thin_seed_sphere24.npz", os.path.getsize(os.path.join(HERE, "thin_seed_sphere24.npz")), "bytes")


if __name__ == "__main__":
    main()
