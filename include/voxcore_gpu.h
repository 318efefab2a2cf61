/* voxcore_gpu.h -- C ABI of libvoxcore_gpu.so, the B200 (sm_100a) implementation of Voxel Cores'
 * data-parallel front end: inside/outside classification of the dense grid, boundary-sample
 * ("site") extraction, the exact closest-site query per grid vertex, and the per-cell medial
 * measures.
 *
 * The reference (danielyan86129/voxel_ma) has no plugin / FFI interface; its boundary is the set
 * of C++ functions the CLI calls.  Each entry point below names the
 * reference function (file:line under the reference tree) whose data-parallel body it replaces;
 * INTEGRATION.md shows the call a maintainer adds at that line.
 *
 * Conventions
 *   - plain C: opaque handle, pointers and sizes; no exceptions, no C++/torch types.
 *   - every call returns vc_status (0 = ok, negative = error); vc_last_error(ctx) has the text.
 *   - host pointers are caller-owned; device memory is library-owned and freed by vc_ctx_destroy.
 *     Pointers marked "host or device" are detected with cudaPointerGetAttributes.
 *   - calls are synchronous for the caller unless a function says otherwise; one host thread per ctx.
 *   - one vc_ctx per GPU (one process per GPU; multi-GPU = one ctx per rank over z-slabs).
 *   - dense arrays are x-fastest: index = x + nx*(y + ny*z)  (MRC payload order,
 *     3rdparty/isosurface_tao/reader.h:232-251).  A ctx owns the grid-vertex planes [z0, z1).
 *   - there is NO CPU fallback anywhere behind this interface: without a CUDA device
 *     vc_ctx_create fails with VC_ERR_CUDA.
 */
#ifndef VOXCORE_GPU_H
#define VOXCORE_GPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct vc_ctx vc_ctx;

typedef enum vc_status
{
    VC_OK = 0,
    VC_ERR_INVALID = -1,     /* bad argument */
    VC_ERR_CUDA = -2,        /* CUDA runtime error (text in vc_last_error) */
    VC_ERR_STATE = -3,       /* call order: a prerequisite stage has not run */
    VC_ERR_NOMEM = -4,       /* device or host allocation failed */
    VC_ERR_UNSUPPORTED = -5  /* valid request this build does not implement */
} vc_status;

/* Starts the CUDA driver and the device's primary context, nothing else (no vc_ctx, no message): a host program may call
 * it from a helper thread at start-up so that the seconds this takes on a cold GPU pass while it reads its input
 * (voxel_ma_b200/host/dropin/gpu_surfacing.cpp does, for the reference CLI).  VC_OK or VC_ERR_CUDA. */
int vc_warmup(int device);
/* ---- context ------------------------------------------------------------------------------ */
int vc_abi_version(void);
int vc_ctx_create(int device, vc_ctx** out);
void vc_ctx_destroy(vc_ctx* ctx);
const char* vc_last_error(const vc_ctx* ctx);
/* the CUDA stream every kernel of this ctx is launched on (cudaStream_t as void*) */
void* vc_stream(vc_ctx* ctx);
int vc_synchronize(vc_ctx* ctx);
/* page-locked host memory for full-speed copies (optional; any host pointer is accepted) */
void* vc_host_alloc(size_t bytes);
void vc_host_free(void* p);

/* ---- grid and volume ------------------------------------------------------------------------
 * Replaces the storage behind Volume3DScalar::getDataAt (include/Volume3DScalar.h:18,
 * include/densevolume.h:36-43, 3rdparty/isosurface_tao/volume.h:217-224).
 * vc_set_grid: global size and the z-slab [z0,z1) of grid-vertex planes this ctx owns
 * (single GPU: z0=0, z1=nz). */
int vc_set_grid(vc_ctx* ctx, int nx, int ny, int nz, int z0, int z1);
/* float32 voxel planes [zlo,zhi) in MRC payload order (x fastest); `planes` points at plane zlo.
 * Must cover [max(z0-1,0), min(z1+1,nz)) -- one halo plane each side of the slab.
 * planes: host or device. */
int vc_volume_upload_f32(vc_ctx* ctx, const float* planes, int zlo, int zhi);
/* The same for an MRC mode 0 payload (signed bytes, x fastest): the reference's reader widens them to double
 * exactly like the float ones (3rdparty/isosurface_tao/reader.h:235-239), so inside <=> value > 0. */
int vc_volume_upload_i8(vc_ctx* ctx, const int8_t* planes, int zlo, int zhi);
/* whole volume in Tao's in-memory order double[x*ny*nz + y*nz + z] (volume.h:217-224);
 * requires the ctx to own the whole grid. */
int vc_volume_upload_f64_zfast(vc_ctx* ctx, const double* vol);

/* ---- stage 1: inside / outside ----------------------------------------------------------------
 * a2: SpaceConverter::get_occupancy_at_vox / voxTaggedAsInside over every voxel
 * (include/spaceinfo.h:53-58,122-129): inside <=> value > 0.0.
 * inside_out (nullable, host or device): uint8 flags of the owned planes [z0,z1). */
int vc_classify_grid(vc_ctx* ctx, uint8_t* inside_out);
/* a4: VoroInfo::tagVert / tagVtsUsingUniformVol (src/voroinfo.cpp:447-494): q = M*p (double 4x4,
 * column-major, homogeneous divide, cast to float -- XForm.h:479-489), voxel = round-half-away(q),
 * then a2.  M == NULL means identity (what tagVert passes).  Needs vc_classify_grid first and the
 * ctx must hold the whole grid.  xyz, out: host. */
int vc_classify_points(vc_ctx* ctx, const float* xyz, int64_t n, const double* M, uint8_t* out);

/* ---- stage 1': inside / outside straight from a closed triangle mesh ------------------------------
 * New surface (SURVEY section 8a "K1'", 8f): the reference only ever sees a voxelised volume.  Parity
 * rule: voxel centre (i,j,k) is inside iff its +x ray crosses the mesh an odd number of times.
 * verts = nv x 3 float32 in model space, tris = nt x 3 uint32, M = model -> voxel transform (double
 * 4x4, column-major, NULL = identity) applied like VoroInfo::tagVert's (XForm.h:479-489); transformed
 * vertices are snapped to 1/256 voxel and every decision is exact integer arithmetic, so the flags
 * are reproducible bit for bit (rules: voxel_ma_b200/csrc/vc_mesh_core.h).  Vertices must lie within
 * [-1024, 3072) voxels.  The flags replace an uploaded volume: the sites / closest / measures stages
 * run on them unchanged.  Needs vc_set_grid.  verts, tris: host or device; inside_out nullable. */
int vc_classify_mesh(vc_ctx* ctx, const float* verts, int64_t nv, const uint32_t* tris, int64_t nt, const double* M,
                     uint8_t* inside_out);

/* ---- stage 1'': boundary samples ("sites") ----------------------------------------------------
 * a3: Surfacer::extractBoundaryVts (src/surfacing.cpp:223-321): the unique corners of all voxel
 * faces that separate a 0-voxel from a 1-voxel, numbered in the reference's first-encounter order
 * (x-major voxel scan, neighbour slot, corner slot).  Single-GPU form: */
int vc_extract_sites(vc_ctx* ctx, int64_t* nsites);
/* float32 xyz triples in site-id order (half-integer coordinates).  xyz_out: host. */
int vc_get_sites(vc_ctx* ctx, float* xyz_out);
/* External sample set (-siteFile / .node, src/voroUtility.cpp:74-76): replaces the extracted
 * sites.  Sites that all lie on the half-integer corner lattice of the grid use the dense
 * transform; any other set is served by the cell-list search. xyz: host. */
int vc_set_sites(vc_ctx* ctx, const float* xyz, int64_t n);
int64_t vc_num_sites(const vc_ctx* ctx);
/* Multi-GPU form (one collective between the two calls): each rank detects the
 * site corners of its own slab, ranks all-gather the (key, corner) records, every rank imports
 * the union and sorts it identically.  Records: key = first-encounter key (uint64), corner =
 * cx | cy<<21 | cz<<42 (uint64). Buffers: host or device. */
int vc_sites_detect_local(vc_ctx* ctx, int64_t* nlocal);
int vc_sites_export_local(vc_ctx* ctx, uint64_t* keys_out, uint64_t* corners_out);
int vc_sites_import_global(vc_ctx* ctx, const uint64_t* keys, const uint64_t* corners, int64_t n);
/* The same exchange over peer memory (NVLink), without a collective library on the data path: the
 * records of a rank's own corner planes are sorted by key on that rank and stored as one sorted run straight
 * into the receive buffer of every rank of the slab group, one warp posts (sequence, count) with a
 * system-scope release, and a rank collects by waiting on its own header and merging the runs by
 * ranking -- no rank sorts the union (voxel_ma_b200/csrc/vc_peer.cu).  Set-up, once per group:
 *   vc_peer_create   allocates this rank's receive buffer for `world` ranks x `cap` records each and
 *                    writes its 64-byte CUDA IPC handle to handle_out (nullable);
 *   vc_peer_open     maps the other ranks' buffers from `handles` (world x 64 bytes, rank order;
 *                    one process per GPU: the handles travel once over torch.distributed / MPI);
 *   vc_peer_open_ptrs  same for contexts living in ONE process: bases[p] = vc_peer_buffer of rank p.
 * Per exchange (needs vc_classify_grid): vc_sites_post_peers is asynchronous; vc_sites_collect_peers
 * waits for all ranks' records, numbers the union (same result as vc_sites_import_global) and
 * returns the global site count.  `cap` and `world` must agree on every rank (the region offsets are computed from
 * them; a mismatch is detected at collect time: VC_ERR_INVALID).  A rank that has not posted within the time bound
 * (vc_peer_set_timeout, else the environment variable VC_PEER_TIMEOUT_MS, else 10 s) turns into VC_ERR_STATE -- the
 * post stays pending and vc_sites_collect_peers may simply be called again -- a rank with more than `cap` records
 * gets VC_ERR_NOMEM from vc_sites_post_peers before anything is stored; neither hangs. */
int vc_peer_create(vc_ctx* ctx, int world, int rank, int64_t cap, void* handle_out);
int vc_peer_open(vc_ctx* ctx, const void* handles);
int vc_peer_open_ptrs(vc_ctx* ctx, void* const* bases);
void* vc_peer_buffer(vc_ctx* ctx);
int vc_peer_close(vc_ctx* ctx);
int vc_peer_set_timeout(vc_ctx* ctx, int64_t milliseconds);
int vc_sites_post_peers(vc_ctx* ctx);
int vc_sites_collect_peers(vc_ctx* ctx, int64_t* n_all);

/* ---- stage 2: closest site ----------------------------------------------------------------------
 * a5: the operator the reference gets from ANN, annkSearch(k=1, eps=0)
 * (3rdparty/ann/src/kd_search.cpp:88-216; call sites src/voroinfo.cpp:344,362,
 * src/voxelapps.cpp:204,217,331,345), with ANNbruteForce's deterministic tie rule
 * (3rdparty/ann/src/brute.cpp:56-82): (squared distance, lowest site id).
 * Dense form: one query per grid vertex of the owned planes.  id_out int32, d2x4_out = 4*d^2 as an
 * exact uint32 (lattice sites; for arbitrary sites it is rounded and d2 comes from
 * vc_closest_points).  Outputs nullable, host or device. */
int vc_closest_grid(vc_ctx* ctx, int32_t* id_out, uint32_t* d2x4_out);
/* Arbitrary query points: drop-in for annkSearch(q, 1, &id, &d2, 0.0); q = n x 3 doubles,
 * d2 = squared distance in double. Host pointers. */
int vc_closest_points(vc_ctx* ctx, const double* q, int64_t n, int32_t* id, double* d2);

/* float32 form: drop-in for trimesh::KDtree::closest_to_pt(p, maxdist2)
 * (3rdparty/trimesh2/libsrc/KDtree.cc:252-292,523-545; call site estimateRadiiField,
 * src/exporters.cpp:629-636).  Distances are the tree's own: sqr(x0-q0) + sqr(x1-q1) + sqr(x2-q2) in
 * float32; only sites with d2 < max_d2 qualify (max_d2 <= 0 or inf: no limit); ties -> lowest id.
 * q = n x 3 floats; id = -1 and d2 = -1 when nothing qualifies.  Host pointers; d2 nullable. */
int vc_closest_points_f32(vc_ctx* ctx, const float* q, int64_t n, float max_d2, int32_t* id, float* d2);

/* Fixed-radius query: drop-in for annkFRSearch(q, sqRad, k, idx, dd, 0.0)
 * (3rdparty/ann/src/kd_fix_rad_search.cpp:58-189; call sites src/voxelapps.cpp:346,353).  sq_rad[i] is
 * the SQUARED radius, inclusive (dist <= sqRad, :172).  count[i] = number of sites in range (what the
 * reference's first call with k = 0 returns).  With off != NULL (int64[n+1], off[0] = 0, host) the
 * rows idx/d2[off[i]..off[i+1]) receive the closest sites in range ordered by (squared distance, site
 * id), padded with -1 / -1.0 -- the reference's second call with k = count.  count, idx, d2 nullable.
 * Host pointers. */
int vc_radius_search(vc_ctx* ctx, const double* q, const double* sq_rad, int64_t n, const int64_t* off, int32_t* count,
                     int32_t* idx, double* d2);

/* ---- stage 3: medial measures -------------------------------------------------------------------
 * a6/a7 on the dense grid (dictionary in SURVEY section 0): per grid vertex 3 edge cells (+x,+y,+z),
 * 3 face cells (xy,xz,yz) and the cube; lambda(2-set) = MeasureForMA::lambdaForFace
 * (include/measureforMA_imp.h:1-4, float32), 4-/8-sets = max over their grid edges
 * (src/voroinfo.cpp:1432-1574); cells with an outside vertex or leaving the grid report 0.
 * Layout: edge3 / face3 = 3 planes-of-volume each (SoA: [c][z][y][x]), cube and radius one each;
 * radius(v) = dist(site(id v), v), the m_r_per_v analogue (src/voroinfo.cpp:301-306).
 * All outputs nullable, host or device; planes [z0,z1). Needs vc_classify_grid + vc_closest_grid. */
int vc_cell_measures_grid(vc_ctx* ctx, float* edge3, float* face3, float* cube, float* radius);
/* a7 on a Voronoi complex: VoroInfo::computeFacesMeasure (src/voroinfo.cpp:1552-1574). Host ptrs. */
int vc_face_lambda(vc_ctx* ctx, const int32_t* site_pairs, int64_t nf, float* out);
/* a8: VoroInfo::computeInfoRelatedtoSites (src/voroinfo.cpp:298-318): r[v] = dist(site, v). */
int vc_vertex_radii(vc_ctx* ctx, const float* v_xyz, int64_t nv, const int32_t* site_of_v, float* r_out);
/* a7 aggregation: computeEdgesMeasure / computeVertexMeasure (src/voroinfo.cpp:1432-1538):
 * out[e] = max over items[off[e]..off[e+1]) of value[item] where valid[item] (valid nullable). */
int vc_segment_max(vc_ctx* ctx, const int32_t* off, const int32_t* items, int64_t n, const float* value,
                   int64_t nvalue, const uint8_t* valid, float* out);

/* ---- K6: the data-parallel part of the thinning --------------------------------
 * CellComplexThinning::prune (src/ccthin.cpp:201-271) before its first queue pop.  The FIFO loop itself
 * stays on the host in the reference's code: its order decides the result.
 * vc_ref_counts: cellcomplex::refCntPerVert / refCntPerEdge (src/cellcomplex.cpp:315-332) as a histogram:
 * out[b] = #{i : idx[i] == b}, e.g. idx = the 2*|E| edge end points, or the edge lists of all faces. */
int vc_ref_counts(vc_ctx* ctx, const int32_t* idx, int64_t n, int64_t nbins, int32_t* out);
/* vc_simple_pairs: the seeding scan (src/ccthin.cpp:246-270).  Edge e with edge_ref[e] == 1 whose first
 * incident face f = edge_face0[e] has face_to_remove[f] (nullable) or face_measure[f] < f_t gives the
 * face-edge pair (1, f, e); vertex v with vert_ref[v] == 1 whose first incident edge e = vert_edge0[v]
 * has edge_measure[e] < l_t gives the edge-vertex pair (0, e, v)  (include/ccthin.h:22-28).
 * pairs_out receives int32 triples (type, idx0, idx1) in the reference's push order -- all edges
 * ascending, then all vertices ascending -- at most `cap` of them; *npairs is the full count.
 * Inputs host or device; pairs_out host or device, nullable (count only). */
int vc_simple_pairs(vc_ctx* ctx, const int32_t* edge_ref, const int32_t* edge_face0, int64_t ne, const float* face_measure,
                    const uint8_t* face_to_remove, int64_t nf, float f_t, const int32_t* vert_ref, const int32_t* vert_edge0,
                    int64_t nv, const float* edge_measure, float l_t, int32_t* pairs_out, int64_t cap, int64_t* npairs);

/* ---- the whole hot path in one call ----------------------------------------------------------------
 * classify -> sites -> closest -> measures on the resident volume, all results left in device
 * memory (fetch with vc_download).  This is one "step" of bench.py. */
int vc_run_dense(vc_ctx* ctx, int64_t* nsites);
/* Stages 2 + 3 for the planes this ctx owns, once its sites are set (vc_extract_sites,
 * vc_sites_import_global or vc_set_sites): the closest-site passes and the measures, pipelined
 * over z chunks on the ctx's worker streams.  This is the second half of a multi-GPU step (the
 * first half being classify + site detection + the site exchange).  Results stay on the device. */
int vc_closest_and_measures(vc_ctx* ctx);
/* Pipeline shape of vc_run_dense / vc_run_dense_host / vc_closest_and_measures: number of worker
 * streams (1..16; 0 = everything on the main stream) and z planes per chunk (0 = automatic).
 * Results do not depend on either.  Defaults: 8 workers, automatic (env VC_WORKERS / VC_ZCHUNK). */
int vc_set_pipeline(vc_ctx* ctx, int workers, int zchunk);
typedef enum vc_array
{
    VC_ARR_INSIDE = 0,   /* uint8  [z][y][x]      */
    VC_ARR_ID = 1,       /* int32  [z][y][x]      */
    VC_ARR_D2X4 = 2,     /* uint32 [z][y][x]      */
    VC_ARR_EDGE3 = 3,    /* float  [3][z][y][x]   */
    VC_ARR_FACE3 = 4,    /* float  [3][z][y][x]   */
    VC_ARR_CUBE = 5,     /* float  [z][y][x]      */
    VC_ARR_RADIUS = 6    /* float  [z][y][x]      */
} vc_array;
/* copy a result array (owned planes) to dst (host or device) */
int vc_download(vc_ctx* ctx, int which, void* dst);
/* the same for the planes [za, zb) of the grid only (global z; inside this ctx's owned planes -- for VC_ARR_ID /
 * VC_ARR_D2X4 also the recomputed halo plane z1 of a slab); EDGE3 / FACE3 arrive as [3][zb-za][y][x] */
int vc_download_planes(vc_ctx* ctx, int which, int za, int zb, void* dst);
/* device pointer of a result array (valid until the next stage call / destroy) */
void* vc_device_ptr(vc_ctx* ctx, int which);
/* ---- stage 3, optional outputs: circumradius and object angle of each cell's closest-point set -----------------
 * north_star names them next to lambda; the reference only sketches a circumradius in comments
 * (src/voroinfo.cpp:1441-1443,1475-1479,1526-1530) and has no angle: PARITY UNPINNED, the definition is the builder's
 * (voxel_ma_b200/csrc/vc_circum.cu; restated in oracle/oracle.c, compared to 1e-12).  Per cell of the 7 anchored at a
 * vertex (edges +x +y +z, faces xy xz yz, cube; 0 unless all its vertices are inside), in float64, with P = the distinct
 * closest sites of the cell's vertices and m = the cell's centre:
 *   circum7[c]  radius of the smallest ball enclosing P;   angle7[c]  max over pairs of angle(p - m, q - m) / 2.
 * Planes [za, zb) of this ctx's owned planes; outputs [7][zb-za][y][x] doubles, host or device pointers. */
int vc_cell_circum_angle_grid(vc_ctx* ctx, int za, int zb, double* circum7, double* angle7);

/* ---- next row 8(f-4): the medial complex of the dense product, read off the id grid ----------------------
 * Replaces (for the DENSE product only) the TetGen Voronoi construction + inside filter of
 * src/highlevelalgo.cpp:503-529 / src/voroinfo.cpp:128-139,624-729: a grid edge (v, v + e_axis) whose end
 * vertices have different closest sites is crossed by the Voronoi face of those two sites; its dual is the quad
 * spanned by the centres of the 4 grid cubes around the edge.  A quad is emitted iff those 4 cubes exist and
 * all their vertices are inside (include/voroinfo_imp.h:26-34), so the quads form a closed cubical 2-complex in
 * the form cellcomplex::finalize / CellComplexThinning take (src/cellcomplex.cpp:364-491, src/ccthin.cpp:201-424).
 * Needs vc_classify_grid + closest sites (vc_closest_grid / vc_run_dense).  Records in ascending (z, y, x, axis):
 *   anchor[i]  linear index of v inside this ctx's owned planes,  axis[i] in {0,1,2} = +x,+y,+z,
 *   site_a/b   closest-site ids of v and v + e_axis,  lambda[i] = lambdaForFace(site_a, site_b) (float32).
 * The complex differs from TetGen's (unit quads instead of general polygons): compared on statistics only. */
int vc_medial_quads_count(vc_ctx* ctx, int64_t* nquads);
int vc_medial_quads(vc_ctx* ctx, int64_t cap, uint32_t* anchor, uint8_t* axis, int32_t* site_a, int32_t* site_b, float* lambda,
                    int64_t* nquads);

/* Host-buffer end-to-end step: H2D of the float32 volume, the hot path, D2H of every non-NULL
 * output, copies and kernels overlapped plane-chunk by plane-chunk.  Pinned buffers
 * (vc_host_alloc) reach PCIe speed. */
int vc_run_dense_host(vc_ctx* ctx, const float* vol, uint8_t* inside, int32_t* id, uint32_t* d2x4,
                      float* edge3, float* face3, float* cube, float* radius, int64_t* nsites);

/* ---- compact product: one record per INSIDE grid vertex ------------------------------------------
 * All the reference keeps of this front end lives on inside elements: Voronoi vertices tagged
 * outside are dropped on load (src/voroinfo.cpp:128-139) and a cell with an outside vertex is
 * invalid, measure 0 (include/voroinfo_imp.h:26-34, src/voroinfo.cpp:1460-1461,1513-1514) -- on the
 * dense grid all 7 measures anchored at an outside vertex are 0 by definition.  The compact product
 * is the occupancy bit rows plus, for the inside vertices in ascending linear index
 * v = x + nx*(y + ny*(z - z0)):  vert[i] = v, id[i], d2x4[i], lambda7[k*cap + i] (k = edge +x,+y,+z,
 * face xy,xz,yz, cube), radius[i] -- the same values as the dense planes at v.  44 B per inside
 * vertex cross the bus instead of 41 B per grid vertex.
 *   vc_compact_count    inside vertices of the owned planes (needs vc_classify_grid)
 *   vc_compact_records  gathers the records from the dense planes (needs vc_run_dense /
 *                       vc_closest_and_measures); outputs host or device, nullable; cap >= count
 *   vc_run_dense_host_compact  host volume in (float32 [z][y][x]), compact product out, all copies
 *                       inside the call: upload classified chunk by chunk as it lands, records of a
 *                       z chunk copied back while the next chunk computes.  inside_bits (nullable):
 *                       uint32 [nz*ny][nx/32+1], bit x&31 of word x>>5.  id_dense / d2x4_dense
 *                       (nullable): the full planes as well.  When vert, id, d2x4, lambda7, radius are the
 *                       rows 0, 1, 2, 3..9, 10 of ONE 11 x cap block of 32-bit words, a chunk's records
 *                       travel in a single copy.  cap < count -> VC_ERR_NOMEM with the
 *                       count in *n_inside.  On a slab ctx (one rank of a peer group, vc_peer_open):
 *                       vol = the slab's resident planes [max(z0-1,0), min(z1+1,nz)), outputs cover the
 *                       owned planes, the site records travel through the peer exchange, and every rank
 *                       of the group must make the call.
 *   vc_set_compact_mode how vc_run_dense_host_compact obtains the records: 1 = dense measure planes,
 *                       then gathered; 2 = computed directly per inside vertex (the 8 dense float
 *                       planes are then NOT produced by that call: vc_download of them fails with
 *                       VC_ERR_STATE); 0 = automatic (2 when at most 1/8 of the vertices are inside).
 *                       Same values either way. */
int vc_set_compact_mode(vc_ctx* ctx, int mode);
int vc_compact_count(vc_ctx* ctx, int64_t* n_inside);
int vc_compact_records(vc_ctx* ctx, int64_t cap, uint32_t* vert, int32_t* id, uint32_t* d2x4, float* lambda7, float* radius);
int vc_run_dense_host_compact(vc_ctx* ctx, const float* vol, uint32_t* inside_bits, int64_t cap, int64_t* n_inside,
                              uint32_t* vert, int32_t* id, uint32_t* d2x4, float* lambda7, float* radius, int32_t* id_dense,
                              uint32_t* d2x4_dense, int64_t* nsites);
/* the same call for an MRC mode 0 (signed byte) host volume: a quarter of the bytes cross the bus */
int vc_run_dense_host_compact_i8(vc_ctx* ctx, const int8_t* vol, uint32_t* inside_bits, int64_t cap, int64_t* n_inside,
                                 uint32_t* vert, int32_t* id, uint32_t* d2x4, float* lambda7, float* radius, int32_t* id_dense,
                                 uint32_t* d2x4_dense, int64_t* nsites);

/* ---- instrumentation --------------------------------------------------------------------------------
 * The reference's only instrumentation is struct timer around stages (include/commondefs.h:110-168);
 * here every kernel launch can be bracketed by CUDA events on the ctx stream. */
int vc_profile_enable(vc_ctx* ctx, int on);
int vc_profile_reset(vc_ctx* ctx);
/* number of distinct kernels seen; then per index: name, total ms, launches */
int vc_profile_count(vc_ctx* ctx);
int vc_profile_get(vc_ctx* ctx, int i, const char** name, double* total_ms, int64_t* launches);
/* kernels launched by this ctx since creation / last reset (counted whether or not profiling is on) */
int64_t vc_launch_count(const vc_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* VOXCORE_GPU_H */
