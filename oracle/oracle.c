/* oracle/oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A plain-C, CPU restatement of the reference's algorithm for the hot path (classify ->
 * boundary samples -> closest sample -> medial measures).  It is the checker the CUDA path is
 * diffed against; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load it.  The product path never calls into this file.
 *
 * PINNING: every function here is checked (tests/test_oracle_pinning.py, tests/golden/) against
 *   - the reference's own code compiled unmodified into oracle/_ref/libvoxref.so
 *     (Surfacer::extractBoundaryVts, SpaceConverter::voxTaggedAsInside, VoroInfo::tagVert,
 *      ANNkd_tree / ANNbruteForce, MeasureForMA::lambdaForFace, VoroInfo::compute*Measure), and
 *   - the only known-answer NN fixture in the reference tree, 3rdparty/ann/sample/sample.save.
 * The dense per-cell measure dictionary (orc_cell_measures_grid) has no reference counterpart as
 * an iteration space; its arithmetic (lambdaForFace + max-aggregation +
 * validity) is pinned through the functions above, its iteration space is builder-defined.
 * Circumradius / object angle (orc_cell_circum_angle_grid) exist nowhere in the reference but in comments:
 * builder-defined, "parity unpinned".
 *
 * Build: gcc -O2 -ffp-contract=off -fopenmp -shared -fPIC oracle/oracle.c -o oracle/_build/liboracle.so -lm
 * (no -march: no FMA contraction, so float results have one meaning).
 *
 * All dense arrays are x-fastest: index = x + nx*(y + ny*z)  (the MRC payload order,
 * 3rdparty/isosurface_tao/reader.h:232-251).
 */
#define _POSIX_C_SOURCE 200809L
#include <stdio.h>
#include <time.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define IDX(x, y, z) ((size_t)(x) + (size_t)nx * ((size_t)(y) + (size_t)ny * (size_t)(z)))

/* ---- a1/a2: getDataAt + get_occupancy_at_vox ---------------------------------------------------
 * include/spaceinfo.h:122-129: in bounds AND getDataAt > 0.0 (double compare; -0.0, 0.0, NaN are
 * outside); include/spaceinfo.h:108-115 bounds test; the MRC reader widens float32 samples to
 * double (reader.h:244-246), which preserves the sign test exactly. */
static inline int occ_f32(const float* vol, int nx, int ny, int nz, int x, int y, int z)
{
    if (x < 0 || x >= nx || y < 0 || y >= ny || z < 0 || z >= nz)
        return 0;
    return ((double)vol[IDX(x, y, z)] > 0.0) ? 1 : 0;
}

/* voxTaggedAsInside over the whole grid (include/spaceinfo.h:53-58). */
void orc_classify_grid_f32(const float* vol, int nx, int ny, int nz, uint8_t* inside)
{
    for (int z = 0; z < nz; ++z)
        for (int y = 0; y < ny; ++y)
            for (int x = 0; x < nx; ++x)
                inside[IDX(x, y, z)] = (uint8_t)occ_f32(vol, nx, ny, nz, x, y, z);
}

/* Same on Tao's in-memory layout double[x*ny*nz + y*nz + z] (3rdparty/isosurface_tao/volume.h:217-224);
 * output is still x-fastest. */
void orc_classify_grid_f64_zfast(const double* vol, int nx, int ny, int nz, uint8_t* inside)
{
    for (int z = 0; z < nz; ++z)
        for (int y = 0; y < ny; ++y)
            for (int x = 0; x < nx; ++x)
                inside[IDX(x, y, z)] = vol[((size_t)x * ny + y) * nz + z] > 0.0 ? 1 : 0;
}

/* ---- a4: VoroInfo::tagVert (src/voroinfo.cpp:447-454) ------------------------------------------
 * q = M * p with the double 4x4 (column-major, homogeneous divide) and a cast of each component
 * back to float (3rdparty/trimesh2/include/XForm.h:479-489); voxel = (int)std::round(q[i]) on the
 * FLOAT value, half away from zero (include/spaceinfo.h:93-105); then a2.  M == NULL is the
 * identity, which is what tagVert always passes. */
void orc_classify_points(const uint8_t* inside, int nx, int ny, int nz, const float* xyz, int64_t n,
                         const double* M, uint8_t* out)
{
    static const double I[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
    const double* xf = M ? M : I;
    for (int64_t i = 0; i < n; ++i)
    {
        double v0 = xyz[3 * i], v1 = xyz[3 * i + 1], v2 = xyz[3 * i + 2];
        double h = 1 / (xf[3] * v0 + xf[7] * v1 + xf[11] * v2 + xf[15]);
        float q0 = (float)(h * (xf[0] * v0 + xf[4] * v1 + xf[8] * v2 + xf[12]));
        float q1 = (float)(h * (xf[1] * v0 + xf[5] * v1 + xf[9] * v2 + xf[13]));
        float q2 = (float)(h * (xf[2] * v0 + xf[6] * v1 + xf[10] * v2 + xf[14]));
        int x = (int)roundf(q0), y = (int)roundf(q1), z = (int)roundf(q2);
        int in = !(x < 0 || x >= nx || y < 0 || y >= ny || z < 0 || z >= nz);
        out[i] = (uint8_t)(in ? inside[IDX(x, y, z)] : 0);
    }
}

/* ---- a3: Surfacer::extractBoundaryVts (src/surfacing.cpp:223-321) ------------------------------
 * Scan voxels x outer / y / z inner (:275-284); for each of the 6 neighbours in the order
 * -x,+x,-y,+y,-z,+z (include/surfacing.h:170-178) whose occupancy differs (out of bounds = 0,
 * include/spaceinfo.h:125), visit the 4 corners of the shared face in the slot order of
 * include/surfacing.h:184-190 and append a corner the first time it is seen (:254-268).
 * Corner slots c0..c7 relative to the voxel centre: include/surfacing.h:97-119.
 * The reference de-duplicates with unordered_map<ivec3,int> keyed by the doubled corner id
 * (include/surfacing.h:121-136); a dense "seen" array over the (nx+1)(ny+1)(nz+1) corner lattice
 * is the same set semantics.
 * Returns the number of sites; writes at most cap of them (float32, half-integer coordinates). */
static const int NB_OFF[6][3] = {{-1, 0, 0}, {1, 0, 0}, {0, -1, 0}, {0, 1, 0}, {0, 0, -1}, {0, 0, 1}};
static const int CORNERS_WRT_NB[6][4] = {{2, 3, 7, 6}, {0, 1, 5, 4}, {4, 6, 2, 0},
                                         {1, 3, 7, 5}, {0, 2, 3, 1}, {5, 7, 6, 4}};
/* corner slot -> (+1 means +0.5, 0 means -0.5) per axis */
static const int CORNER_SIGN[8][3] = {{1, 0, 0}, {1, 1, 0}, {0, 0, 0}, {0, 1, 0},
                                      {1, 0, 1}, {1, 1, 1}, {0, 0, 1}, {0, 1, 1}};

static double orc_now(void)
{
    struct timespec t;
    clock_gettime(CLOCK_MONOTONIC, &t);
    return (double)t.tv_sec + 1e-9 * (double)t.tv_nsec;
}

int64_t orc_extract_sites(const uint8_t* inside, int nx, int ny, int nz, float* out_xyz, int64_t cap)
{
    const int trace = getenv("ORC_TRACE") != NULL;
    double t_0 = orc_now();
    size_t cx = (size_t)nx + 1, cy = (size_t)ny + 1, cz = (size_t)nz + 1;
    /* The scan below runs x outer / z inner like the reference's, over Tao's z-fastest layout
     * (3rdparty/isosurface_tao/volume.h:217-224): a transposed copy of the flags and a z-fastest `seen` array keep
     * the inner loop on consecutive bytes (the same set semantics; 160 s -> seconds at 1024^3). */
    uint8_t* zf = (uint8_t*)malloc((size_t)nx * ny * nz);
#pragma omp parallel for collapse(2) schedule(static)
    for (int j = 0; j < ny; ++j)
        for (int k0 = 0; k0 < nz; k0 += 64)
        {
            uint8_t tile[64][64]; /* [k][i]: rows read along x, columns written along z */
            const int kn = nz - k0 < 64 ? nz - k0 : 64;
            for (int i0 = 0; i0 < nx; i0 += 64)
            {
                const int in = nx - i0 < 64 ? nx - i0 : 64;
                for (int k = 0; k < kn; ++k)
                    memcpy(tile[k], &inside[IDX(i0, j, k0 + k)], (size_t)in);
                for (int i = 0; i < in; ++i)
                {
                    uint8_t* d = &zf[((size_t)(i0 + i) * ny + j) * nz + k0];
                    for (int k = 0; k < kn; ++k)
                        d[k] = tile[k][i];
                }
            }
        }
#define ZF(i, j, k) zf[((size_t)(i) * ny + (j)) * nz + (k)]
    /* Rows (i, j, all k) in which no voxel differs from any of its 6 neighbours emit nothing, so the scan may step
     * over them: rowflag marks the others.  A row is quiet iff it is constant, equals its 4 neighbour rows, and --
     * where a neighbour is out of bounds (read as 0, include/spaceinfo.h:125) -- is all 0.  (Same output, same
     * order: only rows that would not have emitted are skipped; the marking runs on all cores.) */
    uint8_t* rowflag = (uint8_t*)malloc((size_t)nx * ny);
#pragma omp parallel for schedule(static)
    for (int i = 0; i < nx; ++i)
        for (int j = 0; j < ny; ++j)
        {
            const uint8_t* r = &ZF(i, j, 0);
            int quiet = nz < 2 || memcmp(r, r + 1, (size_t)nz - 1) == 0; /* constant along z */
            const int border = i == 0 || i == nx - 1 || j == 0 || j == ny - 1;
            if (quiet && (border || nz >= 1) && r[0] != 0) /* the z ends (k = -1, k = nz) and any side border read 0 */
                quiet = 0;
            if (quiet && i > 0)
                quiet = memcmp(r, &ZF(i - 1, j, 0), nz) == 0;
            if (quiet && i < nx - 1)
                quiet = memcmp(r, &ZF(i + 1, j, 0), nz) == 0;
            if (quiet && j > 0)
                quiet = memcmp(r, &ZF(i, j - 1, 0), nz) == 0;
            if (quiet && j < ny - 1)
                quiet = memcmp(r, &ZF(i, j + 1, 0), nz) == 0;
            rowflag[(size_t)i * ny + j] = (uint8_t)!quiet;
        }
    double t_1 = orc_now();
    uint8_t* seen = (uint8_t*)calloc(cx * cy * cz, 1);
    /* The scan runs on all cores in bands of x-planes and is put back together in scan order.  "First encounter
     * wins" only couples two bands through the corner plane they share (px = the upper band's first x): the upper
     * band keeps its own flags for that plane, and the merge drops what the lower band had already emitted there --
     * de-duplication only ever REMOVES entries, so every band's list is the sequential scan's list for its voxels
     * and the concatenation is the reference's order.
     * Within a live row, act[k] != 0 iff voxel k differs from one of its 6 neighbours (out of bounds reads 0): the
     * others emit nothing and are stepped over.  The marking is a flat byte loop the compiler vectorises; the
     * emitting loop is the reference's, in its order. */
    const int BW = 4;
    const int nband = (nx + BW - 1) / BW;
    typedef struct
    {
        uint64_t* v;
        int64_t n, cap;
    } keylist;
    keylist* lists = (keylist*)calloc((size_t)nband, sizeof(keylist));
    const uint8_t* zrow = (const uint8_t*)calloc((size_t)nz + 1, 1);
#pragma omp parallel
    {
        uint8_t* act = (uint8_t*)malloc((size_t)nz + 1);
        uint8_t* lo = (uint8_t*)malloc(cy * cz);
#pragma omp for schedule(dynamic, 1)
        for (int band = 0; band < nband; ++band)
        {
            const int ia = band * BW, ib = ia + BW < nx ? ia + BW : nx;
            keylist* L = &lists[band];
            int lo_clean = 0;
            for (int i = ia; i < ib; ++i)
                for (int j = 0; j < ny; ++j)
                {
                    if (!rowflag[(size_t)i * ny + j])
                        continue;
                    if (!lo_clean)
                    {
                        memset(lo, 0, cy * cz);
                        lo_clean = 1;
                    }
                    const uint8_t* r = &ZF(i, j, 0);
                    {
                        const uint8_t* xm = i > 0 ? &ZF(i - 1, j, 0) : zrow;
                        const uint8_t* xp = i < nx - 1 ? &ZF(i + 1, j, 0) : zrow;
                        const uint8_t* ym = j > 0 ? &ZF(i, j - 1, 0) : zrow;
                        const uint8_t* yp = j < ny - 1 ? &ZF(i, j + 1, 0) : zrow;
                        for (int k = 0; k < nz; ++k)
                            act[k] = (uint8_t)((r[k] ^ xm[k]) | (r[k] ^ xp[k]) | (r[k] ^ ym[k]) | (r[k] ^ yp[k]));
                        for (int k = 1; k < nz; ++k)
                            act[k] |= (uint8_t)(r[k] ^ r[k - 1]);
                        for (int k = 0; k + 1 < nz; ++k)
                            act[k] |= (uint8_t)(r[k] ^ r[k + 1]);
                        act[0] |= r[0];
                        if (nz > 1)
                            act[nz - 1] |= r[nz - 1];
                    }
                    for (int k = 0; k < nz; ++k)
                    {
                        if (!act[k])
                            continue;
                        int cur = r[k];
                        for (int o = 0; o < 6; ++o)
                        {
                            int a = i + NB_OFF[o][0], b = j + NB_OFF[o][1], c = k + NB_OFF[o][2];
                            int nb = (a < 0 || a >= nx || b < 0 || b >= ny || c < 0 || c >= nz) ? 0 : ZF(a, b, c);
                            if (nb == cur)
                                continue;
                            for (int ii = 0; ii < 4; ++ii)
                            {
                                int ci = CORNERS_WRT_NB[o][ii];
                                int px = i + CORNER_SIGN[ci][0], py = j + CORNER_SIGN[ci][1],
                                    pz = k + CORNER_SIGN[ci][2]; /* corner lattice index: coord = p - 0.5 */
                                size_t key = ((size_t)px * cy + (size_t)py) * cz + (size_t)pz;
                                uint8_t* flag = (band > 0 && px == ia) ? &lo[(size_t)py * cz + (size_t)pz] : &seen[key];
                                if (*flag)
                                    continue;
                                *flag = 1;
                                if (L->n == L->cap)
                                {
                                    L->cap = L->cap ? 2 * L->cap : 1024;
                                    L->v = (uint64_t*)realloc(L->v, (size_t)L->cap * sizeof(uint64_t));
                                }
                                L->v[L->n++] = (uint64_t)key;
                            }
                        }
                    }
                }
        }
        free(lo);
        free(act);
    }
    double t_2 = orc_now();
    int64_t n = 0;
    for (int band = 0; band < nband; ++band)
    {
        const keylist* L = &lists[band];
        const uint64_t shared_lo = (uint64_t)band * BW * cy * cz, shared_hi = shared_lo + cy * cz;
        for (int64_t e = 0; e < L->n; ++e)
        {
            const uint64_t key = L->v[e];
            if (band > 0 && key >= shared_lo && key < shared_hi && seen[key])
                continue; /* the band below met this corner first */
            if (n < cap)
            {
                out_xyz[3 * n] = (float)(key / (cy * cz)) - 0.5f;
                out_xyz[3 * n + 1] = (float)(key / cz % cy) - 0.5f;
                out_xyz[3 * n + 2] = (float)(key % cz) - 0.5f;
            }
            ++n;
        }
        free(L->v);
    }
    free(lists);
    free((void*)zrow);
#undef ZF
    if (trace)
        fprintf(stderr, "[orc_extract_sites] transpose + row marks %.2fs, scan %.2fs, merge %.2fs\n", t_1 - t_0,
                t_2 - t_1, orc_now() - t_2);
    free(seen);
    free(rowflag);
    free(zf);
    return n;
}

/* ---- a5: exact 1-NN, the contract = ANNbruteForce::annkSearch(k=1) -----------------------------
 * 3rdparty/ann/src/brute.cpp:56-82 scans ids ascending; ANNmin_k::insert
 * (3rdparty/ann/src/pr_queue_k.h:109-127) only displaces strictly larger keys, so among equal
 * squared distances the LOWEST id wins.  Distance = annDist = sum over dims of (q-p)^2 in double
 * (3rdparty/ann/src/ANN.cpp:43-58; ANNcoord/ANNdist are double, include/ANN/ANN.h:160-161),
 * accumulated in dimension order.  The kd-tree (kd_search.cpp:88-216) returns the same d2 but
 * its tie choice depends on traversal order; d2 is compared against it, ids
 * against this function. */
void orc_closest_points(const double* sites, int64_t ns, int dim, const double* q, int64_t nq,
                        int32_t* idx, double* d2)
{
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < nq; ++i)
    {
        double best = INFINITY;
        int32_t bi = -1;
        const double* qq = q + (size_t)i * dim;
        for (int64_t s = 0; s < ns; ++s)
        {
            const double* p = sites + (size_t)s * dim;
            double d = 0;
            for (int k = 0; k < dim; ++k)
            {
                double t = qq[k] - p[k];
                d = d + t * t;
            }
            if (d < best)
            {
                best = d;
                bi = (int32_t)s;
            }
        }
        idx[i] = bi;
        if (d2)
            d2[i] = best;
    }
}

/* float32 nearest point: what trimesh::KDtree::closest_to_pt(p, maxdist2) answers
 * (3rdparty/trimesh2/libsrc/KDtree.cc:252-292 leaf test `myd2 < closest_d2`, :523-545 entry with
 * closest_d2 = maxdist2; distance :28-33 = sqr(x0-y0) + sqr(x1-y1) + sqr(x2-y2) in float, x = tree
 * point).  The caller (estimateRadiiField, src/exporters.cpp:629-636) only uses the DISTANCE of the
 * returned point, so the tree's traversal-dependent choice among equidistant points does not show;
 * this restatement reports the lowest index.  max_d2 <= 0 or inf: no limit.  idx = -1, d2 = -1 when
 * no point has d2 < max_d2. */
void orc_closest_points_f32(const float* pts, int64_t n, const float* q, int64_t nq, float max_d2, int32_t* idx, float* d2)
{
    const float lim = (max_d2 > 0.0f && max_d2 < INFINITY) ? max_d2 : INFINITY;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < nq; ++i)
    {
        float best = lim;
        int32_t bi = -1;
        const float* y = q + 3 * (size_t)i;
        for (int64_t s = 0; s < n; ++s)
        {
            const float* x = pts + 3 * (size_t)s;
            float t = x[0] - y[0];
            float d = t * t;
            t = x[1] - y[1];
            d = d + t * t;
            t = x[2] - y[2];
            d = d + t * t;
            if (d < best)
            {
                best = d;
                bi = (int32_t)s;
            }
        }
        idx[i] = bi;
        if (d2)
            d2[i] = bi < 0 ? -1.0f : best;
    }
}

/* Dense query set: one query per grid vertex (integer lattice point) against float32 sites widened
 * to double exactly as the reference widens them for ANN (src/voroinfo.cpp:336-341).  Outputs the
 * id and 4*d2 as an exact integer (sites on the half-integer lattice make 4*d2 integral; for
 * arbitrary sites d2x4 may be NULL and d2 is returned in double). */
void orc_closest_grid(const float* sites_xyz, int64_t ns, int nx, int ny, int nz, int z0, int z1,
                      int32_t* id_out, uint32_t* d2x4_out, double* d2_out)
{
    double* S = (double*)malloc((size_t)ns * 3 * sizeof(double));
    for (int64_t i = 0; i < ns * 3; ++i)
        S[i] = (double)sites_xyz[i];
#pragma omp parallel for schedule(dynamic, 64) collapse(2)
    for (int z = z0; z < z1; ++z)
        for (int y = 0; y < ny; ++y)
            for (int x = 0; x < nx; ++x)
            {
                double best = INFINITY;
                int32_t bi = -1;
                for (int64_t s = 0; s < ns; ++s)
                {
                    double t0 = (double)x - S[3 * s], t1 = (double)y - S[3 * s + 1],
                           t2 = (double)z - S[3 * s + 2];
                    double d = 0;
                    d = d + t0 * t0;
                    d = d + t1 * t1;
                    d = d + t2 * t2;
                    if (d < best)
                    {
                        best = d;
                        bi = (int32_t)s;
                    }
                }
                size_t o = (size_t)x + (size_t)nx * ((size_t)y + (size_t)ny * (size_t)(z - z0));
                id_out[o] = bi;
                if (d2x4_out)
                    d2x4_out[o] = (uint32_t)llround(4.0 * best);
                if (d2_out)
                    d2_out[o] = best;
            }
    free(S);
}

/* The same contract (ANNbruteForce: lowest id among the sites at the minimum squared distance) at a SAMPLE
 * of grid vertices of a grid too large for orc_closest_grid: all sites are scanned for every sampled
 * vertex, in id order, in double, exactly as above; the site coordinates are split into three arrays
 * only so that the compiler can vectorise the scan.  q = integer vertex coordinates (x, y, z). */
void orc_closest_grid_sample(const float* sites_xyz, int64_t ns, const int32_t* q, int64_t nq, int32_t* id_out,
                             uint32_t* d2x4_out)
{
    double* X = (double*)malloc((size_t)ns * sizeof(double));
    double* Y = (double*)malloc((size_t)ns * sizeof(double));
    double* Z = (double*)malloc((size_t)ns * sizeof(double));
    for (int64_t s = 0; s < ns; ++s)
    {
        X[s] = (double)sites_xyz[3 * s];
        Y[s] = (double)sites_xyz[3 * s + 1];
        Z[s] = (double)sites_xyz[3 * s + 2];
    }
#pragma omp parallel for schedule(dynamic, 16)
    for (int64_t i = 0; i < nq; ++i)
    {
        const double x = q[3 * i], y = q[3 * i + 1], z = q[3 * i + 2];
        double best = INFINITY;
        /* pass 1: the minimum (vectorisable), pass 2: the first site that attains it = the lowest id */
        for (int64_t s = 0; s < ns; ++s)
        {
            double t0 = x - X[s], t1 = y - Y[s], t2 = z - Z[s];
            double d = 0;
            d = d + t0 * t0;
            d = d + t1 * t1;
            d = d + t2 * t2;
            best = d < best ? d : best;
        }
        int32_t bi = -1;
        for (int64_t s = 0; s < ns; ++s)
        {
            double t0 = x - X[s], t1 = y - Y[s], t2 = z - Z[s];
            double d = 0;
            d = d + t0 * t0;
            d = d + t1 * t1;
            d = d + t2 * t2;
            if (d == best)
            {
                bi = (int32_t)s;
                break;
            }
        }
        id_out[i] = bi;
        d2x4_out[i] = (uint32_t)llround(4.0 * best);
    }
    free(X);
    free(Y);
    free(Z);
}

/* ---- a6: MeasureForMA::lambdaForFace = trimesh::dist (include/measureforMA_imp.h:1-4,
 * 3rdparty/trimesh2/include/Vec.h:1128-1143): float32, d2 = sqr(b0-a0); d2 += sqr(b_i-a_i);
 * sqrt in float. */
static inline float lambda_f(const float* a, const float* b)
{
    float t = b[0] - a[0];
    float d2 = t * t;
    t = b[1] - a[1];
    d2 += t * t;
    t = b[2] - a[2];
    d2 += t * t;
    return sqrtf(d2);
}

/* a7, face form: VoroInfo::computeFacesMeasure (src/voroinfo.cpp:1552-1574):
 * lambda(f) = lambdaForFace(site[pair[0]], site[pair[1]]). */
void orc_face_lambda(const float* sites_xyz, const int32_t* site_pairs, int64_t nf, float* out)
{
    for (int64_t f = 0; f < nf; ++f)
        out[f] = lambda_f(sites_xyz + 3 * (size_t)site_pairs[2 * f],
                          sites_xyz + 3 * (size_t)site_pairs[2 * f + 1]);
}

/* a8: VoroInfo::computeInfoRelatedtoSites (src/voroinfo.cpp:298-318): r[v] = dist(site, v) with
 * the argument order (site_p, v_p); the caller supplies the site each vertex ends up with. */
void orc_vertex_radii(const float* sites_xyz, const float* v_xyz, int64_t nv, const int32_t* site_of_v,
                      float* r_out)
{
    for (int64_t v = 0; v < nv; ++v)
        r_out[v] = site_of_v[v] < 0 ? 0.0f
                                    : lambda_f(sites_xyz + 3 * (size_t)site_of_v[v], v_xyz + 3 * v);
}

/* a7, aggregation form: computeEdgesMeasure / computeVertexMeasure (src/voroinfo.cpp:1490-1538,
 * 1432-1488): max of the face lambdas over the incident VALID faces, 0 when there are none.
 * CSR adjacency: element e owns items[off[e] .. off[e+1]). */
void orc_segment_max(const int32_t* off, const int32_t* items, int64_t n, const float* face_lambda,
                     const uint8_t* face_valid, float* out)
{
    for (int64_t e = 0; e < n; ++e)
    {
        float m = 0.0f;
        for (int32_t k = off[e]; k < off[e + 1]; ++k)
        {
            int32_t f = items[k];
            if (face_valid && !face_valid[f])
                continue;
            m = face_lambda[f] > m ? face_lambda[f] : m; /* std::max(m, f_lmd) */
        }
        out[e] = m;
    }
}

/* ---- dense cell measures: the grid-cell <-> Voronoi-cell dictionary of SURVEY section 0 ---------
 * 7 cells anchored at each grid vertex v=(x,y,z): edges +x,+y,+z; faces xy,xz,yz; the cube.
 *   lambda_edge(u,w) = lambdaForFace(s(id u), s(id w))                      (src/voroinfo.cpp:1564-1568)
 *   lambda_face      = max over the face's 4 grid edges                     (:1505-1523)
 *   lambda_cube      = max over the cube's 12 grid edges                    (:1447-1470)
 * A cell is valid iff all its vertices are inside (computeFaceValidity, include/voroinfo_imp.h:26-34);
 * invalid cells and cells that would leave the grid report 0 (:1460-1461, 1513-1514).
 * radius(v) = dist(s(id v), v) in float (the m_r_per_v analogue, :301-306).
 * Planes are SoA: edge3[c][v], face3[c][v] with c = 0,1,2 (edges +x,+y,+z; faces xy,xz,yz).
 * id / inside cover planes [z0, z1 + 1) when z1 < nz (one halo plane), outputs cover [z0, z1). */
void orc_cell_measures_grid(const float* sites_xyz, const int32_t* id, const uint8_t* inside, int nx,
                            int ny, int nz, int z0, int z1, float* edge3, float* face3, float* cube,
                            float* radius)
{
    int zh = z1 < nz ? z1 + 1 : z1; /* planes available */
    size_t plane = (size_t)nx * ny, nv = plane * (size_t)(z1 - z0);
#define L(x, y, z) ((size_t)(x) + (size_t)nx * ((size_t)(y) + (size_t)ny * (size_t)((z) - z0)))
#pragma omp parallel for schedule(static)
    for (int z = z0; z < z1; ++z)
        for (int y = 0; y < ny; ++y)
            for (int x = 0; x < nx; ++x)
            {
                size_t o = L(x, y, z);
                /* the 8 cube vertices, bit0 = +x, bit1 = +y, bit2 = +z */
                const float* s[8];
                int in[8], ok[8];
                for (int c = 0; c < 8; ++c)
                {
                    int xx = x + (c & 1), yy = y + ((c >> 1) & 1), zz = z + ((c >> 2) & 1);
                    ok[c] = xx < nx && yy < ny && zz < zh && zz < nz;
                    in[c] = ok[c] ? inside[L(xx, yy, zz)] : 0;
                    s[c] = ok[c] ? sites_xyz + 3 * (size_t)id[L(xx, yy, zz)] : NULL;
                }
                /* 12 cube edges as vertex pairs; the first endpoint is the lower vertex */
                static const int E[12][2] = {{0, 1}, {2, 3}, {4, 5}, {6, 7},  /* x edges */
                                             {0, 2}, {1, 3}, {4, 6}, {5, 7},  /* y edges */
                                             {0, 4}, {1, 5}, {2, 6}, {3, 7}}; /* z edges */
                float le[12];
                for (int e = 0; e < 12; ++e)
                    le[e] = (ok[E[e][0]] && ok[E[e][1]]) ? lambda_f(s[E[e][0]], s[E[e][1]]) : 0.0f;
#define MAX2(a, b) ((a) > (b) ? (a) : (b))
                edge3[0 * nv + o] = (in[0] && in[1]) ? le[0] : 0.0f;
                edge3[1 * nv + o] = (in[0] && in[2]) ? le[4] : 0.0f;
                edge3[2 * nv + o] = (in[0] && in[4]) ? le[8] : 0.0f;
                /* faces: xy = {0,1,2,3}, xz = {0,1,4,5}, yz = {0,2,4,6} */
                float fxy = MAX2(MAX2(le[0], le[1]), MAX2(le[4], le[5]));
                float fxz = MAX2(MAX2(le[0], le[2]), MAX2(le[8], le[9]));
                float fyz = MAX2(MAX2(le[4], le[6]), MAX2(le[8], le[10]));
                face3[0 * nv + o] = (in[0] && in[1] && in[2] && in[3]) ? fxy : 0.0f;
                face3[1 * nv + o] = (in[0] && in[1] && in[4] && in[5]) ? fxz : 0.0f;
                face3[2 * nv + o] = (in[0] && in[2] && in[4] && in[6]) ? fyz : 0.0f;
                float m = 0.0f;
                for (int e = 0; e < 12; ++e)
                    m = MAX2(m, le[e]);
                int all = 1;
                for (int c = 0; c < 8; ++c)
                    all = all && in[c];
                cube[o] = all ? m : 0.0f;
                if (radius)
                {
                    float v[3] = {(float)x, (float)y, (float)z};
                    radius[o] = lambda_f(s[0], v);
                }
            }
#undef L
#undef MAX2
}

/* ---- stage 3 extras: circumradius and object angle of a cell's closest-point set ---------------------------------
 * PARITY UNPINNED: the reference has neither (a circumradius only in comments, src/voroinfo.cpp:1441-1443,1475-1479,
 * 1526-1530); the definition is the builder's, stated in include/voxcore_gpu.h (vc_cell_circum_angle_grid): per cell of
 * the 7 anchored at a vertex, valid iff all its vertices are inside (else 0), P = distinct closest sites of its vertices,
 * m = its centre; circumradius = radius of the smallest ball enclosing P, angle = max over pairs of angle(p-m, q-m)/2.
 * The smallest enclosing ball is written here the textbook way (Welzl's recursion on the boundary set), not as the
 * kernel's enumeration, so the two only share the definition.  id / inside: planes [z0, min(z1+1, nz)); out [7][z1-z0][y][x]. */
typedef struct
{
    double c[3], r2;
    int ok;
} orc_ball;
static orc_ball orc_ball_of(const double (*b)[3], int nb)
{
    orc_ball B = {{0, 0, 0}, -1.0, 1};
    if (nb == 0)
        return B;
    if (nb == 1)
    {
        memcpy(B.c, b[0], sizeof B.c);
        B.r2 = 0;
        return B;
    }
    /* centre = b0 + sum_k l_k (b_k - b0) with  2 (b_j - b0).(centre - b0) = |b_j - b0|^2 : Gram system, Gaussian elimination */
    double A[3][4];
    int m = nb - 1;
    for (int j = 0; j < m; ++j)
    {
        for (int k = 0; k < m; ++k)
        {
            double d = 0;
            for (int t = 0; t < 3; ++t)
                d += (b[j + 1][t] - b[0][t]) * (b[k + 1][t] - b[0][t]);
            A[j][k] = 2 * d;
        }
        double d = 0;
        for (int t = 0; t < 3; ++t)
            d += (b[j + 1][t] - b[0][t]) * (b[j + 1][t] - b[0][t]);
        A[j][m] = d;
    }
    for (int col = 0; col < m; ++col)
    {
        int piv = col;
        for (int r = col + 1; r < m; ++r)
            if (fabs(A[r][col]) > fabs(A[piv][col]))
                piv = r;
        if (fabs(A[piv][col]) < 1e-12)
        {
            B.ok = 0; /* affinely dependent boundary set */
            return B;
        }
        for (int k = 0; k <= m; ++k)
        {
            double t = A[col][k];
            A[col][k] = A[piv][k];
            A[piv][k] = t;
        }
        for (int r = 0; r < m; ++r)
            if (r != col)
            {
                double f = A[r][col] / A[col][col];
                for (int k = col; k <= m; ++k)
                    A[r][k] -= f * A[col][k];
            }
    }
    double off[3] = {0, 0, 0};
    for (int j = 0; j < m; ++j)
        for (int t = 0; t < 3; ++t)
            off[t] += A[j][m] / A[j][j] * (b[j + 1][t] - b[0][t]);
    B.r2 = off[0] * off[0] + off[1] * off[1] + off[2] * off[2];
    for (int t = 0; t < 3; ++t)
        B.c[t] = b[0][t] + off[t];
    return B;
}
static orc_ball orc_welzl(const double (*p)[3], int n, double (*b)[3], int nb)
{
    if (n == 0 || nb == 4)
        return orc_ball_of((const double (*)[3])b, nb);
    orc_ball B = orc_welzl(p, n - 1, b, nb);
    const double* q = p[n - 1];
    double d2 = 0;
    for (int t = 0; t < 3; ++t)
        d2 += (q[t] - B.c[t]) * (q[t] - B.c[t]);
    if (B.ok && B.r2 >= 0 && d2 <= B.r2 * (1.0 + 1e-12) + 1e-12)
        return B;
    memcpy(b[nb], q, 3 * sizeof(double));
    return orc_welzl(p, n - 1, b, nb + 1);
}
void orc_cell_circum_angle_grid(const float* sites_xyz, const int32_t* id, const uint8_t* inside, int nx, int ny, int nz, int z0,
                                int z1, double* circ7, double* ang7)
{
    static const int CELLS[7][8] = {{0, 1, -1}, {0, 2, -1}, {0, 4, -1}, {0, 1, 2, 3, -1}, {0, 1, 4, 5, -1}, {0, 2, 4, 6, -1},
                                    {0, 1, 2, 3, 4, 5, 6, 7}};
    static const int NV[7] = {2, 2, 2, 4, 4, 4, 8};
    const int zh = z1 < nz ? z1 + 1 : z1;
    const size_t plane = (size_t)nx * ny, nv = plane * (size_t)(z1 - z0);
#pragma omp parallel for schedule(dynamic, 4)
    for (int z = z0; z < z1; ++z)
        for (int y = 0; y < ny; ++y)
            for (int x = 0; x < nx; ++x)
            {
                const size_t o = (size_t)x + (size_t)nx * ((size_t)y + (size_t)ny * (size_t)(z - z0));
                for (int cell = 0; cell < 7; ++cell)
                {
                    double P[8][3], m[3] = {0, 0, 0};
                    int ids[8], n = 0, valid = 1;
                    for (int k = 0; k < NV[cell]; ++k)
                    {
                        const int c = CELLS[cell][k];
                        const int xx = x + (c & 1), yy = y + ((c >> 1) & 1), zz = z + (c >> 2);
                        m[0] += xx, m[1] += yy, m[2] += zz;
                        if (!(xx < nx && yy < ny && zz < zh))
                        {
                            valid = 0;
                            continue;
                        }
                        const size_t q = (size_t)xx + (size_t)nx * ((size_t)yy + (size_t)ny * (size_t)(zz - z0));
                        if (!inside[q])
                        {
                            valid = 0;
                            continue;
                        }
                        int seen = 0;
                        for (int j = 0; j < n; ++j)
                            seen |= ids[j] == id[q];
                        if (!seen)
                        {
                            ids[n] = id[q];
                            for (int t = 0; t < 3; ++t)
                                P[n][t] = (double)sites_xyz[3 * (size_t)id[q] + t];
                            ++n;
                        }
                    }
                    double r = 0, a = 0;
                    if (valid && n > 1)
                    {
                        double b[4][3];
                        orc_ball B = orc_welzl((const double (*)[3])P, n, b, 0);
                        r = sqrt(B.r2 > 0 ? B.r2 : 0);
                        for (int t = 0; t < 3; ++t)
                            m[t] /= NV[cell];
                        for (int i = 0; i < n; ++i)
                            for (int j = i + 1; j < n; ++j)
                            {
                                double uu = 0, vv = 0, uv = 0;
                                for (int t = 0; t < 3; ++t)
                                {
                                    const double u = P[i][t] - m[t], v = P[j][t] - m[t];
                                    uu += u * u, vv += v * v, uv += u * v;
                                }
                                if (uu <= 0 || vv <= 0)
                                    continue;
                                double cs = uv / sqrt(uu * vv);
                                cs = cs > 1 ? 1 : (cs < -1 ? -1 : cs);
                                const double h = 0.5 * acos(cs);
                                a = h > a ? h : a;
                            }
                    }
                    circ7[(size_t)cell * nv + o] = r;
                    ang7[(size_t)cell * nv + o] = a;
                }
            }
}

/* ---- next row 8(f-4): dual quads of the crossing grid edges (the medial complex of the dense product) ----------
 * PARITY UNPINNED for the complex (the reference's complex is TetGen's Voronoi diagram, src/highlevelalgo.cpp:503-529;
 * this cubical one can only be compared on statistics); the rules are the builder's, stated in
 * include/voxcore_gpu.h (vc_medial_quads): a grid edge (v, v + e_axis) with different closest sites at its ends gets a
 * quad iff the 4 cubes around it exist and all their vertices are inside (validity as include/voroinfo_imp.h:26-34);
 * lambda = lambdaForFace of the two sites (pinned arithmetic, lambda_f above).  Order: z, y, x, axis ascending.
 * id / inside: planes [zlo, zhi) of the grid with zlo <= z0 - 1 (or 0) and zhi >= z1 + 1 (or nz); anchors are
 * relative to plane z0.  Returns the number of quads; writes at most cap. */
int64_t orc_medial_quads(const float* sites_xyz, const int32_t* id, const uint8_t* inside, int nx, int ny, int nz, int zlo,
                         int zhi, int z0, int z1, int64_t cap, uint32_t* anchor, uint8_t* axis, int32_t* ida, int32_t* idb,
                         float* lam)
{
#define IN(x, y, z) ((x) >= 0 && (x) < nx && (y) >= 0 && (y) < ny && (z) >= zlo && (z) < zhi && (z) >= 0 && (z) < nz && \
                     inside[(size_t)(x) + (size_t)nx * ((size_t)(y) + (size_t)ny * (size_t)((z) - zlo))])
#define ID(x, y, z) id[(size_t)(x) + (size_t)nx * ((size_t)(y) + (size_t)ny * (size_t)((z) - zlo))]
    int64_t n = 0;
    for (int z = z0; z < z1; ++z)
        for (int y = 0; y < ny; ++y)
            for (int x = 0; x < nx; ++x)
                for (int a = 0; a < 3; ++a)
                {
                    /* the 2 x 3 x 3 block of vertices of the 4 cubes around the edge: 2 along the edge, 3 x 3 across */
                    int lo[3] = {x - 1, y - 1, z - 1}, hi[3] = {x + 1, y + 1, z + 1};
                    lo[a] += 1; /* along the edge: the two end vertices only */
                    int ok = 1;
                    for (int zz = lo[2]; zz <= hi[2] && ok; ++zz)
                        for (int yy = lo[1]; yy <= hi[1] && ok; ++yy)
                            for (int xx = lo[0]; xx <= hi[0] && ok; ++xx)
                                ok = IN(xx, yy, zz);
                    if (!ok)
                        continue;
                    const int x1 = x + (a == 0), y1 = y + (a == 1), zq = z + (a == 2);
                    const int32_t i0 = ID(x, y, z), i1 = ID(x1, y1, zq);
                    if (i0 == i1)
                        continue;
                    if (n < cap)
                    {
                        anchor[n] = (uint32_t)((size_t)x + (size_t)nx * ((size_t)y + (size_t)ny * (size_t)(z - z0)));
                        axis[n] = (uint8_t)a;
                        ida[n] = i0;
                        idb[n] = i1;
                        lam[n] = lambda_f(sites_xyz + 3 * (size_t)i0, sites_xyz + 3 * (size_t)i1);
                    }
                    ++n;
                }
#undef IN
#undef ID
    return n;
}

/* ---- 1': parity classification of the grid from a closed triangle mesh --------------------------
 * PARITY UNPINNED: the reference has no mesh voxeliser (stage 1'), so there is
 * nothing of the reference to pin this against.  It is pinned instead (tests/test_mesh_classify.py)
 * against analytic solids (sphere / torus implicit functions away from the surface) and against the
 * brute-force even-odd count below, which is written triangle-major with plain loops and shares no
 * code with the CUDA kernels.
 *
 * Rule (the product states the same in voxel_ma_b200/csrc/vc_mesh_core.h):
 *   vertices: q = M*p as in orc_classify_points (double 4x4, homogeneous divide, cast to float),
 *             snapped to 1/256 voxel, Q = floor(256 q + 0.5);
 *   voxel centre (i,j,k) is inside iff an odd number of triangles T satisfy
 *       (a) (256 j, 256 k) lies in T's (y,z) projection -- a point on an edge U->V of the
 *           counter-clockwise projection counts iff U > V in (y, then z) order, and
 *       (b) 256 i < x_T(j,k), the abscissa of T's plane over that point (an exact rational). */
static long long orc_orient2(long long uy, long long uz, long long vy, long long vz, long long py, long long pz)
{
    return (vy - uy) * (pz - uz) - (vz - uz) * (py - uy);
}
static int orc_edge_owns(long long w, long long uy, long long uz, long long vy, long long vz)
{
    if (w != 0)
        return w > 0;
    return uy > vy || (uy == vy && uz > vz);
}
/* returns 0 on success, 1 when a vertex is out of the supported range [-1024, 3072) voxels, 2 on a bad index */
int orc_classify_mesh(const float* verts, int64_t nv, const uint32_t* tris, int64_t nt, const double* M, int nx, int ny,
                      int nz, uint8_t* inside)
{
    static const double I[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
    const double* xf = M ? M : I;
    long long* Q = (long long*)malloc(sizeof(long long) * 3 * (size_t)(nv > 0 ? nv : 1));
    int rc = 0;
    for (int64_t i = 0; i < nv; ++i)
    {
        double v0 = verts[3 * i], v1 = verts[3 * i + 1], v2 = verts[3 * i + 2];
        double h = 1 / (xf[3] * v0 + xf[7] * v1 + xf[11] * v2 + xf[15]);
        float q[3];
        q[0] = (float)(h * (xf[0] * v0 + xf[4] * v1 + xf[8] * v2 + xf[12]));
        q[1] = (float)(h * (xf[1] * v0 + xf[5] * v1 + xf[9] * v2 + xf[13]));
        q[2] = (float)(h * (xf[2] * v0 + xf[6] * v1 + xf[10] * v2 + xf[14]));
        for (int d = 0; d < 3; ++d)
        {
            double s = floor((double)q[d] * 256.0 + 0.5);
            if (!(s >= -262144.0 && s <= 786431.0))
                rc = 1;
            else
                Q[3 * i + d] = (long long)s;
        }
    }
    memset(inside, 0, (size_t)nx * ny * nz);
    for (int64_t t = 0; t < nt && rc == 0; ++t)
    {
        if (tris[3 * t] >= nv || tris[3 * t + 1] >= nv || tris[3 * t + 2] >= nv)
        {
            rc = 2;
            break;
        }
        const long long* A = Q + 3 * (size_t)tris[3 * t];
        const long long* B = Q + 3 * (size_t)tris[3 * t + 1];
        const long long* C = Q + 3 * (size_t)tris[3 * t + 2];
        long long area = orc_orient2(A[1], A[2], B[1], B[2], C[1], C[2]);
        if (area == 0)
            continue; /* seen edge-on from +x: never crossed */
        if (area < 0)
        {
            const long long* s = B;
            B = C;
            C = s;
            area = -area;
        }
        for (int k = 0; k < nz; ++k)
            for (int j = 0; j < ny; ++j)
            {
                const long long py = 256LL * j, pz = 256LL * k;
                /* cheap reject outside the box of the projection */
                if ((py < A[1] && py < B[1] && py < C[1]) || (py > A[1] && py > B[1] && py > C[1]) ||
                    (pz < A[2] && pz < B[2] && pz < C[2]) || (pz > A[2] && pz > B[2] && pz > C[2]))
                    continue;
                long long wa = orc_orient2(B[1], B[2], C[1], C[2], py, pz);
                long long wb = orc_orient2(C[1], C[2], A[1], A[2], py, pz);
                long long wc = orc_orient2(A[1], A[2], B[1], B[2], py, pz);
                if (!orc_edge_owns(wa, B[1], B[2], C[1], C[2]) || !orc_edge_owns(wb, C[1], C[2], A[1], A[2]) ||
                    !orc_edge_owns(wc, A[1], A[2], B[1], B[2]))
                    continue;
                /* x_T = (wa*Ax + wb*Bx + wc*Cx) / area;  256 i < x_T  <=>  256 i area < that sum (128-bit safe) */
                __int128 sum = (__int128)wa * A[0] + (__int128)wb * B[0] + (__int128)wc * C[0];
                for (int i = 0; i < nx; ++i)
                {
                    if ((__int128)256 * i * area < sum)
                        inside[IDX(i, j, k)] ^= 1;
                    else
                        break;
                }
            }
    }
    free(Q);
    return rc;
}

/* ---- K6: what CellComplexThinning::prune does before its first queue pop ---------------------------
 * refCntPerVert / refCntPerEdge (src/cellcomplex.cpp:315-332) count incidences; the seeding scan
 * (src/ccthin.cpp:246-270) walks edges then vertices in index order and pushes
 *   (FE_PAIR=1, nbFaceofEdge(e,0), e)  when ref_edge[e]==1 and face_edge_pair_below_threshold
 *                                      (:514-521: m_to_remove_face[f] || m_measure[FACE][f] < f_t), then
 *   (EV_PAIR=0, nbEdgeofVert(v,0), v)  when ref_vert[v]==1 and m_measure[EDGE][e] < l_t (:508-512).
 * Pinned by the queue size the reference CLI prints ("after init, q size", tests/golden/cli_sphere64.txt)
 * and, end to end, by the thinned .ply/.r of the GPU CLI being byte-identical to the reference's. */
void orc_ref_counts(const int32_t* idx, int64_t n, int64_t nbins, int32_t* out)
{
    memset(out, 0, sizeof(int32_t) * (size_t)nbins);
    for (int64_t i = 0; i < n; ++i)
        out[idx[i]]++;
}
int64_t orc_simple_pairs(const int32_t* edge_ref, const int32_t* edge_face0, int64_t ne, const float* face_measure,
                         const uint8_t* face_to_remove, float f_t, const int32_t* vert_ref, const int32_t* vert_edge0,
                         int64_t nv, const float* edge_measure, float l_t, int32_t* pairs_out)
{
    int64_t n = 0;
    for (int64_t e = 0; e < ne; ++e)
        if (edge_ref[e] == 1)
        {
            int32_t f = edge_face0[e];
            if ((face_to_remove && face_to_remove[f]) || face_measure[f] < f_t)
            {
                pairs_out[3 * n] = 1, pairs_out[3 * n + 1] = f, pairs_out[3 * n + 2] = (int32_t)e;
                ++n;
            }
        }
    for (int64_t v = 0; v < nv; ++v)
        if (vert_ref[v] == 1)
        {
            int32_t e = vert_edge0[v];
            if (edge_measure[e] < l_t)
            {
                pairs_out[3 * n] = 0, pairs_out[3 * n + 1] = e, pairs_out[3 * n + 2] = (int32_t)v;
                ++n;
            }
        }
    return n;
}

/* ---- a5 (fixed radius): ANNkd_tree::annkFRSearch(q, sqRad, k, idx, dd, 0.0) ----------------------
 * 3rdparty/ann/src/kd_fix_rad_search.cpp:58-189: a point is in range iff its squared distance (double,
 * x,y,z order) is <= sqRad (:172, inclusive); the call returns the number of points in range and the k
 * closest of them.  The reference calls it with k = 0 to count, then with k = count
 * (src/voxelapps.cpp:346-353).  Brute force here; rows are ordered by (distance, id) -- ANN orders
 * equal distances by traversal, so the pin compares rows as sets of (id, distance).
 * off == NULL: count only. */
void orc_radius_search(const double* sites, int64_t ns, const double* q, int64_t nq, const double* sq_rad,
                       const int64_t* off, int32_t* count, int32_t* idx, double* d2)
{
    for (int64_t i = 0; i < nq; ++i)
    {
        int cnt = 0, have = 0;
        int cap = off ? (int)(off[i + 1] - off[i]) : 0;
        int32_t* ri = off ? idx + off[i] : NULL;
        double* rd = off ? d2 + off[i] : NULL;
        for (int k = 0; k < cap; ++k)
            ri[k] = -1, rd[k] = -1.0;
        for (int64_t s = 0; s < ns; ++s)
        {
            double d = 0;
            for (int a = 0; a < 3; ++a)
            {
                double t = q[3 * i + a] - sites[3 * s + a];
                d = d + t * t;
            }
            if (!(d <= sq_rad[i]))
                continue;
            ++cnt;
            if (!cap)
                continue;
            int pos = have;
            if (have == cap)
            {
                if (!(d < rd[cap - 1])) /* ids ascend, so an equal distance never displaces */
                    continue;
                pos = cap - 1;
            }
            else
                ++have;
            while (pos > 0 && rd[pos - 1] > d)
            {
                rd[pos] = rd[pos - 1];
                ri[pos] = ri[pos - 1];
                --pos;
            }
            rd[pos] = d;
            ri[pos] = (int32_t)s;
        }
        if (count)
            count[i] = cnt;
    }
}
