// vc_points.cu -- stage 2 for arbitrary query points and arbitrary (non-lattice) site sets: the
// drop-in for ANNkd_tree::annkSearch(q, k=1, eps=0) (3rdparty/ann/src/kd_search.cpp:88-216; call
// sites src/voroinfo.cpp:362, src/voxelapps.cpp:217,345, src/exporters.cpp:629-636).
//
// Sites are binned into a uniform cell list by the device radix sort (vc_sites.cu); each query does
// an exact expanding-shell search: ring r of cells (Chebyshev distance r from the query's cell) is
// scanned, and the search stops once the best squared distance is strictly below (r*h)^2, the
// least any site in an unvisited ring can have.  Arithmetic is ANN's: double, squared L2, terms
// accumulated x,y,z (3rdparty/ann/src/ANN.cpp:43-58), no FMA contraction; ties go to the lowest
// site id (3rdparty/ann/src/brute.cpp:56-82 -- the kd-tree returns the same distance but an
// order-dependent id).
//
// Memory behaviour (round 2): a cell's sites are packed next to each other as 32-byte records (x, y, z doubles +
// id), read with two 128-bit loads through the read-only path; the cell offsets come from the boundaries of the
// sorted keys (no search per cell); and vc_closest_points SORTS THE QUERIES by cell first (the same device radix
// sort), so the lanes of a warp search the same or neighbouring cells at the same time -- their loads of a site are
// one broadcast out of L1 instead of 32 scattered sector reads.  Staging a cell neighbourhood in shared memory per
// block was weighed and dropped: the reference's query sets (Voronoi vertices, medial-axis vertices, skeleton points:
// about one query per occupied cell) leave nothing to share beyond what the sort already makes L1 do.
#include <cmath>
#include <vector>

#include "vc_internal.h"

struct CellGrid
{
    double org[3];
    double h, inv_h;
    int dim[3];
};

// a site as the searches read it: 32 bytes, aligned, two 128-bit loads
struct __align__(32) CSite
{
    double x, y, z;
    int id, pad;
};
__device__ __forceinline__ CSite cs_load(const CSite* __restrict__ p)
{
    const double2 a = __ldg(reinterpret_cast<const double2*>(p));
    const double2 b = __ldg(reinterpret_cast<const double2*>(p) + 1);
    CSite s;
    s.x = a.x;
    s.y = a.y;
    s.z = b.x;
    s.id = __double_as_longlong(b.y) & 0xFFFFFFFFll;
    s.pad = 0;
    return s;
}

__device__ __forceinline__ int cg_cell(const CellGrid& g, double v, int a)
{
    int k = (int)floor((v - g.org[a]) * g.inv_h);
    return k < 0 ? 0 : (k >= g.dim[a] ? g.dim[a] - 1 : k);
}

__global__ void k_cell_keys(const double* __restrict__ s, int64_t n, CellGrid g, u64* __restrict__ key, u32* __restrict__ val)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    int cx = cg_cell(g, s[3 * i], 0), cy = cg_cell(g, s[3 * i + 1], 1), cz = cg_cell(g, s[3 * i + 2], 2);
    key[i] = ((u64)cz * g.dim[1] + cy) * g.dim[0] + cx;
    val[i] = (u32)i;
}

// ptr[c] = first sorted position whose cell is >= c, from the boundaries of the sorted keys: position i fills the
// cells (key[i-1], key[i]] (position n fills the tail), so every cell is written exactly once and nothing is searched
__global__ void k_cell_ptr(const u64* __restrict__ key_sorted, int64_t n, int64_t ncells, int* __restrict__ ptr)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n)
        return;
    const int64_t lo = i == 0 ? 0 : (int64_t)key_sorted[i - 1] + 1;
    const int64_t hi = i == n ? ncells : (int64_t)key_sorted[i];
    for (int64_t c = lo; c <= hi; ++c)
        ptr[c] = (int)i;
}

// sites of a cell, in ascending id order, re-packed next to each other as 32-byte records
__global__ void k_cell_gather(const double* __restrict__ s, const u32* __restrict__ ids, int64_t n, CSite* __restrict__ packed)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    u32 id = ids[i];
    CSite r;
    r.x = s[3 * (size_t)id];
    r.y = s[3 * (size_t)id + 1];
    r.z = s[3 * (size_t)id + 2];
    r.id = (int)id;
    r.pad = 0;
    packed[i] = r;
}

__device__ __forceinline__ void scan_cell(const CSite* __restrict__ ps, int b, int e, double q0, double q1, double q2, double& best,
                                          int& bid)
{
    for (int k = b; k < e; ++k)
    {
        const CSite p = cs_load(ps + k);
        double t = __dsub_rn(q0, p.x);
        double d = __dmul_rn(t, t);
        t = __dsub_rn(q1, p.y);
        d = __dadd_rn(d, __dmul_rn(t, t));
        t = __dsub_rn(q2, p.z);
        d = __dadd_rn(d, __dmul_rn(t, t));
        const int id = p.id;
        if (d < best || (d == best && id < bid))
        {
            best = d;
            bid = id;
        }
    }
}

// float32 variant: trimesh::KDtree's distance (3rdparty/trimesh2/libsrc/KDtree.cc:28-33), sqr(x0-y0) + sqr(x1-y1) +
// sqr(x2-y2) in float with x = the tree point; sites and query are floats widened exactly, so the casts are lossless
__device__ __forceinline__ void scan_cell_f32(const CSite* __restrict__ ps, int b, int e, double q0, double q1, double q2,
                                              double& best, int& bid)
{
    const float f0 = (float)q0, f1 = (float)q1, f2 = (float)q2;
    for (int k = b; k < e; ++k)
    {
        const CSite p = cs_load(ps + k);
        float t = __fsub_rn((float)p.x, f0);
        float d = __fmul_rn(t, t);
        t = __fsub_rn((float)p.y, f1);
        d = __fadd_rn(d, __fmul_rn(t, t));
        t = __fsub_rn((float)p.z, f2);
        d = __fadd_rn(d, __fmul_rn(t, t));
        const int id = p.id;
        if ((double)d < best || ((double)d == best && id < bid))
        {
            best = (double)d;
            bid = id;
        }
    }
}

// GRID = false: queries are q[3*i..]; GRID = true: query i is the grid vertex (x,y,z) of the slab.
// F32 = true: float32 distances (see scan_cell_f32), only sites with d2 < max_d2 qualify.
template <bool GRID, bool F32 = false>
__global__ void __launch_bounds__(128)
    k_closest_points(const double* __restrict__ q, int64_t n, CellGrid g, const int* __restrict__ ptr,
                     const CSite* __restrict__ ps, const u32* __restrict__ order, int nx, int ny, int z0, int* __restrict__ id_out,
                     double* __restrict__ d2_out, u32* __restrict__ d2x4_out, double max_d2 = INFINITY)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    if (order)
        i = (int64_t)order[i]; // queries are taken in cell order: neighbouring lanes search the same cells
    double q0, q1, q2;
    if (GRID)
    {
        q0 = (double)(i % nx);
        int64_t r = i / nx;
        q1 = (double)(r % ny);
        q2 = (double)(z0 + r / ny);
    }
    else
    {
        q0 = q[3 * i];
        q1 = q[3 * i + 1];
        q2 = q[3 * i + 2];
    }
    const int c0 = cg_cell(g, q0, 0), c1 = cg_cell(g, q1, 1), c2 = cg_cell(g, q2, 2);
    double best = max_d2;
    int bid = -1;
    const int rmax = max(g.dim[0], max(g.dim[1], g.dim[2]));
    // a float32 distance is within a few 1e-7 (relative) of the exact one: stop only with that much room
    const double slack = F32 ? (1.0 - 1e-5) : (1.0 - 1e-12);
    auto scan = [&](int b, int e)
    {
        if (F32)
            scan_cell_f32(ps, b, e, q0, q1, q2, best, bid);
        else
            scan_cell(ps, b, e, q0, q1, q2, best, bid);
    };
    for (int r = 0; r <= rmax; ++r)
    {
        if (r > 0)
        {
            double lim = (double)(r - 1) * g.h; // every unvisited site is at least (r-1)*h away once ring r-1 is done
            if (best < lim * lim * slack)
                break;
        }
        const int zl = c2 - r, zh = c2 + r, yl = c1 - r, yh = c1 + r, xl = c0 - r, xh = c0 + r;
        for (int z = max(zl, 0); z <= min(zh, g.dim[2] - 1); ++z)
        {
            const bool zface = (z == zl || z == zh);
            for (int y = max(yl, 0); y <= min(yh, g.dim[1] - 1); ++y)
            {
                const bool yface = (y == yl || y == yh);
                const int64_t row = ((int64_t)z * g.dim[1] + y) * g.dim[0];
                if (zface || yface)
                { // the whole x-run belongs to the ring: cells are consecutive in the list
                    int xa = max(xl, 0), xb = min(xh, g.dim[0] - 1);
                    if (xa <= xb)
                        scan(ptr[row + xa], ptr[row + xb + 1]);
                }
                else
                {
                    if (xl >= 0)
                        scan(ptr[row + xl], ptr[row + xl + 1]);
                    if (xh < g.dim[0] && r > 0)
                        scan(ptr[row + xh], ptr[row + xh + 1]);
                }
            }
        }
    }
    id_out[i] = bid;
    if (d2_out)
        d2_out[i] = best;
    if (d2x4_out)
        d2x4_out[i] = bid < 0 ? 0xFFFFFFFFu : (u32)__double2ll_rn(4.0 * best);
}

// cell key of every query (for the sort that puts neighbouring queries into neighbouring lanes)
__global__ void k_query_keys(const double* __restrict__ q, int64_t n, CellGrid g, u64* __restrict__ key, u32* __restrict__ val)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    const int cx = cg_cell(g, q[3 * i], 0), cy = cg_cell(g, q[3 * i + 1], 1), cz = cg_cell(g, q[3 * i + 2], 2);
    key[i] = ((u64)cz * g.dim[1] + cy) * g.dim[0] + cx;
    val[i] = (u32)i;
}

static CellGrid g_of(const vc_ctx* c)
{
    CellGrid g;
    for (int a = 0; a < 3; ++a)
    {
        g.org[a] = c->cl_org[a];
        g.dim[a] = c->cl_dim[a];
    }
    g.h = c->cl_h;
    g.inv_h = 1.0 / c->cl_h;
    return g;
}

// device-side cell list from the id-ordered float4 site table (lattice or external sites alike)
static int build_from_device_sites(vc_ctx* c, const double* dsites, int64_t n, const double lo[3], const double hi[3])
{
    double ext[3], vol = 1.0;
    for (int a = 0; a < 3; ++a)
    {
        ext[a] = hi[a] - lo[a];
        if (ext[a] < 1e-9)
            ext[a] = 1e-9;
        vol *= ext[a];
    }
    // sites sample a surface: aim at ~2 sites per occupied cell with h ~ sqrt(area/n) ~ cbrt(vol)/sqrt(n)^(2/3)
    double h = cbrt(vol / (double)(n > 0 ? n : 1)) * 1.5;
    double maxext = fmax(ext[0], fmax(ext[1], ext[2]));
    if (h < maxext / 512.0)
        h = maxext / 512.0;
    if (!(h > 0))
        h = 1.0;
    int64_t ncells = 1;
    for (int a = 0; a < 3; ++a)
    {
        c->cl_org[a] = lo[a];
        c->cl_dim[a] = (int)floor(ext[a] / h) + 1;
        ncells *= c->cl_dim[a];
    }
    c->cl_h = h;
    CellGrid g = g_of(c);
    VC_CUDA(c, c->sk0.ensure((size_t)(n + 1) * 8));
    VC_CUDA(c, c->sk1.ensure((size_t)(n + 1) * 8));
    VC_CUDA(c, c->sv0.ensure((size_t)(n + 1) * 4));
    VC_CUDA(c, c->sv1.ensure((size_t)(n + 1) * 4));
    VC_CUDA(c, c->cl_ptr.ensure((size_t)(ncells + 2) * 4));
    DevBuf packed;
    VC_CUDA(c, packed.ensure((size_t)(n + 1) * sizeof(CSite)));
    u64* k = c->sk0.as<u64>();
    u32* v = c->sv0.as<u32>();
    unsigned blocks = vc_blocks((size_t)(n > 0 ? n : 1), 256);
    if (n)
        VC_LAUNCH(c, "cell_keys", k_cell_keys, blocks, 256, 0, dsites, n, g, k, v);
    int bits = 1;
    while (bits < 62 && ((u64)ncells >> bits))
        ++bits;
    int s = vc_radix_sort_pairs(c, n, bits, &k, &v);
    if (s == VC_OK)
    {
        VC_LAUNCH(c, "cell_ptr", k_cell_ptr, vc_blocks((size_t)n + 1, 256), 256, 0, k, n, ncells, c->cl_ptr.as<int>());
        if (n)
            VC_LAUNCH(c, "cell_gather", k_cell_gather, blocks, 256, 0, dsites, v, n, packed.as<CSite>());
        cudaError_t e = cudaStreamSynchronize(c->stream);
        if (e != cudaSuccess)
            s = vc_fail(c, VC_ERR_CUDA, "cell list", e);
    }
    if (s == VC_OK)
    { // keep the cell-ordered coordinates (replace the id-ordered copy)
        c->gsites.release();
        c->gsites = packed;
    }
    else
        packed.release();
    return s;
}

int st_build_cell_list(vc_ctx* c, const float* xyz_host, int64_t n)
{
    // widen to double exactly as the reference does for ANN (src/voroinfo.cpp:336-341)
    std::vector<double> s((size_t)n * 3);
    double lo[3] = {0, 0, 0}, hi[3] = {0, 0, 0};
    for (int64_t i = 0; i < n; ++i)
        for (int a = 0; a < 3; ++a)
        {
            double v = (double)xyz_host[3 * i + a];
            s[3 * i + a] = v;
            if (i == 0 || v < lo[a])
                lo[a] = v;
            if (i == 0 || v > hi[a])
                hi[a] = v;
        }
    DevBuf ds;
    VC_CUDA(c, ds.ensure((size_t)(n + 1) * 24));
    if (n)
        VC_CUDA(c, cudaMemcpyAsync(ds.p, s.data(), (size_t)n * 24, cudaMemcpyHostToDevice, c->stream));
    // id-ordered float4 table for the measure kernels
    VC_CUDA(c, c->site_xyz.ensure((size_t)(n + 1) * 16));
    {
        std::vector<float> f4((size_t)n * 4, 0.0f);
        for (int64_t i = 0; i < n; ++i)
            for (int a = 0; a < 3; ++a)
                f4[4 * i + a] = xyz_host[3 * i + a];
        if (n)
            VC_CUDA(c, cudaMemcpyAsync(c->site_xyz.p, f4.data(), (size_t)n * 16, cudaMemcpyHostToDevice, c->stream));
        VC_CUDA(c, cudaStreamSynchronize(c->stream));
    }
    int st = build_from_device_sites(c, ds.as<double>(), n, lo, hi);
    ds.release();
    if (st != VC_OK)
        return st;
    c->nsites = n;
    c->lattice = false;
    c->have_sites = true;
    c->have_closest = c->have_measures = false;
    return VC_OK;
}

__global__ void k_sites_to_double(const float4* __restrict__ s, int64_t n, double* __restrict__ out)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    float4 v = s[i];
    out[3 * i] = (double)v.x;
    out[3 * i + 1] = (double)v.y;
    out[3 * i + 2] = (double)v.z;
}

// lattice sites: the cell list is built lazily the first time an arbitrary-point query arrives
static int ensure_cell_list(vc_ctx* c)
{
    if (c->cl_dim[0] > 0 && c->gsites.p)
        return VC_OK;
    int64_t n = c->nsites;
    DevBuf ds;
    VC_CUDA(c, ds.ensure((size_t)(n + 1) * 24));
    if (n)
        VC_LAUNCH(c, "sites_to_double", k_sites_to_double, vc_blocks((size_t)n, 256), 256, 0, c->site_xyz.as<float4>(), n,
                  ds.as<double>());
    double lo[3] = {-0.5, -0.5, -0.5}, hi[3] = {c->nx - 0.5, c->ny - 0.5, c->nz - 0.5};
    int st = build_from_device_sites(c, ds.as<double>(), n, lo, hi);
    ds.release();
    return st;
}

int st_closest_points(vc_ctx* c, const double* q, int64_t n, int32_t* id, double* d2)
{
    if (!c->have_sites)
        return vc_fail(c, VC_ERR_STATE, "vc_closest_points needs sites");
    if (n == 0)
        return VC_OK;
    VC_TRY(ensure_cell_list(c));
    DevBuf dq, did, dd;
    cudaError_t e = dq.ensure((size_t)n * 24);
    if (e == cudaSuccess)
        e = did.ensure((size_t)n * 4);
    if (e == cudaSuccess)
        e = dd.ensure((size_t)n * 8);
    if (e == cudaSuccess)
        e = cudaMemcpyAsync(dq.p, q, (size_t)n * 24, cudaMemcpyDefault, c->stream);
    const u32* order = nullptr;
    if (e == cudaSuccess && n >= 4096 && n < ((int64_t)1 << 31))
    { // the queries in cell order (the cell list's own buffers are free again: it keeps only cl_ptr and the packed sites)
        e = c->sk0.ensure((size_t)(n + 1) * 8);
        e = e == cudaSuccess ? c->sk1.ensure((size_t)(n + 1) * 8) : e;
        e = e == cudaSuccess ? c->sv0.ensure((size_t)(n + 1) * 4) : e;
        e = e == cudaSuccess ? c->sv1.ensure((size_t)(n + 1) * 4) : e;
        if (e == cudaSuccess)
        {
            u64* k = c->sk0.as<u64>();
            u32* v = c->sv0.as<u32>();
            const CellGrid g = g_of(c);
            VC_LAUNCH(c, "query_keys", k_query_keys, vc_blocks((size_t)n, 256), 256, 0, dq.as<double>(), n, g, k, v);
            int bits = 1;
            while (bits < 62 && (((u64)g.dim[0] * g.dim[1] * g.dim[2]) >> bits))
                ++bits;
            if (vc_radix_sort_pairs(c, n, bits, &k, &v) == VC_OK)
                order = v;
        }
    }
    if (e == cudaSuccess)
    {
        VC_LAUNCH(c, "closest_points", (k_closest_points<false>), vc_blocks((size_t)n, 128), 128, 0, dq.as<double>(), n,
                  g_of(c), c->cl_ptr.as<int>(), c->gsites.as<CSite>(), order, 0, 0, 0, did.as<int>(),
                  dd.as<double>(), (u32*)nullptr);
        e = cudaMemcpyAsync(id, did.p, (size_t)n * 4, cudaMemcpyDefault, c->stream);
        if (e == cudaSuccess && d2)
            e = cudaMemcpyAsync(d2, dd.p, (size_t)n * 8, cudaMemcpyDefault, c->stream);
    }
    if (e == cudaSuccess)
        e = cudaStreamSynchronize(c->stream);
    dq.release();
    did.release();
    dd.release();
    if (e != cudaSuccess)
        return vc_fail(c, VC_ERR_CUDA, "closest_points", e);
    return VC_OK;
}

// float32 form: drop-in for trimesh::KDtree::closest_to_pt(p, maxdist2) (3rdparty/trimesh2/libsrc/KDtree.cc:252-292,
// 523-545; call site src/exporters.cpp:629-636).  id = -1 / d2 = -1 when no site has d2 < max_d2.
int st_closest_points_f32(vc_ctx* c, const float* q, int64_t n, float max_d2, int32_t* id, float* d2)
{
    if (!c->have_sites)
        return vc_fail(c, VC_ERR_STATE, "vc_closest_points_f32 needs sites");
    if (n == 0)
        return VC_OK;
    VC_TRY(ensure_cell_list(c));
    std::vector<double> qd((size_t)n * 3);
    for (size_t k = 0; k < qd.size(); ++k)
        qd[k] = (double)q[k];
    std::vector<double> dd((size_t)n);
    DevBuf dq, did, ddv;
    cudaError_t e = dq.ensure((size_t)n * 24);
    if (e == cudaSuccess)
        e = did.ensure((size_t)n * 4);
    if (e == cudaSuccess)
        e = ddv.ensure((size_t)n * 8);
    if (e == cudaSuccess)
        e = cudaMemcpyAsync(dq.p, qd.data(), (size_t)n * 24, cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess)
    {
        const double lim = (max_d2 > 0.0f && max_d2 < INFINITY) ? (double)max_d2 : (double)INFINITY;
        VC_LAUNCH(c, "closest_points_f32", (k_closest_points<false, true>), vc_blocks((size_t)n, 128), 128, 0, dq.as<double>(), n,
                  g_of(c), c->cl_ptr.as<int>(), c->gsites.as<CSite>(), (const u32*)nullptr, 0, 0, 0, did.as<int>(),
                  ddv.as<double>(), (u32*)nullptr, lim);
        e = cudaMemcpyAsync(id, did.p, (size_t)n * 4, cudaMemcpyDefault, c->stream);
        if (e == cudaSuccess)
            e = cudaMemcpyAsync(dd.data(), ddv.p, (size_t)n * 8, cudaMemcpyDeviceToHost, c->stream);
    }
    if (e == cudaSuccess)
        e = cudaStreamSynchronize(c->stream);
    dq.release();
    did.release();
    ddv.release();
    if (e != cudaSuccess)
        return vc_fail(c, VC_ERR_CUDA, "closest_points_f32", e);
    if (d2)
        for (int64_t i = 0; i < n; ++i)
            d2[i] = id[i] < 0 ? -1.0f : (float)dd[(size_t)i];
    return VC_OK;
}

// =============================================================================================
// Fixed-radius query: drop-in for ANNkd_tree::annkFRSearch(q, sqRad, k, idx, dd, eps=0)
// (3rdparty/ann/src/kd_fix_rad_search.cpp:58-189; call sites src/voxelapps.cpp:346,353).
// A site is in range iff its squared distance (double, terms accumulated x,y,z) is <= sqRad
// (inclusive, :172; ANN's early exit on a partial sum is the same predicate because the partial sums
// never exceed the total).  count = number of sites in range; the row off[i]..off[i+1] receives the
// (off[i+1]-off[i]) closest of them ordered by (squared distance, site id), -1 / -1.0 padded.  ANN
// orders equal distances by its traversal; the reference only ever asks for ALL sites in range and
// uses them as a set, so the deterministic order here is a refinement, not a difference.
// =============================================================================================
__global__ void __launch_bounds__(128)
    k_radius_search(const double* __restrict__ q, const double* __restrict__ sq_rad, int64_t n, CellGrid g, const int* __restrict__ ptr,
                    const CSite* __restrict__ ps, const int64_t* __restrict__ off, int* __restrict__ count,
                    int* __restrict__ idx, double* __restrict__ dd)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    const double q0 = q[3 * i], q1 = q[3 * i + 1], q2 = q[3 * i + 2], R2 = sq_rad[i];
    int cnt = 0;
    int64_t row0 = 0;
    int cap = 0, have = 0;
    if (off)
    {
        row0 = off[i];
        cap = (int)(off[i + 1] - off[i]);
        for (int k = 0; k < cap; ++k)
        {
            idx[row0 + k] = -1;
            dd[row0 + k] = -1.0;
        }
    }
    if (R2 >= 0.0)
    { // cells the ball can reach (radius padded by one part in 1e9 against the rounding of sqrt and of the cell map)
        const double r = sqrt(R2) * (1.0 + 1e-9) + 1e-300;
        const int xl = cg_cell(g, q0 - r, 0), xh = cg_cell(g, q0 + r, 0);
        const int yl = cg_cell(g, q1 - r, 1), yh = cg_cell(g, q1 + r, 1);
        const int zl = cg_cell(g, q2 - r, 2), zh = cg_cell(g, q2 + r, 2);
        for (int z = zl; z <= zh; ++z)
            for (int y = yl; y <= yh; ++y)
            {
                const int64_t rowc = ((int64_t)z * g.dim[1] + y) * g.dim[0];
                const int b = ptr[rowc + xl], e = ptr[rowc + xh + 1]; // the x-run is contiguous in the list
                for (int k = b; k < e; ++k)
                {
                    const CSite p = cs_load(ps + k);
                    double t = __dsub_rn(q0, p.x);
                    double d = __dmul_rn(t, t);
                    t = __dsub_rn(q1, p.y);
                    d = __dadd_rn(d, __dmul_rn(t, t));
                    t = __dsub_rn(q2, p.z);
                    d = __dadd_rn(d, __dmul_rn(t, t));
                    if (!(d <= R2))
                        continue;
                    ++cnt;
                    if (!cap)
                        continue;
                    const int id = p.id;
                    // insertion into the sorted row (distance, id); the last one falls off when full
                    int pos = have;
                    if (have == cap)
                    {
                        const double dl = dd[row0 + cap - 1];
                        const int il = idx[row0 + cap - 1];
                        if (!(d < dl || (d == dl && id < il)))
                            continue;
                        pos = cap - 1;
                    }
                    else
                        ++have;
                    while (pos > 0)
                    {
                        const double dp = dd[row0 + pos - 1];
                        const int ip = idx[row0 + pos - 1];
                        if (dp < d || (dp == d && ip < id))
                            break;
                        dd[row0 + pos] = dp;
                        idx[row0 + pos] = ip;
                        --pos;
                    }
                    dd[row0 + pos] = d;
                    idx[row0 + pos] = id;
                }
            }
    }
    if (count)
        count[i] = cnt;
}

int st_radius_search(vc_ctx* c, const double* q, const double* sq_rad, int64_t n, const int64_t* off, int32_t* count, int32_t* idx,
                     double* d2)
{
    if (!c->have_sites)
        return vc_fail(c, VC_ERR_STATE, "vc_radius_search needs sites");
    if (n == 0)
        return VC_OK;
    VC_TRY(ensure_cell_list(c));
    const int64_t total = off ? off[n] : 0;
    if (off && (off[0] != 0 || total < 0))
        return vc_fail(c, VC_ERR_INVALID, "vc_radius_search: off must start at 0 and be non-decreasing");
    DevBuf dq, dr, doff, dcnt, didx, ddd;
    cudaError_t e = dq.ensure((size_t)n * 24);
    if (e == cudaSuccess)
        e = dr.ensure((size_t)n * 8);
    if (e == cudaSuccess)
        e = dcnt.ensure((size_t)n * 4);
    if (e == cudaSuccess && off)
    {
        e = doff.ensure((size_t)(n + 1) * 8);
        if (e == cudaSuccess)
            e = didx.ensure((size_t)(total > 0 ? total : 1) * 4);
        if (e == cudaSuccess)
            e = ddd.ensure((size_t)(total > 0 ? total : 1) * 8);
        if (e == cudaSuccess)
            e = cudaMemcpyAsync(doff.p, off, (size_t)(n + 1) * 8, cudaMemcpyHostToDevice, c->stream);
    }
    if (e == cudaSuccess)
        e = cudaMemcpyAsync(dq.p, q, (size_t)n * 24, cudaMemcpyDefault, c->stream);
    if (e == cudaSuccess)
        e = cudaMemcpyAsync(dr.p, sq_rad, (size_t)n * 8, cudaMemcpyDefault, c->stream);
    if (e == cudaSuccess)
    {
        VC_LAUNCH(c, "radius_search", k_radius_search, vc_blocks((size_t)n, 128), 128, 0, dq.as<double>(), dr.as<double>(), n, g_of(c),
                  c->cl_ptr.as<int>(), c->gsites.as<CSite>(), off ? doff.as<int64_t>() : (const int64_t*)nullptr,
                  dcnt.as<int>(), didx.as<int>(), ddd.as<double>());
        if (count)
            e = cudaMemcpyAsync(count, dcnt.p, (size_t)n * 4, cudaMemcpyDefault, c->stream);
        if (e == cudaSuccess && off && total > 0 && idx)
            e = cudaMemcpyAsync(idx, didx.p, (size_t)total * 4, cudaMemcpyDefault, c->stream);
        if (e == cudaSuccess && off && total > 0 && d2)
            e = cudaMemcpyAsync(d2, ddd.p, (size_t)total * 8, cudaMemcpyDefault, c->stream);
    }
    if (e == cudaSuccess)
        e = cudaStreamSynchronize(c->stream);
    dq.release(), dr.release(), doff.release(), dcnt.release(), didx.release(), ddd.release();
    if (e != cudaSuccess)
        return vc_fail(c, VC_ERR_CUDA, "radius_search", e);
    return VC_OK;
}

// arbitrary (non-lattice) site set queried at every grid vertex of the slab
int st_closest_general_grid(vc_ctx* c)
{
    VC_TRY(ensure_cell_list(c));
    const int nplanes = c->zc - c->z0;
    const size_t nv = (size_t)c->nx * c->ny * nplanes;
    VC_CUDA(c, c->id.ensure(nv * 4));
    VC_CUDA(c, c->d2.ensure(nv * 4));
    VC_LAUNCH(c, "closest_grid_celllist", (k_closest_points<true>), vc_blocks(nv, 128), 128, 0, (const double*)nullptr,
              (int64_t)nv, g_of(c), c->cl_ptr.as<int>(), c->gsites.as<CSite>(), (const u32*)nullptr, c->nx, c->ny, c->z0,
              c->id.as<int>(), (double*)nullptr, c->d2.as<u32>());
    VC_CUDA(c, cudaGetLastError());
    c->have_closest = true;
    c->have_measures = false;
    return VC_OK;
}
