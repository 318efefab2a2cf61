// vc_measures.cu -- stage 3: medial measures, plus the small per-element operators of the
// Voronoi-complex side of the boundary (tagVert, face lambda, vertex radii, max aggregation).
//
// Arithmetic follows the reference in float32, operation by operation, with explicit
// round-to-nearest intrinsics so nvcc cannot contract a multiply-add into an FMA:
//   MeasureForMA::lambdaForFace = trimesh::dist (include/measureforMA_imp.h:1-4,
//   3rdparty/trimesh2/include/Vec.h:1128-1143):  d2 = sqr(b0-a0); d2 += sqr(b1-a1); d2 += sqr(b2-a2); sqrt.
#include "vc_internal.h"

__device__ __forceinline__ float vc_dist2f(float ax, float ay, float az, float bx, float by, float bz)
{
    float t = __fsub_rn(bx, ax);
    float d2 = __fmul_rn(t, t);
    t = __fsub_rn(by, ay);
    d2 = __fadd_rn(d2, __fmul_rn(t, t));
    t = __fsub_rn(bz, az);
    d2 = __fadd_rn(d2, __fmul_rn(t, t));
    return d2;
}

// =============================================================================================
// K5  dense cell measures (dictionary of SURVEY section 0).  The 7 cells anchored at grid vertex v
// need the closest sites of the 8 vertices of v's cube.  sqrt is monotone and correctly rounded,
// so max over edges of sqrt(d2) == sqrt(max d2): the 12 edge terms are reduced as squared
// distances and only the 7 outputs (+ radius) take a square root.
//
// Work split.  A thread owns 4 consecutive x vertices of one row, so every access to the
// per-vertex planes is one 128-bit load / store (8 stores per 4 vertices).  A block owns a
// 128 x 8 (x,y) tile and marches up a chunk of z planes; per plane it parks one 16-byte record
// (site position, flags) per tile vertex (+ a one-vertex halo in x and y) in shared memory,
// in a ring of three plane buffers (a plane is read by the
// iterations z-1 and z, so one barrier per plane suffices with three).
// What is skipped.  A cell is valid only if ALL its vertices are inside, and every one of the 7
// cells anchored at v contains v: so an outside vertex reports 0 everywhere, its site is never
// used by any cell, and neither its id nor its site record is fetched at all.  The occupancy bit
// rows (1 bit / vertex) decide; ids are read (and sites gathered from the L2-resident table, once
// per vertex) only for 4-vertex groups that contain an inside vertex.
// Radius.  r(v) = dist(site, v) in float32.  For lattice sites every term of trimesh::dist2 is an
// exact multiple of 1/4, so while 4*d^2 < 2^24 the float sum equals (float)(4d^2)/4 exactly and the
// radius comes from the d2x4 plane (one more 128-bit load) instead of a gather; beyond that, and
// for arbitrary site sets, the site is fetched and the sum is accumulated in the reference order.
// id, d2: planes [z0, zc) (zc = z1+1 halo when z1 < nz); bits: planes [zlo, zhi).
// =============================================================================================
#define CM_VX 4                 // vertices per thread along x
#define CM_TW (32 * CM_VX)      // tile width  (x)
#define CM_TH 8                 // tile height (y)
#define CM_PITCH (CM_TW + 4)    // record row pitch (halo column + padding)

struct CmRec
{
    float x, y, z;
    u32 f; // bit 0: vertex has a site; bit 1: vertex is inside (records of outside vertices hold f = 0)
};

__device__ __forceinline__ float cm_e2(const CmRec& a, const CmRec& b)
{
    return (a.f & b.f & 1u) ? vc_dist2f(a.x, a.y, a.z, b.x, b.y, b.z) : 0.0f;
}

template <bool VEC>
__global__ void __launch_bounds__(32 * CM_TH, 3)
    k_cell_measures(const int* __restrict__ id, const u32* __restrict__ d2x4, const u32* __restrict__ bits, int wr,
                    const float4* __restrict__ site, int nx, int ny, int z0, int z1, int zc, int zlo, int za, int zb, int zchunk,
                    int radius_from_d2, float* __restrict__ edge3, float* __restrict__ face3, float* __restrict__ cube,
                    float* __restrict__ radius)
{
    extern __shared__ float4 cm_tile[]; // [3][CM_TH + 1][CM_PITCH]
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int x0 = blockIdx.x * CM_TW, y0 = blockIdx.y * CM_TH;
    const int zs = za + blockIdx.z * zchunk, ze = min(zs + zchunk, zb); // anchors [za, zb) of the slab [z0, z1)
    const int x = x0 + CM_VX * tx, y = y0 + ty;
    const size_t plane = (size_t)nx * ny;
    const size_t nv = plane * (size_t)(z1 - z0);
    auto tile = [&](int buf, int row, int col) -> float4& { return cm_tile[(buf * (CM_TH + 1) + row) * CM_PITCH + col]; };

    // inside nibble of the 4 vertices (xx .. xx+3, yy, zz); 0 outside the grid / beyond the halo plane
    auto nibble = [&](int xx, int yy, int zz) -> u32
    {
        if (xx >= nx || yy >= ny || zz >= zc)
            return 0u;
        u32 w = __ldg(bits + ((size_t)(zz - zlo) * ny + yy) * (size_t)wr + (xx >> 5));
        return (w >> (xx & 31)) & 0xFu; // xx is a multiple of 4: the nibble never straddles a word
    };
    // park the records of the 4 vertices (xx.., yy, zz) whose inside nibble is `nib`
    auto park4 = [&](int buf, int row, int col, int xx, int yy, int zz, u32 nib)
    {
        float4 r[CM_VX];
#pragma unroll
        for (int k = 0; k < CM_VX; ++k)
            r[k] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        if (nib)
        {
            const size_t o = (size_t)xx + (size_t)nx * (size_t)yy + plane * (size_t)(zz - z0);
            int s[CM_VX];
            if (VEC)
            {
                int4 v = __ldcs(reinterpret_cast<const int4*>(id + o));
                s[0] = v.x, s[1] = v.y, s[2] = v.z, s[3] = v.w;
            }
            else
            {
#pragma unroll
                for (int k = 0; k < CM_VX; ++k)
                    s[k] = (xx + k < nx) ? __ldcs(id + o + k) : -1;
            }
#pragma unroll
            for (int k = 0; k < CM_VX; ++k)
                if ((nib >> k) & 1u)
                {
                    u32 f = 2u;
                    if (s[k] >= 0)
                    {
                        r[k] = __ldg(site + s[k]);
                        f = 3u;
                    }
                    r[k].w = __uint_as_float(f);
                }
        }
#pragma unroll
        for (int k = 0; k < CM_VX; ++k)
            tile(buf, row, col + k) = r[k];
    };
    // one plane of records: own row, the halo row (warp 0), the halo column (lane 31 of each row)
    auto park_plane = [&](int buf, int zz, u32 nib_own)
    {
        park4(buf, ty, CM_VX * tx, x, y, zz, nib_own);
        if (ty == 0)
            park4(buf, CM_TH, CM_VX * tx, x, y0 + CM_TH, zz, nibble(x, y0 + CM_TH, zz));
        if (tx == 31)
        {
            const int nrow = ty == 0 ? 2 : 1; // row ty, and the corner for ty == 0
            for (int j = 0; j < nrow; ++j)
            {
                const int row = j == 0 ? ty : CM_TH;
                const int xx = x0 + CM_TW, yy = y0 + row;
                float4 r = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                if (xx < nx && yy < ny && zz < zc)
                {
                    u32 w = __ldg(bits + ((size_t)(zz - zlo) * ny + yy) * (size_t)wr + (xx >> 5));
                    if ((w >> (xx & 31)) & 1u)
                    {
                        int sid = __ldg(id + (size_t)xx + (size_t)nx * (size_t)yy + plane * (size_t)(zz - z0));
                        u32 f = 2u;
                        if (sid >= 0)
                        {
                            r = __ldg(site + sid);
                            f = 3u;
                        }
                        r.w = __uint_as_float(f);
                    }
                }
                tile(buf, row, CM_TW) = r;
            }
        }
    };
    auto rec = [&](int buf, int row, int col) -> CmRec
    {
        float4 v = tile(buf, row, col);
        CmRec r;
        r.x = v.x, r.y = v.y, r.z = v.z;
        r.f = __float_as_uint(v.w);
        return r;
    };

    const bool row_live = y < ny && x < nx;
    size_t o = (size_t)x + (size_t)nx * (size_t)y + plane * (size_t)(zs - z0);
    // Streaming path.  When no vertex of the block's tile (halo row and column included) is inside on any of its
    // planes, every lambda it owns is 0 and nothing has to be parked or synchronised: the block only streams the
    // radius (from the d2x4 plane) and the zeros.  Thin solids in a large grid take this path almost everywhere.
    {
        // the tile's occupancy words: planes zs..ze (ze = halo), rows y0..y0+CM_TH (halo row), words of x0..x0+CM_TW
        // (halo column = bit 0 of the fifth word); spread over the block's threads so the loads are independent
        u32 any = 0;
        const int w0 = x0 >> 5, nwords = CM_TW / 32 + 1, per_plane = (CM_TH + 1) * nwords;
        const int total = (ze - zs + 1) * per_plane;
        for (int i = threadIdx.x; i < total; i += 32 * CM_TH)
        {
            const int pz = i / per_plane, r = i - pz * per_plane;
            const int row = r / nwords, w = r - row * nwords;
            const int zz = zs + pz, yy = y0 + row, ww = w0 + w;
            if (zz < zc && yy < ny && ww < wr)
            {
                const u32 word = __ldg(bits + ((size_t)(zz - zlo) * ny + yy) * (size_t)wr + ww);
                any |= (w == nwords - 1) ? (word & 1u) : word;
            }
        }
        if (!__syncthreads_or((int)any))
        {
            if (!row_live)
                return;
            const float4 zero4 = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
#pragma unroll 4
            for (int z = zs; z < ze; ++z, o += plane)
            {
                float lr[CM_VX] = {0.0f, 0.0f, 0.0f, 0.0f};
                if (radius)
                {
                    u32 q[CM_VX];
                    if (VEC)
                    {
                        uint4 v = __ldcs(reinterpret_cast<const uint4*>(d2x4 + o));
                        q[0] = v.x, q[1] = v.y, q[2] = v.z, q[3] = v.w;
                    }
                    else
                    {
#pragma unroll
                        for (int k = 0; k < CM_VX; ++k)
                            q[k] = (x + k < nx) ? __ldcs(d2x4 + o + k) : 0u;
                    }
#pragma unroll
                    for (int k = 0; k < CM_VX; ++k)
                    {
                        if (radius_from_d2 && q[k] < (1u << 24))
                            lr[k] = __fsqrt_rn(__fmul_rn((float)q[k], 0.25f));
                        else if (x + k < nx)
                        {
                            int sid = __ldg(id + o + k);
                            if (sid >= 0)
                            {
                                float4 sp = __ldg(site + sid);
                                lr[k] = __fsqrt_rn(vc_dist2f(sp.x, sp.y, sp.z, (float)(x + k), (float)y, (float)z));
                            }
                        }
                    }
                }
                auto put0 = [&](float* base, const float* v)
                {
                    if (VEC)
                        __stcs(reinterpret_cast<float4*>(base + o), v ? make_float4(v[0], v[1], v[2], v[3]) : zero4);
                    else
                    {
#pragma unroll
                        for (int k = 0; k < CM_VX; ++k)
                            if (x + k < nx)
                                __stcs(base + o + k, v ? v[k] : 0.0f);
                    }
                };
                if (edge3)
                {
                    put0(edge3, nullptr);
                    put0(edge3 + nv, nullptr);
                    put0(edge3 + 2 * nv, nullptr);
                }
                if (face3)
                {
                    put0(face3, nullptr);
                    put0(face3 + nv, nullptr);
                    put0(face3 + 2 * nv, nullptr);
                }
                if (cube)
                    put0(cube, nullptr);
                if (radius)
                    put0(radius, lr);
            }
            return;
        }
    }
    u32 nib_cur = nibble(x, y, zs);
    park_plane(0, zs, nib_cur);
    for (int z = zs; z < ze; ++z, o += plane)
    {
        const int bufa = (z - zs) % 3, bufb = (z - zs + 1) % 3;
        const u32 nib_next = nibble(x, y, z + 1);
        park_plane(bufb, z + 1, nib_next);
        __syncthreads();
        float le[3][CM_VX], lf[3][CM_VX], lc[CM_VX], lr[CM_VX];
#pragma unroll
        for (int k = 0; k < CM_VX; ++k)
            le[0][k] = le[1][k] = le[2][k] = lf[0][k] = lf[1][k] = lf[2][k] = lc[k] = lr[k] = 0.0f;
        if (nib_cur)
        {
#pragma unroll
            for (int k = 0; k < CM_VX; ++k)
                if ((nib_cur >> k) & 1u)
                {
                    const int c0 = CM_VX * tx + k;
                    const CmRec a0 = rec(bufa, ty, c0), a1 = rec(bufa, ty, c0 + 1), a2 = rec(bufa, ty + 1, c0),
                                a3 = rec(bufa, ty + 1, c0 + 1);
                    const CmRec b0 = rec(bufb, ty, c0), b1 = rec(bufb, ty, c0 + 1), b2 = rec(bufb, ty + 1, c0),
                                b3 = rec(bufb, ty + 1, c0 + 1);
                    // x edges {0,1},{2,3},{4,5},{6,7}; y edges {0,2},{1,3},{4,6},{5,7}; z edges {0,4},{1,5},{2,6},{3,7}
                    const float ex0 = cm_e2(a0, a1), ex1 = cm_e2(a2, a3), ex2 = cm_e2(b0, b1), ex3 = cm_e2(b2, b3);
                    const float ey0 = cm_e2(a0, a2), ey1 = cm_e2(a1, a3), ey2 = cm_e2(b0, b2), ey3 = cm_e2(b1, b3);
                    const float w0 = cm_e2(a0, b0), w1 = cm_e2(a1, b1), w2 = cm_e2(a2, b2), w3 = cm_e2(a3, b3);
                    const u32 i1 = a1.f & 2u, i2 = a2.f & 2u, i3 = a3.f & 2u;
                    const u32 i4 = b0.f & 2u, i5 = b1.f & 2u, i6 = b2.f & 2u, i7 = b3.f & 2u;
                    if (i1)
                        le[0][k] = __fsqrt_rn(ex0);
                    if (i2)
                        le[1][k] = __fsqrt_rn(ey0);
                    if (i4)
                        le[2][k] = __fsqrt_rn(w0);
                    if (i1 & i2 & i3)
                        lf[0][k] = __fsqrt_rn(fmaxf(fmaxf(ex0, ex1), fmaxf(ey0, ey1)));
                    if (i1 & i4 & i5)
                        lf[1][k] = __fsqrt_rn(fmaxf(fmaxf(ex0, ex2), fmaxf(w0, w1)));
                    if (i2 & i4 & i6)
                        lf[2][k] = __fsqrt_rn(fmaxf(fmaxf(ey0, ey2), fmaxf(w0, w2)));
                    if (i1 & i2 & i3 & i4 & i5 & i6 & i7)
                    {
                        float m = fmaxf(fmaxf(fmaxf(ex0, ex1), fmaxf(ex2, ex3)), fmaxf(fmaxf(ey0, ey1), fmaxf(ey2, ey3)));
                        lc[k] = __fsqrt_rn(fmaxf(m, fmaxf(fmaxf(w0, w1), fmaxf(w2, w3))));
                    }
                }
        }
        if (row_live)
        {
            if (radius)
            {
                u32 q[CM_VX];
                if (VEC)
                {
                    uint4 v = __ldcs(reinterpret_cast<const uint4*>(d2x4 + o));
                    q[0] = v.x, q[1] = v.y, q[2] = v.z, q[3] = v.w;
                }
                else
                {
#pragma unroll
                    for (int k = 0; k < CM_VX; ++k)
                        q[k] = (x + k < nx) ? __ldcs(d2x4 + o + k) : 0u;
                }
#pragma unroll
                for (int k = 0; k < CM_VX; ++k)
                {
                    if (radius_from_d2 && q[k] < (1u << 24))
                        lr[k] = __fsqrt_rn(__fmul_rn((float)q[k], 0.25f));
                    else if (x + k < nx)
                    { // rare: far beyond 2048 voxels, or an arbitrary site set -- the reference's own sum
                        int sid = __ldg(id + o + k);
                        if (sid >= 0)
                        {
                            float4 sp = __ldg(site + sid);
                            lr[k] = __fsqrt_rn(vc_dist2f(sp.x, sp.y, sp.z, (float)(x + k), (float)y, (float)z));
                        }
                    }
                }
            }
            auto put = [&](float* base, const float* v)
            {
                if (VEC)
                    __stcs(reinterpret_cast<float4*>(base + o), make_float4(v[0], v[1], v[2], v[3]));
                else
                {
#pragma unroll
                    for (int k = 0; k < CM_VX; ++k)
                        if (x + k < nx)
                            __stcs(base + o + k, v[k]);
                }
            };
            if (edge3)
            {
                put(edge3, le[0]);
                put(edge3 + nv, le[1]);
                put(edge3 + 2 * nv, le[2]);
            }
            if (face3)
            {
                put(face3, lf[0]);
                put(face3 + nv, lf[1]);
                put(face3 + 2 * nv, lf[2]);
            }
            if (cube)
                put(cube, lc);
            if (radius)
                put(radius, lr);
        }
        nib_cur = nib_next;
    }
}

int measures_alloc(vc_ctx* c, bool want_radius)
{
    const size_t nv = (size_t)c->nx * c->ny * (size_t)(c->z1 - c->z0);
    VC_CUDA(c, c->edge3.ensure(nv * 3 * 4));
    VC_CUDA(c, c->face3.ensure(nv * 3 * 4));
    VC_CUDA(c, c->cube.ensure(nv * 4));
    if (want_radius)
        VC_CUDA(c, c->radius.ensure(nv * 4));
    const size_t smem = (size_t)3 * (CM_TH + 1) * CM_PITCH * sizeof(float4);
    if (!c->attr_measures)
    {
        VC_CUDA(c, cudaFuncSetAttribute(k_cell_measures<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        VC_CUDA(c, cudaFuncSetAttribute(k_cell_measures<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        c->attr_measures = true;
    }
    return VC_OK;
}

// measures of the anchor planes [za, zb) of the slab, on stream c->cur (buffers from measures_alloc); `alone`: nothing
// else runs beside this launch (a slab transformed as one chunk, vc_cell_measures_grid)
int measures_range(vc_ctx* c, int za, int zb, bool want_radius, bool alone)
{
    // z chunks per block: long enough to amortise the one extra plane a block loads, short enough that the grid is
    // many waves of the SMs.  A launch that has the device to itself pays its tail in full and wants ~ 110 blocks per
    // SM (measured on a 127-plane slab of 1024^2: 1.25 ms at 32 planes per block = 4096 blocks, 1.14 at 16, 1.10 at
    // 8 = 16384 blocks; twist512 whole: 0.95 / 0.89 / 0.86); the launches of a chunk pipeline overlap other kernels and
    // are best at the long chunk (1024^3 in 128-plane chunks: 21.28 ms per step at 32, 21.46 at 8).
    const int nplanes = zb - za;
    const unsigned gx = (c->nx + CM_TW - 1) / CM_TW, gy = (c->ny + CM_TH - 1) / CM_TH;
    const size_t want_blocks = (size_t)c->sm_count * (alone ? 110 : 16);
    int zchunk = 32;
    while (zchunk > (alone ? 8 : 4) && (size_t)gx * gy * ((nplanes + zchunk - 1) / zchunk) < want_blocks)
        zchunk >>= 1;
    static const int zchunk_env = getenv("VC_MEASURE_ZCHUNK") ? atoi(getenv("VC_MEASURE_ZCHUNK")) : 0; // A/B runs
    if (zchunk_env > 0)
        zchunk = zchunk_env;
    dim3 grid(gx, gy, (nplanes + zchunk - 1) / zchunk);
    if (grid.y > 65535u || grid.z > 65535u)
        return vc_fail(c, VC_ERR_UNSUPPORTED, "grid too large for the measures kernel");
    const size_t smem = (size_t)3 * (CM_TH + 1) * CM_PITCH * sizeof(float4);
    if ((c->nx & 3) == 0)
        VC_LAUNCH(c, "cell_measures", k_cell_measures<true>, grid, 32 * CM_TH, smem, c->id.as<int>(), c->d2.as<u32>(),
                  c->bits.as<u32>(), c->wr, c->site_xyz.as<float4>(), c->nx, c->ny, c->z0, c->z1, c->zc, c->zlo, za, zb,
                  zchunk, c->lattice ? 1 : 0, c->edge3.as<float>(), c->face3.as<float>(), c->cube.as<float>(),
                  want_radius ? c->radius.as<float>() : nullptr);
    else
        VC_LAUNCH(c, "cell_measures", k_cell_measures<false>, grid, 32 * CM_TH, smem, c->id.as<int>(), c->d2.as<u32>(),
                  c->bits.as<u32>(), c->wr, c->site_xyz.as<float4>(), c->nx, c->ny, c->z0, c->z1, c->zc, c->zlo, za, zb,
                  zchunk, c->lattice ? 1 : 0, c->edge3.as<float>(), c->face3.as<float>(), c->cube.as<float>(),
                  want_radius ? c->radius.as<float>() : nullptr);
    return VC_OK;
}

int st_measures(vc_ctx* c, bool want_radius)
{
    if (!c->have_closest || !c->have_inside)
        return vc_fail(c, VC_ERR_STATE, "vc_cell_measures_grid needs vc_classify_grid and vc_closest_grid");
    if (c->zhi < c->zc)
        return vc_fail(c, VC_ERR_STATE, "inside flags do not cover the halo plane");
    VC_TRY(measures_alloc(c, want_radius));
    VC_TRY(measures_range(c, c->z0, c->z1, want_radius, true));
    VC_CUDA(c, cudaGetLastError());
    c->have_measures = true;
    return VC_OK;
}

// =============================================================================================
// a4  VoroInfo::tagVert (src/voroinfo.cpp:447-454): q = M*p in double with the homogeneous divide
// of XForm.h:479-489, cast to float, round half away from zero, bounds, flag lookup.
// =============================================================================================
struct Mat16
{
    double m[16];
};

__global__ void k_classify_points(const float* __restrict__ xyz, int64_t n, Mat16 M, const u8* __restrict__ inside,
                                  int nx, int ny, int nz, u8* __restrict__ out)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    const double* xf = M.m;
    double v0 = xyz[3 * i], v1 = xyz[3 * i + 1], v2 = xyz[3 * i + 2];
#define ROW(a, b, cc, d) __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(xf[a], v0), __dmul_rn(xf[b], v1)), __dmul_rn(xf[cc], v2)), xf[d])
    double h = __ddiv_rn(1.0, ROW(3, 7, 11, 15));
    float q0 = __double2float_rn(__dmul_rn(h, ROW(0, 4, 8, 12)));
    float q1 = __double2float_rn(__dmul_rn(h, ROW(1, 5, 9, 13)));
    float q2 = __double2float_rn(__dmul_rn(h, ROW(2, 6, 10, 14)));
#undef ROW
    // (int)std::round(float): half away from zero (include/spaceinfo.h:100-102)
    int x = (int)roundf(q0), y = (int)roundf(q1), z = (int)roundf(q2);
    bool in = x >= 0 && x < nx && y >= 0 && y < ny && z >= 0 && z < nz;
    out[i] = in ? inside[(size_t)x + (size_t)nx * ((size_t)y + (size_t)ny * z)] : (u8)0;
}

int st_classify_points(vc_ctx* c, const float* xyz, int64_t n, const double* M, uint8_t* out)
{
    if (!c->have_inside)
        return vc_fail(c, VC_ERR_STATE, "vc_classify_points needs vc_classify_grid first");
    if (c->zlo != 0 || c->zhi != c->nz)
        return vc_fail(c, VC_ERR_UNSUPPORTED, "vc_classify_points needs the whole grid resident");
    if (n == 0)
        return VC_OK;
    Mat16 m;
    static const double I[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
    for (int i = 0; i < 16; ++i)
        m.m[i] = M ? M[i] : I[i];
    DevBuf din, dout;
    VC_CUDA(c, din.ensure((size_t)n * 12));
    cudaError_t e = dout.ensure((size_t)n);
    if (e == cudaSuccess)
        e = cudaMemcpyAsync(din.p, xyz, (size_t)n * 12, cudaMemcpyDefault, c->stream);
    if (e == cudaSuccess)
    {
        VC_LAUNCH(c, "classify_points", k_classify_points, vc_blocks((size_t)n, 256), 256, 0, din.as<float>(), n, m,
                  c->inside.as<u8>(), c->nx, c->ny, c->nz, dout.as<u8>());
        e = cudaMemcpyAsync(out, dout.p, (size_t)n, cudaMemcpyDefault, c->stream);
    }
    if (e == cudaSuccess)
        e = cudaStreamSynchronize(c->stream);
    din.release();
    dout.release();
    if (e != cudaSuccess)
        return vc_fail(c, VC_ERR_CUDA, "classify_points", e);
    return VC_OK;
}

// =============================================================================================
// a7 / a8 on a Voronoi complex
// =============================================================================================
__global__ void k_face_lambda(const float4* __restrict__ site, const int2* __restrict__ pairs, int64_t nf, int64_t ns,
                              float* __restrict__ out)
{
    int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= nf)
        return;
    int2 p = pairs[f];
    float r = 0.0f;
    if (p.x >= 0 && p.x < ns && p.y >= 0 && p.y < ns)
    {
        float4 a = __ldg(site + p.x), b = __ldg(site + p.y);
        r = __fsqrt_rn(vc_dist2f(a.x, a.y, a.z, b.x, b.y, b.z));
    }
    out[f] = r;
}

__global__ void k_vertex_radii(const float4* __restrict__ site, const float* __restrict__ v, const int* __restrict__ sv,
                               int64_t nv, int64_t ns, float* __restrict__ out)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nv)
        return;
    int s = sv[i];
    float r = 0.0f;
    if (s >= 0 && s < ns)
    {
        float4 a = __ldg(site + s);
        r = __fsqrt_rn(vc_dist2f(a.x, a.y, a.z, v[3 * i], v[3 * i + 1], v[3 * i + 2])); // dist(site_p, v_p)
    }
    out[i] = r;
}

__global__ void k_segment_max(const int* __restrict__ off, const int* __restrict__ items, int64_t n,
                              const float* __restrict__ value, const u8* __restrict__ valid, float* __restrict__ out)
{
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n)
        return;
    float m = 0.0f;
    for (int k = off[e]; k < off[e + 1]; ++k)
    {
        int f = items[k];
        if (valid && !valid[f])
            continue;
        float v = value[f];
        m = v > m ? v : m; // std::max(m, v)
    }
    out[e] = m;
}

// small helper: temp device arrays for host-pointer operators
struct Tmp
{
    std::vector<DevBuf> bufs;
    ~Tmp()
    {
        for (auto& b : bufs)
            b.release();
    }
    void* up(vc_ctx* c, const void* host, size_t bytes, cudaError_t& e)
    {
        bufs.emplace_back();
        if (e == cudaSuccess)
            e = bufs.back().ensure(bytes);
        if (e == cudaSuccess && host)
            e = cudaMemcpyAsync(bufs.back().p, host, bytes, cudaMemcpyDefault, c->stream);
        return bufs.back().p;
    }
};

int st_face_lambda(vc_ctx* c, const int32_t* pairs, int64_t nf, float* out)
{
    if (!c->have_sites)
        return vc_fail(c, VC_ERR_STATE, "vc_face_lambda needs sites");
    if (nf == 0)
        return VC_OK;
    Tmp t;
    cudaError_t e = cudaSuccess;
    int2* dp = (int2*)t.up(c, pairs, (size_t)nf * 8, e);
    float* dout = (float*)t.up(c, nullptr, (size_t)nf * 4, e);
    if (e == cudaSuccess)
    {
        VC_LAUNCH(c, "face_lambda", k_face_lambda, vc_blocks((size_t)nf, 256), 256, 0, c->site_xyz.as<float4>(), dp, nf,
                  c->nsites, dout);
        e = cudaMemcpyAsync(out, dout, (size_t)nf * 4, cudaMemcpyDefault, c->stream);
    }
    if (e == cudaSuccess)
        e = cudaStreamSynchronize(c->stream);
    if (e != cudaSuccess)
        return vc_fail(c, VC_ERR_CUDA, "face_lambda", e);
    return VC_OK;
}

int st_vertex_radii(vc_ctx* c, const float* v, int64_t nv, const int32_t* site_of_v, float* out)
{
    if (!c->have_sites)
        return vc_fail(c, VC_ERR_STATE, "vc_vertex_radii needs sites");
    if (nv == 0)
        return VC_OK;
    Tmp t;
    cudaError_t e = cudaSuccess;
    float* dv = (float*)t.up(c, v, (size_t)nv * 12, e);
    int* ds = (int*)t.up(c, site_of_v, (size_t)nv * 4, e);
    float* dout = (float*)t.up(c, nullptr, (size_t)nv * 4, e);
    if (e == cudaSuccess)
    {
        VC_LAUNCH(c, "vertex_radii", k_vertex_radii, vc_blocks((size_t)nv, 256), 256, 0, c->site_xyz.as<float4>(), dv,
                  ds, nv, c->nsites, dout);
        e = cudaMemcpyAsync(out, dout, (size_t)nv * 4, cudaMemcpyDefault, c->stream);
    }
    if (e == cudaSuccess)
        e = cudaStreamSynchronize(c->stream);
    if (e != cudaSuccess)
        return vc_fail(c, VC_ERR_CUDA, "vertex_radii", e);
    return VC_OK;
}

int st_segment_max(vc_ctx* c, const int32_t* off, const int32_t* items, int64_t n, const float* value,
                   int64_t nvalue, const uint8_t* valid, float* out)
{
    if (n == 0)
        return VC_OK;
    int32_t nitems = off[n];
    Tmp t;
    cudaError_t e = cudaSuccess;
    int* doff = (int*)t.up(c, off, (size_t)(n + 1) * 4, e);
    int* dit = (int*)t.up(c, items, (size_t)(nitems > 0 ? nitems : 1) * 4, e);
    float* dval = (float*)t.up(c, value, (size_t)(nvalue > 0 ? nvalue : 1) * 4, e);
    u8* dvalid = valid ? (u8*)t.up(c, valid, (size_t)(nvalue > 0 ? nvalue : 1), e) : nullptr;
    float* dout = (float*)t.up(c, nullptr, (size_t)n * 4, e);
    if (e == cudaSuccess)
    {
        VC_LAUNCH(c, "segment_max", k_segment_max, vc_blocks((size_t)n, 256), 256, 0, doff, dit, n, dval, dvalid, dout);
        e = cudaMemcpyAsync(out, dout, (size_t)n * 4, cudaMemcpyDefault, c->stream);
    }
    if (e == cudaSuccess)
        e = cudaStreamSynchronize(c->stream);
    if (e != cudaSuccess)
        return vc_fail(c, VC_ERR_CUDA, "segment_max", e);
    return VC_OK;
}
