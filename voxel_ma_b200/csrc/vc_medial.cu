// vc_medial.cu -- SURVEY section 8(f-4): the medial complex of the DENSE product, straight from the id grid.
//
// What it replaces: the reference gets the Voronoi diagram of the boundary samples from TetGen
// (src/highlevelalgo.cpp:503-529, 45-50 % of a vol2ma run) and keeps the part inside the shape
// (src/voroinfo.cpp:128-139, 624-729).  On the dense grid the same object is read off the closest-site ids
// (the grid-cell <-> Voronoi-cell dictionary of SURVEY section 0): a grid EDGE whose two end vertices have different
// closest sites is crossed by the Voronoi FACE of those two sites.  The dual of such an edge is a quad whose four
// corners are the centres of the four grid cubes around the edge; the quads of all crossing edges form a cubical
// 2-complex -- the discrete inside part of the Voronoi diagram -- in exactly the form cellcomplex (src/cellcomplex.cpp:
// 364-491, finalize) and CellComplexThinning (src/ccthin.cpp:201-424) take: vertices + polygon faces with a measure.
//
// Validity mirrors the reference's rule that only cells whose vertices are all inside count
// (include/voroinfo_imp.h:26-34): a grid cube is valid iff its 8 vertices are inside, a quad is emitted iff the
// 4 cubes around its edge exist and are valid (= the 2 x 3 x 3 block of 18 vertices around the edge is inside),
// so every boundary edge and corner of an emitted quad is valid too and the complex is closed.
// Measure: lambda(quad) = lambdaForFace(s(id a), s(id b)) in float32 (include/measureforMA_imp.h:1-4), the value the
// edge3 plane holds for that grid edge.  This complex is NOT the reference's (TetGen's cells are general polygons,
// these are unit quads): it is judged on counts, Euler characteristic and lambda range (tests/test_gpu_medial.py),
// never on bytes -- PARITY UNPINNED for the complex itself, its arithmetic is the pinned lambda.
//
// Order of the records: ascending (z, y, x, axis) -- fixed by a prefix over per-row counts, no atomics.
#include "vc_internal.h"

__device__ __forceinline__ float md_lambda(const float4 a, const float4 b)
{ // trimesh::dist (3rdparty/trimesh2/include/Vec.h:1128-1143): float, x y z in order, no contraction
    float t = __fsub_rn(b.x, a.x);
    float d2 = __fmul_rn(t, t);
    t = __fsub_rn(b.y, a.y);
    d2 = __fadd_rn(d2, __fmul_rn(t, t));
    t = __fsub_rn(b.z, a.z);
    d2 = __fadd_rn(d2, __fmul_rn(t, t));
    return __fsqrt_rn(d2);
}

// bit x of the AND of the bit rows (y, z), y in [ya, yb], z in [za, zb]; rows outside the grid read 0
__device__ __forceinline__ u32 md_and_rows(const u32* __restrict__ bits, int wr, int ny, int nz, int zlo, int ya, int yb, int za,
                                           int zb, int w)
{
    if (w < 0 || w >= wr || ya < 0 || yb >= ny || za < 0 || zb >= nz)
        return 0u;
    u32 r = 0xFFFFFFFFu;
    for (int z = za; z <= zb; ++z)
        for (int y = ya; y <= yb; ++y)
            r &= __ldg(bits + ((size_t)(z - zlo) * ny + y) * (size_t)wr + w);
    return r;
}

// One warp per row (y, z) of the owned planes, 32 vertices x per round.  EMIT = false: cnt[row] = quads of the row.
template <bool EMIT>
__global__ void __launch_bounds__(256)
    k_medial_quads(const u32* __restrict__ bits, int wr, int nx, int ny, int nz, int z0, int z1, int zlo,
                   const int* __restrict__ id, const float4* __restrict__ site, u32* __restrict__ cnt,
                   const u32* __restrict__ rowpre, size_t cap, u32* __restrict__ anchor, u8* __restrict__ axis,
                   int* __restrict__ ida, int* __restrict__ idb, float* __restrict__ lam)
{
    const size_t row = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    const size_t nrows = (size_t)(z1 - z0) * ny;
    if (row >= nrows)
        return;
    const int y = (int)(row % ny), z = z0 + (int)(row / ny);
    const size_t plane = (size_t)nx * ny;
    const int* idrow = id + (size_t)(z - z0) * plane + (size_t)y * nx;
    u32 total = 0;
    size_t pos = EMIT ? rowpre[row] : 0;
    for (int w = 0; w * 32 < nx; ++w)
    {
        // X edge (x, x+1): vertices x, x+1 in rows y-1..y+1, z-1..z+1;  Y edge: x-1..x+1 in rows y..y+1, z-1..z+1;
        // Z edge: x-1..x+1 in rows y-1..y+1, z..z+1
        const u32 ax = md_and_rows(bits, wr, ny, nz, zlo, y - 1, y + 1, z - 1, z + 1, w);
        const u32 axn = md_and_rows(bits, wr, ny, nz, zlo, y - 1, y + 1, z - 1, z + 1, w + 1);
        const u32 mx = ax & ((ax >> 1) | (axn << 31));
        u32 my = 0, mz = 0;
        {
            const u32 c = md_and_rows(bits, wr, ny, nz, zlo, y, y + 1, z - 1, z + 1, w);
            const u32 p = md_and_rows(bits, wr, ny, nz, zlo, y, y + 1, z - 1, z + 1, w - 1);
            const u32 n = md_and_rows(bits, wr, ny, nz, zlo, y, y + 1, z - 1, z + 1, w + 1);
            my = c & ((c >> 1) | (n << 31)) & ((c << 1) | (p >> 31));
        }
        {
            const u32 c = md_and_rows(bits, wr, ny, nz, zlo, y - 1, y + 1, z, z + 1, w);
            const u32 p = md_and_rows(bits, wr, ny, nz, zlo, y - 1, y + 1, z, z + 1, w - 1);
            const u32 n = md_and_rows(bits, wr, ny, nz, zlo, y - 1, y + 1, z, z + 1, w + 1);
            mz = c & ((c >> 1) | (n << 31)) & ((c << 1) | (p >> 31));
        }
        const int x = 32 * w + lane;
        u32 m = 0; // bit a: the dual quad of edge (v, v + e_a) exists
        int i0 = 0, i1[3] = {0, 0, 0};
        if (x < nx && (((mx | my | mz) >> lane) & 1u))
        {
            i0 = __ldg(idrow + x);
            if ((mx >> lane) & 1u)
            {
                i1[0] = __ldg(idrow + x + 1);
                m |= (u32)(i1[0] != i0);
            }
            if ((my >> lane) & 1u)
            {
                i1[1] = __ldg(idrow + x + nx);
                m |= (u32)(i1[1] != i0) << 1;
            }
            if ((mz >> lane) & 1u)
            {
                i1[2] = __ldg(idrow + x + plane);
                m |= (u32)(i1[2] != i0) << 2;
            }
        }
        const int c = __popc(m);
        int incl = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1)
        {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o)
                incl += t;
        }
        const int wtot = __shfl_sync(0xffffffffu, incl, 31);
        if (EMIT && m)
        {
            size_t o = pos + (size_t)(incl - c);
            const float4 s0 = __ldg(site + i0);
            for (int a = 0; a < 3; ++a)
                if ((m >> a) & 1u)
                {
                    if (o < cap)
                    {
                        anchor[o] = (u32)((size_t)(z - z0) * plane + (size_t)y * nx + x);
                        axis[o] = (u8)a;
                        ida[o] = i0;
                        idb[o] = i1[a];
                        lam[o] = md_lambda(s0, __ldg(site + i1[a]));
                    }
                    ++o;
                }
        }
        pos += (size_t)wtot;
        total += (u32)wtot;
    }
    if (!EMIT && lane == 0)
        cnt[row] = total;
}

static int medial_count(vc_ctx* c, int64_t* n)
{
    if (!c->have_closest || !c->have_inside || !c->lattice)
        return vc_fail(c, VC_ERR_STATE, "vc_medial_quads needs vc_classify_grid and the closest sites of a lattice site set");
    if (c->zlo > (c->z0 > 0 ? c->z0 - 1 : 0) || c->zhi < (c->z1 < c->nz ? c->z1 + 1 : c->nz) || c->zc < (c->z1 < c->nz ? c->z1 + 1 : c->nz))
        return vc_fail(c, VC_ERR_STATE, "vc_medial_quads: the planes around the slab are not resident");
    const size_t nrows = (size_t)(c->z1 - c->z0) * c->ny;
    VC_CUDA(c, c->medial_pre.ensure((nrows + 2) * 4));
    u32* pre = c->medial_pre.as<u32>();
    VC_CUDA(c, cudaMemsetAsync(pre + nrows, 0, 4, c->stream));
    VC_LAUNCH(c, "medial_count", k_medial_quads<false>, vc_blocks(nrows * 32, 256), 256, 0, c->bits.as<u32>(), c->wr, c->nx, c->ny,
              c->nz, c->z0, c->z1, c->zlo, c->id.as<int>(), c->site_xyz.as<float4>(), pre, (const u32*)nullptr, (size_t)0, (u32*)nullptr,
              (u8*)nullptr, (int*)nullptr, (int*)nullptr, (float*)nullptr);
    VC_TRY(vc_exclusive_scan_u32(c, pre, (int64_t)nrows + 1));
    u32 tot = 0;
    VC_CUDA(c, cudaMemcpyAsync(&tot, pre + nrows, 4, cudaMemcpyDeviceToHost, c->stream));
    VC_CUDA(c, cudaStreamSynchronize(c->stream));
    *n = (int64_t)tot;
    return VC_OK;
}

extern "C"
{
    int vc_medial_quads_count(vc_ctx* c, int64_t* nquads)
    {
        if (!c || !nquads)
            return VC_ERR_INVALID;
        VC_CUDA(c, cudaSetDevice(c->device));
        return medial_count(c, nquads);
    }

    int vc_medial_quads(vc_ctx* c, int64_t cap, uint32_t* anchor, uint8_t* axis, int32_t* site_a, int32_t* site_b, float* lambda,
                        int64_t* nquads)
    {
        if (!c || cap < 0 || !anchor || !axis || !site_a || !site_b || !lambda)
            return VC_ERR_INVALID;
        VC_CUDA(c, cudaSetDevice(c->device));
        int64_t n = 0;
        VC_TRY(medial_count(c, &n));
        if (nquads)
            *nquads = n;
        if (n > cap)
            return vc_fail(c, VC_ERR_NOMEM, "vc_medial_quads: more quads than the capacity given (vc_medial_quads_count tells how many)");
        if (n == 0)
            return VC_OK;
        DevBuf d;
        VC_CUDA(c, d.ensure((size_t)n * 17 + 64));
        u32* dan = d.as<u32>();
        int* da = (int*)(dan + n);
        int* db = da + n;
        float* dl = (float*)(db + n);
        u8* dax = (u8*)(dl + n);
        const size_t nrows = (size_t)(c->z1 - c->z0) * c->ny;
        VC_LAUNCH(c, "medial_quads", k_medial_quads<true>, vc_blocks(nrows * 32, 256), 256, 0, c->bits.as<u32>(), c->wr, c->nx, c->ny, c->nz,
                  c->z0, c->z1, c->zlo, c->id.as<int>(), c->site_xyz.as<float4>(), (u32*)nullptr, c->medial_pre.as<u32>(), (size_t)n, dan, dax,
                  da, db, dl);
        cudaError_t e = cudaMemcpyAsync(anchor, dan, (size_t)n * 4, cudaMemcpyDefault, c->stream);
        e = e == cudaSuccess ? cudaMemcpyAsync(site_a, da, (size_t)n * 4, cudaMemcpyDefault, c->stream) : e;
        e = e == cudaSuccess ? cudaMemcpyAsync(site_b, db, (size_t)n * 4, cudaMemcpyDefault, c->stream) : e;
        e = e == cudaSuccess ? cudaMemcpyAsync(lambda, dl, (size_t)n * 4, cudaMemcpyDefault, c->stream) : e;
        e = e == cudaSuccess ? cudaMemcpyAsync(axis, dax, (size_t)n, cudaMemcpyDefault, c->stream) : e;
        e = e == cudaSuccess ? cudaStreamSynchronize(c->stream) : e;
        d.release();
        if (e != cudaSuccess)
            return vc_fail(c, VC_ERR_CUDA, "vc_medial_quads copy", e);
        return VC_OK;
    }
}
