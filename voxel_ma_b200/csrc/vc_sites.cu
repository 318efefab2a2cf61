// vc_sites.cu -- stage 1 (classify) and stage 1'' (boundary samples): classification kernels,
// per-corner site detection, the device radix sort that numbers the sites in the reference's
// first-encounter order, and the per-z-line site lists the closest-site transform starts from.
//
// Replaces: SpaceConverter::get_occupancy_at_vox over all voxels (include/spaceinfo.h:122-129) and
// Surfacer::extractBoundaryVts (src/surfacing.cpp:223-321).
#include <cooperative_groups.h>

#include <cstdlib>
#include <cstring>

#include "vc_internal.h"

namespace cg = cooperative_groups;

// =============================================================================================
// K1  classify: inside <=> value > 0.0  (NaN, 0, -0 are outside).  5 B / voxel: one 128-bit load,
// one 32-bit store per 4 voxels.  When rows are whole 32-voxel words (nx % 32 == 0) the same pass
// also packs the flags into the occupancy bit rows the site detection works on: each thread's 4
// flags are a nibble, 8 neighbouring lanes are OR-combined with three shuffles into one word.
// =============================================================================================
template <bool PACK>
__global__ void __launch_bounds__(256)
    k_classify_f32(const float* __restrict__ vol, u8* __restrict__ inside, u32* __restrict__ bits, size_t n, int words_per_row,
                   int wr)
{
    size_t i4 = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    const size_t stride = (size_t)gridDim.x * blockDim.x * 4;
    for (; i4 + 3 < n; i4 += stride)
    {
        float4 v = __ldcs(reinterpret_cast<const float4*>(vol + i4));
        uchar4 o;
        o.x = v.x > 0.0f;
        o.y = v.y > 0.0f;
        o.z = v.z > 0.0f;
        o.w = v.w > 0.0f;
        *reinterpret_cast<uchar4*>(inside + i4) = o;
        if (PACK)
        {
            const int lane = threadIdx.x & 31;
            u32 w = ((u32)o.x | ((u32)o.y << 1) | ((u32)o.z << 2) | ((u32)o.w << 3)) << (4 * (lane & 7));
            const u32 grp = 0xFFu << (lane & 24); // the 8 lanes of one word leave the loop together
            w |= __shfl_xor_sync(grp, w, 1);
            w |= __shfl_xor_sync(grp, w, 2);
            w |= __shfl_xor_sync(grp, w, 4);
            if ((lane & 7) == 0)
            {
                size_t wflat = i4 >> 5;
                size_t row = wflat / (size_t)words_per_row;
                bits[row * (size_t)wr + (wflat - row * (size_t)words_per_row)] = w;
            }
        }
    }
    // tail (n not a multiple of 4): the thread that would own the last partial vector
    if (!PACK && i4 < n && i4 + 3 >= n)
        for (size_t i = i4; i < n; ++i)
            inside[i] = vol[i] > 0.0f;
}

// MRC mode 0 volumes (signed bytes; the reference's reader widens them to double the same way, isosurface_tao/reader.h:
// 235-239): inside <=> value > 0.  2 B / voxel; a thread takes 16 voxels (128-bit load, 128-bit store of the flags), two
// neighbouring lanes make one 32-voxel bit word.  PACK needs whole-word rows (nx % 32 == 0).
template <bool PACK>
__global__ void __launch_bounds__(256)
    k_classify_i8(const signed char* __restrict__ vol, u8* __restrict__ inside, u32* __restrict__ bits, size_t n, int words_per_row, int wr)
{
    if (PACK)
    {
        size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 16;
        const size_t stride = (size_t)gridDim.x * blockDim.x * 16;
        for (; i + 15 < n; i += stride) // n is a multiple of 32: both lanes of a word take the same trips
        {
            const int4 v = __ldcs(reinterpret_cast<const int4*>(vol + i));
            const u32 in[4] = {(u32)v.x, (u32)v.y, (u32)v.z, (u32)v.w};
            u32 out[4], half = 0;
#pragma unroll
            for (int g = 0; g < 4; ++g)
            {
                u32 o = 0;
#pragma unroll
                for (int b = 0; b < 4; ++b)
                {
                    const u32 f = (signed char)((in[g] >> (8 * b)) & 0xFFu) > 0 ? 1u : 0u;
                    o |= f << (8 * b);
                    half |= f << (4 * g + b);
                }
                out[g] = o;
            }
            *reinterpret_cast<uint4*>(inside + i) = make_uint4(out[0], out[1], out[2], out[3]);
            const u32 other = __shfl_xor_sync(0xFFFFFFFFu, half, 1);
            if ((threadIdx.x & 1) == 0)
            {
                const size_t wflat = i >> 5, row = wflat / (size_t)words_per_row;
                bits[row * (size_t)wr + (wflat - row * (size_t)words_per_row)] = half | (other << 16);
            }
        }
    }
    else
    {
        size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
        for (; i < n; i += (size_t)gridDim.x * blockDim.x)
            inside[i] = vol[i] > 0;
    }
}

// Occupancy bit rows from the byte flags, any nx: one thread per word (row, w).
__global__ void __launch_bounds__(256) k_pack_bits(const u8* __restrict__ inside, u32* __restrict__ bits, size_t nrows, int nx, int wr)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nrows * (size_t)wr)
        return;
    size_t row = i / (size_t)wr;
    int w = (int)(i - row * (size_t)wr);
    int cnt = nx - 32 * w;
    cnt = cnt > 32 ? 32 : cnt;
    const u8* src = inside + row * (size_t)nx + 32 * (size_t)w;
    u32 word = 0;
    if ((nx & 3) == 0)
    { // rows are 4-byte aligned and cnt is a multiple of 4
        for (int k = 0; 4 * k < cnt; ++k)
        {
            u32 v = __ldg(reinterpret_cast<const u32*>(src) + k);
            u32 nib = (v & 1u) | ((v >> 7) & 2u) | ((v >> 14) & 4u) | ((v >> 21) & 8u);
            word |= nib << (4 * k);
        }
    }
    else
        for (int b = 0; b < cnt; ++b)
            word |= (__ldg(src + b) ? 1u : 0u) << b;
    bits[i] = word;
}

// flags -> bit rows for the resident planes (used when classification did not pack them itself)
static int pack_bits(vc_ctx* c)
{
    const size_t nrows = (size_t)c->ny * (size_t)(c->zhi - c->zlo);
    c->wr = c->nx / 32 + 1;
    VC_CUDA(c, c->bits.ensure(nrows * (size_t)c->wr * 4 + 16));
    VC_LAUNCH(c, "pack_bits", k_pack_bits, vc_blocks(nrows * (size_t)c->wr, 256), 256, 0, c->inside.as<u8>(), c->bits.as<u32>(),
              nrows, c->nx, c->wr);
    return VC_OK;
}

// Tao's in-memory order double[x][y][z] (z fastest) -> flags [z][y][x]; 32x32 (x,z) tile per y.
__global__ void k_classify_f64_zfast(const double* __restrict__ vol, u8* __restrict__ inside, int nx, int ny, int nz)
{
    __shared__ u8 tile[32][33];
    int y = blockIdx.z;
    int x0 = blockIdx.x * 32, zb = blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += blockDim.y)
    {
        int x = x0 + r, z = zb + threadIdx.x;
        if (x < nx && z < nz)
            tile[r][threadIdx.x] = vol[((size_t)x * ny + y) * nz + z] > 0.0;
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y)
    {
        int z = zb + r, x = x0 + threadIdx.x;
        if (x < nx && z < nz)
            inside[(size_t)x + (size_t)nx * ((size_t)y + (size_t)ny * z)] = tile[threadIdx.x][r];
    }
}

// Classification in two steps so that a host-to-device upload can be classified plane chunk by
// plane chunk while later chunks are still on the wire (vc_compact.cu): st_classify_begin sizes the
// buffers and clears the pad word of every bit row, st_classify_planes handles voxel planes
// [za, zb) of the resident range on stream c->cur.
bool st_classify_chunkable(const vc_ctx* c)
{
    return (c->nx & 31) == 0 || (c->vol_i8 ? true : (((size_t)c->nx * c->ny) & 3) == 0);
}

int st_classify_begin(vc_ctx* c)
{
    if (!c->have_vol)
        return vc_fail(c, VC_ERR_STATE, "vc_classify_grid: no volume uploaded");
    const size_t nrows = (size_t)c->ny * (size_t)(c->zhi - c->zlo);
    VC_CUDA(c, c->inside.ensure((size_t)c->nx * nrows + 16));
    c->wr = c->nx / 32 + 1;
    VC_CUDA(c, c->bits.ensure(nrows * (size_t)c->wr * 4 + 16));
    if ((c->nx & 31) == 0)
        VC_CUDA(c, cudaMemsetAsync(c->bits.p, 0, nrows * (size_t)c->wr * 4, c->cur)); // the pad word of every row
    return VC_OK;
}

int st_classify_planes(vc_ctx* c, int za, int zb)
{
    const size_t nrows = (size_t)c->ny * (size_t)(zb - za), row0 = (size_t)c->ny * (size_t)(za - c->zlo);
    const size_t n = (size_t)c->nx * nrows, off = (size_t)c->nx * row0;
    if (n == 0)
        return VC_OK;
    size_t want = (n / 4 + 255) / 256 + 1, cap = (size_t)c->sm_count * 16;
    unsigned blocks = (unsigned)(want < cap ? want : cap);
    if (c->vol_i8)
    {
        if ((c->nx & 31) == 0)
            VC_LAUNCH(c, "classify_i8", k_classify_i8<true>, blocks, 256, 0, c->vol.as<signed char>() + off, c->inside.as<u8>() + off,
                      c->bits.as<u32>() + row0 * (size_t)c->wr, n, c->nx / 32, c->wr);
        else
        {
            VC_LAUNCH(c, "classify_i8", k_classify_i8<false>, blocks, 256, 0, c->vol.as<signed char>() + off,
                      c->inside.as<u8>() + off, (u32*)nullptr, n, 0, 0);
            VC_LAUNCH(c, "pack_bits", k_pack_bits, vc_blocks(nrows * (size_t)c->wr, 256), 256, 0, c->inside.as<u8>() + off,
                      c->bits.as<u32>() + row0 * (size_t)c->wr, nrows, c->nx, c->wr);
        }
    }
    else if ((c->nx & 31) == 0)
        VC_LAUNCH(c, "classify_f32", k_classify_f32<true>, blocks, 256, 0, c->vol.as<float>() + off, c->inside.as<u8>() + off,
                  c->bits.as<u32>() + row0 * (size_t)c->wr, n, c->nx / 32, c->wr);
    else
    { // (a plane range that does not start on a 16-byte boundary must be the whole resident range: st_classify_chunkable)
        VC_LAUNCH(c, "classify_f32", k_classify_f32<false>, blocks, 256, 0, c->vol.as<float>() + off, c->inside.as<u8>() + off,
                  (u32*)nullptr, n, 0, 0);
        VC_LAUNCH(c, "pack_bits", k_pack_bits, vc_blocks(nrows * (size_t)c->wr, 256), 256, 0, c->inside.as<u8>() + off,
                  c->bits.as<u32>() + row0 * (size_t)c->wr, nrows, c->nx, c->wr);
    }
    VC_CUDA(c, cudaGetLastError());
    return VC_OK;
}

int st_classify(vc_ctx* c)
{
    VC_TRY(st_classify_begin(c));
    VC_TRY(st_classify_planes(c, c->zlo, c->zhi));
    c->have_inside = true;
    c->have_sites = c->have_closest = c->have_measures = false;
    return VC_OK;
}

int st_upload_f64_zfast(vc_ctx* c, const double* vol)
{
    if (!c->have_grid || c->z0 != 0 || c->z1 != c->nz)
        return vc_fail(c, VC_ERR_UNSUPPORTED, "vc_volume_upload_f64_zfast needs a ctx that owns the whole grid");
    size_t n = (size_t)c->nx * c->ny * c->nz;
    DevBuf tmp;
    VC_CUDA(c, tmp.ensure(n * sizeof(double)));
    cudaError_t e = cudaMemcpyAsync(tmp.p, vol, n * sizeof(double), cudaMemcpyDefault, c->stream);
    if (e == cudaSuccess)
        e = c->inside.ensure(n + 16);
    if (e != cudaSuccess)
    {
        tmp.release();
        return vc_fail(c, VC_ERR_CUDA, "upload f64", e);
    }
    c->zlo = 0;
    c->zhi = c->nz;
    dim3 grid((c->nx + 31) / 32, (c->nz + 31) / 32, c->ny), block(32, 8);
    VC_LAUNCH(c, "classify_f64_zfast", k_classify_f64_zfast, grid, block, 0, tmp.as<double>(), c->inside.as<u8>(),
              c->nx, c->ny, c->nz);
    int ps = pack_bits(c);
    e = cudaStreamSynchronize(c->stream);
    tmp.release();
    if (e != cudaSuccess)
        return vc_fail(c, VC_ERR_CUDA, "classify f64", e);
    if (ps != VC_OK)
        return ps;
    c->have_vol = false; // only the flags are kept for a double volume
    c->have_inside = true;
    c->have_sites = c->have_closest = c->have_measures = false;
    return VC_OK;
}

// =============================================================================================
// K2  site detection on the occupancy bit rows.  A corner (cx,cy,cz) is a site iff its 8 incident
// voxels (out of bounds = 0) are neither all 0 nor all 1.  One thread owns the 32 corners
// cx = 32w .. 32w+31 of corner row (cy,cz): with A = OR and B = AND of the four voxel rows
// (y in {cy-1,cy}, z in {cz-1,cz}; a row out of bounds reads 0), and the same words shifted by one
// voxel for x-1,
//     site word = (A | A') & ~(B & B')
// so the count pass is a handful of word operations per 32 corners.  The emit pass expands the set
// bits only: occupancy / in-bounds bytes of the 8 voxels -> first-encounter key (vc_site_key).
// Warp-aggregated append; the order of the appended records does not matter because the keys are
// unique and sorted afterwards.
// =============================================================================================
// MODE 0: count only.  MODE 1: append (key, corner) records to keys[] / corners[] (peers.cap carries their capacity).
// (Round 1 had a MODE 2 that stored every record straight into every rank's receive region; the exchange now posts
// a locally SORTED run instead, vc_peer.cu: k_peer_store_run.)
template <int MODE>
__global__ void __launch_bounds__(256)
    k_detect_sites(const u32* __restrict__ bits, int wr, int nx, int ny, int nz, int zlo, int czb, int cze,
                   u64* __restrict__ keys, u64* __restrict__ corners, u64* __restrict__ counter, VcPeerDst peers)
{
    const int CY = ny + 1;
    const size_t total = (size_t)wr * CY * (size_t)(cze - czb);
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    u32 site = 0, cur[4] = {0, 0, 0, 0}, prv[4] = {0, 0, 0, 0};
    int w = 0, cy = 0, cz = 0;
    u32 rows_in = 0; // bit (b*2+c): voxel row (cy-1+b, cz-1+c) is inside the volume
    if (i < total)
    {
        w = (int)(i % wr);
        size_t r = i / wr;
        cy = (int)(r % CY);
        cz = czb + (int)(r / CY);
#pragma unroll
        for (int k = 0; k < 4; ++k)
        {
            int y = cy - 1 + (k >> 1), z = cz - 1 + (k & 1);
            if (y >= 0 && y < ny && z >= 0 && z < nz)
            {
                rows_in |= 1u << k;
                const u32* row = bits + ((size_t)(z - zlo) * ny + y) * (size_t)wr;
                cur[k] = __ldg(row + w);
                prv[k] = w > 0 ? __ldg(row + w - 1) : 0u;
            }
        }
        const u32 A = cur[0] | cur[1] | cur[2] | cur[3], Ap = prv[0] | prv[1] | prv[2] | prv[3];
        const u32 B = cur[0] & cur[1] & cur[2] & cur[3], Bp = prv[0] & prv[1] & prv[2] & prv[3];
        const u32 any8 = A | (A << 1) | (Ap >> 31);
        const u32 all8 = B & ((B << 1) | (Bp >> 31));
        const int ncorner = nx + 1 - 32 * w; // corners of this word that exist (cx <= nx)
        const u32 cmask = ncorner >= 32 ? 0xFFFFFFFFu : (ncorner <= 0 ? 0u : ((1u << ncorner) - 1u));
        site = any8 & ~all8 & cmask;
    }
    const int cnt = __popc(site);
    // warp-aggregated reservation: exclusive prefix of cnt over the lanes, one atomic per warp
    const int lane = threadIdx.x & 31;
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1)
    {
        int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o)
            incl += t;
    }
    const int warp_total = __shfl_sync(0xffffffffu, incl, 31);
    if (warp_total == 0)
        return;
    u64 base = 0;
    if (lane == 31)
        base = atomicAdd(counter, (u64)warp_total);
    base = __shfl_sync(0xffffffffu, base, 31);
    if (MODE == 0)
        return;
    if (MODE == 1)
    { // local arrays: each lane writes its own records (staging only pays across NVLink)
        u64 pos = base + (u64)(incl - cnt);
        while (site)
        {
            const int b = __ffs(site) - 1;
            site &= site - 1;
            const int cx = 32 * w + b;
            u32 occ = 0, inb = 0;
#pragma unroll
            for (int k = 0; k < 4; ++k)
            { // voxel (cx-1+a, cy-1+(k>>1), cz-1+(k&1)) -> bit a*4 + k of occ / inb
                const u32 lo = b > 0 ? (cur[k] >> (b - 1)) & 1u : prv[k] >> 31; // x = cx-1
                const u32 hi = (cur[k] >> b) & 1u;                              // x = cx
                occ |= (lo << k) | (hi << (4 + k));
                const u32 rin = (rows_in >> k) & 1u;
                inb |= ((cx >= 1 ? rin : 0u) << k) | ((cx < nx ? rin : 0u) << (4 + k));
            }
            if (pos < peers.cap)
            { // (peers.cap carries the capacity of the local arrays in this mode)
                keys[pos] = vc_site_key(occ, inb, cx, cy, cz, ny, nz);
                corners[pos] = vc_pack_corner(cx, cy, cz);
            }
            ++pos;
        }
        return;
    }
}

int st_detect_sites(vc_ctx* c)
{
    if (!c->have_inside)
        return vc_fail(c, VC_ERR_STATE, "site extraction needs vc_classify_grid first");
    // corner planes owned by this slab: [z0,z1) plus the top plane nz for the last slab
    int czb = c->z0, cze = (c->z1 == c->nz) ? c->nz + 1 : c->z1;
    if (c->zlo > (czb > 0 ? czb - 1 : 0) || c->zhi < (cze - 1 < c->nz ? cze : c->nz))
        return vc_fail(c, VC_ERR_STATE, "resident voxel planes do not cover the slab's corner planes");
    size_t total = (size_t)c->wr * (c->ny + 1) * (size_t)(cze - czb);
    VC_CUDA(c, c->scratch.ensure(256));
    u64* counter = c->scratch.as<u64>();
    unsigned blocks = vc_blocks(total, 256);
    // One pass: the records are appended into buffers of a capacity chosen beforehand (the last count plus a quarter,
    // or, the first time, four records per voxel of the three grid faces -- any surface that is not space-filling stays
    // far below), and the counter says afterwards how many there were.  Only when it says "more than the capacity"
    // is the pass repeated, now with the exact size: no separate counting pass in front of the host read-back.
    size_t cap = c->cand_cap_hint;
    if (cap == 0)
        cap = 4 * ((size_t)c->nx * c->ny + (size_t)c->ny * (cze - czb) + (size_t)c->nx * (cze - czb)) + 4096;
    for (int attempt = 0; attempt < 2; ++attempt)
    {
        VC_CUDA(c, c->cand_key.ensure((cap + 1) * 8));
        VC_CUDA(c, c->cand_corner.ensure((cap + 1) * 8));
        VC_CUDA(c, cudaMemsetAsync(counter, 0, 16, c->stream));
        VcPeerDst lim;
        lim.cap = (u64)cap;
        VC_LAUNCH(c, "detect_sites_emit", k_detect_sites<1>, blocks, 256, 0, c->bits.as<u32>(), c->wr, c->nx, c->ny,
                  c->nz, c->zlo, czb, cze, c->cand_key.as<u64>(), c->cand_corner.as<u64>(), counter, lim);
        u64 n = 0;
        VC_CUDA(c, cudaMemcpyAsync(&n, counter, 8, cudaMemcpyDeviceToHost, c->stream));
        VC_CUDA(c, cudaStreamSynchronize(c->stream));
        c->ncand = (int64_t)n;
        if (n <= cap)
        {
            const size_t want = (size_t)n + (size_t)n / 4 + 4096;
            c->cand_cap_hint = want > c->cand_cap_hint ? want : c->cand_cap_hint;
            return VC_OK;
        }
        cap = (size_t)n; // the counter counted everything: the second attempt fits exactly
    }
    return vc_fail(c, VC_ERR_STATE, "site detection: record count changed between two passes over the same flags");
}

// =============================================================================================
// Device radix sort, LSD, 8 bits per pass, stable.  Tile = 256 threads x RS_ITEMS keys.
//   histogram  -> per-(digit, tile) counts, digit-major
//   scan       -> exclusive prefix over that table (one block)
//   scatter    -> per-warp match_any ranking keeps the order (warp, item, lane) = index order
// =============================================================================================
#define RS_ITEMS 16
#define RS_TILE (256 * RS_ITEMS)

__global__ void __launch_bounds__(256) k_rs_hist(const u64* __restrict__ keys, int64_t n, int shift, u32* __restrict__ ghist, int ntiles)
{
    __shared__ u32 h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    int64_t base = (int64_t)blockIdx.x * RS_TILE;
    for (int i = 0; i < RS_ITEMS; ++i)
    {
        int64_t idx = base + (int64_t)i * 256 + threadIdx.x;
        if (idx < n)
            atomicAdd(&h[(u32)(keys[idx] >> shift) & 255u], 1u);
    }
    __syncthreads();
    ghist[(size_t)threadIdx.x * ntiles + blockIdx.x] = h[threadIdx.x];
}

// exclusive scan of `len` u32 values in place, single block of 1024 threads
__global__ void __launch_bounds__(1024) k_rs_scan(u32* __restrict__ a, int64_t len)
{
    __shared__ u32 wsum[32];
    __shared__ u32 carry;
    if (threadIdx.x == 0)
        carry = 0;
    __syncthreads();
    const int PER = 4;
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int64_t base = 0; base < len; base += 1024 * PER)
    {
        int64_t i0 = base + (int64_t)threadIdx.x * PER;
        u32 v[PER], s = 0;
#pragma unroll
        for (int k = 0; k < PER; ++k)
        {
            v[k] = (i0 + k < len) ? a[i0 + k] : 0u;
            s += v[k];
        }
        u32 inc = s;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1)
        {
            u32 t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o)
                inc += t;
        }
        if (lane == 31)
            wsum[warp] = inc;
        __syncthreads();
        if (warp == 0)
        {
            u32 w = wsum[lane], wi = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1)
            {
                u32 t = __shfl_up_sync(0xffffffffu, wi, o);
                if (lane >= o)
                    wi += t;
            }
            wsum[lane] = wi - w; // exclusive
        }
        __syncthreads();
        u32 excl = carry + wsum[warp] + inc - s;
#pragma unroll
        for (int k = 0; k < PER; ++k)
        {
            if (i0 + k < len)
                a[i0 + k] = excl;
            excl += v[k];
        }
        __syncthreads();
        if (threadIdx.x == 1023)
            carry = excl;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(256)
    k_rs_scatter(const u64* __restrict__ kin, const u32* __restrict__ vin, u64* __restrict__ kout,
                 u32* __restrict__ vout, const u32* __restrict__ goff, int64_t n, int shift, int ntiles)
{
    __shared__ u32 wcnt[8][256];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 8 * 256; i += 256)
        (&wcnt[0][0])[i] = 0;
    __syncthreads();
    const int64_t base = (int64_t)blockIdx.x * RS_TILE + (int64_t)warp * (32 * RS_ITEMS);
    u64 k[RS_ITEMS];
    u32 rk[RS_ITEMS];
#pragma unroll
    for (int i = 0; i < RS_ITEMS; ++i)
    {
        int64_t idx = base + (int64_t)i * 32 + lane;
        bool valid = idx < n;
        k[i] = valid ? kin[idx] : 0ull;
        u32 d = (u32)(k[i] >> shift) & 255u;
        unsigned m = __match_any_sync(0xffffffffu, valid ? d : (256u + lane));
        int leader = __ffs(m) - 1;
        u32 old = 0;
        if (valid && lane == leader)
        {
            old = wcnt[warp][d];
            wcnt[warp][d] = old + __popc(m);
        }
        old = __shfl_sync(0xffffffffu, old, leader);
        rk[i] = old + __popc(m & ((1u << lane) - 1));
        __syncwarp();
    }
    __syncthreads();
    {
        int d = threadIdx.x;
        u32 run = goff[(size_t)d * ntiles + blockIdx.x];
#pragma unroll
        for (int w = 0; w < 8; ++w)
        {
            u32 cc = wcnt[w][d];
            wcnt[w][d] = run;
            run += cc;
        }
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < RS_ITEMS; ++i)
    {
        int64_t idx = base + (int64_t)i * 32 + lane;
        if (idx < n)
        {
            u32 d = (u32)(k[i] >> shift) & 255u;
            u32 pos = wcnt[warp][d] + rk[i];
            kout[pos] = k[i];
            vout[pos] = vin[idx];
        }
    }
}

// ---- all passes in ONE launch --------------------------------------------------------------------
// The sort is small (10^5 .. 10^7 records) and made of dependent passes, so three launches per pass
// (27 for the two sorts of a 1024^3 site set) cost more in launch gaps than in work.  This kernel is the
// same algorithm -- same tiles, same per-tile counts, same stable ranking, hence the same result --
// run by a persistent cooperative grid with grid-wide barriers between the phases of a pass:
//   A  per-tile digit histograms                                   (blocks stride over tiles)
//   B  exclusive scan of every digit's row of tile counts, in place  (one warp per digit) + digit totals
//   C  digit bases from the 256 totals (every block, redundantly), then the scatter of its tiles
#define CS_ITEMS 8                 // keys per thread and tile in the cooperative kernel (24 registers of keys + ranks)
#define CS_TILE (256 * CS_ITEMS)
__global__ void __launch_bounds__(256, 4)
    k_rs_sort_coop(u64* __restrict__ k0, u32* __restrict__ v0, u64* __restrict__ k1, u32* __restrict__ v1,
                   u32* __restrict__ ghist, u32* __restrict__ dtot, int64_t n, int nbits, int ntiles)
{
    cg::grid_group grid = cg::this_grid();
    __shared__ u32 wcnt[8][256];
    __shared__ u32 dbase[256];
    __shared__ u32 wsum[8];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    u64* kin = k0;
    u32* vin = v0;
    u64* kout = k1;
    u32* vout = v1;
    for (int shift = 0; shift < nbits; shift += 8)
    {
        // A: per-tile digit counts.  Per-warp counting with match_any (the scatter's own ranking step) instead of
        // shared-memory atomics: the high digits of nearly sorted keys are all equal inside a tile, which would
        // serialise 2048 atomics on one address.
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x)
        {
            for (int i = threadIdx.x; i < 8 * 256; i += 256)
                (&wcnt[0][0])[i] = 0;
            __syncthreads();
            const int64_t base = (int64_t)tile * CS_TILE + (int64_t)warp * (32 * CS_ITEMS);
            u64 k[CS_ITEMS];
#pragma unroll
            for (int i = 0; i < CS_ITEMS; ++i)
            {
                int64_t idx = base + (int64_t)i * 32 + lane;
                k[i] = idx < n ? kin[idx] : ~0ull;
            }
#pragma unroll
            for (int i = 0; i < CS_ITEMS; ++i)
            {
                const bool valid = base + (int64_t)i * 32 + lane < n;
                const u32 d = (u32)(k[i] >> shift) & 255u;
                const unsigned m = __match_any_sync(0xffffffffu, valid ? d : (256u + lane));
                if (valid && lane == __ffs(m) - 1)
                    wcnt[warp][d] += __popc(m);
                __syncwarp();
            }
            __syncthreads();
            u32 tot = 0;
#pragma unroll
            for (int w = 0; w < 8; ++w)
                tot += wcnt[w][threadIdx.x];
            ghist[(size_t)threadIdx.x * ntiles + tile] = tot;
            __syncthreads();
        }
        grid.sync();
        // B
        for (int d = blockIdx.x * 8 + warp; d < 256; d += gridDim.x * 8)
        {
            u32* row = ghist + (size_t)d * ntiles;
            u32 carry = 0;
            for (int t0 = 0; t0 < ntiles; t0 += 32)
            {
                const int t = t0 + lane;
                const u32 v = t < ntiles ? row[t] : 0u;
                u32 inc = v;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1)
                {
                    u32 x = __shfl_up_sync(0xffffffffu, inc, o);
                    if (lane >= o)
                        inc += x;
                }
                if (t < ntiles)
                    row[t] = carry + inc - v;
                carry += __shfl_sync(0xffffffffu, inc, 31);
            }
            if (lane == 0)
                dtot[d] = carry;
        }
        grid.sync();
        // C
        {
            const u32 v = dtot[threadIdx.x];
            u32 inc = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1)
            {
                u32 x = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o)
                    inc += x;
            }
            if (lane == 31)
                wsum[warp] = inc;
            __syncthreads();
            u32 before = 0;
            for (int w = 0; w < warp; ++w)
                before += wsum[w];
            dbase[threadIdx.x] = before + inc - v;
            __syncthreads();
        }
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x)
        {
            for (int i = threadIdx.x; i < 8 * 256; i += 256)
                (&wcnt[0][0])[i] = 0;
            __syncthreads();
            const int64_t base = (int64_t)tile * CS_TILE + (int64_t)warp * (32 * CS_ITEMS);
            u64 k[CS_ITEMS];
            u32 rk[CS_ITEMS], val[CS_ITEMS];
#pragma unroll
            for (int i = 0; i < CS_ITEMS; ++i)
            {
                int64_t idx = base + (int64_t)i * 32 + lane;
                const bool valid = idx < n;
                k[i] = valid ? kin[idx] : 0ull;
                val[i] = valid ? vin[idx] : 0u;
            }
#pragma unroll
            for (int i = 0; i < CS_ITEMS; ++i)
            {
                const bool valid = base + (int64_t)i * 32 + lane < n;
                u32 d = (u32)(k[i] >> shift) & 255u;
                unsigned m = __match_any_sync(0xffffffffu, valid ? d : (256u + lane));
                int leader = __ffs(m) - 1;
                u32 old = 0;
                if (valid && lane == leader)
                {
                    old = wcnt[warp][d];
                    wcnt[warp][d] = old + __popc(m);
                }
                old = __shfl_sync(0xffffffffu, old, leader);
                rk[i] = old + __popc(m & ((1u << lane) - 1));
                __syncwarp();
            }
            __syncthreads();
            {
                const int d = threadIdx.x;
                u32 run = dbase[d] + ghist[(size_t)d * ntiles + tile];
#pragma unroll
                for (int w = 0; w < 8; ++w)
                {
                    u32 cc = wcnt[w][d];
                    wcnt[w][d] = run;
                    run += cc;
                }
            }
            __syncthreads();
#pragma unroll
            for (int i = 0; i < CS_ITEMS; ++i)
            {
                if (base + (int64_t)i * 32 + lane < n)
                {
                    u32 d = (u32)(k[i] >> shift) & 255u;
                    u32 pos = wcnt[warp][d] + rk[i];
                    kout[pos] = k[i];
                    vout[pos] = val[i];
                }
            }
            __syncthreads();
        }
        grid.sync();
        u64* tk = kin;
        kin = kout;
        kout = tk;
        u32* tv = vin;
        vin = vout;
        vout = tv;
    }
}

// note: the histogram kernel tiles by (item, thread) and the scatter by (warp, item, lane); both
// cover exactly [tile*RS_TILE, (tile+1)*RS_TILE), which is all the per-tile counts depend on.
int vc_radix_sort_pairs(vc_ctx* c, int64_t n, int nbits, u64** keys_io, u32** vals_io)
{
    if (n <= 1)
        return VC_OK;
    if (n >= (int64_t)1 << 31)
        return vc_fail(c, VC_ERR_UNSUPPORTED, "radix sort: more than 2^31 records");
    int ntiles = (int)((n + RS_TILE - 1) / RS_TILE);
    VC_CUDA(c, c->shist.ensure(((size_t)256 * ntiles + 256) * sizeof(u32)));
    u64* kin = *keys_io;
    u32* vin = *vals_io;
    u64* kout = (kin == c->sk0.as<u64>()) ? c->sk1.as<u64>() : c->sk0.as<u64>();
    u32* vout = (vin == c->sv0.as<u32>()) ? c->sv1.as<u32>() : c->sv0.as<u32>();
    if (c->coop_sort < 0)
    { // once per ctx: can a cooperative grid be launched here, and how many blocks fit
        int coop = 0, per_sm = 0;
        cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, c->device);
        if (coop && cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_rs_sort_coop, 256, 0) == cudaSuccess && per_sm > 0)
            c->coop_sort = per_sm * c->sm_count;
        else
            c->coop_sort = 0;
        if (const char* e = getenv("VC_SORT"))
            if (!strcmp(e, "legacy"))
                c->coop_sort = 0;
        cudaGetLastError();
    }
    if (c->coop_sort > 0)
    {
        int ctiles = (int)((n + CS_TILE - 1) / CS_TILE);
        VC_CUDA(c, c->shist.ensure(((size_t)256 * ctiles + 256) * sizeof(u32)));
        int grid = ctiles < c->coop_sort ? ctiles : c->coop_sort;
        grid = grid < 32 ? (32 < c->coop_sort ? 32 : c->coop_sort) : grid; // phase B: 256 digit rows, one warp each
        u32* gh = c->shist.as<u32>();
        u32* dtot = gh + (size_t)256 * ctiles;
        int nb = nbits;
        void* args[] = {&kin, &vin, &kout, &vout, &gh, &dtot, &n, &nb, &ctiles};
        {
            ProfScope ps(c, "radix_sort_coop");
            VC_CUDA(c, cudaLaunchCooperativeKernel((void*)k_rs_sort_coop, dim3(grid), dim3(256), args, 0, c->cur));
        }
        const int passes = (nbits + 7) / 8;
        *keys_io = (passes & 1) ? kout : kin;
        *vals_io = (passes & 1) ? vout : vin;
        return VC_OK;
    }
    for (int shift = 0; shift < nbits; shift += 8)
    {
        VC_LAUNCH(c, "radix_hist", k_rs_hist, ntiles, 256, 0, kin, n, shift, c->shist.as<u32>(), ntiles);
        VC_LAUNCH(c, "radix_scan", k_rs_scan, 1, 1024, 0, c->shist.as<u32>(), (int64_t)256 * ntiles);
        VC_LAUNCH(c, "radix_scatter", k_rs_scatter, ntiles, 256, 0, kin, vin, kout, vout, c->shist.as<u32>(), n,
                  shift, ntiles);
        u64* tk = kin;
        kin = kout;
        kout = tk;
        u32* tv = vin;
        vin = vout;
        vout = tv;
    }
    VC_CUDA(c, cudaGetLastError());
    *keys_io = kin;
    *vals_io = vin;
    return VC_OK;
}

// =============================================================================================
// finalize: number the sites (sort by first-encounter key), build id-ordered tables and the
// per-z-line lists (line = cx*(ny+1)+cy, entries (cz<<32|id) ascending in cz).
// =============================================================================================
__global__ void k_iota_copy(const u64* __restrict__ kin, u64* __restrict__ kout, u32* __restrict__ v, int64_t n)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n)
    {
        kout[i] = kin[i];
        v[i] = (u32)i;
    }
}

// site tables in id order + the second sort's keys: key2 = line*(nz+1)+cz, value = id
__global__ void k_site_tables(const u32* __restrict__ order, const u64* __restrict__ corners_in, const u64* __restrict__ keys_sorted,
                              u64* __restrict__ site_key, u64* __restrict__ site_corner, float4* __restrict__ site_xyz,
                              u64* __restrict__ key2, u32* __restrict__ val2, int64_t n, int ny, int nz)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    u64 pc = corners_in[order ? order[i] : i];
    int cx, cy, cz;
    vc_unpack_corner(pc, cx, cy, cz);
    if (site_key)
        site_key[i] = keys_sorted ? keys_sorted[i] : (u64)i;
    site_corner[i] = pc;
    site_xyz[i] = make_float4((float)cx - 0.5f, (float)cy - 0.5f, (float)cz - 0.5f, 0.0f);
    key2[i] = ((u64)cx * (u64)(ny + 1) + (u64)cy) * (u64)(nz + 1) + (u64)cz;
    val2[i] = (u32)i;
}

__global__ void k_line_entries(const u64* __restrict__ key2_sorted, const u32* __restrict__ ids, u64* __restrict__ ent,
                               int64_t n, int nz)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n)
        ent[i] = ((key2_sorted[i] % (u64)(nz + 1)) << 32) | ids[i];
}

// line_ptr[l] = first sorted position whose line >= l  (binary search per line)
__global__ void k_line_ptr(const u64* __restrict__ key2_sorted, int64_t n, int nlines, int nz, int* __restrict__ line_ptr)
{
    int l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l > nlines)
        return;
    u64 want = (u64)l * (u64)(nz + 1);
    int64_t lo = 0, hi = n;
    while (lo < hi)
    {
        int64_t mid = (lo + hi) >> 1;
        if (key2_sorted[mid] < want)
            lo = mid + 1;
        else
            hi = mid;
    }
    line_ptr[l] = (int)lo;
}

// colmask[cy][w]: bit (cx & 31) of word w = cx >> 5 is set when z-line (cx,cy) holds at least one site
__global__ void k_line_mask(const int* __restrict__ line_ptr, int CX, int CY, int nw, u32* __restrict__ mask)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= CY * nw)
        return;
    const int cy = i / nw, w = i - cy * nw;
    u32 word = 0;
    for (int b = 0; b < 32; ++b)
    {
        const int cx = 32 * w + b;
        if (cx < CX)
        {
            const int l = cx * CY + cy;
            word |= (u32)(line_ptr[l + 1] > line_ptr[l]) << b;
        }
    }
    mask[i] = word;
}

// number of adjacent equal keys in a sorted array (duplicate detection for external site sets)
__global__ void k_count_dups(const u64* __restrict__ k, int64_t n, u64* counter)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i + 1 < n && k[i] == k[i + 1])
        atomicAdd(counter, 1ull);
}

// ---- z-line lists by counting (no second sort) -----------------------------------------------------
// The lists only need the sites of each column (cx,cy) in ascending cz.  Counting does that in O(S):
// per-column counts (atomics) -> exclusive scan = line_ptr -> every site takes a slot of its column
// (atomics, arbitrary order) -> each column orders its own few entries by (cz, id).  The final lists
// do not depend on the order in which the atomics were served.
__global__ void k_line_count(const u64* __restrict__ site_corner, int64_t n, int CY, u32* __restrict__ cnt)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    int cx, cy, cz;
    vc_unpack_corner(site_corner[i], cx, cy, cz);
    atomicAdd(&cnt[(size_t)cx * CY + cy], 1u);
}

// exclusive scan of len u32 values, several blocks: (1) each block scans its 4096 values in place and reports its
// total, (2) the totals are scanned by one block (k_rs_scan), (3) every value gets its block's offset
__global__ void __launch_bounds__(1024) k_scan_blocks(u32* __restrict__ a, int64_t len, u32* __restrict__ sums)
{
    __shared__ u32 wsum[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t i0 = ((int64_t)blockIdx.x * 1024 + threadIdx.x) * 4;
    u32 v[4], s = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k)
    {
        v[k] = (i0 + k < len) ? a[i0 + k] : 0u;
        s += v[k];
    }
    u32 inc = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1)
    {
        u32 t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o)
            inc += t;
    }
    if (lane == 31)
        wsum[warp] = inc;
    __syncthreads();
    if (warp == 0)
    {
        u32 w = wsum[lane], wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1)
        {
            u32 t = __shfl_up_sync(0xffffffffu, wi, o);
            if (lane >= o)
                wi += t;
        }
        wsum[lane] = wi - w;
        if (lane == 31)
            sums[blockIdx.x] = wi;
    }
    __syncthreads();
    u32 excl = wsum[warp] + inc - s;
#pragma unroll
    for (int k = 0; k < 4; ++k)
    {
        if (i0 + k < len)
            a[i0 + k] = excl;
        excl += v[k];
    }
}

__global__ void __launch_bounds__(1024) k_scan_add(u32* __restrict__ a, int64_t len, const u32* __restrict__ sums)
{
    const u32 off = sums[blockIdx.x];
    const int64_t i0 = ((int64_t)blockIdx.x * 1024 + threadIdx.x) * 4;
#pragma unroll
    for (int k = 0; k < 4; ++k)
        if (i0 + k < len)
            a[i0 + k] += off;
}

__global__ void k_line_fill(const u64* __restrict__ site_corner, int64_t n, int CY, u32* __restrict__ cursor, u64* __restrict__ tmp)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    int cx, cy, cz;
    vc_unpack_corner(site_corner[i], cx, cy, cz);
    const u32 pos = atomicAdd(&cursor[(size_t)cx * CY + cy], 1u);
    tmp[pos] = ((u64)(u32)cz << 32) | (u32)i;
}

// columns with up to LS_SHORT entries: one thread orders them with a compare-exchange network on the (cz, id)
// words; longer ones are queued by size class for k_line_sort_queued.  dups (nullable) counts equal cz inside a
// column (external sets with repeated points).
// queue[0..3] = number of queued columns of class 0 (5..8 entries), 1 (9..16), 2 (17..32), 3 (more); LsQueues = where
// each class's queue starts inside `queue` (from the bound  #columns with > m entries <= n / (m + 1)).
#define LS_SHORT 4
#define LS_QHDR 4
struct LsQueues
{
    u32 start[4];
};
__device__ __forceinline__ void ls_cx(u64& a, u64& b)
{
    const u64 lo = a < b ? a : b, hi = a < b ? b : a;
    a = lo;
    b = hi;
}
__global__ void __launch_bounds__(256)
    k_line_sort_short(const int* __restrict__ line_ptr, int nlines, const u64* __restrict__ tmp, u64* __restrict__ ent,
                      u32* __restrict__ queue, LsQueues qs, u64* __restrict__ dups)
{
    // queue slots are handed out per block (shared-memory counters, then one global atomic per class and block):
    // one global atomic per long column would serialise 2e5 of them on one address
    __shared__ u32 s_cnt[4], s_base[4];
    if (threadIdx.x < 4)
        s_cnt[threadIdx.x] = 0;
    __syncthreads();
    const int l = blockIdx.x * blockDim.x + threadIdx.x;
    int b = 0, cnt = 0;
    if (l < nlines)
    {
        b = line_ptr[l];
        cnt = line_ptr[l + 1] - b;
    }
    const int cls = cnt <= LS_SHORT ? -1 : cnt <= 8 ? 0 : cnt <= 16 ? 1 : cnt <= 32 ? 2 : 3;
    u32 slot = 0;
    if (cls >= 0)
        slot = atomicAdd(&s_cnt[cls], 1u);
    __syncthreads();
    if (threadIdx.x < 4 && s_cnt[threadIdx.x])
        s_base[threadIdx.x] = atomicAdd(&queue[threadIdx.x], s_cnt[threadIdx.x]);
    __syncthreads();
    if (cls >= 0)
    {
        queue[qs.start[cls] + s_base[cls] + slot] = (u32)l;
        return;
    }
    if (cnt == 0)
        return;
    if (cnt == 1)
    {
        ent[b] = tmp[b];
        return;
    }
    u64 e0 = tmp[b], e1 = tmp[b + 1], e2 = cnt > 2 ? tmp[b + 2] : VC_INF, e3 = cnt > 3 ? tmp[b + 3] : VC_INF;
    ls_cx(e0, e1); // optimal 4-input network: (0,1)(2,3)(0,2)(1,3)(1,2)
    ls_cx(e2, e3);
    ls_cx(e0, e2);
    ls_cx(e1, e3);
    ls_cx(e1, e2);
    ent[b] = e0;
    ent[b + 1] = e1;
    if (cnt > 2)
        ent[b + 2] = e2;
    if (cnt > 3)
        ent[b + 3] = e3;
    if (dups)
    {
        const u32 nd = ((e0 >> 32) == (e1 >> 32)) + (cnt > 2 && (e1 >> 32) == (e2 >> 32)) + (cnt > 3 && (e2 >> 32) == (e3 >> 32));
        if (nd)
            atomicAdd(dups, (u64)nd);
    }
}

// Columns of 5..32 entries: W lanes per column (32 / W columns per warp), lane = entry; the place of an entry is the
// number of smaller (cz, id) words in its group, counted over W shuffles (the words are distinct: ids are).  An entry
// that has a smaller word with the same cz is a repeated point.
template <int W>
__device__ __forceinline__ void ls_rank_group(const int* __restrict__ line_ptr, const u64* __restrict__ tmp, u64* __restrict__ ent,
                                              const u32* __restrict__ q, u32 nq, u32 item, int lane, u64* __restrict__ dups)
{
    const u32 j = item * (32 / W) + (u32)(lane / W);
    const int sub = lane & (W - 1);
    int b = 0, cnt = 0;
    if (j < nq)
    {
        const int l = (int)q[j];
        b = line_ptr[l];
        cnt = line_ptr[l + 1] - b;
    }
    const bool mine = sub < cnt;
    const u64 e = mine ? tmp[b + sub] : VC_INF;
    u32 rank = 0, dup = 0;
#pragma unroll
    for (int s = 0; s < W; ++s)
    {
        const u64 o = __shfl_sync(0xffffffffu, e, s, W);
        rank += o < e;
        dup |= (o < e) & ((u32)(o >> 32) == (u32)(e >> 32));
    }
    if (mine)
        ent[b + rank] = e;
    const u32 nd = __popc(__ballot_sync(0xffffffffu, mine && dup));
    if (nd && dups && lane == 0)
        atomicAdd(dups, (u64)nd);
}

// Longer columns: one warp per column.  The cz of a column's sites are distinct corner indices in [0, nz], so the
// column is ordered by presence: a bitmap of nz+1 bits in shared memory, rank = number of set bits below cz.
// A bit found already set is a repeated point (only possible for external sets): it is counted in dups, which
// makes the caller fall back to the general search -- the lists are then not used.
#define LS_WORDS 65 // 2049 bits + padding: sides up to 2048
__global__ void __launch_bounds__(256)
    k_line_sort_queued(const int* __restrict__ line_ptr, const u64* __restrict__ tmp, u64* __restrict__ ent,
                       const u32* __restrict__ queue, LsQueues qs, u64* __restrict__ dups)
{
    __shared__ u32 bitsm[8][LS_WORDS + 1], pre[8][LS_WORDS + 1];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const u32 n0 = queue[0], n1 = queue[1], n2 = queue[2], n3 = queue[3];
    const u32 i0 = (n0 + 3) >> 2, i1 = i0 + ((n1 + 1) >> 1), i2 = i1 + n2, i3 = i2 + n3; // warp-sized work items
    for (u32 it = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; it < i3; it += (gridDim.x * blockDim.x) >> 5)
    {
        if (it < i0)
        {
            ls_rank_group<8>(line_ptr, tmp, ent, queue + qs.start[0], n0, it, lane, dups);
            continue;
        }
        if (it < i1)
        {
            ls_rank_group<16>(line_ptr, tmp, ent, queue + qs.start[1], n1, it - i0, lane, dups);
            continue;
        }
        if (it < i2)
        {
            ls_rank_group<32>(line_ptr, tmp, ent, queue + qs.start[2], n2, it - i1, lane, dups);
            continue;
        }
        const int l = (int)queue[qs.start[3] + (it - i2)];
        const int b = line_ptr[l], cnt = line_ptr[l + 1] - b;
        for (int w = lane; w <= LS_WORDS; w += 32)
            bitsm[warp][w] = 0;
        __syncwarp();
        u32 nd = 0;
        for (int i = lane; i < cnt; i += 32)
        {
            const u32 cz = (u32)(tmp[b + i] >> 32);
            const u32 old = atomicOr(&bitsm[warp][cz >> 5], 1u << (cz & 31));
            nd += (old >> (cz & 31)) & 1u;
        }
        __syncwarp();
        // exclusive prefix of the word popcounts (LS_WORDS <= 96: three rounds of 32)
        u32 carry = 0;
        for (int w0 = 0; w0 <= LS_WORDS; w0 += 32)
        {
            const int w = w0 + lane;
            const u32 v = w <= LS_WORDS ? __popc(bitsm[warp][w]) : 0u;
            u32 inc = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1)
            {
                u32 t = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o)
                    inc += t;
            }
            if (w <= LS_WORDS)
                pre[warp][w] = carry + inc - v;
            carry += __shfl_sync(0xffffffffu, inc, 31);
        }
        __syncwarp();
        for (int i = lane; i < cnt; i += 32)
        {
            const u64 me = tmp[b + i];
            const u32 cz = (u32)(me >> 32);
            const u32 rank = pre[warp][cz >> 5] + __popc(bitsm[warp][cz >> 5] & ((1u << (cz & 31)) - 1u));
            ent[b + (rank < (u32)cnt ? rank : (u32)cnt - 1)] = me;
        }
        if (nd && dups)
            atomicAdd(dups, (u64)nd);
        __syncwarp();
    }
}

// exclusive prefix of len u32 values in place (device), on stream c->cur: one block for short arrays, otherwise
// block-local scans + a scan of the block totals + an add pass (a single block over 10^5..10^6 values is ~0.1 ms)
int vc_exclusive_scan_u32(vc_ctx* c, u32* a, int64_t len)
{
    if (len <= 8192)
    {
        VC_LAUNCH(c, "scan_u32", k_rs_scan, 1, 1024, 0, a, len);
    }
    else
    {
        const unsigned sblocks = vc_blocks((size_t)len, 4096);
        VC_CUDA(c, c->scan_sums.ensure(((size_t)sblocks + 2) * 4));
        u32* sums = c->scan_sums.as<u32>();
        VC_LAUNCH(c, "scan_u32", k_scan_blocks, sblocks, 1024, 0, a, len, sums);
        VC_LAUNCH(c, "scan_u32", k_rs_scan, 1, 1024, 0, sums, (int64_t)sblocks);
        VC_LAUNCH(c, "scan_u32", k_scan_add, sblocks, 1024, 0, a, len, sums);
    }
    VC_CUDA(c, cudaGetLastError());
    return VC_OK;
}

static int bits_for(u64 maxval)
{
    int b = 1;
    while (b < 64 && (maxval >> b))
        ++b;
    return b;
}

// (first-encounter key, index) pairs of n records in key order: *keys_io / *vals_io (buffers sk0/sk1, sv0/sv1 of the
// ctx, which the caller has sized) come back pointing at the sorted keys and at the permutation
int vc_sort_records_by_key(vc_ctx* c, const u64* keys_dev, int64_t n, u64** keys_io, u32** vals_io)
{
    VC_LAUNCH(c, "sites_iota", k_iota_copy, vc_blocks((size_t)n, 256), 256, 0, keys_dev, *keys_io, *vals_io, n);
    const u64 maxkey = ((u64)c->nx * c->ny * c->nz) * 24ull;
    return vc_radix_sort_pairs(c, n, bits_for(maxkey), keys_io, vals_io);
}

// mode: VC_SITES_EXTERNAL  ids = the order given (external sample set; repeated points are detected),
//       VC_SITES_SORT      ids = rank of the first-encounter key (records in any order),
//       VC_SITES_PRESORTED the records already ARE in key order (the merged runs of the peer exchange): no sort
int st_finalize_sites(vc_ctx* c, const u64* keys_dev, const u64* corners_dev, int64_t n, int mode)
{
    const bool sort_by_key = mode == VC_SITES_SORT;
    const bool external = mode == VC_SITES_EXTERNAL;
    const int nlines = (c->nx + 1) * (c->ny + 1);
    if (n > (int64_t)VC_MAX_SITE_ID + 1)
        return vc_fail(c, VC_ERR_UNSUPPORTED, "more than 2^25 sites: ids no longer fit the packed 25-bit id fields of the dense path");
    c->nsites = n;
    c->lattice = true;
    c->edt_cols_ready = false; // the transform's compact column tables belong to the previous site set
    c->cl_dim[0] = 0; // any cell list of a previous site set is stale
    VC_CUDA(c, c->site_key.ensure((size_t)(n + 1) * 8));
    VC_CUDA(c, c->site_corner.ensure((size_t)(n + 1) * 8));
    VC_CUDA(c, c->site_xyz.ensure((size_t)(n + 1) * 16));
    VC_CUDA(c, c->line_ent.ensure((size_t)(n + 1) * 8));
    VC_CUDA(c, c->line_ptr.ensure((size_t)(nlines + 2) * 4));
    VC_CUDA(c, c->colmask.ensure((size_t)(c->ny + 1) * (size_t)((c->nx + 32) >> 5) * 4 + 16));
    VC_CUDA(c, c->sk0.ensure((size_t)(n + 1) * 8));
    VC_CUDA(c, c->sk1.ensure((size_t)(n + 1) * 8));
    VC_CUDA(c, c->sv0.ensure((size_t)(n + 1) * 4));
    VC_CUDA(c, c->sv1.ensure((size_t)(n + 1) * 4));
    if (n == 0)
    {
        VC_CUDA(c, cudaMemsetAsync(c->line_ptr.p, 0, (size_t)(nlines + 2) * 4, c->stream));
        VC_CUDA(c, cudaMemsetAsync(c->colmask.p, 0, (size_t)(c->ny + 1) * (size_t)((c->nx + 32) >> 5) * 4, c->stream));
        c->have_sites = true;
        c->have_closest = c->have_measures = false;
        return VC_OK;
    }
    unsigned blocks = vc_blocks((size_t)n, 256);
    u64* k = c->sk0.as<u64>();
    u32* v = c->sv0.as<u32>();
    const u32* order = nullptr;
    const u64* ksorted = nullptr;
    if (sort_by_key)
    {
        VC_LAUNCH(c, "sites_iota", k_iota_copy, blocks, 256, 0, keys_dev, k, v, n);
        u64 maxkey = ((u64)c->nx * c->ny * c->nz) * 24ull;
        VC_TRY(vc_radix_sort_pairs(c, n, bits_for(maxkey), &k, &v));
        order = v;
        ksorted = k;
    }
    if (mode == VC_SITES_PRESORTED)
        ksorted = keys_dev;
    // key2/val2 go to the buffers the first sort is NOT currently holding its result in
    u64* k2 = (k == c->sk0.as<u64>()) ? c->sk1.as<u64>() : c->sk0.as<u64>();
    u32* v2 = (v == c->sv0.as<u32>()) ? c->sv1.as<u32>() : c->sv0.as<u32>();
    VC_LAUNCH(c, "site_tables", k_site_tables, blocks, 256, 0, order, corners_dev, ksorted, c->site_key.as<u64>(),
              c->site_corner.as<u64>(), c->site_xyz.as<float4>(), k2, v2, n, c->ny, c->nz);
    u64* counter = c->scratch.as<u64>();
    static const bool lines_by_sort = getenv("VC_LINES") && !strcmp(getenv("VC_LINES"), "sort");
    if (lines_by_sort)
    { // the earlier formulation, kept for A/B runs: a second radix sort on (line, cz)
        u64 maxkey2 = (u64)nlines * (u64)(c->nz + 1);
        VC_TRY(vc_radix_sort_pairs(c, n, bits_for(maxkey2), &k2, &v2));
        VC_LAUNCH(c, "line_entries", k_line_entries, blocks, 256, 0, k2, v2, c->line_ent.as<u64>(), n, c->nz);
        VC_LAUNCH(c, "line_ptr", k_line_ptr, vc_blocks((size_t)nlines + 1, 256), 256, 0, k2, n, nlines, c->nz,
                  c->line_ptr.as<int>());
        if (external)
        {
            VC_CUDA(c, cudaMemsetAsync(counter, 0, 16, c->stream));
            VC_LAUNCH(c, "count_dups", k_count_dups, blocks, 256, 0, k2, n, counter);
        }
    }
    else
    {
        const int CY = c->ny + 1;
        const int64_t len = (int64_t)nlines + 1;
        u32* ptr = c->line_ptr.as<u32>();
        VC_CUDA(c, c->line_cur.ensure((size_t)(nlines + 2) * 4));
        u32* cursor = c->line_cur.as<u32>();
        // queues of the columns with more than LS_SHORT entries, by size class: at most n / (m + 1) columns hold more
        // than m entries
        const u32 qcap[4] = {(u32)(n / 5 + 1), (u32)(n / 9 + 1), (u32)(n / 17 + 1), (u32)(n / 33 + 1)};
        const LsQueues qs = {{LS_QHDR, LS_QHDR + qcap[0], LS_QHDR + qcap[0] + qcap[1], LS_QHDR + qcap[0] + qcap[1] + qcap[2]}};
        VC_CUDA(c, c->shist.ensure(((size_t)qs.start[3] + qcap[3]) * 4));
        u32* queue = c->shist.as<u32>();
        VC_CUDA(c, cudaMemsetAsync(ptr, 0, (size_t)(nlines + 2) * 4, c->stream));
        VC_CUDA(c, cudaMemsetAsync(queue, 0, LS_QHDR * 4, c->stream));
        VC_CUDA(c, cudaMemsetAsync(counter, 0, 16, c->stream));
        VC_LAUNCH(c, "line_count", k_line_count, blocks, 256, 0, c->site_corner.as<u64>(), n, CY, ptr);
        VC_TRY(vc_exclusive_scan_u32(c, ptr, len));
        VC_CUDA(c, cudaMemcpyAsync(cursor, ptr, (size_t)(nlines + 1) * 4, cudaMemcpyDeviceToDevice, c->stream));
        VC_LAUNCH(c, "line_fill", k_line_fill, blocks, 256, 0, c->site_corner.as<u64>(), n, CY, cursor, k2);
        VC_LAUNCH(c, "line_sort", k_line_sort_short, vc_blocks((size_t)nlines, 256), 256, 0, c->line_ptr.as<int>(), nlines, k2,
                  c->line_ent.as<u64>(), queue, qs, external ? counter : (u64*)nullptr);
        VC_LAUNCH(c, "line_sort", k_line_sort_queued, c->sm_count * 4, 256, 0, c->line_ptr.as<int>(), k2, c->line_ent.as<u64>(),
                  queue, qs, external ? counter : (u64*)nullptr);
    }
    {
        const int CX = c->nx + 1, CY = c->ny + 1, nw = (CX + 31) >> 5;
        VC_LAUNCH(c, "line_mask", k_line_mask, vc_blocks((size_t)CY * nw, 256), 256, 0, c->line_ptr.as<int>(), CX, CY, nw,
                  c->colmask.as<u32>());
    }
    if (external)
    { // external set: duplicates on the lattice would need the lowest-id rule inside a list entry
        u64 d = 0;
        VC_CUDA(c, cudaMemcpyAsync(&d, counter, 8, cudaMemcpyDeviceToHost, c->stream));
        VC_CUDA(c, cudaStreamSynchronize(c->stream));
        if (d)
            c->lattice = false;
    }
    VC_CUDA(c, cudaGetLastError());
    c->have_sites = true;
    c->have_closest = c->have_measures = false;
    return VC_OK;
}
