// vc_edt.cu -- stage 2 on the dense grid: the exact closest-site transform.
//
// Replaces one ANNkd_tree::annkSearch(k=1, eps=0) per grid vertex
// (3rdparty/ann/src/kd_search.cpp:88-216) with ANNbruteForce's tie rule
// (3rdparty/ann/src/brute.cpp:56-82): winner = lexicographic min of (d^2, site id).
//
// Sites lie on the corner lattice, so the minimum separates exactly over the axes (vc_core.h).  Round-2
// data flow -- every pass walks only candidates that exist, and the candidates of a warp are UNIFORM:
//   columns   the (cx,cy) z-lines that hold sites, numbered row by row (cy major, cx ascending): `col_x`,
//             `col_line`, `row_ptr[cy]`; the rows that hold any: `live_row` (k_col_rows / k_col_fill)
//   pass Z    G1c[col][vz]       = min over the sites of the column              lanes along vz
//   pass X    G2c[vz][row][vx]   = min over the row's columns  G1c + (2(vx-cx)+1)^2
//             line = (live row, vz); the 32 lanes of a warp are 32 consecutive vz of ONE row, so they all
//             walk the same column list: no dead candidates, no per-lane bitmap, coalesced G1c reads
//   pass Y    out[vz][vy][vx]    = min over the live rows  G2c + (2(vy-cy)+1)^2
//             line = (vz, vx), lanes along vx; the candidate list (the live rows) is the same for every line
// Each 1-D pass is vc_envelope_pruned's scan (Meijster-style on the total order (4d^2, id): the stack keeps
// only candidates that win at >= 1 integer target and the first target each wins at), written out below
// with the device stack in a shared-memory ring that spills to a per-line array (StackRing below).
// The planes of a slab are independent after pass Z: no exchange between GPUs (halo planes are recomputed,
// SURVEY section 8e).
#include "vc_internal.h"

// ---- compact columns ---------------------------------------------------------------------------------
// meta[0] = number of columns with sites, meta[1] = number of live rows
__global__ void __launch_bounds__(1024)
    k_col_rows(const u32* __restrict__ colmask, int CY, int nw, int* __restrict__ row_ptr, int* __restrict__ live_row,
               u32* __restrict__ row_mask, int* __restrict__ meta)
{
    __shared__ int cnt[2052], liv[2052];
    for (int cy = threadIdx.x; cy < CY; cy += blockDim.x)
    {
        int s = 0;
        for (int w = 0; w < nw; ++w)
            s += __popc(colmask[(size_t)cy * nw + w]);
        cnt[cy] = s;
        liv[cy] = s > 0;
    }
    __syncthreads();
    if (threadIdx.x < 32)
    { // one warp scans the <= 2049 rows, 32 at a time
        int carry = 0, lcarry = 0;
        for (int b = 0; b < CY; b += 32)
        {
            const int cy = b + threadIdx.x;
            const int v = cy < CY ? cnt[cy] : 0, l = cy < CY ? liv[cy] : 0;
            int inc = v, linc = l;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1)
            {
                const int t = __shfl_up_sync(0xffffffffu, inc, o), u = __shfl_up_sync(0xffffffffu, linc, o);
                if ((int)threadIdx.x >= o)
                {
                    inc += t;
                    linc += u;
                }
            }
            const u32 lw = __ballot_sync(0xffffffffu, l != 0); // bit cy & 31 of word cy >> 5: row cy is live
            if (threadIdx.x == 0)
                row_mask[b >> 5] = lw;
            if (cy < CY)
            {
                row_ptr[cy] = carry + inc - v;
                if (l)
                    live_row[lcarry + linc - 1] = cy;
            }
            carry += __shfl_sync(0xffffffffu, inc, 31);
            lcarry += __shfl_sync(0xffffffffu, linc, 31);
        }
        if (threadIdx.x == 0)
        {
            row_ptr[CY] = carry;
            meta[0] = carry;
            meta[1] = lcarry;
            row_mask[(CY + 31) >> 5] = 0u; // the word the bitmap walk may read ahead
        }
    }
}

__global__ void k_col_fill(const u32* __restrict__ colmask, const int* __restrict__ row_ptr, int CY, int nw,
                           int* __restrict__ col_x, int* __restrict__ col_line)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= CY * nw)
        return;
    const int cy = i / nw, w = i - cy * nw;
    u32 word = colmask[i];
    if (!word)
        return;
    int o = row_ptr[cy];
    for (int k = 0; k < w; ++k)
        o += __popc(colmask[(size_t)cy * nw + k]);
    while (word)
    {
        const int b = __ffs(word) - 1;
        word &= word - 1;
        const int cx = 32 * w + b;
        col_x[o] = cx;
        col_line[o] = cx * CY + cy; // index into line_ptr (vc_sites.cu: columns are numbered cx * CY + cy)
        ++o;
    }
}

// ---- pass Z --------------------------------------------------------------------------------------------
#ifndef PZ_BLOCKS_PER_SM
#define PZ_BLOCKS_PER_SM 16 // 128-thread blocks per SM of the (grid-stride) launch: 64 resident warps hide the list searches (0.168 -> 0.154 ms on a 127-plane slab of 1024^2 against 8)
#endif
// One warp per column and 32 planes at a time: lane = plane.  The column's sorted list (cz << 32 | id) is
// searched per lane (binary search: a rod's column holds hundreds of sites, most columns < 8).
__global__ void __launch_bounds__(128)
    k_pass_z(const int* __restrict__ line_ptr, const u64* __restrict__ ent, const int* __restrict__ col_line,
             const int* __restrict__ meta, u64* __restrict__ G1c, int zb, int pz)
{
    const int ncol = meta[0];
    const int lane = threadIdx.x & 31;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    for (int ci = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; ci < ncol; ci += nwarps)
    {
        const int l = col_line[ci];
        const int first = line_ptr[l], last = line_ptr[l + 1];
        u64* out = G1c + (size_t)ci * pz;
        for (int v = lane; v < pz; v += 32)
        {
            const int vz = zb + v;
            int lo = first, hi = last; // first entry with cz > vz
            while (lo < hi)
            {
                const int mid = (lo + hi) >> 1;
                if ((int)(ent[mid] >> 32) <= vz)
                    lo = mid + 1;
                else
                    hi = mid;
            }
            __stcs(out + v, vc_nearest_on_zline(ent, first, last, lo - 1, vz));
        }
    }
}

// ---- shared-memory helpers: explicit 32-bit shared addresses, one instruction per access -------------------
__device__ __forceinline__ uint4 lds128(unsigned a)
{
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void sts128(unsigned a, const uint4& v)
{
    asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ u64 lds64(unsigned a)
{
    u64 v;
    asm volatile("ld.shared.u64 %0, [%1];" : "=l"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void cp_async16(unsigned dst, const void* src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async8(unsigned dst, const void* src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

#ifndef XY_THREADS
#define XY_THREADS 128
#endif

// ---- envelope stack storage on the device ----------------------------------------------------------
// Entries are 16 bytes (g, id, p, start): one LDS.128 / STS.128, nothing to pack.  The top of the stack lives
// in registers; of the entries below it the upper SR_R sit in a per-thread ring in shared memory (slot =
// depth mod SR_R, threads interleaved: conflict free whatever the depths), what falls out at the bottom is
// spilled to global memory.
//   forward scan : push = one STS (+ one STG when the ring is full); a pop reads the ring (LDS).  A pop below the
//                  ring (a long run of pops: the scan has just passed a much closer surface) refills the WHOLE
//                  ring with the SR_R entries below in one go -- SR_R independent copies, one round trip to L2 /
//                  HBM per SR_R pops instead of one per pop.
//   spill layout : the stores of the spill are what bounds these kernels (measured: without them pass Y takes
//                  0.6 instead of 1.5 ms), so a warp's 32 stacks are INTERLEAVED, [depth][lane] x 16 B: lanes at
//                  the same depth -- neighbouring lines have nearly the same envelope -- share full 128-byte
//                  lines, one store instruction then costs 4 wavefronts / 16 full sectors instead of 32 wavefronts
//                  / 32 half-written sectors of 32 private arrays (and never more than those).
//   backward scan: the stack is drained strictly downwards, so the slot a pop frees is refilled at once with
//                  the entry SR_R below it by an asynchronous copy (cp.async 16 B, global -> shared).  Each
//                  drain commits exactly one copy group, so `wait_group SR_R-1` before reading a slot is
//                  precisely "the copy issued SR_R drains ago has landed": a pop never waits for HBM.
#ifndef SR_RX
#define SR_RX 8 // ring slots per thread, pass X / pass Y
#endif
#ifndef SR_RY
#define SR_RY 8
#endif
template <int SR_R>
struct StackRing
{
    unsigned sbase; // shared address of this thread's slot 0; slot i at sbase + i * XY_THREADS * 16
    uint4* glob;    // this lane's column of the warp's spill block: depth d at glob[d * 32]
    int lo;         // depths >= lo are in the ring, depths < lo only in global memory

    static __device__ __forceinline__ uint4 pack(const vc_ent& e) { return make_uint4((u32)e.g, e.id, (u32)e.p, (u32)e.start); }
    static __device__ __forceinline__ vc_ent unpack(const uint4& v)
    {
        vc_ent e;
        e.g = (int)v.x;
        e.id = v.y;
        e.p = (int)v.z;
        e.start = (int)v.w;
        return e;
    }
    __device__ __forceinline__ unsigned slot(int d) const { return sbase + (unsigned)(d & (SR_R - 1)) * (XY_THREADS * 16); }
    __device__ __forceinline__ void store(int d, const vc_ent& e)
    {
        if (d - lo >= SR_R)
        { // ring full: depth lo lives in the slot depth d is about to take
#if !defined(WHATIF_NOSPILL) // timing experiments only (results are wrong with these switches)
            glob[lo * 32] = lds128(slot(lo));
#endif
            ++lo;
        }
        sts128(slot(d), pack(e));
    }
    __device__ __forceinline__ vc_ent load(int d)
    {
        if (d < lo)
        { // below the ring: everything above d is popped, so the ring is free -- fetch depths d-SR_R+1 .. d together
            const int first = d - SR_R + 1 > 0 ? d - SR_R + 1 : 0;
#if !defined(WHATIF_NOREFILL)
            for (int e = first; e <= d; ++e)
                cp_async16(slot(e), glob + e * 32);
            cp_commit();
            cp_wait<0>();
#endif
            lo = first;
        }
        return unpack(lds128(slot(d)));
    }
    __device__ __forceinline__ void begin_drain(int dtop)
    {
        const int first = dtop - SR_R + 1;
        for (int d = lo - 1; d >= 0 && d >= first; --d)
            cp_async16(slot(d), glob + d * 32);
        cp_commit();
        cp_wait<0>();
    }
    __device__ __forceinline__ vc_ent drain(int d)
    {
        cp_wait<SR_R - 1>();
        const uint4 e = lds128(slot(d));
        if (d - SR_R >= 0)
            cp_async16(slot(d), glob + (d - SR_R) * 32);
        cp_commit();
        return unpack(e);
    }
};

// ---- candidate stream ----------------------------------------------------------------------------------
// Words: ED_DEPTH candidates ahead in registers, ED_PFD candidates ahead in L2 (prefetch.global.L2: no register, no
// shared memory -- the shared-memory pipe is what bounds these kernels, so the fetch stays out of it).
// Positions: the candidates of a warp are the set bits of a bitmap that is the same for every lane (pass X: the
// row's columns, `colmask`; pass Y: the live rows, `row_mask`), walked with ffs -- no load per candidate at all.
#ifndef ED_PFD
#define ED_PFD 8 // L2 prefetch distance in candidates (0: none)
#endif
#ifndef ED_DEPTH
#define ED_DEPTH 4 // candidates in flight in registers
#endif
struct CandStream
{
    const u64 *gp, *gpf; // word of the next candidate to load / to prefetch
    long stride;
    u64 h[ED_DEPTH];
    const u32* mw; // next bitmap word
    u32 word, wnext;
    int base, ncand;
    __device__ __forceinline__ void init(const u64* src, long stride_, const u32* mask, int ncand_)
    {
        stride = stride_;
        ncand = ncand_;
        gp = src;
#pragma unroll
        for (int i = 0; i < ED_DEPTH; ++i, gp += stride)
            h[i] = i < ncand ? VC_LOAD_STREAM(gp) : 0ull;
        gpf = gp;
        if (ED_PFD > 0)
            for (int i = ED_DEPTH; i < ED_PFD && i < ncand; ++i, gpf += stride)
                VC_PREFETCH_L2(gpf);
        word = 0;
        base = -32;
        wnext = ncand > 0 ? __ldg(mask) : 0u;
        mw = mask + 1;
    }
    __device__ __forceinline__ void get(int k, u64& H, int& p)
    {
        H = h[0];
#pragma unroll
        for (int i = 0; i + 1 < ED_DEPTH; ++i)
            h[i] = h[i + 1];
        if (k + ED_DEPTH < ncand)
            h[ED_DEPTH - 1] = VC_LOAD_STREAM(gp);
        gp += stride;
        if (ED_PFD > 0)
        {
            if (k + ED_PFD < ncand)
                VC_PREFETCH_L2(gpf);
            gpf += stride;
        }
        while (word == 0u)
        { // k < ncand: there is a set bit ahead, so the walk never leaves the bitmap (one word of slack is allocated)
            word = wnext;
            wnext = __ldg(mw++);
            base += 32;
        }
        p = base + __ffs(word) - 1;
        word &= word - 1;
    }
};

// ---- passes X and Y ------------------------------------------------------------------------------------
#ifndef XY_TW
#define XY_TW 8 // pass X: targets per transposed store burst (rows of XY_TW * 8 = 64 bytes = two full sectors)
#endif
#ifndef XY_MINB
#define XY_MINB 8 // resident blocks per SM the register allocation must allow
#endif
#define ST_STRIDE(c) ((size_t)((c)->nx > (c)->ny ? (c)->nx : (c)->ny) + 2) // spill entries per line

// pass X: warp = (live row r, group of 32 planes); lane = plane.  Outputs are staged per warp in shared
// memory and written as rows of 128 bytes of G2c[plane][r][vx].
__global__ void __launch_bounds__(XY_THREADS, XY_MINB)
    k_pass_x(const u64* __restrict__ G1c, u64* __restrict__ G2c, uint4* __restrict__ spill, const int* __restrict__ row_ptr,
             const int* __restrict__ live_row, const u32* __restrict__ colmask, int nw, const int* __restrict__ meta, int pz,
             int ngroups, int nx, size_t spill_stride)
{
    __shared__ u64 tile[XY_THREADS / 32][32][XY_TW + 1];
    __shared__ uint4 rings[SR_RX * XY_THREADS];
    const int nlive = meta[1];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int w = blockIdx.x * (XY_THREADS / 32) + warp;
    const int r = w / ngroups, grp = w - r * ngroups;
    if (r >= nlive)
        return; // warp-uniform
    const int cy = live_row[r], c0 = row_ptr[cy], nc = row_ptr[cy + 1] - c0;
    const int v = grp * 32 + lane; // plane of this lane inside the chunk
    const bool valid = v < pz;
    StackRing<SR_RX> stk;
    stk.sbase = (unsigned)__cvta_generic_to_shared(rings + threadIdx.x);
    stk.glob = spill + (size_t)w * 32 * spill_stride + lane; // the warp's block of 32 interleaved stacks
    stk.lo = 0;
    CandStream src;
    src.init(G1c + (size_t)c0 * pz + (valid ? v : 0), (long)pz, colmask + (size_t)cy * nw, valid ? nc : 0);
    const int rsub = lane / XY_TW, col = lane % XY_TW; // a store instruction covers 32 / XY_TW rows
    const int v0 = grp * 32;
    vc_envelope_pruned(src, valid ? nc : 0, nx, stk,
                       [&](int t, u32 V, u32 id)
                       {
                           tile[warp][lane][t % XY_TW] = ((u64)V << 32) | id;
                           if ((t % XY_TW) == 0)
                           {
                               __syncwarp();
                               const bool colok = t + col < nx;
                               u64* dst = G2c + ((size_t)(v0 + rsub) * nlive + r) * nx + t + col;
#pragma unroll 4
                               for (int rr = rsub; rr < 32; rr += 32 / XY_TW, dst += (size_t)(32 / XY_TW) * nlive * nx)
                                   if (colok && v0 + rr < pz)
                                       *dst = tile[warp][rr][col];
                               __syncwarp();
                           }
                       });
}

// pass Y: line = (plane, vx), lanes along vx; candidates = the live rows; writes id / 4d^2 coalesced.
__global__ void __launch_bounds__(XY_THREADS, XY_MINB)
    k_pass_y(const u64* __restrict__ G2c, int* __restrict__ id_out, u32* __restrict__ d2_out, uint4* __restrict__ spill,
             const u32* __restrict__ row_mask, const int* __restrict__ meta, long nlines, int nx, int ny, size_t spill_stride)
{
    __shared__ uint4 rings[SR_RY * XY_THREADS];
    const int nlive = meta[1];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long g = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = g < nlines;
    const long plane = valid ? g / nx : 0;
    const int vx = valid ? (int)(g - plane * nx) : 0;
    StackRing<SR_RY> stk;
    stk.sbase = (unsigned)__cvta_generic_to_shared(rings + threadIdx.x);
    stk.glob = spill + (size_t)(g - lane) * spill_stride + lane; // the warp's block of 32 interleaved stacks
    stk.lo = 0;
    CandStream src;
    src.init(G2c + (size_t)plane * nlive * nx + vx, (long)nx, row_mask, valid ? nlive : 0);
    const long last = (plane * ny + (ny - 1)) * (long)nx + vx;
    int* pid = id_out + last;
    u32* pd2 = d2_out + last;
    vc_envelope_pruned(src, valid ? nlive : 0, ny, stk,
                       [&](int t, u32 V, u32 id)
                       { // targets arrive as ny-1 .. 0
                           if (valid)
                           {
                               __stcs(pid, (int)id);
                               __stcs(pd2, V);
                           }
                           pid -= nx;
                           pd2 -= nx;
                       });
}

// scratch for the envelope stacks: one region per z chunk of the pipeline (chunks run concurrently), a region
// holds planes x lines x ST_STRIDE 16-byte entries (touched only as deep as a stack outgrows its ring)
static size_t stack_region_entries(const vc_ctx* c, int planes)
{
    const size_t CY = c->ny + 1;
    const size_t lx = CY * (size_t)((planes + 31) / 32 * 32);                          // pass X: <= CY live rows x 32-plane groups
    const size_t ly = ((size_t)planes * c->nx + XY_THREADS - 1) / XY_THREADS * XY_THREADS; // pass Y: whole blocks of lines
    return (lx > ly ? lx : ly) * ST_STRIDE(c);
}

// compact column tables of the current site set (once per site set: st_finalize_sites clears edt_cols_ready)
static int edt_columns(vc_ctx* c)
{
    const int CX = c->nx + 1, CY = c->ny + 1, nw = (CX + 31) >> 5;
    const size_t ncolcap = (size_t)CX * CY;
    if (c->edt_cols_ready)
        return VC_OK;
    VC_CUDA(c, c->row_ptr.ensure((size_t)(CY + 2) * 4));
    VC_CUDA(c, c->live_row.ensure((size_t)(CY + 2) * 4));
    VC_CUDA(c, c->edt_meta.ensure(64));
    const size_t cap = (size_t)c->nsites < ncolcap ? (size_t)c->nsites : ncolcap;
    VC_CUDA(c, c->col_x.ensure((cap + 1) * 4));
    VC_CUDA(c, c->col_line.ensure((cap + 1) * 4));
    VC_CUDA(c, c->row_mask.ensure((size_t)(((CY + 31) >> 5) + 2) * 4));
    VC_LAUNCH(c, "edt_columns", k_col_rows, 1, 1024, 0, c->colmask.as<u32>(), CY, nw, c->row_ptr.as<int>(), c->live_row.as<int>(),
              c->row_mask.as<u32>(), c->edt_meta.as<int>());
    VC_LAUNCH(c, "edt_columns", k_col_fill, vc_blocks((size_t)CY * nw, 256), 256, 0, c->colmask.as<u32>(), c->row_ptr.as<int>(), CY,
              nw, c->col_x.as<int>(), c->col_line.as<int>());
    c->edt_cols_ready = true;
    return VC_OK;
}

// buffers of the transform; `regions` stack regions of `region_planes` planes each
static int edt_alloc(vc_ctx* c, int regions, int region_planes)
{
    if (!c->have_sites)
        return vc_fail(c, VC_ERR_STATE, "vc_closest_grid needs sites (vc_extract_sites / vc_set_sites)");
    const int CX = c->nx + 1, CY = c->ny + 1;
    const int nplanes = c->zc - c->z0;
    const size_t nv = (size_t)c->nx * c->ny * nplanes;
    if ((CX > CY ? CX : CY) > 2049)
        return vc_fail(c, VC_ERR_UNSUPPORTED, "grid side above 2048 is not supported by the dense transform");
    // G1c of a chunk: [columns][planes] <= CX*CY columns; G2c: [planes][live rows][nx] <= CY rows: the dense bounds
    VC_CUDA(c, c->g1.ensure((size_t)nplanes * CX * CY * 8));
    VC_CUDA(c, c->g2.ensure((size_t)nplanes * CY * c->nx * 8));
    VC_CUDA(c, c->id.ensure(nv * 4));
    VC_CUDA(c, c->d2.ensure(nv * 4));
    VC_CUDA(c, c->stk.ensure((size_t)regions * stack_region_entries(c, region_planes) * sizeof(uint4)));
    return VC_OK;
}

// the three passes for the planes [zb, ze) of the slab, on stream c->cur (buffers from edt_alloc)
int edt_range(vc_ctx* c, int zb, int ze, int region, int region_planes)
{
    const int CX = c->nx + 1, CY = c->ny + 1;
    const int pz = ze - zb;
    if (pz <= 0)
        return VC_OK;
    const size_t off = (size_t)(zb - c->z0);
    u64* g1 = c->g1.as<u64>() + off * CX * CY;
    u64* g2 = c->g2.as<u64>() + off * CY * c->nx;
    uint4* spill = c->stk.as<uint4>() + (size_t)region * stack_region_entries(c, region_planes);
    const int* meta = c->edt_meta.as<int>();
    if (c->nsites > 0)
    {
        VC_LAUNCH(c, "edt_pass_z", k_pass_z, c->sm_count * PZ_BLOCKS_PER_SM, 128, 0, c->line_ptr.as<int>(), c->line_ent.as<u64>(),
                  c->col_line.as<int>(), meta, g1, zb, pz);
        const int ngroups = (pz + 31) / 32;
        const size_t warps = (size_t)CY * ngroups; // warps of rows that are not live leave at once
        VC_LAUNCH(c, "edt_pass_x", k_pass_x, vc_blocks(warps, XY_THREADS / 32), XY_THREADS, 0, g1, g2, spill, c->row_ptr.as<int>(),
                  c->live_row.as<int>(), c->colmask.as<u32>(), (CX + 31) >> 5, meta, pz, ngroups, c->nx, ST_STRIDE(c));
    }
    const long nlines = (long)pz * c->nx;
    VC_LAUNCH(c, "edt_pass_y", k_pass_y, vc_blocks((size_t)nlines, XY_THREADS), XY_THREADS, 0, g2,
              c->id.as<int>() + off * c->nx * c->ny, c->d2.as<u32>() + off * c->nx * c->ny, spill, c->row_mask.as<u32>(), meta, nlines,
              c->nx, c->ny, ST_STRIDE(c));
    return VC_OK;
}

int st_closest_lattice(vc_ctx* c)
{
    VC_TRY(edt_alloc(c, 1, c->zc - c->z0));
    VC_TRY(edt_columns(c));
    VC_TRY(edt_range(c, c->z0, c->zc, 0, c->zc - c->z0));
    VC_CUDA(c, cudaGetLastError());
    c->have_closest = true;
    c->have_measures = false;
    return VC_OK;
}

// Closest sites + measures of the whole slab as a pipeline over z chunks.  The transform ranges [zb, ze) of
// the chunks partition the closest planes [z0, zc); the measures of chunk k cover the planes [zb - 1, ze - 1)
// (from z0 for the first chunk, up to z1 for the last): the cells of plane z read the ids of plane z + 1, so a
// chunk's measures need its own transform and the LAST plane of the chunk before it -- one event wait across
// worker streams, no plane is transformed twice and no two streams ever write the same address.
// Chunks are dealt round-robin to the worker streams, so there is no device-wide barrier between the passes:
// the tail of one chunk's stage overlaps whatever the other streams are running.
// With profiling on, everything runs on the main stream so per-kernel event times stay meaningful.
int st_closest_measures_pipelined(vc_ctx* c, bool want_radius)
{
    if (!c->lattice)
    {
        VC_TRY(st_closest_general_grid(c));
        return st_measures(c, want_radius);
    }
    const int nplanes_all = c->zc - c->z0;
    int nw = c->profiling ? 0 : c->nworkers;
    // chunk height: a multiple of 32 planes (pass X puts 32 planes in a warp), >= 64 planes and >= 32k lines per
    // launch so that a chunk's kernels are efficient on their own, and no more chunks than about one per worker stream;
    // VC_ZCHUNK overrides.  Fewer than three chunks do not pay (measured: two are slower than one -- a 129-plane slab of
    // a 1024^2 grid runs 4.11 ms as one chunk, 4.30 as four, 4.40 as two), so a thin slab is one chunk.  (Chunks of a
    // thin slab started one stage apart, so that different stages overlap, are worse still -- 3.8 ms for two, 5.9 for four
    // against 3.05: a pass over a quarter of the planes takes about as long as over all of them, its duration is the
    // sequential scan of the longest lines, not the number of lines.)
    // When every chunk's records are copied to the host as it completes (chunk_hook: vc_run_dense_host_compact), the
    // copies should start early and run throughout: short chunks on few streams, so that chunks complete one after
    // another instead of all at the end (assembly1024: 1.33 GB of records; 111.8 ms per call with 8 chunks on 8 streams,
    // 103.3 ms with 16 chunks on 4 -- the copy-back is then hidden behind the transform except for its last chunk).
    const bool copy_back = c->chunk_hook != nullptr;
    if (copy_back && nw > 4 && c->zchunk <= 0)
        nw = 4;
    int zchunk = c->zchunk;
    if (zchunk <= 0)
    {
        const int side = c->nx < c->ny ? c->nx : c->ny;
        zchunk = (32768 + side - 1) / side;
        zchunk = ((zchunk < 64 ? 64 : zchunk) + 31) / 32 * 32;
        const int per_worker = copy_back ? 0 : ((nplanes_all + (nw > 0 ? nw : 1) - 1) / (nw > 0 ? nw : 1) + 31) / 32 * 32;
        zchunk = zchunk < per_worker ? per_worker : zchunk;
        if (nplanes_all / zchunk < 3)
            zchunk = nplanes_all;
    }
    if (!nw || zchunk > nplanes_all)
        zchunk = nplanes_all;
    // the last chunk takes the remainder (a slab's halo plane, for one) instead of becoming a launch of its own
    const int nchunks = nplanes_all / zchunk > 1 ? nplanes_all / zchunk : 1;
    if (nchunks == 1)
        nw = 0; // nothing to overlap: the fork / join events around a single chunk cost 0.06 ms
    const int region_planes = nplanes_all - (nchunks - 1) * zchunk;
    VC_TRY(edt_alloc(c, nchunks, region_planes));
    if (!c->have_inside)
        return vc_fail(c, VC_ERR_STATE, "measures need vc_classify_grid");
    if (c->zhi < c->zc)
        return vc_fail(c, VC_ERR_STATE, "inside flags do not cover the halo plane");
    const bool dense_measures = !c->skip_dense_measures;
    if (dense_measures)
        VC_TRY(measures_alloc(c, want_radius));
    VC_TRY(edt_columns(c)); // on the main stream, before the fork
    while ((int)c->ev_chunk.size() < nchunks)
    {
        cudaEvent_t e = nullptr;
        VC_CUDA(c, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        c->ev_chunk.push_back(e);
    }
    if (nw)
    {
        VC_CUDA(c, cudaEventRecord(c->ev_fork, c->stream));
        for (int i = 0; i < nw; ++i)
            VC_CUDA(c, cudaStreamWaitEvent(c->workers[i], c->ev_fork, 0));
    }
    int status = VC_OK;
    for (int k = 0; k < nchunks && status == VC_OK; ++k)
    {
        const int zb = c->z0 + k * zchunk;
        const int ze = k == nchunks - 1 ? c->zc : zb + zchunk;
        c->cur = nw ? c->workers[k % nw] : c->stream;
        status = edt_range(c, zb, ze, k, region_planes);
        if (status != VC_OK)
            break;
        if (nw)
        {
            VC_CUDA(c, cudaEventRecord(c->ev_chunk[k], c->cur));
            if (k > 0)
                VC_CUDA(c, cudaStreamWaitEvent(c->cur, c->ev_chunk[k - 1], 0));
        }
        const int ma = k == 0 ? c->z0 : zb - 1;
        const int mb = ze >= c->zc ? c->z1 : ze - 1;
        if (ma < mb && dense_measures)
            status = measures_range(c, ma, mb, want_radius, nchunks == 1);
        if (status == VC_OK && ma < mb && c->chunk_hook)
            status = c->chunk_hook(ma, mb); // e.g. compaction + device-to-host copy of this chunk's records
    }
    c->cur = c->stream;
    if (nw)
        for (int i = 0; i < nw; ++i)
        {
            VC_CUDA(c, cudaEventRecord(c->ev_join[i], c->workers[i]));
            VC_CUDA(c, cudaStreamWaitEvent(c->stream, c->ev_join[i], 0));
        }
    if (status != VC_OK)
        return status;
    VC_CUDA(c, cudaGetLastError());
    c->have_closest = true;
    c->have_measures = dense_measures;
    return VC_OK;
}
