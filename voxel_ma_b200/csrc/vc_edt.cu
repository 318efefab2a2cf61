// vc_edt.cu -- stage 2 on the dense grid: the exact closest-site transform.
//
// Replaces one ANNkd_tree::annkSearch(k=1, eps=0) per grid vertex
// (3rdparty/ann/src/kd_search.cpp:88-216) with ANNbruteForce's tie rule
// (3rdparty/ann/src/brute.cpp:56-82): winner = lexicographic min of (d^2, site id).
//
// Sites lie on the corner lattice, so the minimum separates exactly over the axes (vc_core.h):
//   pass Z  (sparse -> dense)  G1[vz][cx][cy] = min over the sites of z-line (cx,cy)
//   pass X  G2[vz][cy][vx]     = min_cx  G1[vz][cx][cy] + (2(vx-cx)+1)^2      lanes along cy
//   pass Y  out[vz][vy][vx]    = min_cy  G2[vz][cy][vx] + (2(vy-cy)+1)^2      lanes along vx
// Each 1-D pass is a lower envelope of equal-width parabolas (Meijster-style two scans) carried out
// on 64-bit (4d^2<<32 | id) words, so ties in distance resolve to the lowest id inside the pass.
// One thread owns one line; the 32 lanes of a warp own 32 lines adjacent in the fastest-varying
// index of the layout, so every global access of the scans is a coalesced 256-byte row.
// The slab is independent per vz plane after pass Z: no exchange between GPUs (halo planes are
// recomputed).
#include "vc_internal.h"

// ---- pass Z -----------------------------------------------------------------------------------
// One thread walks one z-line's site list over a chunk of PZ_CHUNK planes (blockIdx.y = chunk), so
// the grid has (lines x chunks) threads instead of one per line; lanes are adjacent lines (cy
// fastest), every store is a coalesced 256-byte row of G1.
#define PZ_CHUNK 32
__global__ void __launch_bounds__(256)
    k_pass_z(const int* __restrict__ line_ptr, const u64* __restrict__ ent, u64* __restrict__ G1, int nlines, int z0,
             int zc)
{
    int l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= nlines)
        return;
    const int zb = z0 + blockIdx.y * PZ_CHUNK; // G1 points at plane z0 of this launch
    const int ze = min(zb + PZ_CHUNK, zc);
    const int first = line_ptr[l], last = line_ptr[l + 1];
    const size_t plane = (size_t)nlines;
    u64* out = G1 + l + plane * (size_t)(zb - z0);
    if (first == last)
        return; // a column without sites: pass X knows from the column mask and never reads these entries
    int nxt = first; // index of the first entry with cz > vz
    u64 below = VC_INF, above = ent[first];
    for (int vz = zb; vz < ze; ++vz, out += plane)
    {
        while (nxt < last && (int)(above >> 32) <= vz)
        {
            below = above;
            ++nxt;
            above = nxt < last ? ent[nxt] : (u64)VC_INF;
        }
        u64 H = VC_INF;
        if (below != VC_INF)
        {
            int d = 2 * (vz - (int)(below >> 32)) + 1;
            H = ((u64)(u32)(d * d) << 32) | (u32)below;
        }
        if (nxt < last)
        {
            int d = 2 * ((int)(above >> 32) - vz) - 1;
            u64 H2 = ((u64)(u32)(d * d) << 32) | (u32)above;
            H = H2 < H ? H2 : H;
        }
        __stcs(out, H);
    }
}

// ---- envelope stack storage on the device ----------------------------------------------------------
// The entries of a line's stack below the two in registers.  The upper SR_R of them sit in a
// per-thread ring in shared memory (slot = depth mod SR_R, threads interleaved so a warp's accesses
// are conflict free); what falls out at the bottom is spilled to the line's own contiguous array in
// global memory (8-byte packed entries: consecutive depths of a thread share a 32-byte sector).
//   forward scan : push = one STS (+ one fire-and-forget STG when the ring is full); a pop reads the
//                  ring (LDS) and touches global memory only on underflow.
//   backward scan: the stack is drained strictly downwards and nothing is pushed any more, so the
//                  slot a pop frees is refilled at once with the entry SR_R below it by an
//                  asynchronous copy (cp.async, global -> shared, no register in between).  Each drain
//                  commits exactly one copy group, so `wait_group SR_R-1` before reading a slot is
//                  precisely "the copy issued SR_R drains ago has landed": a pop never waits for HBM.
#ifndef SR_R
#define SR_R 8
#endif
struct StackRing
{
    u64* ring; // this thread's slot 0; slot i at ring[i * nthr]
    u64* glob; // this line's spill array
    int nthr;
    int lo;    // depths >= lo are in the ring, depths < lo only in global memory

    __device__ __forceinline__ u64& slot(int d) { return ring[(d & (SR_R - 1)) * nthr]; }
    __device__ __forceinline__ void store(int d, u64 e)
    {
        if (d - lo >= SR_R)
        { // ring full: depth lo lives in the slot depth d is about to take
            glob[lo] = slot(lo);
            ++lo;
        }
        slot(d) = e;
    }
    __device__ __forceinline__ u64 load(int d)
    {
        if (d >= lo)
            return slot(d);
        lo = d; // underflow: the ring is empty from here on
        return glob[d];
    }
    __device__ __forceinline__ void copy_in(int d)
    {
        unsigned dst = (unsigned)__cvta_generic_to_shared(&slot(d));
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(glob + d) : "memory");
    }
    __device__ __forceinline__ void begin_drain(int dtop)
    {
        int first = dtop - SR_R + 1;
        for (int d = lo - 1; d >= 0 && d >= first; --d)
            copy_in(d);
        asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
    }
    __device__ __forceinline__ u64 drain(int d)
    {
        asm volatile("cp.async.wait_group %0;" ::"n"(SR_R - 1) : "memory");
        u64 e = slot(d);
        if (d - SR_R >= 0)
            copy_in(d - SR_R);
        asm volatile("cp.async.commit_group;" ::: "memory");
        return e;
    }
};

// ---- passes X and Y ------------------------------------------------------------------------------
// TRANSPOSE = true  (pass X): line g = (vz, cy); input G1 + vz*CX*CY + cy, stride CY; the outputs of
//   32 lines x XY_TW targets are staged in shared memory and written as rows of XY_TW*8 = 128 bytes
//   of G2[g][vx].
// TRANSPOSE = false (pass Y): line g = (vz, vx); input G2 + vz*CY*nx + vx, stride nx; outputs go
//   straight to id/d2x4[(vz*ny + vy)*nx + vx], coalesced across the warp.
// Both are capped at 64 registers so that 32 warps are resident per SM.
#define XY_THREADS_T 128 // pass X: 4 warps x 4.25 KB of transpose tile
#define XY_THREADS_D 256 // pass Y
#define XY_TW 16         // targets per transposed store burst
#ifndef XY_MINB_T
#define XY_MINB_T 8 // resident blocks per SM the register allocation must allow (pass X / pass Y)
#define XY_MINB_D 4
#endif
template <bool TRANSPOSE>
__global__ void __launch_bounds__(TRANSPOSE ? XY_THREADS_T : XY_THREADS_D, TRANSPOSE ? XY_MINB_T : XY_MINB_D)
    k_pass_xy(const u64* __restrict__ in, u64* __restrict__ G2, int* __restrict__ id_out, u32* __restrict__ d2_out,
              u64* __restrict__ stack, long nlines_total, int lines_per_plane, long in_plane_stride, long in_stride,
              int ncand, int ntgt, const u32* __restrict__ colmask)
{
    constexpr int NTHR = TRANSPOSE ? XY_THREADS_T : XY_THREADS_D;
    __shared__ u64 tile[TRANSPOSE ? XY_THREADS_T / 32 : 1][TRANSPOSE ? 32 : 1][TRANSPOSE ? XY_TW + 1 : 1];
    __shared__ u64 rings[SR_R * NTHR];
    const long g = (long)blockIdx.x * blockDim.x + threadIdx.x;
    StackRing stk;
    stk.ring = rings + threadIdx.x;
    stk.glob = stack + (size_t)g * (size_t)(ncand + 1); // this line's own contiguous spill array
    stk.nthr = NTHR;
    stk.lo = 0;
    const bool valid = g < nlines_total;
    const long plane = valid ? g / lines_per_plane : 0;
    const int within = valid ? (int)(g - plane * lines_per_plane) : 0;
    const u64* src = in + plane * in_plane_stride + within;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long gwarp = g - lane;

    if (TRANSPOSE)
    {
        const int rsub = lane / XY_TW, col = lane % XY_TW; // a store instruction covers 32/XY_TW rows
        vc_envelope_line(src, in_stride, valid ? ncand : 0, ntgt, stk,
                         [&](int t, u32 V, u32 id)
                         {
                             tile[warp][lane][t % XY_TW] = ((u64)V << 32) | id;
                             if ((t % XY_TW) == 0)
                             {
                                 __syncwarp();
                                 const bool colok = t + col < ntgt;
                                 u64* dst = G2 + (gwarp + rsub) * (long)ntgt + t + col;
#pragma unroll 4
                                 for (int r = rsub; r < 32; r += 32 / XY_TW, dst += (long)(32 / XY_TW) * ntgt)
                                     if (colok && gwarp + r < nlines_total)
                                         *dst = tile[warp][r][col];
                                 __syncwarp();
                             }
                         },
                         colmask ? colmask + (size_t)within * (size_t)((ncand + 31) >> 5) : nullptr);
    }
    else
    {
        const long last = plane * (long)ntgt * lines_per_plane + within + (long)(ntgt - 1) * lines_per_plane;
        int* pid = id_out + last;
        u32* pd2 = d2_out + last;
        vc_envelope_line(src, in_stride, valid ? ncand : 0, ntgt, stk,
                         [&](int t, u32 V, u32 id)
                         { // targets arrive as ntgt-1 .. 0
                             if (valid)
                             {
                                 __stcs(pid, (int)id);
                                 __stcs(pd2, V);
                             }
                             pid -= lines_per_plane;
                             pd2 -= lines_per_plane;
                         });
    }
}

// scratch for the envelope stacks: one region per z chunk of the pipeline (chunks run concurrently),
// a region holds (planes + halo) x lines x (candidates + 1) packed 8-byte entries
static size_t stack_region_entries(const vc_ctx* c, int planes)
{
    const size_t CX = c->nx + 1, CY = c->ny + 1;
    const size_t a = CY * (CX + 1), b = (size_t)c->nx * (CY + 1); // pass X, pass Y entries per plane
    return (size_t)planes * (a > b ? a : b);
}

static void launch_passes(vc_ctx* c, int zb, int nplanes, u64* stack)
{
    const int CX = c->nx + 1, CY = c->ny + 1;
    const size_t off = (size_t)(zb - c->z0);
    u64* g1 = c->g1.as<u64>() + off * CX * CY;
    u64* g2 = c->g2.as<u64>() + off * CY * c->nx;
    // pass X: lines (vz, cy)
    {
        long nlines = (long)nplanes * CY;
        VC_LAUNCH(c, "edt_pass_x", k_pass_xy<true>, vc_blocks((size_t)nlines, XY_THREADS_T), XY_THREADS_T, 0, g1, g2,
                  (int*)nullptr, (u32*)nullptr, stack, nlines, CY, (long)CX * CY, (long)CY, CX, c->nx,
                  c->colmask.as<u32>());
    }
    // pass Y: lines (vz, vx)
    {
        long nlines = (long)nplanes * c->nx;
        VC_LAUNCH(c, "edt_pass_y", k_pass_xy<false>, vc_blocks((size_t)nlines, XY_THREADS_D), XY_THREADS_D, 0, g2,
                  (u64*)nullptr, c->id.as<int>() + off * c->nx * c->ny, c->d2.as<u32>() + off * c->nx * c->ny, stack, nlines,
                  c->nx, (long)CY * c->nx, (long)c->nx, CY, c->ny, (const u32*)nullptr);
    }
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
    VC_CUDA(c, c->g1.ensure((size_t)nplanes * CX * CY * 8));
    VC_CUDA(c, c->g2.ensure((size_t)nplanes * CY * c->nx * 8));
    VC_CUDA(c, c->id.ensure(nv * 4));
    VC_CUDA(c, c->d2.ensure(nv * 4));
    VC_CUDA(c, c->stk.ensure((size_t)regions * stack_region_entries(c, region_planes) * sizeof(u64)));
    return VC_OK;
}

// the three passes for the planes [zb, ze) of the slab, on stream c->cur (buffers from edt_alloc)
int edt_range(vc_ctx* c, int zb, int ze, int region, int region_planes)
{
    const int CX = c->nx + 1, CY = c->ny + 1;
    const int nplanes = ze - zb;
    const int nlines = CX * CY;
    VC_LAUNCH(c, "edt_pass_z", k_pass_z, dim3(vc_blocks((size_t)nlines, 256), (nplanes + PZ_CHUNK - 1) / PZ_CHUNK), 256, 0,
              c->line_ptr.as<int>(), c->line_ent.as<u64>(), c->g1.as<u64>() + (size_t)(zb - c->z0) * nlines, nlines, zb, ze);
    launch_passes(c, zb, nplanes, c->stk.as<u64>() + (size_t)region * stack_region_entries(c, region_planes));
    return VC_OK;
}

int st_closest_lattice(vc_ctx* c)
{
    VC_TRY(edt_alloc(c, 1, c->zc - c->z0));
    VC_TRY(edt_range(c, c->z0, c->zc, 0, c->zc - c->z0));
    VC_CUDA(c, cudaGetLastError());
    c->have_closest = true;
    c->have_measures = false;
    return VC_OK;
}

// Closest sites + measures of the whole slab as a pipeline over z chunks.  Chunk k = planes
// [zb, ze): its three transform passes cover [zb, ze+1) -- the halo plane its measures reach up to
// is recomputed rather than waited for (the same rule as between GPUs; chunk
// k+1 stores the identical values again) -- then its measures run on the same worker stream.
// Chunks are dealt round-robin to the worker streams, so there is no device-wide barrier between
// the passes: the tail of one chunk's stage overlaps whatever the other streams are running.
// With profiling on, everything runs on the main stream so per-kernel event times stay meaningful.
int st_closest_measures_pipelined(vc_ctx* c, bool want_radius)
{
    if (!c->lattice)
    {
        VC_TRY(st_closest_general_grid(c));
        return st_measures(c, want_radius);
    }
    const int nplanes_all = c->zc - c->z0;
    const int nw = c->profiling ? 0 : c->nworkers;
    // chunk height: >= 32k lines per launch keeps a chunk's kernels efficient on their own (measured:
    // 64 planes at 512^2, profiles/), VC_ZCHUNK overrides; one chunk when profiling
    int zchunk = c->zchunk;
    if (zchunk <= 0)
    {
        const int side = c->nx < c->ny ? c->nx : c->ny;
        zchunk = (32768 + side - 1) / side;
        zchunk = zchunk < 8 ? 8 : zchunk;
        // ... and no more chunks than about one per worker stream: with many more (1024^3 on one GPU: 32) the halo
        // recompute and the launch count cost more than the overlap gives (30.6 vs 27.6 ms unpipelined, measured)
        const int per_worker = (nplanes_all + (nw > 0 ? nw : 1) - 1) / (nw > 0 ? nw : 1);
        zchunk = zchunk < per_worker ? per_worker : zchunk;
    }
    if (!nw || zchunk > nplanes_all)
        zchunk = nplanes_all;
    const int nchunks = (nplanes_all + zchunk - 1) / zchunk;
    VC_TRY(edt_alloc(c, nchunks, zchunk + 1));
    if (!c->have_inside)
        return vc_fail(c, VC_ERR_STATE, "measures need vc_classify_grid");
    if (c->zhi < c->zc)
        return vc_fail(c, VC_ERR_STATE, "inside flags do not cover the halo plane");
    const bool dense_measures = !c->skip_dense_measures;
    if (dense_measures)
        VC_TRY(measures_alloc(c, want_radius));
    if (nw)
    {
        VC_CUDA(c, cudaEventRecord(c->ev_fork, c->stream));
        for (int i = 0; i < nw; ++i)
            VC_CUDA(c, cudaStreamWaitEvent(c->workers[i], c->ev_fork, 0));
    }
    int k = 0, status = VC_OK;
    for (int zb = c->z0; zb < c->zc && status == VC_OK; zb += zchunk, ++k)
    {
        const int ze = zb + zchunk < c->zc ? zb + zchunk : c->zc;
        const int zh = ze < c->zc ? ze + 1 : ze; // transform range incl. the halo plane
        c->cur = nw ? c->workers[k % nw] : c->stream;
        status = edt_range(c, zb, zh, k, zchunk + 1);
        const int me = ze < c->z1 ? ze : c->z1;
        if (status == VC_OK && zb < me && dense_measures)
            status = measures_range(c, zb, me, want_radius);
        if (status == VC_OK && zb < me && c->chunk_hook)
            status = c->chunk_hook(zb, me); // e.g. compaction + device-to-host copy of this chunk's records
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
