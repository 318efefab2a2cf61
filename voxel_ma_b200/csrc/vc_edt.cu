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
    const int zb = z0 + blockIdx.y * PZ_CHUNK;
    const int ze = min(zb + PZ_CHUNK, zc);
    const int first = line_ptr[l], last = line_ptr[l + 1];
    const size_t plane = (size_t)nlines;
    u64* out = G1 + l + plane * (size_t)(zb - z0);
    if (first == last)
    {
        for (int vz = zb; vz < ze; ++vz, out += plane)
            __stcs(out, (u64)VC_INF);
        return;
    }
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

// ---- passes X and Y ------------------------------------------------------------------------------
// TRANSPOSE = true  (pass X): line g = (vz, cy); input G1 + vz*CX*CY + cy, stride CY; the outputs of
//   32 lines x XY_TW targets are staged in shared memory and written as rows of XY_TW*8 = 128 bytes
//   of G2[g][vx].
// TRANSPOSE = false (pass Y): line g = (vz, vx); input G2 + vz*CY*nx + vx, stride nx; outputs go
//   straight to id/d2x4[(vz*ny + vy)*nx + vx], coalesced across the warp.
// Both are capped at 64 registers so that 32 warps are resident per SM: each thread keeps
// 2*VC_PF candidate loads in flight (vc_core.h), which covers the HBM latency-bandwidth product.
#define XY_THREADS_T 128 // pass X: 4 warps x 4.25 KB of transpose tile
#define XY_THREADS_D 256 // pass Y
#define XY_TW 16         // targets per transposed store burst
template <int MAXC, bool TRANSPOSE>
__global__ void __launch_bounds__(TRANSPOSE ? XY_THREADS_T : XY_THREADS_D, TRANSPOSE ? 8 : 4)
    k_pass_xy(const u64* __restrict__ in, u64* __restrict__ G2, int* __restrict__ id_out, u32* __restrict__ d2_out,
              long nlines_total, int lines_per_plane, long in_plane_stride, long in_stride, int ncand, int ntgt)
{
    __shared__ u64 tile[TRANSPOSE ? XY_THREADS_T / 32 : 1][TRANSPOSE ? 32 : 1][TRANSPOSE ? XY_TW + 1 : 1];
    u64 stH[MAXC];
    u32 stPT[MAXC];
    const long g = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = g < nlines_total;
    const long plane = valid ? g / lines_per_plane : 0;
    const int within = valid ? (int)(g - plane * lines_per_plane) : 0;
    const u64* src = in + plane * in_plane_stride + within;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long gwarp = g - lane;

    if (TRANSPOSE)
    {
        const int rsub = lane / XY_TW, col = lane % XY_TW; // a store instruction covers 32/XY_TW rows
        vc_envelope_line(src, in_stride, valid ? ncand : 0, ntgt, stH, stPT,
                         [&](int t, u64 v)
                         {
                             tile[warp][lane][t % XY_TW] = v;
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
                         });
    }
    else
    {
        const long last = plane * (long)ntgt * lines_per_plane + within + (long)(ntgt - 1) * lines_per_plane;
        int* pid = id_out + last;
        u32* pd2 = d2_out + last;
        vc_envelope_line(src, in_stride, valid ? ncand : 0, ntgt, stH, stPT,
                         [&](int t, u64 v)
                         { // targets arrive as ntgt-1 .. 0
                             if (valid)
                             {
                                 __stcs(pid, (int)(u32)v);
                                 __stcs(pd2, (u32)(v >> 32));
                             }
                             pid -= lines_per_plane;
                             pd2 -= lines_per_plane;
                         });
    }
}

template <int MAXC>
static int launch_passes(vc_ctx* c, int nplanes)
{
    const int CX = c->nx + 1, CY = c->ny + 1;
    // pass X: lines (vz, cy)
    {
        long nlines = (long)nplanes * CY;
        VC_LAUNCH(c, "edt_pass_x", (k_pass_xy<MAXC, true>), vc_blocks((size_t)nlines, XY_THREADS_T), XY_THREADS_T, 0, c->g1.as<u64>(),
                  c->g2.as<u64>(), (int*)nullptr, (u32*)nullptr, nlines, CY, (long)CX * CY, (long)CY, CX, c->nx);
    }
    // pass Y: lines (vz, vx)
    {
        long nlines = (long)nplanes * c->nx;
        VC_LAUNCH(c, "edt_pass_y", (k_pass_xy<MAXC, false>), vc_blocks((size_t)nlines, XY_THREADS_D), XY_THREADS_D, 0, c->g2.as<u64>(),
                  (u64*)nullptr, c->id.as<int>(), c->d2.as<u32>(), nlines, c->nx, (long)CY * c->nx, (long)c->nx, CY,
                  c->ny);
    }
    return VC_OK;
}

int st_closest_lattice(vc_ctx* c)
{
    if (!c->have_sites)
        return vc_fail(c, VC_ERR_STATE, "vc_closest_grid needs sites (vc_extract_sites / vc_set_sites)");
    const int CX = c->nx + 1, CY = c->ny + 1;
    const int nplanes = c->zc - c->z0;
    const size_t nv = (size_t)c->nx * c->ny * nplanes;
    VC_CUDA(c, c->g1.ensure((size_t)nplanes * CX * CY * 8));
    VC_CUDA(c, c->g2.ensure((size_t)nplanes * CY * c->nx * 8));
    VC_CUDA(c, c->id.ensure(nv * 4));
    VC_CUDA(c, c->d2.ensure(nv * 4));
    const int nlines = CX * CY;
    VC_LAUNCH(c, "edt_pass_z", k_pass_z, dim3(vc_blocks((size_t)nlines, 256), (nplanes + PZ_CHUNK - 1) / PZ_CHUNK), 256, 0, c->line_ptr.as<int>(),
              c->line_ent.as<u64>(), c->g1.as<u64>(), nlines, c->z0, c->zc);
    int m = (CX > CY ? CX : CY) + 1;
    if (m <= 264)
        launch_passes<264>(c, nplanes);
    else if (m <= 520)
        launch_passes<520>(c, nplanes);
    else if (m <= 1032)
        launch_passes<1032>(c, nplanes);
    else if (m <= 2056)
        launch_passes<2056>(c, nplanes);
    else
        return vc_fail(c, VC_ERR_UNSUPPORTED, "grid side above 2048 is not supported by the dense transform");
    VC_CUDA(c, cudaGetLastError());
    c->have_closest = true;
    c->have_measures = false;
    return VC_OK;
}
