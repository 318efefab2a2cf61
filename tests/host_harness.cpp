// tests/host_harness.cpp -- TEST ONLY.  Drives the exact-arithmetic core the CUDA kernels are built
// from (voxel_ma_b200/csrc/vc_core.h: vc_site_key, vc_nearest_on_zline, vc_envelope_line)
// line by line on the CPU, so the algorithm can be checked against the oracle in the `not gpu`
// suite.  Nothing here is reachable from the product library.
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <vector>

#include "../voxel_ma_b200/csrc/vc_core.h"
#include "../voxel_ma_b200/csrc/vc_mesh_core.h"

extern "C"
{
    // site numbering: corners of the whole grid -> sites in reference order. Returns count.
    int64_t hh_sites(const uint8_t* inside, int nx, int ny, int nz, float* out_xyz, int64_t cap)
    {
        std::vector<std::pair<vc_u64, vc_u64>> recs;
        auto occ_at = [&](int x, int y, int z, uint32_t& inb) -> uint32_t
        {
            if (x < 0 || x >= nx || y < 0 || y >= ny || z < 0 || z >= nz)
            {
                inb = 0;
                return 0;
            }
            inb = 1;
            return inside[(size_t)x + (size_t)nx * ((size_t)y + (size_t)ny * z)] ? 1u : 0u;
        };
        for (int cz = 0; cz <= nz; ++cz)
            for (int cy = 0; cy <= ny; ++cy)
                for (int cx = 0; cx <= nx; ++cx)
                {
                    uint32_t occ = 0, inb = 0;
                    for (int bit = 0; bit < 8; ++bit)
                    {
                        uint32_t ib;
                        uint32_t o = occ_at(cx - 1 + (bit >> 2), cy - 1 + ((bit >> 1) & 1), cz - 1 + (bit & 1), ib);
                        occ |= o << bit;
                        inb |= ib << bit;
                    }
                    vc_u64 k = vc_site_key(occ, inb, cx, cy, cz, ny, nz);
                    if (k != VC_INF)
                        recs.emplace_back(k, vc_pack_corner(cx, cy, cz));
                }
        std::sort(recs.begin(), recs.end());
        int64_t n = (int64_t)recs.size();
        for (int64_t i = 0; i < n && i < cap; ++i)
        {
            int cx, cy, cz;
            vc_unpack_corner(recs[i].second, cx, cy, cz);
            out_xyz[3 * i] = (float)cx - 0.5f;
            out_xyz[3 * i + 1] = (float)cy - 0.5f;
            out_xyz[3 * i + 2] = (float)cz - 0.5f;
        }
        return n;
    }

    // separable exact transform with the kernels' data layout:
    //   pass Z: z-line site lists -> G1[vz][cx][cy];  pass X: -> G2[vz][cy][vx];  pass Y: -> out[vz][vy][vx]
    // sites: corner indices (cx,cy,cz) per site id. Output planes [z0,z1).
    void hh_closest_grid(const int32_t* corners, int64_t ns, int nx, int ny, int nz, int z0, int z1,
                         int32_t* id_out, uint32_t* d2x4_out)
    {
        const int CX = nx + 1, CY = ny + 1;
        const int nzs = z1 - z0;
        // line lists
        std::vector<std::vector<vc_u64>> lines((size_t)CX * CY);
        for (int64_t s = 0; s < ns; ++s)
            lines[(size_t)corners[3 * s] * CY + corners[3 * s + 1]].push_back(
                ((vc_u64)corners[3 * s + 2] << 32) | (uint32_t)s);
        for (auto& l : lines)
            std::sort(l.begin(), l.end());
        std::vector<vc_u64> G1((size_t)nzs * CX * CY), G2((size_t)nzs * CY * nx);
        for (int cx = 0; cx < CX; ++cx)
            for (int cy = 0; cy < CY; ++cy)
            {
                const auto& l = lines[(size_t)cx * CY + cy];
                int last = (int)l.size();
                int lo = -1;
                for (int vz = z0; vz < z1; ++vz)
                {
                    while (lo + 1 < last && (int)(l[lo + 1] >> 32) <= vz)
                        ++lo;
                    G1[((size_t)(vz - z0) * CX + cx) * CY + cy] = vc_nearest_on_zline(l.data(), 0, last, lo, vz);
                }
            }
        int maxc = std::max(CX, CY) + 1;
        std::vector<vc_u64> stkv(maxc);
        vc_stack_array stk{stkv.data()};
        for (int vz = 0; vz < nzs; ++vz)
            for (int cy = 0; cy < CY; ++cy)
            {
                vc_u64* row = &G2[((size_t)vz * CY + cy) * nx];
                vc_envelope_line(&G1[(size_t)vz * CX * CY + cy], (long)CY, CX, nx, stk,
                                 [&](int t, uint32_t V, uint32_t id) { row[t] = ((vc_u64)V << 32) | id; });
            }
        for (int vz = 0; vz < nzs; ++vz)
            for (int vx = 0; vx < nx; ++vx)
                vc_envelope_line(&G2[(size_t)vz * CY * nx + vx], (long)nx, CY, ny, stk,
                                 [&](int t, uint32_t V, uint32_t id)
                                 {
                                     size_t o = (size_t)vx + (size_t)nx * ((size_t)t + (size_t)ny * vz);
                                     id_out[o] = (int32_t)id;
                                     d2x4_out[o] = V;
                                 });
    }


    // The round-2 data flow of the transform (vc_edt.cu): compact site-bearing columns, live rows, pruned
    // envelopes.  stats (nullable, 8 x int64): [0] lines X, [1] candidates X, [2] pops X, [3] max depth X,
    // [4..7] the same for pass Y; depth_hist (nullable, 2 x 64 x int64): histogram of the per-line max depth, bin = min(depth, 63)
    void hh_closest_grid2(const int32_t* corners, int64_t ns, int nx, int ny, int nz, int z0, int z1, int32_t* id_out,
                          uint32_t* d2x4_out, int64_t* stats, int64_t* depth_hist)
    {
        const int CX = nx + 1, CY = ny + 1;
        const int nzs = z1 - z0;
        // z-line lists per column, then the compact list of columns that hold sites, row by row (cy major, cx ascending)
        std::vector<std::vector<vc_u64>> lines((size_t)CX * CY);
        for (int64_t s = 0; s < ns; ++s)
            lines[(size_t)corners[3 * s + 1] * CX + corners[3 * s]].push_back(((vc_u64)corners[3 * s + 2] << 32) | (uint32_t)s);
        std::vector<int> rowptr(CY + 1, 0), colx, colline, liverow;
        for (int cy = 0; cy < CY; ++cy)
        {
            for (int cx = 0; cx < CX; ++cx)
                if (!lines[(size_t)cy * CX + cx].empty())
                {
                    std::sort(lines[(size_t)cy * CX + cx].begin(), lines[(size_t)cy * CX + cx].end());
                    colx.push_back(cx);
                    colline.push_back(cy * CX + cx);
                }
            rowptr[cy + 1] = (int)colx.size();
            if (rowptr[cy + 1] > rowptr[cy])
                liverow.push_back(cy);
        }
        const int ncol = (int)colx.size(), nlive = (int)liverow.size();
        // pass Z: G1c[col][vz]
        std::vector<vc_u64> G1((size_t)ncol * nzs), G2((size_t)nzs * std::max(nlive, 1) * nx);
        for (int c = 0; c < ncol; ++c)
        {
            const auto& l = lines[colline[c]];
            int last = (int)l.size(), lo = -1;
            for (int vz = z0; vz < z1; ++vz)
            {
                while (lo + 1 < last && (int)(l[lo + 1] >> 32) <= vz)
                    ++lo;
                G1[(size_t)c * nzs + (vz - z0)] = vc_nearest_on_zline(l.data(), 0, last, lo, vz);
            }
        }
        std::vector<vc_ent> stkv(std::max(CX, CY) + 2);
        int64_t st[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        // pass X: lines (live row, vz) over the row's compact columns -> G2[vz][live row][vx]
        for (int r = 0; r < nlive; ++r)
            for (int vz = 0; vz < nzs; ++vz)
            {
                const int cy = liverow[r], c0 = rowptr[cy], nc = rowptr[cy + 1] - c0;
                vc_u64* row = &G2[((size_t)vz * nlive + r) * nx];
                vc_pstack_array stk{stkv.data()};
                vc_psource_array src{&G1[(size_t)c0 * nzs + vz], (long)nzs, &colx[c0]};
                vc_envelope_pruned(src, nc, nx, stk,
                                   [&](int t, uint32_t V, uint32_t id) { row[t] = ((vc_u64)V << 32) | id; });
                st[0]++, st[1] += nc, st[2] += stk.npop, st[3] = std::max<int64_t>(st[3], stk.maxdepth + 1);
                if (depth_hist)
                    depth_hist[std::min(stk.maxdepth + 1, 63)]++;
            }
        // pass Y: lines (vz, vx) over the live rows -> out[vz][vy][vx]
        for (int vz = 0; vz < nzs; ++vz)
            for (int vx = 0; vx < nx; ++vx)
            {
                vc_pstack_array stk{stkv.data()};
                vc_psource_array src{&G2[(size_t)vz * nlive * nx + vx], (long)nx, liverow.data()};
                vc_envelope_pruned(src, nlive, ny, stk,
                                   [&](int t, uint32_t V, uint32_t id)
                                   {
                                       size_t o = (size_t)vx + (size_t)nx * ((size_t)t + (size_t)ny * vz);
                                       id_out[o] = (int32_t)id;
                                       d2x4_out[o] = V;
                                   });
                st[4]++, st[5] += nlive, st[6] += stk.npop, st[7] = std::max<int64_t>(st[7], stk.maxdepth + 1);
                if (depth_hist)
                    depth_hist[64 + std::min(stk.maxdepth + 1, 63)]++;
            }
        if (stats)
            memcpy(stats, st, sizeof st);
    }

    // one line through the pruned scan: the finite candidates of `in` are compacted (position list), as the kernels
    // only ever see live candidates
    void hh_envelope_pruned(const vc_u64* in, int ncand, int ntgt, vc_u64* out)
    {
        std::vector<vc_u64> h;
        std::vector<int> pos;
        for (int j = 0; j < ncand; ++j)
            if (in[j] != VC_INF)
                h.push_back(in[j]), pos.push_back(j);
        std::vector<vc_ent> stkv(ncand + 2);
        vc_pstack_array stk{stkv.data()};
        vc_psource_array src{h.data(), 1L, pos.data()};
        vc_envelope_pruned(src, (int)h.size(), ntgt, stk,
                           [&](int t, uint32_t V, uint32_t id) { out[t] = ((vc_u64)V << 32) | id; });
    }

    // vc_sep against plain integer floor division: the number of mismatches over all spacings w in [1, 2048], line
    // lengths ntgt and the numerators around every multiple of 8w in the range the scan can produce (N > 0)
    int64_t hh_sep_sweep(void)
    {
        int64_t bad = 0;
        const int ntgts[] = {1, 2, 7, 64, 511, 512, 1024, 2047, 2048};
        for (int w = 1; w <= 2048; ++w)
        {
            const long long d = 8LL * w;
            for (int ntgt : ntgts)
            {
                auto check = [&](long long N)
                { // N = dg + c - 1 + 4w  ->  dg = N + 1 - 4w with c = 0
                    if (N <= 0 || N > (1LL << 28))
                        return;
                    const int got = vc_sep((int)(N + 1 - 4 * w), 0, w, ntgt);
                    const long long want = N / d;
                    if (want >= ntgt ? got < ntgt : got != (int)want)
                        ++bad;
                };
                for (long long k = 0; k <= ntgt + 2; k += (ntgt > 600 && w > 64) ? 7 : 1)
                    for (int e = -2; e <= 9; ++e)
                        check(k * d + e);
                check((1LL << 28) - 1), check((1LL << 27) + 12345);
            }
        }
        return bad;
    }

    // one line of the transform on caller data (robustness tests with 2048-scale coordinates)
    void hh_envelope(const vc_u64* in, int ncand, int ntgt, vc_u64* out)
    {
        std::vector<vc_u64> stkv(ncand + 1);
        vc_stack_array stk{stkv.data()};
        vc_envelope_line(in, 1L, ncand, ntgt, stk, [&](int t, uint32_t V, uint32_t id) { out[t] = ((vc_u64)V << 32) | id; });
    }

    // the same line with the candidate bitmap pass X uses (vc_edt.cu): bit j set <=> candidate j is live; the entries of
    // dead candidates are overwritten with garbage first -- the scan must never look at them
    void hh_envelope_masked(const vc_u64* in, int ncand, int ntgt, vc_u64* out)
    {
        std::vector<uint32_t> mask((size_t)(ncand + 31) / 32 + 1, 0u);
        std::vector<vc_u64> g(in, in + ncand);
        for (int j = 0; j < ncand; ++j)
        {
            if (in[j] != VC_INF)
                mask[j >> 5] |= 1u << (j & 31);
            else
                g[j] = 0x0123456789ABCDEFull * (vc_u64)(j + 1); // what pass Z leaves behind: never-written memory
        }
        std::vector<vc_u64> stkv(ncand + 1);
        vc_stack_array stk{stkv.data()};
        vc_envelope_line(g.data(), 1L, ncand, ntgt, stk, [&](int t, uint32_t V, uint32_t id) { out[t] = ((vc_u64)V << 32) | id; },
                         mask.data());
    }

    // (key, corner) records of one z-slab, as vc_sites_detect_local reports them: corner planes
    // [czb,cze) of a grid nx*ny*nz; `inside` holds voxel planes [zlo, zhi). Returns the count.
    int64_t hh_site_records(const uint8_t* inside, int nx, int ny, int nz, int zlo, int zhi, int czb, int cze,
                            vc_u64* keys, vc_u64* corners, int64_t cap)
    {
        int64_t n = 0;
        for (int cz = czb; cz < cze; ++cz)
            for (int cy = 0; cy <= ny; ++cy)
                for (int cx = 0; cx <= nx; ++cx)
                {
                    uint32_t occ = 0, inb = 0;
                    for (int bit = 0; bit < 8; ++bit)
                    {
                        int x = cx - 1 + (bit >> 2), y = cy - 1 + ((bit >> 1) & 1), z = cz - 1 + (bit & 1);
                        bool in = x >= 0 && x < nx && y >= 0 && y < ny && z >= 0 && z < nz;
                        if (in && (z < zlo || z >= zhi))
                            return -1; // slab does not hold a plane it needs
                        uint32_t o = in ? (inside[(size_t)x + (size_t)nx * ((size_t)y + (size_t)ny * (z - zlo))] ? 1u : 0u) : 0u;
                        occ |= o << bit;
                        inb |= (in ? 1u : 0u) << bit;
                    }
                    vc_u64 k = vc_site_key(occ, inb, cx, cy, cz, ny, nz);
                    if (k == VC_INF)
                        continue;
                    if (n < cap)
                    {
                        keys[n] = k;
                        corners[n] = vc_pack_corner(cx, cy, cz);
                    }
                    ++n;
                }
        return n;
    }
    // Mesh parity classification with the kernels' own rules (vc_mesh_core.h) and data flow: snapped
    // vertices -> one toggle bit per (triangle, covered column) in rows of nx+1 bits -> exclusive suffix
    // parity per row.  q = already transformed float coordinates (identity M).  Returns 0 / 1 (range).
    int hh_classify_mesh(const float* q, int64_t nv, const uint32_t* tris, int64_t nt, int nx, int ny, int nz, uint8_t* inside)
    {
        std::vector<int> Q((size_t)nv * 3);
        for (int64_t i = 0; i < 3 * nv; ++i)
            if (!vc_mesh_snap(q[i], &Q[i]))
                return 1;
        const int wr = nx / 32 + 1;
        std::vector<uint32_t> tog((size_t)ny * nz * wr, 0u);
        for (int64_t t = 0; t < nt; ++t)
        {
            int a[3], b[3], c[3];
            for (int d = 0; d < 3; ++d)
                a[d] = Q[3 * tris[3 * t] + d], b[d] = Q[3 * tris[3 * t + 1] + d], c[d] = Q[3 * tris[3 * t + 2] + d];
            if (!vc_mesh_orient_ccw(&a[0], &a[1], &a[2], &b[0], &b[1], &b[2], &c[0], &c[1], &c[2]))
                continue;
            int j0, j1, k0, k1;
            vc_mesh_columns(a[1], a[2], b[1], b[2], c[1], c[2], ny, 0, nz, &j0, &j1, &k0, &k1);
            for (int k = k0; k <= k1; ++k)
                for (int j = j0; j <= j1; ++j)
                {
                    int T;
                    if (vc_mesh_crossing(a[0], a[1], a[2], b[0], b[1], b[2], c[0], c[1], c[2], j, k, nx, &T) && T > 0)
                        tog[((size_t)k * ny + j) * wr + (T >> 5)] ^= 1u << (T & 31);
                }
        }
        for (size_t row = 0; row < (size_t)ny * nz; ++row)
        {
            int par = 0;
            for (int x = nx; x >= 0; --x)
            { // inside(x) = parity of the toggles strictly above x
                if (x < nx)
                    inside[row * nx + x] = (uint8_t)par;
                par ^= (tog[row * wr + (x >> 5)] >> (x & 31)) & 1u;
            }
        }
        return 0;
    }
}
