// vc_core.h -- the exact integer arithmetic shared by the kernels (host+device inline functions).
//
// Everything the closest-site transform decides is decided here, in integers, so that "bit-exact
// ids" has one definition.  The functions are __host__ __device__ only so that tests/ can drive
// the very same code line by line on the CPU against the oracle (tests/host_harness.cpp); the
// product calls them from CUDA kernels only.
//
// Geometry.  Grid vertices are the integer lattice points v in [0,n)^3 (voxel centres,
// include/surfacing.h:97-119).  Sites are voxel corners at half-integer coordinates c - 0.5 with
// corner index c in [0,n]^3 (include/surfacing.h:121-136).  In doubled coordinates a vertex is 2v
// (even), a site is 2c-1 (odd), so along one axis the offset is  delta(v,c) = 2(v-c)+1  (odd) and
//     4*d^2 = delta_x^2 + delta_y^2 + delta_z^2      is an exact integer (< 2^26 for n <= 2048).
//
// Contract (SURVEY section 7-1, 3rdparty/ann/src/brute.cpp:56-82): the winner at a vertex is the
// lexicographic minimum of (d^2, site id).  A candidate is carried as one 64-bit word
//     H = (4*d^2_so_far << 32) | id
// so a lexicographic compare is one unsigned compare, and the minimum over all sites separates
// exactly over the three axes: min_s (dx^2+dy^2+dz^2, id) = min_cx [dx^2 + min_cy [dy^2 + min_cz (dz^2, id)]].
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define VC_HD __host__ __device__ __forceinline__
#else
#define VC_HD inline
#endif

typedef unsigned long long vc_u64; // (uint64_t is unsigned long on LP64; CUDA atomics want unsigned long long)
#define VC_INF 0xFFFFFFFFFFFFFFFFull

// floor(num / den), den > 0, |num| < 2^27, den < 2^15: float estimate, exact integer fix-up.
VC_HD int vc_floor_div(int num, int den)
{
#if defined(__CUDA_ARCH__)
    int q = __float2int_rd(__fdividef((float)num, (float)den));
#else
    float qf = (float)num / (float)den;
    int q = (int)qf;
    if ((float)q > qf)
        --q;
#endif
    int r = num - q * den;
    while (r < 0)
    {
        --q;
        r += den;
    }
    while (r >= den)
    {
        ++q;
        r -= den;
    }
    return q;
}

// value of candidate (position p on the corner axis, word H) at target vertex t on that axis
VC_HD vc_u64 vc_eval(vc_u64 H, int p, int t)
{
    int d = 2 * (t - p) + 1;
    return H + ((vc_u64)(uint32_t)(d * d) << 32);
}

// First target t at which candidate k (position pk > pi) is lexicographically better than
// candidate i.  With u = 2t+1-pi-pk, m = 4(pk-pi), A = D_k - D_i:
//     f_k(t) - f_i(t) = A - m*u,   so k wins  <=>  m*u > A + e,  e = (id_k - id_i) * 2^-32.
// floor((A+e)/m) = floor((A - [id_k < id_i]) / m), hence t >= ceil((F + pi + pk) / 2).
// (A tie in distance goes to the lower id: this is where the reference's tie rule lives.)
VC_HD int vc_sep(int pi, vc_u64 Hi, int pk, vc_u64 Hk)
{
    int A = (int)(uint32_t)(Hk >> 32) - (int)(uint32_t)(Hi >> 32);
    int c = ((uint32_t)Hk < (uint32_t)Hi) ? 1 : 0;
    int F = vc_floor_div(A - c, 4 * (pk - pi));
    return (F + pi + pk + 1) >> 1; // arithmetic shift: floor((x+1)/2) = ceil(x/2)
}

// One line of the separable transform: lower envelope of the candidates in[j*stride], j in
// [0,ncand), evaluated at targets t in [0,ntgt).  The forward scan keeps the envelope as a stack of
// (word, position, first target it wins); its two upper entries live in registers (top, sec), the
// rest in stH/stPT (thread-local memory on the device), so a pop is a register move plus a load
// whose result is not needed before the NEXT pop -- the dependent chain never waits on memory.
// The backward scan emits targets ntgt-1 .. 0 through `emit(t, value)`.
//
// The candidates are fetched VC_PF at a time into two register banks that alternate: while one
// bank is consumed the loads of the next are already in flight, so a thread keeps up to 2*VC_PF
// independent 8-byte loads outstanding instead of one.
#define VC_PF 4

#if defined(__CUDA_ARCH__)
#define VC_LOAD_STREAM(p) __ldcs(p) // read once: evict-first
#else
#define VC_LOAD_STREAM(p) (*(p))
#endif

struct vc_env_state
{
    int q;        // index of the top entry; -1 = empty.  entry q = top, q-1 = sec, 0..q-2 in memory
    vc_u64 Hs;    // top
    int ps, ts;
    vc_u64 H2;    // second
    int p2, t2;
};

VC_HD void vc_env_pop(vc_env_state& s, const vc_u64* stH, const uint32_t* stPT)
{
    --s.q;
    s.Hs = s.H2;
    s.ps = s.p2;
    s.ts = s.t2;
    if (s.q >= 1)
    {
        s.H2 = stH[s.q - 1];
        uint32_t pt = stPT[s.q - 1];
        s.p2 = (int)(pt & 0xFFFFu);
        s.t2 = (int)(pt >> 16);
    }
}

VC_HD void vc_env_push(vc_env_state& s, vc_u64 H, int j, int ntgt, vc_u64* stH, uint32_t* stPT)
{
    if (H == VC_INF)
        return;
    // the top loses already where its interval starts: it wins nowhere
    while (s.q >= 0 && vc_eval(s.Hs, s.ps, s.ts) > vc_eval(H, j, s.ts))
        vc_env_pop(s, stH, stPT);
    int w = 0;
    if (s.q >= 0)
    {
        w = vc_sep(s.ps, s.Hs, j, H);
        if (w >= ntgt)
            return; // wins only beyond the last target
        if (s.q >= 1)
        {
            stH[s.q - 1] = s.H2;
            stPT[s.q - 1] = (uint32_t)s.p2 | ((uint32_t)s.t2 << 16);
        }
        s.H2 = s.Hs;
        s.p2 = s.ps;
        s.t2 = s.ts;
    }
    ++s.q;
    s.Hs = H;
    s.ps = j;
    s.ts = w;
}

template <class Emit>
VC_HD void vc_envelope_line(const vc_u64* __restrict__ in, long stride, int ncand, int ntgt,
                            vc_u64* stH, uint32_t* stPT, Emit emit)
{
    vc_env_state s;
    s.q = -1;
    s.Hs = s.H2 = 0;
    s.ps = s.ts = s.p2 = s.t2 = 0;
    vc_u64 bankA[VC_PF], bankB[VC_PF];
#define VC_FETCH(bank, j0)                                                                         \
    _Pragma("unroll") for (int k = 0; k < VC_PF; ++k)                                              \
        bank[k] = ((j0) + k < ncand) ? VC_LOAD_STREAM(in + (long)((j0) + k) * stride) : (vc_u64)VC_INF;
#define VC_CONSUME(bank, j0)                                                                       \
    _Pragma("unroll") for (int k = 0; k < VC_PF; ++k) vc_env_push(s, bank[k], (j0) + k, ntgt, stH, stPT);
    VC_FETCH(bankA, 0)
    for (int j0 = 0; j0 < ncand; j0 += 2 * VC_PF)
    {
        VC_FETCH(bankB, j0 + VC_PF)
        VC_CONSUME(bankA, j0)
        VC_FETCH(bankA, j0 + 2 * VC_PF)
        VC_CONSUME(bankB, j0 + VC_PF)
    }
#undef VC_FETCH
#undef VC_CONSUME
    // one uniform backward loop (a line without candidates emits VC_INF) so that a warp whose
    // lanes each own a line stays convergent at the emit() call and may synchronise inside it
    const bool empty = s.q < 0;
    for (int t = ntgt - 1; t >= 0; --t)
    {
        emit(t, empty ? (vc_u64)VC_INF : vc_eval(s.Hs, s.ps, t));
        if (t == s.ts && s.q > 0)
            vc_env_pop(s, stH, stPT);
    }
}

// First pass (along z) straight from a line's sorted site list: entries e[i] = (cz << 32) | id,
// ascending cz.  `lo` = index of the last entry with cz <= vz (or first-1), maintained by the caller.
VC_HD vc_u64 vc_nearest_on_zline(const vc_u64* __restrict__ e, int first, int last, int lo, int vz)
{
    vc_u64 H = VC_INF;
    if (lo >= first)
    {
        vc_u64 en = e[lo];
        int d = 2 * (vz - (int)(en >> 32)) + 1;
        H = ((vc_u64)(uint32_t)(d * d) << 32) | (uint32_t)en;
    }
    if (lo + 1 < last)
    {
        vc_u64 en = e[lo + 1];
        int d = 2 * ((int)(en >> 32) - vz) - 1;
        vc_u64 H2 = ((vc_u64)(uint32_t)(d * d) << 32) | (uint32_t)en;
        H = H2 < H ? H2 : H;
    }
    return H;
}

// ---- site numbering -----------------------------------------------------------------------------
// Surfacer::extractBoundaryVts (src/surfacing.cpp:240-285) numbers a corner by its first encounter
// in the scan: voxels x outer / y / z inner, the 6 neighbours in the order -x,+x,-y,+y,-z,+z
// (include/surfacing.h:170-178), the 4 corners of the shared face in the order of
// include/surfacing.h:184-190.  So  id(corner) = rank of  min over emitters of
//     key = ((x*ny + y)*nz + z) * 24 + o*4 + ii          (validated in SURVEY section 7-2)
// and every emitter of a corner is one of its 8 incident voxels looking at an in-block neighbour.
//
// occ: bit (a*4+b*2+c) = occupancy of voxel (cx-1+a, cy-1+b, cz-1+c), 0 when out of bounds
// inb: same indexing, 1 when that voxel is inside the volume bounds.
// Returns the key, or VC_INF when the corner is not a site.
VC_HD vc_u64 vc_site_key(uint32_t occ, uint32_t inb, int cx, int cy, int cz, int ny, int nz)
{
    if (occ == 0u || (occ & inb) == 0xFFu)
        return VC_INF; // all 8 equal (out-of-bounds voxels read as 0)
    // position of corner slot ci inside m_cornersWRTNbOffset[o], packed 4 bits per (o, ci); 15 = absent
    // o=0:(2,3,7,6) o=1:(0,1,5,4) o=2:(4,6,2,0) o=3:(1,3,7,5) o=4:(0,2,3,1) o=5:(5,7,6,4)
    const uint32_t slot_of[6] = {
        0x23FF10FFu, // o=0: ci2->0 ci3->1 ci7->2 ci6->3        (nibble ci from the right)
        0xFF23FF10u, // o=1: ci0->0 ci1->1 ci5->2 ci4->3
        0xF1F0F2F3u, // o=2: ci4->0 ci6->1 ci2->2 ci0->3
        0x2F3F1F0Fu, // o=3: ci1->0 ci3->1 ci7->2 ci5->3
        0xFFFF2130u, // o=4: ci0->0 ci2->1 ci3->2 ci1->3
        0x1203FFFFu  // o=5: ci5->0 ci7->1 ci6->2 ci4->3
    };
    for (int bit = 0; bit < 8; ++bit)
    { // (a,b,c) lexicographic = increasing x, then y, then z = increasing scan index
        if (!((inb >> bit) & 1u))
            continue;
        int a = bit >> 2, b = (bit >> 1) & 1, c = bit & 1;
        uint32_t me = (occ >> bit) & 1u;
        int o = 6;
        if (((occ >> (bit ^ 1)) & 1u) != me)
            o = c == 0 ? 5 : 4;
        if (((occ >> (bit ^ 2)) & 1u) != me)
            o = b == 0 ? 3 : 2;
        if (((occ >> (bit ^ 4)) & 1u) != me)
            o = a == 0 ? 1 : 0;
        if (o == 6)
            continue;
        // the corner seen from this voxel: +0.5 along an axis when the voxel is the lower one
        int ci = (c == 0 ? 4 : 0) + (a == 0 ? 0 : 2) + (b == 0 ? 1 : 0);
        uint32_t ii = (slot_of[o] >> (4 * ci)) & 0xFu;
        vc_u64 L = ((vc_u64)(cx - 1 + a) * (vc_u64)ny + (vc_u64)(cy - 1 + b)) * (vc_u64)nz +
                     (vc_u64)(cz - 1 + c);
        return L * 24ull + (vc_u64)(o * 4) + ii;
    }
    return VC_INF;
}

// corner record exchanged between ranks: cx | cy<<21 | cz<<42
VC_HD vc_u64 vc_pack_corner(int cx, int cy, int cz)
{
    return (vc_u64)cx | ((vc_u64)cy << 21) | ((vc_u64)cz << 42);
}
VC_HD void vc_unpack_corner(vc_u64 p, int& cx, int& cy, int& cz)
{
    cx = (int)(p & 0x1FFFFFu);
    cy = (int)((p >> 21) & 0x1FFFFFu);
    cz = (int)((p >> 42) & 0x1FFFFFu);
}
