// assembly_node.cuh -- Jacobian assembly with a tangent per element / per Gauss point
// (src/assembly.c:85-108), node-centric: no scatter, no reduction, no atomics.
//
//   Ke[3a+d][3b+c] = wg * sum_gp sum_{p,q} h_a[p] * C_gp[voigt(d,p)][voigt(c,q)] * h_b[q]
//
// (h_n[p] = d N_n / d x_p at the Gauss point: the reference's B matrix, assembly.c:234-253, has
// B[voigt(d,p)][3n+d] = h_n[p] and nothing else).  A warp = one 32-node operator tile x one position
// (d, c) inside the 3 x 3 blocks, lane = node.  Thread (d, c, lane) owns the entry (d, c) of all 27 blocks
// of its node's operator row -- 27 accumulators -- and walks the (up to) eight elements around the node:
// for the element in which the node is local node a
//     T[q]            = sum_p h_a[p] * C[voigt(d,p)][voigt(c,q)]        9 FMA
//     acc[slot(a->b)] += sum_q T[q] * h_b[q]   for the 8 nodes b        24 FMA
// Every operator entry is produced by exactly one thread: nothing is zeroed, nothing is added in
// shared memory, and all nine (d, c) run the same instruction stream (the 3 x 3 sub-matrix of C they
// pick is data; the 24 shape-function derivatives of a Gauss point are 12 128-bit constant-bank loads
// into uniform registers).  2 112 FMA per thread and tile, 19 008 per node (the element-centric form
// of round 2's first half needs 17 280 plus eight shared-memory reduction rounds).  The symmetric
// layout (slots 13..26) keeps the pairs with b >= a only: 1 440 FMA per thread.
//
// Two kernels:
//  * k_assemble_nodes_uniform -- the north_star's case, D from constant memory: the (tile, d, c) jobs are
//    dealt to independent warps (no shared memory, no barrier, 20 warps per SM = 5 per scheduler), the
//    27 entries leave the registers directly (8-byte stores, the two halves of a 16-byte pair come from
//    two warps that work on the same tile at the same time and meet in L2);
//  * k_assemble_nodes_pergp -- tangents per Gauss point: one CTA of 9 warps per tile; the four element
//    rows around the tile (33 elements each) are staged through shared memory one Gauss point ahead
//    with cp.async (36 entries x 132 cells, two buffers), elements that do not exist are staged as zeros;
//    the finished tile is laid out in the buffers in the operator's pair-interleaved order and leaves
//    with ONE bulk copy (cp.async.bulk shared -> global, SASS UBLKCP).
#pragma once

#include "assembly_elem.cuh"

namespace macroc {

constexpr int ASMN_WARPS = 9;
constexpr int ASMN_THREADS = ASMN_WARPS * 32;
constexpr int ASMN_SROW = 33;                               // staged elements per element row
constexpr int ASMN_CELLS = 4 * ASMN_SROW;                   // (ey, ez) in {j-1, j} x {k-1, k}
constexpr int ASMN_BUF_DOUBLES = 36 * ASMN_CELLS;           // one Gauss point: 4 752 doubles
constexpr int ASMN_NBUF = 2;                                // Gauss point gp is integrated while gp+1 lands.  (Measured and rejected: a ring of three
                                                            // with one barrier per Gauss point needs 114.6 KB = one CTA per SM, 74 ms instead of 51.5; no
                                                            // staging at all -- tangents through L1 with prefetch.global.L1 one Gauss point ahead, no
                                                            // barrier -- 99.6 ms: long scoreboard 5.4 per issue, L1 hit rate 63 %.)
constexpr int ASMN_SMEM_PER_GP = ASMN_NBUF * ASMN_BUF_DOUBLES * 8 + ASMN_CELLS * 4;   // 76 560 B: two CTAs per SM
constexpr int ASMU_WARPS = 4, ASMU_CTAS_PER_SM = 5;         // uniform tangent: 20 independent warps per SM (a sixth CTA for the
                                                            // symmetric layout, 80 registers: 15.86 ms instead of 16.15 -- not worth a second setting)

__host__ __device__ __forceinline__ constexpr int node_rank(int n) { return node_px(n) + 2 * node_py(n) + 4 * node_pz(n); }
__host__ __device__ __forceinline__ constexpr int slot_of(int a, int b)
{
    return (node_pz(b) - node_pz(a) + 1) * 9 + (node_py(b) - node_py(a) + 1) * 3 + (node_px(b) - node_px(a) + 1);
}

// one Gauss point of one thread.  hb: the 24 shape-function derivatives (warp-uniform); Cu: the thread's 3 x 3
// sub-matrix of a uniform tangent; PER_GP: the staged tangents instead.  ALL: every element around the node
// exists; otherwise the contributions of the missing ones are zeroed with selects (one basic block either way).
template <bool PER_GP, bool SYM, bool ALL>
__device__ __forceinline__ void asmn_gauss_point(int gp, unsigned ex, const double (&Cu)[3][3], const double *__restrict__ stage_lane,
                                                 const int (&off)[3][3], double (&acc)[SYM ? 14 : 27])
{
    constexpr int S0 = SYM ? 13 : 0;
    const double2 *h2 = reinterpret_cast<const double2 *>(&c_dsh[gp][0][0]);
    double hb[24];
#pragma unroll
    for (int q = 0; q < 12; ++q) { const double2 t = h2[q]; hb[2 * q] = t.x; hb[2 * q + 1] = t.y; }
#pragma unroll
    for (int a = 0; a < 8; ++a) {
        const double h0 = hb[3 * a], h1 = hb[3 * a + 1], h2a = hb[3 * a + 2];
        double T[3];
        if (PER_GP) {
            // cell of the element in which the node is local node a: row (1 - py, 1 - pz), x = lane + 1 - px
            const double *ck = stage_lane + ((1 - node_py(a)) + 2 * (1 - node_pz(a))) * ASMN_SROW + 1 - node_px(a);
#pragma unroll
            for (int q = 0; q < 3; ++q) T[q] = fma(h2a, ck[off[2][q]], fma(h1, ck[off[1][q]], h0 * ck[off[0][q]]));
        } else {
#pragma unroll
            for (int q = 0; q < 3; ++q) T[q] = fma(h2a, Cu[2][q], fma(h1, Cu[1][q], h0 * Cu[0][q]));
            if (!ALL && !((ex >> a) & 1u)) T[0] = T[1] = T[2] = 0.;
        }
#pragma unroll
        for (int b = 0; b < 8; ++b) {
            if (SYM && node_rank(b) < node_rank(a)) continue;              // slot < 13: the mirrored half
            const int s = slot_of(a, b) - S0;
            acc[s] = fma(T[0], hb[3 * b], fma(T[1], hb[3 * b + 1], fma(T[2], hb[3 * b + 2], acc[s])));
        }
    }
}

// node of (tile, lane): local box coordinates (i, j), slab-local plane kl, linear index ln0 of lane 0
struct AsmTile {
    int i, j, kl, nvalid;
    int64_t ln0;
};
template <bool SYM>
__device__ __forceinline__ AsmTile asmn_decode(const GridDev &g, const SymGeom &sg, int64_t tile, int64_t tpp, int lane)
{
    AsmTile t;
    t.i = t.j = t.kl = 0;
    if (SYM) {
        t.kl = (int)((tile + tpp) / tpp) - 1;                              // floor: the ghost plane is -1
        const int rem2 = (int)(tile - (int64_t)t.kl * tpp);
        t.j = rem2 / sg.rt;
        const int x0 = (rem2 - t.j * sg.rt) * 32;
        t.i = x0 + lane;
        t.nvalid = min(32, g.NX - x0);
        t.ln0 = x0 + (int64_t)g.NX * t.j + g.npl * t.kl;
    } else {
        t.ln0 = tile * TILE_NODES;
        t.nvalid = (int)min((int64_t)32, g.nloc - t.ln0);
        if (lane < t.nvalid) {                                             // (a rank's local node count fits 31 bits)
            const unsigned lnu = (unsigned)(t.ln0 + lane), nx = (unsigned)g.NX, npl = (unsigned)g.npl;
            t.kl = (int)(lnu / npl);
            const unsigned inpl = lnu - (unsigned)t.kl * npl;
            t.j = (int)(inpl / nx); t.i = (int)(inpl - (unsigned)t.j * nx);
        }
    }
    return t;
}

// Dirichlet data of the row (dof bits of this node) and of the 27 columns (bit s: dof c of neighbour s is fixed).
// masksum[q] = OR of nodemask[32 q .. 32 q + 31]: most tiles have no Dirichlet dof in reach and load one byte.
template <bool SYM>
__device__ __forceinline__ void asmn_masks(const GridDev &g, const uint8_t *__restrict__ nodemask, const uint8_t *__restrict__ masksum,
                                           int64_t ln0, int lane, bool valid, int c, unsigned &own, unsigned &colmask)
{
    constexpr int S0 = SYM ? 13 : 0;
    own = 0; colmask = 0;
    unsigned any = 0;
    if (lane < 27) {
        // the 34 bytes [G + ln0 - 1 + off, G + ln0 + 32 + off] of row/plane offset o = lane / 3 touch at most 3 chunks
        const int o = lane / 3, kq = lane - 3 * o;
        const int64_t start = g.G + ln0 - 1 + (int64_t)g.NX * (o % 3 - 1) + g.npl * (o / 3 - 1);
        const int64_t q = (start >> 5) + kq;
        if (q >= 0 && q * 32 < g.S && q * 32 <= start + 33) any = masksum[q];
    }
    if (!__any_sync(0xffffffffu, any != 0)) return;
    if (valid) {
        const int64_t base = g.G + ln0 + lane;
        own = nodemask[base];
#pragma unroll
        for (int s = S0; s < 27; ++s) {
            const int64_t idx = base + (s % 3 - 1) + (int64_t)g.NX * ((s / 3) % 3 - 1) + g.npl * (s / 9 - 1);
            if (idx >= 0 && idx < g.S) colmask |= (((unsigned)nodemask[idx] >> c) & 1u) << s;
        }
    }
}

// wg, MatZeroRowsColumns (bcs.c:341-347), the ghost plane's mirrored half; PCJACOBI's diagonal on the way
template <bool SYM>
__device__ __forceinline__ double asmn_entry(int s, double accv, double wg, bool valid, bool rowfixed, unsigned colmask, bool diag_thread,
                                             bool ghost_plane)
{
    double val = valid ? accv * wg : 0.;
    if (rowfixed || ((colmask >> s) & 1u)) val = (s == 13 && diag_thread) ? 1. : 0.;
    if (SYM && ghost_plane && s < 18) val = 0.;                            // only the blocks towards the slab (slots 18..26) survive
    return val;
}

// ---- uniform tangent ------------------------------------------------------------------------------------
template <bool SYM>
__global__ void __launch_bounds__(ASMU_WARPS * 32, ASMU_CTAS_PER_SM)
k_assemble_nodes_uniform(GridDev g, SymGeom sg, ElemRange er, double wg, const uint8_t *__restrict__ nodemask,
                         const uint8_t *__restrict__ masksum, double2 *__restrict__ A, double *__restrict__ dinv, int64_t tile_lo,
                         int64_t tile_hi, int64_t tpp /* tiles per plane (symmetric layout) */)
{
    constexpr int NS = SYM ? 14 : 27, S0 = SYM ? 13 : 0;
    constexpr int TILE_D = SYM ? SYM_TILE_DOUBLES : TILE_DOUBLES;
    const int lane = threadIdx.x & 31;
    const int64_t gw = (int64_t)blockIdx.x * ASMU_WARPS + (threadIdx.x >> 5), nw = (int64_t)gridDim.x * ASMU_WARPS;
    const int64_t njobs = (tile_hi - tile_lo) * 9;
    const int off[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
    // job = (tile, d, c); the nine jobs of a tile run on nine warps at the same time (consecutive warp ids).
    // (All jobs take the same time; starting the CTAs of an SM a fifth of a job apart to keep their warps out of
    // phase changed nothing: 25.43 -> 25.44 ms.)
    for (int64_t job = gw; job < njobs; job += nw) {
        const int64_t tq = job / 9, tile = tile_lo + tq;
        const int e9 = (int)(job - tq * 9), d = e9 / 3, c = e9 - 3 * d;
        double Cu[3][3];
#pragma unroll
        for (int p = 0; p < 3; ++p)
#pragma unroll
            for (int q = 0; q < 3; ++q) Cu[p][q] = c_D[((d == p) ? d : d + p + 2) * 6 + ((c == q) ? c : c + q + 2)];   // [voigt(d,p)][voigt(c,q)]
        const AsmTile t = asmn_decode<SYM>(g, sg, tile, tpp, lane);
        const bool valid = lane < t.nvalid;
        const int k = t.kl + g.zs;
        unsigned own, colmask;
        asmn_masks<SYM>(g, nodemask, masksum, t.ln0, lane, valid, c, own, colmask);
        // the elements around the node that exist on this rank (bit a: the node is local node a)
        unsigned ex = 0;
        if (valid) {
#pragma unroll
            for (int a = 0; a < 8; ++a) {
                const int ei = t.i - node_px(a), ej = t.j - node_py(a), ek = k - node_pz(a);
                if (ei >= 0 && ei < g.NX - 1 && ej >= 0 && ej < g.NY - 1 && ek >= 0 && ek < g.NZ - 1 && ek >= er.ezs &&
                    ek < er.ezs + er.nez_ext)
                    ex |= 1u << a;
            }
        }
        double acc[NS];
#pragma unroll
        for (int s = 0; s < NS; ++s) acc[s] = 0.;
        if (__all_sync(0xffffffffu, ex == 0xffu)) {
#pragma unroll 1
            for (int gp = 0; gp < 8; ++gp) asmn_gauss_point<false, SYM, true>(gp, ex, Cu, nullptr, off, acc);
        } else {
#pragma unroll 1
            for (int gp = 0; gp < 8; ++gp) asmn_gauss_point<false, SYM, false>(gp, ex, Cu, nullptr, off, acc);
        }
        const bool rowfixed = (own >> d) & 1u, ghost_plane = SYM && t.kl < 0;
        double *At = reinterpret_cast<double *>(A) + tile * (int64_t)TILE_D + 2 * lane;
#pragma unroll
        for (int s = S0; s < 27; ++s) {
            const double val = asmn_entry<SYM>(s, acc[s - S0], wg, valid, rowfixed, colmask, d == c, ghost_plane);
            if (s == 13 && d == c && valid && !ghost_plane) dinv[d * g.S + g.G + t.ln0 + lane] = val != 0. ? 1. / val : 1.;
            const int kk = 9 * (s - S0) + e9;
            At[(kk >> 1) * (2 * TILE_NODES) + (kk & 1)] = val;
        }
        if (!SYM && e9 == 8) At[(PAIRS - 1) * (2 * TILE_NODES) + 1] = 0.;  // entry 243: padding
    }
}

// ---- tangents per Gauss point ---------------------------------------------------------------------------
template <bool SYM>
__global__ void __launch_bounds__(ASMN_THREADS, 2)       // (three CTAs per SM for the symmetric layout: 72 registers, spills in the loop, 63 ms instead of 42.6)
k_assemble_nodes_pergp(GridDev g, SymGeom sg, ElemRange er, double wg, const double *__restrict__ ctan_gp,
                       const uint8_t *__restrict__ nodemask, const uint8_t *__restrict__ masksum, double2 *__restrict__ A,
                       double *__restrict__ dinv, int64_t tile_lo, int64_t tile_hi, int64_t tpp /* tiles per plane (rounded up for the full layout) */,
                       int64_t colblock /* tiles of a plane per traversal block (tpp: plain linear order) */)
{
    constexpr int NS = SYM ? 14 : 27, S0 = SYM ? 13 : 0;
    constexpr int TILE_D = SYM ? SYM_TILE_DOUBLES : TILE_DOUBLES;
    extern __shared__ __align__(128) unsigned char smem_asmn[];
    double *stage = reinterpret_cast<double *>(smem_asmn);                 // [ASMN_NBUF][36][ASMN_CELLS]
    double *tileA = stage;                                                 // the outgoing tile re-uses the buffers
    int *cell_ie = reinterpret_cast<int *>(smem_asmn + ASMN_NBUF * ASMN_BUF_DOUBLES * 8);   // element of a cell, -1 = none
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int d = warp / 3, c = warp - 3 * d, e9 = warp;
    const int64_t per_layer = er.nex * er.ney;
    const double Cu[3][3] = {{0., 0., 0.}, {0., 0., 0.}, {0., 0., 0.}};
    const uint64_t store_policy = l2_policy(0);
    int off[3][3];
#pragma unroll
    for (int p = 0; p < 3; ++p)
#pragma unroll
        for (int q = 0; q < 3; ++q) off[p][q] = (((d == p) ? d : d + p + 2) * 6 + ((c == q) ? c : c + q + 2)) * ASMN_CELLS;

    // Traversal: column blocks of `colblock` tiles of a plane, swept through all planes before the next
    // block (an element's tangents are needed by the tiles of two rows and two planes: the second plane
    // then follows within a few MB of traffic instead of a whole plane later).
    const int64_t ntl = tile_hi - tile_lo;
    const int64_t mtot = (ntl + tpp - 1) / tpp, ncb = (tpp + colblock - 1) / colblock;
    const int64_t vper = colblock * mtot, vend = ncb * vper;
    for (int64_t v = blockIdx.x; v < vend; v += gridDim.x) {
        const int64_t cb = v / vper, rem = v - cb * vper;
        const int64_t vrow = rem / colblock;
        const int64_t col = cb * colblock + (rem - vrow * colblock);
        const int64_t tile = tile_lo + col + vrow * tpp;
        if (col >= tpp || tile >= tile_hi) continue;                       // block-uniform
        const AsmTile t = asmn_decode<SYM>(g, sg, tile, tpp, lane);
        const bool valid = lane < t.nvalid;
        // the previous tile's bulk store has to be done reading the buffers before they are refilled
        if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        __syncthreads();
        if (threadIdx.x < ASMN_CELLS) {
            // cell (row r4, x): the element whose lowest corner is the node ln0 - 1 + x - NX py - npl pz
            const int r4 = threadIdx.x / ASMN_SROW, x = threadIdx.x - r4 * ASMN_SROW;
            int ie = -1;
            {
                const int64_t lc = t.ln0 - 1 + x - (int64_t)g.NX * (1 - (r4 & 1)) - g.npl * (1 - (r4 >> 1)) + 3 * g.npl;   // > 0
                const int kc = (int)(lc / g.npl) - 3;
                const int64_t inpl = lc - (int64_t)(kc + 3) * g.npl;
                const int ej = (int)(inpl / g.NX), ei = (int)(inpl - (int64_t)ej * g.NX), ek = kc + g.zs;
                if (ei < g.NX - 1 && ej < g.NY - 1 && ek >= 0 && ek < g.NZ - 1 && ek >= er.ezs && ek < er.ezs + er.nez_ext)
                    ie = (int)((int64_t)(ek - er.ezs) * per_layer + ei + er.nex * (int64_t)ej);
            }
            cell_ie[threadIdx.x] = ie;
        }
        __syncthreads();
        // staging plan: thread -> one cell and every second entry of it (threads 0..271; consecutive threads take
        // consecutive cells = consecutive elements of a row: coalesced), so a copy costs an address increment
        const int scell = threadIdx.x % ASMN_CELLS, shalf = threadIdx.x / ASMN_CELLS;            // shalf 2: idle
        const int sie = cell_ie[scell];
        auto stage_gp = [&](int gp) {
            if (shalf < 2) {
                double *dst = stage + (gp % ASMN_NBUF) * ASMN_BUF_DOUBLES + shalf * ASMN_CELLS + scell;
                if (sie >= 0) {
                    const double *src = ctan_gp + ((int64_t)gp * 36 + shalf) * er.ne_ext + sie;
#pragma unroll
                    for (int q = 0; q < 18; ++q) cp_async8(dst + q * 2 * ASMN_CELLS, src + (int64_t)q * 2 * er.ne_ext);
                } else {
#pragma unroll
                    for (int q = 0; q < 18; ++q) dst[q * 2 * ASMN_CELLS] = 0.;
                }
            }
            cp_async_commit();
        };
        stage_gp(0);
        unsigned own, colmask;
        asmn_masks<SYM>(g, nodemask, masksum, t.ln0, lane, valid, c, own, colmask);
        double acc[NS];
#pragma unroll
        for (int s = 0; s < NS; ++s) acc[s] = 0.;
#pragma unroll 1
        for (int gp = 0; gp < 8; ++gp) {
            if (gp < 7) { stage_gp(gp + 1); asm volatile("cp.async.wait_group 1;" ::: "memory"); }
            else asm volatile("cp.async.wait_group 0;" ::: "memory");
            __syncthreads();                                             // Gauss point gp is in
            asmn_gauss_point<true, SYM, true>(gp, 0u, Cu, stage + (gp % ASMN_NBUF) * ASMN_BUF_DOUBLES + lane, off, acc);
            __syncthreads();                                             // its buffer may be refilled (or become the outgoing tile)
        }
        const bool rowfixed = (own >> d) & 1u, ghost_plane = SYM && t.kl < 0;
#pragma unroll
        for (int s = S0; s < 27; ++s) {
            const double val = asmn_entry<SYM>(s, acc[s - S0], wg, valid, rowfixed, colmask, d == c, ghost_plane);
            if (s == 13 && d == c && valid && !ghost_plane) dinv[d * g.S + g.G + t.ln0 + lane] = val != 0. ? 1. / val : 1.;
            const int kk = 9 * (s - S0) + e9;
            tileA[((kk >> 1) * TILE_NODES + lane) * 2 + (kk & 1)] = val;
        }
        if (!SYM && e9 == 8) tileA[((PAIRS - 1) * TILE_NODES + lane) * 2 + 1] = 0.;      // entry 243: padding
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy writes -> visible to the bulk copy
        __syncthreads();
        if (threadIdx.x == 0) {
            double *dstp = reinterpret_cast<double *>(A) + tile * (int64_t)TILE_D;
            // (evict-first: the finished tile must not push the tangents its neighbours still need out of L2)
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(dstp), "r"(smem_u32(tileA)),
                         "r"(TILE_D * 8), "l"(store_policy)
                         : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
    }
    if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

}  // namespace macroc
