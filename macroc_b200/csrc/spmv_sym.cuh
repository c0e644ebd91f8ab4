// spmv_sym.cuh -- symmetric storage of the assembled operator (opt-in, MACROC_OP_ASSEMBLED_SYM).
//
// A = M K M + I - M is symmetric, so block (i, s) equals the transpose of block (i + off_s, 26 - s).
// Only the 14 slots s = 13..26 (the diagonal block and the 13 neighbours at a non-negative
// linear offset) are stored: 126 entries = 63 double2 pairs per node, 1 008 B/node instead of
// 1 952.  Same tile idea as the full layout: 32 consecutive local nodes, entry
// k' = (s-13)*9 + 3r + c of node `lane` at double index ((k'>>1)*32 + lane)*2 + (k'&1) of its
// 32 256-byte tile; a tile is 7 TMA chunks of two slots each.
//
// SpMV in GATHER form (no atomics, no colouring, bit-reproducible):
//     w_i = sum_{s=13..26} A[i][s] p_{i+off_s}  +  sum_{s=14..26} A[i-off_s][s]^T p_{i-off_s}
// The first sum streams the node's own tile through the TMA ring (read from HBM once); the
// second reads the 13 lower neighbours' blocks with plain loads.  Those blocks were streamed
// moments earlier by the same CTA when the traversal keeps z- and y-neighbours close, so they
// hit L2: a CTA owns a "pencil" (one x-tile, 8 consecutive rows, one warp per row) and sweeps
// it upward in z.  The traversal only affects locality, never the result.
//
// Several ranks: the blocks of a lower neighbour that lives in the ghost plane below the slab
// (z = zs-1) are not local rows, but for a uniform tangent they are a function of node class and
// Dirichlet masks only, so the rank keeps its own copy: `front` extra tiles in front of tile 0
// hold the dz = +1 slots of the ghost plane (node ln < 0 sits in tile floor(ln / 32), lane ln & 31).
// Ghost columns / rows of x / y neighbours are ordinary local nodes and need nothing special: a
// block towards an owned node only sums elements adjacent to that owned node, all of them local.
#pragma once

#include "kernels.cuh"
#include "spmv_tma.cuh"

namespace macroc {

constexpr int SYM_ENTRIES = 126;                       // 14 slots x 9
constexpr int SYM_PAIRS = 63;
constexpr int SYM_TILE_DOUBLES = SYM_PAIRS * 2 * TILE_NODES;      // 4032 doubles = 32 256 B
constexpr int SYM_TILE_BYTES = SYM_TILE_DOUBLES * 8;
constexpr int SYM_CHUNKS = 7;                          // 7 x 9 pairs = 63

__device__ __forceinline__ double sym_entry(const double *__restrict__ A, int64_t node, int kp)
{
    const int64_t tile = node >> 5;
    const int lane = (int)(node & 31);
    return __ldg(A + tile * SYM_TILE_DOUBLES + ((int64_t)(kp >> 1) * TILE_NODES + lane) * 2 + (kp & 1));
}

// Jacobian "assembly" for a uniform tangent into the symmetric layout (cf. k_fill_operator).
__global__ void __launch_bounds__(256)
k_fill_operator_sym(GridDev g, const double *__restrict__ T, const uint8_t *__restrict__ nodemask,
                    double2 *__restrict__ A, double *__restrict__ dinv, int64_t tile_lo /* <= 0: first tile, ghost plane below */)
{
    int64_t tile = tile_lo + (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (tile >= g.ntiles) return;
    int lane = threadIdx.x & 31;
    int64_t ln = tile * TILE_NODES + lane;
    // ln < 0: node of the ghost plane below the slab (only its dz = +1 slots are ever read)
    bool valid = ln < g.nloc && ln >= -g.npl && (ln >= 0 || g.zs > 0);
    int type = 13;
    unsigned own = 0;
    if (valid) {
        const int64_t lg = ln + g.npl;                           // >= 0: index from the start of the ghost plane
        int i = (int)(lg % g.NX), j = (int)((lg / g.NX) % g.NY), k = (int)(lg / g.npl) - 1 + g.zs;
        type = node_class(i, g.NX) + 3 * node_class(j, g.NY) + 9 * node_class(k, g.NZ);
        own = nodemask[g.G + ln];
    }
    const double *Tt = T + type * 243;
    double2 *At = A + tile * (SYM_PAIRS * TILE_NODES) + lane;
    double carry = 0.;
    double diag[3] = {1., 1., 1.};
#pragma unroll
    for (int kp = 0; kp < SYM_ENTRIES; ++kp) {
        const int slot = 13 + kp / 9, rr = (kp % 9) / 3, cc = kp % 3;
        const int ddx = slot % 3 - 1, ddy = (slot / 3) % 3 - 1, ddz = slot / 9 - 1;
        double v = 0.;
        if (valid && (ln >= 0 || ddz == 1)) {
            v = __ldg(Tt + slot * 9 + 3 * rr + cc);
            unsigned nb = nodemask[g.G + ln + ddx + (int64_t)g.NX * ddy + g.npl * ddz];
            if (((own >> rr) & 1u) || ((nb >> cc) & 1u)) v = (slot == 13 && rr == cc) ? 1. : 0.;
            if (slot == 13 && rr == cc) diag[rr] = v;
        }
        if (kp & 1) At[(kp >> 1) * TILE_NODES] = make_double2(carry, v);
        else carry = v;
    }
    if (valid && ln >= 0) {
#pragma unroll
        for (int d = 0; d < 3; ++d) dinv[d * g.S + g.G + ln] = diag[d] != 0. ? 1. / diag[d] : 1.;
    }
}

// full 27-slot view of owned nodes from the symmetric storage (export / tests)
__global__ void k_export_blocks_sym(GridDev g, const double *__restrict__ A, int64_t node0, int64_t nnodes,
                                    double *__restrict__ out /* [nnodes][243] */, int64_t jmin /* first stored node (<= 0) */)
{
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nnodes * 243) return;
    const int64_t ln = owned_to_local(g, node0 + e / 243);
    const int kk = (int)(e % 243), slot = kk / 9, rr = (kk % 9) / 3, cc = kk % 3;
    double v;
    if (slot >= 13) v = sym_entry(A, ln, (slot - 13) * 9 + 3 * rr + cc);
    else {
        // block (i, s) = transpose of block (i + off_s, 26 - s), stored with node i + off_s
        const int ddx = slot % 3 - 1, ddy = (slot / 3) % 3 - 1, ddz = slot / 9 - 1;
        const int64_t j = ln + ddx + (int64_t)g.NX * ddy + g.npl * ddz;
        v = j >= jmin ? sym_entry(A, j, (26 - slot - 13) * 9 + 3 * cc + rr) : 0.;
    }
    out[e] = v;
}

// L2 eviction-priority policies: 0 normal, 1 evict_first, 2 evict_last
__device__ __forceinline__ uint64_t l2_policy(int kind)
{
    uint64_t pol;
    if (kind == 1) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    else if (kind == 2) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    else asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ double2 ldg_hint(const double2 *ptr, uint64_t pol)
{
    double2 v;
    asm("ld.global.nc.L2::cache_hint.v2.f64 {%0, %1}, [%2], %3;" : "=d"(v.x), "=d"(v.y) : "l"(ptr), "l"(pol));
    return v;
}

template <int WARPS, int NSTAGE, int MINB, bool DOT>
__global__ void __launch_bounds__(WARPS * 32, MINB)
k_spmv_sym(GridDev g, const double2 *__restrict__ A, const double *__restrict__ p, double *__restrict__ w,
           int64_t tile0, int64_t ntiles_range, int64_t tpp /* tiles per plane (rounded up) */,
           int64_t rt /* tiles per x-row (rounded up) */, int nseg, double *__restrict__ partial,
           const int *__restrict__ done, int64_t jmin /* first stored node: 0, or -(front tiles * 32) */,
           int hint /* L2 policies, 2 bits each: [1:0] z-1 gathers, [3:2] stream, [5:4] same-plane gathers */)
{
    static_assert(NSTAGE >= 2 && NSTAGE <= SYM_CHUNKS, "ring depth");
    extern __shared__ __align__(128) unsigned char smem_ring[];
    if (done && *done) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned char *ring = smem_ring + (size_t)warp * NSTAGE * CHUNK_BYTES;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem_ring + (size_t)WARPS * NSTAGE * CHUNK_BYTES) + warp * NSTAGE;
    double *red = reinterpret_cast<double *>(smem_ring + (size_t)WARPS * NSTAGE * CHUNK_BYTES + (size_t)WARPS * NSTAGE * 8);
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < NSTAGE; ++s) mbar_init(&bars[s], 1);
        fence_barrier_init();
    }
    __syncwarp();
    const unsigned char *Ab = reinterpret_cast<const unsigned char *>(A);
    const double *Ad = reinterpret_cast<const double *>(A);
    // every block is used twice: streamed by the TMA engine with its own tile, then read once
    // more as the transposed block of a neighbour.  After the second use the line is dead.
    const uint64_t pol_stream = l2_policy((hint >> 2) & 3), pol_again = l2_policy(hint & 3);
    const uint64_t pol_plane = l2_policy((hint >> 4) & 3);
    const int64_t NX = g.NX, npl = g.npl;
    const int64_t tile_end = tile0 + ntiles_range;
    // work items: pencil (x-tile xt, block of WARPS rows yb) x z-segment; column of warp = xt + rt*(yb*WARPS + warp)
    const int64_t rows = (tpp + rt - 1) / rt, yblocks = (rows + WARPS - 1) / WARPS;
    const int64_t mtot = (g.ntiles + tpp - 1) / tpp;                  // tiles per column (planes)
    const int64_t mseg = (mtot + nseg - 1) / nseg;
    const int64_t items = rt * yblocks * nseg;
    double dot = 0.;
    int64_t c = 0;                                                   // chunk counter of this warp (ring phase)
    for (int64_t item = blockIdx.x; item < items; item += gridDim.x) {
        const int64_t seg = item / (rt * yblocks), pen = item % (rt * yblocks);
        const int64_t xt = pen % rt, yb = pen / rt;
        const int64_t col = xt + rt * (yb * WARPS + warp);
        if (col >= tpp) continue;                                    // warp-uniform
        const int64_t m0 = seg * mseg, m1 = min(mtot, m0 + mseg);
        // this warp's tile sequence: col + m*tpp, m in [m0, m1), clipped to [tile0, tile_end)
        int64_t first = -1, count = 0;
        for (int64_t m = m0; m < m1; ++m) {
            const int64_t t = col + m * tpp;
            if (t >= tile0 && t < tile_end) { if (first < 0) first = m; count++; }
        }
        if (count == 0) continue;
        const int64_t nch = count * SYM_CHUNKS;
        const int64_t c0 = c;                                        // ring position at the start of the item
        auto issue = [&](int64_t qi) {                               // qi-th chunk of this item's sequence
            const int64_t tq = col + (first + qi / SYM_CHUNKS) * tpp;
            const int ch = (int)(qi % SYM_CHUNKS);
            const int stage = (int)((c0 + qi) % NSTAGE);
            mbar_arrive_expect_tx(&bars[stage], CHUNK_BYTES);
            tma_load_bulk(ring + stage * CHUNK_BYTES, Ab + tq * (int64_t)SYM_TILE_BYTES + (int64_t)ch * CHUNK_BYTES,
                          (uint32_t)CHUNK_BYTES, &bars[stage], pol_stream);
        };
        if (lane == 0)
            for (int64_t qi = 0; qi < NSTAGE && qi < nch; ++qi) issue(qi);
        int64_t q = 0;
        for (int64_t mm = 0; mm < count; ++mm) {
            const int64_t tile = col + (first + mm) * tpp;
            const int64_t ln = tile * TILE_NODES + lane;
            const double *p0 = p + g.G + ln, *p1 = p0 + g.S, *p2 = p1 + g.S;
            double a0 = 0., a1 = 0., a2 = 0., pc0 = 0., pc1 = 0., pc2 = 0.;
            // (1) transposed blocks of the 13 lower neighbours: plain loads, expected to hit L2
#pragma unroll
            for (int s = 14; s < 27; ++s) {
                const int ddx = s % 3 - 1, ddy = (s / 3) % 3 - 1, ddz = s / 9 - 1;
                const int64_t off = ddx + NX * ddy + npl * ddz;
                const int64_t j = ln - off;
                // branch-free: outside the operator (below the first tile / beyond the last) the
                // vector operand is zeroed and the block is read from a valid dummy location, so
                // the loads of all 13 slots can be in flight together
#ifndef MACROC_SYM_PROBE
                const bool okj = j >= jmin && j < g.ntiles * TILE_NODES;
#else
                // measurement build only (make EXTRA=-DMACROC_SYM_PROBE; results are wrong): hint bit 8
                // drops the gathers whose block was streamed by this CTA, bit 9 the ones streamed by
                // another CTA / an earlier z segment
                bool okj = j >= jmin && j < g.ntiles * TILE_NODES;
                if (hint & 0x300) {
                    const bool intra = (lane - ddx) >= 0 && (lane - ddx) < 32 && (warp - ddy) >= 0 && (warp - ddy) < WARPS &&
                                       (ddz == 0 || mm > 0);
                    if (((hint & 0x100) && intra) || ((hint & 0x200) && !intra)) okj = false;
                }
#endif
                const int64_t jc = okj ? j : ln;
                const double x0 = okj ? __ldg(p0 - off) : 0., x1 = okj ? __ldg(p1 - off) : 0., x2 = okj ? __ldg(p2 - off) : 0.;
                const double2 *bj = reinterpret_cast<const double2 *>(Ad) + (jc >> 5) * (SYM_PAIRS * TILE_NODES) + (jc & 31);
                const int k0s = (s - 13) * 9;
                double2 pr[5];
#ifndef MACROC_SYM_PROBE
#pragma unroll
                for (int e = 0; e < 5; ++e) pr[e] = ldg_hint(bj + ((k0s >> 1) + e) * TILE_NODES, ddz ? pol_again : pol_plane);
#else
#pragma unroll
                for (int e = 0; e < 5; ++e) pr[e] = make_double2(0., 0.);
                if (okj || !(hint & 0x300)) {
#pragma unroll
                    for (int e = 0; e < 5; ++e) pr[e] = ldg_hint(bj + ((k0s >> 1) + e) * TILE_NODES, ddz ? pol_again : pol_plane);
                }
#endif
                const double *m = reinterpret_cast<const double *>(pr) + (k0s & 1);
                // w_i[c] += sum_r A[j][s][r][c] * p_j[r]
                a0 = fma(m[0], x0, a0); a0 = fma(m[3], x1, a0); a0 = fma(m[6], x2, a0);
                a1 = fma(m[1], x0, a1); a1 = fma(m[4], x1, a1); a1 = fma(m[7], x2, a1);
                a2 = fma(m[2], x0, a2); a2 = fma(m[5], x1, a2); a2 = fma(m[8], x2, a2);
            }
            // (2) own upper blocks (slots 13..26) from the TMA ring
#pragma unroll
            for (int ch = 0; ch < SYM_CHUNKS; ++ch, ++q, ++c) {
                const int stage = (int)(c % NSTAGE);
                const uint32_t parity = (uint32_t)((c / NSTAGE) & 1);
                double xv[2][3];
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int slot = 13 + 2 * ch + h;
                    const int ddx = slot % 3 - 1, ddy = (slot / 3) % 3 - 1, ddz = slot / 9 - 1;
                    const int64_t off = ddx + NX * ddy + npl * ddz;
                    xv[h][0] = __ldg(p0 + off); xv[h][1] = __ldg(p1 + off); xv[h][2] = __ldg(p2 + off);
                    if (slot == 13) { pc0 = xv[h][0]; pc1 = xv[h][1]; pc2 = xv[h][2]; }
                }
                mbar_wait(&bars[stage], parity);
                const double2 *sv = reinterpret_cast<const double2 *>(ring + stage * CHUNK_BYTES) + lane;
                double2 v[9];
#pragma unroll
                for (int e = 0; e < 9; ++e) v[e] = sv[e * TILE_NODES];
                const double *ev = reinterpret_cast<const double *>(v);
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const double *m = ev + 9 * h;
                    const double x0 = xv[h][0], x1 = xv[h][1], x2 = xv[h][2];
                    a0 = fma(m[0], x0, a0); a0 = fma(m[1], x1, a0); a0 = fma(m[2], x2, a0);
                    a1 = fma(m[3], x0, a1); a1 = fma(m[4], x1, a1); a1 = fma(m[5], x2, a1);
                    a2 = fma(m[6], x0, a2); a2 = fma(m[7], x1, a2); a2 = fma(m[8], x2, a2);
                }
                __syncwarp();
                if (lane == 0 && q + NSTAGE < nch) issue(q + NSTAGE);
            }
            if (ln < g.nloc) {
                double *w0 = w + g.G + ln;
                w0[0] = a0; w0[g.S] = a1; w0[2 * g.S] = a2;
                if (DOT && owned_node(g, ln)) dot += a0 * pc0 + a1 * pc1 + a2 * pc2;
            }
        }
    }
    if (DOT) {
        dot = warp_sum(dot);
        if (lane == 0) red[warp] = dot;
        __syncthreads();
        if (threadIdx.x == 0) {
            double s = 0.;
#pragma unroll
            for (int qq = 0; qq < WARPS; ++qq) s += red[qq];
            partial[blockIdx.x] = s;
        }
    }
}

}  // namespace macroc
