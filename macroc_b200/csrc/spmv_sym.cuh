// spmv_sym.cuh -- symmetric storage of the assembled operator (MACROC_OP_ASSEMBLED_SYM).
//
// A = M K M + I - M is symmetric, so block (i, s) equals the transpose of block (i + off_s, 26 - s).
// Only the 14 slots s = 13..26 (the diagonal block and the 13 neighbours at a non-negative linear
// offset) are stored: 126 entries = 63 double2 pairs per node, 1 008 B/node instead of 1 952
// (PETSc's SBAIJ idea; the reference uses MATAIJ, src/init.c:92).
//
// Layout: ROW tiles.  Tile (xt, y, z) holds the nodes x = 32 xt + lane of grid row (y, z) of the
// local box (lanes with x >= NX are padding: zero blocks), tile index xt + rt*(y + NY*z), rt =
// ceil(NX/32); entry k' = (s-13)*9 + 3r + c of node `lane` at double index ((k'>>1)*32 + lane)*2 +
// (k'&1) of its 32 256-byte tile = 7 TMA chunks of two slots.  Unlike the full layout (32
// consecutive nodes in linear order) a lane keeps its x across rows and planes, which is what
// makes the sweep below possible.  With a lower z neighbour, plane z = -1 (the ghost plane) is
// stored too: only its dz = +1 slots are ever non-zero / used (rows of the lower rank, but a
// function of this rank's own element layer ezs = zs-1).
//
// SpMV: every stored block is used twice,
//     w_i           += A[i][s]   p_{i+off_s}      (row i, "gather")
//     w_{i+off_s}   += A[i][s]^T p_i              (row i+off_s, "scatter", s >= 14)
// and must leave HBM once.  A warp owns a BAND of R consecutive rows of one x-tile and sweeps it
// upward in z, streaming its tiles through a private TMA ring (cp.async.bulk + mbarrier, no
// CTA-wide synchronisation, as in spmv_tma.cuh).  Scatter contributions travel
//   * to x +- 1 by warp shuffles,
//   * to the next row of the same plane in registers (`carry`),
//   * to the next plane through a per-warp shared-memory plane of accumulators (R x 32 x 3
//     doubles; three running register rows N[3] make every cell written once and read once).
// Contributions that would cross the band (x tile edge: lanes 0 / 31; first / last row of the band)
// are instead re-computed by the RECEIVING band from the neighbour's blocks read with ordinary loads
// -- the only blocks fetched twice (~ (9/13)/R + 2/32 of the gathers).  A z segment starts with
// a scatter-only pass over the plane below it.  No atomics; the summation order is fixed, so
// results are bit-reproducible run to run.
#pragma once

#include "kernels.cuh"
#include "spmv_tma.cuh"

namespace macroc {

constexpr int SYM_ENTRIES = 126;                       // 14 slots x 9
constexpr int SYM_PAIRS = 63;
constexpr int SYM_TILE_DOUBLES = SYM_PAIRS * 2 * TILE_NODES;      // 4032 doubles = 32 256 B
constexpr int SYM_TILE_BYTES = SYM_TILE_DOUBLES * 8;
constexpr int SYM_CHUNKS = 7;                          // 7 x 9 pairs = 63
constexpr int SYM_PRE_CH0 = 2;                         // first chunk that holds a dz = +1 slot (chunk 2: slots 17, 18)

struct SymGeom {
    int rt;                 // x tiles per grid row
    int zmin;               // first stored plane: -1 with a lower z neighbour, else 0
};
__host__ __device__ __forceinline__ int64_t sym_tile_index(const GridDev &g, const SymGeom &sg, int xt, int y, int z)
{
    return xt + (int64_t)sg.rt * (y + (int64_t)g.NY * z);
}
__host__ __device__ __forceinline__ int64_t sym_tiles_per_plane(const GridDev &g, const SymGeom &sg) { return (int64_t)sg.rt * g.NY; }

// entry kp of the node at (x, y, z) of the local box (z >= zmin)
__device__ __forceinline__ double sym_entry(const GridDev &g, const SymGeom &sg, const double *__restrict__ A, int x, int y, int z, int kp)
{
    const int64_t tile = sym_tile_index(g, sg, x >> 5, y, z);
    return __ldg(A + tile * SYM_TILE_DOUBLES + ((int64_t)(kp >> 1) * TILE_NODES + (x & 31)) * 2 + (kp & 1));
}

// Jacobian "assembly" for a uniform tangent into the symmetric layout (cf. k_fill_operator).
// One warp per tile, tiles [tile_lo, ntiles): tile_lo = -tiles_per_plane when the ghost plane is stored.
__global__ void __launch_bounds__(256)
k_fill_operator_sym(GridDev g, SymGeom sg, const double *__restrict__ T, const uint8_t *__restrict__ nodemask,
                    double2 *__restrict__ A, double *__restrict__ dinv, int64_t tile_lo, int64_t tile_hi)
{
    const int64_t tile = tile_lo + (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (tile >= tile_hi) return;
    const int lane = threadIdx.x & 31;
    const int64_t tpp = sym_tiles_per_plane(g, sg);
    const int z = (int)((tile + tpp) / tpp) - 1;                  // floor division for the ghost plane
    const int64_t rem = tile - (int64_t)z * tpp;
    const int y = (int)(rem / sg.rt), x = (int)(rem % sg.rt) * 32 + lane;
    const bool valid = x < g.NX;
    const int64_t ln = x + (int64_t)g.NX * y + g.npl * z;
    int type = 13;
    unsigned own = 0;
    if (valid) {
        type = node_class(x, g.NX) + 3 * node_class(y, g.NY) + 9 * node_class(z + g.zs, g.NZ);
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
        // a neighbour outside the local box does not exist: the class stencil has a zero block there,
        // except for ghost-plane nodes, whose class is exact but whose in-plane slots are never used
        if (valid && (z >= 0 || ddz == 1)) {
            v = __ldg(Tt + slot * 9 + 3 * rr + cc);
            const unsigned nb = nodemask[g.G + ln + ddx + (int64_t)g.NX * ddy + g.npl * ddz];
            if (((own >> rr) & 1u) || ((nb >> cc) & 1u)) v = (slot == 13 && rr == cc) ? 1. : 0.;
            if (slot == 13 && rr == cc) diag[rr] = v;
        }
        if (kp & 1) At[(kp >> 1) * TILE_NODES] = make_double2(carry, v);
        else carry = v;
    }
    if (valid && z >= 0) {
#pragma unroll
        for (int d = 0; d < 3; ++d) dinv[d * g.S + g.G + ln] = diag[d] != 0. ? 1. / diag[d] : 1.;
    }
}

// full 27-slot view of owned nodes from the symmetric storage (export / tests)
__global__ void k_export_blocks_sym(GridDev g, SymGeom sg, const double *__restrict__ A, int64_t node0, int64_t nnodes,
                                    double *__restrict__ out /* [nnodes][243] */)
{
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nnodes * 243) return;
    const int64_t ln = owned_to_local(g, node0 + e / 243);
    const int x = (int)(ln % g.NX), y = (int)((ln / g.NX) % g.NY), z = (int)(ln / g.npl);
    const int kk = (int)(e % 243), slot = kk / 9, rr = (kk % 9) / 3, cc = kk % 3;
    double v;
    if (slot >= 13) v = sym_entry(g, sg, A, x, y, z, (slot - 13) * 9 + 3 * rr + cc);
    else {
        // block (i, s) = transpose of block (i + off_s, 26 - s), stored with node i + off_s
        const int ddx = slot % 3 - 1, ddy = (slot / 3) % 3 - 1, ddz = slot / 9 - 1;
        const int xj = x + ddx, yj = y + ddy, zj = z + ddz;
        const bool ok = xj >= 0 && xj < g.NX && yj >= 0 && yj < g.NY && zj >= sg.zmin;
        v = ok ? sym_entry(g, sg, A, xj, yj, zj, (26 - slot - 13) * 9 + 3 * cc + rr) : 0.;
    }
    out[e] = v;
}

// x-edge staging: per edge lane 5 neighbour blocks of 5 double2 (the 9 entries + 1) and 5 x 3 vector entries
constexpr int SYM_EDGE_SLOTS = 5;
constexpr int SYM_EDGE_DOUBLES = SYM_EDGE_SLOTS * (10 + 3) + 1;             // +1: 16-byte alignment of the next lane's area

__device__ __forceinline__ void cp_async16(void *dst_smem, const void *src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async8(void *dst_smem, const void *src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// L2 eviction-priority policies: 0 evict_first, 1 normal, 2 evict_last
__device__ __forceinline__ uint64_t l2_policy(int kind)
{
    uint64_t pol;
    if (kind == 0) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    else if (kind == 2) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    else asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ double ldg_f64_hint(const double *ptr, uint64_t pol)
{
    double v;
    asm("ld.global.nc.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(ptr), "l"(pol));
    return v;
}

template <int WARPS, int NSTAGE, int RMAX>
struct SpmvSymSmem {
    static constexpr int ring_bytes = WARPS * NSTAGE * CHUNK_BYTES;
    static constexpr int acc_bytes = WARPS * RMAX * 3 * TILE_NODES * 8;
    static constexpr int edge_bytes = WARPS * 2 * SYM_EDGE_DOUBLES * 8;      // x-edge staging of lanes 0 and 31
    static constexpr int bar_bytes = WARPS * NSTAGE * 8;
    static constexpr int red_bytes = WARPS * 8;
    static constexpr int total = ring_bytes + acc_bytes + edge_bytes + bar_bytes + red_bytes;
};

__device__ __forceinline__ double shfl_from_left(double v, int lane)      // value of lane-1, 0 for lane 0
{
    const double t = __shfl_up_sync(0xffffffffu, v, 1);
    return lane >= 1 ? t : 0.;
}
__device__ __forceinline__ double shfl_from_right(double v, int lane)     // value of lane+1, 0 for lane 31
{
    const double t = __shfl_down_sync(0xffffffffu, v, 1);
    return lane <= 30 ? t : 0.;
}

// w = A p on the owned planes [zA, zB) of the slab (+ partial of p.w).  R = rows per band (<= RMAX),
// nseg = z segments per band column.
template <int WARPS, int NSTAGE, int RMAX, bool DOT>
__global__ void __launch_bounds__(WARPS * 32, 1)
k_spmv_sym(GridDev g, SymGeom sg, const double2 *__restrict__ A, const double *__restrict__ p, double *__restrict__ w,
           int zA, int zB, int R, int nseg, double *__restrict__ partial, const int *__restrict__ done,
           int hint /* L2 policies, 2 bits each: [1:0] operator stream (0 evict_first), [3:2] vector loads (0 default) */,
           CgFuse fuse /* single rank: the last block folds the partials and updates the CG scalars */)
{
    using SM = SpmvSymSmem<WARPS, NSTAGE, RMAX>;
    extern __shared__ __align__(128) unsigned char smem_ring[];
    if (done && *done) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned char *ring = smem_ring + (size_t)warp * NSTAGE * CHUNK_BYTES;
    double *acc = reinterpret_cast<double *>(smem_ring + SM::ring_bytes) + (size_t)warp * RMAX * 3 * TILE_NODES + lane;
    // lanes 0 and 31 stage the blocks of their x neighbours (other x tile) one tile step ahead
    double *edge = reinterpret_cast<double *>(smem_ring + SM::ring_bytes + SM::acc_bytes) +
                   ((size_t)warp * 2 + (lane == 31 ? 1 : 0)) * SYM_EDGE_DOUBLES;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem_ring + SM::ring_bytes + SM::acc_bytes + SM::edge_bytes) + warp * NSTAGE;
    double *red = reinterpret_cast<double *>(smem_ring + SM::ring_bytes + SM::acc_bytes + SM::edge_bytes + SM::bar_bytes);
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < NSTAGE; ++s) mbar_init(&bars[s], 1);
        fence_barrier_init();
    }
    __syncwarp();
    const unsigned char *Ab = reinterpret_cast<const unsigned char *>(A);
    const uint64_t policy = l2_policy(hint & 3);
    const uint64_t ppol = l2_policy(((hint >> 2) & 3) == 0 ? 1 : ((hint >> 2) & 3) == 1 ? 2 : 0);
    const int64_t NX = g.NX, npl = g.npl;
    const int ybands = (g.NY + R - 1) / R;
    const int64_t items = (int64_t)sg.rt * ybands * nseg;
    const int nplanes = zB - zA, lseg = (nplanes + nseg - 1) / nseg;
    double dot = 0.;
    int64_t c = 0;                                                   // chunk counter of this warp (ring phase)
    for (int64_t item = (int64_t)blockIdx.x * WARPS + warp; item < items; item += (int64_t)gridDim.x * WARPS) {
        const int seg = (int)(item / ((int64_t)sg.rt * ybands)), pen = (int)(item % ((int64_t)sg.rt * ybands));
        const int xt = pen % sg.rt, y0 = (pen / sg.rt) * R;
        const int rows = min(R, g.NY - y0);
        const int z0 = zA + seg * lseg, z1 = min(zB, z0 + lseg);
        if (z0 >= z1) continue;                                      // warp-uniform
        const bool pre = z0 - 1 >= sg.zmin;                          // scatter-only pass over the plane below the segment
        const int zfirst = pre ? z0 - 1 : z0;
        // the scatter-only pass needs the dz = +1 slots only: chunks SYM_PRE_CH0..6 (slots 17..26)
        const int64_t nch = (int64_t)(z1 - z0) * rows * SYM_CHUNKS + (pre ? (int64_t)rows * (SYM_CHUNKS - SYM_PRE_CH0) : 0);
        const int64_t c0 = c;
        // the item's chunk sequence: tiles (xt, y0 + r, z), r fastest, 7 chunks each.  Lane 0 walks it
        // with counters (no divisions on the critical path); iq = chunks issued so far
        int64_t iq = 0;
        int i_ch = pre ? SYM_PRE_CH0 : 0, i_r = 0, i_z = zfirst;
        auto issue_next = [&]() {
            const int64_t tq = sym_tile_index(g, sg, xt, y0 + i_r, i_z);
            const int stage = (int)((c0 + iq) % NSTAGE);
            mbar_arrive_expect_tx(&bars[stage], CHUNK_BYTES);
            tma_load_bulk(ring + stage * CHUNK_BYTES, Ab + tq * (int64_t)SYM_TILE_BYTES + (int64_t)i_ch * CHUNK_BYTES,
                          (uint32_t)CHUNK_BYTES, &bars[stage], policy);
            ++iq;
            if (++i_ch == SYM_CHUNKS) {
                if (++i_r == rows) { i_r = 0; ++i_z; }
                i_ch = i_z < z0 ? SYM_PRE_CH0 : 0;
            }
        };
        if (lane == 0)
            for (int qi = 0; qi < NSTAGE && qi < nch; ++qi) issue_next();
        // the accumulator plane starts empty (also orders it after the previous item's last reads)
        for (int r = 0; r < rows; ++r) { acc[(r * 3 + 0) * TILE_NODES] = 0.; acc[(r * 3 + 1) * TILE_NODES] = 0.; acc[(r * 3 + 2) * TILE_NODES] = 0.; }
        const int x = xt * 32 + lane;
        const bool xvalid = x < g.NX;
        // lanes whose x neighbour lives in another x tile: lane 0 looks left (slots with ddx = +1),
        // lane 31 looks right (ddx = -1)
        const bool edge_lane = (lane == 0 && x > 0) || (lane == 31 && x + 1 < g.NX);
        const int exj = lane == 0 ? x - 1 : x + 1;
        auto edge_slot = [&](int e, int yy, int zz, int &sl, int &yj, int &zj) -> bool {
            // lane 0: slots 14 17 20 23 26; lane 31: slots 15 18 21 24
            sl = lane == 0 ? 14 + 3 * e : 15 + 3 * e;
            if (sl > 26) return false;
            yj = yy - ((sl / 3) % 3 - 1); zj = zz - (sl / 9 - 1);
            return yj >= 0 && yj < g.NY && zj >= sg.zmin;
        };
        auto edge_stage = [&](int yy, int zz) {
#pragma unroll
            for (int e = 0; e < SYM_EDGE_SLOTS; ++e) {
                int sl, yj, zj;
                if (edge_slot(e, yy, zz, sl, yj, zj)) {
                    const double2 *bj = A + sym_tile_index(g, sg, exj >> 5, yj, zj) * (SYM_PAIRS * TILE_NODES) + (exj & 31);
                    const int k0s = (sl - 13) * 9;
#pragma unroll
                    for (int qq = 0; qq < 5; ++qq) cp_async16(edge + e * 10 + 2 * qq, bj + ((k0s >> 1) + qq) * TILE_NODES);
                    const double *pj = p + g.G + exj + NX * yj + npl * zj;
#pragma unroll
                    for (int d = 0; d < 3; ++d) cp_async8(edge + 10 * SYM_EDGE_SLOTS + 3 * e + d, pj + d * g.S);
                }
            }
            cp_async_commit();
        };
        auto edge_apply = [&](int yy, int zz, double &a0, double &a1, double &a2) {
#pragma unroll
            for (int e = 0; e < SYM_EDGE_SLOTS; ++e) {
                int sl, yj, zj;
                if (edge_slot(e, yy, zz, sl, yj, zj)) {
                    const double *m = edge + e * 10 + (((sl - 13) * 9) & 1);
                    const double x0 = edge[10 * SYM_EDGE_SLOTS + 3 * e], x1 = edge[10 * SYM_EDGE_SLOTS + 3 * e + 1],
                                 x2 = edge[10 * SYM_EDGE_SLOTS + 3 * e + 2];
                    // w_i[c] += sum_r A[j][s][r][c] * p_j[r]
                    a0 = fma(m[0], x0, a0); a0 = fma(m[3], x1, a0); a0 = fma(m[6], x2, a0);
                    a1 = fma(m[1], x0, a1); a1 = fma(m[4], x1, a1); a1 = fma(m[7], x2, a1);
                    a2 = fma(m[2], x0, a2); a2 = fma(m[5], x1, a2); a2 = fma(m[8], x2, a2);
                }
            }
        };
        // slots sA, sA+1, sA+2 (ddx = -1, 0, +1; common ddy, ddz) of the row above / below the band,
        // sources inside this x tile only (the other lanes are served by the x-edge staging)
        auto gather3 = [&](int sA, int yy, int zz, double &a0, double &a1, double &a2) {
            const int ddy = (sA / 3) % 3 - 1, ddz = sA / 9 - 1;
            const int yj = yy - ddy, zj = zz - ddz;
            if (yj < 0 || yj >= g.NY || zj < sg.zmin) return;       // warp-uniform
            const double2 *bt = A + sym_tile_index(g, sg, xt, yj, zj) * (SYM_PAIRS * TILE_NODES);
            const double *pt = p + g.G + (int64_t)xt * 32 + NX * yj + npl * zj;
            double2 pr[3][5];
            double xx[3][3];
            bool ok[3];
#pragma unroll
            for (int u = 0; u < 3; ++u) {
                const int sl = lane - (u - 1);                       // source lane
                ok[u] = xvalid && sl >= 0 && sl <= 31;
                const int slc = ok[u] ? sl : lane;
                const int k0s = (sA + u - 13) * 9;
#pragma unroll
                for (int e = 0; e < 5; ++e) pr[u][e] = __ldg(bt + ((k0s >> 1) + e) * TILE_NODES + slc);
#pragma unroll
                for (int d = 0; d < 3; ++d) xx[u][d] = ok[u] ? __ldg(pt + d * g.S + slc) : 0.;
            }
#pragma unroll
            for (int u = 0; u < 3; ++u) {
                const double *m = reinterpret_cast<const double *>(pr[u]) + (((sA + u - 13) * 9) & 1);
                a0 = fma(m[0], xx[u][0], a0); a0 = fma(m[3], xx[u][1], a0); a0 = fma(m[6], xx[u][2], a0);
                a1 = fma(m[1], xx[u][0], a1); a1 = fma(m[4], xx[u][1], a1); a1 = fma(m[7], xx[u][2], a1);
                a2 = fma(m[2], xx[u][0], a2); a2 = fma(m[5], xx[u][1], a2); a2 = fma(m[8], xx[u][2], a2);
            }
        };
        // the first step of the item gathers already (no scatter-only pass below): stage it now
        if (!pre && edge_lane) edge_stage(y0, z0);
        int64_t q = 0;
        // vector operands are fetched well ahead of their use (the mbarrier wait is a scheduling fence,
        // loads issued after it would expose their L2 latency): the operands of the first two chunks of
        // a step are loaded at the end of the previous step, the rest two chunks ahead.
        // load_xv(q0, ch) reads the operands of chunk ch's two slots of the node whose p is at q0.
        auto load_xv = [&](const double *q0, int ch, double (&xq)[2][3]) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int slot = 13 + 2 * ch + h;
                const int ddx = slot % 3 - 1, ddy = (slot / 3) % 3 - 1, ddz = slot / 9 - 1;
                const int64_t off = ddx + NX * ddy + npl * ddz;
                xq[h][0] = ldg_f64_hint(q0 + off, ppol); xq[h][1] = ldg_f64_hint(q0 + g.S + off, ppol); xq[h][2] = ldg_f64_hint(q0 + 2 * g.S + off, ppol);
            }
        };
        double xa[2][3], xb[2][3];                                   // operands of the next two chunks to be processed
        {
            const double *q0 = p + g.G + x + NX * y0 + npl * (int64_t)zfirst;
            if (pre) { load_xv(q0, SYM_PRE_CH0, xa); load_xv(q0, SYM_PRE_CH0 + 1, xb); }
            else { load_xv(q0, 0, xa); load_xv(q0, 1, xb); }
        }
        for (int z = zfirst; z < z1; ++z) {
            const bool scatter_only = z < z0;
            double carry0 = 0., carry1 = 0., carry2 = 0.;            // to the next row of this plane
            double N[3][3];                                          // to rows r-1, r, r+1 of the next plane
#pragma unroll
            for (int a = 0; a < 3; ++a) N[a][0] = N[a][1] = N[a][2] = 0.;
            for (int r = 0; r < rows; ++r) {
                const int y = y0 + r;
                const int64_t ln = x + NX * y + npl * z;
                const double *p0 = p + g.G + ln, *p1 = p0 + g.S, *p2 = p1 + g.S;
                // the next step (row r+1 of this plane, or row 0 of the next plane) and where it starts
                const bool last_row = r + 1 == rows;
                const int zn = last_row ? z + 1 : z, yn = last_row ? y0 : y + 1;
                const bool has_next = zn < z1;
                const double *pn = p + g.G + x + NX * yn + npl * (int64_t)zn;
                const int chn0 = zn < z0 ? SYM_PRE_CH0 : 0;          // first chunk of the next step
                // the node's own p is slot 13's operand (chunk 0); a scatter-only step does not stream chunk 0
                double pc0, pc1, pc2;
                if (scatter_only) { pc0 = ldg_f64_hint(p0, ppol); pc1 = ldg_f64_hint(p1, ppol); pc2 = ldg_f64_hint(p2, ppol); }
                else { pc0 = xa[0][0]; pc1 = xa[0][1]; pc2 = xa[0][2]; }
                // what the plane below and the previous row of this plane scattered to this node
                double a0 = acc[(r * 3 + 0) * TILE_NODES] + carry0, a1 = acc[(r * 3 + 1) * TILE_NODES] + carry1,
                       a2 = acc[(r * 3 + 2) * TILE_NODES] + carry2;
                // (1a) x tile edge (lanes 0 / 31): the neighbour's blocks were staged one tile step ahead
                if (!scatter_only && edge_lane) {
                    cp_async_wait_all();
                    edge_apply(y, z, a0, a1, a2);
                }
                if (edge_lane && has_next && zn >= z0) edge_stage(yn, zn);      // stage the next step's
                // (1b) first / last row of the band: the rows above / below belong to another band.  Three
                // slots (ddx = -1, 0, +1) per group, all loads of a group in flight together
                if (!scatter_only) {
                    if (r == 0) { gather3(15, y, z, a0, a1, a2); gather3(24, y, z, a0, a1, a2); }
                    if (r == rows - 1) gather3(18, y, z, a0, a1, a2);
                }
                // (2) own blocks (slots 13..26) from the TMA ring: row i, and the transposed use for row i + off
                double nc0 = 0., nc1 = 0., nc2 = 0.;                 // carry for the next row
#pragma unroll
                for (int ch = 0; ch < SYM_CHUNKS; ++ch) {
                    if (scatter_only && ch < SYM_PRE_CH0) continue;      // warp-uniform: these chunks were not streamed
                    const int stage = (int)(c % NSTAGE);
                    const uint32_t parity = (uint32_t)((c / NSTAGE) & 1);
                    ++q; ++c;
                    double xv[2][3];
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        xv[h][0] = xa[h][0]; xv[h][1] = xa[h][1]; xv[h][2] = xa[h][2];
                        xa[h][0] = xb[h][0]; xa[h][1] = xb[h][1]; xa[h][2] = xb[h][2];
                    }
                    // two chunks ahead: still this step, or the first chunks of the next one
                    if (ch + 2 < SYM_CHUNKS) load_xv(p0, ch + 2, xb);
                    else if (has_next) {
                        if (chn0 == 0) load_xv(pn, ch + 2 - SYM_CHUNKS, xb); else load_xv(pn, ch + 2 - SYM_CHUNKS + SYM_PRE_CH0, xb);
                    }
                    mbar_wait(&bars[stage], parity);
                    const double2 *sv = reinterpret_cast<const double2 *>(ring + stage * CHUNK_BYTES) + lane;
                    double2 v[9];
#pragma unroll
                    for (int e = 0; e < 9; ++e) v[e] = sv[e * TILE_NODES];
                    const double *ev = reinterpret_cast<const double *>(v);
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int slot = 13 + 2 * ch + h;
                        const int ddx = slot % 3 - 1, ddy = (slot / 3) % 3 - 1, ddz = slot / 9 - 1;
                        const double *m = ev + 9 * h;
                        const double x0 = xv[h][0], x1 = xv[h][1], x2 = xv[h][2];
                        a0 = fma(m[0], x0, a0); a0 = fma(m[1], x1, a0); a0 = fma(m[2], x2, a0);
                        a1 = fma(m[3], x0, a1); a1 = fma(m[4], x1, a1); a1 = fma(m[5], x2, a1);
                        a2 = fma(m[6], x0, a2); a2 = fma(m[7], x1, a2); a2 = fma(m[8], x2, a2);
                        if (slot >= 14) {
                            // t = A[i][s]^T p_i belongs to node i + off_s: lane + ddx, row r + ddy, plane z + ddz
                            double t0 = fma(m[6], pc2, fma(m[3], pc1, m[0] * pc0));
                            double t1 = fma(m[7], pc2, fma(m[4], pc1, m[1] * pc0));
                            double t2 = fma(m[8], pc2, fma(m[5], pc1, m[2] * pc0));
                            if (ddx == 1) { t0 = shfl_from_left(t0, lane); t1 = shfl_from_left(t1, lane); t2 = shfl_from_left(t2, lane); }
                            if (ddx == -1) { t0 = shfl_from_right(t0, lane); t1 = shfl_from_right(t1, lane); t2 = shfl_from_right(t2, lane); }
                            if (ddz == 0 && ddy == 0) { a0 += t0; a1 += t1; a2 += t2; }
                            else if (ddz == 0) { nc0 += t0; nc1 += t1; nc2 += t2; }
                            else { N[ddy + 1][0] += t0; N[ddy + 1][1] += t1; N[ddy + 1][2] += t2; }
                        }
                    }
                    __syncwarp();
                    if (lane == 0 && iq < nch) issue_next();
                }
                carry0 = nc0; carry1 = nc1; carry2 = nc2;
                if (!scatter_only && xvalid) {
                    double *w0 = w + g.G + ln;
                    w0[0] = a0; w0[g.S] = a1; w0[2 * g.S] = a2;
                    if (DOT && owned_node(g, ln)) dot += a0 * pc0 + a1 * pc1 + a2 * pc2;
                }
                // row r-1 of the next plane is complete: park it where this plane's row r-1 was (already consumed)
                if (r >= 1) { acc[((r - 1) * 3 + 0) * TILE_NODES] = N[0][0]; acc[((r - 1) * 3 + 1) * TILE_NODES] = N[0][1]; acc[((r - 1) * 3 + 2) * TILE_NODES] = N[0][2]; }
#pragma unroll
                for (int d = 0; d < 3; ++d) { N[0][d] = N[1][d]; N[1][d] = N[2][d]; N[2][d] = 0.; }
            }
            acc[((rows - 1) * 3 + 0) * TILE_NODES] = N[0][0]; acc[((rows - 1) * 3 + 1) * TILE_NODES] = N[0][1]; acc[((rows - 1) * 3 + 2) * TILE_NODES] = N[0][2];
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
        if (fuse.ticket) cg_last_block<WARPS, 1>(partial, gridDim.x, fuse, red);
    }
}

}  // namespace macroc
