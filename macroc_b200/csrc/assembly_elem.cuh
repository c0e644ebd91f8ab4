// assembly_elem.cuh -- per-element residual and Jacobian kernels.
//
// These are the general form of the reference's element loops
// (src/assembly.c:85-108 and :142-162): every Gauss point has its own stress
// (6 doubles) and tangent (36 doubles, row-major) -- the reference's
// gpi = ie*8 + gp data (assembly.c:58,91,148), kept on the device as SoA over
// elements (see k_strain_stress) -- the arrays a constitutive plug-in in
// MicroPP's role fills on the device.  With the homogenised linear law they are filled by
// k_homogenize_linear (stress = D strain, ctan = D); with MACROC_MAT_UNIFORM the
// same kernels take D from constant memory and never touch the arrays.
//
// Scatter without colouring or atomics:
//  * residual: pass 1 writes the 24 element forces to a scratch array that is
//    SoA over elements (coalesced), pass 2 lets every owned node add its <= 8
//    contributions in increasing element order;
//  * Jacobian: one CTA per 32-node operator tile, warp a = local node a of the
//    element, lane = node.  Thread (a, lane) integrates the 3x24 row block of
//    "its" element (the one in which the node is local node a) in registers --
//    no flop is done twice -- then the 8 warps add their blocks into the tile in
//    shared memory in 8 conflict-free rounds (round b: block column b; within a
//    round distinct a hit distinct stencil slots), the Dirichlet mask is applied
//    and the tile leaves as one contiguous 62 KB store.
#pragma once

#include "kernels.cuh"
#include "spmv_sym.cuh"

namespace macroc {

struct ElemRange {
    int ezs;          // first element layer stored (global z index)
    int nez_ext;      // stored layers: owned ones plus the upper neighbour's first layer
    int64_t nex, ney; // elements per row / rows per layer
    int64_t ne_ext;   // stored elements = pitch of the SoA Gauss-point arrays
};

// MicroPP stand-in on the device: sigma = D eps, C = D for every Gauss point of the DMDA-owned
// elements only (ex < onex, ey < oney), like a real material model; the upper neighbours'
// layers arrive through the Gauss-point halo.  SoA arrays, pitch ne_ext.
__global__ void k_homogenize_linear(int64_t ne, int64_t lnex, int64_t lney, int64_t onex, int64_t oney, int64_t ne_ext,
                                    const double *__restrict__ strain, double *__restrict__ stress,
                                    double *__restrict__ ctan)
{
    int64_t ie = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (ie >= ne) return;
    if (ie % lnex >= onex || (ie / lnex) % lney >= oney) return;
#pragma unroll 1
    for (int gp = 0; gp < 8; ++gp) {
        double e[6];
#pragma unroll
        for (int j = 0; j < 6; ++j) e[j] = strain[(gp * 6 + j) * ne_ext + ie];
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            double s = 0.;
#pragma unroll
            for (int j = 0; j < 6; ++j) s = fma(c_D[i * 6 + j], e[j], s);
            stress[(gp * 6 + i) * ne_ext + ie] = s;
        }
#pragma unroll
        for (int q = 0; q < 36; ++q) ctan[(gp * 36 + q) * ne_ext + ie] = c_D[q];
    }
}

// pass 1: be[24] of every stored element in layers [l0, l0+nl) of the range -> scratch[q][e_local]
template <bool PER_GP>
__global__ void __launch_bounds__(128)
k_elem_forces(GridDev g, ElemRange er, int l0, int nl, double wg, const double *__restrict__ u,
              const double *__restrict__ stress_gp, double *__restrict__ scratch)
{
    const int64_t per_layer = er.nex * er.ney, n = per_layer * nl;
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    int ei = (int)(e % er.nex), ej = (int)((e / er.nex) % er.ney), el = (int)(e / per_layer) + l0;
    double be[24];
#pragma unroll
    for (int q = 0; q < 24; ++q) be[q] = 0.;
    double ue[8][3];
    if (!PER_GP) gather_element(u, g, g.G + ei + (int64_t)g.NX * ej + g.npl * (er.ezs + el - g.zs), ue);
    const double *sg = PER_GP ? stress_gp + ((int64_t)el * per_layer + (e % per_layer)) : nullptr;
#pragma unroll 1
    for (int gp = 0; gp < 8; ++gp) {
        double sig[6];
        if (PER_GP) {
#pragma unroll
            for (int q = 0; q < 6; ++q) sig[q] = __ldg(sg + (gp * 6 + q) * er.ne_ext);
        } else {
            double eps[6];
            element_strain(ue, gp, eps);
            stress_of(eps, sig);
        }
#pragma unroll
        for (int a = 0; a < 8; ++a) {
            const double hx = c_dsh[gp][a][0], hy = c_dsh[gp][a][1], hz = c_dsh[gp][a][2];
            // be[i] += B[j][i]*stress[j]*wg, j ascending (assembly.c:151-153)
            be[3 * a + 0] += hx * sig[0] * wg; be[3 * a + 0] += hy * sig[3] * wg; be[3 * a + 0] += hz * sig[4] * wg;
            be[3 * a + 1] += hy * sig[1] * wg; be[3 * a + 1] += hx * sig[3] * wg; be[3 * a + 1] += hz * sig[5] * wg;
            be[3 * a + 2] += hz * sig[2] * wg; be[3 * a + 2] += hx * sig[4] * wg; be[3 * a + 2] += hy * sig[5] * wg;
        }
    }
#pragma unroll
    for (int q = 0; q < 24; ++q) scratch[q * n + e] = be[q];
}

// pass 2: nodes of planes [k0, k0+nk) (slab-local) add their element forces, apply the
// Dirichlet mask and the sign (bcs.c:350-362, assembly.c:173) and accumulate |b|^2.
__global__ void __launch_bounds__(256)
k_gather_forces(GridDev g, ElemRange er, int l0, int nl, int k0, int nk, const double *__restrict__ scratch,
                const uint8_t *__restrict__ nodemask, double *__restrict__ b, double *__restrict__ partial)
{
    __shared__ double sm[8];
    const int64_t per_layer = er.nex * er.ney, n = per_layer * nl;
    int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    double sq = 0.;
    if (q < g.npl * nk) {
        int64_t ln = (int64_t)k0 * g.npl + q;
        int i = (int)(ln % g.NX), j = (int)((ln / g.NX) % g.NY), k = (int)(ln / g.npl) + g.zs;
        double r0 = 0., r1 = 0., r2 = 0.;
        for (int oz = -1; oz <= 0; ++oz)
            for (int oy = -1; oy <= 0; ++oy)
                for (int ox = -1; ox <= 0; ++ox) {
                    int ei = i + ox, ej = j + oy, ek = k + oz;
                    if (ei < 0 || ei >= g.NX - 1 || ej < 0 || ej >= g.NY - 1 || ek < 0 || ek >= g.NZ - 1) continue;
                    int el = ek - er.ezs - l0;                  // layer inside the scratch chunk
                    if (el < 0 || el >= nl) continue;           // (cannot happen for a correct chunking)
                    int a = local_node_of_pos(-ox, -oy, -oz);
                    int64_t e = ei + er.nex * (ej + er.ney * (int64_t)el);
                    r0 += scratch[(3 * a + 0) * n + e];
                    r1 += scratch[(3 * a + 1) * n + e];
                    r2 += scratch[(3 * a + 2) * n + e];
                }
        unsigned own = nodemask[g.G + ln];
        r0 = (own & 1u) ? 0. : -r0;
        r1 = (own & 2u) ? 0. : -r1;
        r2 = (own & 4u) ? 0. : -r2;
        double *b0 = b + g.G + ln;
        b0[0] = r0; b0[g.S] = r1; b0[2 * g.S] = r2;
        if (owned_node(g, ln)) sq = r0 * r0 + r1 * r1 + r2 * r2;
    }
    double s = block_sum<8>(sq, sm);
    if (threadIdx.x == 0) partial[blockIdx.x] = s;
}

// kk = slot*9 + 3*row + col  ->  slot | row << 5 | col << 7 | diagonal << 9   (entry 243 is padding)
struct KkInfo {
    unsigned short v[244];
};
constexpr KkInfo make_kkinfo()
{
    KkInfo t{};
    for (int kk = 0; kk < 243; ++kk) {
        const int slot = kk / 9, rr = (kk % 9) / 3, cc = kk % 3;
        t.v[kk] = (unsigned short)(slot | (rr << 5) | (cc << 7) | ((slot == 13 && rr == cc) ? 1 << 9 : 0));
    }
    t.v[243] = 0x8000;
    return t;
}
__constant__ KkInfo c_kkinfo = make_kkinfo();

// General Jacobian assembly (assembly.c:85-108 with a tangent per Gauss point), one CTA of 8 warps
// per operator tile: warp a = local node a of the element, lane = node.  SYM: the tile is a row
// tile of the symmetric layout (spmv_sym.cuh) and only the slots 13..26 are stored -- the tangent
// must then be symmetric (Ke is; the reference never relies on it, MATAIJ stores both halves).
// The Gauss-point loop is fully unrolled so that every shape-function derivative of the block
// columns is an immediate constant-bank operand of its DFMA; wg is applied once, when a thread adds
// its 3 x 24 row block to the tile.
constexpr int ASM_WARPS = 24;                                             // 8 local nodes x 3 block rows
constexpr int ASM_THREADS = ASM_WARPS * 32;
constexpr int ASM_STAGE_OFFSET = TILE_DOUBLES * 8 + 27 * 32 + 32;          // 16-byte aligned, behind the tile and the masks
constexpr int ASM_SMEM_UNIFORM = TILE_DOUBLES * 8 + 27 * 32;
constexpr int ASM_SMEM_PER_GP = ASM_STAGE_OFFSET + 2 * 36 * 256 * 8;       // + two staging buffers [36][8 nodes a][32 lanes]

// Row D of the 3 x 24 row block (B_a^T C B) of one element, one Gauss point:
//   T[k]      = sum_r B_a[r][D] C[r][k]            -- B_a's column D has three non-zeros
//   blk[3b+c] += T[k] B_b[k][c]                     -- B_b's row k has one or two non-zeros per node b
// 18 + 72 FMA; the three rows of a block are three different threads, so no product is computed twice
// and a thread holds 24 accumulators instead of 72.  ck: the tangent, C[r][k] at ck[(6 r + k) * cstride].
template <bool PER_GP, int D, int GP>
__device__ __forceinline__ void integrate_gp(int gp, int a, const double *__restrict__ ck, int cstride, double (&blk)[24])
{
    // rows of C that meet column D of B_a, and the shape-function derivative that multiplies each:
    // D=0: (0,hx) (3,hy) (4,hz)   D=1: (1,hy) (3,hx) (5,hz)   D=2: (2,hz) (4,hx) (5,hy)
    constexpr int R0 = D, R1 = D == 2 ? 4 : 3, R2 = D == 0 ? 4 : 5;
    constexpr int I0 = D, I1 = D == 0 ? 1 : 0, I2 = D == 2 ? 1 : 2;
    const int g = GP >= 0 ? GP : gp;               // GP >= 0: compile-time Gauss point (every B_b entry an immediate)
    const double h0 = c_dsh[g][a][I0], h1 = c_dsh[g][a][I1], h2 = c_dsh[g][a][I2];
#pragma unroll
    for (int kc = 0; kc < 6; ++kc) {
        const double c0 = PER_GP ? ck[(R0 * 6 + kc) * cstride] : c_D[R0 * 6 + kc];
        const double c1 = PER_GP ? ck[(R1 * 6 + kc) * cstride] : c_D[R1 * 6 + kc];
        const double c2 = PER_GP ? ck[(R2 * 6 + kc) * cstride] : c_D[R2 * 6 + kc];
        const double T = fma(h2, c2, fma(h1, c1, h0 * c0));
#pragma unroll
        for (int b = 0; b < 8; ++b) {
            const double bx = c_dsh[g][b][0], by = c_dsh[g][b][1], bz = c_dsh[g][b][2];
            // row kc of B_b: (0: c0=bx) (1: c1=by) (2: c2=bz) (3: c0=by, c1=bx) (4: c0=bz, c2=bx) (5: c1=bz, c2=by)
            if (kc == 0) blk[3 * b + 0] = fma(T, bx, blk[3 * b + 0]);
            if (kc == 1) blk[3 * b + 1] = fma(T, by, blk[3 * b + 1]);
            if (kc == 2) blk[3 * b + 2] = fma(T, bz, blk[3 * b + 2]);
            if (kc == 3) { blk[3 * b + 0] = fma(T, by, blk[3 * b + 0]); blk[3 * b + 1] = fma(T, bx, blk[3 * b + 1]); }
            if (kc == 4) { blk[3 * b + 0] = fma(T, bz, blk[3 * b + 0]); blk[3 * b + 2] = fma(T, bx, blk[3 * b + 2]); }
            if (kc == 5) { blk[3 * b + 1] = fma(T, bz, blk[3 * b + 1]); blk[3 * b + 2] = fma(T, by, blk[3 * b + 2]); }
        }
    }
}

// uniform tangent: all eight Gauss points unrolled
template <int D>
__device__ __forceinline__ void integrate_uniform(int a, double (&blk)[24])
{
    integrate_gp<false, D, 0>(0, a, nullptr, 0, blk); integrate_gp<false, D, 1>(1, a, nullptr, 0, blk);
    integrate_gp<false, D, 2>(2, a, nullptr, 0, blk); integrate_gp<false, D, 3>(3, a, nullptr, 0, blk);
    integrate_gp<false, D, 4>(4, a, nullptr, 0, blk); integrate_gp<false, D, 5>(5, a, nullptr, 0, blk);
    integrate_gp<false, D, 6>(6, a, nullptr, 0, blk); integrate_gp<false, D, 7>(7, a, nullptr, 0, blk);
}

// General Jacobian assembly (assembly.c:85-108 with a tangent per Gauss point), one CTA of 24 warps per
// operator tile: warp (a, d) = local node a of the element x row d of its 3 x 24 row block, lane = node.
// Thread (a, d, lane) integrates row d of the block of the element in which its node is local node a
// (24 accumulators, 720 FMA), the 24 warps add their rows into the tile in shared memory in 8
// conflict-free rounds (round b: block column b; distinct (a, d) hit distinct entries), the Dirichlet
// mask is applied -- tiles with no Dirichlet dof in reach, the vast majority, skip it -- and the tile
// leaves as one contiguous store.  Per-Gauss-point tangents are staged through shared memory one Gauss
// point ahead (cp.async; the three row threads of a node share one copy, two barriers per Gauss point):
// 8 memory round trips per tile, all of them overlapped, instead of 48 exposed ones.
// SYM: the tile is a row tile of the symmetric layout (spmv_sym.cuh) and only the slots 13..26 are
// stored -- the tangent must then be symmetric (Ke is; the reference never relies on it, MATAIJ stores
// both halves).  wg is applied once, when a thread adds its row to the tile.
// Measured and rejected (256^3, uniform / per-GP tangents): 8 warps x 72 accumulators (42 / 63 ms; 2 warps
// per scheduler cannot keep the FP64 pipe busy), this kernel with the 24 warps split into two co-resident
// 12-warp CTAs that take two local nodes each (33.8 / 72.7 ms: no better, 3 integrating warps per
// scheduler do not saturate the pipe either), grids larger than the resident CTAs (the traversal's locality
// is lost: 75 GB of tangent reads instead of ~40).
template <bool PER_GP, bool SYM>
__global__ void __launch_bounds__(ASM_THREADS, 1)
k_assemble_elements(GridDev g, SymGeom sg, ElemRange er, double wg, const double *__restrict__ ctan_gp,
                    const uint8_t *__restrict__ nodemask, double2 *__restrict__ A, double *__restrict__ dinv,
                    int64_t tile_lo, int64_t tile_hi, int64_t tpp /* tiles per plane (rounded up for the full layout) */,
                    int64_t colblock /* tiles of a plane per traversal block (tpp: plain linear order) */, int stream_stores)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // staging tile: entry kk = slot*9 + 3 r + c of node `lane` at tileA[kk*32 + lane] -- lanes 8 bytes apart, so
    // the read-modify-write rounds are bank-conflict free (the operator's own pair-interleaved layout is
    // produced by the final pass)
    double *tileA = reinterpret_cast<double *>(smem_raw);                  // 244 x 32 doubles = TILE_DOUBLES
    uint8_t *nbmask = smem_raw + TILE_DOUBLES * sizeof(double);            // [27][32]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int a = warp / 3, d = warp - 3 * a;
    const int apx = node_px(a), apy = node_py(a), apz = node_pz(a);
    const int64_t per_layer = er.nex * er.ney;
    // staging of the tangents (PER_GP): [2 buffers][36 entries][a][lane]; thread (a, d, lane) copies the
    // entries 12 d .. 12 d + 11 of its element and reads the 18 its row needs after the barrier
    double *stage = reinterpret_cast<double *>(smem_raw + ASM_STAGE_OFFSET) + a * 32 + lane;
    // where this thread's three cells of block column b live in the tile: entry kk = slot(a -> b) * 9 + 3 d + c
    int cell0[8];
#pragma unroll
    for (int b = 0; b < 8; ++b) cell0[b] = ((node_pz(b) - apz + 1) * 9 + (node_py(b) - apy + 1) * 3 + (node_px(b) - apx + 1)) * 9 + 3 * d;

    // Traversal: column blocks of `colblock` tiles of a plane, swept through all planes before the next
    // block (an element's tangents are needed by the tiles of two rows and two planes: the second plane
    // then follows within a few MB of traffic instead of a whole plane later).
    const int64_t ntl = tile_hi - tile_lo;
    const int64_t mtot = (ntl + tpp - 1) / tpp, ncb = (tpp + colblock - 1) / colblock;
    for (int64_t v = blockIdx.x; v < ncb * colblock * mtot; v += gridDim.x) {
        const int64_t cb = v / (colblock * mtot), rem = v - cb * (colblock * mtot);
        const int64_t col = cb * colblock + rem % colblock;
        const int64_t tile = tile_lo + col + (rem / colblock) * tpp;
        if (col >= tpp || tile >= tile_hi) continue;                       // block-uniform
        // node of (tile, lane): local box coordinates (i, j), slab-local plane kl, linear index ln0 of lane 0
        int i = 0, j = 0, kl = 0, nvalid = 0;
        int64_t ln0;
        if (SYM) {
            kl = (int)((tile + tpp) / tpp) - 1;                            // floor: the ghost plane is -1
            const int64_t rem2 = tile - (int64_t)kl * tpp;
            j = (int)(rem2 / sg.rt);
            const int x0 = (int)(rem2 % sg.rt) * 32;
            i = x0 + lane;
            nvalid = min(32, g.NX - x0);
            ln0 = x0 + (int64_t)g.NX * j + g.npl * kl;
        } else {
            ln0 = tile * TILE_NODES;
            const int64_t ln = ln0 + lane;
            nvalid = (int)min((int64_t)32, g.nloc - ln0);
            // (a rank's local node count fits 31 bits: 32-bit divisions instead of 64-bit calls)
            if (lane < nvalid) {
                const unsigned lnu = (unsigned)ln, nx = (unsigned)g.NX, npl = (unsigned)g.npl;
                kl = (int)(lnu / npl);
                const unsigned inpl = lnu - (unsigned)kl * npl;
                j = (int)(inpl / nx); i = (int)(inpl - (unsigned)j * nx);
            }
        }
        const bool valid = lane < nvalid;
        const int k = kl + g.zs;
        // the element in which this node is local node a (it must be one whose tangents this rank holds)
        const int ei = i - apx, ej = j - apy, ek = k - apz;
        const bool exists = valid && ei >= 0 && ei < g.NX - 1 && ej >= 0 && ej < g.NY - 1 && ek >= 0 && ek < g.NZ - 1 &&
                            ek >= er.ezs && ek < er.ezs + er.nez_ext;
        const double *cg = (PER_GP && exists) ? ctan_gp + ((int64_t)(ek - er.ezs) * per_layer + ei + er.nex * (int64_t)ej) : nullptr;
        auto stage_gp = [&](int gp) {               // this thread's third of the element's 36 entries of Gauss point gp
            if (exists) {
                double *dst = stage + (gp & 1) * (36 * 256) + (12 * d) * 256;
                const double *src = cg + (int64_t)(gp * 36 + 12 * d) * er.ne_ext;
#pragma unroll
                for (int q = 0; q < 12; ++q) cp_async8(dst + q * 256, src + (int64_t)q * er.ne_ext);
            }
            cp_async_commit();
        };
        if (PER_GP) stage_gp(0);
        {
            double2 *z2 = reinterpret_cast<double2 *>(tileA);
            for (int q = threadIdx.x; q < TILE_DOUBLES / 2; q += blockDim.x) z2[q] = make_double2(0., 0.);
        }
        unsigned anymask = 0;
        for (int q = threadIdx.x; q < 27 * 32; q += blockDim.x) {
            const int slot = q >> 5, l2 = q & 31;
            const int ddx = slot % 3 - 1, ddy = (slot / 3) % 3 - 1, ddz = slot / 9 - 1;
            // (the ghost plane of the symmetric layout would look two planes below the slab)
            const int64_t idx = g.G + ln0 + l2 + ddx + (int64_t)g.NX * ddy + g.npl * ddz;
            const uint8_t mk = (l2 < nvalid && idx >= 0 && idx < g.S) ? nodemask[idx] : 0;
            nbmask[q] = mk;
            anymask |= mk;
        }
        double blk[24];
#pragma unroll
        for (int q = 0; q < 24; ++q) blk[q] = 0.;
        if (PER_GP) {
#pragma unroll 1
            for (int gp = 0; gp < 8; ++gp) {
                if (gp < 7) { stage_gp(gp + 1); asm volatile("cp.async.wait_group 1;" ::: "memory"); }
                else asm volatile("cp.async.wait_group 0;" ::: "memory");
                __syncthreads();                                         // the three thirds of every element are in
                const double *ck = stage + (gp & 1) * (36 * 256);
                if (exists) {
                    if (d == 0) integrate_gp<true, 0, -1>(gp, a, ck, 256, blk);          // warp-uniform
                    else if (d == 1) integrate_gp<true, 1, -1>(gp, a, ck, 256, blk);
                    else integrate_gp<true, 2, -1>(gp, a, ck, 256, blk);
                }
                __syncthreads();                                         // buffer (gp & 1) may be refilled two trips later
            }
        } else if (exists) {
            if (d == 0) integrate_uniform<0>(a, blk);                                    // warp-uniform
            else if (d == 1) integrate_uniform<1>(a, blk);
            else integrate_uniform<2>(a, blk);
        }
        // does any node of the tile, or any of its neighbours, carry a Dirichlet dof?  (block-uniform)
        const int masked_tile = __syncthreads_or(anymask != 0);          // also: the tile is zeroed, the masks are in
        // 8 rounds: in round b every warp adds its row of block column b; for a fixed b the 24 warps
        // (different (a, d)) target 24 different (slot, row) pairs, so no two threads touch the same entry
#pragma unroll
        for (int b = 0; b < 8; ++b) {
            if (exists) {
#pragma unroll
                for (int cc = 0; cc < 3; ++cc) {
                    double *cell = tileA + (cell0[b] + cc) * TILE_NODES + lane;
                    *cell = fma(blk[3 * b + cc], wg, *cell);
                }
            }
            __syncthreads();
        }
        // PCJACOBI: the inverse diagonal (after MatZeroRowsColumns a Dirichlet row's diagonal is 1)
        if (warp < 3 && valid && (!SYM || kl >= 0)) {
            double v = tileA[(13 * 9 + 4 * warp) * TILE_NODES + lane];
            if ((nbmask[13 * 32 + lane] >> warp) & 1u) v = 1.;
            dinv[warp * g.S + g.G + ln0 + lane] = v != 0. ? 1. / v : 1.;
        }
        // MatZeroRowsColumns (bcs.c:341-347) + coalesced store; warp w takes the entry pairs w, w+24, ...
        const unsigned own = nbmask[13 * 32 + lane];
        auto apply_mask = [&](unsigned info, double v) -> double {
            if (info & 0x8000u) return 0.;
            const int slot = info & 31, rr = (info >> 5) & 3, cc = (info >> 7) & 3;
            const unsigned nb = nbmask[slot * 32 + lane];
            if (((own >> rr) & 1u) || ((nb >> cc) & 1u)) v = ((info >> 9) & 1u) ? 1. : 0.;
            return v;
        };
        const bool ghost_plane = SYM && kl < 0;                           // only the blocks towards the slab (slots 18..26) survive
        if (SYM) {
            double2 *At = A + tile * (SYM_PAIRS * TILE_NODES) + lane;
            for (int pr = warp; pr < SYM_PAIRS; pr += ASM_WARPS) {
                const int kk0 = 117 + 2 * pr;
                double v0 = tileA[kk0 * TILE_NODES + lane], v1 = tileA[(kk0 + 1) * TILE_NODES + lane];
                if (masked_tile) { v0 = apply_mask(c_kkinfo.v[kk0], v0); v1 = apply_mask(c_kkinfo.v[kk0 + 1], v1); }
                if (ghost_plane) { if (kk0 < 18 * 9) v0 = 0.; if (kk0 + 1 < 18 * 9) v1 = 0.; }
                if (stream_stores) __stcs(At + pr * TILE_NODES, make_double2(v0, v1));
                else At[pr * TILE_NODES] = make_double2(v0, v1);
            }
        } else {
            double2 *At = A + tile * (PAIRS * TILE_NODES) + lane;
            const unsigned *info2 = reinterpret_cast<const unsigned *>(c_kkinfo.v);      // two 16-bit entries per pair
            for (int pr = warp; pr < PAIRS; pr += ASM_WARPS) {
                double2 o2 = make_double2(tileA[(2 * pr) * TILE_NODES + lane], tileA[(2 * pr + 1) * TILE_NODES + lane]);
                if (masked_tile) {
                    const unsigned info = info2[pr];
                    o2 = make_double2(apply_mask(info & 0xffffu, o2.x), apply_mask(info >> 16, o2.y));
                }
                if (stream_stores) __stcs(At + pr * TILE_NODES, o2);
                else At[pr * TILE_NODES] = o2;
            }
        }
        __syncthreads();
    }
}

}  // namespace macroc
