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
constexpr int ASM_COLBLOCK = 64;

template <bool PER_GP, bool SYM>
__global__ void __launch_bounds__(256, 1)
k_assemble_elements(GridDev g, SymGeom sg, ElemRange er, double wg, const double *__restrict__ ctan_gp,
                    const uint8_t *__restrict__ nodemask, double2 *__restrict__ A, double *__restrict__ dinv,
                    int64_t tile_lo, int64_t tile_hi, int64_t tpp /* tiles per plane (rounded up for the full layout) */)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *tileA = reinterpret_cast<double *>(smem_raw);                  // TILE_DOUBLES, full 27-slot indexing
    uint8_t *nbmask = smem_raw + TILE_DOUBLES * sizeof(double);            // [27][32]
    const int lane = threadIdx.x & 31, a = threadIdx.x >> 5;
    const int apx = node_px(a), apy = node_py(a), apz = node_pz(a);
    const int64_t per_layer = er.nex * er.ney;

    // Traversal: column blocks of ASM_COLBLOCK tiles of a plane, swept through all planes before the next
    // block.  An element's tangents are needed by the tiles of two rows and two planes; with the plain
    // linear order the second plane comes a whole plane (150 MB of tangents + 128 MB of operator at 256^3)
    // later and misses L2 (ncu: 81 GB read for 38 GB of tangents).  Here the CTAs in flight cover a few
    // planes of one column block, so the second use follows the first within a few MB of traffic.
    const int64_t ntl = tile_hi - tile_lo;
    const int64_t mtot = (ntl + tpp - 1) / tpp, ncb = (tpp + ASM_COLBLOCK - 1) / ASM_COLBLOCK;
    for (int64_t v = blockIdx.x; v < ncb * ASM_COLBLOCK * mtot; v += gridDim.x) {
        const int64_t cb = v / (ASM_COLBLOCK * mtot), rem = v - cb * (ASM_COLBLOCK * mtot);
        const int64_t col = cb * ASM_COLBLOCK + rem % ASM_COLBLOCK;
        const int64_t tile = tile_lo + col + (rem / ASM_COLBLOCK) * tpp;
        if (col >= tpp || tile >= tile_hi) continue;                       // block-uniform
        // node of (tile, lane): local box coordinates (i, j), slab-local plane kl, linear index ln0 of lane 0
        int i = 0, j = 0, kl = 0, nvalid = 0;
        int64_t ln0;
        if (SYM) {
            kl = (int)((tile + tpp) / tpp) - 1;                            // floor: the ghost plane is -1
            const int64_t rem = tile - (int64_t)kl * tpp;
            j = (int)(rem / sg.rt);
            const int x0 = (int)(rem % sg.rt) * 32;
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
        for (int q = threadIdx.x; q < TILE_DOUBLES; q += blockDim.x) tileA[q] = 0.;
        for (int q = threadIdx.x; q < 27 * 32; q += blockDim.x) {
            const int slot = q >> 5, l2 = q & 31;
            const int ddx = slot % 3 - 1, ddy = (slot / 3) % 3 - 1, ddz = slot / 9 - 1;
            // (the ghost plane of the symmetric layout would look two planes below the slab)
            const int64_t idx = g.G + ln0 + l2 + ddx + (int64_t)g.NX * ddy + g.npl * ddz;
            nbmask[q] = (l2 < nvalid && idx >= 0 && idx < g.S) ? nodemask[idx] : 0;
        }
        // the element in which this node is local node a (it must be one whose tangents this rank holds)
        const int ei = i - apx, ej = j - apy, ek = k - apz;
        const bool exists = valid && ei >= 0 && ei < g.NX - 1 && ej >= 0 && ej < g.NY - 1 && ek >= 0 && ek < g.NZ - 1 &&
                            ek >= er.ezs && ek < er.ezs + er.nez_ext;
        double blk[3][24];
#pragma unroll
        for (int d = 0; d < 3; ++d)
#pragma unroll
            for (int q = 0; q < 24; ++q) blk[d][q] = 0.;
        if (exists) {
            const double *cg = PER_GP ? ctan_gp + ((int64_t)(ek - er.ezs) * per_layer + ei + er.nex * (int64_t)ej) : nullptr;
            // uniform tangent: fully unrolled (every B entry an immediate operand).  Per-Gauss-point
            // tangents: one Gauss point per loop trip -- unrolled, the 288 loads of a thread serialise
            // behind the register allocator (measured 84 ms against 67 ms at 256^3)
#pragma unroll (PER_GP ? 1 : 8)
            for (int gp = 0; gp < 8; ++gp) {
                const double hx = c_dsh[gp][a][0], hy = c_dsh[gp][a][1], hz = c_dsh[gp][a][2];
                // stream the tangent one column k at a time: T[d] = (B_a^T C)[d][k], then every block
                // column b takes T[d] * B_b[k][.]  (B has at most two non-zeros per (k, b))
#pragma unroll
                for (int kc = 0; kc < 6; ++kc) {
                    double ck[6];
#pragma unroll
                    for (int r = 0; r < 6; ++r) ck[r] = PER_GP ? __ldg(cg + (int64_t)(gp * 36 + r * 6 + kc) * er.ne_ext) : c_D[r * 6 + kc];
                    const double T0 = fma(hz, ck[4], fma(hy, ck[3], hx * ck[0]));
                    const double T1 = fma(hz, ck[5], fma(hx, ck[3], hy * ck[1]));
                    const double T2 = fma(hy, ck[5], fma(hx, ck[4], hz * ck[2]));
#pragma unroll
                    for (int b = 0; b < 8; ++b) {
                        const double bx = c_dsh[gp][b][0], by = c_dsh[gp][b][1], bz = c_dsh[gp][b][2];
                        // column 3b+c of B: row k non-zero for (k,c) in {(0,0),(1,1),(2,2),(3,0)=by,(3,1)=bx,(4,0)=bz,(4,2)=bx,(5,1)=bz,(5,2)=by}
                        if (kc == 0) { blk[0][3 * b + 0] = fma(T0, bx, blk[0][3 * b + 0]); blk[1][3 * b + 0] = fma(T1, bx, blk[1][3 * b + 0]); blk[2][3 * b + 0] = fma(T2, bx, blk[2][3 * b + 0]); }
                        if (kc == 1) { blk[0][3 * b + 1] = fma(T0, by, blk[0][3 * b + 1]); blk[1][3 * b + 1] = fma(T1, by, blk[1][3 * b + 1]); blk[2][3 * b + 1] = fma(T2, by, blk[2][3 * b + 1]); }
                        if (kc == 2) { blk[0][3 * b + 2] = fma(T0, bz, blk[0][3 * b + 2]); blk[1][3 * b + 2] = fma(T1, bz, blk[1][3 * b + 2]); blk[2][3 * b + 2] = fma(T2, bz, blk[2][3 * b + 2]); }
                        if (kc == 3) {
                            blk[0][3 * b + 0] = fma(T0, by, blk[0][3 * b + 0]); blk[1][3 * b + 0] = fma(T1, by, blk[1][3 * b + 0]); blk[2][3 * b + 0] = fma(T2, by, blk[2][3 * b + 0]);
                            blk[0][3 * b + 1] = fma(T0, bx, blk[0][3 * b + 1]); blk[1][3 * b + 1] = fma(T1, bx, blk[1][3 * b + 1]); blk[2][3 * b + 1] = fma(T2, bx, blk[2][3 * b + 1]);
                        }
                        if (kc == 4) {
                            blk[0][3 * b + 0] = fma(T0, bz, blk[0][3 * b + 0]); blk[1][3 * b + 0] = fma(T1, bz, blk[1][3 * b + 0]); blk[2][3 * b + 0] = fma(T2, bz, blk[2][3 * b + 0]);
                            blk[0][3 * b + 2] = fma(T0, bx, blk[0][3 * b + 2]); blk[1][3 * b + 2] = fma(T1, bx, blk[1][3 * b + 2]); blk[2][3 * b + 2] = fma(T2, bx, blk[2][3 * b + 2]);
                        }
                        if (kc == 5) {
                            blk[0][3 * b + 1] = fma(T0, bz, blk[0][3 * b + 1]); blk[1][3 * b + 1] = fma(T1, bz, blk[1][3 * b + 1]); blk[2][3 * b + 1] = fma(T2, bz, blk[2][3 * b + 1]);
                            blk[0][3 * b + 2] = fma(T0, by, blk[0][3 * b + 2]); blk[1][3 * b + 2] = fma(T1, by, blk[1][3 * b + 2]); blk[2][3 * b + 2] = fma(T2, by, blk[2][3 * b + 2]);
                        }
                    }
                }
            }
        }
        __syncthreads();
        // 8 rounds: in round b every warp adds block column b; for a fixed b the 8 warps
        // (different a) target 8 different slots, so no two threads touch the same entry
#pragma unroll
        for (int b = 0; b < 8; ++b) {
            if (exists) {
                const int slot = (node_pz(b) - apz + 1) * 9 + (node_py(b) - apy + 1) * 3 + (node_px(b) - apx + 1);
#pragma unroll
                for (int d = 0; d < 3; ++d)
#pragma unroll
                    for (int cc = 0; cc < 3; ++cc) {
                        const int kk = slot * 9 + 3 * d + cc;
                        double *cell = tileA + ((kk >> 1) * TILE_NODES + lane) * 2 + (kk & 1);
                        *cell = fma(blk[d][3 * b + cc], wg, *cell);
                    }
            }
            __syncthreads();
        }
        // MatZeroRowsColumns (bcs.c:341-347) + PCJACOBI diagonal + coalesced store; warp a takes the
        // entry pairs a, a+8, ...: the entry decoding (slot, row, col) is a warp-uniform table lookup
        const unsigned own = nbmask[13 * 32 + lane];
        auto apply_mask = [&](unsigned info, double v) -> double {
            if (info & 0x8000u) return 0.;
            const int slot = info & 31, rr = (info >> 5) & 3, cc = (info >> 7) & 3;
            const unsigned nb = nbmask[slot * 32 + lane];
            const bool isdiag = (info >> 9) & 1u;
            if (((own >> rr) & 1u) || ((nb >> cc) & 1u)) v = isdiag ? 1. : 0.;
            if (SYM && kl < 0 && slot < 18) v = 0.;                  // ghost plane: only the blocks towards the slab
            if (isdiag && valid && (!SYM || kl >= 0)) dinv[rr * g.S + g.G + ln0 + lane] = v != 0. ? 1. / v : 1.;
            return v;
        };
        // streaming stores: the operator must not push the tangents of the next plane out of L2
        if (SYM) {
            double2 *At = A + tile * (SYM_PAIRS * TILE_NODES) + lane;
            for (int pr = a; pr < SYM_PAIRS; pr += 8) {
                const int kk0 = 117 + 2 * pr;                          // odd: the pair straddles two pairs of the staging tile
                const double v0 = apply_mask(c_kkinfo.v[kk0], tileA[((kk0 >> 1) * TILE_NODES + lane) * 2 + 1]);
                const double v1 = apply_mask(c_kkinfo.v[kk0 + 1], tileA[(((kk0 + 1) >> 1) * TILE_NODES + lane) * 2]);
                __stcs(At + pr * TILE_NODES, make_double2(v0, v1));
            }
        } else {
            double2 *At = A + tile * (PAIRS * TILE_NODES) + lane;
            const unsigned *info2 = reinterpret_cast<const unsigned *>(c_kkinfo.v);      // two 16-bit entries per pair
            for (int pr = a; pr < PAIRS; pr += 8) {
                const unsigned info = info2[pr];
                const double2 t2 = *reinterpret_cast<const double2 *>(tileA + (pr * TILE_NODES + lane) * 2);
                __stcs(At + pr * TILE_NODES, make_double2(apply_mask(info & 0xffffu, t2.x), apply_mask(info >> 16, t2.y)));
            }
        }
        __syncthreads();
    }
}

}  // namespace macroc
