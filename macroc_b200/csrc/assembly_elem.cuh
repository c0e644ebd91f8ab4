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
//  * Jacobian: assembly_node.cuh (node-centric: one thread per operator entry).  integrate_gp below is
//    the element-centric row integration it replaced (24 warps per tile + eight shared-memory reduction
//    rounds: 34.1 ms against 25.0 at 256^3, profiles/r2_jac_history.md); it stays as the DFMA side of the
//    DMMA A/B (dmma_ab.cuh).
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

}  // namespace macroc
