// strict_fp.cuh -- verification mode (cfg.strict_fp): the hot path with the REFERENCE'S ROUNDING.
//
// The production kernels use fused multiply-adds and tree reductions; the reference (compiled for
// plain x86-64 SSE2, one rank) rounds every product and sums in loop order.  Both are correct to
// rounding, but CG amplifies rounding differences, so at the reference's own rtol = 1e-5 the two
// displacement fields agree only to solver tolerance.  This file restates the same algorithm with
// the reference's operation order and no contraction:
//   * residual:   strain, stress and B^T sigma wg exactly as assembly.c:45-56,142-154 round them;
//   * SpMV:       row sums in CSR column order (slot, then column), s = s + (a * x)  (MatMult_SeqAIJ);
//   * CG updates: x += a p, r += (-a) w, z = r / d, p = z + b p  with separately rounded products;
//   * dots/norms: ONE thread adds in dof order (VecDot / VecNorm of a sequential Vec).
// With it the time loop reproduces the reference binary BIT FOR BIT (tests/test_gpu_parity.py), which
// turns "the differences are rounding order, not arithmetic" from an explanation into a measurement.
// One rank, uniform tangent, full-storage operator; orders of magnitude slower than the product path.
#pragma once

#include "kernels.cuh"
#include "assembly_elem.cuh"

namespace macroc {

__device__ __forceinline__ double madd_rn(double a, double b, double c) { return __dadd_rn(c, __dmul_rn(a, b)); }   // c + (a*b)

// element forces, the reference's rounding (cf. k_elem_forces<false>)
__global__ void __launch_bounds__(128)
k_elem_forces_strict(GridDev g, ElemRange er, int l0, int nl, double wg, const double *__restrict__ u,
                     double *__restrict__ scratch)
{
    const int64_t per_layer = er.nex * er.ney, n = per_layer * nl;
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    int ei = (int)(e % er.nex), ej = (int)((e / er.nex) % er.ney), el = (int)(e / per_layer) + l0;
    double be[24];
#pragma unroll
    for (int q = 0; q < 24; ++q) be[q] = 0.;
    double ue[8][3];
    gather_element(u, g, g.G + ei + (int64_t)g.NX * ej + g.npl * (er.ezs + el - g.zs), ue);
#pragma unroll 1
    for (int gp = 0; gp < 8; ++gp) {
        // strain[i] = sum_j B[i][j] u_e[j], j ascending (assembly.c:52-54); zero entries of B add +-0
        double e0 = 0., e1 = 0., e2 = 0., e3 = 0., e4 = 0., e5 = 0.;
#pragma unroll
        for (int nn = 0; nn < 8; ++nn) {
            const double hx = c_dsh[gp][nn][0], hy = c_dsh[gp][nn][1], hz = c_dsh[gp][nn][2];
            e0 = madd_rn(hx, ue[nn][0], e0);
            e1 = madd_rn(hy, ue[nn][1], e1);
            e2 = madd_rn(hz, ue[nn][2], e2);
            e3 = madd_rn(hy, ue[nn][0], e3); e3 = madd_rn(hx, ue[nn][1], e3);
            e4 = madd_rn(hz, ue[nn][0], e4); e4 = madd_rn(hx, ue[nn][2], e4);
            e5 = madd_rn(hz, ue[nn][1], e5); e5 = madd_rn(hy, ue[nn][2], e5);
        }
        const double eps[6] = {e0, e1, e2, e3, e4, e5};
        double sig[6];
#pragma unroll
        for (int i = 0; i < 6; ++i) {               // sigma = D eps (the MicroPP stand-in's loop)
            double t = 0.;
#pragma unroll
            for (int j = 0; j < 6; ++j) t = madd_rn(c_D[i * 6 + j], eps[j], t);
            sig[i] = t;
        }
#pragma unroll
        for (int a = 0; a < 8; ++a) {
            const double hx = c_dsh[gp][a][0], hy = c_dsh[gp][a][1], hz = c_dsh[gp][a][2];
            // be[i] += B[j][i]*stress[j]*wg, j ascending (assembly.c:151-153): ((B * s) * wg), then the add
            auto acc = [&](double &b, double h, double s) { b = __dadd_rn(b, __dmul_rn(__dmul_rn(h, s), wg)); };
            acc(be[3 * a + 0], hx, sig[0]); acc(be[3 * a + 0], hy, sig[3]); acc(be[3 * a + 0], hz, sig[4]);
            acc(be[3 * a + 1], hy, sig[1]); acc(be[3 * a + 1], hx, sig[3]); acc(be[3 * a + 1], hz, sig[5]);
            acc(be[3 * a + 2], hz, sig[2]); acc(be[3 * a + 2], hx, sig[4]); acc(be[3 * a + 2], hy, sig[5]);
        }
    }
#pragma unroll
    for (int q = 0; q < 24; ++q) scratch[q * n + e] = be[q];
}

// sum_i a_i * b_i over the dofs in natural order (node-major, component fastest), one thread.
// mode 0: out[0] = a.b      mode 1: with z = a (*) dinv rounded:  out[0] = z.z, out[1] = z.a   (a = r)
__global__ void k_seq_dots(GridDev g, int mode, const double *__restrict__ a, const double *__restrict__ b,
                           double *__restrict__ out)
{
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    double s0 = 0., s1 = 0.;
    if (mode == 0) {
        for (int64_t ln = 0; ln < g.nloc; ++ln)
            for (int d = 0; d < 3; ++d) {
                const int64_t q = d * g.S + g.G + ln;
                s0 = madd_rn(a[q], b[q], s0);
            }
    } else {
        // the reference forms z first (VecPointwiseMult), then VecNorm(z), then VecDot(z, r): two passes
        for (int64_t ln = 0; ln < g.nloc; ++ln)
            for (int d = 0; d < 3; ++d) {
                const int64_t q = d * g.S + g.G + ln;
                const double z = __dmul_rn(a[q], b[q]);
                s0 = madd_rn(z, z, s0);
                s1 = madd_rn(z, a[q], s1);
            }
    }
    out[0] = s0; out[1] = s1;
}

// w = A p, row sums in CSR column order without contraction (full 27-slot tile layout)
__global__ void __launch_bounds__(256)
k_spmv_strict(GridDev g, const double2 *__restrict__ A, const double *__restrict__ p, double *__restrict__ w)
{
    const int64_t ln = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (ln >= g.nloc) return;
    const int64_t tile = ln / TILE_NODES;
    const int lane = (int)(ln % TILE_NODES);
    const double *At = reinterpret_cast<const double *>(A) + tile * TILE_DOUBLES;
    double acc[3] = {0., 0., 0.};
    for (int slot = 0; slot < 27; ++slot) {
        const int ddx = slot % 3 - 1, ddy = (slot / 3) % 3 - 1, ddz = slot / 9 - 1;
        const int64_t q = g.G + ln + ddx + (int64_t)g.NX * ddy + g.npl * ddz;
        const double x0 = p[q], x1 = p[g.S + q], x2 = p[2 * g.S + q];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            const int kk = slot * 9 + 3 * r;
            const double m0 = At[(((kk + 0) >> 1) * TILE_NODES + lane) * 2 + ((kk + 0) & 1)];
            const double m1 = At[(((kk + 1) >> 1) * TILE_NODES + lane) * 2 + ((kk + 1) & 1)];
            const double m2 = At[(((kk + 2) >> 1) * TILE_NODES + lane) * 2 + ((kk + 2) & 1)];
            acc[r] = madd_rn(m0, x0, acc[r]); acc[r] = madd_rn(m1, x1, acc[r]); acc[r] = madd_rn(m2, x2, acc[r]);
        }
    }
    w[g.G + ln] = acc[0]; w[g.S + g.G + ln] = acc[1]; w[2 * g.S + g.G + ln] = acc[2];
}

// x = 0, r = b
__global__ void k_cg_init_strict(GridDev g, const double *__restrict__ b, double *__restrict__ x, double *__restrict__ r)
{
    const int64_t ln = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (ln >= g.nloc) return;
    for (int d = 0; d < 3; ++d) { const int64_t q = d * g.S + g.G + ln; x[q] = 0.; r[q] = b[q]; }
}

// p = z (first iteration) or z + (beta/betaold) p, z = r (*) dinv      (VecCopy / VecAYPX)
__global__ void k_cg_update_p_strict(GridDev g, const CgScalars *__restrict__ s, const double *__restrict__ r,
                                     const double *__restrict__ dinv, double *__restrict__ p)
{
    if (s->done) return;
    const int64_t ln = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (ln >= g.nloc) return;
    const bool first = s->its == 0;
    const double bb = s->beta / s->betaold;
    for (int d = 0; d < 3; ++d) {
        const int64_t q = d * g.S + g.G + ln;
        const double z = __dmul_rn(r[q], dinv[q]);
        p[q] = first ? z : __dadd_rn(z, __dmul_rn(bb, p[q]));
    }
}

// a = beta / p.w;  x += a p;  r += (-a) w       (two VecAXPY)
__global__ void k_cg_update_xr_strict(GridDev g, const CgScalars *__restrict__ s, const double *__restrict__ p,
                                      const double *__restrict__ w, double *__restrict__ x, double *__restrict__ r)
{
    if (s->done) return;
    const int64_t ln = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (ln >= g.nloc) return;
    const double a = s->beta / s->pw, ma = -a;
    for (int d = 0; d < 3; ++d) {
        const int64_t q = d * g.S + g.G + ln;
        x[q] = __dadd_rn(x[q], __dmul_rn(a, p[q]));
        r[q] = __dadd_rn(r[q], __dmul_rn(ma, w[q]));
    }
}

// |b|^2 in dof order (VecNorm of the sequential Vec, main.c:67)
__global__ void k_seq_norm2(GridDev g, const double *__restrict__ b, double *__restrict__ out)
{
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    double s = 0.;
    for (int64_t ln = 0; ln < g.nloc; ++ln)
        for (int d = 0; d < 3; ++d) { const double v = b[d * g.S + g.G + ln]; s = madd_rn(v, v, s); }
    out[0] = s;
}

}  // namespace macroc
