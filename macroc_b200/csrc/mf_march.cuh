// mf_march.cuh -- matrix-free operator y = (M K M + I - M) x from the 27 class stencils, z-marching form.
//
// A CTA of 4 warps owns a column block of 32 x 16 nodes and marches through a segment of z planes; a thread
// owns FOUR nodes adjacent in y (warp = 8 x-positions x 4 row groups).  Step s uses ONE plane of M x -- the
// (32+2) x (16+2) x 3 patch, staged two steps ahead with cp.async into a ring of four slots (the rare
// Dirichlet entries are zeroed in shared memory by the thread that copied them, before the step's only
// barrier; a fourth slot keeps plane s-1 for the output) -- and every thread adds the plane's contributions to THREE output planes it keeps in registers:
//     y[s-1] += T[.., dz=+1] x[s],   y[s] += T[.., dz=0] x[s],   y[s+1] += T[.., dz=-1] x[s]
// (output-stationary, 36 accumulators).  Per step and thread: 54 shared-memory loads and 135 128-bit
// constant-bank loads feed 972 FMA -- a 3 x 3 stencil block is loaded once and used for the four nodes -- where
// the patch form of kernels.cuh (k_apply_mf3d: input-stationary, two CTA barriers around a 52 KB stage) needs
// 162 + 486.  Every x value leaves HBM once per column block.  Interior warps take the stencil entries from the
// constant bank; warps with a node on a face read the class tables from global memory (L1 resident).
#pragma once

#include "kernels.cuh"
#include "spmv_sym.cuh"     // cp_async8

namespace macroc {

constexpr int MZ_BX = 32, MZ_BY = 16, MZ_NY = 4;            // nodes per CTA in x, y; nodes per thread in y
constexpr int MZ_THREADS = MZ_BX * MZ_BY / MZ_NY;          // 128
constexpr int MZ_PITCH = 34;                               // = MZ_BX + 2; the row groups of a half-warp are 4 rows apart: 4 * 34 = 8 (mod 16) doubles, no bank conflicts
constexpr int MZ_ROWS = MZ_BY + 2;
constexpr int MZ_PLANE = MZ_PITCH * MZ_ROWS;               // doubles per component
constexpr int MZ_POINTS = (MZ_BX + 2) * MZ_ROWS;           // 612 staged points per plane
constexpr int MZ_NIT = (MZ_POINTS + MZ_THREADS - 1) / MZ_THREADS;   // 5
constexpr int MZ_SLOTS = 4;                               // planes s-1 (read by the output), s, s+1 and s+2 (in flight)
constexpr int MZ_MASKB = (MZ_PLANE + 15) / 16 * 16;        // Dirichlet bytes of a staged plane
constexpr int MZ_SLOT_BYTES = 3 * MZ_PLANE * 8 + MZ_MASKB;
constexpr int MZ_SMEM = MZ_SLOTS * MZ_SLOT_BYTES;          // 61 312 B: three CTAs per SM

template <bool DOT>
__global__ void __launch_bounds__(MZ_THREADS, 3)
k_apply_mf_march(GridDev g, const double *__restrict__ Tg, const uint8_t *__restrict__ nodemask, const double *__restrict__ x,
                 double *__restrict__ y, int k0, int k1, int bx, int by, int nseg, double *__restrict__ partial,
                 const int *__restrict__ done)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    auto sxp = [&](int slot, int comp) -> double * { return reinterpret_cast<double *>(smem_raw + slot * MZ_SLOT_BYTES) + comp * MZ_PLANE; };
    auto smk = [&](int slot) -> uint8_t * { return smem_raw + slot * MZ_SLOT_BYTES + 3 * MZ_PLANE * 8; };
    __shared__ double sm[MZ_THREADS / 32];
    if (done && *done) return;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tx = warp * 8 + (lane & 7), ty0 = (lane >> 3) * MZ_NY;
    const int cb = (ty0 + 1) * MZ_PITCH + tx + 1;          // the thread's first node inside a staged plane
    const int seglen = (k1 - k0 + nseg - 1) / nseg;
    const int nitems = bx * by * nseg;
    double dot = 0.;
    for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
        const int seg = item / (bx * by), rem = item - seg * (bx * by);
        const int i0 = (rem % bx) * MZ_BX, j0 = (rem / bx) * MZ_BY;
        const int zlo = k0 + seg * seglen, zhi = min(k1, zlo + seglen);          // slab-local output planes [zlo, zhi)
        if (zlo >= zhi) continue;                                                  // block-uniform
        const int i = i0 + tx, jb = j0 + ty0;
        const int cx = node_class(i, g.NX);
        bool xy_interior = i < g.NX ? cx == 1 : true;
        int cxy[MZ_NY];
#pragma unroll
        for (int jj = 0; jj < MZ_NY; ++jj) {
            const bool v = i < g.NX && jb + jj < g.NY;
            cxy[jj] = v ? cx + 3 * node_class(jb + jj, g.NY) : 4;
            xy_interior = xy_interior && cxy[jj] == 4;
        }
        // staging: plane s -> slot.  The Dirichlet bytes of this thread's points are loaded into registers and first
        // looked at one step later (an immediate use would expose the load latency once per step).
        unsigned mkb[MZ_NIT];
        auto stage_plane = [&](int s, int slot) {
            double *d0 = sxp(slot, 0), *d1 = sxp(slot, 1), *d2 = sxp(slot, 2);
#pragma unroll
            for (int it = 0; it < MZ_NIT; ++it) {
                const int q = threadIdx.x + it * MZ_THREADS;
                mkb[it] = 0;
                if (q < MZ_POINTS) {
                    const int py = q / (MZ_BX + 2), px = q - py * (MZ_BX + 2);
                    const int ii = i0 + px - 1;
                    const int64_t ln = (int64_t)s * g.npl + (int64_t)(j0 + py - 1) * g.NX + ii;
                    const bool ok = ii <= g.NX && ln >= -(int64_t)g.G && g.G + ln < g.S;
                    const int o = py * MZ_PITCH + px;
                    if (ok) {
                        const int64_t idx = g.G + ln;
                        mkb[it] = nodemask[idx];
                        cp_async8(d0 + o, x + idx); cp_async8(d1 + o, x + g.S + idx); cp_async8(d2 + o, x + 2 * g.S + idx);
                    } else {
                        d0[o] = 0.; d1[o] = 0.; d2[o] = 0.;
                    }
                }
            }
            cp_async_commit();
        };
        auto mask_plane = [&](int slot) {                  // M x: zero the Dirichlet entries this thread copied; park the bytes
            double *d0 = sxp(slot, 0), *d1 = sxp(slot, 1), *d2 = sxp(slot, 2);
            uint8_t *mb = smk(slot);
#pragma unroll
            for (int it = 0; it < MZ_NIT; ++it) {
                const int q = threadIdx.x + it * MZ_THREADS;
                if (q < MZ_POINTS) {
                    const int py = q / (MZ_BX + 2), px = q - py * (MZ_BX + 2);
                    const int o = py * MZ_PITCH + px;
                    const unsigned mk = mkb[it];
                    mb[o] = (uint8_t)mk;
                    if (mk & 1u) d0[o] = 0.;
                    if (mk & 2u) d1[o] = 0.;
                    if (mk & 4u) d2[o] = 0.;
                }
            }
        };
        __syncthreads();                                   // the previous item's planes have been read
        stage_plane(zlo - 1, (zlo - 1 + 4) % MZ_SLOTS);
        cp_async_wait_all();
        mask_plane((zlo - 1 + 4) % MZ_SLOTS);
        stage_plane(zlo, (zlo + 4) % MZ_SLOTS);            // masked at the top of the first step
        double am[MZ_NY][3], ac[MZ_NY][3], ap[MZ_NY][3];   // output planes s-1, s, s+1
#pragma unroll
        for (int jj = 0; jj < MZ_NY; ++jj)
#pragma unroll
            for (int r = 0; r < 3; ++r) am[jj][r] = ac[jj][r] = ap[jj][r] = 0.;
        for (int s = zlo - 1; s <= zhi; ++s) {
            const int cur = (s + 4) % MZ_SLOTS, prv = (s + 3) % MZ_SLOTS;
            cp_async_wait_all();                           // this thread's copies of plane s+1 (issued one step ago) are in
            if (s + 1 <= zhi) mask_plane((s + 5) % MZ_SLOTS);
            __syncthreads();                               // planes s and s+1 are complete; the slot of plane s-2 has been read
            if (s + 2 <= zhi) stage_plane(s + 2, (s + 6) % MZ_SLOTS);
            // classes of the three output planes (global z = slab-local + zs); planes outside the grid are never written
            const int zg = s + g.zs;
            const int czm = 9 * node_class(zg - 1, g.NZ), czc = 9 * node_class(zg, g.NZ), czp = 9 * node_class(zg + 1, g.NZ);
            const bool interior = xy_interior && czm == 9 && czc == 9 && czp == 9;
            const double *s0 = sxp(cur, 0) + cb, *s1 = sxp(cur, 1) + cb, *s2 = sxp(cur, 2) + cb;
            if (__all_sync(0xffffffffu, interior)) {
#pragma unroll
                for (int dx = -1; dx <= 1; ++dx) {
                    double xr[MZ_NY + 2][3];               // the six rows this thread's nodes see at this dx
#pragma unroll
                    for (int py = 0; py < MZ_NY + 2; ++py) {
                        const int o = (py - 1) * MZ_PITCH + dx;
                        xr[py][0] = s0[o]; xr[py][1] = s1[o]; xr[py][2] = s2[o];
                    }
#pragma unroll
                    for (int dy = -1; dy <= 1; ++dy) {
                        const int sl = ((dy + 1) * 3 + (dx + 1)) * 9;
                        constexpr int base = 13 * 243;
#pragma unroll
                        for (int r = 0; r < 3; ++r) {
                            const double p0 = c_T[base + sl + 3 * r], p1 = c_T[base + sl + 3 * r + 1], p2 = c_T[base + sl + 3 * r + 2];
                            const double c0 = c_T[base + 81 + sl + 3 * r], c1 = c_T[base + 81 + sl + 3 * r + 1], c2 = c_T[base + 81 + sl + 3 * r + 2];
                            const double m0 = c_T[base + 162 + sl + 3 * r], m1 = c_T[base + 162 + sl + 3 * r + 1], m2 = c_T[base + 162 + sl + 3 * r + 2];
#pragma unroll
                            for (int jj = 0; jj < MZ_NY; ++jj) {
                                const double x0 = xr[jj + 1 + dy][0], x1 = xr[jj + 1 + dy][1], x2 = xr[jj + 1 + dy][2];
                                ap[jj][r] = fma(p2, x2, fma(p1, x1, fma(p0, x0, ap[jj][r])));
                                ac[jj][r] = fma(c2, x2, fma(c1, x1, fma(c0, x0, ac[jj][r])));
                                am[jj][r] = fma(m2, x2, fma(m1, x1, fma(m0, x0, am[jj][r])));
                            }
                        }
                    }
                }
            } else {
                // boundary classes: per-node class offsets into the global copy of the table (L1 resident; lanes of
                // one class broadcast, divergent constant-bank reads would serialise)
#pragma unroll 1
                for (int dx = -1; dx <= 1; ++dx) {
                    double xr[MZ_NY + 2][3];
#pragma unroll
                    for (int py = 0; py < MZ_NY + 2; ++py) {
                        const int o = (py - 1) * MZ_PITCH + dx;
                        xr[py][0] = s0[o]; xr[py][1] = s1[o]; xr[py][2] = s2[o];
                    }
#pragma unroll
                    for (int dy = -1; dy <= 1; ++dy) {
                        const int sl = ((dy + 1) * 3 + (dx + 1)) * 9;
#pragma unroll
                        for (int jj = 0; jj < MZ_NY; ++jj) {
                            const double *Tp = Tg + (cxy[jj] + czp) * 243 + sl, *Tc = Tg + (cxy[jj] + czc) * 243 + 81 + sl,
                                         *Tm = Tg + (cxy[jj] + czm) * 243 + 162 + sl;
                            const double x0 = xr[jj + 1 + dy][0], x1 = xr[jj + 1 + dy][1], x2 = xr[jj + 1 + dy][2];
#pragma unroll
                            for (int r = 0; r < 3; ++r) {
                                ap[jj][r] = fma(__ldg(Tp + 3 * r + 2), x2, fma(__ldg(Tp + 3 * r + 1), x1, fma(__ldg(Tp + 3 * r), x0, ap[jj][r])));
                                ac[jj][r] = fma(__ldg(Tc + 3 * r + 2), x2, fma(__ldg(Tc + 3 * r + 1), x1, fma(__ldg(Tc + 3 * r), x0, ac[jj][r])));
                                am[jj][r] = fma(__ldg(Tm + 3 * r + 2), x2, fma(__ldg(Tm + 3 * r + 1), x1, fma(__ldg(Tm + 3 * r), x0, am[jj][r])));
                            }
                        }
                    }
                }
            }
            if (s - 1 >= zlo && i < g.NX) {                // plane s-1 has all three contributions (s-1 < zhi always)
                const double *q0 = sxp(prv, 0) + cb, *q1 = sxp(prv, 1) + cb, *q2 = sxp(prv, 2) + cb;   // M x of plane s-1 (still resident)
                const uint8_t *qm = smk(prv) + cb;
#pragma unroll
                for (int jj = 0; jj < MZ_NY; ++jj) {
                    if (jb + jj >= g.NY) continue;
                    const int64_t ln = (int64_t)(s - 1) * g.npl + (int64_t)(jb + jj) * g.NX + i;
                    const unsigned own = qm[jj * MZ_PITCH];
                    double a0 = am[jj][0], a1 = am[jj][1], a2 = am[jj][2];
                    double p0 = q0[jj * MZ_PITCH], p1 = q1[jj * MZ_PITCH], p2 = q2[jj * MZ_PITCH];
                    if (own) {                             // Dirichlet rows are identity rows: y = x (unmasked)
                        const double *xo = x + g.G + ln;
                        if (own & 1u) { p0 = __ldg(xo); a0 = p0; }
                        if (own & 2u) { p1 = __ldg(xo + g.S); a1 = p1; }
                        if (own & 4u) { p2 = __ldg(xo + 2 * g.S); a2 = p2; }
                    }
                    if (DOT && owned_node(g, ln)) dot += a0 * p0 + a1 * p1 + a2 * p2;
                    double *y0 = y + g.G + ln;
                    y0[0] = a0; y0[g.S] = a1; y0[2 * g.S] = a2;
                }
            }
#pragma unroll
            for (int jj = 0; jj < MZ_NY; ++jj)
#pragma unroll
                for (int r = 0; r < 3; ++r) { am[jj][r] = ac[jj][r]; ac[jj][r] = ap[jj][r]; ap[jj][r] = 0.; }
        }
    }
    if (DOT) {
        double s = block_sum<MZ_THREADS / 32>(dot, sm);
        if (threadIdx.x == 0) partial[blockIdx.x] = s;
    }
}

}  // namespace macroc
