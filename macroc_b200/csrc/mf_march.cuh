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
// 162 + 486.  Every x value leaves HBM once per column block.  The marching kernel knows one stencil only, the
// interior class, taken from the constant bank; the nodes on a face of the box (2.3 % at 256^3: their rows sum fewer
// elements, 26 other classes) are skipped by it and done by k_apply_mf_faces, one thread per node, class table
// from global memory.  (A per-lane class inside the marching loop cost more than everything else: a fifth of
// the warps held a face node and ran a 972-load path 5.8 times slower than their neighbours, who waited for
// them at the barrier -- ncu, profiles/r2_mf_history.md.)
#pragma once

#include "kernels.cuh"
#include "spmv_sym.cuh"     // cp_async8

namespace macroc {

#ifndef MACROC_MZ_NY
#define MACROC_MZ_NY 4                                      // measurement builds: -DMACROC_MZ_NY=3 (12-row blocks, four CTAs per SM)
#endif
constexpr int MZ_NY = MACROC_MZ_NY;                        // nodes per thread in y
constexpr int MZ_BX = 32, MZ_BY = 4 * MZ_NY;               // nodes per CTA in x, y (a warp = 8 x-positions x 4 row groups)
constexpr int MZ_CTAS = MZ_NY == 4 ? 3 : 4;                // resident CTAs per SM (registers)
constexpr int MZ_THREADS = MZ_BX * MZ_BY / MZ_NY;          // 128
// row pitch >= MZ_BX + 2 with (MZ_NY * pitch) = 8 (mod 16): the row groups of a half-warp fall into disjoint bank halves
constexpr int MZ_PITCH = MZ_NY == 4 ? 34 : 40;
static_assert((MZ_NY * MZ_PITCH) % 16 == 8 && MZ_PITCH >= MZ_BX + 2, "bank-conflict-free row pitch");
constexpr int MZ_ROWS = MZ_BY + 2;
constexpr int MZ_PLANE = MZ_PITCH * MZ_ROWS;               // doubles per component
constexpr int MZ_POINTS = (MZ_BX + 2) * MZ_ROWS;           // 612 staged points per plane
constexpr int MZ_NIT = (MZ_POINTS + MZ_THREADS - 1) / MZ_THREADS;   // 5
constexpr int MZ_SLOTS = 4;                               // planes s-1 (read by the output), s, s+1 and s+2 (in flight)
constexpr int MZ_MASKB = (MZ_PLANE + 15) / 16 * 16;        // Dirichlet bytes of a staged plane
constexpr int MZ_SLOT_BYTES = 3 * MZ_PLANE * 8 + MZ_MASKB;
constexpr int MZ_SMEM = MZ_SLOTS * MZ_SLOT_BYTES;          // 61 312 B: three CTAs per SM

template <bool DOT>
__global__ void __launch_bounds__(MZ_THREADS, MZ_CTAS)
k_apply_mf_march(GridDev g, const double *__restrict__ Tg, const uint8_t *__restrict__ nodemask, const double *__restrict__ x,
                 double *__restrict__ y, int k0, int k1, int bx, int by, int nseg, double *__restrict__ partial /* all partials of this apply */,
                 int part0 /* this kernel's first slot in it (the face kernel's come before) */, const int *__restrict__ done, CgFuse fuse)
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
        // nodes on an x or y face of the box belong to k_apply_mf_faces
        bool face_xy[MZ_NY];
#pragma unroll
        for (int jj = 0; jj < MZ_NY; ++jj) face_xy[jj] = i == 0 || i == g.NX - 1 || jb + jj == 0 || jb + jj == g.NY - 1;
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
            const double *s0 = sxp(cur, 0) + cb, *s1 = sxp(cur, 1) + cb, *s2 = sxp(cur, 2) + cb;
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
            if (s - 1 >= zlo && i < g.NX) {                // plane s-1 has all three contributions (s-1 < zhi always)
                const double *q0 = sxp(prv, 0) + cb, *q1 = sxp(prv, 1) + cb, *q2 = sxp(prv, 2) + cb;   // M x of plane s-1 (still resident)
                const uint8_t *qm = smk(prv) + cb;
                const int zgm = s - 1 + g.zs;
                const bool zface = zgm == 0 || zgm == g.NZ - 1;
#pragma unroll
                for (int jj = 0; jj < MZ_NY; ++jj) {
                    if (jb + jj >= g.NY || face_xy[jj] || zface) continue;
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
        if (threadIdx.x == 0) partial[part0 + blockIdx.x] = s;
        if (fuse.ticket) cg_last_block<MZ_THREADS / 32, 1>(partial, part0 + gridDim.x, fuse, sm);
    }
}

// The nodes on a face of the box, planes [k0, k1): x faces (all their nodes), y faces (without the x-face nodes),
// global z faces (without both).  One thread per node, its class stencil from the global table.
__host__ __device__ __forceinline__ int64_t mf_face_count(const GridDev &g, int k0, int k1)
{
    const int64_t nzp = k1 - k0, nx2 = g.NX > 2 ? g.NX - 2 : 0, ny2 = g.NY > 2 ? g.NY - 2 : 0;
    const int zf = (k0 + g.zs <= 0 && 0 < k1 + g.zs ? 1 : 0) + (g.NZ > 1 && k0 + g.zs <= g.NZ - 1 && g.NZ - 1 < k1 + g.zs ? 1 : 0);
    return (g.NX > 1 ? 2 : 1) * (int64_t)g.NY * nzp + (g.NY > 1 ? 2 : 1) * nx2 * nzp + zf * nx2 * ny2;
}

template <bool DOT>
__global__ void __launch_bounds__(128)
k_apply_mf_faces(GridDev g, const double *__restrict__ Tg, const uint8_t *__restrict__ nodemask, const double *__restrict__ x,
                 double *__restrict__ y, int k0, int k1, double *__restrict__ partial, const int *__restrict__ done)
{
    __shared__ double sm[4];
    if (done && *done) return;
    const int64_t nzp = k1 - k0, nx2 = g.NX > 2 ? g.NX - 2 : 0, ny2 = g.NY > 2 ? g.NY - 2 : 0;
    const int sxn = g.NX > 1 ? 2 : 1, syn = g.NY > 1 ? 2 : 1;
    const int64_t nA = sxn * (int64_t)g.NY * nzp, nB = syn * nx2 * nzp, total = mf_face_count(g, k0, k1);
    const bool zlow = k0 + g.zs <= 0 && 0 < k1 + g.zs;
    double dot = 0.;
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < total; q += (int64_t)gridDim.x * blockDim.x) {
        int i, j, kl;
        if (q < nA) {
            const int64_t per = (int64_t)g.NY * nzp;
            const int side = (int)(q / per); const int64_t r = q - side * per;
            i = side ? g.NX - 1 : 0; j = (int)(r % g.NY); kl = k0 + (int)(r / g.NY);
        } else if (q < nA + nB) {
            const int64_t qq = q - nA, per = nx2 * nzp;
            const int side = (int)(qq / per); const int64_t r = qq - side * per;
            j = side ? g.NY - 1 : 0; i = 1 + (int)(r % nx2); kl = k0 + (int)(r / nx2);
        } else {
            const int64_t qq = q - nA - nB, per = nx2 * ny2;
            const int side = (int)(qq / per); const int64_t r = qq - side * per;
            kl = (side == 0 && zlow) ? -g.zs : g.NZ - 1 - g.zs;
            i = 1 + (int)(r % nx2); j = 1 + (int)(r / nx2);
        }
        const int64_t ln = (int64_t)kl * g.npl + (int64_t)j * g.NX + i;
        const int type = node_class(i, g.NX) + 3 * node_class(j, g.NY) + 9 * node_class(kl + g.zs, g.NZ);
        const unsigned own = nodemask[g.G + ln];
        const double *Tt = Tg + type * 243;
        // a warp whose nodes are all of one class (the inside of a face) takes the table from the constant bank: the
        // kernel is bound by load instructions, and these are two thirds of them
        const unsigned am = __activemask();
        const int type0 = __shfl_sync(am, type, __ffs(am) - 1);
        const bool one_class = __all_sync(am, type == type0);
        double a0 = 0., a1 = 0., a2 = 0.;
        // a node with all three dofs prescribed is an identity row (the bending case fixes the whole x faces, whose
        // stride-NX neighbourhoods are the expensive ones here): no stencil
#pragma unroll 9                                            // nine slots' loads in flight together (the kernel is latency-bound)
        for (int sl = 0; sl < (own == 7u ? 0 : 27); ++sl) {
            const int dx = sl % 3 - 1, dy = (sl / 3) % 3 - 1, dz = sl / 9 - 1;
            const int64_t lj = ln + dx + (int64_t)g.NX * dy + g.npl * dz;
            const bool ok = i + dx >= -1 && i + dx <= g.NX && lj >= -(int64_t)g.G && g.G + lj < g.S;
            // unconditional loads from a clamped index (a load predicated on the mask byte would wait for it)
            const int64_t idx = g.G + (ok ? lj : ln);
            const unsigned mk = ok ? (unsigned)nodemask[idx] : 7u;
            const double v0 = __ldg(x + idx), v1 = __ldg(x + g.S + idx), v2 = __ldg(x + 2 * g.S + idx);
            const double x0 = (mk & 1u) ? 0. : v0, x1 = (mk & 2u) ? 0. : v1, x2 = (mk & 4u) ? 0. : v2;
            if (one_class) {
                const double *m = c_T + type0 * 243 + sl * 9;
                a0 = fma(m[2], x2, fma(m[1], x1, fma(m[0], x0, a0)));
                a1 = fma(m[5], x2, fma(m[4], x1, fma(m[3], x0, a1)));
                a2 = fma(m[8], x2, fma(m[7], x1, fma(m[6], x0, a2)));
            } else {
                const double *m = Tt + sl * 9;
                a0 = fma(__ldg(m + 2), x2, fma(__ldg(m + 1), x1, fma(__ldg(m + 0), x0, a0)));
                a1 = fma(__ldg(m + 5), x2, fma(__ldg(m + 4), x1, fma(__ldg(m + 3), x0, a1)));
                a2 = fma(__ldg(m + 8), x2, fma(__ldg(m + 7), x1, fma(__ldg(m + 6), x0, a2)));
            }
        }
        const double *xo = x + g.G + ln;
        const double p0 = __ldg(xo), p1 = __ldg(xo + g.S), p2 = __ldg(xo + 2 * g.S);
        if (own & 1u) a0 = p0;                              // Dirichlet rows are identity rows
        if (own & 2u) a1 = p1;
        if (own & 4u) a2 = p2;
        double *y0 = y + g.G + ln;
        y0[0] = a0; y0[g.S] = a1; y0[2 * g.S] = a2;
        if (DOT && owned_node(g, ln)) dot += a0 * p0 + a1 * p1 + a2 * p2;
    }
    if (DOT) {
        double s = block_sum<4>(dot, sm);
        if (threadIdx.x == 0) partial[blockIdx.x] = s;
    }
}

}  // namespace macroc
