// dmma_ab.cuh -- MEASUREMENT kernels, not part of the solve path.
//
// north_star: "fp64 DMMA tensor cores are used for the 24x24 element contractions only if ncu shows
// a win over FFMA".  These kernels answer that with numbers (macroc_contraction_ab, macroc_dmma_probe):
//   * k_dmma_probe:   the DMMA issue rate of the device (mma.sync.m8n8k4.f64, SASS DMMA.8x8x4);
//   * k_ab_dfma:      Ke = sum_gp B^T C_gp B of every element from its per-Gauss-point tangents with the
//                     sparsity-aware DFMA form of the product kernels (integrate_gp: 2 160 FMA per Gauss
//                     point: B has 9 non-zeros per node);
//   * k_ab_dmma:      the same contraction as dense 8x8x4 tensor-core tiles: W^T = B^T C (24x6, 6 DMMA),
//                     Ke += W^T B (24x24, 18 DMMA) per Gauss point -- 6 144 FMA-equivalents, of which the
//                     zero padding of K = 6 -> 8 and of B's structural zeros is most.
// Both stage the tangents the same way (cp.async, all 8 Gauss points of 32 consecutive elements, double
// buffered), both write the same result: 24 weighted row sums of Ke per element (coalesced) and, for the
// first n_full elements, the whole matrix (tests compare it with B^T D B from macroc_calc_B).
#pragma once

#include "assembly_elem.cuh"

namespace macroc {

constexpr int AB_ELEMS = 32;                           // elements per batch
constexpr int AB_PITCH = 33;                           // doubles between two tangent entries of the batch (bank spread for the DMMA reads)
constexpr int AB_BUF = 288 * AB_PITCH;                 // 8 Gauss points x 36 entries
constexpr int AB_DFMA_THREADS = 24 * 32;
constexpr int AB_DMMA_WARPS = 16;
constexpr int AB_SMEM_DFMA = 2 * AB_BUF * 8;
constexpr int AB_SMEM_DMMA = 2 * AB_BUF * 8 + 8 * 8 * 24 * 8 + 24 * 32 * 8;     // + B table + row-sum transposition

__device__ __forceinline__ void ab_stage(double *Cs, const double *__restrict__ ctan, int64_t pitch, int64_t e0, int64_t ne)
{
    for (int q = threadIdx.x; q < 288 * AB_ELEMS; q += blockDim.x) {
        const int entry = q >> 5, l = q & 31;
        if (e0 + l < ne) cp_async8(Cs + entry * AB_PITCH + l, ctan + (int64_t)entry * pitch + e0 + l);
        else Cs[entry * AB_PITCH + l] = 0.;
    }
    cp_async_commit();
}

// sparsity-aware DFMA form: warp (a, d) = row d of the 3 x 24 row block of local node a, lane = element
__global__ void __launch_bounds__(AB_DFMA_THREADS, 1)
k_ab_dfma(const double *__restrict__ ctan, int64_t pitch, int64_t ne, double wg, double *__restrict__ rowsum,
          double *__restrict__ full, int n_full)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *Cs = reinterpret_cast<double *>(smem_raw);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int a = warp / 3, d = warp - 3 * a;
    const int64_t nbatch = (ne + AB_ELEMS - 1) / AB_ELEMS;
    int buf = 0;
    if ((int64_t)blockIdx.x < nbatch) ab_stage(Cs, ctan, pitch, (int64_t)blockIdx.x * AB_ELEMS, ne);
    for (int64_t bt = blockIdx.x; bt < nbatch; bt += gridDim.x, buf ^= 1) {
        const int64_t nxt = bt + gridDim.x;
        if (nxt < nbatch) { ab_stage(Cs + (buf ^ 1) * AB_BUF, ctan, pitch, nxt * AB_ELEMS, ne); asm volatile("cp.async.wait_group 1;" ::: "memory"); }
        else asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();
        const double *ck = Cs + buf * AB_BUF + lane;
        double blk[24];
#pragma unroll
        for (int q = 0; q < 24; ++q) blk[q] = 0.;
#pragma unroll 1
        for (int gp = 0; gp < 8; ++gp) {
            if (d == 0) integrate_gp<true, 0, -1>(gp, a, ck + gp * 36 * AB_PITCH, AB_PITCH, blk);
            else if (d == 1) integrate_gp<true, 1, -1>(gp, a, ck + gp * 36 * AB_PITCH, AB_PITCH, blk);
            else integrate_gp<true, 2, -1>(gp, a, ck + gp * 36 * AB_PITCH, AB_PITCH, blk);
        }
        const int64_t e = bt * AB_ELEMS + lane;
        if (e < ne) {
            double s = 0.;
#pragma unroll
            for (int q = 0; q < 24; ++q) s = fma(blk[q], (double)(q + 1), s);
            rowsum[(int64_t)(3 * a + d) * ne + e] = s * wg;
            if (e < n_full)
#pragma unroll
                for (int q = 0; q < 24; ++q) full[e * 576 + (3 * a + d) * 24 + q] = blk[q] * wg;
        }
        __syncthreads();
    }
}

__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// dense tensor-core form: one warp per element.  Fragments of mma.m8n8k4 (gid = lane / 4, t = lane % 4):
// A[gid][t], B[t][gid], C[gid][2t], C[gid][2t+1].  The k index of k-step s is permuted to 2t + s, so the
// accumulator fragments of W^T = B^T C (columns 2t, 2t+1) ARE the A fragments of the second product and the
// fragments of B^T serve as A operand of the first and as B operand of the second product: no shuffles.
template <bool UPPER>
__global__ void __launch_bounds__(AB_DMMA_WARPS * 32, 1)
k_ab_dmma(const double *__restrict__ ctan, int64_t pitch, int64_t ne, double wg, double *__restrict__ rowsum,
          double *__restrict__ full, int n_full)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *Cs = reinterpret_cast<double *>(smem_raw);
    double *Bt = Cs + 2 * AB_BUF;                      // [gp][k 0..7][col 0..23], rows 6, 7 zero
    double *rows = Bt + 8 * 8 * 24;                    // [24][32]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, gid = lane >> 2, t = lane & 3;
    for (int q = threadIdx.x; q < 8 * 8 * 24; q += blockDim.x) {
        const int gp = q / 192, k = (q / 24) % 8, col = q % 24;
        Bt[q] = k < 6 ? Bentry(&c_dsh[0][0][0], gp, k, col) : 0.;
    }
    const int64_t nbatch = (ne + AB_ELEMS - 1) / AB_ELEMS;
    int buf = 0;
    if ((int64_t)blockIdx.x < nbatch) ab_stage(Cs, ctan, pitch, (int64_t)blockIdx.x * AB_ELEMS, ne);
    for (int64_t bt = blockIdx.x; bt < nbatch; bt += gridDim.x, buf ^= 1) {
        const int64_t nxt = bt + gridDim.x;
        if (nxt < nbatch) { ab_stage(Cs + (buf ^ 1) * AB_BUF, ctan, pitch, nxt * AB_ELEMS, ne); asm volatile("cp.async.wait_group 1;" ::: "memory"); }
        else asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();
        const double *Cb = Cs + buf * AB_BUF;
        for (int l = warp; l < AB_ELEMS; l += AB_DMMA_WARPS) {
            double ke[3][3][2];
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
                for (int j = 0; j < 3; ++j) ke[i][j][0] = ke[i][j][1] = 0.;
#pragma unroll 1
            for (int gp = 0; gp < 8; ++gp) {
                double af[3][2], cf[2], wt[3][2];
#pragma unroll
                for (int s = 0; s < 2; ++s) {
                    const int k = 2 * t + s;
#pragma unroll
                    for (int i = 0; i < 3; ++i) af[i][s] = Bt[(gp * 8 + k) * 24 + 8 * i + gid];
                    cf[s] = (k < 6 && gid < 6) ? Cb[(gp * 36 + k * 6 + gid) * AB_PITCH + l] : 0.;
                }
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    wt[i][0] = wt[i][1] = 0.;
                    dmma884(wt[i][0], wt[i][1], af[i][0], cf[0]);
                    dmma884(wt[i][0], wt[i][1], af[i][1], cf[1]);
                }
#pragma unroll
                for (int i = 0; i < 3; ++i)
#pragma unroll
                    for (int j = 0; j < 3; ++j) {
                        if (UPPER && j < i) continue;              // symmetric tangent: the lower tiles are mirrors
                        dmma884(ke[i][j][0], ke[i][j][1], wt[i][0], af[j][0]);
                        dmma884(ke[i][j][0], ke[i][j][1], wt[i][1], af[j][1]);
                    }
            }
            const int64_t e = bt * AB_ELEMS + l;
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                double s = 0.;
#pragma unroll
                for (int j = 0; j < 3; ++j)
#pragma unroll
                    for (int u = 0; u < 2; ++u) s = fma(ke[i][j][u], (double)(8 * j + 2 * t + u + 1), s);
                s += __shfl_xor_sync(0xffffffffu, s, 1);
                s += __shfl_xor_sync(0xffffffffu, s, 2);
                if (t == 0) rows[(8 * i + gid) * 32 + l] = s * wg;
                if (e < n_full)
#pragma unroll
                    for (int j = 0; j < 3; ++j)
#pragma unroll
                        for (int u = 0; u < 2; ++u) full[e * 576 + (8 * i + gid) * 24 + 8 * j + 2 * t + u] = ke[i][j][u] * wg;
            }
        }
        __syncthreads();
        for (int q = threadIdx.x; q < 24 * 32; q += blockDim.x) {
            const int64_t e = bt * AB_ELEMS + (q & 31);
            if (e < ne) rowsum[(int64_t)(q >> 5) * ne + e] = rows[q];
        }
        __syncthreads();
    }
}

// DMMA issue rate: 8 independent accumulator tiles per warp
constexpr int DMMA_PROBE_ITERS = 2048, DMMA_PROBE_CHAINS = 8;
__global__ void __launch_bounds__(256)
k_dmma_probe(double *out, double seed)
{
    double c[DMMA_PROBE_CHAINS][2];
    const double a = 1.0 + 1e-9 * threadIdx.x, b = seed * 1e-9;
#pragma unroll
    for (int q = 0; q < DMMA_PROBE_CHAINS; ++q) c[q][0] = c[q][1] = seed + q;
    for (int it = 0; it < DMMA_PROBE_ITERS; ++it) {
#pragma unroll
        for (int q = 0; q < DMMA_PROBE_CHAINS; ++q) dmma884(c[q][0], c[q][1], a, b);
    }
    double s = 0.;
#pragma unroll
    for (int q = 0; q < DMMA_PROBE_CHAINS; ++q) s += c[q][0] + c[q][1];
    if (s == 12345.678) out[0] = s;            // never true: keeps the chains alive
}

}  // namespace macroc
