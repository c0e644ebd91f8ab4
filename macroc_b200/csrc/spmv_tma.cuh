// spmv_tma.cuh -- assembled block-stencil SpMV with the operator streamed by the
// TMA engine (cp.async.bulk global -> shared, mbarrier completion) instead of
// per-lane LDG: the bytes in flight per SM are set by the shared-memory ring
// (WARPS x NSTAGE x 4.5 KB), not by how many loads the resident warps can keep
// outstanding in registers.
//
// Every warp owns a private ring of NSTAGE stages and a private set of
// mbarriers, so there is no CTA-wide synchronisation in the steady state: lane
// 0 re-arms a stage (arrive.expect_tx + one bulk copy of the next 4 608-byte
// chunk of the warp's tile sequence) as soon as the warp has consumed it.
// A tile (32 nodes, 62 464 B) is 13 chunks of two stencil slots (9 entry pairs
// x 32 lanes x 16 B) plus one of 5 pairs (slot 26 + padding).
#pragma once

#include "kernels.cuh"

namespace macroc {

constexpr int CHUNK_PAIRS = 9;
constexpr int CHUNK_BYTES = CHUNK_PAIRS * TILE_NODES * 16;       // 4608
constexpr int LAST_CHUNK_BYTES = 5 * TILE_NODES * 16;            // 2560
constexpr int CHUNKS_PER_TILE = 14;
constexpr int TILE_BYTES = TILE_DOUBLES * 8;                     // 62464

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ uint64_t l2_evict_first_policy()
{
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
// TMA bulk copy global -> shared::cta, completion counted in bytes on `bar`
__device__ __forceinline__ void tma_load_bulk(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar,
                                              uint64_t policy)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
        ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "l"(policy) : "memory");
}

template <int WARPS, int NSTAGE>
struct SpmvTmaSmem {
    static constexpr int ring_bytes = WARPS * NSTAGE * CHUNK_BYTES;
    static constexpr int bar_bytes = WARPS * NSTAGE * 8;
    static constexpr int red_bytes = WARPS * 8;
    static constexpr int total = ring_bytes + bar_bytes + red_bytes;
};

template <int WARPS, int NSTAGE, bool DOT>
__global__ void __launch_bounds__(WARPS * 32, 1)
k_spmv_tma(GridDev g, const double2 *__restrict__ A, const double *__restrict__ p, double *__restrict__ w,
           int64_t tile0, int64_t ntiles, double *__restrict__ partial, const int *__restrict__ done, CgFuse fuse)
{
    static_assert(NSTAGE >= 2 && NSTAGE < CHUNKS_PER_TILE, "ring depth");
    extern __shared__ __align__(128) unsigned char smem_ring[];
    if (done && *done) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned char *ring = smem_ring + (size_t)warp * NSTAGE * CHUNK_BYTES;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem_ring + SpmvTmaSmem<WARPS, NSTAGE>::ring_bytes) + warp * NSTAGE;
    double *red = reinterpret_cast<double *>(smem_ring + SpmvTmaSmem<WARPS, NSTAGE>::ring_bytes +
                                             SpmvTmaSmem<WARPS, NSTAGE>::bar_bytes);
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < NSTAGE; ++s) mbar_init(&bars[s], 1);
        fence_barrier_init();
    }
    __syncwarp();

    const int64_t wstride = (int64_t)gridDim.x * WARPS;
    const int64_t first = tile0 + (int64_t)blockIdx.x * WARPS + warp;
    const int64_t last = tile0 + ntiles;
    const int64_t my_tiles = first < last ? (last - first + wstride - 1) / wstride : 0;
    const int64_t my_chunks = my_tiles * CHUNKS_PER_TILE;
    const unsigned char *Ab = reinterpret_cast<const unsigned char *>(A);
    const uint64_t policy = l2_evict_first_policy();

    // issue chunk `c` of this warp's sequence (lane 0 only)
    auto issue = [&](int64_t c) {
        const int64_t t = c / CHUNKS_PER_TILE;
        const int ch = (int)(c - t * CHUNKS_PER_TILE);
        const int stage = (int)(c % NSTAGE);
        const uint32_t bytes = ch < CHUNKS_PER_TILE - 1 ? CHUNK_BYTES : LAST_CHUNK_BYTES;
        const unsigned char *src = Ab + (first + t * wstride) * (int64_t)TILE_BYTES + (int64_t)ch * CHUNK_BYTES;
        mbar_arrive_expect_tx(&bars[stage], bytes);
        tma_load_bulk(ring + stage * CHUNK_BYTES, src, bytes, &bars[stage], policy);
    };
    if (lane == 0)
        for (int64_t c = 0; c < NSTAGE && c < my_chunks; ++c) issue(c);

    const int64_t NX = g.NX, npl = g.npl;
    double dot = 0.;
    int64_t c = 0;                          // chunk counter of this warp
    for (int64_t t = 0; t < my_tiles; ++t) {
        const int64_t tile = first + t * wstride;
        const int64_t ln = tile * TILE_NODES + lane;
        const double *p0 = p + g.G + ln, *p1 = p0 + g.S, *p2 = p1 + g.S;
        double a0 = 0., a1 = 0., a2 = 0., pc0 = 0., pc1 = 0., pc2 = 0.;
#pragma unroll
        for (int ch = 0; ch < CHUNKS_PER_TILE; ++ch, ++c) {
            const int stage = (int)(c % NSTAGE);
            const uint32_t parity = (uint32_t)((c / NSTAGE) & 1);
            // the vector operands of this chunk's slots do not depend on the operator: fetch them
            // before blocking on the barrier
            double xv[2][3];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int slot = 2 * ch + h;
                if (slot < 27) {
                    const int ddx = slot % 3 - 1, ddy = (slot / 3) % 3 - 1, ddz = slot / 9 - 1;
                    const int64_t off = ddx + NX * ddy + npl * ddz;
                    xv[h][0] = __ldg(p0 + off); xv[h][1] = __ldg(p1 + off); xv[h][2] = __ldg(p2 + off);
                    if (slot == 13) { pc0 = xv[h][0]; pc1 = xv[h][1]; pc2 = xv[h][2]; }
                }
            }
            mbar_wait(&bars[stage], parity);
            const double2 *sv = reinterpret_cast<const double2 *>(ring + stage * CHUNK_BYTES) + lane;
            double2 v[9];
#pragma unroll
            for (int q = 0; q < 9; ++q)
                if (ch < CHUNKS_PER_TILE - 1 || q < 5) v[q] = sv[q * TILE_NODES];
            const double *e = reinterpret_cast<const double *>(v);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int slot = 2 * ch + h;
                if (slot < 27) {
                    const double *m = e + 9 * h;
                    const double x0 = xv[h][0], x1 = xv[h][1], x2 = xv[h][2];
                    a0 = fma(m[0], x0, a0); a0 = fma(m[1], x1, a0); a0 = fma(m[2], x2, a0);
                    a1 = fma(m[3], x0, a1); a1 = fma(m[4], x1, a1); a1 = fma(m[5], x2, a1);
                    a2 = fma(m[6], x0, a2); a2 = fma(m[7], x1, a2); a2 = fma(m[8], x2, a2);
                }
            }
            __syncwarp();                   // every lane has read the stage: hand it back to the TMA engine
            if (lane == 0 && c + NSTAGE < my_chunks) issue(c + NSTAGE);
        }
        if (ln < g.nloc) {
            double *w0 = w + g.G + ln;
            w0[0] = a0; w0[g.S] = a1; w0[2 * g.S] = a2;
            if (DOT && owned_node(g, ln)) dot += a0 * pc0 + a1 * pc1 + a2 * pc2;
        }
    }
    if (DOT) {
        dot = warp_sum(dot);
        if (lane == 0) red[warp] = dot;
        __syncthreads();
        if (threadIdx.x == 0) {
            double s = 0.;
#pragma unroll
            for (int q = 0; q < WARPS; ++q) s += red[q];
            partial[blockIdx.x] = s;
        }
        if (fuse.ticket) cg_last_block<WARPS, 1>(partial, gridDim.x, fuse, red);
    }
}

}  // namespace macroc
