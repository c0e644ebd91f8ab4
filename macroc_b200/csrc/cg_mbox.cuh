// cg_mbox.cuh -- the CG dot-product all-reduces as part of the reduction kernels themselves.
//
// KSPCG needs two global reductions per iteration (p.w, then z.z and z.r; reference: the three
// VecDot/VecNorm inside KSPSolve, src/assembly.c:185).  With NCCL each one costs a reduce kernel, an
// ncclAllReduce of 1-2 doubles and a 1-thread scalar kernel, all latency on the critical path.
// Here every rank owns a MAILBOX in its HBM that all ranks of the node can write (cudaIpc peer
// mapping over NVLink).  The one-block kernel that folds the per-CTA partials
//   1. stores the rank's partial sums, then a sequence number, into slot [seq & 1][me] of EVERY
//      rank's mailbox (st.global over the peer mapping, __threadfence_system between data and flag),
//   2. polls its own mailbox until all N slots carry the sequence number,
//   3. adds the N contributions in rank order (every rank computes bit-identical sums) and updates
//      the CG scalars -- no separate all-reduce, no separate scalar kernel.
// Two parities are enough: a rank can be at most one reduction ahead of the slowest one, because
// finishing reduction k needs everybody's contribution k.  The poll is bounded (about ten
// seconds): a missing rank ends the solve with reason MACROC_KSP_DIVERGED_COMM instead of hanging.
// Ranks are processes on different GPUs: the kernels that wait for one another never share a device
// (the in-process loopback communicator keeps its host-side sum).
#pragma once

#include "kernels.cuh"

namespace macroc {

constexpr int MBOX_MAX_RANKS = 64;
constexpr int MBOX_SLOT_DOUBLES = 4;                  // v0, v1, sequence number (as bits), pad: 32 B
constexpr int KSP_DIVERGED_COMM = -100;

struct MboxDev {
    double *const *peer;      // device table: peer[q] = base of rank q's mailbox as mapped into this process
    int me, n;
};

__device__ __forceinline__ void st_release_sys_u64(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys_u64(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// Called by ALL threads of a one-block kernel (blockDim.x >= n).  in0/in1: this rank's partial sums
// (valid in thread 0).  Returns true and the rank-ordered global sums in out[0..1] (all threads),
// false if a rank stayed silent.
__device__ __forceinline__ bool mbox_allreduce2(const MboxDev &mb, unsigned long long seq, double in0, double in1, double (&out)[2])
{
    __shared__ double sh_in[2];
    __shared__ double sh_v[2][MBOX_MAX_RANKS];
    __shared__ int sh_fail;
    const int t = threadIdx.x;
    if (t == 0) { sh_in[0] = in0; sh_in[1] = in1; sh_fail = 0; }
    __syncthreads();
    const int par = (int)(seq & 1ull);
    if (t < mb.n) {
        // my contribution -> rank t's mailbox, slot [par][me]
        double *slot = mb.peer[t] + ((size_t)par * mb.n + mb.me) * MBOX_SLOT_DOUBLES;
        slot[0] = sh_in[0];
        slot[1] = sh_in[1];
        __threadfence_system();
        st_release_sys_u64(reinterpret_cast<unsigned long long *>(slot + 2), seq);
        // rank t's contribution <- my mailbox, slot [par][t]
        const double *mine = mb.peer[mb.me] + ((size_t)par * mb.n + t) * MBOX_SLOT_DOUBLES;
        const long long t0 = clock64();
        bool ok = true;
        while (ld_acquire_sys_u64(reinterpret_cast<const unsigned long long *>(mine + 2)) != seq) {
            if (clock64() - t0 > (1ll << 34)) { ok = false; break; }
            __nanosleep(64);
        }
        if (ok) {
            sh_v[0][t] = reinterpret_cast<const volatile double *>(mine)[0];
            sh_v[1][t] = reinterpret_cast<const volatile double *>(mine)[1];
        } else
            sh_fail = 1;
    }
    __syncthreads();
    double s0 = 0., s1 = 0.;
    for (int q = 0; q < mb.n; ++q) { s0 += sh_v[0][q]; s1 += sh_v[1][q]; }     // rank order on every rank
    out[0] = s0; out[1] = s1;
    return sh_fail == 0;
}

// partials of p.w -> global p.w -> CG scalars (replaces k_reduce + ncclAllReduce + k_cg_scalars_pw)
__global__ void __launch_bounds__(256)
k_cg_reduce_pw_mbox(const double *__restrict__ partial, int nblk, CgScalars *s, MboxDev mb, unsigned long long seq)
{
    __shared__ double sm[8];
    if (s->done) return;
    double v = 0.;
    for (int q = threadIdx.x; q < nblk; q += blockDim.x) v += partial[q];
    v = block_sum<8>(v, sm);
    double out[2];
    const bool ok = mbox_allreduce2(mb, seq, v, 0., out);
    if (threadIdx.x == 0) {
        if (ok) cg_scalars_pw_body(s, out[0]);
        else { s->done = 1; s->reason = KSP_DIVERGED_COMM; }
    }
}

// partials of (z.z, z.r) -> global sums -> convergence test and beta rotation
__global__ void __launch_bounds__(256)
k_cg_reduce_iter_mbox(const double *__restrict__ partial, int nblk, CgScalars *s, MboxDev mb, unsigned long long seq)
{
    __shared__ double sm[8];
    if (s->done) return;
    double s0 = 0., s1 = 0.;
    for (int q = threadIdx.x; q < nblk; q += blockDim.x) { s0 += partial[q]; s1 += partial[nblk + q]; }
    s0 = block_sum<8>(s0, sm);
    s1 = block_sum<8>(s1, sm);
    double out[2];
    const bool ok = mbox_allreduce2(mb, seq, s0, s1, out);
    if (threadIdx.x == 0) {
        if (ok) cg_scalars_iter_body(s, out[0], out[1]);
        else { s->done = 1; s->reason = KSP_DIVERGED_COMM; }
    }
}

// KSPSolve prologue: partials of (z.z, z.r) of the initial residual -> k_cg_scalars_init's bookkeeping
__global__ void __launch_bounds__(256)
k_cg_reduce_init_mbox(const double *__restrict__ partial, int nblk, CgScalars *s, MboxDev mb, unsigned long long seq)
{
    __shared__ double sm[8];
    double s0 = 0., s1 = 0.;
    for (int q = threadIdx.x; q < nblk; q += blockDim.x) { s0 += partial[q]; s1 += partial[nblk + q]; }
    s0 = block_sum<8>(s0, sm);
    s1 = block_sum<8>(s1, sm);
    double out[2];
    const bool ok = mbox_allreduce2(mb, seq, s0, s1, out);
    if (threadIdx.x == 0) {
        if (ok) cg_scalars_init_body(s, out[0], out[1]);
        else { s->its = 0; s->done = 1; s->reason = KSP_DIVERGED_COMM; }
    }
}

}  // namespace macroc
