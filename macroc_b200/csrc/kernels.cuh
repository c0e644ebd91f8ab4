// kernels.cuh -- fp64 CUDA kernels (sm_100a) for MacroC's macro-scale hot path.
//
// Data layout in HBM (per rank / z-slab):
//   * nodal vectors are SoA by displacement component, v[c*S + G + ln], ln the
//     owned node in DMDA natural order, G >= NX*NY + NX + 1 nodes of padding on
//     both sides that also holds the ghost planes (so every 27-point neighbour
//     is at a uniform linear offset and always in bounds);
//   * the assembled operator is a fixed 27-slot 3x3-block stencil ("index-free
//     block-DIA"): tiles of 32 consecutive owned nodes, inside a tile entry
//     k = slot*9 + 3*r + c of node `lane` sits at double index
//     ((k>>1)*32 + lane)*2 + (k&1); a tile is one contiguous 62 464-byte chunk
//     (244 entries per node, entry 243 is padding), read with one coalesced
//     128-bit load per lane per entry pair.
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

namespace macroc {

constexpr int TILE_NODES = 32;
constexpr int ENTRIES = 243;                 // 27 slots x 9
constexpr int PAIRS = 122;                   // ceil(243/2)
constexpr int TILE_DOUBLES = PAIRS * 2 * TILE_NODES;   // 7808 doubles = 62 464 B

// Signs of the 8 hex nodes / Gauss points in natural coordinates
// (reference include/macroc.h:61-69 and the dsh table of assembly.c:200-232).
__constant__ int c_sgn[8][3] = {{-1, -1, -1}, {+1, -1, -1}, {+1, +1, -1}, {-1, +1, -1},
                                {-1, -1, +1}, {+1, -1, +1}, {+1, +1, +1}, {-1, +1, +1}};
// dsh[gp][n][d]: shape-function derivatives of the unit-cube element
// (assembly.c:198-232), filled once by k_init_dsh and then read-only.
__constant__ __align__(16) double c_dsh[8][8][3];
__constant__ double c_D[36];                 // homogenised tangent (row-major 6x6)
// class stencils T[27 classes][27 slots][3][3] (see k_stencil_table); interior class = 13
__constant__ double c_T[27 * 243];

struct GridDev {
    int NX, NY, NZ;          // LOCAL box extents in x and y (ghost columns/rows of x/y neighbours
                             // included; = the global grid for z-slabs) and the global NZ
    int zs, nzl;             // owned planes
    int64_t npl, nloc;       // local nodes per plane / local nodes of the owned planes
    int64_t S;               // SoA component stride (doubles)
    int G;                   // padding (nodes) in front of local node 0
    int64_t ntiles;
    int Xs, Ys;              // global coordinates of local node (0, 0)
    int ox0, oy0, xm, ym;    // owned sub-box of the local box (DMDAGetCorners, local coordinates)
    const uint8_t *ghost;    // 1 for local nodes owned by an x/y neighbour (excluded from reductions);
                             // nullptr for z-slabs
};

__device__ __forceinline__ bool owned_node(const GridDev &g, int64_t ln)
{
    return g.ghost == nullptr || g.ghost[g.G + ln] == 0;
}

struct CgScalars {
    double beta, betaold, pw, zz, zr, dp, dp0, ttol, rtol, abstol, dtol;
    int its, maxits, done, reason;
};

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Deterministic block reduction (fixed order): returns the block sum in thread 0.
template <int NW>
__device__ __forceinline__ double block_sum(double v, double *sm)
{
    v = warp_sum(v);
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) sm[w] = v;
    __syncthreads();
    double s = 0.;
    if (threadIdx.x == 0)
        for (int q = 0; q < NW; ++q) s += sm[q];
    __syncthreads();
    return s;
}

// ---------------------------------------------------------------------------
// KSPCG scalar bookkeeping (used by the vector kernels further down and by the operator kernels' fused reductions)
// ---------------------------------------------------------------------------
__device__ __forceinline__ void cg_scalars_pw_body(CgScalars *s, double pw)
{
    // KSPSolve_CG: KSPCheckDot(dpi) -> KSP_DIVERGED_NANORINF; dpi == 0 or a sign change against the
    // previous dpi -> KSP_DIVERGED_INDEFINITE_MAT
    const double old = s->pw;
    const int i = s->its;
    s->pw = pw;
    s->its += 1;                               // ksp->its = i+1 at the top of the loop body
    if (!isfinite(pw)) { s->done = 1; s->reason = -9; }
    else if (pw == 0. || (i > 0 && ((pw > 0.) != (old > 0.)))) { s->done = 1; s->reason = -10; }
}

__device__ __forceinline__ void cg_scalars_iter_body(CgScalars *s, double zz, double zr)
{
    double dp = sqrt(zz);
    s->dp = dp;
    if (!isfinite(dp) || !isfinite(zr)) { s->done = 1; s->reason = -9; return; }      // KSP_DIVERGED_NANORINF
    if (dp <= s->ttol) { s->done = 1; s->reason = dp <= s->abstol ? 3 : 2; return; }
    if (dp >= s->dtol * s->dp0) { s->done = 1; s->reason = -4; return; }
    if (s->its >= s->maxits) { s->done = 1; s->reason = -3; return; }
    s->betaold = s->beta;
    s->beta = zr;
    if (zr < 0.) { s->done = 1; s->reason = -8; return; }             // KSP_DIVERGED_INDEFINITE_PC
    if (s->beta == 0.) { s->its += 1; s->done = 1; s->reason = 3; }   // KSP_CONVERGED_ATOL at the next top
}

// Single rank: the reduction of the per-block partials needs no launch of its own.  The block that draws the last
// ticket folds all partials in a fixed order (bit-reproducible for a given grid) and updates the CG scalars; every
// other block has read what it needs of them long before (it draws its ticket after its work).  All threads of a
// block call this after thread 0 has written the block's partial(s).  NV = 1: p.w, NV = 2: (z.z, z.r).
struct CgFuse {
    CgScalars *sc;
    unsigned *ticket;            // nullptr: not fused (several ranks, or a caller that wants the plain sums)
};
template <int NW, int NV>
__device__ __forceinline__ void cg_last_block(const double *partial, int nblk, const CgFuse &f, double *sm /* NW doubles */)
{
    __shared__ unsigned s_last;
    if (threadIdx.x == 0) {
        __threadfence();                                   // the partial before the ticket
        s_last = atomicAdd(f.ticket, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!s_last) return;                                   // block-uniform
    __threadfence();
    double v0 = 0., v1 = 0.;
    for (int q = threadIdx.x; q < nblk; q += blockDim.x) {
        v0 += __ldcg(partial + q);
        if (NV == 2) v1 += __ldcg(partial + nblk + q);
    }
    v0 = block_sum<NW>(v0, sm);
    if (NV == 2) v1 = block_sum<NW>(v1, sm);
    if (threadIdx.x == 0) {
        *f.ticket = 0;
        if (NV == 1) cg_scalars_pw_body(f.sc, v0);
        else cg_scalars_iter_body(f.sc, v0, v1);
    }
}

// ---------------------------------------------------------------------------
// Element constants
// ---------------------------------------------------------------------------

// calc_B's dsh table (assembly.c:195-232) with explicit round-to-nearest ops so
// the bits match the CPU evaluation of the same expressions.
__global__ void k_make_dsh(double *out /* [8][8][3] */, double hx, double hy, double hz)
{
    int t = threadIdx.x;
    if (t >= 64) return;
    int gp = t >> 3, n = t & 7;
    const double CONSTXG = 0.577350269189626;
    double xi = c_sgn[gp][0] * CONSTXG, eta = c_sgn[gp][1] * CONSTXG, zeta = c_sgn[gp][2] * CONSTXG;
    double fx = __dadd_rn(1., c_sgn[n][0] * xi), fy = __dadd_rn(1., c_sgn[n][1] * eta),
           fz = __dadd_rn(1., c_sgn[n][2] * zeta);
    // hx = hy = hz = 1 reproduces the reference (assembly.c:198 shadows the global dx, dy, dz)
    out[(gp * 8 + n) * 3 + 0] = __ddiv_rn(c_sgn[n][0] * __dmul_rn(fy, fz) / 8. * 2., hx);
    out[(gp * 8 + n) * 3 + 1] = __ddiv_rn(c_sgn[n][1] * __dmul_rn(fx, fz) / 8. * 2., hy);
    out[(gp * 8 + n) * 3 + 2] = __ddiv_rn(c_sgn[n][2] * __dmul_rn(fx, fy) / 8. * 2., hz);
}

__device__ __forceinline__ double Bentry(const double *__restrict__ dsh, int gp, int row, int col)
{
    // B[row][3n+d] of assembly.c:234-253
    int n = col / 3, d = col % 3;
    const double *h = dsh + (gp * 8 + n) * 3;
    switch (row) {
        case 0: return d == 0 ? h[0] : 0.;
        case 1: return d == 1 ? h[1] : 0.;
        case 2: return d == 2 ? h[2] : 0.;
        case 3: return d == 0 ? h[1] : (d == 1 ? h[0] : 0.);
        case 4: return d == 0 ? h[2] : (d == 2 ? h[0] : 0.);
        default: return d == 1 ? h[2] : (d == 2 ? h[1] : 0.);
    }
}

// Ke = sum_gp B^T C B wg for a tangent that is the same at the 8 Gauss points
// (assembly.c:87-101).  One thread per entry, the reference's summation order
// (gp, k, l) and rounding (no FMA contraction) -> bitwise the CPU value.
// dsh, D: global-memory copies (the context's own; the __constant__ symbols may be bound to
// another context while this one is being created).
__global__ void k_element_matrix(const double *__restrict__ dsh /* [8][8][3] */, const double *__restrict__ D /* [36] */,
                                 double wg, double *Ke /* [24][24] */)
{
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= 576) return;
    int i = t / 24, j = t % 24;
    double acc = 0.;
    for (int gp = 0; gp < 8; ++gp)
        for (int k = 0; k < 6; ++k) {
            double bki = Bentry(dsh, gp, k, i);
            for (int l = 0; l < 6; ++l) {
                double term = __dmul_rn(__dmul_rn(__dmul_rn(bki, D[k * 6 + l]), Bentry(dsh, gp, l, j)), wg);
                acc = __dadd_rn(acc, term);
            }
        }
    Ke[t] = acc;
}

__device__ __host__ __forceinline__ int local_node_of_pos(int px, int py, int pz)
{
    return (py ? (px ? 2 : 3) : (px ? 1 : 0)) + 4 * pz;
}
// position (0/1 per axis) of local node n inside its element (same table as c_sgn)
__device__ __host__ __forceinline__ constexpr int node_px(int n) { return ((n & 3) == 1 || (n & 3) == 2) ? 1 : 0; }
__device__ __host__ __forceinline__ constexpr int node_py(int n) { return (n & 3) >= 2 ? 1 : 0; }
__device__ __host__ __forceinline__ constexpr int node_pz(int n) { return n >> 2; }

// Pre-summed 27-slot stencils for the 27 node classes (lower face / interior /
// upper face per axis): T[type][slot][3][3] = sum over the elements that exist
// around a node of that class, in increasing element order (the order
// MatSetValuesLocal(ADD_VALUES) accumulates them, assembly.c:85-108).
__global__ void k_stencil_table(const double *__restrict__ Ke, double *__restrict__ T /* [27][27][9] */)
{
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= 27 * 27 * 9) return;
    int cc = t % 3, rr = (t / 3) % 3, slot = (t / 9) % 27, type = t / 243;
    int tx = type % 3, ty = (type / 3) % 3, tz = type / 9;
    int ddx = slot % 3 - 1, ddy = (slot / 3) % 3 - 1, ddz = slot / 9 - 1;
    double acc = 0.;
    for (int oz = -1; oz <= 0; ++oz)
        for (int oy = -1; oy <= 0; ++oy)
            for (int ox = -1; ox <= 0; ++ox) {
                // element with origin node + o exists?
                if ((ox == -1 && tx == 0) || (ox == 0 && tx == 2)) continue;
                if ((oy == -1 && ty == 0) || (oy == 0 && ty == 2)) continue;
                if ((oz == -1 && tz == 0) || (oz == 0 && tz == 2)) continue;
                int bx = ddx - ox, by = ddy - oy, bz = ddz - oz;   // neighbour's position in the element
                if (bx < 0 || bx > 1 || by < 0 || by > 1 || bz < 0 || bz > 1) continue;
                int a = local_node_of_pos(-ox, -oy, -oz), b = local_node_of_pos(bx, by, bz);
                acc = __dadd_rn(acc, Ke[(a * 3 + rr) * 24 + b * 3 + cc]);
            }
    T[t] = acc;
}

__device__ __forceinline__ int node_class(int c, int N) { return c == 0 ? 0 : (c == N - 1 ? 2 : 1); }

// ---------------------------------------------------------------------------
// Jacobian "assembly" for a tangent that is uniform over the grid: expand the
// class stencils into the tile-blocked operator and apply MatZeroRowsColumns
// (A <- M A M + (I-M), bcs.c:341-347) on the fly.  Pure HBM-write bound:
// 1 952 B per node.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_fill_operator(GridDev g, const double *__restrict__ T, const uint8_t *__restrict__ nodemask,
                double2 *__restrict__ A, double *__restrict__ dinv)
{
    int64_t tile = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (tile >= g.ntiles) return;
    int lane = threadIdx.x & 31;
    int64_t ln = tile * TILE_NODES + lane;
    bool valid = ln < g.nloc;
    int i = 0, j = 0, k = 0, type = 13;
    unsigned own = 0;
    if (valid) {
        i = (int)(ln % g.NX);
        j = (int)((ln / g.NX) % g.NY);
        k = (int)(ln / g.npl) + g.zs;
        type = node_class(i, g.NX) + 3 * node_class(j, g.NY) + 9 * node_class(k, g.NZ);
        own = nodemask[g.G + ln];
    }
    const double *Tt = T + type * 243;
    double2 *At = A + tile * (PAIRS * TILE_NODES) + lane;
    double carry = 0.;
    double diag[3] = {1., 1., 1.};
#pragma unroll
    for (int kk = 0; kk < 244; ++kk) {
        double v = 0.;
        if (kk < ENTRIES) {
            const int slot = kk / 9, rr = (kk % 9) / 3, cc = kk % 3;
            const int ddx = slot % 3 - 1, ddy = (slot / 3) % 3 - 1, ddz = slot / 9 - 1;
            if (valid) {
                v = __ldg(Tt + kk);
                unsigned nb = nodemask[g.G + ln + ddx + (int64_t)g.NX * ddy + g.npl * ddz];
                if (((own >> rr) & 1u) || ((nb >> cc) & 1u)) v = (slot == 13 && rr == cc) ? 1. : 0.;
                if (slot == 13 && rr == cc) diag[rr] = v;
            }
        }
        if (kk & 1) At[(kk >> 1) * TILE_NODES] = make_double2(carry, v);
        else carry = v;
    }
    if (valid) {
#pragma unroll
        for (int d = 0; d < 3; ++d) dinv[d * g.S + g.G + ln] = diag[d] != 0. ? 1. / diag[d] : 1.;   // PCJACOBI
    }
}

// ---------------------------------------------------------------------------
// Assembled block-stencil SpMV  w = A p  (+ fused partial of p.w)
// HBM bound: 1 952 B of operator per node against 48 B of vectors.
// ---------------------------------------------------------------------------
template <bool DOT, int MINB>
__global__ void __launch_bounds__(256, MINB)
k_spmv(GridDev g, const double2 *__restrict__ A, const double *__restrict__ p, double *__restrict__ w,
       int64_t tile0, int64_t ntiles, double *__restrict__ partial, const int *__restrict__ done)
{
    __shared__ double sm[8];
    if (done && *done) return;
    const int lane = threadIdx.x & 31;
    const int64_t wstride = (int64_t)gridDim.x * (blockDim.x >> 5);
    const int64_t NX = g.NX, npl = g.npl;
    double dot = 0.;
    for (int64_t tile = tile0 + (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
         tile < tile0 + ntiles; tile += wstride) {
        int64_t ln = tile * TILE_NODES + lane;
        const double2 *At = A + tile * (PAIRS * TILE_NODES) + lane;
        const double *p0 = p + g.G + ln, *p1 = p0 + g.S, *p2 = p1 + g.S;
        double a0 = 0., a1 = 0., a2 = 0., pc0 = 0., pc1 = 0., pc2 = 0.;
        // two slots (18 entries = 9 pairs) per step; slot 26 + padding at the end
#pragma unroll
        for (int gq = 0; gq < 13; ++gq) {
            double2 v[9];
#pragma unroll
            for (int q = 0; q < 9; ++q) v[q] = __ldcs(At + (gq * 9 + q) * TILE_NODES);
            const double *e = reinterpret_cast<const double *>(v);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int slot = 2 * gq + h;
                const int ddx = slot % 3 - 1, ddy = (slot / 3) % 3 - 1, ddz = slot / 9 - 1;
                const int64_t off = ddx + NX * ddy + npl * ddz;
                double x0 = __ldg(p0 + off), x1 = __ldg(p1 + off), x2 = __ldg(p2 + off);
                if (slot == 13) { pc0 = x0; pc1 = x1; pc2 = x2; }
                const double *m = e + 9 * h;
                a0 = fma(m[0], x0, a0); a0 = fma(m[1], x1, a0); a0 = fma(m[2], x2, a0);
                a1 = fma(m[3], x0, a1); a1 = fma(m[4], x1, a1); a1 = fma(m[5], x2, a1);
                a2 = fma(m[6], x0, a2); a2 = fma(m[7], x1, a2); a2 = fma(m[8], x2, a2);
            }
        }
        {
            double2 v[5];
#pragma unroll
            for (int q = 0; q < 5; ++q) v[q] = __ldcs(At + (117 + q) * TILE_NODES);
            const double *m = reinterpret_cast<const double *>(v);
            const int64_t off = 1 + NX + npl;    // slot 26 = (+1,+1,+1)
            double x0 = __ldg(p0 + off), x1 = __ldg(p1 + off), x2 = __ldg(p2 + off);
            a0 = fma(m[0], x0, a0); a0 = fma(m[1], x1, a0); a0 = fma(m[2], x2, a0);
            a1 = fma(m[3], x0, a1); a1 = fma(m[4], x1, a1); a1 = fma(m[5], x2, a1);
            a2 = fma(m[6], x0, a2); a2 = fma(m[7], x1, a2); a2 = fma(m[8], x2, a2);
        }
        if (ln < g.nloc) {
            double *w0 = w + g.G + ln;
            w0[0] = a0; w0[g.S] = a1; w0[2 * g.S] = a2;
            if (DOT && owned_node(g, ln)) dot += a0 * pc0 + a1 * pc1 + a2 * pc2;
        }
    }
    if (DOT) {
        double s = block_sum<8>(dot, sm);
        if (threadIdx.x == 0) partial[blockIdx.x] = s;
    }
}

// ---------------------------------------------------------------------------
// Matrix-free apply  y = (M K M + I - M) x  with the class stencils
// (27 x 243 doubles): 16 B/DOF of HBM traffic, FP64-pipe / issue bound.
// ---------------------------------------------------------------------------
// 3-D tiled matrix-free apply  y = (M K M + I - M) x  on planes [k0, k1) of the slab.
// A CTA owns MF_TX x MF_TY x MF_TZ nodes and stages the (TX+2)(TY+2)(TZ+2) x 3 patch of M x in
// shared memory once (Dirichlet columns are zeroed while staging, so each x value is read
// from L2 ~1.5 times instead of 27).  A thread computes a column of MF_TY nodes in y, so every
// shared-memory operand feeds up to three output nodes (40 LDS.64 per node instead of 81);
// for interior warps the stencil entries are immediate constant-bank operands of the DFMAs.
constexpr int MF_TX = 32, MF_TY = 4, MF_TZ = 8;
// patch pitch: PX = TX + 2 rounded so that PY*PX = 8 (mod 16) doubles -- a warp is 8 (x) x 4 (z)
// lanes and its four z-rows then fall into disjoint shared-memory bank halves (no conflicts)
constexpr int MF_PX = MF_TX + 4, MF_PY = MF_TY + 2, MF_PZ = MF_TZ + 2;
static_assert((MF_PX * MF_PY) % 16 == 8, "bank-conflict-free z pitch");
constexpr int MF_PATCH = MF_PX * MF_PY * MF_PZ;
constexpr int MF_THREADS = MF_TX * MF_TZ;
constexpr int MF_SMEM = 3 * MF_PATCH * (int)sizeof(double);   // the staged patch of M x

template <bool DOT>
__global__ void __launch_bounds__(MF_THREADS, 2)
k_apply_mf3d(GridDev g, const double *__restrict__ Tg, const uint8_t *__restrict__ nodemask,
             const double *__restrict__ x, double *__restrict__ y,
             int k0, int k1, int tiles_x, int tiles_y, double *__restrict__ partial /* all partials of this apply */,
             int part0 /* this kernel's first slot (the face kernel's come before) */, const int *__restrict__ done, CgFuse fuse)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double(*sx)[MF_PATCH] = reinterpret_cast<double(*)[MF_PATCH]>(smem_raw);
    __shared__ double sm[8];
    if (done && *done) return;
    // warp w = x-block (w % 4) of 8 nodes, z-block (w / 4) of 4 planes; lane = 8 (x) x 4 (z): only
    // 2 of the 32 x-blocks of a row touch the x faces (a 32-wide warp would put 2 of 8 there)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tx = (warp & 3) * 8 + (lane & 7), tz = (warp >> 2) * 4 + (lane >> 3);
    const int64_t ntile = (int64_t)tiles_x * tiles_y * ((k1 - k0 + MF_TZ - 1) / MF_TZ);
    double dot = 0.;
    for (int64_t t = blockIdx.x; t < ntile; t += gridDim.x) {
        const int i0 = (int)(t % tiles_x) * MF_TX, j0 = (int)((t / tiles_x) % tiles_y) * MF_TY;
        const int kk0 = k0 + (int)(t / ((int64_t)tiles_x * tiles_y)) * MF_TZ;       // slab-local plane
        __syncthreads();
        {
            // all loads of the patch are issued before the first use (independent of the masks)
            constexpr int NIT = (MF_PATCH + MF_THREADS - 1) / MF_THREADS;
            double v0[NIT], v1[NIT], v2[NIT];
            unsigned mk[NIT];
#pragma unroll
            for (int it = 0; it < NIT; ++it) {
                const int q = threadIdx.x + it * MF_THREADS;
                const int px = q % MF_PX, py = (q / MF_PX) % MF_PY, pz = q / (MF_PX * MF_PY);
                const int i = i0 + px - 1;
                const int64_t ln = (int64_t)(kk0 + pz - 1) * g.npl + (int64_t)(j0 + py - 1) * g.NX + i;
                const bool ok = q < MF_PATCH && i <= g.NX && ln >= -(int64_t)g.G && g.G + ln < g.S;
                const int64_t idx = g.G + (ok ? ln : 0);
                mk[it] = ok ? nodemask[idx] : 7u;
                v0[it] = __ldg(x + idx); v1[it] = __ldg(x + g.S + idx); v2[it] = __ldg(x + 2 * g.S + idx);
            }
#pragma unroll
            for (int it = 0; it < NIT; ++it) {
                const int q = threadIdx.x + it * MF_THREADS;
                if (q < MF_PATCH) {
                    sx[0][q] = (mk[it] & 1u) ? 0. : v0[it];
                    sx[1][q] = (mk[it] & 2u) ? 0. : v1[it];
                    sx[2][q] = (mk[it] & 4u) ? 0. : v2[it];
                }
            }
        }
        __syncthreads();
        const int i = i0 + tx, kl = kk0 + tz;
        const bool col_valid = i < g.NX && kl < k1;
        const int cx = node_class(i, g.NX), cz = node_class(kl + g.zs, g.NZ);
        int type[MF_TY];
#pragma unroll
        for (int jj = 0; jj < MF_TY; ++jj) {
            const int j = j0 + jj;
            type[jj] = (col_valid && j < g.NY) ? cx + 3 * node_class(j, g.NY) + 9 * cz : 13;
        }
        double acc[MF_TY][3];
#pragma unroll
        for (int jj = 0; jj < MF_TY; ++jj) acc[jj][0] = acc[jj][1] = acc[jj][2] = 0.;
        const int cbase = ((tz + 1) * MF_PY) * MF_PX + tx + 1;          // patch row py = 0 of this column
        // every node takes the interior stencil here; the nodes on a face of the box (another class) are skipped below and
        // done by k_apply_mf_faces (mf_march.cuh)
#pragma unroll
        for (int dz = -1; dz <= 1; ++dz)
#pragma unroll
            for (int py = 0; py < MF_PY; ++py)
#pragma unroll
                for (int dx = -1; dx <= 1; ++dx) {
                    const int o = cbase + (dz * MF_PY + py) * MF_PX + dx;
                    const double x0 = sx[0][o], x1 = sx[1][o], x2 = sx[2][o];
#pragma unroll
                    for (int jj = 0; jj < MF_TY; ++jj) {
                        const int ddy = py - 1 - jj;
                        if (ddy < -1 || ddy > 1) continue;
                        const int ct = 13 * 243 + ((dz + 1) * 9 + (ddy + 1) * 3 + (dx + 1)) * 9;
                        acc[jj][0] = fma(c_T[ct + 0], x0, acc[jj][0]); acc[jj][0] = fma(c_T[ct + 1], x1, acc[jj][0]); acc[jj][0] = fma(c_T[ct + 2], x2, acc[jj][0]);
                        acc[jj][1] = fma(c_T[ct + 3], x0, acc[jj][1]); acc[jj][1] = fma(c_T[ct + 4], x1, acc[jj][1]); acc[jj][1] = fma(c_T[ct + 5], x2, acc[jj][1]);
                        acc[jj][2] = fma(c_T[ct + 6], x0, acc[jj][2]); acc[jj][2] = fma(c_T[ct + 7], x1, acc[jj][2]); acc[jj][2] = fma(c_T[ct + 8], x2, acc[jj][2]);
                    }
                }
        if (col_valid) {
#pragma unroll
            for (int jj = 0; jj < MF_TY; ++jj) {
                const int j = j0 + jj;
                if (j >= g.NY || type[jj] != 13) continue;
                const int64_t ln = (int64_t)kl * g.npl + (int64_t)j * g.NX + i;
                const unsigned own = nodemask[g.G + ln];
                const int c0 = cbase + (jj + 1) * MF_PX;
                double xc0 = sx[0][c0], xc1 = sx[1][c0], xc2 = sx[2][c0];
                double a0 = acc[jj][0], a1 = acc[jj][1], a2 = acc[jj][2];
                if (own) {                       // Dirichlet rows are identity rows: y = x (unmasked)
                    const double *xo = x + g.G + ln;
                    if (own & 1u) { xc0 = __ldg(xo); a0 = xc0; }
                    if (own & 2u) { xc1 = __ldg(xo + g.S); a1 = xc1; }
                    if (own & 4u) { xc2 = __ldg(xo + 2 * g.S); a2 = xc2; }
                }
                double *y0 = y + g.G + ln;
                y0[0] = a0; y0[g.S] = a1; y0[2 * g.S] = a2;
                if (DOT && owned_node(g, ln)) dot += a0 * xc0 + a1 * xc1 + a2 * xc2;
            }
        }
    }
    if (DOT) {
        double s = block_sum<MF_THREADS / 32>(dot, sm);
        if (threadIdx.x == 0) partial[part0 + blockIdx.x] = s;
        if (fuse.ticket) cg_last_block<MF_THREADS / 32, 1>(partial, part0 + gridDim.x, fuse, sm);
    }
}

// Jacobi diagonal for the matrix-free operator (PCJACOBI needs diag(A)).
__global__ void k_mf_diag(GridDev g, const double *__restrict__ T, const uint8_t *__restrict__ nodemask,
                          double *__restrict__ dinv)
{
    int64_t ln = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (ln >= g.nloc) return;
    int i = (int)(ln % g.NX), j = (int)((ln / g.NX) % g.NY), k = (int)(ln / g.npl) + g.zs;
    int type = node_class(i, g.NX) + 3 * node_class(j, g.NY) + 9 * node_class(k, g.NZ);
    unsigned own = nodemask[g.G + ln];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        double v = ((own >> d) & 1u) ? 1. : T[type * 243 + 13 * 9 + d * 4];
        dinv[d * g.S + g.G + ln] = v != 0. ? 1. / v : 1.;
    }
}

// ---------------------------------------------------------------------------
// Residual  b = -(sum_e sum_gp B^T sigma wg), Dirichlet rows -> 0, + |b|^2
// (set_strains assembly.c:25-66, sigma = D eps, assembly_res :120-176).
// The kernels are in assembly_elem.cuh (element forces -> per-node gather in
// increasing element order: no atomics, no colouring, no reverse halo -- the
// element layer above the slab is integrated redundantly from the upper ghost
// plane of u).  The element-level helpers below are shared with them.
// ---------------------------------------------------------------------------
__device__ __forceinline__ void element_strain(const double (&ue)[8][3], int gp, double (&eps)[6])
{
    double e0 = 0., e1 = 0., e2 = 0., e3 = 0., e4 = 0., e5 = 0.;
#pragma unroll
    for (int n = 0; n < 8; ++n) {
        const double hx = c_dsh[gp][n][0], hy = c_dsh[gp][n][1], hz = c_dsh[gp][n][2];
        e0 = fma(hx, ue[n][0], e0);
        e1 = fma(hy, ue[n][1], e1);
        e2 = fma(hz, ue[n][2], e2);
        e3 = fma(hy, ue[n][0], e3); e3 = fma(hx, ue[n][1], e3);
        e4 = fma(hz, ue[n][0], e4); e4 = fma(hx, ue[n][2], e4);
        e5 = fma(hz, ue[n][1], e5); e5 = fma(hy, ue[n][2], e5);
    }
    eps[0] = e0; eps[1] = e1; eps[2] = e2; eps[3] = e3; eps[4] = e4; eps[5] = e5;
}

__device__ __forceinline__ void stress_of(const double (&eps)[6], double (&sig)[6])
{
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        double t = 0.;
#pragma unroll
        for (int j = 0; j < 6; ++j) t = fma(c_D[i * 6 + j], eps[j], t);
        sig[i] = t;
    }
}

__device__ __forceinline__ void gather_element(const double *__restrict__ u, const GridDev &g,
                                               int64_t base /* SoA index of element origin */,
                                               double (&ue)[8][3])
{
    const int64_t NX = g.NX, npl = g.npl;
#pragma unroll
    for (int n = 0; n < 8; ++n) {
        const int64_t q = base + node_px(n) + NX * node_py(n) + npl * node_pz(n);
#pragma unroll
        for (int d = 0; d < 3; ++d) ue[n][d] = __ldg(u + d * g.S + q);
    }
}

// set_strains with materialisation.  Device layout of the Gauss-point arrays is SoA over the
// rank's stored elements (coalesced for every kernel that walks elements with consecutive
// lanes):  strain/stress[(gp*6 + i)*ne_ext + ie],  ctan[((gp*6 + k)*6 + l)*ne_ext + ie],
// ie = rank-local element in DMDAGetElements order.  The reference's gpi = ie*8+gp AoS view
// (assembly.c:58,91,148) is what the host-facing copies of the C ABI speak (k_gp_aos_soa).
__global__ void __launch_bounds__(128)
k_strain_stress(GridDev g, int ezs, int nez, int64_t ne_ext, const double *__restrict__ u,
                double *__restrict__ strain, double *__restrict__ stress)
{
    int64_t ie = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t nex = g.NX - 1, ney = g.NY - 1;
    if (ie >= nex * ney * nez) return;
    int ei = (int)(ie % nex), ej = (int)((ie / nex) % ney), ek = (int)(ie / (nex * ney)) + ezs;
    int64_t base = g.G + ei + (int64_t)g.NX * ej + g.npl * (ek - g.zs);
    double ue[8][3];
    gather_element(u, g, base, ue);
#pragma unroll
    for (int gp = 0; gp < 8; ++gp) {
        double eps[6], sig[6];
        element_strain(ue, gp, eps);
        stress_of(eps, sig);
#pragma unroll
        for (int q = 0; q < 6; ++q) {
            strain[(gp * 6 + q) * ne_ext + ie] = eps[q];
            if (stress) stress[(gp * 6 + q) * ne_ext + ie] = sig[q];
        }
    }
}

// AoS (gpi-major, [ie][gp][n], ie over the rank's DMDA-OWNED elements in DMDAGetElements order)
// <-> SoA ([(gp*n + i)][le], pitch ne_ext, le over the LOCAL box's elements) for n = 6 or 36.
// (lex0, ley0) = first owned element in local-box coordinates; for z-slabs owned == local.
__global__ void k_gp_aos_soa(int n, int64_t onex, int64_t oney, int64_t nez, int lex0, int ley0, int64_t lnex,
                             int64_t lney, int64_t ne_ext, const double *__restrict__ in, double *__restrict__ out,
                             int to_soa)
{
    const int64_t ne = onex * oney * nez;
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ne * 8 * n) return;
    const int64_t ie = t % ne, q = t / ne;
    const int64_t ex = ie % onex, ey = (ie / onex) % oney, ez = ie / (onex * oney);
    const int64_t le = (lex0 + ex) + lnex * ((ley0 + ey) + lney * ez);
    if (to_soa) out[q * ne_ext + le] = in[ie * 8 * n + q];
    else out[ie * 8 * n + q] = in[q * ne_ext + le];
}

// one element layer (normal to `axis`) of a SoA Gauss-point array <-> contiguous buffer
// (Gauss-point halo).  Local elements are indexed ex + lnex*(ey + lney*ez), ez < nlay.
__global__ void k_gp_face_copy(int nq, int64_t lnex, int64_t lney, int64_t nlay, int axis, int64_t layer,
                               int64_t ne_ext, double *__restrict__ arr, double *__restrict__ buf, int pack)
{
    const int64_t d0 = axis == 0 ? lney : lnex, d1 = axis == 2 ? lney : nlay;   // the two in-face extents
    const int64_t face = d0 * d1;
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= face * nq) return;
    const int64_t f = t % face, q = t / face, a0 = f % d0, a1 = f / d0;
    int64_t ex, ey, ez;
    if (axis == 0) { ex = layer; ey = a0; ez = a1; }
    else if (axis == 1) { ex = a0; ey = layer; ez = a1; }
    else { ex = a0; ey = a1; ez = layer; }
    const int64_t e = ex + lnex * (ey + lney * ez);
    if (pack) buf[t] = arr[q * ne_ext + e];
    else arr[q * ne_ext + e] = buf[t];
}

// forces.c:58-106 / :115-166: sum of the 8 Gauss-point stresses of the elements
// next to the loaded boundary, component [3]*dy*dz (bending) or [1]*dx*dz (circle).
__global__ void __launch_bounds__(128)
k_force(GridDev g, int ezs, int nez, int lex0, int ley0, int onex, int oney, int bc_type, double dx, double dy,
        double dz, double lx, double lz, double rad, const double *__restrict__ u,
        const double *__restrict__ stress_gp, int64_t ne_ext, double *__restrict__ partial)
{
    // (lex0, ley0): first DMDA-owned element of the rank in local-box coordinates; onex x oney
    // owned elements per layer (forces.c:69,126); nex, ney below are the local box's elements
    __shared__ double sm[4];
    int64_t nex = g.NX - 1, ney = g.NY - 1;
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    double f = 0.;
    int64_t count = bc_type == 0 ? (int64_t)oney * nez : (int64_t)onex * nez;
    if (t < count) {
        int ei, ej, ek;
        bool take = true;
        if (bc_type == 0) { ei = lex0 + onex - 1; ej = ley0 + (int)(t % oney); ek = (int)(t / oney) + ezs; }
        else {
            ei = lex0 + (int)(t % onex); ej = ley0 + oney - 1; ek = (int)(t / onex) + ezs;
            // forces.c:138-141: ghost start + local element index
            int sk = ezs;
            double xx = __dsub_rn(lx / 2., __dadd_rn(__dmul_rn((double)(g.Xs + ei), dx), dx / 2.));
            double zz = __dsub_rn(lz / 2., __dadd_rn(__dmul_rn((double)(sk + (ek - ezs)), dz), dz / 2.));
            take = (__dadd_rn(__dmul_rn(xx, xx), __dmul_rn(zz, zz))) < rad * rad;
        }
        if (take) {
            double ave = 0.;
            const int comp = bc_type == 0 ? 3 : 1;
            if (stress_gp) {                      // Gauss-point stresses of the material plug-in (forces.c:85,149)
                const int64_t e = ei + nex * (ej + ney * (int64_t)(ek - ezs));
                for (int gp = 0; gp < 8; ++gp) ave += stress_gp[(gp * 6 + comp) * ne_ext + e];
            } else {
                int64_t base = g.G + ei + (int64_t)g.NX * ej + g.npl * (ek - g.zs);
                double ue[8][3];
                gather_element(u, g, base, ue);
#pragma unroll
                for (int gp = 0; gp < 8; ++gp) {
                    double eps[6], sig[6];
                    element_strain(ue, gp, eps);
                    stress_of(eps, sig);
                    ave += sig[comp];
                }
            }
            f = bc_type == 0 ? ave * dy * dz : ave * dx * dz;
        }
    }
    double s = block_sum<4>(f, sm);
    if (threadIdx.x == 0) partial[blockIdx.x] = s;
}

// ---------------------------------------------------------------------------
// Small vector kernels (owned range only; SoA, blockIdx.y = component)
// ---------------------------------------------------------------------------
__global__ void k_scatter_bc(const int64_t *__restrict__ idx, const double *__restrict__ coef, int n, double U,
                             double *__restrict__ u)
{
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) u[idx[t]] = coef[t] * U;            // VecSetValues(INSERT) bcs.c:85,140
}

__global__ void k_axpy1(GridDev g, double *__restrict__ y, const double *__restrict__ x)
{
    int64_t ln = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (ln >= g.nloc) return;
    int64_t q = blockIdx.y * g.S + g.G + ln;
    y[q] += 1. * x[q];                              // VecAXPY(u, 1., du) main.c:79
}

// boundary layout (owned box, x fastest, dof-interleaved = the rank's part of the DMDA global
// vector) <-> local SoA arrays
__device__ __forceinline__ int64_t owned_to_local(const GridDev &g, int64_t on)
{
    const int64_t i = on % g.xm, j = (on / g.xm) % g.ym, k = on / ((int64_t)g.xm * g.ym);
    return (g.ox0 + i) + (int64_t)g.NX * (g.oy0 + j) + g.npl * k;
}

__global__ void k_aos_to_soa(GridDev g, const double *__restrict__ in, double *__restrict__ out)
{
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= 3 * (int64_t)g.xm * g.ym * g.nzl) return;
    int64_t ln = owned_to_local(g, e / 3); int d = (int)(e % 3);
    out[d * g.S + g.G + ln] = in[e];
}

__global__ void k_soa_to_aos(GridDev g, const double *__restrict__ in, double *__restrict__ out)
{
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= 3 * (int64_t)g.xm * g.ym * g.nzl) return;
    int64_t ln = owned_to_local(g, e / 3); int d = (int)(e % 3);
    out[e] = in[d * g.S + g.G + ln];
}

// owned + ghost planes -> interleaved host layout; node0 may be negative (lower ghost plane)
__global__ void k_soa_to_aos_range(GridDev g, const double *__restrict__ in, int64_t node0, int64_t nnodes,
                                   double *__restrict__ out)
{
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= 3 * nnodes) return;
    int64_t ln = node0 + e / 3; int d = (int)(e % 3);
    out[e] = in[d * g.S + g.G + ln];
}

__global__ void k_export_blocks(GridDev g, const double *__restrict__ A, int64_t node0, int64_t nnodes,
                                double *__restrict__ out /* [nnodes][243] */)
{
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nnodes * 243) return;
    int64_t ln = owned_to_local(g, node0 + e / 243); int kk = (int)(e % 243);   // node0, nnodes: owned-box order
    int64_t tile = ln / TILE_NODES; int lane = (int)(ln % TILE_NODES);
    out[e] = A[tile * TILE_DOUBLES + ((int64_t)(kk >> 1) * TILE_NODES + lane) * 2 + (kk & 1)];
}

// x / y halo of the general DMDA box: one column (x) or row (y) of the owned planes, all three
// components, to / from a contiguous buffer.  axis 0: buf[(c*nzl + k)*NY + j] <-> v(idx, j, k);
// axis 1: buf[(c*nzl + k)*NX + i] <-> v(i, idx, k).
__global__ void k_halo_pack_xy(GridDev g, double *__restrict__ v, double *__restrict__ buf, int axis, int idx, int pack)
{
    const int64_t len = axis == 0 ? g.NY : g.NX, n = 3 * (int64_t)g.nzl * len;
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const int64_t a = t % len, k = (t / len) % g.nzl, c = t / (len * g.nzl);
    const int64_t ln = axis == 0 ? idx + g.NX * a + g.npl * k : a + (int64_t)g.NX * idx + g.npl * k;
    if (pack) buf[t] = v[c * g.S + g.G + ln];
    else v[c * g.S + g.G + ln] = buf[t];
}

// fixed-order reduction of per-block partials: out[0..nout) = sum over blocks
__global__ void k_reduce(const double *__restrict__ partial, int nblocks, double *__restrict__ out)
{
    __shared__ double sm[8];
    double s = 0.;
    for (int q = threadIdx.x; q < nblocks; q += blockDim.x) s += partial[q];
    s = block_sum<8>(s, sm);
    if (threadIdx.x == 0) out[0] = s;
}

__global__ void k_fill(double *p, int64_t n, double v)
{
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e < n) p[e] = v;
}

// deterministic synthetic vector of SURVEY.md 8d: x = sin(0.37*gdof) + 0.1 on free dofs
__global__ void k_fill_pattern(GridDev g, const uint8_t *__restrict__ nodemask, double *__restrict__ v)
{
    int64_t ln = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (ln >= g.nloc) return;
    unsigned own = nodemask[g.G + ln];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        double gd = (double)((ln + (int64_t)g.zs * g.npl) * 3 + d) + 1e3 * g.Xs + 1e5 * g.Ys;
        v[d * g.S + g.G + ln] = ((own >> d) & 1u) ? 0. : sin(0.37 * gd) + 0.1;
    }
}

// ---------------------------------------------------------------------------
// PCG (KSPCG + PCJACOBI, PETSc semantics -- SURVEY.md 8c item 6)
// ---------------------------------------------------------------------------

// The vector kernels below walk the owned range two nodes at a time with 128-bit accesses
// (S and G are multiples of 32, so component bases are 16-byte aligned); npair = ceil(nloc/2),
// the odd tail element is handled by predication inside the pair.

// r = b, x = 0, z = r*dinv; partials of z.z and z.r
__global__ void __launch_bounds__(256)
k_cg_init(GridDev g, const double *__restrict__ b, const double *__restrict__ dinv, double *__restrict__ x,
          double *__restrict__ r, double *__restrict__ partial /* [2][nblk] */, int nblk)
{
    __shared__ double sm[8];
    double zz = 0., zr = 0.;
    const int64_t npair = (g.nloc + 1) >> 1;
    for (int64_t pr = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; pr < npair; pr += (int64_t)gridDim.x * blockDim.x) {
        const bool two = 2 * pr + 1 < g.nloc;
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            const int64_t q = d * g.S + g.G + 2 * pr;
            const double2 bv = *reinterpret_cast<const double2 *>(b + q), dv = *reinterpret_cast<const double2 *>(dinv + q);
            double2 rv = bv;
            if (!two) rv.y = 0.;
            const double z0 = rv.x * dv.x, z1 = two ? rv.y * dv.y : 0.;
            if (two) {
                *reinterpret_cast<double2 *>(x + q) = make_double2(0., 0.);
                *reinterpret_cast<double2 *>(r + q) = rv;
            } else { x[q] = 0.; r[q] = rv.x; }
            if (g.ghost) {       // x/y ghost rows do not count
                if (g.ghost[g.G + 2 * pr] == 0) { zz = fma(z0, z0, zz); zr = fma(z0, rv.x, zr); }
                if (two && g.ghost[g.G + 2 * pr + 1] == 0) { zz = fma(z1, z1, zz); zr = fma(z1, rv.y, zr); }
            } else {
                zz = fma(z0, z0, zz); zr = fma(z0, rv.x, zr);
                zz = fma(z1, z1, zz); zr = fma(z1, rv.y, zr);
            }
        }
    }
    double s0 = block_sum<8>(zz, sm), s1 = block_sum<8>(zr, sm);
    if (threadIdx.x == 0) { partial[blockIdx.x] = s0; partial[nblk + blockIdx.x] = s1; }
}

// scalar bookkeeping after the initial residual (KSPSolve_CG prologue +
// KSPConvergedDefault at it 0)
__device__ __forceinline__ void cg_scalars_init_body(CgScalars *s, double zz, double zr)
{
    double dp = sqrt(zz);
    s->dp = s->dp0 = dp;
    s->ttol = fmax(s->rtol * dp, s->abstol);
    s->beta = zr; s->betaold = 1.;
    s->its = 0; s->done = 0; s->reason = 0;
    if (!isfinite(dp) || !isfinite(zr)) { s->done = 1; s->reason = -9; }             // KSPCheckNorm / KSPCheckDot
    else if (dp <= s->ttol) { s->done = 1; s->reason = dp <= s->abstol ? 3 : 2; }
    else if (s->beta == 0.) { s->its = 1; s->done = 1; s->reason = 3; }
}
__global__ void k_cg_scalars_init(CgScalars *s, const double *sums /* zz, zr */)
{
    cg_scalars_init_body(s, sums[0], sums[1]);
}

// p = z + (beta/betaold) p   (first iteration: p = z), z = r*dinv recomputed
__global__ void __launch_bounds__(256)
k_cg_update_p(GridDev g, const CgScalars *__restrict__ s, const double *__restrict__ r,
              const double *__restrict__ dinv, double *__restrict__ p)
{
    if (s->done) return;
    const bool first = s->its == 0;
    const double bb = s->beta / s->betaold;
    const int64_t npair = (g.nloc + 1) >> 1;
    for (int64_t pr = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; pr < npair; pr += (int64_t)gridDim.x * blockDim.x) {
        const bool two = 2 * pr + 1 < g.nloc;
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            const int64_t q = d * g.S + g.G + 2 * pr;
            const double2 rv = *reinterpret_cast<const double2 *>(r + q), dv = *reinterpret_cast<const double2 *>(dinv + q);
            double2 pv = *reinterpret_cast<const double2 *>(p + q);
            const double z0 = rv.x * dv.x, z1 = rv.y * dv.y;
            pv.x = first ? z0 : fma(bb, pv.x, z0);
            pv.y = first ? z1 : fma(bb, pv.y, z1);
            if (two) *reinterpret_cast<double2 *>(p + q) = pv;
            else p[q] = pv.x;
        }
    }
}

// a = beta/(p.w); x += a p; r -= a w; z = r*dinv; partials z.z, z.r
__global__ void __launch_bounds__(256)
k_cg_update_xr(GridDev g, const CgScalars *__restrict__ s, const double *__restrict__ p,
               const double *__restrict__ w, const double *__restrict__ dinv, double *__restrict__ x,
               double *__restrict__ r, double *__restrict__ partial, int nblk, CgFuse fuse)
{
    __shared__ double sm[8];
    if (s->done) return;
    const double pw = s->pw;
    if (pw == 0.) return;                      // KSP_DIVERGED_INDEFINITE_MAT, flagged by the scalar update
    const double a = s->beta / pw;
    double zz = 0., zr = 0.;
    const int64_t npair = (g.nloc + 1) >> 1;
    for (int64_t pr = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; pr < npair; pr += (int64_t)gridDim.x * blockDim.x) {
        const bool two = 2 * pr + 1 < g.nloc;
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            const int64_t q = d * g.S + g.G + 2 * pr;
            const double2 pv = *reinterpret_cast<const double2 *>(p + q), wv = *reinterpret_cast<const double2 *>(w + q);
            const double2 dv = *reinterpret_cast<const double2 *>(dinv + q);
            double2 xv = *reinterpret_cast<const double2 *>(x + q), rv = *reinterpret_cast<const double2 *>(r + q);
            xv.x = fma(a, pv.x, xv.x); xv.y = fma(a, pv.y, xv.y);
            rv.x = fma(-a, wv.x, rv.x); rv.y = fma(-a, wv.y, rv.y);
            if (two) {
                *reinterpret_cast<double2 *>(x + q) = xv;
                *reinterpret_cast<double2 *>(r + q) = rv;
            } else { x[q] = xv.x; r[q] = rv.x; rv.y = 0.; }
            const double z0 = rv.x * dv.x, z1 = two ? rv.y * dv.y : 0.;
            if (g.ghost) {
                if (g.ghost[g.G + 2 * pr] == 0) { zz = fma(z0, z0, zz); zr = fma(z0, rv.x, zr); }
                if (two && g.ghost[g.G + 2 * pr + 1] == 0) { zz = fma(z1, z1, zz); zr = fma(z1, rv.y, zr); }
            } else {
                zz = fma(z0, z0, zz); zr = fma(z0, rv.x, zr);
                zz = fma(z1, z1, zz); zr = fma(z1, rv.y, zr);
            }
        }
    }
    double s0 = block_sum<8>(zz, sm), s1 = block_sum<8>(zr, sm);
    if (threadIdx.x == 0) { partial[blockIdx.x] = s0; partial[nblk + blockIdx.x] = s1; }
    if (fuse.ticket) cg_last_block<8, 2>(partial, nblk, fuse, sm);
}

// after the p.w reduction: store it (and flag an indefinite matrix)
__global__ void k_cg_scalars_pw(CgScalars *s, const double *sum)
{
    if (s->done) return;
    cg_scalars_pw_body(s, sum[0]);
}

// after the (z.z, z.r) reduction: convergence test and beta rotation
__global__ void k_cg_scalars_iter(CgScalars *s, const double *sums)
{
    if (s->done) return;
    cg_scalars_iter_body(s, sums[0], sums[1]);
}

// single-rank fast path: partial reduction and scalar update in one launch
__global__ void k_cg_reduce_pw(const double *__restrict__ partial, int nblk, CgScalars *s)
{
    __shared__ double sm[8];
    if (s->done) return;
    double v = 0.;
    for (int q = threadIdx.x; q < nblk; q += blockDim.x) v += partial[q];
    v = block_sum<8>(v, sm);
    if (threadIdx.x == 0) cg_scalars_pw_body(s, v);
}

__global__ void k_cg_reduce_iter(const double *__restrict__ partial, int nblk, CgScalars *s)
{
    __shared__ double sm[8];
    if (s->done) return;
    double s0 = 0., s1 = 0.;
    for (int q = threadIdx.x; q < nblk; q += blockDim.x) { s0 += partial[q]; s1 += partial[nblk + q]; }
    s0 = block_sum<8>(s0, sm);
    s1 = block_sum<8>(s1, sm);
    if (threadIdx.x == 0) cg_scalars_iter_body(s, s0, s1);
}

// reduce two partial arrays at once: out[0], out[1]
__global__ void k_reduce2(const double *__restrict__ partial, int nblk, double *__restrict__ out)
{
    __shared__ double sm[8];
    double s0 = 0., s1 = 0.;
    for (int q = threadIdx.x; q < nblk; q += blockDim.x) { s0 += partial[q]; s1 += partial[nblk + q]; }
    s0 = block_sum<8>(s0, sm);
    s1 = block_sum<8>(s1, sm);
    if (threadIdx.x == 0) { out[0] = s0; out[1] = s1; }
}

// FP64 pipe probe: NCHAIN independent DFMA chains per thread, register operands only.  Used to
// MEASURE the DFMA rate the FP64-bound kernels (matrix-free apply, element assembly) are compared with.
constexpr int FP64_PROBE_CHAINS = 16, FP64_PROBE_ITERS = 4096;
__global__ void __launch_bounds__(256)
k_fp64_probe(double *out, double seed)
{
    double acc[FP64_PROBE_CHAINS];
    const double m = 1.0 + seed * 1e-9, b = 1e-9 * (threadIdx.x + 1);
#pragma unroll
    for (int q = 0; q < FP64_PROBE_CHAINS; ++q) acc[q] = seed + q;
    for (int it = 0; it < FP64_PROBE_ITERS; ++it) {
#pragma unroll
        for (int q = 0; q < FP64_PROBE_CHAINS; ++q) acc[q] = fma(acc[q], m, b);
    }
    double s = 0.;
#pragma unroll
    for (int q = 0; q < FP64_PROBE_CHAINS; ++q) s += acc[q];
    if (s == 12345.678) out[0] = s;            // never true: keeps the chains alive
}

// the same probe with the multiplier taken from the constant bank (the form the unrolled element kernel uses)
__global__ void __launch_bounds__(256)
k_fp64_probe_const(double *out, double seed)
{
    double acc[FP64_PROBE_CHAINS];
    const double b = 1e-9 * (threadIdx.x + 1);
#pragma unroll
    for (int q = 0; q < FP64_PROBE_CHAINS; ++q) acc[q] = seed + q;
    for (int it = 0; it < FP64_PROBE_ITERS; ++it) {
#pragma unroll
        for (int q = 0; q < FP64_PROBE_CHAINS; ++q) acc[q] = fma(acc[q], c_dsh[q & 7][(q >> 1) & 7][q % 3], b);
    }
    double s = 0.;
#pragma unroll
    for (int q = 0; q < FP64_PROBE_CHAINS; ++q) s += acc[q];
    if (s == 12345.678) out[0] = s;
}

}  // namespace macroc
