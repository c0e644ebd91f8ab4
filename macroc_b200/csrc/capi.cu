// capi.cu -- the C ABI of include/macroc_b200.h over the kernels in kernels.cuh.
//
// One macroc_ctx replaces the reference's file-scope globals
// (include/macroc.h:71-128): it owns the slab's device vectors u, du, b, the
// assembled operator, the Dirichlet bookkeeping and the KSP work vectors, a
// CUDA stream and (for nranks > 1) an NCCL communicator used for the z-plane
// halo exchange and the CG dot-product all-reduces.
// No CPU fallback: compute entry points return MACROC_ERR_NO_DEVICE without a GPU.
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <condition_variable>
#include <mutex>
#include <string>
#include <vector>

#include <cuda_runtime.h>
#include <nccl.h>

#include "../../include/macroc_b200.h"
#include "grid.h"
#include "kernels.cuh"
#include "spmv_tma.cuh"
#include "assembly_elem.cuh"
#include "assembly_node.cuh"
#include "mf_march.cuh"
#include "dmma_ab.cuh"
#include "spmv_sym.cuh"
#include "loopback.h"
#include "cg_mbox.cuh"
#include "strict_fp.cuh"

using namespace macroc;

static thread_local std::string g_last_error;

enum { V_U = 0, V_DU = 1, V_B = 2, V_R = 3, V_P = 4, V_W = 5, V_DINV = 6, V_COUNT = 7 };

struct macroc_ctx {
    macroc_config cfg;
    Slab slab;
    Geometry geo;
    GridDev g;
    int device = 0;
    cudaStream_t stream = nullptr, comm_stream = nullptr;
    cudaEvent_t ev_ready = nullptr, ev_halo = nullptr, ev_t0 = nullptr, ev_t1 = nullptr, ev_chk[2] = {nullptr, nullptr};
    ncclComm_t comm = nullptr;
    LoopGroup *loop = nullptr;       // in-process communicator (loopback.h) instead of NCCL
    // CG all-reduces through peer-mapped mailboxes (cg_mbox.cuh); NCCL all-reduce when the mapping is unavailable
    double *mbox = nullptr;          // this rank's mailbox: [2 parities][nranks][4 doubles]
    double **mbox_peers = nullptr;   // device table of every rank's mailbox as mapped here
    void *mbox_opened[MBOX_MAX_RANKS] = {nullptr};
    bool mbox_on = false;
    unsigned long long mbox_seq = 0;
    double *vec[V_COUNT] = {nullptr};
    double2 *A = nullptr;
    double2 *Asym = nullptr;         // symmetric storage (14 of 27 slots), MACROC_OP_ASSEMBLED_SYM; points at tile 0
    double2 *Asym_alloc = nullptr;   // start of the allocation: the ghost plane below (if any), then the slab
    SymGeom sg = {1, 0};
    unsigned *tickets = nullptr;                   // [2] last-block tickets of the fused single-rank reductions
    CgFuse fuse_pw = {nullptr, nullptr};           // set by apply_operator for the launch it is about to make
    bool fuse_pw_done = false;
    int mf_variant = 0, mf_nseg = 0;               // matrix-free apply: 0 auto, 1 patch form, 2 z-marching (mf_march.cuh); segments override (MACROC_MF_*)
    int asm_colblock = 64, asm_ctas_per_sm = 2;    // per-GP element-Jacobian knobs (MACROC_ASM_*)   // element-Jacobian knobs (MACROC_ASM_*)
    int sym_R = 0, sym_nseg = 0, sym_variant = 0, sym_hint = 0;   // tuning overrides (MACROC_SYM_R / _NSEG / _VARIANT / _HINT, read once at create)
    bool A_valid = false, mf_ready = false, Asym_valid = false;
    double *Ke = nullptr, *T = nullptr;
    uint8_t *nodemask = nullptr, *masksum = nullptr, *ghostflag = nullptr;   // masksum[q] = OR of nodemask[32 q .. 32 q + 31]
    double *xy_halo = nullptr;       // 4 send + 4 receive staging buffers of the x / y halo
    size_t xy_halo_stride = 0;
    double *consts = nullptr;        // device copy of {dsh[192], D[36], T[6561]} for bind_constants
    std::vector<double> consts_host; // the same values on the host (compared on an owner change)
    uint64_t id = 0;
    int64_t *bc_idx = nullptr;
    double *bc_coef = nullptr;
    int nbc = 0;
    double *partial = nullptr;
    int partial_cap = 0;
    double *sums = nullptr;          // device: 4 doubles
    double *sums_host = nullptr;     // pinned: 4 doubles
    CgScalars *sc = nullptr;         // device
    CgScalars *sc_host = nullptr;    // pinned [2]
    double *stage = nullptr;         // device staging, 3*nloc doubles (boundary layout)
    double *strain = nullptr, *stress = nullptr, *ctan = nullptr;   // Gauss-point arrays, gpi = ie*8+gp
    double *gp_halo = nullptr;       // send/recv staging of one element layer of a Gauss-point array
    double *scratch = nullptr;       // element forces of one z-chunk, SoA over elements
    int chunk_planes = 0;
    ElemRange er;
    int64_t ne_owned = 0, ne_ext = 0;
    double *flush = nullptr;
    size_t flush_bytes = 0;
    uint64_t launches = 0;
    int ksp_reason = 0;
    int vec_blocks = 0, spmv_blocks = 0;
    cudaGraphExec_t cg_graph[3] = {nullptr, nullptr, nullptr};   // `check` PCG iterations per launch, per operator
    uint64_t cg_graph_launches[3] = {0, 0, 0};
    int spmv_variant = 10;           // 10: TMA ring 8 warps x 4 stages (default); 11: 8x5, 14: 6x6; 0: per-lane LDG
                                     // (kept for A/B measurements, tools/spmv_sweep.py; MACROC_SPMV_VARIANT)
    cudaEvent_t ev_user[8] = {nullptr};
    // live profile of the operator application (ring of event pairs)
    static constexpr int PROF_RING = 128;
    cudaEvent_t prof_ev[2 * PROF_RING] = {nullptr};
    bool prof_on = false;
    int prof_stride = 1, prof_used = 0;
    int64_t prof_counter = 0, prof_samples = 0;
    double prof_ms = 0.;
    double prof_solve_ms = 0.;       // device time of the PCG solves while the profile is on
    int64_t prof_solve_its = 0;
    std::string err;
};

#define FAIL(ctx, code, ...)                                   \
    do {                                                       \
        char _b[512];                                          \
        snprintf(_b, sizeof(_b), __VA_ARGS__);                 \
        g_last_error = _b;                                     \
        if (ctx) (ctx)->err = _b;                              \
        return (code);                                         \
    } while (0)

#define CU(ctx, call)                                                                       \
    do {                                                                                    \
        cudaError_t _e = (call);                                                            \
        if (_e != cudaSuccess)                                                              \
            FAIL(ctx, MACROC_ERR_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(_e)); \
    } while (0)

#define NC(ctx, call)                                                                       \
    do {                                                                                    \
        ncclResult_t _e = (call);                                                           \
        if (_e != ncclSuccess)                                                              \
            FAIL(ctx, MACROC_ERR_NCCL, "%s:%d %s -> %s", __FILE__, __LINE__, #call, ncclGetErrorString(_e)); \
    } while (0)

#define LAUNCH(ctx, kernel, grid, block, ...)                         \
    do {                                                              \
        kernel<<<(grid), (block), 0, (ctx)->stream>>>(__VA_ARGS__);   \
        (ctx)->launches++;                                            \
    } while (0)

static inline int cdiv64(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// __constant__ symbols (c_dsh, c_D, c_T) are per device and shared by every context of the
// process on that device.  ConstLease is held for the duration of every entry point that
// launches kernels reading them:
//   * contexts whose constants are byte-identical (the ranks of a loopback group, repeated runs
//     of one configuration) share the symbols, also from different host threads;
//   * a context with different constants waits until no such entry point is in flight, drains
//     the device (kernels of the previous owner may still be running: set_strains, assembly_jac
//     and update_u return without a sync), uploads its own values synchronously and takes over.
static std::atomic<uint64_t> g_next_ctx_id{1};
struct ConstSlot {
    std::mutex mu;
    std::condition_variable cv;
    int users = 0;                   // entry points in flight that read the symbols
    uint64_t owner = 0;              // context whose values are bound (0: none)
    std::vector<double> content;     // the bound values
};
static ConstSlot g_const_slot[64];

static int const_acquire(macroc_ctx *c)
{
    ConstSlot &sl = g_const_slot[c->device & 63];
    std::unique_lock<std::mutex> lk(sl.mu);
    if (sl.content != c->consts_host) {
        sl.cv.wait(lk, [&] { return sl.users == 0 || sl.content == c->consts_host; });
        if (sl.content != c->consts_host) {
            cudaError_t e = cudaDeviceSynchronize();          // the previous owner's kernels
            if (e == cudaSuccess) e = cudaMemcpyToSymbol(c_dsh, c->consts_host.data(), sizeof(double) * 192);
            if (e == cudaSuccess) e = cudaMemcpyToSymbol(c_D, c->consts_host.data() + 192, sizeof(double) * 36);
            if (e == cudaSuccess) e = cudaMemcpyToSymbol(c_T, c->consts_host.data() + 228, sizeof(double) * 27 * 243);
            if (e == cudaSuccess) e = cudaDeviceSynchronize();
            if (e != cudaSuccess) {
                sl.content.clear(); sl.owner = 0;
                FAIL(c, MACROC_ERR_CUDA, "binding the element constants: %s", cudaGetErrorString(e));
            }
            sl.content = c->consts_host;
        }
    }
    sl.owner = c->id;
    sl.users++;
    return MACROC_OK;
}
static void const_release(macroc_ctx *c)
{
    ConstSlot &sl = g_const_slot[c->device & 63];
    std::lock_guard<std::mutex> lk(sl.mu);
    if (sl.users > 0) sl.users--;
    if (sl.users == 0) sl.cv.notify_all();
}
struct ConstLease {
    macroc_ctx *c;
    int rc;
    explicit ConstLease(macroc_ctx *ctx) : c(ctx), rc(const_acquire(ctx)) {}
    ~ConstLease() { if (rc == MACROC_OK) const_release(c); }
    ConstLease(const ConstLease &) = delete;
    ConstLease &operator=(const ConstLease &) = delete;
};
#define BIND_CONSTANTS(ctx) ConstLease _lease(ctx); if (_lease.rc) return _lease.rc

extern "C" int macroc_version(void) { return 100; }

extern "C" const char *macroc_last_error(const macroc_ctx *ctx)
{
    return ctx ? ctx->err.c_str() : g_last_error.c_str();
}

extern "C" int macroc_default_config(macroc_config *cfg)
{
    if (!cfg) return MACROC_ERR_ARG;
    memset(cfg, 0, sizeof(*cfg));
    cfg->NX = 40; cfg->NY = 3; cfg->NZ = 40;                 // macroc.h:44-46
    cfg->lx = 50.; cfg->ly = 1.; cfg->lz = 50.;              // macroc.h:47-49
    cfg->bc_type = MACROC_BC_CIRCLE;                         // init.c:64
    cfg->ts = 1; cfg->dt = 0.001; cfg->final_time = 1.0;     // macroc.h:40-43
    cfg->vtu_freq = -1;                                      // macroc.h:42
    cfg->newton_max_its = 5; cfg->newton_min_tol = 1.0e-1; cfg->newton_rel_tol = 1.0e-4;   // macroc.h:36-38
    cfg->ksp_rtol = 1.0e-5; cfg->ksp_abstol = 1.0e-50; cfg->ksp_dtol = 1.0e4; cfg->ksp_maxits = 10000;  // init.c:147-148
    cfg->E = 1.0e7; cfg->nu = 0.25;                          // init.c:31
    cfg->op = MACROC_OP_ASSEMBLED;
    cfg->device = -1;
    return MACROC_OK;
}

extern "C" int macroc_config_from_args(macroc_config *cfg, int argc, const char *const *argv)
{
    if (!cfg) return MACROC_ERR_ARG;
    for (int i = 0; i + 1 < argc; ++i) {
        const char *k = argv[i], *v = argv[i + 1];
        if (!k || k[0] != '-') continue;
        auto is = [&](const char *name) { return strcmp(k, name) == 0; };
        if (is("-da_grid_x")) cfg->NX = atoi(v);
        else if (is("-da_grid_y")) cfg->NY = atoi(v);
        else if (is("-da_grid_z")) cfg->NZ = atoi(v);
        else if (is("-da_processors_x")) cfg->px = atoi(v);
        else if (is("-da_processors_y")) cfg->py = atoi(v);
        else if (is("-da_processors_z")) cfg->pz = atoi(v);
        else if (is("-dt")) cfg->dt = atof(v);
        else if (is("-lx")) cfg->lx = atof(v);
        else if (is("-ly")) cfg->ly = atof(v);
        else if (is("-lz")) cfg->lz = atof(v);
        else if (is("-ts")) cfg->ts = atoi(v);
        else if (is("-vtu_freq")) cfg->vtu_freq = atoi(v);
        else if (is("-newton_min_tol") || is("-new_tol")) cfg->newton_min_tol = atof(v);
        else if (is("-newton_rel_tol")) cfg->newton_rel_tol = atof(v);
        else if (is("-newton_max_its") || is("-new_its")) cfg->newton_max_its = atoi(v);
        else if (is("-bc_type")) cfg->bc_type = atoi(v);
        else if (is("-ksp_rtol")) cfg->ksp_rtol = atof(v);
        else if (is("-ksp_atol")) cfg->ksp_abstol = atof(v);
        else if (is("-ksp_divtol")) cfg->ksp_dtol = atof(v);
        else if (is("-ksp_max_it")) cfg->ksp_maxits = atoi(v);
        else if (is("-micro_mat_1")) {                       // E,nu,Ka,Sy (init.c:82)
            double a = 0, b = 0;
            if (sscanf(v, "%lf,%lf", &a, &b) == 2) { cfg->E = a; cfg->nu = b; }
        } else if (is("-mat_free")) cfg->op = atoi(v) ? MACROC_OP_MATRIX_FREE : MACROC_OP_ASSEMBLED;
        else if (is("-strict_fp")) cfg->strict_fp = atoi(v);
        else if (is("-physical_B")) cfg->physical_B = atoi(v);
        else if (is("-ksp_type")) { if (strcmp(v, "cg") != 0) return MACROC_ERR_UNSUPPORTED; }
        else if (is("-pc_type")) { if (strcmp(v, "jacobi") != 0) return MACROC_ERR_UNSUPPORTED; }
    }
    return MACROC_OK;
}

extern "C" int macroc_partition(const macroc_config *cfg, int rank, int nranks, int32_t out[15])
{
    if (!cfg || !out) return MACROC_ERR_ARG;
    Slab s;
    int rc = make_slab(*cfg, rank, nranks, &s);
    if (rc) return rc;
    int32_t v[15] = {s.xs, s.ys, s.zs, s.xm, s.ym, s.nzl, s.Xs, s.Ys, s.Zs, s.Xm, s.Ym, s.Zm, s.nex, s.ney, s.nez};
    memcpy(out, v, sizeof(v));
    return MACROC_OK;
}

extern "C" int macroc_bc_lists(const macroc_config *cfg, int rank, int nranks, int32_t *idx, double *coef, int32_t *n)
{
    if (!cfg || !n) return MACROC_ERR_ARG;
    Slab s;
    int rc = make_slab(*cfg, rank, nranks, &s);
    if (rc) return rc;
    std::vector<int32_t> ix;
    std::vector<double> cf;
    build_bc_lists(*cfg, s, ix, cf);
    if (idx) memcpy(idx, ix.data(), sizeof(int32_t) * ix.size());
    if (coef) memcpy(coef, cf.data(), sizeof(double) * cf.size());
    *n = (int32_t)ix.size();
    return MACROC_OK;
}

extern "C" int macroc_calc_B(int gp, double *B)
{
    // assembly.c:195-254 (host helper; the device keeps the same table in c_dsh)
    if (gp < 0 || gp >= 8 || !B) return MACROC_ERR_ARG;
    static const int sg[8][3] = {{-1, -1, -1}, {+1, -1, -1}, {+1, +1, -1}, {-1, +1, -1},
                                 {-1, -1, +1}, {+1, -1, +1}, {+1, +1, +1}, {-1, +1, +1}};
    const double c = 0.577350269189626;
    memset(B, 0, sizeof(double) * 6 * 24);
    for (int n = 0; n < 8; ++n) {
        double fx = 1 + sg[n][0] * (sg[gp][0] * c), fy = 1 + sg[n][1] * (sg[gp][1] * c), fz = 1 + sg[n][2] * (sg[gp][2] * c);
        double h0 = sg[n][0] * fy * fz / 8. * 2., h1 = sg[n][1] * fx * fz / 8. * 2., h2 = sg[n][2] * fx * fy / 8. * 2.;
        B[0 * 24 + 3 * n + 0] = h0; B[1 * 24 + 3 * n + 1] = h1; B[2 * 24 + 3 * n + 2] = h2;
        B[3 * 24 + 3 * n + 0] = h1; B[3 * 24 + 3 * n + 1] = h0;
        B[4 * 24 + 3 * n + 0] = h2; B[4 * 24 + 3 * n + 2] = h0;
        B[5 * 24 + 3 * n + 1] = h2; B[5 * 24 + 3 * n + 2] = h1;
    }
    return MACROC_OK;
}

extern "C" int macroc_get_unique_id(void *id128)
{
    if (!id128) return MACROC_ERR_ARG;
    ncclUniqueId id;
    ncclResult_t e = ncclGetUniqueId(&id);
    if (e != ncclSuccess) { g_last_error = ncclGetErrorString(e); return MACROC_ERR_NCCL; }
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    memcpy(id128, &id, 128);
    return MACROC_OK;
}

extern "C" int macroc_loopback_id(int nranks, void *id128)
{
    if (!id128 || nranks < 2) return MACROC_ERR_ARG;
    LoopGroup *grp = new LoopGroup(nranks);
    if (const char *v = getenv("MACROC_LOOPBACK_TIMEOUT")) grp->timeout_s = std::max(1, atoi(v));
    memset(id128, 0, 128);
    memcpy(id128, LoopGroup::MAGIC, 16);
    memcpy((unsigned char *)id128 + 16, &grp, sizeof(grp));
    return MACROC_OK;
}

static void isotropic_D(double E, double nu, double *D)
{
    double lambda = E * nu / ((1. + nu) * (1. - 2. * nu));
    double mu = E / (2. * (1. + nu));
    memset(D, 0, 36 * sizeof(double));
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) D[i * 6 + j] = lambda + (i == j ? 2. * mu : 0.);
    for (int i = 3; i < 6; ++i) D[i * 6 + i] = mu;
}

static int ctx_free(macroc_ctx *c)
{
    if (!c) return MACROC_OK;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    for (void *ptr : c->mbox_opened) if (ptr) cudaIpcCloseMemHandle(ptr);
    cudaFree(c->mbox_peers); cudaFree(c->mbox);
    if (c->comm) ncclCommDestroy(c->comm);
    if (c->loop) {
        // The last member to leave frees the group; a member that leaves early breaks it for the rest.
        // Announcing the departure (left++) is this rank's LAST access to the group: once it is visible
        // another thread may delete the object.
        LoopGroup *grp = c->loop;
        LoopGroup::Member &me = grp->m[(size_t)c->slab.rank];
        if (me.ev_ready) cudaEventDestroy(me.ev_ready);
        if (me.ev_done) cudaEventDestroy(me.ev_done);
        me.ev_ready = me.ev_done = nullptr;
        bool last;
        {
            std::lock_guard<std::mutex> lk(grp->mu);
            grp->left++;
            last = grp->left == grp->joined;
            if (!last) { grp->broken = true; grp->cv.notify_all(); }
        }
        if (last) { if (grp->host_part) cudaFreeHost(grp->host_part); delete grp; }
        c->loop = nullptr;
    }
    for (cudaGraphExec_t ge : c->cg_graph) if (ge) cudaGraphExecDestroy(ge);
    for (int i = 0; i < V_COUNT; ++i) cudaFree(c->vec[i]);
    cudaFree(c->A); cudaFree(c->Asym_alloc); cudaFree(c->Ke); cudaFree(c->T); cudaFree(c->nodemask); cudaFree(c->masksum); cudaFree(c->tickets); cudaFree(c->bc_idx); cudaFree(c->bc_coef);
    cudaFree(c->partial); cudaFree(c->sums); cudaFree(c->sc); cudaFree(c->stage); cudaFree(c->strain); cudaFree(c->stress);
    cudaFree(c->ctan); cudaFree(c->scratch); cudaFree(c->consts); cudaFree(c->gp_halo); cudaFree(c->ghostflag); cudaFree(c->xy_halo);
    cudaFree(c->flush);
    if (c->sums_host) cudaFreeHost(c->sums_host);
    if (c->sc_host) cudaFreeHost(c->sc_host);
    for (cudaEvent_t e : {c->ev_ready, c->ev_halo, c->ev_t0, c->ev_t1, c->ev_chk[0], c->ev_chk[1]})
        if (e) cudaEventDestroy(e);
    for (cudaEvent_t e : c->ev_user) if (e) cudaEventDestroy(e);
    for (cudaEvent_t e : c->prof_ev) if (e) cudaEventDestroy(e);
    if (c->comm_stream) cudaStreamDestroy(c->comm_stream);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
    return MACROC_OK;
}

extern "C" int macroc_destroy(macroc_ctx *ctx) { return ctx_free(ctx); }

// Peer-mapped mailboxes for the CG all-reduces (cg_mbox.cuh): every rank allocates one, the cudaIpc
// handles travel through an ncclAllGather, each rank maps its peers' mailboxes.  Every step is
// optional: if any rank cannot map its peers (no IPC in the container, no P2P between the GPUs,
// MACROC_ALLREDUCE=nccl), ALL ranks fall back to ncclAllReduce -- the decision is itself all-reduced.
static int mbox_setup(macroc_ctx *c)
{
    const int n = c->slab.nranks, me = c->slab.rank;
    int ok = 1;
    if (const char *v = getenv("MACROC_ALLREDUCE")) if (strcmp(v, "nccl") == 0) ok = 0;
    if (n > MBOX_MAX_RANKS) ok = 0;
    const size_t bytes = sizeof(double) * 2 * (size_t)n * MBOX_SLOT_DOUBLES;
    cudaIpcMemHandle_t mine;
    memset(&mine, 0, sizeof(mine));
    unsigned char *stage = nullptr;                   // [n][64] handles + [n] flags
    const size_t hs = sizeof(cudaIpcMemHandle_t);
    CU(c, cudaMalloc(&stage, (size_t)n * hs + sizeof(int) * 2));
    if (ok && (cudaMalloc(&c->mbox, bytes) != cudaSuccess || cudaMemset(c->mbox, 0, bytes) != cudaSuccess ||
               cudaIpcGetMemHandle(&mine, c->mbox) != cudaSuccess)) { cudaGetLastError(); ok = 0; }
    std::vector<unsigned char> all((size_t)n * hs);
    cudaError_t e = cudaMemcpy(stage + (size_t)me * hs, &mine, hs, cudaMemcpyHostToDevice);
    ncclResult_t ne = ncclSuccess;
    if (e == cudaSuccess) ne = ncclAllGather(stage + (size_t)me * hs, stage, hs, ncclUint8, c->comm, c->stream);
    if (e == cudaSuccess && ne == ncclSuccess) e = cudaStreamSynchronize(c->stream);
    if (e == cudaSuccess && ne == ncclSuccess) e = cudaMemcpy(all.data(), stage, (size_t)n * hs, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess || ne != ncclSuccess) { cudaFree(stage); FAIL(c, MACROC_ERR_NCCL, "mailbox setup: exchanging the IPC handles failed"); }
    std::vector<double *> peers((size_t)n, nullptr);
    if (ok) {
        for (int q = 0; q < n && ok; ++q) {
            if (q == me) { peers[(size_t)q] = c->mbox; continue; }
            cudaIpcMemHandle_t h;
            memcpy(&h, all.data() + (size_t)q * hs, hs);
            void *ptr = nullptr;
            if (cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); ok = 0; break; }
            c->mbox_opened[q] = ptr;
            peers[(size_t)q] = (double *)ptr;
        }
    }
    // unanimous?
    int *flag = reinterpret_cast<int *>(stage + (size_t)n * hs);
    e = cudaMemcpy(flag, &ok, sizeof(int), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) ne = ncclAllReduce(flag, flag + 1, 1, ncclInt, ncclMin, c->comm, c->stream);
    if (e == cudaSuccess && ne == ncclSuccess) e = cudaStreamSynchronize(c->stream);
    int all_ok = 0;
    if (e == cudaSuccess && ne == ncclSuccess) e = cudaMemcpy(&all_ok, flag + 1, sizeof(int), cudaMemcpyDeviceToHost);
    cudaFree(stage);
    if (e != cudaSuccess || ne != ncclSuccess) FAIL(c, MACROC_ERR_NCCL, "mailbox setup: agreeing on the all-reduce path failed");
    if (all_ok) {
        CU(c, cudaMalloc(&c->mbox_peers, sizeof(double *) * (size_t)n));
        CU(c, cudaMemcpy(c->mbox_peers, peers.data(), sizeof(double *) * (size_t)n, cudaMemcpyHostToDevice));
        c->mbox_on = true;
    }
    return MACROC_OK;
}

extern "C" int macroc_create(const macroc_config *cfg, int rank, int nranks, const void *id128, macroc_ctx **out)
{
    if (!cfg || !out) FAIL((macroc_ctx *)nullptr, MACROC_ERR_ARG, "macroc_create: null argument");
    *out = nullptr;
    Slab slab;
    int rc = make_slab(*cfg, rank, nranks, &slab);
    if (rc) FAIL((macroc_ctx *)nullptr, rc, "macroc_create: -da_processors_x/y/z do not factor the rank count or exceed the grid");
    if (cfg->bc_type != MACROC_BC_BENDING && cfg->bc_type != MACROC_BC_CIRCLE)
        FAIL((macroc_ctx *)nullptr, MACROC_ERR_ARG, "macroc_create: bc_type must be 0 or 1");
    if (nranks > 1 && !id128) FAIL((macroc_ctx *)nullptr, MACROC_ERR_ARG, "macroc_create: nranks > 1 needs a unique id");
    if (cfg->op != MACROC_OP_ASSEMBLED && cfg->op != MACROC_OP_MATRIX_FREE && cfg->op != MACROC_OP_ASSEMBLED_SYM)
        FAIL((macroc_ctx *)nullptr, MACROC_ERR_ARG, "macroc_create: unknown operator %d", (int)cfg->op);
    if (cfg->material != MACROC_MAT_UNIFORM && cfg->material != MACROC_MAT_PER_GP)
        FAIL((macroc_ctx *)nullptr, MACROC_ERR_ARG, "macroc_create: unknown material source %d", (int)cfg->material);
    if (cfg->jac_mode != MACROC_JAC_AUTO && cfg->jac_mode != MACROC_JAC_ELEMENT)
        FAIL((macroc_ctx *)nullptr, MACROC_ERR_ARG, "macroc_create: unknown jac_mode %d", (int)cfg->jac_mode);
    if (cfg->ksp_maxits < 0 || cfg->newton_max_its < 0)
        FAIL((macroc_ctx *)nullptr, MACROC_ERR_ARG, "macroc_create: negative iteration limit");
    if (cfg->strict_fp && (nranks != 1 || cfg->material != MACROC_MAT_UNIFORM || cfg->op != MACROC_OP_ASSEMBLED ||
                           cfg->jac_mode != MACROC_JAC_AUTO))
        FAIL((macroc_ctx *)nullptr, MACROC_ERR_UNSUPPORTED, "macroc_create: strict_fp is a one-rank verification mode "
             "(uniform tangent, full-storage assembled operator)");
    if (cfg->material == MACROC_MAT_PER_GP && cfg->op == MACROC_OP_MATRIX_FREE)
        FAIL((macroc_ctx *)nullptr, MACROC_ERR_UNSUPPORTED, "macroc_create: per-Gauss-point tangents need an assembled operator");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        FAIL((macroc_ctx *)nullptr, MACROC_ERR_NO_DEVICE, "macroc_create: no CUDA device (this library has no CPU path)");
    }
    macroc_ctx *c = new macroc_ctx();
    c->id = g_next_ctx_id++;
    c->cfg = *cfg; c->slab = slab; c->geo = make_geometry(*cfg);
    if (const char *v = getenv("MACROC_SPMV_VARIANT")) c->spmv_variant = atoi(v);
    if (const char *v = getenv("MACROC_SYM_R")) c->sym_R = atoi(v);
    if (const char *v = getenv("MACROC_SYM_NSEG")) c->sym_nseg = atoi(v);
    if (const char *v = getenv("MACROC_SYM_VARIANT")) c->sym_variant = atoi(v);
    if (const char *v = getenv("MACROC_SYM_HINT")) c->sym_hint = atoi(v);
    if (const char *v = getenv("MACROC_MF_VARIANT")) c->mf_variant = atoi(v);
    if (const char *v = getenv("MACROC_MF_NSEG")) c->mf_nseg = atoi(v);
    if (const char *v = getenv("MACROC_ASM_COLBLOCK")) c->asm_colblock = atoi(v);
    if (const char *v = getenv("MACROC_ASM_CTAS")) c->asm_ctas_per_sm = atoi(v);
    if (cfg->device >= 0) c->device = cfg->device;
    else if (cudaGetDevice(&c->device) != cudaSuccess) c->device = 0;
#define CUC(call) do { cudaError_t _e = (call); if (_e != cudaSuccess) { g_last_error = std::string(#call) + " -> " + cudaGetErrorString(_e); ctx_free(c); return MACROC_ERR_CUDA; } } while (0)
    CUC(cudaSetDevice(c->device));
    CUC(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    CUC(cudaStreamCreateWithFlags(&c->comm_stream, cudaStreamNonBlocking));
    for (cudaEvent_t *e : {&c->ev_ready, &c->ev_halo, &c->ev_chk[0], &c->ev_chk[1]}) CUC(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
    CUC(cudaEventCreate(&c->ev_t0)); CUC(cudaEventCreate(&c->ev_t1));
    for (cudaEvent_t &e : c->ev_user) CUC(cudaEventCreate(&e));

    GridDev &g = c->g;
    g.NX = slab.NX; g.NY = slab.NY; g.NZ = slab.NZ; g.zs = slab.zs; g.nzl = slab.nzl;
    g.npl = slab.npl; g.nloc = slab.nloc;
    g.Xs = slab.Xs; g.Ys = slab.Ys; g.ox0 = slab.xs - slab.Xs; g.oy0 = slab.ys - slab.Ys; g.xm = slab.xm; g.ym = slab.ym;
    g.ghost = nullptr;
    c->sg.rt = (slab.NX + TILE_NODES - 1) / TILE_NODES;
    c->sg.zmin = slab.has_lower() ? -1 : 0;
    g.G = (int)(((slab.npl + slab.NX + 1 + 31) / 32) * 32);
    g.ntiles = (slab.nloc + TILE_NODES - 1) / TILE_NODES;
    g.S = g.ntiles * TILE_NODES + 2 * (int64_t)g.G + 32;

    for (int i = 0; i < V_COUNT; ++i) {
        CUC(cudaMalloc(&c->vec[i], sizeof(double) * 3 * g.S));
        CUC(cudaMemsetAsync(c->vec[i], 0, sizeof(double) * 3 * g.S, c->stream));
    }
    CUC(cudaMalloc(&c->Ke, sizeof(double) * 576));
    CUC(cudaMalloc(&c->T, sizeof(double) * 27 * 243));
    CUC(cudaMalloc(&c->sums, sizeof(double) * 4));
    CUC(cudaMalloc(&c->sc, sizeof(CgScalars)));
    CUC(cudaMallocHost(&c->sums_host, sizeof(double) * 4));
    CUC(cudaMallocHost(&c->sc_host, sizeof(CgScalars) * 2));
    CUC(cudaMalloc(&c->stage, sizeof(double) * 3 * (size_t)g.nloc));
    c->vec_blocks = std::min<int64_t>(cdiv64(g.nloc, 256), 148 * 8);
    c->spmv_blocks = std::min<int64_t>(cdiv64(g.ntiles, 8), 148 * 8);
    c->partial_cap = std::max<int64_t>({(int64_t)cdiv64(g.nloc, 128) + slab.nzl + 8, (int64_t)2 * c->vec_blocks, (int64_t)3 * c->spmv_blocks + 8, (int64_t)4096});
    CUC(cudaMalloc(&c->partial, sizeof(double) * c->partial_cap));
    CUC(cudaMalloc(&c->tickets, 2 * sizeof(unsigned)));
    CUC(cudaMemset(c->tickets, 0, 2 * sizeof(unsigned)));

    // element bookkeeping: DMDA-owned layers plus, for the gather-form assembly, the first layer
    // of the upper neighbour (integrated redundantly / received as Gauss-point halo)
    // (in x and y the kernels integrate every element of the local box; for z-slabs these are
    // exactly the DMDA-owned ones)
    c->er.ezs = slab.ezs; c->er.nex = slab.lnex; c->er.ney = slab.lney;
    c->er.nez_ext = slab.nez + (slab.has_upper() ? 1 : 0);
    c->ne_owned = (int64_t)slab.lnex * slab.lney * slab.nez;
    c->ne_ext = (int64_t)slab.lnex * slab.lney * c->er.nez_ext;
    c->er.ne_ext = std::max<int64_t>(c->ne_ext, 1);
    {
        // residual scratch: 24 doubles per element of a chunk of node planes (<= ~256 MB)
        int64_t per_layer = std::max<int64_t>((int64_t)slab.lnex * slab.lney, 1);
        int64_t max_layers = std::max<int64_t>(2, ((int64_t)256 << 20) / (per_layer * 24 * 8));
        c->chunk_planes = (int)std::min<int64_t>(slab.nzl, max_layers - 1);
        if (c->chunk_planes < 1) c->chunk_planes = 1;
        CUC(cudaMalloc(&c->scratch, sizeof(double) * 24 * per_layer * (c->chunk_planes + 1)));
    }
    // Dirichlet bookkeeping: the reference's lists (bc_init over the ghosted box) -> per-node dof
    // mask over the padded local array (ghost planes included) + (index, coef) pairs of the local
    // nodes in the owned planes; x/y ghost flags for the reductions.
    {
        // The reference's per-rank lists are merged by PETSc (VecAssembly / MatZeroRowsColumns
        // forward off-rank rows); the merged set equals the one-rank list, which is what the mask
        // of the local nodes -- ghosts included -- is cut from.
        std::vector<BcEntry> ent;
        int nbcs = 0;
        Slab whole;
        macroc_config serial = *cfg;
        serial.px = serial.py = serial.pz = 1;
        make_slab(serial, 0, 1, &whole);
        build_bc_entries(*cfg, whole, ent, &nbcs);
        std::vector<uint8_t> mask((size_t)g.S, 0);
        std::vector<int64_t> own_idx; std::vector<double> own_coef;
        for (const BcEntry &e : ent) {
            if (e.i < slab.Xs || e.i >= slab.Xs + slab.Xm || e.j < slab.Ys || e.j >= slab.Ys + slab.Ym) continue;
            const int64_t ln = (e.i - slab.Xs) + (int64_t)slab.NX * (e.j - slab.Ys) + slab.npl * (int64_t)(e.k - slab.zs);
            const int64_t pos = g.G + ln;
            if (pos < 0 || pos >= g.S) continue;
            mask[(size_t)pos] |= (uint8_t)(1u << e.d);
            if (e.k >= slab.zs && e.k < slab.zs + slab.nzl) { own_idx.push_back(e.d * g.S + pos); own_coef.push_back(e.coef); }
        }
        c->nbc = (int)own_idx.size();
        CUC(cudaMalloc(&c->nodemask, (size_t)g.S));
        CUC(cudaMemcpyAsync(c->nodemask, mask.data(), (size_t)g.S, cudaMemcpyHostToDevice, c->stream));
        std::vector<uint8_t> msum((size_t)(g.S / 32 + 1), 0);
        for (int64_t q = 0; q < g.S; ++q) msum[(size_t)(q >> 5)] |= mask[(size_t)q];
        CUC(cudaMalloc(&c->masksum, msum.size()));
        CUC(cudaMemcpy(c->masksum, msum.data(), msum.size(), cudaMemcpyHostToDevice));
        if (c->nbc) {
            CUC(cudaMalloc(&c->bc_idx, sizeof(int64_t) * c->nbc));
            CUC(cudaMalloc(&c->bc_coef, sizeof(double) * c->nbc));
            CUC(cudaMemcpyAsync(c->bc_idx, own_idx.data(), sizeof(int64_t) * c->nbc, cudaMemcpyHostToDevice, c->stream));
            CUC(cudaMemcpyAsync(c->bc_coef, own_coef.data(), sizeof(double) * c->nbc, cudaMemcpyHostToDevice, c->stream));
        }
        if (slab.xy_split()) {
            std::vector<uint8_t> gh((size_t)g.S, 0);
            for (int k = 0; k < slab.nzl; ++k)
                for (int j = 0; j < slab.NY; ++j)
                    for (int i = 0; i < slab.NX; ++i) {
                        bool owned = i >= g.ox0 && i < g.ox0 + g.xm && j >= g.oy0 && j < g.oy0 + g.ym;
                        gh[(size_t)(g.G + i + (int64_t)slab.NX * j + slab.npl * k)] = owned ? 0 : 1;
                    }
            CUC(cudaMalloc(&c->ghostflag, (size_t)g.S));
            CUC(cudaMemcpyAsync(c->ghostflag, gh.data(), (size_t)g.S, cudaMemcpyHostToDevice, c->stream));
            g.ghost = c->ghostflag;
            c->xy_halo_stride = (size_t)3 * slab.nzl * std::max(slab.NX, slab.NY);
            CUC(cudaMalloc(&c->xy_halo, sizeof(double) * 8 * c->xy_halo_stride));
        }
        CUC(cudaStreamSynchronize(c->stream));
    }
    // element constants: dsh table, D, Ke, class stencils (kept in c->consts, bound on demand)
    {
        double D[36];
        if (cfg->use_D) memcpy(D, cfg->D, sizeof(D)); else isotropic_D(cfg->E, cfg->nu, D);
        CUC(cudaMalloc(&c->consts, sizeof(double) * (228 + 27 * 243)));
        CUC(cudaMemcpyAsync(c->consts + 192, D, sizeof(D), cudaMemcpyHostToDevice, c->stream));
        const bool phys = cfg->physical_B != 0;
        LAUNCH(c, k_make_dsh, 1, 64, c->consts, phys ? c->geo.dx : 1., phys ? c->geo.dy : 1., phys ? c->geo.dz : 1.);
        LAUNCH(c, k_element_matrix, 3, 192, c->consts, c->consts + 192, c->geo.wg, c->Ke);
        LAUNCH(c, k_stencil_table, cdiv64(27 * 243, 256), 256, c->Ke, c->T);
        CUC(cudaMemcpyAsync(c->consts + 228, c->T, sizeof(double) * 27 * 243, cudaMemcpyDeviceToDevice, c->stream));
        // the __constant__ copies are bound by the first entry point that needs them (ConstLease)
        c->consts_host.resize(228 + 27 * 243);
        CUC(cudaMemcpyAsync(c->consts_host.data(), c->consts, sizeof(double) * c->consts_host.size(), cudaMemcpyDeviceToHost, c->stream));
        CUC(cudaStreamSynchronize(c->stream));
        CUC(cudaGetLastError());
    }
    if (nranks > 1 && is_loopback_id(id128)) {
        // in-process ranks (loopback.h): join the group, meet the other members once
        LoopGroup *grp = loopback_group_of(id128);
        if (!grp || grp->n != nranks) { g_last_error = "macroc_create: loopback id was made for another rank count"; ctx_free(c); return MACROC_ERR_ARG; }
        LoopGroup::Member &me = grp->m[(size_t)rank];
        CUC(cudaEventCreateWithFlags(&me.ev_ready, cudaEventDisableTiming));
        CUC(cudaEventCreateWithFlags(&me.ev_done, cudaEventDisableTiming));
        me.device = c->device;
        { std::lock_guard<std::mutex> lk(grp->mu); grp->joined++; }
        c->loop = grp;
        if (!grp->barrier()) { g_last_error = "macroc_create: loopback group did not assemble (every rank needs its own host thread)"; ctx_free(c); return MACROC_ERR_NCCL; }
    } else if (nranks > 1) {
        ncclUniqueId id;
        memcpy(&id, id128, 128);
        ncclResult_t e = ncclCommInitRank(&c->comm, nranks, id, rank);
        if (e != ncclSuccess) { g_last_error = std::string("ncclCommInitRank -> ") + ncclGetErrorString(e); ctx_free(c); return MACROC_ERR_NCCL; }
        int mrc = mbox_setup(c);
        if (mrc) { ctx_free(c); return mrc; }
    }
#undef CUC
    *out = c;
    return MACROC_OK;
}

// ---------------------------------------------------------------------------
// halo + reductions
// ---------------------------------------------------------------------------

static inline bool has_comm(const macroc_ctx *c) { return c->comm != nullptr || c->loop != nullptr; }

// One grouped exchange with the neighbours: NCCL send/recv, or the loopback rendezvous.  Matching
// is by order per peer pair, like ncclSend/ncclRecv inside one group.  Every rank of the
// communicator must reach the same sequence of exchanges (the call sites below skip an exchange
// only on conditions that are global: a processor-grid extent of 1).
static int comm_exchange(macroc_ctx *c, const std::vector<LoopXfer> &sends, const std::vector<LoopXfer> &recvs, cudaStream_t st)
{
    if (c->comm) {
        if (sends.empty() && recvs.empty()) return MACROC_OK;
        NC(c, ncclGroupStart());
        for (const LoopXfer &x : sends) NC(c, ncclSend(x.buf, x.cnt, ncclDouble, x.peer, c->comm, st));
        for (const LoopXfer &x : recvs) NC(c, ncclRecv(x.buf, x.cnt, ncclDouble, x.peer, c->comm, st));
        NC(c, ncclGroupEnd());
        return MACROC_OK;
    }
    LoopGroup *grp = c->loop;
    if (!grp) return MACROC_OK;
    const int me = c->slab.rank;
    LoopGroup::Member &mine = grp->m[(size_t)me];
    mine.sends = sends;
    CU(c, cudaEventRecord(mine.ev_ready, st));                    // my send buffers are final after this point of `st`
    if (!grp->barrier()) FAIL(c, MACROC_ERR_NCCL, "loopback exchange: a rank is missing (timeout or failure of another rank)");
    std::vector<size_t> taken((size_t)grp->n, 0);
    for (const LoopXfer &r : recvs) {
        const LoopGroup::Member &src = grp->m[(size_t)r.peer];
        const LoopXfer *match = nullptr;
        size_t seen = 0;
        for (const LoopXfer &sx : src.sends)
            if (sx.peer == me && seen++ == taken[(size_t)r.peer]) { match = &sx; break; }
        if (!match || match->cnt != r.cnt) { grp->fail(); FAIL(c, MACROC_ERR_NCCL, "loopback exchange: rank %d has no matching send for rank %d", r.peer, me); }
        taken[(size_t)r.peer]++;
        CU(c, cudaStreamWaitEvent(st, src.ev_ready, 0));
        CU(c, cudaMemcpyAsync(r.buf, match->buf, sizeof(double) * r.cnt, cudaMemcpyDefault, st));
    }
    CU(c, cudaEventRecord(mine.ev_done, st));                     // everything I had to fetch is enqueued before this
    if (!grp->barrier()) FAIL(c, MACROC_ERR_NCCL, "loopback exchange: a rank is missing (timeout or failure of another rank)");
    for (const LoopXfer &x : sends) CU(c, cudaStreamWaitEvent(st, grp->m[(size_t)x.peer].ev_done, 0));   // do not overwrite before it was read
    return MACROC_OK;
}

// x / y phases of DMGlobalToLocal for a general DMDA box: the owned boundary column (row) of the
// owned planes goes to the neighbour's ghost column (row).  Run x, then y (rows include the x
// ghost columns just received), then z (planes include both): edge and corner ghosts arrive
// without diagonal messages.
static int halo_exchange_xy(macroc_ctx *c, double *v, cudaStream_t st)
{
    const GridDev &g = c->g;
    const Slab &s = c->slab;
    if (!has_comm(c) || !s.xy_split()) return MACROC_OK;
    for (int axis = 0; axis < 2; ++axis) {
        if ((axis == 0 ? s.px : s.py) == 1) continue;             // global: no rank has a neighbour along this axis
        const int lo = s.nb[2 * axis], hi = s.nb[2 * axis + 1];
        const int len = axis == 0 ? g.NY : g.NX, ext = axis == 0 ? g.NX : g.NY;
        const int first_owned = axis == 0 ? g.ox0 : g.oy0, last_owned = first_owned + (axis == 0 ? g.xm : g.ym) - 1;
        const size_t cnt = (size_t)3 * g.nzl * len;
        double *sb_lo = c->xy_halo + (4 * axis + 0) * c->xy_halo_stride, *sb_hi = c->xy_halo + (4 * axis + 1) * c->xy_halo_stride;
        double *rb_lo = c->xy_halo + (4 * axis + 2) * c->xy_halo_stride, *rb_hi = c->xy_halo + (4 * axis + 3) * c->xy_halo_stride;
        const int blocks = cdiv64((int64_t)cnt, 256);
        if (lo >= 0) { k_halo_pack_xy<<<blocks, 256, 0, st>>>(g, v, sb_lo, axis, first_owned, 1); c->launches++; }
        if (hi >= 0) { k_halo_pack_xy<<<blocks, 256, 0, st>>>(g, v, sb_hi, axis, last_owned, 1); c->launches++; }
        std::vector<LoopXfer> sends, recvs;
        if (lo >= 0) { sends.push_back({lo, sb_lo, cnt}); recvs.push_back({lo, rb_lo, cnt}); }
        if (hi >= 0) { sends.push_back({hi, sb_hi, cnt}); recvs.push_back({hi, rb_hi, cnt}); }
        int rc = comm_exchange(c, sends, recvs, st);
        if (rc) return rc;
        if (lo >= 0) { k_halo_pack_xy<<<blocks, 256, 0, st>>>(g, v, rb_lo, axis, 0, 0); c->launches++; }
        if (hi >= 0) { k_halo_pack_xy<<<blocks, 256, 0, st>>>(g, v, rb_hi, axis, ext - 1, 0); c->launches++; }
    }
    return MACROC_OK;
}

// DMGlobalToLocal (assembly.c:40-41 and inside every MatMult), z phase: one node plane per side
// and component over NCCL send/recv.
static int halo_exchange_z(macroc_ctx *c, double *v, cudaStream_t st)
{
    if (!has_comm(c)) return MACROC_OK;
    const GridDev &g = c->g;
    const Slab &s = c->slab;
    if (s.pz == 1) return MACROC_OK;
    std::vector<LoopXfer> sends, recvs;
    for (int d = 0; d < 3; ++d) {
        double *base = v + d * g.S + g.G;
        if (s.has_lower()) {
            sends.push_back({s.nb[4], base, (size_t)g.npl});
            recvs.push_back({s.nb[4], base - g.npl, (size_t)g.npl});
        }
        if (s.has_upper()) {
            sends.push_back({s.nb[5], base + g.nloc - g.npl, (size_t)g.npl});
            recvs.push_back({s.nb[5], base + g.nloc, (size_t)g.npl});
        }
    }
    return comm_exchange(c, sends, recvs, st);
}

static int halo_exchange(macroc_ctx *c, double *v, cudaStream_t st)
{
    int rc = halo_exchange_xy(c, v, st);
    if (rc) return rc;
    return halo_exchange_z(c, v, st);
}

// sum of c->sums[0..n) over the ranks, in place, on the context's stream
static int allreduce_sums(macroc_ctx *c, int n)
{
    if (c->comm) {
        NC(c, ncclAllReduce(c->sums, c->sums, (size_t)n, ncclDouble, ncclSum, c->comm, c->stream));
        return MACROC_OK;
    }
    LoopGroup *grp = c->loop;
    if (!grp) return MACROC_OK;
    const int me = c->slab.rank;
    if (me == 0 && !grp->host_part) CU(c, cudaMallocHost(&grp->host_part, sizeof(double) * 4 * (size_t)grp->n));
    if (!grp->barrier()) FAIL(c, MACROC_ERR_NCCL, "loopback all-reduce: a rank is missing");
    CU(c, cudaMemcpyAsync(grp->host_part + 4 * me, c->sums, sizeof(double) * n, cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    if (!grp->barrier()) FAIL(c, MACROC_ERR_NCCL, "loopback all-reduce: a rank is missing");
    for (int q = 0; q < n; ++q) {
        double acc = 0.;
        for (int r = 0; r < grp->n; ++r) acc += grp->host_part[4 * r + q];     // rank order: bit-reproducible
        c->sums_host[q] = acc;
    }
    if (!grp->barrier()) FAIL(c, MACROC_ERR_NCCL, "loopback all-reduce: a rank is missing");   // everyone has read the partials
    CU(c, cudaMemcpyAsync(c->sums, c->sums_host, sizeof(double) * n, cudaMemcpyHostToDevice, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));          // sums_host is reused by the callers
    return MACROC_OK;
}

// ---------------------------------------------------------------------------
// hot path
// ---------------------------------------------------------------------------

extern "C" double macroc_get_displacement(const macroc_ctx *ctx, int time_s)
{
    // bcs.c:52-58 (intended value; the reference function lacks its return)
    if (!ctx) return NAN;
    double time = time_s * ctx->cfg.dt;
    return -1.0 * (time / ctx->cfg.final_time);
}

extern "C" int macroc_apply_bc_on_u(macroc_ctx *c, double U)
{
    if (!c) return MACROC_ERR_ARG;
    CU(c, cudaSetDevice(c->device));
    if (c->nbc) LAUNCH(c, k_scatter_bc, cdiv64(c->nbc, 256), 256, c->bc_idx, c->bc_coef, c->nbc, U, c->vec[V_U]);
    CU(c, cudaGetLastError());
    return MACROC_OK;
}

static int ensure_gp_arrays(macroc_ctx *c, bool need_ctan)
{
    size_t n = (size_t)std::max<int64_t>(c->ne_ext, 1);
    if (!c->strain) {
        CU(c, cudaMalloc(&c->strain, sizeof(double) * 48 * n));
        CU(c, cudaMalloc(&c->stress, sizeof(double) * 48 * n));
        CU(c, cudaMemsetAsync(c->strain, 0, sizeof(double) * 48 * n, c->stream));
        CU(c, cudaMemsetAsync(c->stress, 0, sizeof(double) * 48 * n, c->stream));
    }
    if (need_ctan && !c->ctan) {
        cudaError_t e = cudaMalloc(&c->ctan, sizeof(double) * 288 * n);
        if (e != cudaSuccess) { cudaGetLastError(); FAIL(c, MACROC_ERR_MEM, "per-Gauss-point tangents need %.2f GB", 288. * 8 * n / 1e9); }
        CU(c, cudaMemsetAsync(c->ctan, 0, sizeof(double) * 288 * n, c->stream));
    }
    return MACROC_OK;
}

// Gauss-point halo.  The local box holds, beside the DMDA-owned elements, the first owned
// element layer of every upper neighbour (x+, y+, z+): the layer the gather-form assembly
// integrates redundantly.  Three phases (x, y, z; a phase sends layer 0 along its axis over the
// full local extent of the other two, ghosts received earlier included), so edge and corner
// elements arrive without diagonal messages.  The arrays are SoA over elements: a layer is
// packed into / unpacked from a contiguous buffer.
static int halo_gp_layer(macroc_ctx *c, double *arr, int nq)
{
    if (!has_comm(c)) return MACROC_OK;
    const Slab &s = c->slab;
    const int64_t lnex = s.lnex, lney = s.lney, nlay = c->er.nez_ext;
    const int64_t maxface = std::max({lney * nlay, lnex * nlay, lnex * lney});
    if (!c->gp_halo) CU(c, cudaMalloc(&c->gp_halo, sizeof(double) * 2 * (size_t)maxface * 288));
    double *sbuf = c->gp_halo, *rbuf = c->gp_halo + (size_t)maxface * 288;
    const int64_t owned[3] = {s.nex, s.ney, s.nez};          // DMDA-owned element layers per axis
    const int pgrid[3] = {s.px, s.py, s.pz};
    for (int axis = 0; axis < 3; ++axis) {
        if (pgrid[axis] == 1) continue;                       // global: nobody has a neighbour along this axis
        const int lo = s.nb[2 * axis], hi = s.nb[2 * axis + 1];
        const int64_t face = axis == 0 ? lney * nlay : (axis == 1 ? lnex * nlay : lnex * lney);
        const size_t cnt = (size_t)face * nq;
        const int blocks = cdiv64((int64_t)cnt, 256);
        if (lo >= 0) LAUNCH(c, k_gp_face_copy, blocks, 256, nq, lnex, lney, nlay, axis, (int64_t)0, c->er.ne_ext, arr, sbuf, 1);
        std::vector<LoopXfer> sends, recvs;
        if (lo >= 0) sends.push_back({lo, sbuf, cnt});
        if (hi >= 0) recvs.push_back({hi, rbuf, cnt});
        int rc = comm_exchange(c, sends, recvs, c->stream);
        if (rc) return rc;
        if (hi >= 0) LAUNCH(c, k_gp_face_copy, blocks, 256, nq, lnex, lney, nlay, axis, owned[axis], c->er.ne_ext, arr, rbuf, 0);
    }
    return MACROC_OK;
}

extern "C" int macroc_set_strains(macroc_ctx *c, int materialize)
{
    if (!c) return MACROC_ERR_ARG;
    CU(c, cudaSetDevice(c->device));
    BIND_CONSTANTS(c);
    int rc = halo_exchange(c, c->vec[V_U], c->stream);
    if (rc) return rc;
    if (materialize || c->cfg.material == MACROC_MAT_PER_GP) {
        const Slab &s = c->slab;
        if ((rc = ensure_gp_arrays(c, false))) return rc;
        // strain of the owned elements; stress = D strain as a by-product for the uniform law
        if (c->ne_owned > 0)
            LAUNCH(c, k_strain_stress, cdiv64(c->ne_owned, 128), 128, c->g, s.ezs, s.nez, c->er.ne_ext, c->vec[V_U], c->strain,
                   c->cfg.material == MACROC_MAT_PER_GP ? nullptr : c->stress);
        CU(c, cudaGetLastError());
    }
    return MACROC_OK;
}

extern "C" int macroc_homogenize(macroc_ctx *c)
{
    if (!c) return MACROC_ERR_ARG;
    if (c->cfg.material != MACROC_MAT_PER_GP) return MACROC_OK;
    CU(c, cudaSetDevice(c->device));
    BIND_CONSTANTS(c);
    int rc = ensure_gp_arrays(c, true);
    if (rc) return rc;
    if (c->ne_owned > 0)
        LAUNCH(c, k_homogenize_linear, cdiv64(c->ne_owned, 128), 128, c->ne_owned, (int64_t)c->slab.lnex, (int64_t)c->slab.lney,
               (int64_t)c->slab.nex, (int64_t)c->slab.ney, c->er.ne_ext, c->strain, c->stress, c->ctan);
    CU(c, cudaGetLastError());
    return MACROC_OK;
}

extern "C" int macroc_gp_arrays(macroc_ctx *c, double **strain, double **stress, double **ctan, int64_t *n_gp,
                                int64_t *pitch)
{
    if (!c) return MACROC_ERR_ARG;
    CU(c, cudaSetDevice(c->device));
    int rc = ensure_gp_arrays(c, ctan != nullptr);
    if (rc) return rc;
    if (strain) *strain = c->strain;
    if (stress) *stress = c->stress;
    if (ctan) *ctan = c->ctan;
    if (n_gp) *n_gp = c->ne_owned * 8;
    if (pitch) *pitch = c->er.ne_ext;
    return MACROC_OK;
}

// host (reference AoS view, gpi = ie*8+gp) -> device SoA arrays
static int gp_upload(macroc_ctx *c, const double *host, double *dev, int n)
{
    const Slab &s = c->slab;
    const int64_t ne = (int64_t)s.nex * s.ney * s.nez;           // DMDA-owned elements
    if (!host || ne == 0) return MACROC_OK;
    double *tmp = nullptr;
    CU(c, cudaMalloc(&tmp, sizeof(double) * 8 * n * (size_t)ne));
    cudaError_t e = cudaMemcpyAsync(tmp, host, sizeof(double) * 8 * n * (size_t)ne, cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess) {
        LAUNCH(c, k_gp_aos_soa, cdiv64(ne * 8 * n, 256), 256, n, (int64_t)s.nex, (int64_t)s.ney, (int64_t)s.nez, s.exs - s.Xs,
               s.eys - s.Ys, (int64_t)s.lnex, (int64_t)s.lney, c->er.ne_ext, tmp, dev, 1);
        e = cudaStreamSynchronize(c->stream);
    }
    cudaFree(tmp);
    if (e != cudaSuccess) FAIL(c, MACROC_ERR_CUDA, "gp_upload: %s", cudaGetErrorString(e));
    return MACROC_OK;
}

static int gp_download(macroc_ctx *c, const double *dev, double *host, int n)
{
    const Slab &s = c->slab;
    const int64_t ne = (int64_t)s.nex * s.ney * s.nez;           // DMDA-owned elements
    if (!host || ne == 0) return MACROC_OK;
    double *tmp = nullptr;
    CU(c, cudaMalloc(&tmp, sizeof(double) * 8 * n * (size_t)ne));
    LAUNCH(c, k_gp_aos_soa, cdiv64(ne * 8 * n, 256), 256, n, (int64_t)s.nex, (int64_t)s.ney, (int64_t)s.nez, s.exs - s.Xs,
           s.eys - s.Ys, (int64_t)s.lnex, (int64_t)s.lney, c->er.ne_ext, dev, tmp, 0);
    cudaError_t e = cudaMemcpyAsync(host, tmp, sizeof(double) * 8 * n * (size_t)ne, cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    cudaFree(tmp);
    if (e != cudaSuccess) FAIL(c, MACROC_ERR_CUDA, "gp_download: %s", cudaGetErrorString(e));
    return MACROC_OK;
}

extern "C" int macroc_set_gp_data(macroc_ctx *c, const double *stress_host, const double *ctan_host)
{
    if (!c) return MACROC_ERR_ARG;
    CU(c, cudaSetDevice(c->device));
    int rc = ensure_gp_arrays(c, ctan_host != nullptr);
    if (rc) return rc;
    if ((rc = gp_upload(c, stress_host, c->stress, 6))) return rc;
    if ((rc = gp_upload(c, ctan_host, c->ctan, 36))) return rc;
    return MACROC_OK;
}

// b = -(sum_e B^T sigma wg) on the owned nodes, two passes per chunk of node planes
static int residual_launch(macroc_ctx *c, int *nparts_out)
{
    const GridDev &g = c->g;
    const Slab &s = c->slab;
    const bool per_gp = c->cfg.material == MACROC_MAT_PER_GP;
    if (per_gp) {
        if (!c->stress) FAIL(c, MACROC_ERR_ARG, "assembly_res: no Gauss-point stresses (call set_strains + homogenize)");
        int rc = halo_gp_layer(c, c->stress, 48);   // 8 gp x 6
        if (rc) return rc;
    }
    int nparts = 0;
    const int last_stored = c->er.ezs + c->er.nez_ext - 1;
    for (int k0 = 0; k0 < s.nzl; k0 += c->chunk_planes) {
        const int nk = std::min(c->chunk_planes, s.nzl - k0);
        const int lay_lo = std::max(s.zs + k0 - 1, c->er.ezs), lay_hi = std::min(s.zs + k0 + nk - 1, last_stored);
        const int l0 = lay_lo - c->er.ezs, nl = lay_hi - lay_lo + 1;
        if (nl > 0) {
            int64_t n = (int64_t)s.lnex * s.lney * nl;
            if (c->cfg.strict_fp) LAUNCH(c, k_elem_forces_strict, cdiv64(n, 128), 128, g, c->er, l0, nl, c->geo.wg, c->vec[V_U], c->scratch);
            else if (per_gp) LAUNCH(c, k_elem_forces<true>, cdiv64(n, 128), 128, g, c->er, l0, nl, c->geo.wg, c->vec[V_U], c->stress, c->scratch);
            else LAUNCH(c, k_elem_forces<false>, cdiv64(n, 128), 128, g, c->er, l0, nl, c->geo.wg, c->vec[V_U], c->stress, c->scratch);
        }
        int blocks = cdiv64(g.npl * nk, 256);
        LAUNCH(c, k_gather_forces, blocks, 256, g, c->er, l0, std::max(nl, 0), k0, nk, c->scratch, c->nodemask, c->vec[V_B], c->partial + nparts);
        nparts += blocks;
    }
    *nparts_out = nparts;
    CU(c, cudaGetLastError());
    return MACROC_OK;
}

extern "C" int macroc_assembly_res(macroc_ctx *c, double *norm)
{
    if (!c) return MACROC_ERR_ARG;
    CU(c, cudaSetDevice(c->device));
    BIND_CONSTANTS(c);
    int nparts = 0;
    int rc = residual_launch(c, &nparts);
    if (rc) return rc;
    if (c->cfg.strict_fp) LAUNCH(c, k_seq_norm2, 1, 32, c->g, c->vec[V_B], c->sums);     // the sequential Vec's summation order
    else LAUNCH(c, k_reduce, 1, 256, c->partial, nparts, c->sums);
    rc = allreduce_sums(c, 1);
    if (rc) return rc;
    CU(c, cudaMemcpyAsync(c->sums_host, c->sums, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    if (norm) *norm = sqrt(c->sums_host[0]);           // VecNorm(b, NORM_2) main.c:67
    return MACROC_OK;
}

static int ensure_operator_storage(macroc_ctx *c)
{
    if (c->A) return MACROC_OK;
    size_t bytes = sizeof(double) * (size_t)TILE_DOUBLES * (size_t)c->g.ntiles;
    cudaError_t e = cudaMalloc(&c->A, bytes);
    if (e != cudaSuccess) {
        cudaGetLastError();
        FAIL(c, MACROC_ERR_MEM, "operator needs %.2f GB of device memory: %s", bytes / 1e9, cudaGetErrorString(e));
    }
    return MACROC_OK;
}

// Launch of the per-element Jacobian kernel into the full (27-slot) or the symmetric (14-slot) layout.
template <bool SYM>
static int launch_assemble_elements(macroc_ctx *c, bool per_gp, double2 *A, int64_t tile_lo, int64_t tile_hi)
{
    const int64_t tpp = SYM ? sym_tiles_per_plane(c->g, c->sg) : std::max<int64_t>(1, (c->g.npl + TILE_NODES - 1) / TILE_NODES);
    const int64_t colblock = c->asm_colblock > 0 ? std::min<int64_t>(c->asm_colblock, tpp) : tpp;
    // node-centric kernels (assembly_node.cuh)
    static bool configured[64] = {false};
    if (!configured[c->device & 63]) {
        CU(c, cudaFuncSetAttribute(k_assemble_nodes_pergp<SYM>, cudaFuncAttributeMaxDynamicSharedMemorySize, ASMN_SMEM_PER_GP));
        configured[c->device & 63] = true;
    }
    if (per_gp) {
        // exactly the resident CTAs, so that neighbouring rows and planes are in flight together and share their tangents in L2
        const int blocks = (int)std::min<int64_t>(tile_hi - tile_lo, 148 * std::max(1, c->asm_ctas_per_sm));
        k_assemble_nodes_pergp<SYM><<<blocks, ASMN_THREADS, ASMN_SMEM_PER_GP, c->stream>>>(c->g, c->sg, c->er, c->geo.wg, c->ctan, c->nodemask, c->masksum, A, c->vec[V_DINV], tile_lo, tile_hi, tpp, colblock);
    } else {
        const int blocks = (int)std::min<int64_t>(cdiv64((tile_hi - tile_lo) * 9, ASMU_WARPS), 148 * ASMU_CTAS_PER_SM);
        k_assemble_nodes_uniform<SYM><<<blocks, ASMU_WARPS * 32, 0, c->stream>>>(c->g, c->sg, c->er, c->geo.wg, c->nodemask, c->masksum, A, c->vec[V_DINV], tile_lo, tile_hi, tpp);
    }
    c->launches++;
    return MACROC_OK;
}

extern "C" int macroc_assembly_jac(macroc_ctx *c)
{
    if (!c) return MACROC_ERR_ARG;
    CU(c, cudaSetDevice(c->device));
    BIND_CONSTANTS(c);
    if (c->cfg.op == MACROC_OP_ASSEMBLED_SYM) {
        const int64_t tpp = sym_tiles_per_plane(c->g, c->sg);
        const int64_t tile_lo = c->sg.zmin * tpp, tile_hi = tpp * c->g.nzl;
        if (!c->Asym_alloc) {
            // with a lower z neighbour the ghost plane's dz = +1 blocks are kept as a private copy (plane -1)
            size_t bytes = (size_t)SYM_TILE_BYTES * (size_t)(tile_hi - tile_lo);
            cudaError_t e = cudaMalloc(&c->Asym_alloc, bytes);
            if (e != cudaSuccess) { cudaGetLastError(); FAIL(c, MACROC_ERR_MEM, "operator needs %.2f GB of device memory", bytes / 1e9); }
            c->Asym = c->Asym_alloc - tile_lo * (int64_t)(SYM_PAIRS * TILE_NODES);
        }
        const bool per_gp = c->cfg.material == MACROC_MAT_PER_GP;
        if (per_gp || c->cfg.jac_mode == MACROC_JAC_ELEMENT) {
            // per-Gauss-point tangents (assumed symmetric, like Ke itself): the element kernel writes slots 13..26
            if (per_gp) {
                if (!c->ctan) FAIL(c, MACROC_ERR_ARG, "assembly_jac: no Gauss-point tangents (call homogenize)");
                int rc = halo_gp_layer(c, c->ctan, 288);   // 8 gp x 36
                if (rc) return rc;
            }
            int rc = launch_assemble_elements<true>(c, per_gp, c->Asym, tile_lo, tile_hi);
            if (rc) return rc;
        } else
            LAUNCH(c, k_fill_operator_sym, cdiv64(tile_hi - tile_lo, 8), 256, c->g, c->sg, c->T, c->nodemask, c->Asym, c->vec[V_DINV],
                   tile_lo, tile_hi);
        c->Asym_valid = true;
    } else if (c->cfg.op == MACROC_OP_MATRIX_FREE) {
        if (c->cfg.material != MACROC_MAT_UNIFORM) FAIL(c, MACROC_ERR_UNSUPPORTED, "matrix-free operator needs the uniform tangent");
        LAUNCH(c, k_mf_diag, cdiv64(c->g.nloc, 256), 256, c->g, c->T, c->nodemask, c->vec[V_DINV]);
        c->mf_ready = true;
    } else {
        int rc = ensure_operator_storage(c);
        if (rc) return rc;
        const bool per_gp = c->cfg.material == MACROC_MAT_PER_GP;
        if (per_gp || c->cfg.jac_mode == MACROC_JAC_ELEMENT) {
            if (per_gp) {
                if (!c->ctan) FAIL(c, MACROC_ERR_ARG, "assembly_jac: no Gauss-point tangents (call homogenize)");
                if ((rc = halo_gp_layer(c, c->ctan, 288))) return rc;   // 8 gp x 36
            }
            if ((rc = launch_assemble_elements<false>(c, per_gp, c->A, 0, c->g.ntiles))) return rc;
        } else
            LAUNCH(c, k_fill_operator, cdiv64(c->g.ntiles, 8), 256, c->g, c->T, c->nodemask, c->A, c->vec[V_DINV]);
        c->A_valid = true;
    }
    CU(c, cudaGetLastError());
    return MACROC_OK;
}

template <int WARPS, int NSTAGE>
static int spmv_launch_tma(macroc_ctx *c, double *p, double *w, int64_t first, int64_t count, double *partial,
                           bool with_dot, const int *done)
{
    using SM = SpmvTmaSmem<WARPS, NSTAGE>;
    static bool configured[64] = {false};          // function attributes are per device
    if (!configured[c->device & 63]) {
        cudaError_t e1 = cudaFuncSetAttribute(k_spmv_tma<WARPS, NSTAGE, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM::total);
        cudaError_t e2 = cudaFuncSetAttribute(k_spmv_tma<WARPS, NSTAGE, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM::total);
        if (e1 != cudaSuccess || e2 != cudaSuccess) return -1;
        configured[c->device & 63] = true;
    }
    int per_sm = std::max(1, (227 * 1024) / (SM::total + 1024));
    int blocks = (int)std::min<int64_t>(cdiv64(count, WARPS), (int64_t)148 * per_sm);
    if (with_dot) k_spmv_tma<WARPS, NSTAGE, true><<<blocks, WARPS * 32, SM::total, c->stream>>>(c->g, c->A, p, w, first, count, partial, done, c->fuse_pw);
    else k_spmv_tma<WARPS, NSTAGE, false><<<blocks, WARPS * 32, SM::total, c->stream>>>(c->g, c->A, p, w, first, count, partial, done, CgFuse{nullptr, nullptr});
    if (with_dot && c->fuse_pw.ticket) c->fuse_pw_done = true;
    c->launches++;
    return blocks;
}

// one assembled SpMV launch over tiles [first, first+count); returns the number of partials written
static int spmv_launch(macroc_ctx *c, double *p, double *w, int64_t first, int64_t count, double *partial,
                       bool with_dot, const int *done)
{
    const GridDev &g = c->g;
    switch (c->spmv_variant) {
        case 10: return spmv_launch_tma<8, 4>(c, p, w, first, count, partial, with_dot, done);
        case 11: return spmv_launch_tma<8, 5>(c, p, w, first, count, partial, with_dot, done);
        case 14: return spmv_launch_tma<6, 6>(c, p, w, first, count, partial, with_dot, done);
        default: {
            int blocks = (int)std::min<int64_t>(cdiv64(count, 8), 148 * 8);
            if (with_dot) LAUNCH(c, (k_spmv<true, 1>), blocks, 256, g, c->A, p, w, first, count, partial, done);
            else LAUNCH(c, (k_spmv<false, 1>), blocks, 256, g, c->A, p, w, first, count, partial, done);
            return blocks;
        }
    }
}

// symmetric-storage SpMV over the owned planes [first, first + count): bands of R rows x z
// segments, about one item per resident warp.  Returns the number of partials written (< 0: error).
template <int WARPS, int NSTAGE, int RMAX>
static int spmv_sym_launch(macroc_ctx *c, double *p, double *w, int first, int count, double *partial, bool with_dot,
                           const int *done)
{
    using SM = SpmvSymSmem<WARPS, NSTAGE, RMAX>;
    const GridDev &g = c->g;
    static bool configured[64] = {false};          // function attributes are per device
    if (!configured[c->device & 63]) {
        cudaError_t e1 = cudaFuncSetAttribute(k_spmv_sym<WARPS, NSTAGE, RMAX, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM::total);
        cudaError_t e2 = cudaFuncSetAttribute(k_spmv_sym<WARPS, NSTAGE, RMAX, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM::total);
        if (e1 != cudaSuccess || e2 != cudaSuccess) return -1;
        configured[c->device & 63] = true;
    }
    // Work items = bands (x tile x R rows) x z segments; one item per resident warp is the goal.
    // Cost model (tile loads, relative): a warp's items run back to back, an item streams R rows of
    // lseg planes plus the scatter-only pass below it (5 of 7 chunks), and re-reads (9/13 of the)
    // neighbour blocks of the rows above / below the band: pick the (R, nseg) with the cheapest
    // slowest warp; 256^3 lands on R = 7, nseg = 4 -> exactly 148 x 8 items.
    const int64_t target = (int64_t)148 * WARPS;
    int R = 1, nseg = 1;
    {
        double best = 1e300;
        const int r_lo = c->sym_R > 0 ? std::min(c->sym_R, RMAX) : 1, r_hi = c->sym_R > 0 ? r_lo : std::min(RMAX, g.NY);
        for (int r = r_hi; r >= r_lo; --r) {
            const int64_t bnds = (int64_t)c->sg.rt * ((g.NY + r - 1) / r);
            const int s_hi = c->sym_nseg > 0 ? c->sym_nseg : (int)std::min<int64_t>(count, 4 * target / bnds + 1);
            for (int sgm = c->sym_nseg > 0 ? c->sym_nseg : 1; sgm <= s_hi; ++sgm) {
                const int ls = (count + sgm - 1) / sgm, eff = (count + ls - 1) / ls;
                const int64_t waves = (bnds * eff + target - 1) / target;
                const double cost = (double)waves * r * (ls + 0.7) * (1. + 0.64 / r);
                if (cost < best * (1. - 1e-9)) { best = cost; R = r; nseg = eff; }
            }
        }
    }
    const int64_t bands = (int64_t)c->sg.rt * ((g.NY + R - 1) / R);
    const int blocks = (int)std::min<int64_t>(cdiv64(bands * nseg, WARPS), 148);
    if (with_dot) k_spmv_sym<WARPS, NSTAGE, RMAX, true><<<blocks, WARPS * 32, SM::total, c->stream>>>(g, c->sg, c->Asym, p, w, first, first + count, R, nseg, partial, done, c->sym_hint, c->fuse_pw);
    else k_spmv_sym<WARPS, NSTAGE, RMAX, false><<<blocks, WARPS * 32, SM::total, c->stream>>>(g, c->sg, c->Asym, p, w, first, first + count, R, nseg, partial, done, c->sym_hint, CgFuse{nullptr, nullptr});
    if (with_dot && c->fuse_pw.ticket) c->fuse_pw_done = true;
    c->launches++;
    return blocks;
}

// w = A p on the context's stream; p's halo is exchanged on comm_stream while
// the rows that do not touch a ghost plane are computed.
// z-marching matrix-free apply: the segment count whose slowest CTA marches through the fewest planes (a segment costs
// its planes + 2), and whether the form pays at all: on small grids the segments get so short that the two extra planes
// dominate, and the patch form (k_apply_mf3d) is used (mf_variant: 0 auto, 1 patch form, 2 marching form).
static int mf_march_segments(const macroc_ctx *c, int nz)
{
    const int bx = (c->g.NX + MZ_BX - 1) / MZ_BX, by = (c->g.NY + MZ_BY - 1) / MZ_BY, slots = 148 * MZ_CTAS;
    int nseg = 1;
    int64_t best = INT64_MAX;
    for (int q = 1; q <= std::min(nz, 64); ++q) {
        const int64_t items = (int64_t)bx * by * q, rounds = (items + slots - 1) / slots;
        const int64_t cost = rounds * ((nz + q - 1) / q + 2);
        if (cost < best) { best = cost; nseg = q; }
    }
    return nseg;
}
static bool mf_use_march(const macroc_ctx *c, int k0, int k1)
{
    if (c->mf_variant == 1) return false;
    if (c->mf_variant == 2) return true;
    const int nz = k1 - k0, nseg = mf_march_segments(c, nz);
    return (nz + nseg - 1) / nseg >= 6;
}

static int apply_operator(macroc_ctx *c, int op, double *p, double *w, bool with_dot, const int *done,
                          bool fuse_pw_scalars = false, int *nparts_out = nullptr)
{
    const GridDev &g = c->g;
    const bool comm = has_comm(c);
    const bool mf = op == MACROC_OP_MATRIX_FREE;
    const bool symop = op == MACROC_OP_ASSEMBLED_SYM;
    // interior range in nodes (matrix-free), planes (symmetric storage) or tiles (full storage)
    const int64_t total = mf ? g.nloc : (symop ? (int64_t)g.nzl : g.ntiles);
    int64_t lo_end = 0, hi_begin = total;
    if (comm && g.nzl >= 3) {
        if (mf) { lo_end = g.npl; hi_begin = g.nloc - g.npl; }
        else if (symop) { lo_end = 1; hi_begin = g.nzl - 1; }
        else { lo_end = (g.npl + TILE_NODES - 1) / TILE_NODES; hi_begin = (g.nloc - g.npl) / TILE_NODES; }
        if (hi_begin <= lo_end) { lo_end = 0; hi_begin = total; }
    }
    const bool split = comm && lo_end > 0;
    if (comm) {
        // x / y ghost columns are needed by every tile: exchange them first, in stream order
        int rc = halo_exchange_xy(c, p, c->stream);
        if (rc) return rc;
        if (split) {
            CU(c, cudaEventRecord(c->ev_ready, c->stream));
            CU(c, cudaStreamWaitEvent(c->comm_stream, c->ev_ready, 0));
            rc = halo_exchange_z(c, p, c->comm_stream);
            if (rc) return rc;
            CU(c, cudaEventRecord(c->ev_halo, c->comm_stream));
        } else {
            rc = halo_exchange_z(c, p, c->stream);
            if (rc) return rc;
        }
    }
    int nparts = 0, rc_run = MACROC_OK;
    // single rank, one launch: the kernel's last block folds the p.w partials and updates the CG scalars itself
    c->fuse_pw_done = false;
    c->fuse_pw = (with_dot && fuse_pw_scalars && !nparts_out && !split) ? CgFuse{c->sc, c->tickets} : CgFuse{nullptr, nullptr};
    auto run = [&](int64_t first, int64_t count) {
        if (count <= 0) return;
        int blocks;
        if (symop) {
            // planes [first, first + count)
            switch (c->sym_variant) {
                case 1: blocks = spmv_sym_launch<8, 3, 6>(c, p, w, (int)first, (int)count, c->partial + nparts, with_dot, done); break;
                case 2: blocks = spmv_sym_launch<8, 4, 4>(c, p, w, (int)first, (int)count, c->partial + nparts, with_dot, done); break;
                case 3: blocks = spmv_sym_launch<8, 2, 16>(c, p, w, (int)first, (int)count, c->partial + nparts, with_dot, done); break;
                case 4: blocks = spmv_sym_launch<6, 4, 10>(c, p, w, (int)first, (int)count, c->partial + nparts, with_dot, done); break;
                case 5: blocks = spmv_sym_launch<8, 3, 10>(c, p, w, (int)first, (int)count, c->partial + nparts, with_dot, done); break;
                case 6: blocks = spmv_sym_launch<4, 6, 32>(c, p, w, (int)first, (int)count, c->partial + nparts, with_dot, done); break;
                case 8: blocks = spmv_sym_launch<8, 3, 16>(c, p, w, (int)first, (int)count, c->partial + nparts, with_dot, done); break;
                // default: 8 warps x 3 ring stages, bands of up to 8 rows: 168 KB of shared memory leave 32 KB of L1 for p
                default: blocks = spmv_sym_launch<8, 3, 8>(c, p, w, (int)first, (int)count, c->partial + nparts, with_dot, done); break;
            }
            if (blocks < 0) { rc_run = MACROC_ERR_CUDA; return; }
        } else if (mf && mf_use_march(c, (int)(first / g.npl), (int)((first + count) / g.npl))) {
            // z-marching form: work items = column blocks x z segments, dealt round-robin to the resident CTAs
            const int k0 = (int)(first / g.npl), k1 = (int)((first + count) / g.npl);
            const int bx = (g.NX + MZ_BX - 1) / MZ_BX, by = (g.NY + MZ_BY - 1) / MZ_BY, nz = k1 - k0;
            const int slots = 148 * MZ_CTAS;
            int nseg = mf_march_segments(c, nz);
            if (c->mf_nseg > 0) nseg = std::min(c->mf_nseg, nz);
            blocks = (int)std::min<int64_t>((int64_t)bx * by * nseg, slots);
            static bool mz_configured[64] = {false};
            if (!mz_configured[c->device & 63]) {
                cudaError_t e1 = cudaFuncSetAttribute(k_apply_mf_march<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, MZ_SMEM);
                cudaError_t e2 = cudaFuncSetAttribute(k_apply_mf_march<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, MZ_SMEM);
                if (e1 != cudaSuccess || e2 != cudaSuccess) { rc_run = MACROC_ERR_CUDA; return; }
                mz_configured[c->device & 63] = true;
            }
            // the nodes on a face of the box first (their partials come first), then the marching kernel for the rest
            const int64_t nface = mf_face_count(g, k0, k1);
            const int fblocks = (int)std::min<int64_t>(cdiv64(nface, 128), 148 * 16);   // one node per thread up to 303 k face nodes: the kernel lives on loads in flight
            double *pbase = c->partial + nparts;
            if (fblocks > 0) {
                if (with_dot) k_apply_mf_faces<true><<<fblocks, 128, 0, c->stream>>>(g, c->T, c->nodemask, p, w, k0, k1, pbase, done);
                else k_apply_mf_faces<false><<<fblocks, 128, 0, c->stream>>>(g, c->T, c->nodemask, p, w, k0, k1, pbase, done);
                c->launches++;
            }
            if (with_dot) k_apply_mf_march<true><<<blocks, MZ_THREADS, MZ_SMEM, c->stream>>>(g, c->T, c->nodemask, p, w, k0, k1, bx, by, nseg, pbase, fblocks, done, c->fuse_pw);
            else k_apply_mf_march<false><<<blocks, MZ_THREADS, MZ_SMEM, c->stream>>>(g, c->T, c->nodemask, p, w, k0, k1, bx, by, nseg, pbase, fblocks, done, CgFuse{nullptr, nullptr});
            if (with_dot && c->fuse_pw.ticket) c->fuse_pw_done = true;
            c->launches++;
            blocks += fblocks;
        } else if (mf) {
            // node ranges are whole planes here
            const int k0 = (int)(first / g.npl), k1 = (int)((first + count) / g.npl);
            const int tiles_x = (g.NX + MF_TX - 1) / MF_TX, tiles_y = (g.NY + MF_TY - 1) / MF_TY;
            const int64_t ntile = (int64_t)tiles_x * tiles_y * ((k1 - k0 + MF_TZ - 1) / MF_TZ);
            static bool configured[64] = {false};
            if (!configured[c->device & 63]) {
                cudaError_t e1 = cudaFuncSetAttribute(k_apply_mf3d<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, MF_SMEM);
                cudaError_t e2 = cudaFuncSetAttribute(k_apply_mf3d<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, MF_SMEM);
                if (e1 != cudaSuccess || e2 != cudaSuccess) { rc_run = MACROC_ERR_CUDA; return; }
                configured[c->device & 63] = true;
            }
            // an odd grid: the tile -> CTA map must not be periodic in the 8 x-tiles of a row, or the
            // CTAs that always get a boundary column finish last
            blocks = (int)std::min<int64_t>(ntile, 148 * 6 - 1);
            // the nodes on a face of the box first (their partials come first), then the patch kernel for the rest
            const int64_t nface = mf_face_count(g, k0, k1);
            const int fblocks = (int)std::min<int64_t>(cdiv64(nface, 128), 148 * 16);
            double *pbase = c->partial + nparts;
            if (fblocks > 0) {
                if (with_dot) k_apply_mf_faces<true><<<fblocks, 128, 0, c->stream>>>(g, c->T, c->nodemask, p, w, k0, k1, pbase, done);
                else k_apply_mf_faces<false><<<fblocks, 128, 0, c->stream>>>(g, c->T, c->nodemask, p, w, k0, k1, pbase, done);
                c->launches++;
            }
            if (with_dot) k_apply_mf3d<true><<<blocks, MF_THREADS, MF_SMEM, c->stream>>>(g, c->T, c->nodemask, p, w, k0, k1, tiles_x, tiles_y, pbase, fblocks, done, c->fuse_pw);
            else k_apply_mf3d<false><<<blocks, MF_THREADS, MF_SMEM, c->stream>>>(g, c->T, c->nodemask, p, w, k0, k1, tiles_x, tiles_y, pbase, fblocks, done, CgFuse{nullptr, nullptr});
            if (with_dot && c->fuse_pw.ticket) c->fuse_pw_done = true;
            c->launches++;
            blocks += fblocks;
        } else {
            blocks = spmv_launch(c, p, w, first, count, c->partial + nparts, with_dot, done);
        }
        nparts += blocks;
    };
    if (split) {
        run(lo_end, hi_begin - lo_end);
        CU(c, cudaStreamWaitEvent(c->stream, c->ev_halo, 0));
        run(0, lo_end);
        run(hi_begin, total - hi_begin);
    } else
        run(0, total);
    c->fuse_pw = CgFuse{nullptr, nullptr};
    if (rc_run) FAIL(c, rc_run, "apply_operator: kernel configuration failed (shared memory)");
    if (with_dot) {
        if (nparts_out) *nparts_out = nparts;        // the caller folds the partials (mailbox all-reduce kernel)
        else if (fuse_pw_scalars) { if (!c->fuse_pw_done) LAUNCH(c, k_cg_reduce_pw, 1, 256, c->partial, nparts, c->sc); }
        else LAUNCH(c, k_reduce, 1, 256, c->partial, nparts, c->sums);
    }
    CU(c, cudaGetLastError());
    return MACROC_OK;
}

// one PCG iteration with the reference's rounding (strict_fp.cuh)
static int cg_iteration_strict(macroc_ctx *c)
{
    const GridDev &g = c->g;
    const int nbk = cdiv64(g.nloc, 256);
    LAUNCH(c, k_cg_update_p_strict, nbk, 256, g, c->sc, c->vec[V_R], c->vec[V_DINV], c->vec[V_P]);
    LAUNCH(c, k_spmv_strict, nbk, 256, g, c->A, c->vec[V_P], c->vec[V_W]);
    LAUNCH(c, k_seq_dots, 1, 32, g, 0, c->vec[V_P], c->vec[V_W], c->sums);
    LAUNCH(c, k_cg_scalars_pw, 1, 1, c->sc, c->sums);
    LAUNCH(c, k_cg_update_xr_strict, nbk, 256, g, c->sc, c->vec[V_P], c->vec[V_W], c->vec[V_DU], c->vec[V_R]);
    LAUNCH(c, k_seq_dots, 1, 32, g, 1, c->vec[V_R], c->vec[V_DINV], c->sums);
    LAUNCH(c, k_cg_scalars_iter, 1, 1, c->sc, c->sums);
    CU(c, cudaGetLastError());
    return MACROC_OK;
}

static int cg_iteration(macroc_ctx *c, int op)
{
    if (c->cfg.strict_fp) return cg_iteration_strict(c);
    const GridDev &g = c->g;
    const int nb = c->vec_blocks;
    LAUNCH(c, k_cg_update_p, nb, 256, g, c->sc, c->vec[V_R], c->vec[V_DINV], c->vec[V_P]);
    const bool sample = c->prof_on && c->prof_used < macroc_ctx::PROF_RING && (c->prof_counter++ % c->prof_stride) == 0;
    if (sample) CU(c, cudaEventRecord(c->prof_ev[2 * c->prof_used], c->stream));
    const bool single = !has_comm(c);            // no all-reduce between reduction and scalar update
    const MboxDev mb = {c->mbox_peers, c->slab.rank, c->slab.nranks};
    int nparts = 0;
    int rc = apply_operator(c, op, c->vec[V_P], c->vec[V_W], true, &c->sc->done, single, c->mbox_on ? &nparts : nullptr);
    if (rc) return rc;
    if (sample) { CU(c, cudaEventRecord(c->prof_ev[2 * c->prof_used + 1], c->stream)); c->prof_used++; }
    if (c->mbox_on) {
        // partials -> mailboxes of all ranks -> rank-ordered sum -> CG scalars, one launch
        LAUNCH(c, k_cg_reduce_pw_mbox, 1, 256, c->partial, nparts, c->sc, mb, ++c->mbox_seq);
    } else if (!single) {
        rc = allreduce_sums(c, 1);
        if (rc) return rc;
        LAUNCH(c, k_cg_scalars_pw, 1, 1, c->sc, c->sums);
    }
    LAUNCH(c, k_cg_update_xr, nb, 256, g, c->sc, c->vec[V_P], c->vec[V_W], c->vec[V_DINV], c->vec[V_DU], c->vec[V_R], c->partial, nb,
           single ? CgFuse{c->sc, c->tickets + 1} : CgFuse{nullptr, nullptr});
    if (single) {
        // (z.z, z.r) and the convergence test: folded by the last block of k_cg_update_xr
    } else if (c->mbox_on) {
        LAUNCH(c, k_cg_reduce_iter_mbox, 1, 256, c->partial, nb, c->sc, mb, ++c->mbox_seq);
    } else {
        LAUNCH(c, k_reduce2, 1, 256, c->partial, nb, c->sums);
        rc = allreduce_sums(c, 2);
        if (rc) return rc;
        LAUNCH(c, k_cg_scalars_iter, 1, 1, c->sc, c->sums);
    }
    return MACROC_OK;
}

static int cg_begin(macroc_ctx *c, double rtol, double abstol, double dtol, int maxits)
{
    const GridDev &g = c->g;
    const int nb = c->vec_blocks;
    CgScalars h;
    memset(&h, 0, sizeof(h));
    h.rtol = rtol; h.abstol = abstol; h.dtol = dtol; h.maxits = maxits;
    c->sc_host[0] = h;
    CU(c, cudaMemcpyAsync(c->sc, &c->sc_host[0], sizeof(CgScalars), cudaMemcpyHostToDevice, c->stream));
    if (c->cfg.strict_fp) {
        LAUNCH(c, k_cg_init_strict, cdiv64(g.nloc, 256), 256, g, c->vec[V_B], c->vec[V_DU], c->vec[V_R]);
        LAUNCH(c, k_seq_dots, 1, 32, g, 1, c->vec[V_R], c->vec[V_DINV], c->sums);
        LAUNCH(c, k_cg_scalars_init, 1, 1, c->sc, c->sums);
        return MACROC_OK;
    }
    LAUNCH(c, k_cg_init, nb, 256, g, c->vec[V_B], c->vec[V_DINV], c->vec[V_DU], c->vec[V_R], c->partial, nb);
    if (c->mbox_on) {
        const MboxDev mb = {c->mbox_peers, c->slab.rank, c->slab.nranks};
        LAUNCH(c, k_cg_reduce_init_mbox, 1, 256, c->partial, nb, c->sc, mb, ++c->mbox_seq);
        return MACROC_OK;
    }
    LAUNCH(c, k_reduce2, 1, 256, c->partial, nb, c->sums);
    int rc = allreduce_sums(c, 2);
    if (rc) return rc;
    LAUNCH(c, k_cg_scalars_init, 1, 1, c->sc, c->sums);
    return MACROC_OK;
}

extern "C" int macroc_solve_Ax(macroc_ctx *c, int *its, double *rnorm)
{
    if (!c) return MACROC_ERR_ARG;
    CU(c, cudaSetDevice(c->device));
    BIND_CONSTANTS(c);
    const int op = c->cfg.op;
    if (op == MACROC_OP_ASSEMBLED && !c->A_valid) FAIL(c, MACROC_ERR_ARG, "solve_Ax: assembly_jac has not been called");
    if (op == MACROC_OP_ASSEMBLED_SYM && !c->Asym_valid) FAIL(c, MACROC_ERR_ARG, "solve_Ax: assembly_jac has not been called");
    if (op == MACROC_OP_MATRIX_FREE && !c->mf_ready) FAIL(c, MACROC_ERR_ARG, "solve_Ax: assembly_jac has not been called");
    if (c->prof_on) CU(c, cudaEventRecord(c->ev_t0, c->stream));
    int rc = cg_begin(c, c->cfg.ksp_rtol, c->cfg.ksp_abstol, c->cfg.ksp_dtol, c->cfg.ksp_maxits);
    if (rc) return rc;
    // The iteration count lives on the device; the host only polls a "done"
    // flag every `check` iterations (one check behind), kernels launched after
    // convergence are no-ops, so the count is exactly KSPCG's.
    const int check = 8;
    int pending = -1, slot = 0;
    bool finished = false;
    // Launch-bound small grids (BASELINE configs[1]): after the first batch, `check` iterations are
    // replayed as one CUDA graph (all kernel arguments are iteration-invariant: the CG state lives
    // on the device).  Single rank only; not while the live profile brackets launches with events.
    const bool use_graph = !has_comm(c) && !c->cfg.strict_fp && !c->prof_on && c->g.nloc <= ((int64_t)1 << 21);
    for (int it = 0; it < c->cfg.ksp_maxits + 1 && !finished; ++it) {
        if (use_graph && it >= check && it % check == 0) {
            if (!c->cg_graph[op]) {
                cudaGraph_t graph = nullptr;
                uint64_t before = c->launches;
                CU(c, cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
                for (int q = 0; q < check && !rc; ++q) rc = cg_iteration(c, op);
                cudaError_t e = cudaStreamEndCapture(c->stream, &graph);
                if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
                CU(c, e);
                c->cg_graph_launches[op] = c->launches - before;
                c->launches = before;
                CU(c, cudaGraphInstantiate(&c->cg_graph[op], graph, 0));
                cudaGraphDestroy(graph);
            }
            CU(c, cudaGraphLaunch(c->cg_graph[op], c->stream));
            c->launches += c->cg_graph_launches[op];
            it += check - 1;
        } else {
            rc = cg_iteration(c, op);
            if (rc) return rc;
        }
        if ((it + 1) % check == 0) {
            CU(c, cudaMemcpyAsync(&c->sc_host[slot], c->sc, sizeof(CgScalars), cudaMemcpyDeviceToHost, c->stream));
            CU(c, cudaEventRecord(c->ev_chk[slot], c->stream));
            if (pending >= 0) {
                CU(c, cudaEventSynchronize(c->ev_chk[pending]));
                if (c->sc_host[pending].done) finished = true;
            }
            pending = slot;
            slot ^= 1;
        }
    }
    CU(c, cudaMemcpyAsync(&c->sc_host[0], c->sc, sizeof(CgScalars), cudaMemcpyDeviceToHost, c->stream));
    if (c->prof_on) CU(c, cudaEventRecord(c->ev_t1, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    if (c->prof_on) {
        float sms = 0.f;
        CU(c, cudaEventElapsedTime(&sms, c->ev_t0, c->ev_t1));
        c->prof_solve_ms += sms; c->prof_solve_its += c->sc_host[0].its;
        // samples taken after convergence bracket no-op launches: keep the first `its` only
        int64_t stride = c->prof_stride;
        for (int q = 0; q < c->prof_used; ++q) {
            if ((int64_t)q * stride >= c->sc_host[0].its) break;
            float ms = 0.f;
            CU(c, cudaEventElapsedTime(&ms, c->prof_ev[2 * q], c->prof_ev[2 * q + 1]));
            c->prof_ms += ms; c->prof_samples++;
        }
        c->prof_used = 0; c->prof_counter = 0;
    }
    if (its) *its = c->sc_host[0].its;
    if (rnorm) *rnorm = c->sc_host[0].dp;       // KSPGetResidualNorm: last preconditioned norm
    c->ksp_reason = c->sc_host[0].reason;
    if (c->ksp_reason == KSP_DIVERGED_COMM)
        FAIL(c, MACROC_ERR_NCCL, "solve_Ax: a rank did not deliver its part of a CG dot product (mailbox all-reduce timed out)");
    return MACROC_OK;
}

extern "C" int macroc_set_operator(macroc_ctx *c, int op)
{
    if (!c || (op != MACROC_OP_ASSEMBLED && op != MACROC_OP_MATRIX_FREE && op != MACROC_OP_ASSEMBLED_SYM)) return MACROC_ERR_ARG;
    if (op == MACROC_OP_MATRIX_FREE && c->cfg.material != MACROC_MAT_UNIFORM)
        FAIL(c, MACROC_ERR_UNSUPPORTED, "matrix-free operator needs the uniform tangent");
    if (c->cfg.strict_fp && op != MACROC_OP_ASSEMBLED)
        FAIL(c, MACROC_ERR_UNSUPPORTED, "strict_fp needs the full-storage assembled operator");
    c->cfg.op = op;
    return MACROC_OK;
}

extern "C" int macroc_ksp_reason(const macroc_ctx *c, int *reason)
{
    if (!c || !reason) return MACROC_ERR_ARG;
    *reason = c->ksp_reason;
    return MACROC_OK;
}

extern "C" int macroc_update_u(macroc_ctx *c)
{
    if (!c) return MACROC_ERR_ARG;
    CU(c, cudaSetDevice(c->device));
    dim3 grid(cdiv64(c->g.nloc, 256), 3);
    LAUNCH(c, k_axpy1, grid, 256, c->g, c->vec[V_U], c->vec[V_DU]);
    CU(c, cudaGetLastError());
    return MACROC_OK;
}

extern "C" int macroc_calc_force(macroc_ctx *c, double *force)
{
    if (!c || !force) return MACROC_ERR_ARG;
    CU(c, cudaSetDevice(c->device));
    BIND_CONSTANTS(c);
    const Slab &s = c->slab;
    { int _rc = halo_exchange(c, c->vec[V_U], c->stream); if (_rc) return _rc; }   // ghost nodes of u may be stale after update_u
    // forces.c:75 (bending: ranks that own the X = LX face) / :133 (circle: ghost start + owned
    // count, the reference's mix) decide which ranks contribute
    const bool on_face = c->cfg.bc_type == MACROC_BC_BENDING ? (s.xs + s.xm == s.gNX) : (s.Ys + s.ym == s.gNY);
    int64_t count = !on_face ? 0 : (c->cfg.bc_type == MACROC_BC_BENDING ? (int64_t)s.ney * s.nez : (int64_t)s.nex * s.nez);
    double local = 0.;
    if (count > 0) {
        int nblk = cdiv64(count, 128);
        LAUNCH(c, k_force, nblk, 128, c->g, s.ezs, s.nez, s.exs - s.Xs, s.eys - s.Ys, s.nex, s.ney, c->cfg.bc_type, c->geo.dx, c->geo.dy, c->geo.dz, c->cfg.lx,
               c->cfg.lz, c->geo.rad, c->vec[V_U],
               (c->cfg.material == MACROC_MAT_PER_GP && c->stress) ? c->stress : (const double *)nullptr, c->er.ne_ext,
               c->partial);
        LAUNCH(c, k_reduce, 1, 256, c->partial, nblk, c->sums);
    } else
        CU(c, cudaMemsetAsync(c->sums, 0, sizeof(double), c->stream));
    int rc = allreduce_sums(c, 1);               // MPI_Reduce(SUM) forces.c:47 (every rank gets it)
    if (rc) return rc;
    CU(c, cudaMemcpyAsync(c->sums_host, c->sums, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    local = c->sums_host[0];
    *force = local;
    return MACROC_OK;
}

extern "C" int macroc_time_step(macroc_ctx *c, int time_s, int *newton_its, double *res_norms, int *n_res,
                                int *ksp_its, double *ksp_rnorms)
{
    if (!c) return MACROC_ERR_ARG;
    // main.c:53-82
    double U = macroc_get_displacement(c, time_s);
    int rc = macroc_apply_bc_on_u(c, U);
    if (rc) return rc;
    int newton_it = 0, nres = 0;
    double norm = 0., norm_0 = 0.;
    while (newton_it < c->cfg.newton_max_its) {
        if ((rc = macroc_set_strains(c, 0))) return rc;
        if ((rc = macroc_homogenize(c))) return rc;                  /* main.c:62 */
        if ((rc = macroc_assembly_res(c, &norm))) return rc;
        if (res_norms) res_norms[nres] = norm;
        nres++;
        if (newton_it == 0) norm_0 = norm;
        if (norm < c->cfg.newton_min_tol || norm < norm_0 * c->cfg.newton_rel_tol) break;
        if ((rc = macroc_assembly_jac(c))) return rc;
        int its = 0; double rn = 0.;
        if ((rc = macroc_solve_Ax(c, &its, &rn))) return rc;
        if (ksp_its) ksp_its[newton_it] = its;
        if (ksp_rnorms) ksp_rnorms[newton_it] = rn;
        if ((rc = macroc_update_u(c))) return rc;
        newton_it++;
    }
    if (newton_its) *newton_its = newton_it;
    if (n_res) *n_res = nres;
    return MACROC_OK;
}

// ---------------------------------------------------------------------------
// boundary copies
// ---------------------------------------------------------------------------

extern "C" int64_t macroc_local_ndof(const macroc_ctx *c) { return c ? 3 * (int64_t)c->g.xm * c->g.ym * c->g.nzl : 0; }
extern "C" int64_t macroc_global_ndof(const macroc_ctx *c) { return c ? 3 * (int64_t)c->slab.gNX * c->slab.gNY * c->slab.gNZ : 0; }

static double *which_vec(macroc_ctx *c, int which)
{
    switch (which) {
        case MACROC_VEC_U: return c->vec[V_U];
        case MACROC_VEC_DU: return c->vec[V_DU];
        case MACROC_VEC_B: return c->vec[V_B];
        default: return nullptr;
    }
}

extern "C" int macroc_set_vec(macroc_ctx *c, int which, const double *host)
{
    if (!c || !host) return MACROC_ERR_ARG;
    double *v = which_vec(c, which);
    if (!v) FAIL(c, MACROC_ERR_ARG, "set_vec: unknown vector %d", which);
    CU(c, cudaSetDevice(c->device));
    int64_t n = macroc_local_ndof(c);
    CU(c, cudaMemcpyAsync(c->stage, host, sizeof(double) * n, cudaMemcpyHostToDevice, c->stream));
    LAUNCH(c, k_aos_to_soa, cdiv64(n, 256), 256, c->g, c->stage, v);
    CU(c, cudaStreamSynchronize(c->stream));
    return MACROC_OK;
}

extern "C" int macroc_get_vec(macroc_ctx *c, int which, double *host)
{
    if (!c || !host) return MACROC_ERR_ARG;
    double *v = which_vec(c, which);
    if (!v) FAIL(c, MACROC_ERR_ARG, "get_vec: unknown vector %d", which);
    CU(c, cudaSetDevice(c->device));
    int64_t n = macroc_local_ndof(c);
    LAUNCH(c, k_soa_to_aos, cdiv64(n, 256), 256, c->g, v, c->stage);
    CU(c, cudaMemcpyAsync(host, c->stage, sizeof(double) * n, cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    return MACROC_OK;
}

extern "C" int macroc_get_matrix_blocks(macroc_ctx *c, double *host)
{
    if (!c || !host) return MACROC_ERR_ARG;
    const bool sym = c->cfg.op == MACROC_OP_ASSEMBLED_SYM;
    if (sym ? !c->Asym_valid : !c->A_valid) FAIL(c, MACROC_ERR_ARG, "get_matrix_blocks: no assembled operator");
    CU(c, cudaSetDevice(c->device));
    const int64_t chunk = 1 << 16;                 // nodes per export chunk
    double *tmp = nullptr;
    CU(c, cudaMalloc(&tmp, sizeof(double) * 243 * chunk));
    const int64_t nown = (int64_t)c->g.xm * c->g.ym * c->g.nzl;
    for (int64_t n0 = 0; n0 < nown; n0 += chunk) {
        int64_t nn = std::min<int64_t>(chunk, nown - n0);
        if (sym) LAUNCH(c, k_export_blocks_sym, cdiv64(nn * 243, 256), 256, c->g, c->sg, reinterpret_cast<const double *>(c->Asym), n0, nn, tmp);
        else LAUNCH(c, k_export_blocks, cdiv64(nn * 243, 256), 256, c->g, reinterpret_cast<const double *>(c->A), n0, nn, tmp);
        cudaError_t e = cudaMemcpyAsync(host + n0 * 243, tmp, sizeof(double) * 243 * nn, cudaMemcpyDeviceToHost, c->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
        if (e != cudaSuccess) { cudaFree(tmp); FAIL(c, MACROC_ERR_CUDA, "get_matrix_blocks: %s", cudaGetErrorString(e)); }
    }
    cudaFree(tmp);
    return MACROC_OK;
}

extern "C" int macroc_matmult(macroc_ctx *c, int op, const double *x_host, double *y_host)
{
    if (!c || !x_host || !y_host) return MACROC_ERR_ARG;
    CU(c, cudaSetDevice(c->device));
    BIND_CONSTANTS(c);
    if (op != MACROC_OP_ASSEMBLED && op != MACROC_OP_MATRIX_FREE && op != MACROC_OP_ASSEMBLED_SYM)
        FAIL(c, MACROC_ERR_ARG, "matmult: unknown operator %d", op);
    if (op == MACROC_OP_ASSEMBLED && !c->A_valid) FAIL(c, MACROC_ERR_ARG, "matmult: no assembled operator");
    if (op == MACROC_OP_ASSEMBLED_SYM && !c->Asym_valid) FAIL(c, MACROC_ERR_ARG, "matmult: no symmetric operator");
    int64_t n = macroc_local_ndof(c);
    CU(c, cudaMemcpyAsync(c->stage, x_host, sizeof(double) * n, cudaMemcpyHostToDevice, c->stream));
    LAUNCH(c, k_aos_to_soa, cdiv64(n, 256), 256, c->g, c->stage, c->vec[V_P]);
    int rc = apply_operator(c, op, c->vec[V_P], c->vec[V_W], false, nullptr);
    if (rc) return rc;
    LAUNCH(c, k_soa_to_aos, cdiv64(n, 256), 256, c->g, c->vec[V_W], c->stage);
    CU(c, cudaMemcpyAsync(y_host, c->stage, sizeof(double) * n, cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    return MACROC_OK;
}

extern "C" int macroc_get_strain_stress(macroc_ctx *c, double *strain, double *stress, int64_t *n_gp)
{
    if (!c) return MACROC_ERR_ARG;
    CU(c, cudaSetDevice(c->device));
    int64_t ne = (int64_t)c->slab.nex * c->slab.ney * c->slab.nez;    // DMDA-owned elements
    if (n_gp) *n_gp = ne * 8;
    if ((strain || stress) && !c->strain && ne > 0) FAIL(c, MACROC_ERR_ARG, "get_strain_stress: call set_strains(ctx, 1) first");
    int rc = gp_download(c, c->strain, strain, 6);
    if (!rc) rc = gp_download(c, c->stress, stress, 6);
    return rc;
}

// write_pvtu (src/output.c:25-267).  Host-side formatting of data fetched through the same
// device paths the solver uses; the numbers printed are computed on the GPU.
extern "C" int macroc_write_pvtu(macroc_ctx *c, const char *file_prefix)
{
    if (!c || !file_prefix) return MACROC_ERR_ARG;
    CU(c, cudaSetDevice(c->device));
    BIND_CONSTANTS(c);
    const Slab &s = c->slab;
    const GridDev &g = c->g;
    char name[4096];
    if (s.rank == 0) {
        snprintf(name, sizeof(name), "%s.pvtu", file_prefix);
        FILE *fp = fopen(name, "w");
        if (!fp) FAIL(c, 65, "write_pvtu: cannot open %.400s", name);
        fprintf(fp,
                "<?xml version=\"1.0\"?>\n"
                "<VTKFile type=\"PUnstructuredGrid\" version=\"0.1\" byte_order=\"LittleEndian\">\n"
                "<PUnstructuredGrid GhostLevel=\"0\">\n"
                "<PPoints>\n"
                "  <PDataArray type=\"Float64\" Name=\"Position\"   NumberOfComponents=\"3\"/>\n"
                "</PPoints>\n"
                "<PCells>\n"
                "  <PDataArray type=\"Int32\" Name=\"connectivity\" NumberOfComponents=\"1\"/>\n"
                "  <PDataArray type=\"Int32\" Name=\"offsets\"      NumberOfComponents=\"1\"/>\n"
                "  <PDataArray type=\"UInt8\" Name=\"types\"        NumberOfComponents=\"1\"/>\n"
                "</PCells>\n"
                "<PPointData Vectors=\"displ\">\n"
                "  <PDataArray type=\"Float64\" Name=\"displ\"      NumberOfComponents=\"3\" />\n"
                "</PPointData>\n"
                "<PCellData>\n"
                "  <PDataArray type=\"Int32\"   Name=\"part\"       NumberOfComponents=\"1\"/>\n"
                "  <PDataArray type=\"Float64\" Name=\"cost\"       NumberOfComponents=\"1\"/>\n"
                "  <PDataArray type=\"Int32\"   Name=\"non-linear\" NumberOfComponents=\"1\"/>\n"
                "<PDataArray type=\"Float64\" Name=\"strain\"       NumberOfComponents=\"6\"/>\n"
                "<PDataArray type=\"Float64\" Name=\"stress\"       NumberOfComponents=\"6\"/>\n"
                "</PCellData>\n");
        for (int i = 0; i < s.nranks; ++i) fprintf(fp, "  <Piece Source=\"%s-subdo-%d.vtu\"/>\n", file_prefix, i);
        fprintf(fp, "</PUnstructuredGrid>\n</VTKFile>\n");
        fclose(fp);
    }
    // ghosted displacement (DMGlobalToLocal, output.c:150-153) and Gauss-point strain / stress
    int rc = halo_exchange(c, c->vec[V_U], c->stream);
    if (rc) return rc;
    const int64_t N = g.npl * s.Zm, node0 = (int64_t)(s.Zs - s.zs) * g.npl;   // the ghosted box (x/y ghosts are local nodes)
    std::vector<double> u_loc((size_t)3 * N);
    {
        double *tmp = nullptr;
        CU(c, cudaMalloc(&tmp, sizeof(double) * 3 * (size_t)N));
        LAUNCH(c, k_soa_to_aos_range, cdiv64(3 * N, 256), 256, g, c->vec[V_U], node0, N, tmp);
        cudaError_t e = cudaMemcpyAsync(u_loc.data(), tmp, sizeof(double) * 3 * (size_t)N, cudaMemcpyDeviceToHost, c->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
        cudaFree(tmp);
        if (e != cudaSuccess) FAIL(c, MACROC_ERR_CUDA, "write_pvtu: %s", cudaGetErrorString(e));
    }
    const int64_t nelem = (int64_t)s.nex * s.ney * s.nez;                      // DMDA-owned elements
    std::vector<double> eps((size_t)48 * std::max<int64_t>(nelem, 1)), sig((size_t)48 * std::max<int64_t>(nelem, 1));
    if (c->cfg.material == MACROC_MAT_PER_GP) {
        // strain from u (output.c:219-231), stress as the material model left it (output.c:245)
        if ((rc = ensure_gp_arrays(c, false))) return rc;
        if (c->ne_owned > 0) LAUNCH(c, k_strain_stress, cdiv64(c->ne_owned, 128), 128, g, s.ezs, s.nez, c->er.ne_ext, c->vec[V_U], c->strain, (double *)nullptr);
    } else if ((rc = macroc_set_strains(c, 1)))
        return rc;
    if ((rc = gp_download(c, c->strain, eps.data(), 6))) return rc;
    if ((rc = gp_download(c, c->stress, sig.data(), 6))) return rc;

    snprintf(name, sizeof(name), "%s-subdo-%d.vtu", file_prefix, s.rank);
    FILE *fp = fopen(name, "w");
    if (!fp) FAIL(c, 65, "write_pvtu: cannot open %.400s", name);
    fprintf(fp,
            "<?xml version=\"1.0\"?>\n"
            "<VTKFile type=\"UnstructuredGrid\" version=\"0.1\" byte_order=\"LittleEndian\">\n"
            "<UnstructuredGrid>\n"
            "<Piece NumberOfPoints=\"%d\" NumberOfCells=\"%d\">\n"
            "<Points>\n", (int)N, (int)nelem);
    fprintf(fp, "<DataArray type=\"Float64\" Name=\"Position\" NumberOfComponents=\"3\" format=\"ascii\">\n");
    for (int k = s.Zs; k < s.Zs + s.Zm; ++k)
        for (int j = s.Ys; j < s.Ys + s.Ym; ++j)
            for (int i = s.Xs; i < s.Xs + s.Xm; ++i) fprintf(fp, "%01.6e\t%01.6e\t%01.6e\n", i * c->geo.dx, j * c->geo.dy, k * c->geo.dz);
    fprintf(fp, "</DataArray>\n</Points>\n<Cells>\n");
    fprintf(fp, "<DataArray type=\"Int32\" Name=\"connectivity\" NumberOfComponents=\"1\" format=\"ascii\">\n");
    for (int ek = 0; ek < s.nez; ++ek)
        for (int ej = 0; ej < s.ney; ++ej)
            for (int ei = 0; ei < s.nex; ++ei) {
                // DMDAGetElements: local ghosted ids, counter-clockwise bottom face then top face
                const int64_t sx = 1, sy = s.NX, sz = g.npl;
                const int64_t b0 = (s.exs - s.Xs + ei) + sy * (s.eys - s.Ys + ej) + sz * (int64_t)(s.ezs + ek - s.Zs);
                const int64_t ids[8] = {b0, b0 + sx, b0 + sx + sy, b0 + sy, b0 + sz, b0 + sx + sz, b0 + sx + sy + sz, b0 + sy + sz};
                for (int n = 0; n < 8; ++n) fprintf(fp, "%-6d\t", (int)ids[n]);
                fprintf(fp, "\n");
            }
    fprintf(fp, "</DataArray>\n");
    fprintf(fp, "<DataArray type=\"Int32\" Name=\"offsets\" NumberOfComponents=\"1\" format=\"ascii\">\n");
    for (int64_t e = 1; e < nelem + 1; ++e) fprintf(fp, "%d\t", (int)(e * 8));
    fprintf(fp, "\n</DataArray>\n");
    fprintf(fp, "<DataArray type=\"UInt8\"  Name=\"types\" NumberOfComponents=\"1\" format=\"ascii\">\n");
    for (int64_t e = 0; e < nelem; ++e) fprintf(fp, "12\t");
    fprintf(fp, "\n</DataArray>\n</Cells>\n<PointData Vectors=\"displ\">\n");
    fprintf(fp, "<DataArray type=\"Float64\" Name=\"displ\" NumberOfComponents=\"3\" format=\"ascii\" >\n");
    for (int64_t n = 0; n < N; ++n) fprintf(fp, "%01.6e\t%01.6e\t%01.6e\n", u_loc[3 * n], u_loc[3 * n + 1], u_loc[3 * n + 2]);
    fprintf(fp, "</DataArray>\n</PointData>\n<CellData>\n");
    fprintf(fp, "<DataArray type=\"Int32\" Name=\"part\" NumberOfComponents=\"1\" format=\"ascii\">\n");
    for (int64_t e = 0; e < nelem; ++e) fprintf(fp, "%d\t", s.rank);
    fprintf(fp, "\n</DataArray>\n");
    fprintf(fp, "<DataArray type=\"Float64\" Name=\"cost\" NumberOfComponents=\"1\" format=\"ascii\">\n");
    for (int64_t e = 0; e < nelem; ++e) fprintf(fp, "%lf\t", 0.);
    fprintf(fp, "\n</DataArray>\n");
    fprintf(fp, "<DataArray type=\"Int32\" Name=\"non-linear\" NumberOfComponents=\"1\" format=\"ascii\">\n");
    for (int64_t e = 0; e < nelem; ++e) fprintf(fp, "%d\t", 0);
    fprintf(fp, "\n</DataArray>\n");
    for (int pass = 0; pass < 2; ++pass) {
        const std::vector<double> &v = pass == 0 ? eps : sig;
        fprintf(fp, "<DataArray type=\"Float64\" Name=\"%s\" NumberOfComponents=\"6\" format=\"ascii\">", pass == 0 ? "strain" : "stress");
        for (int64_t e = 0; e < nelem; ++e) {
            double acc[6] = {0, 0, 0, 0, 0, 0};
            for (int gp = 0; gp < 8; ++gp)
                for (int i = 0; i < 6; ++i) acc[i] += v[(size_t)(e * 8 + gp) * 6 + i] * c->geo.wg;     // output.c:230,247
            for (int i = 0; i < 6; ++i) fprintf(fp, "%e\t", acc[i]);
        }
        fprintf(fp, "\n</DataArray>\n");
    }
    fprintf(fp, "</CellData>\n</Piece>\n</UnstructuredGrid>\n</VTKFile>\n");
    fclose(fp);
    return MACROC_OK;
}

// ---------------------------------------------------------------------------
// measurement
// ---------------------------------------------------------------------------

extern "C" uint64_t macroc_launch_count(const macroc_ctx *c) { return c ? c->launches : 0; }

extern "C" int macroc_event_record(macroc_ctx *c, int slot)
{
    if (!c || slot < 0 || slot >= 8) return MACROC_ERR_ARG;
    CU(c, cudaSetDevice(c->device));
    CU(c, cudaEventRecord(c->ev_user[slot], c->stream));
    return MACROC_OK;
}

extern "C" int macroc_event_elapsed_ms(macroc_ctx *c, int a, int b, double *ms)
{
    if (!c || !ms || a < 0 || a >= 8 || b < 0 || b >= 8) return MACROC_ERR_ARG;
    CU(c, cudaSetDevice(c->device));
    CU(c, cudaEventSynchronize(c->ev_user[b]));
    float f = 0.f;
    CU(c, cudaEventElapsedTime(&f, c->ev_user[a], c->ev_user[b]));
    *ms = f;
    return MACROC_OK;
}

extern "C" int macroc_profile_enable(macroc_ctx *c, int enable, int stride)
{
    if (!c) return MACROC_ERR_ARG;
    CU(c, cudaSetDevice(c->device));
    if (enable && !c->prof_ev[0])
        for (cudaEvent_t &e : c->prof_ev) CU(c, cudaEventCreate(&e));
    c->prof_on = enable != 0;
    c->prof_stride = stride > 0 ? stride : 1;
    c->prof_used = 0; c->prof_counter = 0; c->prof_samples = 0; c->prof_ms = 0.;
    c->prof_solve_ms = 0.; c->prof_solve_its = 0;
    return MACROC_OK;
}

extern "C" int macroc_profile_get_solve(macroc_ctx *c, double *solve_ms_total, int64_t *iterations)
{
    if (!c) return MACROC_ERR_ARG;
    if (solve_ms_total) *solve_ms_total = c->prof_solve_ms;
    if (iterations) *iterations = c->prof_solve_its;
    return MACROC_OK;
}

extern "C" int macroc_profile_get(macroc_ctx *c, double *apply_ms_mean, int64_t *samples)
{
    if (!c) return MACROC_ERR_ARG;
    if (apply_ms_mean) *apply_ms_mean = c->prof_samples ? c->prof_ms / (double)c->prof_samples : 0.;
    if (samples) *samples = c->prof_samples;
    return MACROC_OK;
}

extern "C" int macroc_device_synchronize(macroc_ctx *c)
{
    if (!c) return MACROC_ERR_ARG;
    CU(c, cudaSetDevice(c->device));
    CU(c, cudaStreamSynchronize(c->stream));
    CU(c, cudaStreamSynchronize(c->comm_stream));
    return MACROC_OK;
}

extern "C" int macroc_allreduce_path(const macroc_ctx *c)
{
    if (!c) return -1;
    if (c->loop) return 3;
    if (!c->comm) return 0;
    return c->mbox_on ? 2 : 1;
}

extern "C" int macroc_fp64_probe(macroc_ctx *c, double *tflops)
{
    if (!c || !tflops) return MACROC_ERR_ARG;
    CU(c, cudaSetDevice(c->device));
    const int blocks = 148 * 8;                      // 8 CTAs of 256 threads per SM: every scheduler has 16 warps to pick from
    BIND_CONSTANTS(c);
    double best = 0.;
    // two operand forms: three register operands (register-port bound, ~33 TFLOP/s) and a constant-bank
    // multiplier (the form of the unrolled element kernel; ~37 TFLOP/s): the better one is the pipe's rate
    for (int mode = 0; mode < 2; ++mode)
        for (int r = 0; r < 6; ++r) {
            CU(c, cudaEventRecord(c->ev_t0, c->stream));
            if (mode) LAUNCH(c, k_fp64_probe_const, blocks, 256, c->sums, 1.0 + r);
            else LAUNCH(c, k_fp64_probe, blocks, 256, c->sums, 1.0 + r);
            CU(c, cudaEventRecord(c->ev_t1, c->stream));
            CU(c, cudaEventSynchronize(c->ev_t1));
            float ms = 0.f;
            CU(c, cudaEventElapsedTime(&ms, c->ev_t0, c->ev_t1));
            const double flops = 2.0 * FP64_PROBE_CHAINS * (double)FP64_PROBE_ITERS * 256.0 * blocks;
            if (r >= 1 && ms > 0.f) best = std::max(best, flops / (ms * 1e-3) / 1e12);   // first launch: warm-up
        }
    CU(c, cudaGetLastError());
    *tflops = best;
    return MACROC_OK;
}

extern "C" int macroc_dmma_probe(macroc_ctx *c, double *tflops)
{
    if (!c || !tflops) return MACROC_ERR_ARG;
    CU(c, cudaSetDevice(c->device));
    const int blocks = 148 * 8;
    double best = 0.;
    for (int r = 0; r < 6; ++r) {
        CU(c, cudaEventRecord(c->ev_t0, c->stream));
        LAUNCH(c, k_dmma_probe, blocks, 256, c->sums, 1.0 + r);
        CU(c, cudaEventRecord(c->ev_t1, c->stream));
        CU(c, cudaEventSynchronize(c->ev_t1));
        float ms = 0.f;
        CU(c, cudaEventElapsedTime(&ms, c->ev_t0, c->ev_t1));
        const double flops = 512.0 * DMMA_PROBE_CHAINS * (double)DMMA_PROBE_ITERS * 8.0 * blocks;   // 8 warps per CTA, 2 x 8 x 8 x 4 per DMMA
        if (r >= 1 && ms > 0.f) best = std::max(best, flops / (ms * 1e-3) / 1e12);
    }
    CU(c, cudaGetLastError());
    *tflops = best;
    return MACROC_OK;
}

extern "C" int macroc_contraction_ab(macroc_ctx *c, int variant, int reps, int n_full, double *full_host, double *ms_mean)
{
    if (!c || !ms_mean || reps <= 0 || variant < 0 || variant > 2 || n_full < 0 || (n_full > 0 && !full_host)) return MACROC_ERR_ARG;
    if (!c->ctan) FAIL(c, MACROC_ERR_ARG, "contraction_ab: no Gauss-point tangents (material MACROC_MAT_PER_GP, call homogenize)");
    CU(c, cudaSetDevice(c->device));
    BIND_CONSTANTS(c);
    const int64_t ne = c->er.ne_ext;
    double *rowsum = nullptr, *full = nullptr;
    CU(c, cudaMalloc(&rowsum, sizeof(double) * 24 * (size_t)ne));
    if (cudaMalloc(&full, sizeof(double) * 576 * (size_t)std::max(1, n_full)) != cudaSuccess) { cudaFree(rowsum); FAIL(c, MACROC_ERR_MEM, "contraction_ab: out of memory"); }
    cudaError_t ea = cudaSuccess;
    if (variant == 0) ea = cudaFuncSetAttribute(k_ab_dfma, cudaFuncAttributeMaxDynamicSharedMemorySize, AB_SMEM_DFMA);
    else if (variant == 1) ea = cudaFuncSetAttribute(k_ab_dmma<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, AB_SMEM_DMMA);
    else ea = cudaFuncSetAttribute(k_ab_dmma<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, AB_SMEM_DMMA);
    int rc = MACROC_OK;
    double total = 0.;
    const int blocks = (int)std::min<int64_t>((ne + AB_ELEMS - 1) / AB_ELEMS, 148);
    for (int r = 0; r < reps + 1 && ea == cudaSuccess; ++r) {              // the first launch is a warm-up
        cudaEventRecord(c->ev_t0, c->stream);
        if (variant == 0) k_ab_dfma<<<blocks, AB_DFMA_THREADS, AB_SMEM_DFMA, c->stream>>>(c->ctan, ne, ne, c->geo.wg, rowsum, full, n_full);
        else if (variant == 1) k_ab_dmma<false><<<blocks, AB_DMMA_WARPS * 32, AB_SMEM_DMMA, c->stream>>>(c->ctan, ne, ne, c->geo.wg, rowsum, full, n_full);
        else k_ab_dmma<true><<<blocks, AB_DMMA_WARPS * 32, AB_SMEM_DMMA, c->stream>>>(c->ctan, ne, ne, c->geo.wg, rowsum, full, n_full);
        c->launches++;
        cudaEventRecord(c->ev_t1, c->stream);
        ea = cudaEventSynchronize(c->ev_t1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, c->ev_t0, c->ev_t1);
        if (r >= 1) total += ms;
    }
    if (ea == cudaSuccess) ea = cudaGetLastError();
    if (ea == cudaSuccess && n_full > 0) ea = cudaMemcpy(full_host, full, sizeof(double) * 576 * (size_t)n_full, cudaMemcpyDeviceToHost);
    cudaFree(rowsum); cudaFree(full);
    if (ea != cudaSuccess) FAIL(c, MACROC_ERR_CUDA, "contraction_ab: %s", cudaGetErrorString(ea));
    *ms_mean = total / reps;
    return rc;
}

extern "C" int macroc_time_kernel(macroc_ctx *c, int what, int reps, int flush_l2, double *ms_mean)
{
    if (!c || !ms_mean || reps <= 0) return MACROC_ERR_ARG;
    CU(c, cudaSetDevice(c->device));
    BIND_CONSTANTS(c);
    const GridDev &g = c->g;
    if ((what == 0 || what == 2) && !c->A_valid) FAIL(c, MACROC_ERR_ARG, "time_kernel: assemble first");
    if ((what == 8 || what == 9) && !c->Asym_valid) FAIL(c, MACROC_ERR_ARG, "time_kernel: assemble the symmetric operator first");
    if (flush_l2 && !c->flush) {
        c->flush_bytes = (size_t)256 << 20;
        CU(c, cudaMalloc(&c->flush, c->flush_bytes));
    }
    if (what == 0 || what == 1 || what == 8) LAUNCH(c, k_fill_pattern, cdiv64(g.nloc, 256), 256, g, c->nodemask, c->vec[V_P]);
    if (what == 2 || what == 5 || what == 9) {
        // a never-converging PCG on a synthetic right-hand side: every iteration does the full work
        if (what == 5) LAUNCH(c, k_mf_diag, cdiv64(g.nloc, 256), 256, g, c->T, c->nodemask, c->vec[V_DINV]);
        LAUNCH(c, k_fill_pattern, cdiv64(g.nloc, 256), 256, g, c->nodemask, c->vec[V_B]);
        int rc = cg_begin(c, 0., 0., 1e300, 1 << 30);
        if (rc) return rc;
    }
    double total = 0.;
    for (int r = 0; r < reps; ++r) {
        if (flush_l2) CU(c, cudaMemsetAsync(c->flush, r & 0xff, c->flush_bytes, c->stream));
        CU(c, cudaEventRecord(c->ev_t0, c->stream));
        int rc = MACROC_OK;
        switch (what) {
            case 0: rc = apply_operator(c, MACROC_OP_ASSEMBLED, c->vec[V_P], c->vec[V_W], true, nullptr); break;
            case 1: rc = apply_operator(c, MACROC_OP_MATRIX_FREE, c->vec[V_P], c->vec[V_W], true, nullptr); break;
            case 2: rc = cg_iteration(c, MACROC_OP_ASSEMBLED); break;
            case 5: rc = cg_iteration(c, MACROC_OP_MATRIX_FREE); break;
            case 8: rc = apply_operator(c, MACROC_OP_ASSEMBLED_SYM, c->vec[V_P], c->vec[V_W], true, nullptr); break;
            case 9: rc = cg_iteration(c, MACROC_OP_ASSEMBLED_SYM); break;
            case 3: {
                int save = c->cfg.op; c->cfg.op = MACROC_OP_ASSEMBLED;
                rc = macroc_assembly_jac(c);
                c->cfg.op = save;
                break;
            }
            case 4: {
                int nparts = 0;
                rc = residual_launch(c, &nparts);
                if (!rc) LAUNCH(c, k_reduce, 1, 256, c->partial, nparts, c->sums);
                break;
            }
            case 7: case 17: {                          // per-element Jacobian kernel (tangent source per cfg.material); 17: symmetric layout
                int save_op = c->cfg.op, save_j = c->cfg.jac_mode;
                c->cfg.op = what == 17 ? MACROC_OP_ASSEMBLED_SYM : MACROC_OP_ASSEMBLED; c->cfg.jac_mode = MACROC_JAC_ELEMENT;
                rc = macroc_assembly_jac(c);
                c->cfg.op = save_op; c->cfg.jac_mode = save_j;
                break;
            }
            default: FAIL(c, MACROC_ERR_ARG, "time_kernel: unknown kernel %d", what);
        }
        if (rc) return rc;
        CU(c, cudaEventRecord(c->ev_t1, c->stream));
        CU(c, cudaEventSynchronize(c->ev_t1));
        float ms = 0.f;
        CU(c, cudaEventElapsedTime(&ms, c->ev_t0, c->ev_t1));
        total += ms;
    }
    *ms_mean = total / reps;
    return MACROC_OK;
}
