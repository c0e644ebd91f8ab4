// grid.h -- host-side partition and Dirichlet bookkeeping (no CUDA here).
//
// Restates what the reference inherits from PETSc's DMDA for the z-slab
// decompositions this build shards over (-da_processors_x 1 -da_processors_y 1
// -da_processors_z P) and what bc_init builds
// (reference src/bcs.c:154-338).  Node numbering is the DMDA natural one,
// n = i + NX*(j + NY*k); with z-slabs PETSc's global numbering equals it.
#pragma once

#include <cstdint>
#include <cstring>
#include <vector>

#include "../../include/macroc_b200.h"

namespace macroc {

struct Slab {
    int NX = 0, NY = 0, NZ = 0;
    int rank = 0, nranks = 1;
    int zs = 0, nzl = 0;        // owned node planes [zs, zs+nzl)   (DMDAGetCorners)
    int Zs = 0, Zm = 0;         // ghosted plane range              (DMDAGetGhostCorners)
    int nex = 0, ney = 0, nez = 0, ezs = 0;   // DMDAGetElementsSizes; first element layer
    int64_t npl = 0, nloc = 0;  // nodes per plane, owned nodes
    bool has_lower() const { return zs > 0; }
    bool has_upper() const { return zs + nzl < NZ; }
};

// PETSc ownership rule along one axis: M/m + ((M % m) > i)
inline void split_axis(int M, int m, int i, int *start, int *count)
{
    int s = 0;
    for (int q = 0; q < i; ++q) s += M / m + ((M % m) > q);
    *start = s;
    *count = M / m + ((M % m) > i);
}

inline int make_slab(const macroc_config &cfg, int rank, int nranks, Slab *out)
{
    int px = cfg.px > 0 ? cfg.px : 1, py = cfg.py > 0 ? cfg.py : 1;
    int pz = cfg.pz > 0 ? cfg.pz : nranks;
    if (px != 1 || py != 1) return MACROC_ERR_UNSUPPORTED;     // z-slabs only
    if (pz != nranks || rank < 0 || rank >= nranks) return MACROC_ERR_ARG;
    if (cfg.NX < 2 || cfg.NY < 2 || cfg.NZ < 2 || nranks > cfg.NZ) return MACROC_ERR_ARG;
    Slab s;
    s.NX = cfg.NX; s.NY = cfg.NY; s.NZ = cfg.NZ; s.rank = rank; s.nranks = nranks;
    split_axis(cfg.NZ, pz, rank, &s.zs, &s.nzl);
    s.Zs = s.zs > 0 ? s.zs - 1 : 0;
    int Ze = s.zs + s.nzl < cfg.NZ ? s.zs + s.nzl + 1 : cfg.NZ;
    s.Zm = Ze - s.Zs;
    // DMDAGetElements: a rank owns the cells whose upper corner node it owns
    s.ezs = s.zs != s.Zs ? s.zs - 1 : s.zs;
    s.nex = cfg.NX - 1; s.ney = cfg.NY - 1; s.nez = (s.zs + s.nzl) - s.ezs - 1;
    if (s.nez < 0) s.nez = 0;
    s.npl = (int64_t)cfg.NX * cfg.NY;
    s.nloc = s.npl * s.nzl;
    *out = s;
    return MACROC_OK;
}

struct Geometry {
    double dx, dy, dz, wg, rad;   // init.c:137-141
};

inline Geometry make_geometry(const macroc_config &cfg)
{
    Geometry g;
    g.dx = cfg.lx / (cfg.NX - 1);
    g.dy = cfg.ly / (cfg.NY - 1);
    g.dz = cfg.lz / (cfg.NZ - 1);
    g.wg = g.dx * g.dy * g.dz / 8;
    g.rad = 1.;
    return g;
}

// The cell-centre-like circle test of bcs.c:132-134 / :324-327 / forces.c:138-141.
inline bool in_circle(const macroc_config &cfg, const Geometry &g, int gi, int gk)
{
    double x = cfg.lx / 2. - (gi * g.dx + g.dx / 2.);
    double z = cfg.lz / 2. - (gk * g.dz + g.dz / 2.);
    return (x * x + z * z) < (g.rad * g.rad);
}

// bc_init_bending (bcs.c:198-251) / bc_init_circle (bcs.c:254-338) over the
// rank's GHOSTED box, with the values of bc_apply_on_u_* (bcs.c:61-146)
// expressed as coef*U.  idx holds global dof ids, -1 padded to nbcs.
inline void build_bc_lists(const macroc_config &cfg, const Slab &s, std::vector<int32_t> &idx,
                           std::vector<double> &coef)
{
    const Geometry g = make_geometry(cfg);
    const int nxg = s.NX, nyg = s.NY, nzg = s.Zm;         // ghost box of a z-slab
    const int si = 0, sj = 0, sk = s.Zs;
    auto gdof = [&](int i, int j, int k, int d) {          // local ghosted -> global dof
        return (int32_t)(((int64_t)(si + i) + (int64_t)s.NX * ((sj + j) + (int64_t)s.NY * (sk + k))) * 3 + d);
    };
    idx.clear(); coef.clear();
    if (cfg.bc_type == MACROC_BC_BENDING) {
        int nbcs = 2 * nyg * nzg * 3;
        // X = 0 : (0,0,0);  X = LX : (0,U,0).  Both faces lie in every z-slab.
        for (int face = 0; face < 2; ++face) {
            int i = face == 0 ? 0 : nxg - 1;
            for (int k = 0; k < nzg; ++k)
                for (int j = 0; j < nyg; ++j)
                    for (int d = 0; d < 3; ++d) {
                        idx.push_back(gdof(i, j, k, d));
                        coef.push_back(face == 1 && d == 1 ? 1. : 0.);
                    }
        }
        idx.resize(nbcs, -1); coef.resize(nbcs, 0.);
    } else {
        int nbcs = (2 * nxg + 2 * nzg) * 3 + nxg * nzg;
        for (int k = 0; k < nzg; ++k)                       // X=0 & Y=0 along z
            for (int d = 0; d < 3; ++d) { idx.push_back(gdof(0, 0, k, d)); coef.push_back(0.); }
        for (int k = 0; k < nzg; ++k)                       // X=LX & Y=0 along z
            for (int d = 0; d < 3; ++d) { idx.push_back(gdof(nxg - 1, 0, k, d)); coef.push_back(0.); }
        if (sk == 0)                                        // Z=0 & Y=0 along x
            for (int i = 1; i < nxg - 1; ++i)
                for (int d = 0; d < 3; ++d) { idx.push_back(gdof(i, 0, 0, d)); coef.push_back(0.); }
        if (sk + nzg == s.NZ)                               // Z=LZ & Y=0 along x
            for (int i = 1; i < nxg - 1; ++i)
                for (int d = 0; d < 3; ++d) { idx.push_back(gdof(i, 0, nzg - 1, d)); coef.push_back(0.); }
        for (int i = 0; i < nxg; ++i)                       // circle on Y = LY, dof y only
            for (int k = 0; k < nzg; ++k)
                if (in_circle(cfg, g, si + i, sk + k)) { idx.push_back(gdof(i, nyg - 1, k, 1)); coef.push_back(1.); }
        idx.resize(nbcs, -1); coef.resize(nbcs, 0.);
    }
}

}  // namespace macroc
