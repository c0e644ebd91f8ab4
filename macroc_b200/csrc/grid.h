// grid.h -- host-side partition and Dirichlet bookkeeping (no CUDA here).
//
// Restates what the reference inherits from PETSc's DMDA (init.c:85-94,167-171): the
// processor grid (-da_processors_x/y/z or PETSC_DECIDE), the ownership split, ghost corners,
// element ownership, the rank-contiguous global numbering, and what bc_init builds on top of it
// (reference src/bcs.c:154-338).
//
// What the kernels see: a rank's LOCAL box.  In x and y it is the ghosted extent [Xs, Xs+Xm) x
// [Ys, Ys+Ym) (ghost columns/rows of x/y neighbours are ordinary local nodes whose rows are
// computed but never used), in z it is the owned planes with the ghost planes outside the
// owned range (z-slabs are the sharding the benchmarks use; x/y splits are the
// general DMDA case).  Because a ghost layer exists wherever a neighbour exists, an OWNED node
// sits on the edge of the local box only if it sits on the edge of the global grid, so node
// classes and element existence evaluated in local x/y coordinates are exact for owned nodes.
#pragma once

#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../include/macroc_b200.h"

namespace macroc {

struct Slab {
    int gNX = 0, gNY = 0, gNZ = 0;              // global grid
    int px = 1, py = 1, pz = 1, pi = 0, pj = 0, pk = 0;
    int rank = 0, nranks = 1;
    int xs = 0, xm = 0, ys = 0, ym = 0, zs = 0, nzl = 0;     // owned box        (DMDAGetCorners)
    int Xs = 0, Xm = 0, Ys = 0, Ym = 0, Zs = 0, Zm = 0;      // ghosted box      (DMDAGetGhostCorners)
    int NX = 0, NY = 0, NZ = 0;                 // local array extents in x, y (= Xm, Ym); NZ = global
    int nex = 0, ney = 0, nez = 0;              // DMDAGetElementsSizes
    int exs = 0, eys = 0, ezs = 0;              // first owned element per axis (global index)
    int lnex = 0, lney = 0;                     // elements of the local box per row / rows per layer
    int64_t npl = 0, nloc = 0;                  // local nodes per plane, local nodes of the owned planes
    int nb[6] = {-1, -1, -1, -1, -1, -1};       // neighbour ranks: x-, x+, y-, y+, z-, z+
    bool has_lower() const { return nb[4] >= 0; }
    bool has_upper() const { return nb[5] >= 0; }
    bool xy_split() const { return px * py > 1; }
};

// PETSc ownership rule along one axis: M/m + ((M % m) > i)
inline void split_axis(int M, int m, int i, int *start, int *count)
{
    int s = 0;
    for (int q = 0; q < i; ++q) s += M / m + ((M % m) > q);
    *start = s;
    *count = M / m + ((M % m) > i);
}

// PETSC_DECIDE processor grid of a 3-D DMDA (restated from PETSc's DMSetUp_DA_3D; not verified
// against a PETSc build -- none exists in this image).  m, n, p <= 0 mean "decide".
inline int decide_proc_grid(int M, int N, int P, int size, int *m_, int *n_, int *p_)
{
    int m = *m_ > 0 ? *m_ : 0, n = *n_ > 0 ? *n_ : 0, p = *p_ > 0 ? *p_ : 0;
    auto squarish = [&](double A, double B, int fixed, int *a, int *b) {
        // a ~ sqrt(A*size/(B*fixed)), the largest a <= that with a*b*fixed == size
        int aa = (int)(0.5 + std::sqrt(A * (double)size / (B * (double)fixed)));
        if (!aa) aa = 1;
        int bb = 0;
        while (aa > 0) { bb = size / (aa * fixed); if (aa * bb * fixed == size) break; aa--; }
        if (!aa) return false;
        if (A > B && aa < bb) { int t = aa; aa = bb; bb = t; }
        *a = aa; *b = bb;
        return true;
    };
    if (m && n && p) { /* fully specified */ }
    else if (!m && n && p) m = size / (n * p);
    else if (m && !n && p) n = size / (m * p);
    else if (m && n && !p) p = size / (m * n);
    else if (!m && !n && p) { if (!squarish(M, N, p, &m, &n)) return MACROC_ERR_ARG; }
    else if (!m && n && !p) { if (!squarish(M, P, n, &m, &p)) return MACROC_ERR_ARG; }
    else if (m && !n && !p) { if (!squarish(N, P, m, &n, &p)) return MACROC_ERR_ARG; }
    else {
        n = (int)(0.5 + std::pow(((double)N * N) * ((double)size) / ((double)P * M), 1. / 3.));
        if (!n) n = 1;
        while (n > 0) { int pm = size / n; if (n * pm == size) break; n--; }
        if (!n) n = 1;
        m = (int)(0.5 + std::sqrt(((double)M) * ((double)size) / ((double)P * n)));
        if (!m) m = 1;
        while (m > 0) { p = size / (m * n); if (m * n * p == size) break; m--; }
        if (!m) return MACROC_ERR_ARG;
        if (M > P && m < p) { int t = m; m = p; p = t; }
    }
    if (m < 1 || n < 1 || p < 1 || m * n * p != size || m > M || n > N || p > P) return MACROC_ERR_ARG;
    *m_ = m; *n_ = n; *p_ = p;
    return MACROC_OK;
}

inline int make_slab(const macroc_config &cfg, int rank, int nranks, Slab *out)
{
    if (cfg.NX < 2 || cfg.NY < 2 || cfg.NZ < 2 || nranks < 1 || rank < 0 || rank >= nranks) return MACROC_ERR_ARG;
    Slab s;
    s.gNX = cfg.NX; s.gNY = cfg.NY; s.gNZ = cfg.NZ; s.rank = rank; s.nranks = nranks;
    s.px = cfg.px; s.py = cfg.py; s.pz = cfg.pz;
    int rc = decide_proc_grid(cfg.NX, cfg.NY, cfg.NZ, nranks, &s.px, &s.py, &s.pz);
    if (rc) return rc;
    s.pi = rank % s.px; s.pj = (rank / s.px) % s.py; s.pk = rank / (s.px * s.py);   // rank = i + j m + k m n
    split_axis(cfg.NX, s.px, s.pi, &s.xs, &s.xm);
    split_axis(cfg.NY, s.py, s.pj, &s.ys, &s.ym);
    split_axis(cfg.NZ, s.pz, s.pk, &s.zs, &s.nzl);
    // ghost corners: box stencil width 1, DM_BOUNDARY_NONE (init.c:85-90)
    s.Xs = s.xs > 0 ? s.xs - 1 : 0; s.Ys = s.ys > 0 ? s.ys - 1 : 0; s.Zs = s.zs > 0 ? s.zs - 1 : 0;
    s.Xm = (s.xs + s.xm < cfg.NX ? s.xs + s.xm + 1 : cfg.NX) - s.Xs;
    s.Ym = (s.ys + s.ym < cfg.NY ? s.ys + s.ym + 1 : cfg.NY) - s.Ys;
    s.Zm = (s.zs + s.nzl < cfg.NZ ? s.zs + s.nzl + 1 : cfg.NZ) - s.Zs;
    // DMDAGetElements: a rank owns the cells whose upper corner node it owns
    s.exs = s.xs != s.Xs ? s.xs - 1 : s.xs; s.eys = s.ys != s.Ys ? s.ys - 1 : s.ys; s.ezs = s.zs != s.Zs ? s.zs - 1 : s.zs;
    s.nex = s.xs + s.xm - s.exs - 1; s.ney = s.ys + s.ym - s.eys - 1; s.nez = s.zs + s.nzl - s.ezs - 1;
    if (s.nex < 0) s.nex = 0;
    if (s.ney < 0) s.ney = 0;
    if (s.nez < 0) s.nez = 0;
    s.NX = s.Xm; s.NY = s.Ym; s.NZ = cfg.NZ;
    s.lnex = s.NX - 1; s.lney = s.NY - 1;
    s.npl = (int64_t)s.NX * s.NY;
    s.nloc = s.npl * s.nzl;
    auto rk = [&](int i, int j, int k) { return i + j * s.px + k * s.px * s.py; };
    if (s.pi > 0) s.nb[0] = rk(s.pi - 1, s.pj, s.pk);
    if (s.pi < s.px - 1) s.nb[1] = rk(s.pi + 1, s.pj, s.pk);
    if (s.pj > 0) s.nb[2] = rk(s.pi, s.pj - 1, s.pk);
    if (s.pj < s.py - 1) s.nb[3] = rk(s.pi, s.pj + 1, s.pk);
    if (s.pk > 0) s.nb[4] = rk(s.pi, s.pj, s.pk - 1);
    if (s.pk < s.pz - 1) s.nb[5] = rk(s.pi, s.pj, s.pk + 1);
    *out = s;
    return MACROC_OK;
}

// PETSc's global node numbering of the DMDA: rank-contiguous, x fastest inside a rank's owned box.
struct GlobalNumbering {
    int NX, NY, NZ, px, py, pz;
    std::vector<int> ox, oy, oz, cx, cy, cz;     // start / count per processor coordinate
    std::vector<int64_t> off;                    // first node of each rank
    GlobalNumbering(const Slab &s) : NX(s.gNX), NY(s.gNY), NZ(s.gNZ), px(s.px), py(s.py), pz(s.pz)
    {
        ox.resize(px); cx.resize(px); oy.resize(py); cy.resize(py); oz.resize(pz); cz.resize(pz);
        for (int i = 0; i < px; ++i) split_axis(NX, px, i, &ox[i], &cx[i]);
        for (int i = 0; i < py; ++i) split_axis(NY, py, i, &oy[i], &cy[i]);
        for (int i = 0; i < pz; ++i) split_axis(NZ, pz, i, &oz[i], &cz[i]);
        off.resize((size_t)px * py * pz);
        int64_t o = 0;
        for (int r = 0; r < px * py * pz; ++r) {
            off[r] = o;
            o += (int64_t)cx[r % px] * cy[(r / px) % py] * cz[r / (px * py)];
        }
    }
    static int owner(const std::vector<int> &o, const std::vector<int> &c, int v)
    {
        for (size_t q = 0; q < o.size(); ++q) if (v < o[q] + c[q]) return (int)q;
        return (int)o.size() - 1;
    }
    int64_t node(int i, int j, int k) const
    {
        int a = owner(ox, cx, i), b = owner(oy, cy, j), c = owner(oz, cz, k);
        int r = a + b * px + c * px * py;
        return off[r] + (i - ox[a]) + (int64_t)(j - oy[b]) * cx[a] + (int64_t)(k - oz[c]) * cx[a] * cy[b];
    }
};

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

struct BcEntry {
    int i, j, k, d;     // GLOBAL node coordinates and dof
    double coef;        // value = coef * U  (bc_apply_on_u_*, bcs.c:61-146)
};

// bc_init_bending (bcs.c:198-251) / bc_init_circle (bcs.c:254-338) over the rank's GHOSTED box,
// in the reference's order; *nbcs is the reference's allocation size (the list is -1 padded to it).
inline void build_bc_entries(const macroc_config &cfg, const Slab &s, std::vector<BcEntry> &out, int *nbcs)
{
    const Geometry g = make_geometry(cfg);
    const int nxg = s.Xm, nyg = s.Ym, nzg = s.Zm, si = s.Xs, sj = s.Ys, sk = s.Zs;
    out.clear();
    if (cfg.bc_type == MACROC_BC_BENDING) {
        *nbcs = 2 * nyg * nzg * 3;
        for (int face = 0; face < 2; ++face) {
            bool on = face == 0 ? (si == 0) : (si + nxg == cfg.NX);
            if (!on) continue;
            int i = face == 0 ? 0 : nxg - 1;
            for (int k = 0; k < nzg; ++k)
                for (int j = 0; j < nyg; ++j)
                    for (int d = 0; d < 3; ++d) out.push_back({si + i, sj + j, sk + k, d, (face == 1 && d == 1) ? 1. : 0.});
        }
    } else {
        *nbcs = (2 * nxg + 2 * nzg) * 3 + nxg * nzg;
        if (si == 0 && sj == 0)                                   // X=0 & Y=0 along z
            for (int k = 0; k < nzg; ++k)
                for (int d = 0; d < 3; ++d) out.push_back({0, 0, sk + k, d, 0.});
        if (si + nxg == cfg.NX && sj == 0)                        // X=LX & Y=0 along z
            for (int k = 0; k < nzg; ++k)
                for (int d = 0; d < 3; ++d) out.push_back({cfg.NX - 1, 0, sk + k, d, 0.});
        if (sk == 0 && sj == 0)                                   // Z=0 & Y=0 along x
            for (int i = 1; i < nxg - 1; ++i)
                for (int d = 0; d < 3; ++d) out.push_back({si + i, 0, 0, d, 0.});
        if (sk + nzg == cfg.NZ && sj == 0)                        // Z=LZ & Y=0 along x
            for (int i = 1; i < nxg - 1; ++i)
                for (int d = 0; d < 3; ++d) out.push_back({si + i, 0, cfg.NZ - 1, d, 0.});
        if (sj + nyg == cfg.NY)                                   // circle on Y = LY, dof y only
            for (int i = 0; i < nxg; ++i)
                for (int k = 0; k < nzg; ++k)
                    if (in_circle(cfg, g, si + i, sk + k)) out.push_back({si + i, cfg.NY - 1, sk + k, 1, 1.});
    }
}

// index_dirichlet as the reference holds it: PETSc global dof ids, -1 padded to nbcs
inline void build_bc_lists(const macroc_config &cfg, const Slab &s, std::vector<int32_t> &idx, std::vector<double> &coef)
{
    std::vector<BcEntry> e;
    int nbcs = 0;
    build_bc_entries(cfg, s, e, &nbcs);
    GlobalNumbering gn(s);
    idx.clear(); coef.clear();
    for (const BcEntry &b : e) {
        idx.push_back((int32_t)(gn.node(b.i, b.j, b.k) * 3 + b.d));
        coef.push_back(b.coef);
    }
    idx.resize(nbcs, -1); coef.resize(nbcs, 0.);
}

}  // namespace macroc
