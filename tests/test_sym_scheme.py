"""The symmetric-storage scheme on several ranks, restated in numpy.

The CUDA path (`spmv_sym.cuh`) stores 14 of the 27 block slots per local node and claims that a rank
needs nothing from its neighbours to apply the transposed half:

* the blocks of the ghost plane below a slab (rows of the lower z neighbour) are a function of the
  node class and the Dirichlet masks only, so the rank keeps a private copy;
* ghost columns / rows of x / y neighbours are ordinary local nodes whose node class is taken from
  the *local* box -- wrong for their own rows, but exact for every block that points at an owned
  node, because such a block only sums elements adjacent to that owned node, all of them local.

This file checks those two claims against the oracle's assembled operator for every rank of
several processor grids, and -- independently of the hardware -- the band-sweep algorithm of
`k_spmv_sym` (scatter through lane shifts, row carry and the accumulator plane; gathers only across
band edges; scatter-only pass below each z segment).  It mirrors `k_stencil_table`,
`k_fill_operator_sym` and `k_spmv_sym` step for step with a configurable tile width (no GPU needed);
the kernels themselves are checked by the `gpu` tests.
"""
import numpy as np
import pytest

from oracle import oracle as O

POS = [(0, 0, 0), (1, 0, 0), (1, 1, 0), (0, 1, 0), (0, 0, 1), (1, 0, 1), (1, 1, 1), (0, 1, 1)]   # hex8 local nodes
LOCAL_OF = {p: a for a, p in enumerate(POS)}


def node_class(c, n):
    return 0 if c == 0 else (2 if c == n - 1 else 1)


def class_stencils(Ke):
    """T[27 classes][27 slots][3][3]: k_stencil_table (kernels.cuh), same element order."""
    T = np.zeros((27, 27, 3, 3))
    for typ in range(27):
        tx, ty, tz = typ % 3, (typ // 3) % 3, typ // 9
        for slot in range(27):
            ddx, ddy, ddz = slot % 3 - 1, (slot // 3) % 3 - 1, slot // 9 - 1
            for oz in (-1, 0):
                for oy in (-1, 0):
                    for ox in (-1, 0):
                        if (ox == -1 and tx == 0) or (ox == 0 and tx == 2): continue
                        if (oy == -1 and ty == 0) or (oy == 0 and ty == 2): continue
                        if (oz == -1 and tz == 0) or (oz == 0 and tz == 2): continue
                        b = (ddx - ox, ddy - oy, ddz - oz)
                        if b not in LOCAL_OF: continue
                        a = LOCAL_OF[(-ox, -oy, -oz)]
                        T[typ, slot] += Ke[3 * a:3 * a + 3, 3 * LOCAL_OF[b]:3 * LOCAL_OF[b] + 3]
    return T


def rank_apply_sym(o, cfg, rank, T, mask, x, R=2, nseg=2, TW=4):
    """w on the owned nodes of `rank` from the rank's own 14-slot storage (fill + band sweep)."""
    NX, NY, NZ = cfg.NX, cfg.NY, cfg.NZ
    xs, ys, zs, xm, ym, zm = o.corners(rank)
    Xs, Ys, Zs, Xm, Ym, Zm = o.ghost_corners(rank)
    lo = 1 if zs > 0 else 0                         # a ghost plane below is stored (dz = +1 slots only)
    LX, LY, LZ = Xm, Ym, zm + lo                    # stored local box: ghosted in x / y, owned planes (+ lower ghost)

    def gid(i, j, k):                               # local (i, j, k) -> global natural node, or -1 outside the grid
        gi, gj, gk = Xs + i, Ys + j, zs - lo + k
        return gi + NX * (gj + NY * gk) if 0 <= gi < NX and 0 <= gj < NY and 0 <= gk < NZ else -1

    # k_fill_operator_sym: upper half from the class of the node in the LOCAL box (x, y) / global grid (z)
    up = np.zeros((LZ, LY, LX, 14, 3, 3))
    for k in range(LZ):
        for j in range(LY):
            for i in range(LX):
                g = gid(i, j, k)
                typ = node_class(i, LX) + 3 * node_class(j, LY) + 9 * node_class(zs - lo + k, NZ)
                own = mask[g]
                for s in range(13, 27):
                    ddx, ddy, ddz = s % 3 - 1, (s // 3) % 3 - 1, s // 9 - 1
                    if k < lo and ddz != 1: continue            # ghost plane: only the blocks towards the slab
                    v = T[typ, s].copy()
                    # neighbour mask as the padded device array sees it: zero outside the local box
                    ii, jj, kk = i + ddx, j + ddy, k + ddz
                    gn = gid(ii, jj, kk) if 0 <= ii < LX and 0 <= jj < LY and 0 <= kk <= LZ else -1
                    nb = mask[gn] if gn >= 0 else np.zeros(3, bool)
                    for r in range(3):
                        for c in range(3):
                            if own[r] or nb[c]:
                                v[r, c] = 1.0 if (s == 13 and r == c) else 0.0
                    up[k, j, i, s - 13] = v
    X = x.reshape(NZ, NY, NX, 3)

    def xval(i, j, k):                              # p with ghosts (halo) and zero padding
        g = gid(i, j, k)
        return X.reshape(-1, 3)[g] if g >= 0 and 0 <= i < LX and 0 <= j < LY else np.zeros(3)

    W = band_sweep(up, xval, LX, LY, LZ, lo, R, nseg, TW)
    out = {}
    for k in range(lo, LZ):
        for j in range(ys - Ys, ys - Ys + ym):
            for i in range(xs - Xs, xs - Xs + xm):
                out[gid(i, j, k)] = W[k, j, i]
    return out


def band_sweep(up, xval, LX, LY, LZ, lo, R, nseg, TW):
    """k_spmv_sym, step for step: a "warp" of TW lanes owns a band of R rows of one x-tile and sweeps
    it upward in z; scatter contributions travel by lane shifts (x), a carry (next row), the running
    rows N[3] and the accumulator plane `acc` (next plane); what would cross the band is gathered by
    the receiving band from the neighbour's blocks; a segment starts with a scatter-only pass over the
    plane below it.  Planes are indexed as stored: k = 0 is the ghost plane when lo = 1."""
    W = np.full((LZ, LY, LX, 3), np.nan)
    rt = -(-LX // TW)
    nplanes = LZ - lo
    lseg = -(-nplanes // nseg)
    offs = {s: (s % 3 - 1, (s // 3) % 3 - 1, s // 9 - 1) for s in range(13, 27)}
    for seg in range(nseg):
        k0, k1 = lo + seg * lseg, min(LZ, lo + (seg + 1) * lseg)
        if k0 >= k1:
            continue
        pre = k0 - 1 >= 0
        for y0 in range(0, LY, R):
            rows = min(R, LY - y0)
            for xt in range(rt):
                acc = np.zeros((R, TW, 3))
                for k in range(k0 - 1 if pre else k0, k1):
                    scatter_only = k < k0
                    carry = np.zeros((TW, 3))
                    N = np.zeros((3, TW, 3))
                    for r in range(rows):
                        j = y0 + r
                        a = acc[r] + carry
                        if not scatter_only:
                            for lane in range(TW):
                                i = xt * TW + lane
                                if i >= LX:
                                    continue
                                for s in range(14, 27):
                                    ddx, ddy, ddz = offs[s]
                                    outside = not (0 <= lane - ddx < TW) or not (0 <= r - ddy < rows)
                                    ii, jj, kk = i - ddx, j - ddy, k - ddz
                                    if outside and 0 <= ii < LX and 0 <= jj < LY and kk >= 0:
                                        a[lane] += up[kk, jj, ii, s - 13].T @ xval(ii, jj, kk)
                        nc = np.zeros((TW, 3))
                        for s in range(13, 27):
                            ddx, ddy, ddz = offs[s]
                            t = np.zeros((TW, 3))
                            for lane in range(TW):
                                i = xt * TW + lane
                                blk = up[k, j, i, s - 13] if i < LX else np.zeros((3, 3))
                                a[lane] += blk @ xval(i + ddx, j + ddy, k + ddz)
                                t[lane] = blk.T @ xval(i, j, k)
                            if s == 13:
                                continue
                            sh = np.zeros((TW, 3))                     # the shuffle: lane -> lane + ddx, edges dropped
                            for lane in range(TW):
                                if 0 <= lane + ddx < TW:
                                    sh[lane + ddx] = t[lane]
                            if ddz == 0 and ddy == 0:
                                a += sh
                            elif ddz == 0:
                                nc += sh
                            else:
                                N[ddy + 1] += sh
                        carry = nc
                        if not scatter_only:
                            for lane in range(TW):
                                if xt * TW + lane < LX:
                                    assert np.isnan(W[k, j, xt * TW + lane]).all()     # every node written once
                                    W[k, j, xt * TW + lane] = a[lane]
                        if r >= 1:
                            acc[r - 1] = N[0]
                        N[0], N[1], N[2] = N[1].copy(), N[2].copy(), np.zeros((TW, 3))
                    acc[rows - 1] = N[0]
    return W


@pytest.mark.parametrize("grid,bc,nranks,pg", [
    ((6, 5, 7), 0, 2, (1, 1, 2)),        # z-slabs
    ((6, 5, 9), 1, 3, (1, 1, 3)),
    ((7, 6, 5), 0, 2, (2, 1, 1)),        # x split: ghost columns
    ((6, 7, 5), 1, 2, (1, 2, 1)),        # y split: ghost rows
    ((7, 6, 6), 0, 4, (2, 2, 1)),
    ((6, 6, 7), 1, 8, (2, 2, 2)),        # corners and edges
    ((6, 4, 5), 0, 1, (1, 1, 1)),
])
@pytest.mark.parametrize("R,nseg,TW", [(2, 2, 4), (1, 1, 32), (3, 3, 4), (16, 1, 3)])
def test_rank_local_symmetric_storage_reproduces_the_operator(grid, bc, nranks, pg, R, nseg, TW):
    NX, NY, NZ = grid
    cfg = O.Config(NX=NX, NY=NY, NZ=NZ, bc_type=bc, nranks=nranks, px=pg[0], py=pg[1], pz=pg[2], lx=4., ly=1., lz=4.,
                   faithful_ke=0)
    o = O.Oracle(cfg)
    o.set_strains(); o.homogenize(); o.assembly_jac()
    dx, dy, dz = 4. / (NX - 1), 1. / (NY - 1), 4. / (NZ - 1)
    Ke = O.elem_jac(np.tile(O.isotropic_D().reshape(-1), 8), dx * dy * dz / 8.)
    T = class_stencils(Ke)
    mask = o.dirichlet_mask_natural().reshape(-1, 3)
    x = np.sin(0.37 * np.arange(3 * NX * NY * NZ)) + 0.1
    y_ref = o.matmult(x).reshape(-1, 3)
    seen = 0
    for r in range(nranks):
        for g, w in rank_apply_sym(o, cfg, r, T, mask, x, R, nseg, TW).items():
            assert np.allclose(w, y_ref[g], rtol=0, atol=1e-13 * np.abs(y_ref).max()), (r, g, w, y_ref[g])
            seen += 1
    assert seen == NX * NY * NZ           # every node owned by exactly one rank
