"""CPU mirror of the index logic of the matrix-free kernels (macroc_b200/csrc/mf_march.cuh): the face-node enumeration of
k_apply_mf_faces covers every node on a face of the box exactly once and nothing else (the marching / patch kernels skip
exactly those), for slabs anywhere in the global grid; the z-marching pipeline (plane s contributes to the outputs s-1, s,
s+1; ring of four slots; segments) reproduces a plain 27-point stencil application.  The CUDA kernels themselves are
checked on the GPU against the assembled operator (tests/test_gpu_parity.py)."""
import numpy as np
import pytest


def face_count(NX, NY, NZ, zs, k0, k1):
    nzp = k1 - k0
    nx2 = NX - 2 if NX > 2 else 0
    ny2 = NY - 2 if NY > 2 else 0
    zf = (1 if k0 + zs <= 0 < k1 + zs else 0) + (1 if NZ > 1 and k0 + zs <= NZ - 1 < k1 + zs else 0)
    return (2 if NX > 1 else 1) * NY * nzp + (2 if NY > 1 else 1) * nx2 * nzp + zf * nx2 * ny2


def face_node(q, NX, NY, NZ, zs, k0, k1):
    """k_apply_mf_faces: q -> (i, j, kl)"""
    nzp = k1 - k0
    nx2 = NX - 2 if NX > 2 else 0
    ny2 = NY - 2 if NY > 2 else 0
    sxn = 2 if NX > 1 else 1
    syn = 2 if NY > 1 else 1
    nA, nB = sxn * NY * nzp, syn * nx2 * nzp
    zlow = k0 + zs <= 0 < k1 + zs
    if q < nA:
        per = NY * nzp
        side, r = divmod(q, per)
        return (NX - 1 if side else 0, r % NY, k0 + r // NY)
    if q < nA + nB:
        per = nx2 * nzp
        side, r = divmod(q - nA, per)
        return (1 + r % nx2, NY - 1 if side else 0, k0 + r // nx2)
    per = nx2 * ny2
    side, r = divmod(q - nA - nB, per)
    kl = -zs if (side == 0 and zlow) else NZ - 1 - zs
    return (1 + r % nx2, 1 + r // nx2, kl)


@pytest.mark.parametrize("NX,NY,NZ,zs,k0,k1", [
    (4, 4, 2, 0, 0, 2), (33, 5, 4, 0, 0, 4), (7, 3, 9, 0, 0, 9), (2, 2, 2, 0, 0, 2),
    (9, 6, 12, 0, 0, 4), (9, 6, 12, 4, 0, 4), (9, 6, 12, 8, 0, 4),          # three z-slabs of a 12-plane grid
    (9, 6, 12, 4, 1, 3), (9, 6, 12, 8, 3, 4), (9, 6, 12, 0, 0, 1),          # interior / boundary plane ranges of a split apply
    (5, 4, 3, 0, 0, 3), (3, 3, 3, 0, 1, 2)])
def test_face_enumeration_is_a_partition_of_the_face_nodes(NX, NY, NZ, zs, k0, k1):
    want = {(i, j, kl) for kl in range(k0, k1) for j in range(NY) for i in range(NX)
            if i in (0, NX - 1) or j in (0, NY - 1) or kl + zs in (0, NZ - 1)}
    n = face_count(NX, NY, NZ, zs, k0, k1)
    got = [face_node(q, NX, NY, NZ, zs, k0, k1) for q in range(n)]
    assert len(set(got)) == len(got), "a face node is visited twice"
    assert set(got) == want


def test_z_marching_pipeline_equals_a_plain_stencil():
    """Output-stationary marching over segments: step s adds plane s to the outputs s-1, s, s+1; an output plane leaves
    after the step of the plane above it; slot of plane p = (p + 4) % 4 holds it until plane p+1's output has used it."""
    rng = np.random.default_rng(2)
    NX, NY, NZ = 6, 5, 11
    x = rng.standard_normal((NZ + 2, NY + 2, NX + 2))          # one zero-padded ghost layer all around
    x[0] = x[-1] = 0; x[:, 0] = x[:, -1] = 0; x[:, :, 0] = x[:, :, -1] = 0
    T = rng.standard_normal((3, 3, 3))                         # T[dz+1][dy+1][dx+1]
    ref = np.zeros((NZ, NY, NX))
    for dz in range(3):
        for dy in range(3):
            for dx in range(3):
                ref += T[dz, dy, dx] * x[dz:dz + NZ, dy:dy + NY, dx:dx + NX]
    for nseg in (1, 2, 3, 11):
        out = np.full((NZ, NY, NX), np.nan)
        seglen = -(-NZ // nseg)
        for seg in range(nseg):
            zlo, zhi = seg * seglen, min(NZ, (seg + 1) * seglen)
            if zlo >= zhi:
                continue
            slots = {}
            am = np.zeros((NY, NX)); ac = np.zeros((NY, NX)); ap = np.zeros((NY, NX))
            slots[(zlo - 1 + 4) % 4] = zlo - 1
            slots[(zlo + 4) % 4] = zlo
            for s in range(zlo - 1, zhi + 1):
                if s + 2 <= zhi:
                    slots[(s + 6) % 4] = s + 2                  # staged two planes ahead into the slot of plane s-2
                assert slots[(s + 4) % 4] == s                  # the plane being used is resident ...
                if s - 1 >= zlo:
                    assert slots[(s + 3) % 4] == s - 1          # ... and so is the one the output still reads
                xs = x[s + 1]                                   # plane s (x carries one ghost plane below)
                for dy in range(3):
                    for dx in range(3):
                        v = xs[dy:dy + NY, dx:dx + NX]
                        ap += T[0, dy, dx] * v                  # output s+1 sees plane s through dz = -1
                        ac += T[1, dy, dx] * v
                        am += T[2, dy, dx] * v                  # output s-1 sees plane s through dz = +1
                if s - 1 >= zlo:
                    out[s - 1] = am
                am, ac, ap = ac, ap, np.zeros((NY, NX))
        assert np.allclose(out, ref, rtol=1e-13, atol=1e-13)
