"""CPU mirror of the algebra the node-centric element-Jacobian kernels rely on (macroc_b200/csrc/assembly_node.cuh)
against the oracle's element routine (reference loop src/assembly.c:94-99 with calc_B, :195-254):

    Ke[3a+d][3b+c] = wg * sum_gp sum_{p,q} h_a[p] * C_gp[voigt(d,p)][voigt(c,q)] * h_b[q],
    B[voigt(d,p)][3n+d] = h_n[p]  (and nothing else),

the slot a thread accumulates into, the half the symmetric layout keeps, and the k-permutation of the DMMA
contraction (csrc/dmma_ab.cuh).  The CUDA kernels themselves are checked on the GPU (tests/test_gpu_parity.py)."""
import numpy as np
import pytest

from oracle import oracle as O

PX = [0, 1, 1, 0, 0, 1, 1, 0]          # position of local node n inside its element (kernels.cuh node_px/py/pz)
PY = [0, 0, 1, 1, 0, 0, 1, 1]
PZ = [0, 0, 0, 0, 1, 1, 1, 1]


def voigt(i, j):
    return i if i == j else i + j + 2


def slot_of(a, b):
    return (PZ[b] - PZ[a] + 1) * 9 + (PY[b] - PY[a] + 1) * 3 + (PX[b] - PX[a] + 1)


def node_rank(n):
    return PX[n] + 2 * PY[n] + 4 * PZ[n]


def shape_derivatives():
    """h[gp][n][p] read off the oracle's B matrix: B[p][3n+p] = h_n[p]."""
    h = np.zeros((8, 8, 3))
    for gp in range(8):
        B = O.calc_B(gp)
        for n in range(8):
            for p in range(3):
                h[gp, n, p] = B[p, 3 * n + p]
    return h


def test_B_has_exactly_the_voigt_structure():
    h = shape_derivatives()
    for gp in range(8):
        B = O.calc_B(gp)
        R = np.zeros_like(B)
        for n in range(8):
            for d in range(3):
                for p in range(3):
                    R[voigt(d, p), 3 * n + d] = h[gp, n, p]
        assert np.array_equal(R, B)


@pytest.mark.parametrize("symmetric", [True, False])
def test_node_centric_formula_equals_the_reference_element_matrix(symmetric):
    rng = np.random.default_rng(7)
    h = shape_derivatives()
    Q = rng.standard_normal((8, 6, 6))
    C = Q @ Q.transpose(0, 2, 1) + 6 * np.eye(6) if symmetric else Q + 6 * np.eye(6)
    wg = 0.37
    ref = O.elem_jac(C.reshape(8, 36), wg).reshape(24, 24)
    Ke = np.zeros((24, 24))
    for d in range(3):
        for c in range(3):                                   # one (d, c) = one warp of the kernel
            sub = np.array([[[C[gp, voigt(d, p), voigt(c, q)] for q in range(3)] for p in range(3)] for gp in range(8)])
            for a in range(8):
                for gp in range(8):
                    T = h[gp, a] @ sub[gp]                   # T[q] = sum_p h_a[p] C[voigt(d,p)][voigt(c,q)]
                    for b in range(8):
                        Ke[3 * a + d, 3 * b + c] += T @ h[gp, b]
    Ke *= wg
    assert np.abs(Ke - ref).max() <= 1e-13 * np.abs(ref).max()


def test_slots_and_the_symmetric_half():
    """Thread (node, d, c) adds the pair (a, b) to slot_of(a, b): the 64 pairs of an element cover every slot offset of the
    octant exactly once, the mirrored slot is 26 - s, and slot >= 13 (what the symmetric layout stores) is b's rank >= a's."""
    for a in range(8):
        seen = set()
        for b in range(8):
            s = slot_of(a, b)
            assert 0 <= s < 27 and s not in seen
            seen.add(s)
            assert slot_of(b, a) == 26 - s
            assert (s >= 13) == (node_rank(b) >= node_rank(a))
    assert sum(1 for a in range(8) for b in range(8) if slot_of(a, b) >= 13) == 36


def test_dmma_k_permutation():
    """dmma_ab.cuh: with the k index of k-step s taken as 2t + s (t = lane % 4), the accumulator fragment of
    W^T = B^T C (columns 2t, 2t+1) is the A fragment of Ke += W^T B, and the fragments of B^T serve both products.
    Emulate mma.m8n8k4 with that operand placement and compare with B^T C B."""
    rng = np.random.default_rng(11)
    B = np.zeros((8, 24)); B[:6] = O.calc_B(3)               # rows 6, 7: zero padding of K = 6 -> 8
    C = np.zeros((8, 8)); C[:6, :6] = rng.standard_normal((6, 6))

    def mma(acc, afrag, bfrag):                              # afrag[gid][t] = A[gid][k_t], bfrag[t][gid] = B[k_t][gid]
        return acc + afrag @ bfrag

    Ke = np.zeros((24, 24))
    for i in range(3):
        Wt = np.zeros((8, 8))                                # rows 8i..8i+7 of W^T, all 8 columns
        for s in range(2):
            ks = [2 * t + s for t in range(4)]
            Wt = mma(Wt, B[ks][:, 8 * i:8 * i + 8].T, C[ks])
        for j in range(3):
            acc = np.zeros((8, 8))
            for s in range(2):
                ks = [2 * t + s for t in range(4)]
                acc = mma(acc, Wt[:, ks], B[ks][:, 8 * j:8 * j + 8])   # A fragment = the accumulator's own columns 2t + s
            Ke[8 * i:8 * i + 8, 8 * j:8 * j + 8] = acc
    ref = B[:6].T @ C[:6, :6] @ B[:6]
    assert np.abs(Ke - ref).max() <= 1e-13 * np.abs(ref).max()
