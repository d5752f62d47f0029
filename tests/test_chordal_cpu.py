"""CPU tests of the host-side chordal decomposition (fixture G15):
the decomposed and the undecomposed problem have the same optimum (Agler's theorem), cliques cover
the pattern, and reverse_decomposition sums the blocks back."""
import numpy as np
import pytest
import scipy.sparse as sp

import cosmo_b200
from cosmo_b200 import chordal
from oracle import cosmo_oracle as O


def _solve_oracle(P, q, A, b, sets, **kw):
    return O.solve(P, q, A, b, cosmo_b200.problems.to_oracle_cones(sets), O.Settings(**kw))


def test_svec_index_roundtrip():
    k = np.arange(0, 5000)
    i, j = chordal.svec_to_ij(k)
    assert np.all(i <= j) and np.array_equal(j * (j + 1) // 2 + i, k)


def test_cliques_are_a_chordal_cover():
    rows, cols, w = cosmo_b200.problems.banded_random_graph(300, 3.0, 8, seed=2)
    tree = chordal.chordal_cliques(300, rows, cols)
    cl = [set(c.tolist()) for c in tree.cliques]
    for a, b in zip(rows, cols):          # every edge lives in a clique
        assert any(a in c and b in c for c in cl)
    assert set().union(*cl) == set(range(300))
    for k, p in enumerate(tree.parent):   # running intersection: separator = clique ∩ parent
        if p >= 0:
            assert set(tree.sep[k].tolist()) == cl[k] & cl[p] and len(tree.sep[k]) < len(cl[k])
    assert sum(1 for p in tree.parent if p < 0) >= 1


def test_g15_small_maxcut_primal_equals_decomposed_dual():
    # examples/maxcut.jl:8-13 (n = 4, weights 1, 8, 2, 10, 6); documented cliques {1,2,4}, {2,3,4}
    rows = np.array([0, 0, 1, 1, 2]); cols = np.array([1, 3, 2, 3, 3]); w = np.array([1.0, 8, 2, 10, 6])
    P, q, A, b, sets = cosmo_b200.problems.maxcut_dual_sdp(4, rows, cols, w)
    P2, q2, A2, b2, sets2, info = chordal.decompose(P, q, A, b, sets, merge="none")
    assert sorted(sorted(c.tolist()) for _, c in info.blocks[0]) == [[0, 1, 3], [1, 2, 3]]
    ref = _solve_oracle(P, q, A, b, sets, eps_abs=1e-7, eps_rel=1e-7, scaling=0)
    dec = _solve_oracle(P2, q2, A2, b2, sets2, eps_abs=1e-7, eps_rel=1e-7, scaling=0)
    assert ref.status == dec.status == "Solved"
    assert abs(ref.obj_val - dec.obj_val) < 1e-4          # chordal_decomposition_triangle.jl:141-190 tolerance
    x, s, mu = chordal.reverse(info, dec.x, dec.s, -dec.y)
    assert np.allclose(x, ref.x, atol=1e-3)
    assert np.allclose(s, ref.s, atol=1e-3)               # S blocks sum back to the full slack
    # primal MAXCUT SDP optimum equals the dual optimum (strong duality): max 1/4 <L, Y>, Y_ii = 1
    L = np.zeros((4, 4))
    for a, c_, ww in zip(rows, cols, w):
        L[a, a] += ww; L[c_, c_] += ww; L[a, c_] -= ww; L[c_, a] -= ww
    assert abs(dec.obj_val - ref.obj_val) < 1e-4 and dec.obj_val > 0.25 * L.diagonal().sum() / 2


@pytest.mark.parametrize("merge", ["none", "parent_child"])
def test_random_banded_maxcut_decomposition(merge):
    nv = 40
    rows, cols, w = cosmo_b200.problems.banded_random_graph(nv, 3.0, 5, seed=4)
    P, q, A, b, sets = cosmo_b200.problems.maxcut_dual_sdp(nv, rows, cols, w)
    P2, q2, A2, b2, sets2, info = chordal.decompose(P, q, A, b, sets, merge=merge)
    assert len(sets2) > 3 and max(info.clique_sizes) < nv
    assert A2.shape == (sum(S.dim for S in sets2), nv + info.num_overlaps)
    ref = _solve_oracle(P, q, A, b, sets, eps_abs=1e-6, eps_rel=1e-6)
    dec = _solve_oracle(P2, q2, A2, b2, sets2, eps_abs=1e-6, eps_rel=1e-6)
    assert ref.status == dec.status == "Solved"
    assert abs(ref.obj_val - dec.obj_val) < 1e-3 * max(1, abs(ref.obj_val))
    x, s, mu = chordal.reverse(info, dec.x, dec.s, -dec.y)
    assert np.allclose(x, ref.x, atol=2e-3 * max(1, np.abs(ref.x).max()))
    # the reassembled slack is PSD and satisfies the original equality A x + s = b
    assert np.max(np.abs(A @ x + s - b)) < 1e-3
