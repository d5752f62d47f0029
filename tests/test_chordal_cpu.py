"""CPU tests of the host-side chordal decomposition (fixture G15):
the decomposed and the undecomposed problem have the same optimum (Agler's theorem), cliques cover
the pattern, and reverse_decomposition sums the blocks back."""
import numpy as np
import pytest
import scipy.sparse as sp

import cosmo_b200
from cosmo_b200 import chordal
from oracle import cosmo_oracle as O
from oracle.bridge import to_oracle_cones


def _solve_oracle(P, q, A, b, sets, **kw):
    return O.solve(P, q, A, b, to_oracle_cones(sets), O.Settings(**kw))


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


def test_psd_completion_of_a_banded_matrix():
    # psd_complete! (chordal_decomposition.jl:262-311): keep the entries of a positive definite matrix on a
    # chordal pattern, complete the rest -> PSD, pattern entries untouched, inverse has the pattern's zeros
    # (the maximum-determinant completion), and a matrix that is already "its own completion" is reproduced
    rng = np.random.default_rng(3)
    nv = 30
    rows, cols, _ = cosmo_b200.problems.banded_random_graph(nv, 3.0, 4, seed=5)
    for merge in ("none", "parent_child"):
        tree = chordal.chordal_cliques(nv, rows, cols)
        if merge == "parent_child":
            tree = chordal.parent_child_merge(tree)
        B = rng.standard_normal((nv, nv))
        X = B @ B.T + nv * np.eye(nv)
        mask = np.zeros((nv, nv), dtype=bool)
        for c in tree.cliques:
            mask[np.ix_(c, c)] = True
        Y = chordal.psd_complete(np.where(mask, X, 0.0), tree)
        assert np.allclose(Y, Y.T) and np.allclose(Y[mask], X[mask])
        assert np.linalg.eigvalsh(Y).min() > 1e-8
        Yinv = np.linalg.inv(Y)
        assert np.max(np.abs(Yinv[~mask])) < 1e-9 * np.abs(Yinv).max()      # zeros of the inverse off the pattern
        Y2 = chordal.psd_complete(np.where(mask, Y, 0.0), tree)             # idempotent
        assert np.allclose(Y2, Y, atol=1e-9)


def test_g15_maxcut_completed_dual_recovers_the_primal_matrix():
    # examples/maxcut.jl:73-83: the dual variable of the PSD constraint of the dual MAXCUT SDP is the primal
    # matrix Y (Y_ii = 1, Y PSD, 1/4 <L, Y> = optimum); after the decomposition only its clique entries are
    # determined and `complete_dual` fills the rest
    rows = np.array([0, 0, 1, 1, 2]); cols = np.array([1, 3, 2, 3, 3]); w = np.array([1.0, 8, 2, 10, 6])
    P, q, A, b, sets = cosmo_b200.problems.maxcut_dual_sdp(4, rows, cols, w)
    P2, q2, A2, b2, sets2, info = chordal.decompose(P, q, A, b, sets, merge="none")
    dec = _solve_oracle(P2, q2, A2, b2, sets2, eps_abs=1e-8, eps_rel=1e-8, scaling=0)
    x, s, mu = chordal.reverse(info, dec.x, dec.s, -dec.y, complete_dual=True)
    x0, s0, mu0 = chordal.reverse(info, dec.x, dec.s, -dec.y)
    Y = chordal._svec_to_mat(-mu, 4)
    Y0 = chordal._svec_to_mat(-mu0, 4)
    assert Y0[0, 2] == 0.0 and abs(Y[0, 2]) > 1e-3                  # (1,3) is the only entry outside the cliques
    keep = np.ones((4, 4), dtype=bool); keep[0, 2] = keep[2, 0] = False
    assert np.allclose(Y[keep], Y0[keep])
    assert np.linalg.eigvalsh(Y).min() > -1e-6 and np.allclose(np.diag(Y), 1.0, atol=1e-4)
    L = np.zeros((4, 4))
    for a, c_, ww in zip(rows, cols, w):
        L[a, a] += ww; L[c_, c_] += ww; L[a, c_] -= ww; L[c_, a] -= ww
    assert abs(0.25 * np.sum(L * Y) - dec.obj_val) < 1e-3
    # the undecomposed solve returns a PSD dual matrix with the same clique entries
    ref = _solve_oracle(P, q, A, b, sets, eps_abs=1e-8, eps_rel=1e-8, scaling=0)
    Yref = chordal._svec_to_mat(ref.y, 4)
    assert np.allclose(Yref[keep], Y[keep], atol=1e-3)


def _mine(cons):
    out = []
    for c in cons:
        S = c.convex_set
        out.append((c.A, c.b, getattr(cosmo_b200, type(S).__name__)(S.dim)))
    return out


def _assemble_mine(builder):
    P, q, cons = builder()
    model = cosmo_b200.Model()
    cosmo_b200.assemble(model, P, q, [cosmo_b200.Constraint(A, b, S) for A, b, S in _mine(cons)], cosmo_b200.Settings())
    return model.P0, model.q0, model.A0, model.b0, model.sets0


def test_g6_documented_cliques_and_decomposed_optimum():
    from tests import golden_problems as G
    P, q, A, b, sets = _assemble_mine(G.g6_chordal_sdp)
    P2, q2, A2, b2, sets2, info = chordal.decompose(P, q, A, b, sets, merge="none")
    # docs/src/decomposition.md:43: {1,3,6}, {2,3}, {3,6,7,8}, {4,5,8}, {6,7,8,9}
    assert sorted(sorted(c.tolist()) for _, c in info.blocks[0]) == sorted(G.G6_CLIQUES)
    ref = _solve_oracle(P, q, A, b, sets, eps_abs=1e-7, eps_rel=1e-7)
    dec = _solve_oracle(P2, q2, A2, b2, sets2, eps_abs=1e-7, eps_rel=1e-7)
    assert ref.status == dec.status == "Solved" and abs(ref.obj_val - dec.obj_val) < 1e-4
    x, s, mu = chordal.reverse(info, dec.x, dec.s, -dec.y, complete_dual=True)
    assert np.allclose(x, ref.x, atol=1e-3)
    A1, A2m, B, c = G.g6_chordal_sdp_data()
    S = B - A1 * x[0] - A2m * x[1]
    assert np.linalg.eigvalsh(S).min() > -1e-4                      # primal feasible
    Y = chordal._svec_to_mat(-mu, 9)                                 # completed dual (examples/...: complete_dual = true)
    assert np.linalg.eigvalsh(Y).min() > -1e-4
    assert abs(np.sum(S * Y)) < 1e-2 * max(1.0, np.abs(S).max() * np.abs(Y).max())   # complementary slackness
    # stationarity of  min c'x  s.t.  B - A1 x1 - A2 x2 = S >= 0 :  c_i + <A_i, Y> = 0
    assert abs(np.sum(A1 * Y) + c[0]) < 1e-3 and abs(np.sum(A2m * Y) + c[1]) < 1e-3


def test_g5_sigma_max_lmi_decomposed():
    # nuclear_norm_minimization.jl:16-41 runs with decompose = true, compact_transformation = true
    from tests import golden_problems as G
    P, q, A, b, sets = _assemble_mine(G.g5_sigma_max_lmi)
    P2, q2, A2, b2, sets2, info = chordal.decompose(P, q, A, b, sets, merge="none")
    assert len(sets2) > len(sets)                                    # the 6x6 arrow pattern does decompose
    dec = _solve_oracle(P2, q2, A2, b2, sets2)
    assert dec.status == "Solved"
    x, s, mu = chordal.reverse(info, dec.x, dec.s, -dec.y)
    Y = x[1:].reshape(3, 3, order="F")
    assert Y[1, 0] <= 4 + 1e-6 and Y[1, 1] >= 3 - 1e-6 and Y.sum() - 12.0 >= -1e-3
    assert abs(np.linalg.svd(Y, compute_uv=False).max() - x[0]) <= 1e-3


# ---------------------------------------------------------------------------
# clique merging: the reference's literal fixtures, 1-based there, 0-based here
# ---------------------------------------------------------------------------
def _example_tree():
    """test/UnitTests/DecompositionTests/clique_merging_example.jl:6-24"""
    snd = [{15, 16, 17}, {5, 9}, {3, 4}, {1}, {2}, {6}, {7, 8}, {12, 13, 14}, {10, 11}]
    sep = [set(), {15, 16}, {5, 15}, {3}, {3, 4}, {9, 16}, {9, 15}, {16, 17}, {13, 14, 17}]
    parents = [0, 1, 2, 3, 3, 2, 2, 1, 8]
    cliques = [np.array(sorted(a | b)) for a, b in zip(snd, sep)]
    tree = chordal.CliqueTree(cliques, [p - 1 for p in parents], [np.array(sorted(x), dtype=np.int64) for x in sep],
                              np.arange(1, 18))
    return tree, snd, sep


def test_g10_parent_child_merge_log():
    # clique_merging_example.jl:88-97: pairs [1 2; 1 3; 1 4; 1 5; 1 6; 1 7; 1 8; 8 9],
    # decisions [T T T T T F F T], 6 merges
    tree, snd, sep = _example_tree()
    out, pairs, dec = chordal.parent_child_merge_reference(tree, snd_post=list(range(8, -1, -1)), return_log=True)
    assert [(a + 1, b + 1) for a, b in pairs] == [(1, 2), (1, 3), (1, 4), (1, 5), (1, 6), (1, 7), (1, 8), (8, 9)]
    assert dec == [True, True, True, True, True, False, False, True] and sum(dec) == 6
    got = sorted(sorted(c.tolist()) for c in out.cliques)
    assert got == sorted([sorted({15, 16, 17, 5, 9, 3, 4, 1, 2, 6}), [7, 8, 9, 15], sorted({12, 13, 14, 16, 17, 10, 11})])
    # first merge (:72-84): fill-in 2, supernode size 3
    assert ((3 + 0) - 2) * ((2 + 2) - 2) == 2


def test_g10_clique_graph_of_the_example_tree():
    # clique_merging_example.jl:46-62, 102-121: edges of the reduced clique graph with their complexity weights;
    # all weights are negative -> no merge, and the recomputed clique tree has the original parents
    tree, snd, sep = _example_tree()
    g = chordal.CliqueGraph([set(c.tolist()) for c in tree.cliques], [set(x.tolist()) for x in tree.sep])
    rows = [2, 8, 3, 6, 7, 4, 5, 5, 9]
    cols = [1, 1, 2, 2, 2, 3, 3, 4, 8]
    ref = {}
    for r, c in zip(rows, cols):
        a, b = set(tree.cliques[r - 1].tolist()), set(tree.cliques[c - 1].tolist())
        ref[(r - 1, c - 1)] = float(len(a) ** 3 + len(b) ** 3 - len(a | b) ** 3)
    assert g.edges == ref and all(w < 0 for w in ref.values())
    g.run()
    assert sum(1 for e in g.log if e[2]) == 0
    out = g.clique_tree(tree.order)
    assert [p + 1 for p in out.parent] == [0, 1, 2, 3, 3, 2, 2, 1, 8]


def _habib_stacho():
    """reduced_clique_graph.jl:6-8 (Habib & Stacho 2011, Fig. 1)"""
    snd = [{4, 5}, {1, 4, 6}, {1, 7}, {1, 8}, {1, 3, 4}, {1, 2, 3}, {2, 3, 9}, {3, 4, 11}, {3, 10}]
    sep = [{1, 3}, {1, 4}, {2, 3}, {3, 4}, {1}, {3}, {4}]
    return snd, sep


def _valid_clique_tree(t):
    """running intersection property (check_clique_tree of the reference's test utilities)"""
    cl = [set(c.tolist()) for c in t.cliques]
    for k, p in enumerate(t.parent):
        if p < 0:
            continue
        anc, q = set(), p
        while q >= 0:
            anc |= cl[q]
            q = t.parent[q]
        if not (cl[k] & anc) <= cl[p]:
            return False
    return sum(1 for p in t.parent if p < 0) >= 1      # a forest when the pattern is disconnected


def test_g10_reduced_clique_graph_habib_stacho():
    # reduced_clique_graph.jl:10-43: edges, permissible edges
    snd, sep = _habib_stacho()
    rows, cols = chordal.reduced_clique_graph(snd, sep)
    edges_ref = [(2, 1), (5, 1), (8, 1), (9, 8), (9, 5), (9, 7), (7, 6), (6, 4), (5, 4), (4, 2), (4, 3), (3, 2), (5, 3),
                 (6, 3), (9, 6), (8, 5), (5, 2), (6, 5)]
    got = {(r + 1, c + 1) for r, c in zip(rows, cols)}
    assert got <= set(edges_ref) and got == set(edges_ref)
    g = chordal.CliqueGraph(snd, sep)
    permissible_ref = {edges_ref[i - 1] for i in (7, 11, 16, 17, 18)}
    for e in g.edges:
        if g.permissible(e):
            assert (e[0] + 1, e[1] + 1) in permissible_ref


def test_g10_merging_two_cliques_updates_the_graph():
    # reduced_clique_graph.jl:46-87
    snd, sep = _habib_stacho()
    g = chordal.CliqueGraph(snd, sep)
    assert g.edges[(4, 1)] < 0                                   # cand [5, 2]: evaluate is false ...
    g.merge((4, 1))                                              # ... the test merges it anyway
    assert g.snd[1] == set() and g.snd[4] == {1, 3, 4, 6}
    assert 1 not in g.adj and not any(1 in s for s in g.adj.values())
    g = chordal.CliqueGraph(snd, sep)
    assert g.edges[(6, 5)] < 0
    g.merge((6, 5))                                              # cand [7, 6]
    assert g.snd[5] == set() and g.snd[6] == {1, 2, 3, 9}
    assert 5 not in g.adj and not any(5 in s for s in g.adj.values())
    t = g.clique_tree(np.arange(1, 12))                          # recomputation step -> a valid clique tree
    assert len(t.cliques) == 8 and _valid_clique_tree(t)


@pytest.mark.parametrize("merge", ["clique_graph", "parent_child_reference"])
def test_merge_strategies_give_valid_decompositions(merge):
    nv = 40
    rows, cols, w = cosmo_b200.problems.banded_random_graph(nv, 3.0, 5, seed=4)
    tree0 = chordal.chordal_cliques(nv, rows, cols)
    tree = chordal.clique_graph_merge(tree0) if merge == "clique_graph" else chordal.parent_child_merge_reference(tree0)
    assert len(tree.cliques) <= len(tree0.cliques) and _valid_clique_tree(tree)
    cl0 = [set(c.tolist()) for c in tree0.cliques]
    cl = [set(c.tolist()) for c in tree.cliques]
    assert all(any(c <= d for d in cl) for c in cl0)             # merging only enlarges cliques
    if merge == "clique_graph":                                  # every merge saved projection work
        assert sum(len(c) ** 3 for c in cl) <= sum(len(c) ** 3 for c in cl0)
    P, q, A, b, sets = cosmo_b200.problems.maxcut_dual_sdp(nv, rows, cols, w)
    P2, q2, A2, b2, sets2, info = chordal.decompose(P, q, A, b, sets, merge=merge)
    ref = _solve_oracle(P, q, A, b, sets, eps_abs=1e-6, eps_rel=1e-6)
    dec = _solve_oracle(P2, q2, A2, b2, sets2, eps_abs=1e-6, eps_rel=1e-6)
    assert ref.status == dec.status == "Solved"
    assert abs(ref.obj_val - dec.obj_val) < 1e-3 * max(1, abs(ref.obj_val))
    x, s, mu = chordal.reverse(info, dec.x, dec.s, -dec.y, complete_dual=True)
    assert np.allclose(x, ref.x, atol=2e-3 * max(1, np.abs(ref.x).max()))
    assert np.linalg.eigvalsh(chordal._svec_to_mat(-mu, nv)).min() > -1e-3


@pytest.mark.parametrize("nv,deg,band,seed", [(60, 3.0, 6, 1), (200, 3.0, 20, 2), (150, 5.0, 12, 3), (120, 4.0, 30, 4), (80, 2.0, 40, 4)])
def test_clique_graph_merge_incremental_order_equals_the_literal_restatement(nv, deg, band, seed):
    """CliqueGraph keeps its candidate order (weight descending, ties in CSC order) incrementally; the literal
    restatement of traverse (clique_merging.jl:242-259) sorts all edges at every step.  Both must pick the same edge at
    every step of the merge, and the bookkeeping (edges, adjacency) must stay consistent."""
    rows, cols, _ = cosmo_b200.problems.banded_random_graph(nv, deg, band, seed=seed)
    tree0 = chordal.chordal_cliques(nv, rows, cols)
    g = chordal.CliqueGraph([set(c.tolist()) for c in tree0.cliques], [set(x.tolist()) for x in tree0.sep])
    steps = 0
    while g.num > 1:
        cand = g.traverse()
        assert cand == g.traverse_by_sorting()
        assert sorted((-w, e[1], e[0]) for e, w in g.edges.items()) == list(g._ranked)
        assert all(e[0] in g.adj[e[1]] and e[1] in g.adj[e[0]] for e in g.edges)
        if cand is None or g.edges[cand] < 0:
            break
        g.merge(cand)
        steps += 1
        assert cand[1] not in g.adj and not any(cand[1] in st for st in g.adj.values())
        assert not any(cand[1] in e for e in g.edges)
    assert steps > 0 or deg <= 2.0        # a very sparse pattern may offer no merge that saves work
    assert _valid_clique_tree(g.clique_tree(tree0.order))


@pytest.mark.parametrize("nv,deg,band,seed", [(40, 3.0, 5, 4), (200, 3.0, 20, 2), (150, 5.0, 12, 3)])
@pytest.mark.parametrize("merge", ["none", "parent_child", "clique_graph"])
def test_psd_complete_renumbered_equals_the_literal_restatement(nv, deg, band, seed, merge):
    """psd_complete works on a renumbered matrix (visited vertices = a leading block); the literal restatement of
    psd_complete! (chordal_decomposition.jl:262-311) gathers index sets per clique.  Same completion, known entries
    bit-identical, result positive definite."""
    rows, cols, _ = cosmo_b200.problems.banded_random_graph(nv, deg, band, seed=seed)
    tree = chordal.chordal_cliques(nv, rows, cols)
    if merge == "parent_child":
        tree = chordal.parent_child_merge(tree)
    elif merge == "clique_graph":
        tree = chordal.clique_graph_merge(tree)
    rng = np.random.default_rng(seed)
    B = rng.standard_normal((nv, nv))
    X = B @ B.T + nv * np.eye(nv)
    mask = np.zeros((nv, nv), dtype=bool)
    for c in tree.cliques:
        mask[np.ix_(c, c)] = True
    Y0 = np.where(mask, X, 0.0)
    got = chordal.psd_complete(Y0, tree)
    ref = chordal._psd_complete_reference(Y0, tree)
    assert np.array_equal(got[mask], X[mask])
    assert np.max(np.abs(got - ref)) <= 1e-12 * np.max(np.abs(ref))
    assert np.array_equal(got, got.T) and np.linalg.eigvalsh(got).min() > 0
