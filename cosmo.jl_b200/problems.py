"""Synthetic problem generators for the BASELINE.json configs.

All data come from ``numpy.random.default_rng(seed)``; the same arrays are fed
to the CUDA engine and to the CPU oracle.  Problems are returned in COSMO's
model form  min 1/2 x'Px + q'x  s.t.  A x + s = b, s in K  as
``(P, q, A, b, sets)`` with ``sets`` a list of cosmo_b200 set objects.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp

from . import model as M


def _stratified_csr(rng, nrows, ncols, per_row):
    """Random sparse matrix with exactly `per_row` entries per row: one uniformly
    placed column in each of `per_row` equal strata (sorted, duplicate-free),
    N(0,1) values.  Same density / gather pattern as sprandn(m, n, per_row/n)."""
    per_row = int(min(per_row, ncols))
    edges = np.linspace(0, ncols, per_row + 1).astype(np.int64)
    width = np.diff(edges)
    cols = edges[:-1][None, :] + (rng.random((nrows, per_row)) * width[None, :]).astype(np.int64)
    vals = rng.standard_normal((nrows, per_row))
    indptr = np.arange(nrows + 1, dtype=np.int64) * per_row
    return sp.csr_matrix((vals.ravel(), cols.ravel().astype(np.int32), indptr), shape=(nrows, ncols))


def _stratified_csc(rng, nrows, ncols, per_col):
    """Transpose-layout twin of `_stratified_csr`: exactly `per_col` entries per column
    (one per row stratum), returned directly as CSC so that no format conversion of the
    5e7-entry matrix is needed (CSC is what COSMO / the C ABI ingest)."""
    return _stratified_csr(rng, ncols, nrows, per_col).T.tocsc(copy=False)


def random_sparse_qp(n=50_000, m=100_000, density=0.01, seed=2, p_nnz_per_row=10):
    """BASELINE config C2: random sparse QP, Nonnegatives(m/2) + Box(m/2).

    A = sprandn-like (m x n, `density`: m*density entries per column at stratified random rows,
    N(0,1) values), P = sparse diagonally dominant symmetric (~p_nnz_per_row per row, PSD by
    construction), strictly feasible around a random x0.
    """
    rng = np.random.default_rng(seed)
    per_col = max(1, int(round(density * m)))
    A = _stratified_csc(rng, m, n, per_col)
    k = max(1, p_nnz_per_row // 2)
    B = _stratified_csr(rng, n, n, k) * 0.1
    S = (B + B.T).tocsr()
    rowabs = np.asarray(abs(S).sum(axis=1)).ravel()
    P = (S + sp.diags(rowabs + rng.uniform(0.1, 1.0, n))).tocsc()
    x0 = rng.standard_normal(n)
    Ax0 = A @ x0
    m1 = m // 2
    # rows [0, m1): Nonnegatives, s = b - A x >= 0 with slack U(0,1) at x0
    b = np.empty(m)
    b[:m1] = Ax0[:m1] + rng.uniform(0.0, 1.0, m1)
    # rows [m1, m): Box, model form A x + s = 0  =>  s = -A x in [l, u]
    b[m1:] = 0.0
    l = -Ax0[m1:] - rng.uniform(0.0, 1.0, m - m1)
    u = -Ax0[m1:] + rng.uniform(0.0, 1.0, m - m1)
    q = -(P @ x0) + rng.standard_normal(n)
    sets = [M.Nonnegatives(m1), M.Box(l, u)]
    return P, q, A, b, sets


def portfolio_socp(n=20_000, k=2_000, seed=1, gamma=1.0):
    """BASELINE config C3: portfolio SOCP (examples/portfolio_optimisation.jl:64-70 scaled):
    min -mu'x  s.t.  |M'x| <= gamma, 1'x = 1, x >= 0, M' = [D^0.5; F'],
    rows ordered Zero(1), Nonneg(n), SOC(1+n+k) as sort_sets would."""
    rng = np.random.default_rng(seed)
    Ddiag = rng.random(n) * np.sqrt(k)
    F = sp.random(n, k, density=0.5, random_state=rng, data_rvs=rng.standard_normal, format="csc")
    mu = (3.0 + 9.0 * rng.random(n)) / 100.0
    Mt = sp.vstack([sp.diags(np.sqrt(Ddiag)), F.T], format="csr")
    # constraints in user form A_c x + b_c in K; model form uses A = -A_c, b = b_c
    A_zero = sp.csr_matrix(np.ones((1, n)))
    b_zero = np.array([-1.0])          # sum(x) - 1 == 0
    A_nn = sp.identity(n, format="csr")
    b_nn = np.zeros(n)
    A_soc = sp.vstack([sp.csr_matrix((1, n)), Mt], format="csr")
    b_soc = np.concatenate([[gamma], np.zeros(n + k)])
    A = sp.vstack([-A_zero, -A_nn, -A_soc], format="csc")
    b = np.concatenate([b_zero, b_nn, b_soc])
    P = sp.csc_matrix((n, n))
    q = -mu
    sets = [M.ZeroSet(1), M.Nonnegatives(n), M.SecondOrderCone(1 + n + k)]
    return P, q, A, b, sets


def svec_index(i, j):
    return j * (j + 1) // 2 + i


def closest_correlation_sdp(N=2000, seed=12345):
    """BASELINE config C4: closest correlation matrix (examples/closest_correlation_matrix.jl:21-33),
    svec/triangle form: min 1/2|X - C|_F^2, X_ii = 1, X PSD, x = svec(X)."""
    rng = np.random.default_rng(seed)
    C = -1.0 + 2.0 * rng.random((N, N))
    Cs = (C + C.T) / 2.0
    d = N * (N + 1) // 2
    iu = np.triu_indices(N)
    order = np.lexsort((iu[0], iu[1]))
    r, c = iu[0][order], iu[1][order]
    svecC = np.where(r == c, Cs[r, c], np.sqrt(2.0) * Cs[r, c])
    # 1/2 |X - C|_F^2 = 1/2 x'x - svec(C)'x + const
    P = sp.identity(d, format="csc")
    q = -svecC
    diag_pos = np.array([svec_index(i, i) for i in range(N)])
    A_zero = sp.csr_matrix((np.ones(N), (np.arange(N), diag_pos)), shape=(N, d))
    b_zero = -np.ones(N)
    A = sp.vstack([-A_zero, -sp.identity(d, format="csr")], format="csc")
    b = np.concatenate([b_zero, np.zeros(d)])
    sets = [M.ZeroSet(N), M.PsdConeTriangle(d)]
    return P, q, A, b, sets


def banded_random_graph(nv=10_000, mean_degree=3.0, bandwidth=20, seed=1):
    """BASELINE config C5 graph: random sparse graph with mean degree ~3 whose edges join vertices at
    index distance <= `bandwidth` (bounded treewidth => the chordal extension has small cliques, the
    regime chordal decomposition targets: many small PSD blocks instead of one huge one).
    Integer weights U{1..10}.  Returns (rows, cols, weights) with rows < cols, duplicate-free."""
    rng = np.random.default_rng(seed)
    ne = int(round(nv * mean_degree / 2.0 * 1.15))
    i = rng.integers(0, nv, size=ne)
    j = i + rng.integers(1, bandwidth + 1, size=ne)
    keep = j < nv
    i, j = i[keep], j[keep]
    key = np.unique(i.astype(np.int64) * nv + j)
    key = key[: int(round(nv * mean_degree / 2.0))] if len(key) > nv * mean_degree / 2.0 else key
    i, j = key // nv, key % nv
    w = rng.integers(1, 11, size=len(i)).astype(np.float64)
    return i, j, w


def maxcut_dual_sdp(nv, rows, cols, weights):
    """Dual MAXCUT SDP (examples/maxcut.jl:73-77): min sum(gamma) s.t. diag(gamma) - L/4 = S, S PSD,
    as a PsdConeTriangle constraint on svec(S).  Model form A x + s = b with x = gamma."""
    W = sp.coo_matrix((weights, (rows, cols)), shape=(nv, nv))
    deg = np.asarray((W + W.T).sum(axis=1)).ravel()
    d = nv * (nv + 1) // 2
    diag_pos = np.arange(nv, dtype=np.int64) * (np.arange(nv, dtype=np.int64) + 1) // 2 + np.arange(nv)
    # constraint A_c x + b_c in K with A_c = selector of the diagonal, b_c = svec(-L/4)
    A_c = sp.csr_matrix((np.ones(nv), (diag_pos, np.arange(nv))), shape=(d, nv))
    b_c = np.zeros(d)
    b_c[diag_pos] = -deg / 4.0
    lo, hi = np.minimum(rows, cols).astype(np.int64), np.maximum(rows, cols).astype(np.int64)
    b_c[hi * (hi + 1) // 2 + lo] = np.sqrt(2.0) * weights / 4.0      # -(-w_ij)/4, off-diagonals scaled by sqrt 2
    P = sp.csc_matrix((nv, nv))
    q = np.ones(nv)
    return P, q, (-A_c).tocsc(), b_c, [M.PsdConeTriangle(d)]
