"""Host-side chordal decomposition -- the producer of the clique batch.

Stays on the host like the reference's `src/chordal_decomposition/` (north-star);
what the GPU consumes is only the augmented `(P', q', A', b')`, the cone list with
one `PsdConeTriangle` per clique, and the row map used to sum the blocks back.

Restated (not ported) from the reference:
  * aggregate sparsity pattern of a PSD cone   chordal_decomposition.jl:100-115
  * chordal extension + elimination tree        trees.jl:634-642 (the reference calls QDLDL with an AMD
    ordering; neither is available here, so a minimum-degree elimination with explicit fill is used --
    any perfect elimination ordering gives a valid decomposition, only the clique set differs)
  * supernodes / cliques / clique tree          trees.jl:390-513 (Pothen-Sun rule: v joins a child's
    supernode iff |hadj(child)| = |hadj(v)| + 1)
  * merge strategy                              NoMerge; ParentChildMerge(t_fill, t_size) in two flavours --
    `parent_child` (bottom-up variant with the clique-size rule, the one the C5 measurements use) and
    `parent_child_reference` (the reference's top-down walk, clique_merging.jl:262-283, 641-648);
    `clique_graph` = CliqueGraphMerge, the reference default (reduced clique graph, complexity weights,
    permissible edges, clique tree rebuilt by Kruskal; clique_graph.jl, clique_merging.jl:34-67, 204-259, 305-600)
  * compact ("clique tree based") augmentation  transformations.jl:152-374: every clique becomes a
    PsdConeTriangle block; an entry (i,j) inside the separator of a clique gets a new variable with +1 in
    the clique's row and -1 in the parent's row of the same (i,j)
  * reverse_decomposition!                       chordal_decomposition.jl:129-213 (x truncated, s = sum of
    blocks, mu = block value)
  * psd_completion! / psd_complete!               chordal_decomposition.jl:215-311 (`complete_dual`): the entries
    of the dual matrix outside the pattern are chosen so that Y = -mat(mu) is positive semidefinite -- the
    maximum-determinant completion, clique by clique from the root of the clique tree
    (Vandenberghe & Andersen, Chordal Graphs and Semidefinite Optimization, alg. 10.2)
"""
from __future__ import annotations

import heapq
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import numpy as np
from sortedcontainers import SortedList
import scipy.sparse as sp

from . import model as M


def svec_index(i: int, j: int) -> int:
    """position of (i,j), i<=j (0-based), in the column-major upper triangle (convexset.jl:432-442)"""
    return j * (j + 1) // 2 + i


def svec_to_ij(k: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """inverse of svec_index (svec_to_mat, trees.jl:697-719)"""
    k = np.asarray(k, dtype=np.int64)
    j = ((np.sqrt(8.0 * k + 1.0) - 1.0) / 2.0).astype(np.int64)
    j = np.where((j + 1) * (j + 2) // 2 <= k, j + 1, j)
    j = np.where(j * (j + 1) // 2 > k, j - 1, j)
    return k - j * (j + 1) // 2, j


@dataclass
class CliqueTree:
    cliques: List[np.ndarray]          # sorted vertex lists
    parent: List[int]                  # -1 for roots
    sep: List[np.ndarray]              # clique ∩ parent clique (sorted)
    order: np.ndarray                  # elimination order used


def chordal_cliques(nv: int, rows: np.ndarray, cols: np.ndarray) -> CliqueTree:
    """Minimum-degree elimination with explicit fill -> maximal cliques and their tree."""
    adj: List[set] = [set() for _ in range(nv)]
    for a, b in zip(rows.tolist(), cols.tolist()):
        if a != b:
            adj[a].add(b)
            adj[b].add(a)
    heap = [(len(adj[v]), v) for v in range(nv)]
    heapq.heapify(heap)
    done = np.zeros(nv, dtype=bool)
    pos = np.empty(nv, dtype=np.int64)
    order: List[int] = []
    hadj: List[Optional[List[int]]] = [None] * nv
    while heap:
        d, v = heapq.heappop(heap)
        if done[v] or d != len(adj[v]):
            continue
        done[v] = True
        pos[v] = len(order)
        order.append(v)
        nb = list(adj[v])
        hadj[v] = nb
        for a in nb:
            adj[a].discard(v)
        for ia in range(len(nb)):      # fill: the higher neighbourhood becomes a clique
            a = nb[ia]
            sa = adj[a]
            for ib in range(ia + 1, len(nb)):
                b = nb[ib]
                if b not in sa:
                    sa.add(b)
                    adj[b].add(a)
        for a in nb:
            heapq.heappush(heap, (len(adj[a]), a))
        adj[v] = set()
    # elimination tree: parent = earliest-eliminated higher neighbour
    par = np.full(nv, -1, dtype=np.int64)
    for v in range(nv):
        if hadj[v]:
            par[v] = min(hadj[v], key=lambda u: pos[u])
    # Pothen-Sun supernodes
    absorbing_child = np.full(nv, -1, dtype=np.int64)
    for u in order:
        p = par[u]
        if p >= 0 and absorbing_child[p] < 0 and len(hadj[u]) == len(hadj[p]) + 1:
            absorbing_child[p] = u
    snode = np.full(nv, -1, dtype=np.int64)
    lowest: List[int] = []
    top: List[int] = []
    for v in order:
        c = absorbing_child[v]
        if c >= 0:
            snode[v] = snode[c]
            top[snode[v]] = v
        else:
            snode[v] = len(lowest)
            lowest.append(v)
            top.append(v)
    cliques, parent, seps = [], [], []
    for k, v in enumerate(lowest):
        cl = np.array(sorted([v] + list(hadj[v])), dtype=np.int64)
        cliques.append(cl)
        t = top[k]
        p = par[t]
        parent.append(int(snode[p]) if p >= 0 else -1)
    for k, cl in enumerate(cliques):
        if parent[k] < 0:
            seps.append(np.zeros(0, dtype=np.int64))
        else:
            seps.append(np.intersect1d(cl, cliques[parent[k]]))
    return CliqueTree(cliques, parent, seps, np.array(order, dtype=np.int64))


def parent_child_merge(tree: CliqueTree, t_fill: int = 8, t_size: int = 8) -> CliqueTree:
    """ParentChildMerge (clique_merging.jl:278-285, 641-648): children are merged into their parent
    when the fill-in (|C_par|-|sep|)(|C|-|sep|) <= t_fill or both supernodes are small."""
    n = len(tree.cliques)
    cl = [set(c.tolist()) for c in tree.cliques]
    parent = list(tree.parent)
    alive = [True] * n
    children: List[List[int]] = [[] for _ in range(n)]
    for k, p in enumerate(parent):
        if p >= 0:
            children[p].append(k)
    # visit children before parents
    depth = [0] * n
    for k in range(n):
        d, p = 0, parent[k]
        while p >= 0:
            d += 1
            p = parent[p]
        depth[k] = d
    for k in sorted(range(n), key=lambda i: -depth[i]):
        p = parent[k]
        if p < 0 or not alive[k]:
            continue
        while not alive[p]:
            p = parent[p]
        sep = len(cl[k] & cl[p])
        fill = (len(cl[p]) - sep) * (len(cl[k]) - sep)
        if fill <= t_fill or max(len(cl[k]) - sep, len(cl[p]) - sep) <= t_size:
            cl[p] |= cl[k]
            alive[k] = False
            for ch in children[k]:
                parent[ch] = p
                children[p].append(ch)
    idx = {k: i for i, k in enumerate([k for k in range(n) if alive[k]])}
    cliques, par2, seps = [], [], []
    for k in range(n):
        if not alive[k]:
            continue
        p = parent[k]
        while p >= 0 and not alive[p]:
            p = parent[p]
        c = np.array(sorted(cl[k]), dtype=np.int64)
        cliques.append(c)
        par2.append(idx[p] if p >= 0 else -1)
    for k, c in enumerate(cliques):
        seps.append(np.intersect1d(c, cliques[par2[k]]) if par2[k] >= 0 else np.zeros(0, dtype=np.int64))
    return CliqueTree(cliques, par2, seps, tree.order)


def _tree_from_sets(cl: List[set], parent: List[int], order: np.ndarray) -> CliqueTree:
    cliques = [np.array(sorted(c), dtype=np.int64) for c in cl]
    seps = [np.intersect1d(c, cliques[parent[k]]) if parent[k] >= 0 else np.zeros(0, dtype=np.int64)
            for k, c in enumerate(cliques)]
    return CliqueTree(cliques, list(parent), seps, order)


def _post_order(parent: List[int]) -> List[int]:
    """children before parents, roots last (post_order, trees.jl)"""
    n = len(parent)
    children: List[List[int]] = [[] for _ in range(n)]
    roots = []
    for k, p in enumerate(parent):
        (children[p] if p >= 0 else roots).append(k)
    out: List[int] = []
    for r in roots:
        stack = [(r, 0)]
        while stack:
            v, i = stack.pop()
            if i < len(children[v]):
                stack.append((v, i + 1))
                stack.append((children[v][i], 0))
            else:
                out.append(v)
    return out


def parent_child_merge_reference(tree: CliqueTree, t_fill: int = 8, t_size: int = 8, snd_post: Optional[List[int]] = None,
                                 return_log: bool = False):
    """ParentChildMerge exactly as the reference walks it (clique_merging.jl:98-106 initialise!, :262-283
    traverse / evaluate, :177-201 merge_child!, :295-303 update_strategy!): the supernode tree is traversed in
    descending topological order (root first); clique c is merged into its *current* parent when
        (|C_par| - |sep_c|) (|C_c| - |sep_c|) <= t_fill   or   max(|snd_c|, |snd_par|) <= t_size ,
    with snd = clique minus its separator.  A merge moves only the child's supernode into the parent's."""
    n = len(tree.cliques)
    sep = [set(x.tolist()) for x in tree.sep]
    snd = [set(c.tolist()) - sep[k] for k, c in enumerate(tree.cliques)]
    parent = list(tree.parent)
    children: List[set] = [set() for _ in range(n)]
    for k, p in enumerate(parent):
        if p >= 0:
            children[p].add(k)
    post = list(snd_post) if snd_post is not None else _post_order(parent)
    log_pairs, log_dec = [], []
    for ind in range(n - 2, -1, -1):             # clique_ind = length(snd) - 1 ... 1 (1-based)
        c = post[ind]
        par = parent[c]
        if par < 0:                               # a second root (disconnected pattern): nothing to merge into
            continue
        d_snd, d_sep = len(snd[c]), len(sep[c])
        p_snd, p_sep = len(snd[par]), len(sep[par])
        fill = ((p_snd + p_sep) - d_sep) * ((d_snd + d_sep) - d_sep)
        do_merge = fill <= t_fill or max(d_snd, p_snd) <= t_size
        log_pairs.append((par, c))
        log_dec.append(do_merge)
        if do_merge:                              # merge_child!
            snd[par] |= snd[c]
            snd[c] = set()
            sep[c] = set()
            for g in children[c]:
                parent[g] = par
            parent[c] = -2                        # removed
            children[par].discard(c)
            children[par] |= children[c]
            children[c] = set()
    alive = [k for k in range(n) if parent[k] != -2]
    idx = {k: i for i, k in enumerate(alive)}
    cl = [snd[k] | sep[k] for k in alive]
    par2 = [idx[parent[k]] if parent[k] >= 0 else -1 for k in alive]
    out = _tree_from_sets(cl, par2, tree.order)
    return (out, log_pairs, log_dec) if return_log else out


# ---- CliqueGraphMerge (the reference's default merge strategy) ------------------------------------------
def reduced_clique_graph(cliques: List[set], seps: List[set]) -> Tuple[List[int], List[int]]:
    """compute_reduced_clique_graph! (clique_graph.jl:19-49, Habib & Stacho): for every separator, largest
    first, the cliques that contain it form the separator graph H (an edge when two of them intersect in more
    than the separator); two such cliques get an edge of the reduced clique graph iff they lie in different
    connected components of H.  Returns (rows, cols) with row > col; separators that occur several times in
    `seps` contribute their edges several times (the reference then *adds* the duplicate weights)."""
    rows: List[int] = []
    cols: List[int] = []
    members: Dict[int, set] = {}
    for k, c in enumerate(cliques):
        for v in c:
            members.setdefault(v, set()).add(k)
    for separator in sorted(seps, key=len, reverse=True):      # sort! is stable, as is sorted()
        if not separator:
            # the empty separator of a root would link every pair of cliques from different connected components
            # of the pattern (O(p^2) edges of weight n1^3 + n2^3 - (n1+n2)^3 < 0 that can never be merged and never
            # make an edge impermissible); they are left out, so a disconnected pattern keeps a forest
            continue
        # cliques that contain the separator, in index order (inverted index instead of a scan over all cliques)
        holders = sorted((members[v] for v in separator), key=len)
        ind = sorted(set.intersection(*holders)) if holders else []
        H: Dict[int, List[int]] = {v: [] for v in ind}
        for a in range(len(ind)):
            for b_ in range(a + 1, len(ind)):
                ca, cb = ind[a], ind[b_]
                if (cliques[ca] & cliques[cb]) != separator:    # inter_equal, clique_graph.jl:113-131
                    H[ca].append(cb)
                    H[cb].append(ca)
        comp: Dict[int, int] = {}
        for v in ind:                                            # find_components (DFS)
            if v in comp:
                continue
            comp[v] = v
            stack = [v]
            while stack:
                u = stack.pop()
                for w in H[u]:
                    if w not in comp:
                        comp[w] = v
                        stack.append(w)
        for a in range(len(ind)):
            for b_ in range(a + 1, len(ind)):
                ca, cb = ind[a], ind[b_]
                if comp[ca] != comp[cb]:
                    rows.append(max(ca, cb))
                    cols.append(min(ca, cb))
    return rows, cols


def _complexity_weight(c_a: set, c_b: set) -> float:
    """ComplexityWeight: |Ca|^3 + |Cb|^3 - |Ca u Cb|^3 (clique_merging.jl:24-31, 395-405)"""
    return float(len(c_a) ** 3 + len(c_b) ** 3 - len(c_a | c_b) ** 3)


class CliqueGraph:
    """State of CliqueGraphMerge (clique_merging.jl:55-67): weighted edges of the reduced clique graph and the
    adjacency table.  Edges are keyed (row, col) with row > col and iterated in CSC order (col, then row), the
    order Julia's `findmax(edges.nzval)` / `findnz` see them in."""

    def __init__(self, cliques: List[set], seps: List[set]):
        self.snd = [set(c) for c in cliques]
        self.num = len(cliques)
        rows, cols = reduced_clique_graph(self.snd, [set(x) for x in seps])
        self.edges: Dict[Tuple[int, int], float] = {}
        for r, c in zip(rows, cols):                             # sparse(rows, cols, weights): duplicates add up
            self.edges[(r, c)] = self.edges.get((r, c), 0.0) + _complexity_weight(self.snd[r], self.snd[c])
        self.edges = {e: w for e, w in self.edges.items() if w != 0.0}
        # the candidate order of traverse(): weight descending, ties in CSC order (col, then row) -- kept incrementally
        self._ranked = SortedList((-w, e[1], e[0]) for e, w in self.edges.items())
        self.adj: Dict[int, set] = {k: set() for k in range(self.num)}
        for (r, c) in self.edges:
            self.adj[r].add(c)
            self.adj[c].add(r)
        self.log: List[Tuple[int, int, bool]] = []

    def _csc(self):
        return sorted(self.edges, key=lambda e: (e[1], e[0]))

    def permissible(self, edge) -> bool:                        # ispermissible, clique_graph.jl:149-158
        c1, c2 = edge
        for nb in self.adj[c1] & self.adj[c2]:
            if (self.snd[c1] & self.snd[nb]) != (self.snd[c2] & self.snd[nb]):
                return False
        return True

    def traverse(self):                                          # clique_merging.jl:242-259
        for (_, c, r) in self._ranked:                           # = sortperm(weights, rev = true), stable in CSC order
            if self.permissible((r, c)):
                return (r, c)
        return None

    def traverse_by_sorting(self):
        """the literal restatement (sort all edges at every call); kept as the cross-check of the incremental order"""
        order = self._csc()
        if not order:
            return None
        by_weight = sorted(order, key=lambda e: -self.edges[e])   # stable: ties keep CSC order
        for e in by_weight:
            if self.permissible(e):
                return e
        return None

    def _set_edge(self, key, w: float) -> None:
        old = self.edges.pop(key, None)
        if old is not None:
            self._ranked.remove((-old, key[1], key[0]))
        if w != 0.0:                                             # dropzeros!
            self.edges[key] = w
            self._ranked.add((-w, key[1], key[0]))

    def merge(self, edge) -> None:                              # merge_two_cliques! + update_strategy!
        c1, removed = edge
        self.snd[c1] |= self.snd[removed]
        self.snd[removed] = set()
        self.num -= 1
        neighbors = set(self.adj[c1])
        new_neighbors = self.adj[removed] - neighbors - {c1}
        for nb in (neighbors - {removed}) | new_neighbors:
            self._set_edge((max(c1, nb), min(c1, nb)), _complexity_weight(self.snd[c1], self.snd[nb]))
        for nb in self.adj[removed]:                           # every edge of `removed` (adj is a superset of edges)
            self._set_edge((max(removed, nb), min(removed, nb)), 0.0)
        self.adj[c1] |= new_neighbors
        for nb in new_neighbors:
            self.adj[nb].add(c1)
        for nb in self.adj[removed]:
            self.adj[nb].discard(removed)
        del self.adj[removed]

    def run(self) -> None:                                      # _merge_cliques!, clique_merging.jl:112-133
        while self.num > 1:
            cand = self.traverse()
            if cand is None:
                break
            do_merge = self.edges[cand] >= 0                    # evaluate, :286-293
            self.log.append((cand[0], cand[1], do_merge))
            if not do_merge:
                break
            self.merge(cand)

    def clique_tree(self, order: np.ndarray) -> CliqueTree:
        """clique_tree_from_graph! (clique_merging.jl:577-600): maximum-weight spanning tree of the clique
        graph under the weights |Ci & Cj| (Kruskal, :480-506), rooted at the clique that holds the vertex
        eliminated last (:532-550)."""
        alive = [k for k in range(len(self.snd)) if self.snd[k]]
        inter = {e: float(len(self.snd[e[0]] & self.snd[e[1]])) for e in self._csc()}
        edges_sorted = sorted(inter, key=lambda e: -inter[e])     # sortperm(V, rev = true), stable
        uf = {k: k for k in alive}

        def find(a):
            while uf[a] != a:
                uf[a] = uf[uf[a]]
                a = uf[a]
            return a

        mst: Dict[int, List[int]] = {k: [] for k in alive}
        found = 0
        for (r, c) in edges_sorted:
            if found >= len(alive) - 1:
                break
            ra, rb = find(r), find(c)
            if ra != rb:
                uf[ra] = rb
                mst[r].append(c)
                mst[c].append(r)
                found += 1
        parent = {k: -1 for k in alive}
        seen = set()
        last = int(order[-1]) if len(order) else -1
        roots = [k for k in alive if last in self.snd[k]][:1] + alive
        for r in roots:                                         # the pattern may be disconnected: several roots
            if r in seen:
                continue
            seen.add(r)
            stack = [r]
            while stack:
                u = stack.pop()
                for v in sorted(mst[u]):
                    if v not in seen:
                        seen.add(v)
                        parent[v] = u
                        stack.append(v)
        idx = {k: i for i, k in enumerate(alive)}
        return _tree_from_sets([self.snd[k] for k in alive], [idx[parent[k]] if parent[k] >= 0 else -1 for k in alive], order)


def clique_graph_merge(tree: CliqueTree) -> CliqueTree:
    """CliqueGraphMerge(edge_weight = ComplexityWeight()), the reference's default `merge_strategy`
    (settings.jl; clique_merging.jl:34-67, 147-166)."""
    g = CliqueGraph([set(c.tolist()) for c in tree.cliques], [set(x.tolist()) for x in tree.sep])
    g.run()
    return g.clique_tree(tree.order)


@dataclass
class DecompositionInfo:
    n_orig: int
    m_orig: int
    sets_orig: list
    # for every decomposed cone: list of (new_row_start, clique vertices)
    blocks: Dict[int, List[Tuple[int, np.ndarray]]] = field(default_factory=dict)
    row_map_plain: List[Tuple[int, int, int]] = field(default_factory=list)   # (old_start, new_start, dim)
    cone_offsets: Dict[int, int] = field(default_factory=dict)
    num_overlaps: int = 0
    clique_sizes: List[int] = field(default_factory=list)
    trees: Dict[int, "CliqueTree"] = field(default_factory=dict)   # clique tree of every decomposed cone


def decompose(P, q, A, b, sets, merge: str = "parent_child", min_dim: int = 3):
    """chordal_decomposition!(ws) for PsdConeTriangle cones (compact transformation).
    Returns (P', q', A', b', sets', info)."""
    A = sp.csr_matrix(A)
    b = np.asarray(b, dtype=np.float64)
    m, n = A.shape
    info = DecompositionInfo(n, m, list(sets))
    rows_new: List[np.ndarray] = []
    cols_new: List[np.ndarray] = []
    vals_new: List[np.ndarray] = []
    b_new: List[np.ndarray] = []
    sets_new = []
    row_ptr = 0
    n_new = n
    off = 0
    Acoo_by_row = A  # csr
    for k, S in enumerate(sets):
        dim = S.dim
        decomposable = isinstance(S, M.PsdConeTriangle) and S.sqrt_dim >= min_dim
        if decomposable:
            N = S.sqrt_dim
            # rows of the cone that hold an entry of A or b -- straight from the row pointers (slicing the CSR matrix
            # would copy a 50-million-row cone: C5 has N = 10 000)
            ip = A.indptr
            row_nnz = ip[off + 1:off + dim + 1] - ip[off:off + dim]
            a_rows = np.nonzero(row_nnz)[0]
            nz_rows = np.unique(np.concatenate([a_rows, np.nonzero(b[off:off + dim])[0]]))
            ii, jj = svec_to_ij(nz_rows)
            diag_rows = np.arange(N, dtype=np.int64) * (np.arange(N, dtype=np.int64) + 1) // 2 + np.arange(N)
            if len(np.union1d(nz_rows, diag_rows)) >= dim:   # dense pattern: keep the cone (chordal_decomposition.jl:53-60)
                decomposable = False
        if not decomposable:
            sub = Acoo_by_row[off:off + dim].tocoo()
            rows_new.append(sub.row + row_ptr)
            cols_new.append(sub.col)
            vals_new.append(sub.data)
            b_new.append(b[off:off + dim])
            sets_new.append(S)
            info.row_map_plain.append((off, row_ptr, dim))
            row_ptr += dim
            off += dim
            continue
        tree = chordal_cliques(N, ii, jj)
        if merge == "parent_child":
            tree = parent_child_merge(tree)
        elif merge == "parent_child_reference":
            tree = parent_child_merge_reference(tree)
        elif merge == "clique_graph":
            tree = clique_graph_merge(tree)
        elif merge != "none":
            raise ValueError("merge must be 'none', 'parent_child', 'parent_child_reference' or 'clique_graph'")
        # row offsets of the clique blocks
        starts = []
        for c in tree.cliques:
            starts.append(row_ptr)
            nc = len(c)
            row_ptr += nc * (nc + 1) // 2
        info.cone_offsets[k] = off
        info.trees[k] = tree
        info.blocks[k] = [(starts[t], tree.cliques[t]) for t in range(len(tree.cliques))]
        info.clique_sizes += [len(c) for c in tree.cliques]
        # owner (clique, local row) of every pattern entry; overlaps get +1/-1 columns
        owner: Dict[int, int] = {}
        for t, c in enumerate(tree.cliques):
            nc = len(c)
            in_sep = np.isin(c, tree.sep[t])
            loc = {int(v): a for a, v in enumerate(c)}
            par_t = tree.parent[t]
            par_loc = {int(v): a for a, v in enumerate(tree.cliques[par_t])} if par_t >= 0 else None
            ov_rows, ov_cols, ov_vals = [], [], []
            for bj in range(nc):
                for ai in range(bj + 1):
                    gi, gj = int(c[ai]), int(c[bj])
                    new_row = starts[t] + svec_index(ai, bj)
                    if in_sep[ai] and in_sep[bj]:
                        pa, pb = par_loc[gi], par_loc[gj]
                        if pa > pb:
                            pa, pb = pb, pa
                        ov_rows += [new_row, starts[par_t] + svec_index(pa, pb)]
                        ov_cols += [n_new, n_new]
                        ov_vals += [1.0, -1.0]
                        n_new += 1
                    else:
                        owner[svec_index(gi, gj)] = new_row
            if ov_rows:
                rows_new.append(np.array(ov_rows, dtype=np.int64))
                cols_new.append(np.array(ov_cols, dtype=np.int64))
                vals_new.append(np.array(ov_vals))
            sets_new.append(M.PsdConeTriangle(nc * (nc + 1) // 2))
        lo, hi = int(ip[off]), int(ip[off + dim])
        sub_row = np.repeat(a_rows, row_nnz[a_rows])             # cone-local row of every entry, CSR order
        mapped = np.array([owner[int(r)] for r in sub_row], dtype=np.int64) if hi > lo else np.zeros(0, dtype=np.int64)
        rows_new.append(mapped)
        cols_new.append(A.indices[lo:hi].astype(np.int64))
        vals_new.append(A.data[lo:hi].copy())
        bseg = np.zeros(row_ptr - starts[0])
        nzb = np.nonzero(b[off:off + dim])[0]
        for r in nzb:
            bseg[owner[int(r)] - starts[0]] = b[off + r]
        b_new.append(bseg)
        off += dim
    info.num_overlaps = n_new - n
    rows_c = np.concatenate(rows_new) if rows_new else np.zeros(0, dtype=np.int64)
    cols_c = np.concatenate(cols_new) if cols_new else np.zeros(0, dtype=np.int64)
    vals_c = np.concatenate(vals_new) if vals_new else np.zeros(0)
    A2 = sp.csc_matrix((vals_c, (rows_c, cols_c)), shape=(row_ptr, n_new))
    b2 = np.concatenate(b_new) if b_new else np.zeros(0)
    P2 = sp.block_diag([sp.csc_matrix(P), sp.csc_matrix((n_new - n, n_new - n))], format="csc")
    q2 = np.concatenate([np.asarray(q, dtype=np.float64), np.zeros(n_new - n)])
    return P2, q2, A2, b2, sets_new, info


def psd_complete(Y: np.ndarray, tree: CliqueTree, assume_symmetric: bool = False) -> np.ndarray:
    """psd_complete! (chordal_decomposition.jl:262-311): fill the entries of the symmetric matrix `Y` that lie
    outside the cliques of `tree` so that the result is positive semidefinite (given that every clique block
    is).  Cliques are visited parents first; for clique k with separator alpha = C_k & C_parent and residual
    nu = C_k minus alpha, the unknown entries between nu and eta = (vertices of the cliques visited so far) minus C_k are
        Y[eta, nu] = Y[eta, alpha] Y[alpha, alpha]^-1 Y[alpha, nu]
    (pseudo-inverse when the separator block is singular, as the reference's try/catch does).

    The vertices are renumbered in the order the traversal first meets them: the visited set is then a leading block,
    the residual of the current clique the next few indices, and the update is one product of a contiguous row block
    (rows of alpha are recomputed to themselves and restored, so known entries stay bit-identical)."""
    W = np.array(Y, dtype=np.float64)
    if not assume_symmetric:                 # the reference reads the upper triangle
        W = np.triu(W) + np.triu(W, 1).T
    N = W.shape[0]
    ncl = len(tree.cliques)
    children: List[List[int]] = [[] for _ in range(ncl)]
    roots = []
    for k, p in enumerate(tree.parent):
        (children[p] if p >= 0 else roots).append(k)
    # traversal order (parents first) and the renumbering it induces
    order_k: List[int] = []
    stack = list(reversed(roots))
    while stack:
        k = stack.pop()
        order_k.append(k)
        stack.extend(reversed(children[k]))
    new_of = np.full(N, -1, dtype=np.int64)
    nxt = 0
    residuals: Dict[int, np.ndarray] = {}
    for k in order_k:
        c = np.asarray(tree.cliques[k], dtype=np.int64)
        fresh = c[new_of[c] < 0]
        residuals[k] = fresh
        new_of[fresh] = np.arange(nxt, nxt + len(fresh))
        nxt += len(fresh)
    rest = np.nonzero(new_of < 0)[0]                       # vertices in no clique (cannot happen for a clique tree)
    new_of[rest] = np.arange(nxt, nxt + len(rest))
    perm = np.argsort(new_of)                              # perm[new] = old
    W = W[np.ix_(perm, perm)]
    seen = 0
    for k in order_k:
        nu_old = residuals[k]
        nn = len(nu_old)
        alpha_old = tree.sep[k] if tree.parent[k] >= 0 else np.zeros(0, dtype=np.int64)
        # the reference's nu = C_k minus alpha; vertices of alpha are always met before (they belong to the parent)
        lo, hi = seen, seen + nn                           # new indices of nu
        if seen and nn:
            if len(alpha_old):
                al = new_of[np.asarray(alpha_old, dtype=np.int64)]
                Waa = W[np.ix_(al, al)]
                Wan = W[al, lo:hi].copy()
                try:
                    Z = np.linalg.solve(Waa, Wan)
                    if not np.all(np.isfinite(Z)):
                        raise np.linalg.LinAlgError
                except np.linalg.LinAlgError:
                    Z = np.linalg.pinv(Waa) @ Wan
                blk = W[:seen, al] @ Z
                blk[al, :] = Wan                           # known entries (alpha x nu lies inside the clique) stay exact
            else:   # a new connected component: no coupling with what was completed before
                blk = np.zeros((seen, nn))
            # entries between nu and the OTHER members of its own clique are known as well: keep them
            c_new = new_of[np.asarray(tree.cliques[k], dtype=np.int64)]
            c_seen = c_new[c_new < seen]
            blk[c_seen, :] = W[c_seen, lo:hi]
            W[:seen, lo:hi] = blk
            W[lo:hi, :seen] = blk.T
        seen = hi
    inv = new_of                                           # old -> new
    return W[np.ix_(inv, inv)]


def _psd_complete_reference(Y: np.ndarray, tree: CliqueTree) -> np.ndarray:
    """the literal restatement (index sets per clique); kept as the cross-check of the renumbered version"""
    W = np.array(Y, dtype=np.float64)
    W = np.triu(W) + np.triu(W, 1).T
    ncl = len(tree.cliques)
    children: List[List[int]] = [[] for _ in range(ncl)]
    roots = []
    for k, p in enumerate(tree.parent):
        (children[p] if p >= 0 else roots).append(k)
    seen = np.zeros(W.shape[0], dtype=bool)
    stack = list(reversed(roots))
    while stack:
        k = stack.pop()
        c = tree.cliques[k]
        alpha = tree.sep[k] if tree.parent[k] >= 0 else np.zeros(0, dtype=np.int64)
        nu = np.setdiff1d(c, alpha)
        eta = np.nonzero(seen)[0]
        eta = np.setdiff1d(eta, c)
        if len(eta) and len(nu):
            if len(alpha):
                Waa = W[np.ix_(alpha, alpha)]
                Wan = W[np.ix_(alpha, nu)]
                try:
                    Z = np.linalg.solve(Waa, Wan)
                    if not np.all(np.isfinite(Z)):
                        raise np.linalg.LinAlgError
                except np.linalg.LinAlgError:
                    Z = np.linalg.pinv(Waa) @ Wan
                blk = W[np.ix_(eta, alpha)] @ Z
            else:
                blk = np.zeros((len(eta), len(nu)))
            W[np.ix_(eta, nu)] = blk
            W[np.ix_(nu, eta)] = blk.T
        seen[c] = True
        stack.extend(reversed(children[k]))
    return W


def _svec_to_mat(v: np.ndarray, N: int) -> np.ndarray:
    """populate_upper_triangle!(X, v, 1/sqrt 2) + symmetrise (convexset.jl:432-442).  The column-major upper triangle
    (0,0), (0,1), (1,1), (0,2) ... is the row-major lower triangle of the transpose: one masked assignment."""
    L = np.zeros((N, N))
    L[np.tri(N, dtype=bool)] = v
    d = np.diagonal(L).copy()
    X = L + L.T
    X *= 1.0 / np.sqrt(2.0)
    np.fill_diagonal(X, d)
    return X


def _mat_to_svec(X: np.ndarray) -> np.ndarray:
    """extract_upper_triangle!(X, v, sqrt 2) of a symmetric X: X[c, r] for r = 0.., c <= r, i.e. its row-major lower triangle"""
    N = X.shape[0]
    v = X[np.tri(N, dtype=bool)] * np.sqrt(2.0)
    j = np.arange(N, dtype=np.int64)
    v[j * (j + 1) // 2 + j] = np.diagonal(X)
    return v


def reverse(info: DecompositionInfo, x2, s2, mu2, complete_dual: bool = False):
    """reverse_decomposition! (chordal_decomposition.jl:129-213): x = x'[1:n]; s = sum of clique blocks;
    mu = the clique block's value (overlaps carry equal values at optimality).  With `complete_dual`
    (settings.complete_dual, :146) the entries of every decomposed dual matrix outside its cliques are
    filled by `psd_complete` so that y = -mu is in the PSD cone."""
    x = np.asarray(x2)[:info.n_orig].copy()
    s = np.zeros(info.m_orig)
    mu = np.zeros(info.m_orig)
    for old, new, dim in info.row_map_plain:
        s[old:old + dim] = s2[new:new + dim]
        mu[old:old + dim] = mu2[new:new + dim]
    for k, blocks in info.blocks.items():
        off = info.cone_offsets[k]
        for start, c in blocks:
            nc = len(c)
            for bj in range(nc):
                gj = int(c[bj])
                gi = c[:bj + 1]
                orig = off + gj * (gj + 1) // 2 + gi
                seg = slice(start + svec_index(0, bj), start + svec_index(bj, bj) + 1)
                s[orig] += s2[seg]
                mu[orig] = mu2[seg]
        if complete_dual:   # complete!(mu, ::PsdConeTriangle, ...), chordal_decomposition.jl:245-257
            S = info.sets_orig[k]
            N = S.sqrt_dim
            seg = slice(off, off + S.dim)
            Y = psd_complete(_svec_to_mat(-mu[seg], N), info.trees[k], assume_symmetric=True)
            mu[seg] = -_mat_to_svec(Y)
    return x, s, mu
