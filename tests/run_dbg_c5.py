import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import cosmo_b200
from cosmo_b200 import chordal, sharding
from oracle import cosmo_oracle as O
from oracle.bridge import to_oracle_cones
nv = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
rows, cols, w = cosmo_b200.problems.banded_random_graph(nv, 3.0, 20, seed=1)
P, q, A, b, sets = cosmo_b200.problems.maxcut_dual_sdp(nv, rows, cols, w)
P2, q2, A2, b2, sets2, info = chordal.decompose(P, q, A, b, sets, merge="parent_child")
cones = to_oracle_cones(sets2)
n, m = A2.shape[1], A2.shape[0]
print("n", n, "m", m, "cones", len(sets2), "max clique", max(info.clique_sizes))
for kk in (1, 2, 3):
    stk = dict(scaling=0, adaptive_rho=False, max_iter=kk, eps_abs=0.0, eps_rel=0.0)
    ref = O.solve(P2, q2, A2, b2, cones, O.Settings(kkt_solver="cg", **stk))
    sh = sharding.make_shard(P2, q2, A2, b2, sets2, 0, 1)
    eng = sharding.create_engine(sh, cosmo_b200.Settings(**stk), device=0)
    o = eng.solve()
    wg = eng.w()
    d = np.abs(wg - ref.w)
    sc = np.max(np.abs(ref.w))
    offs = np.concatenate([[0], np.cumsum([S.dim for S in sets2])])
    worst = int(np.argmax(d[n:]))
    k = int(np.searchsorted(offs, worst, side="right") - 1)
    print("iters", kk, "w_rel x-part %.2e s-part %.2e" % (d[:n].max() / sc, d[n:].max() / sc), "cg", o.kkt_inner_iterations, int(np.sum(ref.kkt.inner_iterations)),
          "worst row in cone", k, type(sets2[k]).__name__, sets2[k].dim, "ds %.2e dx(out) %.2e" % (np.max(np.abs(o.s - ref.s)), np.max(np.abs(o.x - ref.x))), flush=True)
    eng.close()
