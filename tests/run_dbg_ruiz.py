import sys
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np, scipy.sparse as sp
import cosmo_b200
from cosmo_b200 import engine as E
pr = cosmo_b200.problems
for name, (P, q, A, b, sets) in (("qp_box", pr.random_sparse_qp(300, 500, 0.05, seed=0)), ("qp_wide", pr.random_sparse_qp(30000, 4000, 0.002, seed=5))):
    st = cosmo_b200.Settings()
    eng = E.Engine(P, q, A, b, [cosmo_b200.model.set_tuple(S) for S in sets], st.to_struct(), equilibrate=True)
    D, Ev, c = eng.scaling()
    rng = np.random.default_rng(1)
    x, y = rng.standard_normal(A.shape[1]), rng.standard_normal(A.shape[0])
    A = sp.csc_matrix(A)
    for which, v in ((0, x), (1, y)):
        got = eng.spmv(which, v)
        M = A if which == 0 else A.T
        l, r = (Ev, D) if which == 0 else (D, Ev)
        cands = {"full": l * (M @ (r * v)), "none": M @ v, "left": l * (M @ v), "right": M @ (r * v), "full2": l * l * (M @ (r * r * v))}
        print(name, which, {k: float(np.max(np.abs(got - c_))) for k, c_ in cands.items()}, "windowed?", A.shape)
