import numpy as np, sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cosmo_b200
from cosmo_b200 import engine as E
P, q, A, b, sets = cosmo_b200.problems.portfolio_socp(n=20_000, k=2_000, seed=1)
eng = E.Engine(P, q, A, b, [(S.code, S.dim, getattr(S, "l", None), getattr(S, "u", None)) for S in sets],
               cosmo_b200.Settings(scaling=0).to_struct())
for which, name in ((0, "A pass"), (1, "At plain"), (3, "At + P op")):
    ms, nb = eng.spmv_bench(which, 20)
    print("%-10s %.1f us  %.2f TB/s (algorithmic %.0f MB)" % (name, ms * 1e3, nb / ms / 1e9, nb / 1e6), flush=True)
rows = np.diff(A.tocsr().indptr); cols = np.diff(A.indptr)
print("A rows: max %d, >256: %d ; At rows: max %d mean %.0f" % (rows.max(), (rows > 256).sum(), cols.max(), cols.mean()))
