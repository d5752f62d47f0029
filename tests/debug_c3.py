import numpy as np, sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cosmo_b200
from oracle import cosmo_oracle as O
P, q, A, b, sets = cosmo_b200.problems.portfolio_socp(n=400, k=40, seed=1)
ref = O.solve(P, q, A, b, cosmo_b200.problems.to_oracle_cones(sets), O.Settings(kkt_solver="cg", max_iter=3000))
print("oracle", ref.status, ref.iter, ref.obj_val, sum(ref.kkt.inner_iterations), flush=True)
for t in range(int(sys.argv[1]) if len(sys.argv) > 1 else 5):
    model = cosmo_b200.Model()
    model.set(P, q, A, b, sets, cosmo_b200.Settings(max_iter=int(sys.argv[2]) if len(sys.argv) > 2 else 3000))
    t0 = time.time()
    res = model.optimize()
    print(t, res.status, res.iter, res.obj_val, res.kkt_inner_iterations, res.info.rho_updates, "%.2fs" % (time.time() - t0), flush=True)
