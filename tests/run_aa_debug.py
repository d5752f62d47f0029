"""Debug helper: accelerated trajectories engine vs oracle for a range of iteration counts."""
import sys
import numpy as np
sys.path.insert(0, ".")
import cosmo_b200
from oracle import cosmo_oracle as O

P, q, A, b, sets = cosmo_b200.problems.random_sparse_qp(40, 70, 0.15, seed=7)
cones = cosmo_b200.problems.to_oracle_cones(sets)
for iters in list(range(14, 36)) + [40, 41, 42, 45]:
    ost = O.Settings(kkt_solver="cg", scaling=0, max_iter=iters, eps_abs=1e-14, eps_rel=1e-14, accelerator="anderson")
    ref = O.solve(P, q, A, b, cones, ost)
    model = cosmo_b200.Model()
    model.set(P, q, A, b, sets, cosmo_b200.Settings(scaling=0, max_iter=iters, eps_abs=1e-14, eps_rel=1e-14,
                                                    accelerator="AndersonAccelerator"))
    res = model.optimize()
    w = model.engine.w()
    print(iters, "sg", res.safeguarding_iter, ref.safeguarding_iter, "relerr %.3e" % (np.linalg.norm(w - ref.w) / np.linalg.norm(ref.w)),
          "rho", len(res.info.rho_updates), flush=True)
