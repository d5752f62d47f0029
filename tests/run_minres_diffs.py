"""Measured distance between the engine's and the oracle's MINRES solves (bound of test_minres_solve_matches_oracle)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np
import cosmo_b200
from oracle import cosmo_oracle as O
from oracle.bridge import to_oracle_cones
import test_gpu_parity as TG
for seed in (12, 13, 14):
    for name, kind in (("MINRESIndirectKKTSolver", "minres"), ("IndirectReducedKKTSolver:MINRES", "minres_reduced")):
        P, q, A, b, sets = TG._small_qp(seed=seed)
        ref = O.solve(P, q, A, b, to_oracle_cones(sets), O.Settings(kkt_solver=kind, max_iter=300))
        model = cosmo_b200.Model()
        model.set(P, q, A, b, sets, cosmo_b200.Settings(kkt_solver=name, max_iter=300))
        res = model.optimize()
        print(seed, kind, res.status, ref.status, res.iter, ref.iter, "dobj_rel %.3e" % (abs(res.obj_val - ref.obj_val) / max(1, abs(ref.obj_val))),
              "dx_rel %.3e" % (np.max(np.abs(res.x - ref.x)) / max(1, np.abs(ref.x).max())), "r_prim %.2e/%.2e" % (res.info.r_prim, ref.info.r_prim), flush=True)
