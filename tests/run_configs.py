"""BASELINE configs C1, C3, C4 at full size on one GPU: ADMM iterations/s next to the oracle port
(CPU, same box).  One JSON line per config.  Usage: python tests/run_configs.py [c1] [c3] [c4]"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cosmo_b200
from oracle import cosmo_oracle as O
from oracle.bridge import to_oracle_cones


def timed(name, P, q, A, b, sets, iters, cpu_iters, **kw):
    st = dict(scaling=0, adaptive_rho=False, eps_abs=0.0, eps_rel=0.0)
    st.update(kw)
    model = cosmo_b200.Model()
    model.set(P, q, A, b, sets, cosmo_b200.Settings(max_iter=iters, **st))
    t0 = time.time()
    res = model.optimize()                     # includes engine creation (setup)
    setup_plus = time.time() - t0
    model.engine.reset()
    model.x[:] = 0; model.s[:] = 0; model.mu[:] = 0
    res = model.optimize()
    dev = res.times["iter_time_device"]
    line = {"config": name, "n": A.shape[1], "m": A.shape[0], "nnz_A": int(A.nnz), "iters": iters,
            "iter_per_s": iters / dev, "ms_per_iter": 1e3 * dev / iters,
            "kkt_inner_per_iter": res.kkt_inner_iterations / max(res.iter, 1), "first_solve_incl_setup_s": setup_plus,
            "obj": res.obj_val, "r_prim": res.info.r_prim, "r_dual": res.info.r_dual}
    if cpu_iters:
        cones = to_oracle_cones(sets)
        t0 = time.time()
        ref = O.solve(P, q, A, b, cones, O.Settings(kkt_solver="cg", max_iter=cpu_iters, **st))
        cpu = time.time() - t0
        line["cpu_oracle_iter_per_s"] = cpu_iters / cpu
        line["speedup_vs_cpu_oracle"] = line["iter_per_s"] / line["cpu_oracle_iter_per_s"]
        if cpu_iters == iters:
            line["obj_cpu"] = ref.obj_val
            line["max_abs_dx"] = float(np.max(np.abs(res.x - ref.x)))
        # parity at full size: the operator variable w after the oracle's iterations on identical arrays
        model.engine.update_settings(cosmo_b200.Settings(max_iter=cpu_iters, **st).to_struct())
        model.engine.reset()
        model.engine.warm_start(np.zeros(A.shape[1]), np.zeros(A.shape[0]), np.zeros(A.shape[0]))
        model.engine.solve()
        w_gpu = model.engine.w()
        line["parity_w_rel"] = float(np.max(np.abs(w_gpu - ref.w)) / max(np.max(np.abs(ref.w)), 1e-300))
        line["parity_iters"] = cpu_iters
        line["parity_ok"] = bool(line["parity_w_rel"] <= 1e-8)
        line["psd_stats"] = model.engine.psd_stats()
    print(json.dumps(line), flush=True)


def main():
    which = [a for a in sys.argv[1:]] or ["c1", "c3", "c4"]
    pr = cosmo_b200.problems
    if "c1" in which:
        import scipy.sparse as sp
        P = sp.csc_matrix(np.array([[4.0, 1.0], [1.0, 2.0]]))
        q = np.array([1.0, 1.0])
        Am = np.array([[1.0, 1.0], [1.0, 0.0], [0.0, 1.0]])
        A = sp.csc_matrix(np.vstack([Am, -Am]))          # model form of examples/qp.jl:19-21 (A = -Aa)
        b = np.array([1.0, 0.7, 0.7, -1.0, 0.0, 0.0])
        timed("C1 examples/qp.jl (n=2, m=6, Nonnegatives)", P, q, A, b, [cosmo_b200.Nonnegatives(6)], 375, 375)
    if "c3" in which:
        P, q, A, b, sets = pr.portfolio_socp(n=20_000, k=2_000, seed=1)
        timed("C3 portfolio SOCP n=20000 k=2000 (Zero(1)+Nonneg(n)+SOC(1+n+k))", P, q, A, b, sets, 50, 3)
    if "c4" in which:
        P, q, A, b, sets = pr.closest_correlation_sdp(N=2000, seed=12345)
        timed("C4 closest correlation N=2000 (Zero(N)+PsdConeTriangle(2001000))", P, q, A, b, sets, 20, 2)


if __name__ == "__main__":
    main()
