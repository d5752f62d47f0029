"""Measurement helper (not a test): cost and benefit of the Anderson accelerator on one GPU.
Prints one JSON line per run: iterations to the default tolerance and device time per iteration,
with accelerator = EmptyAccelerator / AndersonAccelerator on the same problem."""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import cosmo_b200

pr = cosmo_b200.problems
cases = [("qp_n20k_m40k", lambda: pr.random_sparse_qp(20000, 40000, 0.005, seed=2), dict(scaling=0)),
         ("socp_n2000_k200", lambda: pr.portfolio_socp(n=2000, k=200, seed=1), dict(scaling=0, max_iter=5000)),
         ("sdp_closest_corr_N200", lambda: pr.closest_correlation_sdp(N=200, seed=3), dict(scaling=0))]
for name, gen, kw in cases:
    P, q, A, b, sets = gen()
    for acc in ("EmptyAccelerator", "AndersonAccelerator"):
        model = cosmo_b200.Model()
        model.set(P, q, A, b, sets, cosmo_b200.Settings(accelerator=acc, **kw))
        t0 = time.perf_counter()
        res = model.optimize()
        wall = time.perf_counter() - t0
        dev = res.times["iter_time_device"]
        print(json.dumps({"case": name, "accelerator": acc, "status": res.status, "iter": res.iter,
                          "safeguarding_iter": res.safeguarding_iter, "obj": res.obj_val,
                          "device_s": round(dev, 5), "ms_per_iter": round(1e3 * dev / max(res.iter, 1), 4),
                          "wall_s": round(wall, 3), "kkt_inner": res.kkt_inner_iterations}), flush=True)
        model.empty_model()
