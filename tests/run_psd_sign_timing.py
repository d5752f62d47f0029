"""Measurement helper (not a test): the GEMM-only PSD projections (csrc/psd_tc.cuh on the tensor cores -- the default
for N >= 192 -- and csrc/psd_sign.cuh on the FP64 FMA pipe) against the block-Jacobi path on the same matrices -- accuracy vs the CPU oracle, wall time
of `Engine.project` (includes the host <-> device copies of the N(N+1)/2 vector) and the step counts printed by
COSMO_B200_PSD_DEBUG=1.

    COSMO_B200_PSD_DEBUG=1 python tests/run_psd_sign_timing.py 300 1000 2000
"""
import json
import os
import sys
import time

import numpy as np
import scipy.sparse as sp

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cosmo_b200  # noqa: E402
from cosmo_b200 import engine as E  # noqa: E402
from oracle import cosmo_oracle as O  # noqa: E402
from oracle.bridge import to_oracle_cones  # noqa: E402


def matrices(N, rng):
    B = rng.standard_normal((N, N))
    yield "wigner", (B + B.T) / 2
    k = max(N // 10, 1)
    yield "rank_deficient", B[:, :k] @ B[:, :k].T - B[:, k:2 * k] @ B[:, k:2 * k].T
    # an ADMM-like iterate: PSD part plus a scaled negative part (w_s = s + mu / rho)
    Q, _ = np.linalg.qr(B)
    lam = np.concatenate([np.abs(rng.standard_normal(N // 2)), -10.0 * np.abs(rng.standard_normal(N - N // 2))])
    yield "split_spectrum", (Q * lam) @ Q.T


def svec(X):
    N = X.shape[0]
    iu = np.triu_indices(N)
    order = np.lexsort((iu[0], iu[1]))
    i, j = iu[0][order], iu[1][order]
    v = X[i, j].copy()
    v[i != j] *= np.sqrt(2.0)
    return v


for N in [int(a) for a in sys.argv[1:]] or [300, 1000, 2000]:
    rng = np.random.default_rng(N)
    d = N * (N + 1) // 2
    sets = [cosmo_b200.PsdConeTriangle(d)]
    for name, X in matrices(N, rng):
        ws = svec(X)
        t0 = time.time()
        ref = ws.copy()
        O.project(ref, to_oracle_cones(sets))
        t_cpu = time.time() - t0
        row = {"N": N, "matrix": name, "cpu_dsyevr_s": round(t_cpu, 4)}
        for mode in ("0", "1", "tc"):     # block Jacobi (cold) / Newton-Schulz on the FP64 FMA pipe / on tcgen05 int8 slices
            os.environ["COSMO_B200_PSD_SIGN"] = "1" if mode == "1" else "0"
            os.environ["COSMO_B200_PSD_TC"] = "1" if mode == "tc" else "0"
            os.environ["COSMO_B200_PSD_WARM"] = "0"
            eng = E.Engine(sp.identity(1, format="csc"), np.zeros(1), sp.csc_matrix((d, 1)), np.zeros(d),
                           [cosmo_b200.model.set_tuple(S) for S in sets], cosmo_b200.Settings(scaling=0).to_struct())
            eng.project(ws)                      # warm-up (allocations, first launches)
            t0 = time.time()
            got = eng.project(ws)
            dt = time.time() - t0
            key = {"0": "jacobi_cold", "1": "sign_fp64fma", "tc": "sign_tc"}[mode]
            row[key + "_s"] = round(dt, 4)
            row[key + "_relerr"] = float(np.linalg.norm(got - ref) / (np.linalg.norm(ws) + 1e-300))
            eng.close()
        print(json.dumps(row), flush=True)
