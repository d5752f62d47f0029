"""Measurement / bring-up helper (not a test): the int8-sliced tcgen05 product kernel against numpy dgemm.

    python tests/run_tc_gemm.py [N ...]
"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cosmo_b200  # noqa: E402
from cosmo_b200 import engine as E  # noqa: E402

Ns = [int(a) for a in sys.argv[1:]] or [128, 200, 256, 1000, 2000]
variants = [(8, 10), (8, 8), (7, 7), (6, 8), (4, 6)]
if os.environ.get("TC_VARIANTS"):
    variants = [tuple(int(x) for x in v.split(",")) for v in os.environ["TC_VARIANTS"].split(";")]
for N in Ns:
    rng = np.random.default_rng(N)
    G = rng.standard_normal((N, N))
    A = (G + G.T) / np.sqrt(2.0 * N)
    A[0, :] *= 1e-3; A[:, 0] *= 1e-3          # a row with a much smaller scale
    B = A @ A                                  # commutes with A
    B = (B + B.T) / 2
    ref = A @ B
    nrm = np.abs(A) @ np.abs(B)
    for (k, groups) in variants:
        try:
            Cm, ms, fr = E.tc_gemm(A, B, slices=k, groups=groups, reps=(10 if N >= 1000 else 2))
        except Exception as ex:   # noqa: BLE001
            print(json.dumps({"N": N, "k": k, "groups": groups, "error": str(ex)}), flush=True)
            continue
        err = float(np.max(np.abs(Cm - ref) / nrm))
        row = {"N": N, "k": k, "groups": groups, "max_rel_err_vs_absAabsB": err,
               "fro_rel_err": float(np.linalg.norm(Cm - ref) / np.linalg.norm(ref)),
               "asym": float(np.max(np.abs(Cm - Cm.T))), "ms": round(ms, 4),
               "frob2_ok": bool(abs(fr[0] - np.sum(Cm * Cm)) <= 1e-10 * np.sum(Cm * Cm)),
               "fp64_equiv_TFlops": round(2.0 * (N ** 3) / (ms * 1e-3) / 1e12, 1) if ms > 0 else None}
        print(json.dumps(row), flush=True)
