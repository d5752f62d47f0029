import numpy as np, sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import scipy.sparse as sp
import cosmo_b200
from cosmo_b200 import engine as E
from oracle import cosmo_oracle as O
from oracle.bridge import to_oracle_cones
for N in [int(a) for a in sys.argv[1:]] or [200, 500, 1000, 2000]:
    rng = np.random.default_rng(N)
    d = N * (N + 1) // 2
    sets = [cosmo_b200.PsdConeTriangle(d)]
    eng = E.Engine(sp.identity(1, format="csc"), np.zeros(1), sp.csc_matrix((d, 1)), np.zeros(d),
                   [(S.code, S.dim, None, None) for S in sets], cosmo_b200.Settings(scaling=0).to_struct())
    ws = rng.standard_normal(d)
    t0 = time.time(); ref = ws.copy(); O.project(ref, to_oracle_cones(sets)); tcpu = time.time() - t0
    eng.project(ws)
    t0 = time.time(); got = eng.project(ws); tgpu = time.time() - t0
    # a nearby matrix (ADMM-like small change)
    ws2 = got + 1e-3 * rng.standard_normal(d)
    t0 = time.time(); got2 = eng.project(ws2); tgpu2 = time.time() - t0
    ref2 = ws2.copy(); O.project(ref2, to_oracle_cones(sets))
    err2 = np.linalg.norm(got2 - ref2) / np.linalg.norm(ws2)
    ws3 = got2 + 1e-5 * rng.standard_normal(d)
    t0 = time.time(); got3 = eng.project(ws3); tgpu3 = time.time() - t0
    print("   nearby relerr %.2e, third (1e-5 away) %.3fs" % (err2, tgpu3))
    print("N=%d  cpu dsyevr+syrk %.3fs  gpu %.3fs (nearby %.3fs)  relerr %.2e" % (
        N, tcpu, tgpu, tgpu2, np.linalg.norm(got - ref) / np.linalg.norm(ws)), flush=True)
