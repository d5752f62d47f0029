import sys, time, os
sys.path.insert(0, '.')
import numpy as np, scipy.sparse as sp
import cosmo_b200
from oracle import fast_matvec as F
P, q, A, b, sets = cosmo_b200.problems.random_sparse_qp(50000, 100000, 0.01, 2)
T = F.ThreadedCsr(A)
x = np.ones(A.shape[1])
lib = F.load()
for t in (1, 2, 4, 8, 16, 32, 64, 128):
    lib.oracle_spmv_set_threads(t)
    T @ x
    t0 = time.perf_counter()
    for _ in range(3): T @ x
    print(t, (time.perf_counter() - t0) / 3, flush=True)
Ac = sp.csc_matrix(A)
t0 = time.perf_counter(); Ac @ x; print('scipy csc', time.perf_counter() - t0)
print(open('/sys/fs/cgroup/cpu.max').read() if os.path.exists('/sys/fs/cgroup/cpu.max') else 'no cpu.max')
