"""Multi-threaded sparse mat-vec for the CPU reference arm -- TEST / BENCH INFRASTRUCTURE, NOT THE PRODUCT.

``threaded(ws)`` swaps the sparse operators of an oracle ``Workspace`` (after ``setup()``) for row-parallel CSR
kernels (oracle/spmv_omp.c, OpenMP): the oracle's arithmetic is unchanged per row, rows run on all host threads.
Only bench.py (``--impl reference`` and ``cpu_baseline``) and tests/ may import this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np
import scipy.sparse as sp

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "spmv_omp.c")
LIB = os.path.join(HERE, "_build", "liboracle_spmv.so")
_lib = None


def build(force=False):
    """gcc -O3 -fopenmp the C restatement (called by __graft_entry__.build())."""
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(SRC):
        return LIB
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    subprocess.check_call(["gcc", "-O3", "-march=x86-64-v2", "-fopenmp", "-shared", "-fPIC", "-o", LIB, SRC])
    return LIB


def load():
    global _lib
    if _lib is None:
        build()
        os.environ.setdefault("OMP_PROC_BIND", "false")
        _lib = C.CDLL(LIB)
        _lib.oracle_spmv_set_threads.argtypes = [C.c_int]
        _lib.oracle_spmv_set_threads.restype = None
        _lib.oracle_csr_matvec.argtypes = [C.c_int64] + [C.c_void_p] * 5
        _lib.oracle_csr_matvec.restype = None
        _lib.oracle_spmv_max_threads.restype = C.c_int
    return _lib


def max_threads():
    return int(load().oracle_spmv_max_threads())


class ThreadedCsr:
    """Duck-types the two things the oracle does with a sparse matrix: ``M @ x`` and ``M.shape``."""

    def __init__(self, M):
        M = sp.csr_matrix(M)
        M.sort_indices()
        self.shape = M.shape
        self.indptr = np.ascontiguousarray(M.indptr, dtype=np.int32)
        self.indices = np.ascontiguousarray(M.indices, dtype=np.int32)
        self.data = np.ascontiguousarray(M.data, dtype=np.float64)
        self._lib = load()

    def __matmul__(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.empty(self.shape[0], dtype=np.float64)
        self._lib.oracle_csr_matvec(self.shape[0], self.indptr.ctypes.data, self.indices.ctypes.data, self.data.ctypes.data,
                                    x.ctypes.data, y.ctypes.data)
        return y


_threads = None


def usable_cpus():
    """CPUs this process may actually burn: the affinity mask capped by the cgroup CPU quota (a container on the GPU
    box sees 128 cores and is allowed 16; more threads than that only buys throttling)."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        quota, period = open("/sys/fs/cgroup/cpu.max").read().split()[:2]
        if quota != "max":
            n = max(1, min(n, int(int(quota) // int(period))))
    except Exception:
        pass
    return n


def calibrate(M, candidates=None, reps=2):
    """Pick the thread count that is actually fastest for y = M x on this host (a container can see 128 cores and be
    allowed 8): returns (threads, seconds per product).  The choice is kept for all later products."""
    import time
    global _threads
    lib = load()
    T = M if isinstance(M, ThreadedCsr) else ThreadedCsr(M)
    ncores = usable_cpus()
    cand = candidates or [t for t in (1, 2, 4, 8, 16, 32, 64, 128, 256) if t <= ncores]
    x = np.ones(T.shape[1])
    best = (None, float("inf"))
    for t in cand:
        lib.oracle_spmv_set_threads(t)
        T @ x                                   # start the team
        t0 = time.perf_counter()
        for _ in range(reps):
            T @ x
        dt = (time.perf_counter() - t0) / reps
        if dt < best[1]:
            best = (t, dt)
    _threads = best[0]
    lib.oracle_spmv_set_threads(_threads)
    return best


def set_threads(t):
    """Fix the thread count of the following products (in-situ calibration by bench.py)."""
    global _threads
    _threads = int(t)
    load().oracle_spmv_set_threads(_threads)


def thread_candidates():
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    q = usable_cpus()
    return sorted({t for t in (q // 2, q, 2 * q, 4 * q) if 1 <= t <= n})


def threads_in_use():
    return _threads if _threads is not None else max_threads()


def threaded(ws):
    """Replace the operators of the KKT solver of an oracle Workspace (call after ws.setup()); the first call
    calibrates the thread count on the workspace's own A."""
    k = ws.kkt
    A = k.A
    TA = ThreadedCsr(A)
    if _threads is None:
        calibrate(TA)
    k.A, k.At, k.P = TA, ThreadedCsr(sp.csr_matrix(A.T)), ThreadedCsr(k.P)
    return ws
