"""CPU tests of the N>1 host logic: row/cone partitioning and the
exchange pattern of the sharded reduced-KKT operator, run with world_size 2 over
`gloo` (no GPU).  The GPU path issues the same reductions through NCCL inside
libcosmo_b200.so."""
import os
import socket

import numpy as np
import pytest
import scipy.sparse as sp

import cosmo_b200
from cosmo_b200 import sharding


def _mixed_problem(seed=1):
    P, q, A, b, sets = cosmo_b200.problems.random_sparse_qp(120, 401, 0.08, seed=seed)
    box = sets[1]
    sets = [cosmo_b200.ZeroSet(0), sets[0],
            cosmo_b200.Box(box.l[:99], box.u[:99]), cosmo_b200.SecondOrderCone(60),
            cosmo_b200.PsdConeTriangle(36), cosmo_b200.ExponentialCone(), cosmo_b200.DualPowerCone(0.4),
            cosmo_b200.Box(box.l[99:99], box.u[99:99])]
    assert sum(S.dim for S in sets) == A.shape[0]
    return P, q, A, b, sets


@pytest.mark.parametrize("world", [1, 2, 3, 4, 8])
def test_partition_covers_rows_and_keeps_cones_whole(world):
    P, q, A, b, sets = _mixed_problem()
    rows = []
    for r in range(world):
        sh = sharding.make_shard(P, q, A, b, sets, r, world)
        assert sum(S.dim for S in sh.sets) == sh.A.shape[0] == len(sh.rows) == len(sh.b)
        assert np.array_equal(sh.A.toarray(), A.toarray()[sh.rows])
        assert np.array_equal(sh.b, b[sh.rows])
        for S in sh.sets:   # SOC / PSD cones are never split
            if isinstance(S, cosmo_b200.SecondOrderCone):
                assert S.dim == 60
            if isinstance(S, cosmo_b200.PsdConeTriangle):
                assert S.dim == 36
            if isinstance(S, (cosmo_b200.ExponentialCone, cosmo_b200.PowerCone)):
                assert S.dim == 3
        rows.append(sh.rows)
    assert np.array_equal(np.concatenate(rows), np.arange(A.shape[0]))  # contiguous, in cone order


def test_partition_balances_nnz():
    P, q, A, b, sets = cosmo_b200.problems.random_sparse_qp(500, 4000, 0.05, seed=3)
    nnz = [sharding.make_shard(P, q, A, b, sets, r, 4).A.nnz for r in range(4)]
    assert max(nnz) <= 1.1 * (A.nnz / 4)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import cosmo_oracle as O
        from oracle.bridge import to_oracle_cones
        P, q, A, b, sets = _mixed_problem(seed=5)
        sh = sharding.make_shard(P, q, A, b, sets, rank, world)
        n = A.shape[1]
        rng = np.random.default_rng(0)          # replicated n-vector, identical on every rank
        u = rng.standard_normal(n)
        rho_full = rng.uniform(0.05, 2.0, A.shape[0])
        rho = rho_full[sh.rows]
        sigma = 1e-6
        # rank-local part of the reduced operator: A_g'(rho_g .* (A_g u)); rank 0 adds P u + sigma u;
        # the partial dot u'c_g rides in the same buffer (what the engine appends at cb[n])
        c = sh.A.T @ (rho * (sh.A @ u))
        if rank == 0:
            c = c + P @ u + sigma * u
        buf = torch.from_numpy(np.concatenate([c, [u @ c]]))
        dist.all_reduce(buf, op=dist.ReduceOp.SUM)          # the one collective per operator application
        full = A.T @ (rho_full * (A @ u)) + P @ u + sigma * u
        ok = np.allclose(buf[:-1].numpy(), full, rtol=1e-12, atol=1e-12) and np.isclose(buf[-1].item(), u @ full, rtol=1e-12)
        # residual norms: row-local maxima + one allreduce(max)
        x, s_full = rng.standard_normal(n), rng.standard_normal(A.shape[0])
        loc = np.max(np.abs(sh.A @ x + s_full[sh.rows] - sh.b)) if len(sh.rows) else 0.0
        t = torch.tensor([loc], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ok = ok and np.isclose(t.item(), np.max(np.abs(A @ x + s_full - b)), rtol=1e-13)
        # projection is rank-local: cones are whole
        w = rng.standard_normal(A.shape[0])
        ref = w.copy()
        O.project(ref, to_oracle_cones(sets))
        mine = w[sh.rows].copy()
        O.project(mine, to_oracle_cones(sh.sets))
        ok = ok and np.allclose(mine, ref[sh.rows], rtol=0, atol=1e-14)
        # Anderson acceleration over sharded rows (aa.cuh / Engine::aa_update, aa_accelerate): every rank keeps
        # [w_x (replicated); w_s restricted to its rows]; inner products run over [lo, dim) with lo = 0 on rank 0
        # and lo = n elsewhere, followed by an allreduce -> the same R, eta and candidate as the unsharded method
        m_full = A.shape[0]
        dim_full = n + m_full
        idx = np.concatenate([np.arange(n), n + sh.rows])          # this rank's slice of the operator variable
        lo = 0 if rank == 0 else n

        def gdot(a, b_):
            t_ = torch.tensor([float(a[lo:] @ b_[lo:])], dtype=torch.float64)
            dist.all_reduce(t_, op=dist.ReduceOp.SUM)
            return t_.item()

        mem = 5
        aa = O.AndersonAccelerator(dim_full, mem)                   # the unsharded restatement, on full vectors
        Gm, Qm, Rm = np.zeros((len(idx), mem)), np.zeros((len(idx), mem)), np.zeros((mem, mem))
        g_last = f_last = None
        it = 0
        Mop = rng.standard_normal((dim_full, dim_full)) * (0.5 / np.sqrt(dim_full))
        cst = rng.standard_normal(dim_full)
        xk = rng.standard_normal(dim_full)
        for k in range(2 * mem + 3):                                # crosses one memory restart
            gk = Mop @ xk + cst
            aa.update(gk, xk, k + 2)
            g_acc = gk.copy()
            aa.accelerate(g_acc, xk, k + 2)
            # ---- sharded emulation of the same step ----
            gl, xl = gk[idx], xk[idx]
            fl = xl - gl
            cand = gl.copy()
            if g_last is not None:
                j = it % mem
                if j == 0 and it != 0:
                    it = 0
                Gm[:, j] = gl - g_last
                qv = fl - f_last
                for i in range(j):
                    Rm[i, j] = gdot(Qm[:, i], qv)
                    qv = qv - Rm[i, j] * Qm[:, i]
                Rm[j, j] = np.sqrt(gdot(qv, qv))
                Qm[:, j] = qv / Rm[j, j]
                it += 1
                l = min(it, mem)
                if l >= 3:
                    eta = np.array([gdot(Qm[:, c_], fl) for c_ in range(l)])
                    eta = np.linalg.solve(np.triu(Rm[:l, :l]), eta)
                    if np.linalg.norm(eta) <= 1e4:
                        cand = gl - Gm[:, :l] @ eta
            g_last, f_last = gl.copy(), fl.copy()
            ok = ok and np.allclose(cand, g_acc[idx], rtol=1e-9, atol=1e-9)
            xk = g_acc                                               # continue from the accelerated point
        ok = ok and aa.num_accelerated_steps >= mem
        # the 128-byte ncclUniqueId travels as a python object over the plumbing backend
        obj = [bytes(range(128)) if rank == 0 else None]
        dist.broadcast_object_list(obj, src=0)
        ok = ok and obj[0] == bytes(range(128))
        out[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def test_sharded_operator_matches_full_gloo_world2():
    import torch.multiprocessing as mp
    world = 2
    port = _free_port()
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
        assert dict(out) == {0: True, 1: True}
