"""BASELINE config C5: dual MAXCUT SDP on a random sparse graph |V| = 10k, host-side chordal
decomposition, clique PSD cones sharded over the available GPUs.  Prints one JSON line (rank 0)."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import cosmo_b200
from cosmo_b200 import chordal, sharding


def main():
    nv = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 200
    merge = sys.argv[3] if len(sys.argv) > 3 else "parent_child"
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    t0 = time.time()
    rows, cols, w = cosmo_b200.problems.banded_random_graph(nv, 3.0, 20, seed=1)
    P, q, A, b, sets = cosmo_b200.problems.maxcut_dual_sdp(nv, rows, cols, w)
    P2, q2, A2, b2, sets2, info = chordal.decompose(P, q, A, b, sets, merge=merge)
    graph_time = time.time() - t0
    cs = np.array(info.clique_sizes)
    st = cosmo_b200.Settings(scaling=0, adaptive_rho=False, max_iter=iters, eps_abs=0.0, eps_rel=0.0)
    sh = sharding.make_shard(P2, q2, A2, b2, sets2, rank, world)
    eng = sharding.create_engine(sh, st, device=local_rank, dist=dist)
    eng.solve()          # warm-up (same number of iterations)
    eng.reset()
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    out = eng.solve()
    dev = out.times["iter_time_device"]
    if dist is not None:
        t = torch.tensor([dev], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev = float(t.item())
    line = {"workload": "C5 MAXCUT dual SDP |V|=%d |E|=%d, chordal decomposition (%s): %d cliques, sizes min/median/max %d/%d/%d, "
                        "sum|c|^3=%.3g, n'=%d m'=%d" % (nv, len(rows), merge, len(cs), cs.min(), int(np.median(cs)), cs.max(),
                                                       float((cs.astype(float) ** 3).sum()), A2.shape[1], A2.shape[0]),
            "n_gpus": world, "iters": iters, "iter_per_s": iters / dev, "ms_per_iter": 1e3 * dev / iters,
            "cg_iters_per_admm_iter": out.kkt_inner_iterations / max(out.iter, 1), "graph_time_s": graph_time,
            "kernel_launches": int(out.kernel_launches)}
    if rank == 0 and world == 1 and "--cpu" in sys.argv:
        from oracle import cosmo_oracle as O
        from oracle.bridge import to_oracle_cones
        cones = to_oracle_cones(sets2)
        k = min(iters, 20)
        t0 = time.time()
        ref = O.solve(P2, q2, A2, b2, cones, O.Settings(kkt_solver="cg", scaling=0, adaptive_rho=False, max_iter=k, eps_abs=0.0, eps_rel=0.0))
        line["cpu_oracle_iter_per_s"] = k / (time.time() - t0)
        # parity at full size: the operator variable w after the same iterations on the identical arrays,
        # at 3 iterations (like C3 / C4) and at k; the CG tolerance of iteration j is 1 / j^1.5 relative to |rhs|, so one
        # inner iteration more or less on either side moves w by ~1e-6: the inner-iteration totals are reported with it
        # The reduced KKT matrix of this P = 0 problem makes CG lose orthogonality within ~20 iterations: the oracle with
        # exactly rounded inner products (math.fsum) instead of np.dot ends 1e-4 away from itself at equal iteration
        # counts (measured, profiles/c5_r2_1gpu.json), so the bound here is that self-sensitivity, not 1e-8.
        import math

        def cg_fsum(x, L, bvec, abstol, reltol=0.0, maxiter=None):
            nn = bvec.shape[0]
            maxiter = nn if maxiter is None else maxiter
            r = bvec - L(x)
            u = np.zeros(nn)
            residual, prev, it = math.sqrt(math.fsum(r * r)), 1.0, 0
            tol = max(reltol * residual, abstol)
            while it < maxiter and not (residual <= tol):
                u = r + (residual ** 2 / prev ** 2) * u
                c = L(u)
                alpha = residual ** 2 / math.fsum(u * c)
                x += alpha * u
                r -= alpha * c
                prev, residual, it = residual, math.sqrt(math.fsum(r * r)), it + 1
            return x, it, it + 1

        for kk in (3, k):
            stk = O.Settings(kkt_solver="cg", scaling=0, adaptive_rho=False, max_iter=kk, eps_abs=0.0, eps_rel=0.0)
            refk = ref if kk == k else O.solve(P2, q2, A2, b2, cones, stk)
            self_rel = None
            if kk == 3:
                orig = O.cg_solve
                O.cg_solve = cg_fsum
                try:
                    alt = O.solve(P2, q2, A2, b2, cones, stk)
                finally:
                    O.cg_solve = orig
                self_rel = float(np.max(np.abs(alt.w - refk.w)) / max(np.max(np.abs(refk.w)), 1e-300))
            eng.update_settings(cosmo_b200.Settings(scaling=0, adaptive_rho=False, max_iter=kk, eps_abs=0.0, eps_rel=0.0).to_struct())
            eng.reset()
            eng.warm_start(np.zeros(A2.shape[1]), np.zeros(A2.shape[0]), np.zeros(A2.shape[0]))
            o = eng.solve()
            rel = float(np.max(np.abs(eng.w() - refk.w)) / max(np.max(np.abs(refk.w)), 1e-300))
            line["parity_%d_iters" % kk] = {"w_rel": rel, "cg_total_engine": int(o.kkt_inner_iterations),
                                            "cg_total_oracle": int(np.sum(refk.kkt.inner_iterations)),
                                            "oracle_np_dot_vs_fsum_w_rel": self_rel,
                                            "ok": bool(int(o.kkt_inner_iterations) == int(np.sum(refk.kkt.inner_iterations)) and
                                                       rel <= max(1e-8, 2.0 * (self_rel or 1e-4)))}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
