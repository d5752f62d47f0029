"""Regenerates the committed golden fixtures of this directory.

    python tests/golden/make_golden.py

* ``reference_known_answers.json`` -- the literal expectations of the reference's own tests / examples for
  the hot path (status, objective, tolerance, file:line), as data.  Julia is not available here, so these
  literals (not outputs of a run of the reference) are what pins the CPU oracle.
* ``oracle_iterates.npz`` -- outputs of the pinned CPU oracle (oracle/cosmo_oracle.py) on small seeded
  problems: the operator variable w after k ADMM iterations and the final (x, s, y), for plain and for
  accelerated runs.  The GPU tests compare the engine with these arrays, the CPU tests check that the oracle
  still reproduces them (so an accidental change of the oracle is noticed).
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import cosmo_b200  # noqa: E402  (problem generators only; no GPU needed)
from oracle import cosmo_oracle as O  # noqa: E402
from oracle.bridge import to_oracle_cones
from tests import golden_problems as G  # noqa: E402

KNOWN = [
    dict(id="G1", problem="g1_qp_nonneg", status="Solved", obj=1.88, x=[0.3, 0.7], atol=1e-3, ref="examples/qp.jl:13-44; test/UnitTests/simple.jl:21-47"),
    dict(id="G1b", problem="g1_qp_box", status="Solved", obj=1.88, x=[0.3, 0.7], atol=1e-3, ref="examples/qp.jl:31-44"),
    dict(id="G2", problem="g2_box_feasible", status="Solved", obj=-0.5, atol=1e-5, ref="test/UnitTests/qp-box.jl:15-32"),
    dict(id="G2i1", problem="g2_box_primal_infeasible_1", status="Primal_infeasible", ref="qp-box.jl:35-52"),
    dict(id="G2i2", problem="g2_box_primal_infeasible_2", status="Primal_infeasible", ref="qp-box.jl:54-71"),
    dict(id="G2d", problem="g2_box_dual_infeasible", status="Dual_infeasible", settings=dict(check_infeasibility=20, scaling=0), ref="qp-box.jl:73-106"),
    dict(id="G3", problem="g3_hs21", status="Solved", obj=0.04, x=[-2.0, 0.0], atol=1e-3, ref="test/UnitTests/moi_wrapper.jl:219-276 (obj -99.96 with r = -100)"),
    dict(id="G12", problem="g12_lp", status="Solved", obj=20.0, x=[3.0, 5.0, 1.0, 1.0], atol=1e-2, ref="examples/lp.jl:17-46"),
    dict(id="G13", problem="g13_lovasz_petersen", status="Solved", obj=-4.0, atol=1e-3, ref="examples/lovasz_petersen.jl:22-60"),
    dict(id="G11", problem="g11_iteration_limit", status="Max_iter_reached", settings=dict(max_iter=2), rho_updates=[0.1], ref="moi_wrapper.jl:201-217"),
] + [dict(id="G15/16:" + name, problem=b.__name__, status=status, obj=obj, atol=atol, settings=kw, ref="test/UnitTests/exp_cone.jl, pow_cone.jl")
     for name, b, status, obj, atol, kw in G.G15_G16]


def _run(P, q, A, b, cones, **kw):
    return O.solve(P, q, A, b, cones, O.Settings(kkt_solver="cg", **kw))


def iterates():
    out = {}
    P, q, A, b, sets = cosmo_b200.problems.random_sparse_qp(40, 70, 0.15, seed=7)
    cones = to_oracle_cones(sets)
    for acc in ("empty", "anderson"):
        for scaling in (0, 10):
            for iters in (5, 14, 33):
                r = _run(P, q, A, b, cones, scaling=scaling, max_iter=iters, eps_abs=1e-14, eps_rel=1e-14, accelerator=acc)
                key = "qp40x70_seed7/%s/scaling%d/it%d" % (acc, scaling, iters)
                out[key + "/w"] = r.w
                out[key + "/iter_sg"] = np.array([r.iter, r.safeguarding_iter])
    for name, builder, kw in (("g1_qp_nonneg", G.g1_qp_nonneg, dict(scaling=0)), ("g13_lovasz_petersen", G.g13_lovasz_petersen, dict(eps_abs=1e-6, eps_rel=1e-6)),
                              ("g15_exp_feasible", G.g15_exp_feasible, dict(eps_abs=1e-4, eps_rel=1e-4))):
        Pb, qb, cons = builder()
        Pm, qm, Am, bm, cn = O.assemble(Pb, qb, cons)
        r = _run(Pm, qm, Am, bm, cn, **kw)
        out[name + "/x"], out[name + "/s"], out[name + "/y"] = r.x, r.s, r.y
        out[name + "/iter_obj"] = np.array([r.iter, r.obj_val])
    return out


def main():
    with open(os.path.join(HERE, "reference_known_answers.json"), "w") as f:
        json.dump(KNOWN, f, indent=1)
    np.savez_compressed(os.path.join(HERE, "oracle_iterates.npz"), **iterates())


if __name__ == "__main__":
    main()
