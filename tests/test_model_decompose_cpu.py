"""CPU test of the host glue around the engine for `Settings(decompose=True)`: Model.optimize() must
decompose -> hand the augmented problem to the engine -> reverse (and complete the dual).  The CUDA engine is
replaced by a stand-in that solves the problem it is given with the CPU oracle, so only the glue is under test
(the engine itself is covered by the GPU tests)."""
import numpy as np
import pytest
import scipy.sparse as sp

import cosmo_b200
from cosmo_b200 import chordal, engine as E, model as M
from oracle import cosmo_oracle as O
from tests import golden_problems as G

_CODE = {E.ZERO: O.ZeroSet, E.NONNEG: O.Nonnegatives, E.SOC: O.SecondOrderCone, E.PSD_SQUARE: O.PsdCone,
         E.PSD_TRIANGLE: O.PsdConeTriangle}


class _OracleEngine:
    """stand-in with the call surface Model uses: ctor, update_settings, warm_start, update_qb, solve, close"""
    instances = []

    def __init__(self, P, q, A, b, sets, settings, D=None, E=None, c=1.0, dtype=np.float64, device=0, equilibrate=False):
        assert D is None and E is None and not equilibrate         # the glue test runs with scaling = 0
        self.P, self.q, self.A, self.b = sp.csc_matrix(P), np.array(q), sp.csc_matrix(A), np.array(b)
        self.cones = [O.Box(t[2], t[3]) if t[0] == E_BOX else _CODE[t[0]](t[1]) for t in sets]
        self.st = settings
        self.ws = None
        _OracleEngine.instances.append(self)

    def update_settings(self, st):
        self.st = st

    def warm_start(self, x, s, mu):
        self._warm = (np.array(x), np.array(s), np.array(mu))

    def update_qb(self, q, b):
        raise AssertionError("decomposed models re-decompose instead of updating b in place")

    def solve(self):
        st = O.Settings(scaling=0, eps_abs=self.st.eps_abs, eps_rel=self.st.eps_rel, max_iter=self.st.max_iter)
        r = O.solve(self.P, self.q, self.A, self.b, self.cones, st)
        out = E.SolveOutput()
        out.x, out.s, out.mu = r.x, r.s, -r.y
        out.obj_val, out.iter, out.safeguarding_iter, out.status = r.obj_val, r.iter, 0, r.status
        out.r_prim, out.r_dual, out.max_norm_prim, out.max_norm_dual = r.info.r_prim, r.info.r_dual, 0.0, 0.0
        out.rho, out.rho_updates, out.times = 0.1, [0.1], {"iter_time_device": 0.0}
        out.kkt_inner_iterations = out.kkt_multiplications = out.kernel_launches = 0
        return out

    def close(self):
        pass


E_BOX = E.BOX


@pytest.fixture
def oracle_engine(monkeypatch):
    _OracleEngine.instances.clear()
    monkeypatch.setattr(M._eng, "Engine", _OracleEngine)
    return _OracleEngine


def _g6_model(**kw):
    P, q, cons = G.g6_chordal_sdp()
    model = cosmo_b200.Model()
    mine = [cosmo_b200.Constraint(c.A, c.b, cosmo_b200.PsdConeTriangle(c.convex_set.dim)) for c in cons]
    cosmo_b200.assemble(model, P, q, mine, cosmo_b200.Settings(scaling=0, eps_abs=1e-7, eps_rel=1e-7, **kw))
    return model


@pytest.mark.parametrize("merge", ["NoMerge", "ParentChildMerge", "CliqueGraphMerge"])
def test_model_decomposes_solves_and_reverses(oracle_engine, merge):
    plain = _g6_model().optimize()
    assert oracle_engine.instances[-1].A.shape == (45, 2)                       # undecomposed: one 9x9 cone
    model = _g6_model(decompose=True, merge_strategy=merge, complete_dual=True)
    res = model.optimize()
    eng = oracle_engine.instances[-1]
    if merge == "NoMerge":   # docs/src/decomposition.md:43: five cliques, sizes 3, 2, 4, 3, 4
        assert sorted(c.dim for c in eng.cones) == sorted([6, 3, 10, 6, 10])
    assert eng.A.shape[0] == sum(c.dim for c in eng.cones) and eng.A.shape[1] == 2 + model._dec.num_overlaps
    assert res.status == "Solved" == plain.status and abs(res.obj_val - plain.obj_val) < 1e-4
    assert res.x.shape == (2,) and res.s.shape == (45,) and res.y.shape == (45,)
    assert np.allclose(res.x, plain.x, atol=1e-3) and np.allclose(res.s, plain.s, atol=1e-3)
    Y = chordal._svec_to_mat(res.y, 9)                                           # completed dual: PSD
    assert np.linalg.eigvalsh(Y).min() > -1e-4
    # a second optimize! reuses the engine and warm-starts it in the decomposed space
    n_eng = len(oracle_engine.instances)
    res2 = model.optimize()
    assert len(oracle_engine.instances) == n_eng and res2.status == "Solved"
    assert eng._warm[0].shape == (eng.A.shape[1],) and eng._warm[1].shape == (eng.A.shape[0],)
    assert np.any(eng._warm[1] != 0)                                             # ... from the previous decomposed iterates
    # a warm start in the original coordinates restarts the decomposed iterates from x0 (clique copies of s, mu: zero)
    model.warm_start_primal(np.array([0.25, -0.5]))
    assert model.optimize().status == "Solved" and len(oracle_engine.instances) == n_eng
    assert np.array_equal(eng._warm[0][:2], [0.25, -0.5]) and not np.any(eng._warm[0][2:]) and not np.any(eng._warm[1]) and not np.any(eng._warm[2])
    # update!(b) drops the engine: the clique row map is rebuilt
    model.update(b=model.b0 * 1.0)
    assert model.engine is None
    assert model.optimize().status == "Solved" and len(oracle_engine.instances) == n_eng + 1


def test_model_without_decomposable_cones_is_untouched(oracle_engine):
    P, q, cons = G.g1_qp_nonneg()
    model = cosmo_b200.Model()
    cosmo_b200.assemble(model, P, q, [cosmo_b200.Constraint(c.A, c.b, cosmo_b200.Nonnegatives(c.convex_set.dim)) for c in cons],
                        cosmo_b200.Settings(scaling=0, decompose=True))
    res = model.optimize()
    assert model._dec is None and res.status == "Solved" and np.max(np.abs(res.x - G.G1_X)) < 1e-3
