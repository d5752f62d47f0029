"""Host mirror of the reference's model-building interface (no GPU): the cases of the reference's own
test/UnitTests/constraints.jl and test/UnitTests/interface.jl that do not need a solve, so that tests written
against COSMO.Model / Constraint / assemble! / set! read the same here (names, argument meaning, error behaviour)."""
import numpy as np
import pytest
import scipy.sparse as sp

import cosmo_b200
from cosmo_b200 import Constraint


def test_constraint_constructors_accept_the_reference_input_shapes():
    # constraints.jl:41-49: scalars, sparse, row vector with scalar b, column vector, dense matrix
    rng = np.random.default_rng(1872381)
    for A, b, rows, cols in ((4, 2, 1, 1), (4.0, 2.0, 1, 1), (np.uint64(4), np.uint64(2), 1, 1),
                             (sp.random(10, 2, 0.4, random_state=rng), rng.random(10), 10, 2),
                             (np.array([[1.0, 2, 3, 4]]), 1.0, 1, 4), (np.array([1.0, 2, 3, 4]), np.array([4.0, 3, 2, 1]), 4, 1),
                             (rng.random((10, 10)), sp.csc_matrix(rng.random((10, 1))).toarray(), 10, 10)):
        c = Constraint(A, b, cosmo_b200.ZeroSet)
        assert c.A.shape == (rows, cols) and c.b.shape == (rows,) and c.A.dtype == np.float64
        assert isinstance(c.convex_set, cosmo_b200.ZeroSet) and c.convex_set.dim == rows


def test_constraint_indices_embed_the_block():
    # constraints.jl:52-60: cs.A[:, 3:5] == A for dim = 10
    rng = np.random.default_rng(3)
    A, b = rng.random((3, 3)), rng.random(3)
    cs = Constraint(A, b, cosmo_b200.ZeroSet, 10, (3, 5))
    assert cs.A.shape == (3, 10) and np.array_equal(cs.A.toarray()[:, 2:5], A) and cs.A.nnz == 9
    with pytest.raises(ValueError):      # constraint.jl:67: increasing, positive range
        Constraint(A, b, cosmo_b200.ZeroSet, 10, (0, 2))
    with pytest.raises(ValueError):      # constraint.jl:68: dim >= stop
        Constraint(A, b, cosmo_b200.ZeroSet, 4, (3, 5))


def test_constraint_dimension_errors():
    with pytest.raises(ValueError, match="don't match"):          # constraint.jl:62
        Constraint(np.ones((3, 2)), np.ones(2), cosmo_b200.Nonnegatives)
    with pytest.raises(ValueError, match="row dimension"):        # constraint.jl:63
        Constraint(np.ones((3, 2)), np.ones(3), cosmo_b200.Nonnegatives(2))
    with pytest.raises(ValueError, match="Box"):                  # constraint.jl:91: argument cones need an object
        Constraint(np.ones((3, 2)), np.ones(3), cosmo_b200.Box)
    with pytest.raises(ValueError):                               # convexset.jl:826-830
        cosmo_b200.Box([0.0, 2.0], [1.0, 1.0])
    with pytest.raises(ValueError):                               # convexset.jl:631-634
        cosmo_b200.PowerCone(1.5)
    with pytest.raises(ValueError):                               # convexset.jl:275-279
        cosmo_b200.PsdCone(10)
    with pytest.raises(ValueError):
        cosmo_b200.PsdConeTriangle(7)


def test_assemble_merges_sorts_and_negates_like_the_reference():
    # constraints.jl:62-90 (merge_constraints!) + interface.jl:44-57 (model.p.A == -A, model.p.b == b)
    rng = np.random.default_rng(5)
    A1, b1, A2, b2 = rng.random((10, 4)), rng.random(10), rng.random((10, 4)), rng.random(10)
    soc = Constraint(rng.random((3, 4)), rng.random(3), cosmo_b200.SecondOrderCone)
    for S in (cosmo_b200.ZeroSet, cosmo_b200.Nonnegatives):
        model = cosmo_b200.Model()
        cosmo_b200.assemble(model, np.eye(4), np.ones(4), [Constraint(A1, b1, S), soc, Constraint(A2, b2, S)])
        assert [type(s) for s in model.sets0] == [S, cosmo_b200.SecondOrderCone] and model.sets0[0].dim == 20
        assert np.array_equal(model.A0.toarray()[:20], -np.vstack([A1, A2])) and np.array_equal(model.b0[:20], np.concatenate([b1, b2]))
        assert np.array_equal(model.A0.toarray()[20:], -soc.A.toarray())


@pytest.mark.parametrize("P,q", [(1.0, [1.0]), ([1.0], 1.0), (1.0, 1.0), ([1.0], [[1.0]])])
def test_assemble_accepts_scalar_and_vector_P_q(P, q):
    # interface.jl:63-85: P number / vector, q number / vector / matrix
    rng = np.random.default_rng(2)
    con = Constraint(rng.random((5, 1)), rng.random(5), cosmo_b200.Nonnegatives)
    model = cosmo_b200.Model()
    cosmo_b200.assemble(model, P, q, con)
    assert model.P0.toarray().tolist() == [[1.0]] and model.q0.tolist() == [1.0] and model.A0.shape == (5, 1)


def test_set_and_assemble_reject_inconsistent_dimensions():
    # interface.jl:28-37 (DimensionMismatch) and :88-95
    P = np.array([[4.0, 1], [1, 2]])
    q = np.array([1.0, 1])
    A = np.array([[1.0, 1], [1, 0], [0, 1]])
    Aa, b = np.vstack([A, -A]), np.concatenate([[1, 0.7, 0.7], [-1.0, 0, 0]])
    sets = [cosmo_b200.Nonnegatives(3), cosmo_b200.Nonnegatives(3)]
    model = cosmo_b200.Model()
    model.set(P, q, Aa, b, sets)
    assert model.m == 6 and model.n == 2
    for args in ((P, np.ones(3), Aa, b), (np.zeros((1, 1)), q, Aa, b), (P, q, np.array([[1.0, 2], [1, 2]]), b), (P, q, Aa, np.array([1.0, 2]))):
        with pytest.raises(ValueError):
            cosmo_b200.Model().set(*args, sets)
    with pytest.raises(ValueError, match="inconsistent"):
        cosmo_b200.assemble(cosmo_b200.Model(), sp.identity(2), np.ones(2), [Constraint(1.0, 0.0, cosmo_b200.Nonnegatives)])
    with pytest.raises(RuntimeError):                              # optimize! before assemble!
        cosmo_b200.Model().optimize()


def test_warm_start_and_update_argument_checks():
    # interface.jl:117-211
    P, q = np.eye(2), np.ones(2)
    model = cosmo_b200.Model()
    with pytest.raises(RuntimeError):
        model.update(q=[1.0, 2.0])
    cosmo_b200.assemble(model, P, q, Constraint(np.eye(2), np.zeros(2), cosmo_b200.Nonnegatives))
    model.warm_start_primal([1.0, 2.0])
    assert np.array_equal(model.x, [1.0, 2.0]) and np.array_equal(model.s, model.b0 - model.A0 @ model.x)
    model.warm_start_dual([0.5, 0.25])
    assert np.array_equal(model.mu, [-0.5, -0.25])
    for bad in ([1.0], [1.0, 2.0, 3.0]):
        with pytest.raises(ValueError):
            model.warm_start_primal(bad)
        with pytest.raises(ValueError):
            model.warm_start_dual(bad)
        with pytest.raises(ValueError):
            model.update(q=bad)
        with pytest.raises(ValueError):
            model.update(b=bad)
    model.update(q=[2.0, 3.0], b=[1.0, 1.0])
    assert np.array_equal(model.q0, [2.0, 3.0]) and np.array_equal(model.b0, [1.0, 1.0])


def test_settings_reject_what_the_engine_does_not_implement():
    with pytest.raises(cosmo_b200.EngineError):                   # direct LDL' factorisations are CPU plugins
        cosmo_b200.Settings(kkt_solver="QdldlKKTSolver").to_struct()
    with pytest.raises(cosmo_b200.EngineError):
        cosmo_b200.Settings(accelerator="AndersonAccelerator{Type1}").to_struct()
    st = cosmo_b200.Settings(kkt_solver="MINRESIndirectKKTSolver").to_struct()
    assert st.kkt_solver == cosmo_b200.engine.KKT_MINRES
