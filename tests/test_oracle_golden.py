"""Pin the CPU oracle against the reference's own literal known answers
(G1, G2, G3, G8, G12, G13, G14) before it is trusted as the checker."""
import numpy as np
import pytest
import scipy.sparse as sp

from oracle import cosmo_oracle as O
from tests import golden_problems as G


def _solve(builder, **kw):
    P, q, cons = builder()
    Pm, qm, A, b, cones = O.assemble(P, q, cons)
    return O.solve(Pm, qm, A, b, cones, O.Settings(**kw)), cones


@pytest.mark.parametrize("builder", [G.g1_qp_nonneg, G.g1_qp_box])
@pytest.mark.parametrize("kkt", ["direct", "cg", "minres", "minres_reduced"])
@pytest.mark.parametrize("scaling", [0, 10])
def test_g1_simple_qp(builder, kkt, scaling):
    # examples/qp.jl:41-44, test/UnitTests/simple.jl:21-47 (tol 1e-3)
    res, _ = _solve(builder, kkt_solver=kkt, scaling=scaling)
    if kkt in ("direct", "cg"):
        assert res.status == "Solved"
    else:
        # MINRES: the reference passes abstol = tol_k / |initial residual|
        # (kktsolver_indirect.jl:72-73,151-152), which *loosens* as the initial
        # residual shrinks; the restated recurrence therefore plateaus near 3e-4
        # and exits with Max_iter_reached.  The reference has no active test for
        # this path (kktsolver.jl:7-8) -> parity unpinned; only accuracy is asserted.
        assert res.status in ("Solved", "Max_iter_reached")
    assert np.max(np.abs(res.x - G.G1_X)) < 1e-3
    assert abs(res.obj_val - G.G1_OBJ) < 1e-3


def test_g1_matches_survey_iteration_count():
    # 375 iterations at eps 1e-5 with scaling=0, fixed rho, dense KKT
    res, _ = _solve(G.g1_qp_nonneg, scaling=0, adaptive_rho=False)
    assert res.status == "Solved" and res.iter == 375


def test_g2_box():
    res, _ = _solve(G.g2_box_feasible)
    assert res.status == "Solved"
    assert abs(res.obj_val - (-0.5)) < 1e-5  # qp-box.jl:31
    res, _ = _solve(G.g2_box_primal_infeasible_1)
    assert res.status == "Primal_infeasible"  # qp-box.jl:51
    res, _ = _solve(G.g2_box_primal_infeasible_2)
    assert res.status == "Primal_infeasible"  # qp-box.jl:70
    res, _ = _solve(G.g2_box_dual_infeasible, check_infeasibility=20, scaling=0)
    assert res.status == "Dual_infeasible"    # qp-box.jl:88
    res, _ = _solve(G.g2_box_dual_infeasible, check_infeasibility=40, scaling=10)
    assert res.status == "Dual_infeasible"    # qp-box.jl:105


def test_g3_hs21_set_merging():
    res, cones = _solve(G.g3_hs21)
    # moi_wrapper.jl:266-271 (native assemble keeps the two Box sets separate)
    kinds = [type(c).__name__ for c in cones]
    assert kinds == ["ZeroSet", "Nonnegatives", "Box", "Box", "SecondOrderCone"]
    assert cones[0].dim == 3 and cones[1].dim == 3
    assert res.status == "Solved"
    assert abs(res.obj_val - G.G3_OBJ) < 1e-3
    assert np.max(np.abs(res.x - G.G3_X)) < 1e-3


@pytest.mark.parametrize("kkt", ["direct", "cg"])
def test_g12_lp(kkt):
    res, _ = _solve(G.g12_lp, eps_abs=1e-4, eps_rel=1e-5, kkt_solver=kkt)
    assert res.status == "Solved"
    assert np.max(np.abs(res.x - G.G12_X)) < 1e-2  # lp.jl:44
    assert abs(res.obj_val - G.G12_OBJ) < 1e-2     # lp.jl:46


@pytest.mark.parametrize("scaling", [0, 10])
def test_g13_lovasz_petersen(scaling):
    res, _ = _solve(G.g13_lovasz_petersen, scaling=scaling, eps_abs=1e-6, eps_rel=1e-6)
    assert res.status == "Solved"
    assert abs(res.obj_val - G.G13_OBJ) < 1e-3


def test_g14_model_updates():
    # model_modifications.jl:20-31: second optimize! warm-starts from the first
    P, q, cons = G.g1_qp_nonneg()
    Pm, qm, A, b, cones = O.assemble(P, q, cons)
    ws = O.Workspace(Pm, qm, A, b, cones, O.Settings(check_termination=1))
    r1 = ws.optimize()
    r2 = ws.optimize()
    assert abs(r1.obj_val - r2.obj_val) <= 1e-3 and r2.iter <= r1.iter
    # :33-43 update!(q=[2,3]) -> obj 3.5, x=[0.5,0.5]
    ws = O.Workspace(Pm, qm, A, b, cones, O.Settings())
    ws.optimize()
    ws.update(q=np.array([2.0, 3.0]))
    r = ws.optimize()
    assert abs(r.obj_val - 3.5) < 1e-3 and np.linalg.norm(r.x - [0.5, 0.5]) < 1e-3
    # :45-61 LP, update!(b=[0,1]) -> x=[0,-1]
    Pm, qm, A, b, cones = O.assemble(np.zeros((2, 2)), np.array([1.0, 1.0]),
                                     [O.Constraint(np.eye(2), np.array([-2.0, -3.0]), O.Nonnegatives(2))])
    ws = O.Workspace(Pm, qm, A, b, cones, O.Settings(check_termination=20))
    r = ws.optimize()
    assert np.linalg.norm(r.x - [2.0, 3.0]) < 1e-3
    ws.update(b=np.array([0.0, 1.0]))
    r2 = ws.optimize()
    assert np.linalg.norm(r2.x - [0.0, -1.0]) < 1e-4


def test_g8_algebra_kats():
    # test/UnitTests/algebra.jl:31-60
    E = np.array([1.0, 2.0, 3.0])
    v = np.array([-1.0, 5.0, 4.0])
    assert O.scaled_norm(E, v, 1) == 23
    assert O.scaled_norm(E, v, 2) == np.linalg.norm([-1.0, 10.0, 12.0])
    assert O.scaled_norm(E, v, np.inf) == 12
    with pytest.raises(ValueError):
        O.scaled_norm(E, v, 3)
    A = sp.csc_matrix(np.array([[1.0, 2, 3, 4], [2, -1, 30, 4.1]]))
    B = sp.csc_matrix(np.array([[1.0, 3, -2, 4.2], [-100, -1, -2, -100]]))
    vv = O.col_norms(A)
    assert np.array_equal(vv, [2, 2, 30, 4.1])
    O.col_norms(B, vv)
    assert np.array_equal(vv, [100, 3, 30, 100])
    vv = O.row_norms(A.T)
    O.row_norms(B.T, vv)
    assert np.array_equal(vv, [100, 3, 30, 100])
    x = np.linspace(-5, 5, 50); x[9] = -8; x[19] = 8
    xc = O.clip(x, -2, 2)
    assert xc.max() <= 2 and xc.min() >= -2
    xc2 = O.clip(x, -2, 2, -100, 100)
    assert xc2[9] == -100 and xc2[19] == 100


def test_g9_cone_membership_and_projection():
    # test/UnitTests/sets.jl:27-112 — projections land in the cone (property tests)
    rng = np.random.default_rng(0)
    x = rng.standard_normal(10)
    O.project_cone(x, O.SecondOrderCone(10))
    assert np.linalg.norm(x[1:]) <= x[0] + 1e-12
    X = rng.standard_normal((4, 4)); X = X + X.T
    xs = X.reshape(-1, order="F").copy()
    O.project_cone(xs, O.PsdCone(16))
    assert np.linalg.eigvalsh(xs.reshape(4, 4, order="F")).min() >= -1e-9
    xt = O.extract_upper_triangle(X, np.sqrt(2.0))
    O.project_cone(xt, O.PsdConeTriangle(10))
    Xp = O.populate_upper_triangle(xt, 4, 1 / np.sqrt(2.0))
    Xp = Xp + np.triu(Xp, 1).T
    assert np.linalg.eigvalsh(Xp).min() >= -1e-9
    # triangle and square projections agree
    assert np.allclose(Xp, xs.reshape(4, 4, order="F"), atol=1e-12)
    # exact eigen-clip reference
    w, V = np.linalg.eigh(X)
    assert np.allclose(Xp, (V * np.maximum(w, 0)) @ V.T, atol=1e-12)
    one = np.array([-3.0]); O.project_cone(one, O.PsdConeTriangle(1)); assert one[0] == 0.0


def test_kkt_indirect_matches_direct():
    # the reference's (disabled) test kktsolver.jl:96-125: CG/MINRES vs dense K\b, before/after update_rho!
    rng = np.random.default_rng(1)
    n, m = 10, 15
    A = sp.random(m, n, density=0.4, random_state=2, format="csc")
    Pd = sp.random(n, n, density=0.4, random_state=3); P = sp.csc_matrix(Pd @ Pd.T)
    rho = 0.1 * np.ones(m)
    st = O.Settings()
    rhs = rng.standard_normal(n + m)
    for name in ["cg", "minres_reduced", "minres"]:
        S = O.make_kkt_solver(name, P, A, 1e-6, rho.copy(), st)
        S.iteration_counter = 10 ** 6  # tight tolerance
        D = O.DirectKKT(P, A, 1e-6, rho)
        assert np.allclose(S.solve(rhs), D.solve(rhs), atol=1e-5), name
        rho2 = rng.uniform(0.05, 0.5, m)
        S.update_rho(rho2); D.update_rho(rho2)
        for _ in range(3):  # maxiter = size(A,2) per call; warm start carries over
            S.iteration_counter = 10 ** 6
            sol = S.solve(rhs)
        assert np.allclose(sol, D.solve(rhs), atol=1e-5), name


@pytest.mark.parametrize("name,builder,status,obj,atol,kw", G.G15_G16, ids=[g[0] for g in G.G15_G16])
def test_g15_g16_exp_pow_cone_problems(name, builder, status, obj, atol, kw):
    # test/UnitTests/exp_cone.jl, pow_cone.jl: statuses and objective values with the reference's tolerances
    res, _ = _solve(builder, **kw)
    assert res.status == status
    if obj is not None:
        assert abs(res.obj_val - obj) < atol


def test_exp_pow_projections_land_in_cone():
    # test/UnitTests/sets.jl:84-110: 100 random points in [-25, 25]^3, projection in the cone (tol 1e-4);
    # plus the projection's optimality conditions (v - Pi v in the polar cone, orthogonal to Pi v)
    rng = np.random.default_rng(7)
    for i in range(100):
        alpha = 0.1 + 0.85 * rng.random()
        for cone in (O.ExponentialCone(), O.PowerCone(alpha), O.DualExponentialCone(), O.DualPowerCone(alpha)):
            v0 = -25.0 + 50.0 * rng.random(3)
            v = v0.copy()
            O.project_cone(v, cone)
            assert O.in_cone(v, cone, 1e-4)
            d = v - v0                       # must lie in the dual cone, orthogonal to v
            assert O.in_dual(d, cone, 1e-3)
            assert abs(v @ d) < 1e-5 * (1.0 + v0 @ v0)
            w = v.copy()
            O.project_cone(w, cone)          # idempotence (up to the iteration tolerances)
            assert np.max(np.abs(w - v)) < 1e-3 * (1.0 + np.max(np.abs(v)))


# ---------------------------------------------------------------------------
# Accelerated runs (the reference's default: Anderson type-II, QR, restarted memory 15, safeguarded).
# COSMOAccelerators.jl is not part of the reference tree -> iterate-level parity unpinned; what the
# reference's own tests pin is behaviour, reproduced here.
# ---------------------------------------------------------------------------
_AA = dict(accelerator="anderson")


@pytest.mark.parametrize("builder,x,obj,tol", [(G.g1_qp_nonneg, G.G1_X, G.G1_OBJ, 1e-3), (G.g1_qp_box, G.G1_X, G.G1_OBJ, 1e-3),
                                               (G.g12_lp, G.G12_X, G.G12_OBJ, 1e-2), (G.g3_hs21, G.G3_X, G.G3_OBJ, 1e-3),
                                               (G.g13_lovasz_petersen, None, G.G13_OBJ, 1e-3)])
def test_accelerated_runs_reach_the_reference_answers(builder, x, obj, tol):
    # simple.jl:21-47, examples/lp.jl, moi_wrapper.jl:219-276, lovasz_petersen.jl run with default settings
    # (accelerator on) in the reference; test/UnitTests/AccelerationTests/anderson_accelerator.jl asks :Solved
    plain, _ = _solve(builder)
    res, _ = _solve(builder, **_AA)
    assert res.status == "Solved" and abs(res.obj_val - obj) < tol
    if x is not None:
        assert np.max(np.abs(res.x - x)) < tol
    assert res.iter <= plain.iter + 1      # never slower than plain ADMM on these, usually 2-6x faster
    assert res.iter == (res.iter - res.safeguarding_iter) + res.safeguarding_iter and res.safeguarding_iter >= 0


def test_accelerated_statuses_of_infeasible_problems():
    assert _solve(G.g2_box_primal_infeasible_1, **_AA)[0].status == "Primal_infeasible"
    assert _solve(G.g2_box_primal_infeasible_2, **_AA)[0].status == "Primal_infeasible"
    assert _solve(G.g2_box_dual_infeasible, check_infeasibility=20, scaling=0, **_AA)[0].status == "Dual_infeasible"
    for name, builder, status, obj, atol, kw in G.G15_G16:
        res, _ = _solve(builder, **kw, **_AA)
        assert res.status == status, name
        if obj is not None:
            assert abs(res.obj_val - obj) < atol, name


def test_accelerator_restarts_and_max_rho_adaptions():
    P, q, cons = G.g1_qp_nonneg()
    Pm, qm, A, b, cones = O.assemble(P, q, cons)
    # AccelerationTests/max_rho_adaption.jl:21-36 (default accelerator): exactly 2, then exactly 1 adaption
    ws = O.Workspace(Pm, qm, A, b, cones, O.Settings(adaptive_rho_interval=25, adaptive_rho_max_adaptions=2, rho=1e-6,
                                                     eps_abs=1e-6, eps_rel=1e-4, **_AA))
    ws.optimize()
    assert len(ws.rho_updates) - 1 == 2
    ws = O.Workspace(Pm, qm, A, b, cones, O.Settings(adaptive_rho_interval=25, adaptive_rho_max_adaptions=1, rho=1e-6,
                                                     eps_abs=1e-4, eps_rel=1e-4, **_AA))
    ws.optimize()
    assert len(ws.rho_updates) - 1 == 1
    # AccelerationTests/adaptive_rho_acc_restarts.jl:19-25: one accelerator restart per rho adaption
    ws = O.Workspace(Pm, qm, A, b, cones, O.Settings(adaptive_rho_interval=23, rho=1e-4, safeguard=False,
                                                     accelerator_mem=5, **_AA))
    res = ws.optimize()
    assert res.status == "Solved"
    assert sum(1 for e in ws.accelerator.log if e[1] == "rho_adapted") == len(ws.rho_updates) - 1 >= 1


def test_anderson_qr_solves_the_least_squares_problem():
    # the updated QR factorisation reproduces numpy's least squares solution of min |f - F eta|
    rng = np.random.default_rng(3)
    dim, mem = 40, 6
    aa = O.AndersonAccelerator(dim, mem)
    M = rng.standard_normal((dim, dim)) * 0.1
    x = rng.standard_normal(dim)
    fs, gs = [], []
    for k in range(mem + 1):
        g = M @ x + 1.0
        aa.update(g, x, k + 2)
        fs.append(x - g); gs.append(g.copy())
        x = g
    F = np.column_stack([fs[i + 1] - fs[i] for i in range(mem)])
    Gm = np.column_stack([gs[i + 1] - gs[i] for i in range(mem)])
    assert np.allclose(aa.Q[:, :mem] @ aa.R[:mem, :mem], F, atol=1e-12)
    eta = np.linalg.lstsq(F, fs[-1], rcond=None)[0]
    g_acc = gs[-1].copy()
    aa.accelerate(g_acc, x, mem + 2)
    assert aa.success and np.allclose(aa.eta[:mem], eta, rtol=1e-8, atol=1e-10)
    assert np.allclose(g_acc, gs[-1] - Gm @ eta, atol=1e-9)


def test_g4_small_sdp_constraint_primals():
    # moi_wrapper.jl:39-106: <A1, X> = 11, <A2, X> = 19 at atol 1e-3 (check_termination = 1 there), X PSD
    res, _ = _solve(G.g4_small_sdp, check_termination=1)
    assert res.status == "Solved"
    assert abs(G.G4_A1 @ res.x - 11.0) < 1e-3 and abs(G.G4_A2 @ res.x - 19.0) < 1e-3
    X = np.zeros((3, 3))
    X[np.triu_indices(3)] = res.x[[0, 1, 3, 2, 4, 5]]     # x = (X11, X12, X22, X13, X23, X33)
    X = X + np.triu(X, 1).T
    assert np.linalg.eigvalsh(X).min() > -1e-4


def test_g5_sigma_max_lmi():
    # nuclear_norm_minimization.jl:31-40 (undecomposed here; the decomposed run is in test_chordal_cpu.py)
    res, _ = _solve(G.g5_sigma_max_lmi)
    assert res.status == "Solved"
    Y = res.x[1:].reshape(3, 3, order="F")
    assert Y[1, 0] <= 4 + 1e-6 and Y[1, 1] >= 3 - 1e-6 and Y.sum() - 12.0 >= -1e-3
    assert abs(np.linalg.svd(Y, compute_uv=False).max() - res.x[0]) <= 1e-3


def test_g11_iteration_limit_keeps_rho():
    res, _ = _solve(G.g11_iteration_limit, max_iter=2)
    assert res.status == "Max_iter_reached" and res.iter == 2
    assert list(res.info.rho_updates) == [0.1]


def test_g17_complex_psd_least_eigenvalue():
    # least_eigenvalue.jl:33-39: obj = 1 - sqrt 2 at atol = rtol = 1e-4 (eps 1e-5)
    res, _ = _solve(G.g17_complex_least_eigenvalue)
    assert res.status == "Solved" and abs(res.obj_val - G.G17_OBJ) < 1e-4 + 1e-4 * abs(G.G17_OBJ)
    res, _ = _solve(G.g17_complex_least_eigenvalue, accelerator="anderson")
    assert res.status == "Solved" and abs(res.obj_val - G.G17_OBJ) < 1e-4 + 1e-4 * abs(G.G17_OBJ)


def test_complex_psd_projection_and_its_real_embedding():
    # the device projects PsdConeTriangle{T, Complex{T}} through the real embedding [[A, -B], [B, A]] of X = A + iB:
    # Pi(embedding) = embedding(Pi), checked here against the Hermitian eigendecomposition
    rng = np.random.default_rng(9)
    for N in (1, 2, 5, 12):
        Z = rng.standard_normal((N, N)) + 1j * rng.standard_normal((N, N))
        X = (Z + Z.conj().T) / 2
        x = O.extract_upper_triangle_complex(X, np.sqrt(2.0))
        assert x.shape == (N * N,)
        assert np.allclose(np.triu(O.populate_upper_triangle_complex(x, N, 1 / np.sqrt(2.0))), np.triu(X))
        p = x.copy()
        O.project_cone(p, O.ComplexPsdConeTriangle(N * N))
        w, V = np.linalg.eigh(X)
        Xp = (V * np.maximum(w, 0)) @ V.conj().T
        assert np.allclose(p, O.extract_upper_triangle_complex(Xp, np.sqrt(2.0)), atol=1e-12)
        A, B = X.real, X.imag
        M = np.block([[A, -B], [B, A]])
        wm, Vm = np.linalg.eigh(M)
        Mp = (Vm * np.maximum(wm, 0)) @ Vm.T
        assert np.allclose(Mp[:N, :N], Xp.real, atol=1e-12) and np.allclose(Mp[N:, :N], Xp.imag, atol=1e-12)
        assert np.allclose(Mp[N:, N:], Xp.real, atol=1e-12) and np.allclose(Mp[:N, N:], -Xp.imag, atol=1e-12)
        # membership predicates (convexset.jl:415-425)
        assert O.in_dual(p, O.ComplexPsdConeTriangle(N * N), 1e-8) and O.in_pol_recc(-p, O.ComplexPsdConeTriangle(N * N), 1e-8)
