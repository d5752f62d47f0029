"""CPU oracle: a NumPy/SciPy restatement of COSMO.jl's ADMM iteration.

THIS IS TEST INFRASTRUCTURE, NOT THE PRODUCT.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import it, and only as the checker / the CPU baseline.
The shipped path is the CUDA library behind ``include/cosmo_b200.h``.

Every function cites the reference file:line it restates (paths relative to
the COSMO.jl tree, v0.8.11).

Parity pinning
--------------
* The recurrence, cones, residuals, infeasibility tests, rho rules and Ruiz
  scaling are pinned against the reference's literal known answers
  (tests/test_oracle_golden.py: examples/qp.jl, test/UnitTests/qp-box.jl,
  examples/lp.jl, examples/lovasz_petersen.jl, test/UnitTests/algebra.jl,
  test/UnitTests/model_modifications.jl ...).
* The CG / MINRES inner solvers live in IterativeSolvers.jl (^0.9), which is
  NOT vendored under the reference tree and whose unit tests are disabled in
  the reference (test/UnitTests/kktsolver.jl:7-8).  ``cg_solve`` and
  ``minres_solve`` restate the published v0.9 algorithm (src/cg.jl,
  src/minres.jl) from the call sites kktsolver_indirect.jl:68-73,149-152:
  **parity unpinned** at iterate level for these two; they are pinned at
  solution level against the direct KKT solve (as the reference's own disabled
  test kktsolver.jl:104-109 demands).
* The exponential / power cones are pinned on the ten known-answer problems of
  test/UnitTests/exp_cone.jl and pow_cone.jl and the property test of sets.jl:84-110.
* The accelerator lives in COSMOAccelerators.jl (^0.1.0), NOT vendored under the
  reference tree.  ``AndersonAccelerator`` restates the reference's default variant
  (type-II, QR-updated least squares, restarted memory, no regulariser) from the
  published method and the call sites accelerator_interface.jl:58-130: **parity
  unpinned** at iterate level; pinned behaviourally as the reference's own
  AccelerationTests do (status, #restarts == #rho adaptions, max adaptions) and by
  the accelerated runs reaching the reference's known answers.  The default stays
  ``accelerator = "empty"`` (EmptyAccelerator), the iterate-level pinned setting.
"""
from __future__ import annotations

import math
import time
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla
from scipy.linalg import lapack as _lapack


# --------------------------------------------------------------------------
# Cones (src/convexset.jl)
# --------------------------------------------------------------------------
@dataclass
class ZeroSet:  # convexset.jl:16-23
    dim: int


@dataclass
class Nonnegatives:  # convexset.jl:52-60
    dim: int
    constr_type: Optional[np.ndarray] = None  # loose rows (bool), set by classify


@dataclass
class Box:  # convexset.jl:803-821
    l: np.ndarray
    u: np.ndarray
    constr_type: Optional[np.ndarray] = None  # -1 loose, 0 ineq, 1 eq

    def __post_init__(self):
        self.l = np.array(self.l, dtype=float).copy()
        self.u = np.array(self.u, dtype=float).copy()
        if np.any(self.l > self.u):  # convexset.jl:826-830
            raise ValueError("Box set: inconsistent lower/upper bounds")

    @property
    def dim(self):
        return self.l.shape[0]


@dataclass
class SecondOrderCone:  # convexset.jl:92-98
    dim: int


@dataclass
class PsdCone:  # convexset.jl:271-284 (column-major square)
    dim: int

    @property
    def sqrt_dim(self):
        r = math.isqrt(self.dim)
        if r * r != self.dim:
            raise ValueError("dimension must be a square")
        return r


@dataclass
class PsdConeTriangle:  # convexset.jl:362-377 (svec upper triangle)
    dim: int

    @property
    def sqrt_dim(self):
        return (math.isqrt(1 + 8 * self.dim) - 1) // 2


@dataclass
class ComplexPsdConeTriangle:  # PsdConeTriangle{T, Complex{T}}, convexset.jl:344-379: dim = N^2
    dim: int

    @property
    def sqrt_dim(self):
        r = math.isqrt(self.dim)
        if r * r != self.dim:
            raise ValueError("dimension must be a square")
        return r


@dataclass
class ExponentialCone:  # convexset.jl:497-507  K_exp = cl{(x,y,z) | y > 0, y e^(x/y) <= z}
    dim: int = 3
    MAX_ITER: int = 100
    EXP_TOL: float = 1e-8


@dataclass
class DualExponentialCone:  # convexset.jl:749-758
    dim: int = 3
    MAX_ITER: int = 100
    EXP_TOL: float = 1e-8


@dataclass
class PowerCone:  # convexset.jl:625-636  K_pow = {(x,y,z) | x^a y^(1-a) >= |z|, x,y >= 0}
    alpha: float
    MAX_ITER: int = 20
    POW_TOL: float = 1e-8
    dim: int = 3

    def __post_init__(self):
        if self.alpha <= 0 or self.alpha >= 1:
            raise ValueError("The exponent alpha of the power cone has to be in (0, 1).")


@dataclass
class DualPowerCone:  # convexset.jl:765-775
    alpha: float
    MAX_ITER: int = 20
    POW_TOL: float = 1e-8
    dim: int = 3

    def __post_init__(self):
        if self.alpha <= 0 or self.alpha >= 1:
            raise ValueError("The exponent alpha of the dual power cone has to be in (0, 1).")


SCALAR_SCALED_CONES = (SecondOrderCone, PsdCone, PsdConeTriangle, ComplexPsdConeTriangle, ExponentialCone, DualExponentialCone, PowerCone,
                       DualPowerCone)  # rectify_scaling!, convexset.jl:955-957


def row_ranges(cones) -> List[slice]:
    """get_set_indices, convexset.jl:985-993."""
    out, s = [], 0
    for c in cones:
        out.append(slice(s, s + c.dim))
        s += c.dim
    return out


# ---- svec helpers (convexset.jl:432-472) ---------------------------------
def populate_upper_triangle(x: np.ndarray, N: int, scaling: float) -> np.ndarray:
    """convexset.jl:432-442 — column-major upper triangle, off-diagonals scaled."""
    X = np.zeros((N, N), dtype=x.dtype)
    iu = np.triu_indices(N)
    # column-major order of the upper triangle: sort by (col, row)
    order = np.lexsort((iu[0], iu[1]))
    r, c = iu[0][order], iu[1][order]
    X[r, c] = np.where(r == c, x, scaling * x)
    return X


def extract_upper_triangle(X: np.ndarray, scaling: float) -> np.ndarray:
    """convexset.jl:462-472."""
    N = X.shape[0]
    iu = np.triu_indices(N)
    order = np.lexsort((iu[0], iu[1]))
    r, c = iu[0][order], iu[1][order]
    v = X[r, c]
    return np.where(r == c, v, scaling * v)


def populate_upper_triangle_complex(x: np.ndarray, N: int, scaling: float) -> np.ndarray:
    """populate_upper_triangle!(A::Matrix{Complex}, x, scaling), convexset.jl:456-472 (upper triangle only)."""
    A = np.zeros((N, N), dtype=complex)
    k = 0
    for j in range(N):
        for i in range(j):
            A[i, j] = scaling * x[k]
            k += 1
        A[j, j] = x[k]
        k += 1
    for j in range(N):
        for i in range(j):
            A[i, j] += 1j * scaling * x[k]
            k += 1
    return A


def extract_upper_triangle_complex(A: np.ndarray, scaling: float) -> np.ndarray:
    """extract_upper_triangle!(A::Matrix{Complex}, x, scaling), convexset.jl:474-490."""
    N = A.shape[0]
    re, im = [], []
    for j in range(N):
        for i in range(j):
            re.append(scaling * A[i, j].real)
            im.append(scaling * A[i, j].imag)
        re.append(A[j, j].real)
    return np.array(re + im)


def _psd_project_hermitian(Xu: np.ndarray) -> np.ndarray:
    """_project! for a Hermitian matrix given by its upper triangle (zheevr + rank-k update, convexset.jl:219-263);
    numpy's eigh (zheevd) stands in for zheevr."""
    X = np.triu(Xu) + np.triu(Xu, 1).conj().T
    w, Z = np.linalg.eigh(X)
    pos = w > 0
    V = Z[:, pos] * np.sqrt(w[pos])
    return V @ V.conj().T


def _psd_project_upper(X: np.ndarray) -> np.ndarray:
    """_project! + _syevr! + rank_k_update!, convexset.jl:163-189,219-263.

    Same LAPACK entry point and flags as the reference: ?syevr('V','A','U',
    abstol=-1), then scale the positive-eigenvalue columns by sqrt(lambda) and
    syrk('U','N') them.  Only the upper triangle of the result is defined.
    """
    N = X.shape[0]
    if X.dtype == np.float32:
        syevr, syrk = _lapack.ssyevr, None
    else:
        syevr, syrk = _lapack.dsyevr, None
    Xf = np.asfortranarray(X)
    w, Z, m_found, isuppz, info = syevr(Xf, compute_v=1, range="A", lower=0, abstol=-1.0)
    if info != 0:
        raise RuntimeError("LAPACK syevr info=%d" % info)  # convexset.jl:186
    pos = w > 0  # strict, convexset.jl:250
    nnz = int(pos.sum())
    out = np.zeros_like(Xf)
    if nnz > 0:
        # relies on ascending order: trailing nnz columns (convexset.jl:258-261)
        V = Z[:, N - nnz:] * np.sqrt(w[N - nnz:])
        out = np.triu(V @ V.T)
    return out


# ---- 3-d exponential / power cones (convexset.jl:510-747) -------------------
def _exp_in_cone(v, tol):  # convexset.jl:600-605
    x, y, z = v
    with np.errstate(over="ignore", invalid="ignore", divide="ignore"):
        return bool((y > 0 and y * np.exp(x / y) <= z + tol) or (x <= tol and y == 0.0 and z >= -tol))


def _exp_in_dual(v, tol):  # convexset.jl:607-612
    x, y, z = v
    with np.errstate(over="ignore", invalid="ignore", divide="ignore"):
        return bool((x < 0 and -x * np.exp(y / x) - math.e * z <= tol) or (abs(x) <= tol and y >= -tol and z >= -tol))


def _exp_find_min_t(lam, s0, t0, tol):  # find_min_t, convexset.jl:578-597 (Newton on dt = t - t0)
    dt = max(-t0, tol)
    for _ in range(150):
        f = dt * (dt + t0) / lam ** 2 - s0 / lam + math.log(dt / lam) + 1.0
        grad_f = (2.0 * dt + t0) / lam ** 2 + 1.0 / dt
        dt = dt - f / grad_f
        if dt <= -t0:
            dt = -t0
            break
        elif dt <= 0:
            dt = 0.0
            break
        elif abs(f) < tol:
            break
    return dt + t0


def _exp_grad_dual(lam, v, v0, tol):  # grad_dual! + find_minimizers!, convexset.jl:559-573
    v[2] = _exp_find_min_t(lam, v0[1], v0[2], tol)
    v[1] = (1.0 / lam) * (v[2] - v0[2]) * v[2]
    v[0] = v0[0] - lam
    if v[1] == 0:
        return v[0]
    with np.errstate(divide="ignore", invalid="ignore"):
        return v[0] + v[1] * np.log(v[1] / v[2])


def _project_exp(v, cone):  # project!, convexset.jl:510-532
    if _exp_in_cone(v, 0.0):
        return
    if _exp_in_dual(-v, 0.0):
        v[:] = 0.0
        return
    if v[0] < 0 and v[1] < 0:
        v[1] = 0.0
        v[2] = max(v[2], 0.0)
        return
    # project_exp!, convexset.jl:537-557: bisection on the dual variable lambda
    v0 = v.copy()
    tol = cone.EXP_TOL
    l, lam = 0.0, 0.125
    g = _exp_grad_dual(lam, v, v0, tol)
    while g > 0:
        l = lam
        lam *= 2
        g = _exp_grad_dual(lam, v, v0, tol)
    u = lam
    for _ in range(cone.MAX_ITER):
        lam = (u + l) / 2
        g = _exp_grad_dual(lam, v, v0, tol)
        if g > 0:
            l = lam
        else:
            u = lam
        if u - l < tol:
            break


def _pow_in_cone(v, alpha, tol):  # convexset.jl:719-725
    x, y, z = v
    return bool(x >= 0 and y >= 0 and x ** alpha * y ** (1 - alpha) >= abs(z) - tol)


def _pow_in_dual(v, alpha, tol):  # convexset.jl:728-734  (NaN powers of negative s, t compare false)
    s_, t_, w_ = v
    if not (s_ >= -tol and t_ >= -tol):
        return False
    if s_ < 0 or t_ < 0:
        return False  # Julia raises a DomainError here; unreachable for tol = 0
    return bool(s_ ** alpha * t_ ** (1 - alpha) >= abs(w_) * alpha ** alpha * (1 - alpha) ** (1 - alpha) - tol)


def _project_pow(v, cone):  # project!, convexset.jl:646-712
    a = cone.alpha
    if _pow_in_cone(v, a, 0.0):
        return
    if _pow_in_dual(-v, a, 0.0):
        v[:] = 0.0
        return
    if abs(v[2]) <= cone.POW_TOL:
        v[0] = max(v[0], 0.0)
        v[1] = max(v[1], 0.0)
        return
    x0, y0, z0 = float(v[0]), float(v[1]), float(v[2])
    az = abs(z0)

    def phic(c0, r, al):  # convexset.jl:698-700
        return max(0.5 * (c0 + math.sqrt(c0 * c0 + 4.0 * al * r * (az - r))), 1e-10)

    r = az / 2.0
    px = py = 0.0
    for _ in range(cone.MAX_ITER):
        px = phic(x0, r, a)
        py = phic(y0, r, 1.0 - a)
        pw = px ** a * py ** (1.0 - a)
        phi = pw - r
        if abs(phi) < cone.POW_TOL:
            break
        dpx = a / (2.0 * px - x0) * (az - 2.0 * r)           # convexset.jl:702-704
        dpy = (1.0 - a) / (2.0 * py - y0) * (az - 2.0 * r)
        dphi = pw * (a * dpx / px + (1.0 - a) * dpy / py) - 1.0   # convexset.jl:711-713
        r = r - phi / dphi
        r = min(max(r, 0.0), az)
    v[0] = px
    v[1] = py
    v[2] = z0 * r / az


def in_cone(x, cone, tol) -> bool:
    if isinstance(cone, ExponentialCone):
        return _exp_in_cone(x, tol)
    if isinstance(cone, DualExponentialCone):  # convexset.jl:779
        return _exp_in_dual(x, tol)
    if isinstance(cone, PowerCone):
        return _pow_in_cone(x, cone.alpha, tol)
    if isinstance(cone, DualPowerCone):
        return _pow_in_dual(x, cone.alpha, tol)
    raise TypeError(cone)


def project_cone(x: np.ndarray, cone) -> None:
    """project!(x, cone) in place on a contiguous view."""
    if isinstance(cone, ZeroSet):  # convexset.jl:25-28
        x[:] = 0.0
    elif isinstance(cone, Nonnegatives):  # convexset.jl:71-74
        np.maximum(x, 0.0, out=x)
    elif isinstance(cone, Box):  # convexset.jl:844-847, algebra.jl:5-7
        # clip(s,l,u) = s<l ? l : (s>u ? u : s)
        x[:] = np.where(x < cone.l, cone.l, np.where(x > cone.u, cone.u, x))
    elif isinstance(cone, SecondOrderCone):  # convexset.jl:100-114
        t = x[0]
        nx = np.linalg.norm(x[1:], 2)
        if nx <= t:
            pass
        elif nx <= -t:
            x[:] = 0.0
        else:
            x[0] = (nx + t) / 2.0
            x[1:] = (nx + t) / (2.0 * nx) * x[1:]
    elif isinstance(cone, PsdCone):  # convexset.jl:303-321
        n = cone.sqrt_dim
        if x.shape[0] == 1:
            x[:] = max(x[0], 0.0)
        else:
            X = x.reshape(n, n, order="F").copy()
            # symmetrize_upper!, algebra.jl:201-208
            Xu = np.triu((X + X.T) / 2.0)
            Xp = _psd_project_upper(Xu)
            full = Xp + np.triu(Xp, 1).T  # mirror, convexset.jl:316-318
            x[:] = full.reshape(-1, order="F")
    elif isinstance(cone, PsdConeTriangle):  # convexset.jl:402-412
        if x.shape[0] == 1:
            x[:] = max(x[0], 0.0)
        else:
            N = cone.sqrt_dim
            X = populate_upper_triangle(x, N, 1.0 / math.sqrt(2.0))
            Xp = _psd_project_upper(X)
            x[:] = extract_upper_triangle(Xp, math.sqrt(2.0))
    elif isinstance(cone, ComplexPsdConeTriangle):  # convexset.jl:402-412 with R = Complex{T}
        if x.shape[0] == 1:
            x[:] = max(x[0], 0.0)
        else:
            X = populate_upper_triangle_complex(x, cone.sqrt_dim, 1.0 / math.sqrt(2.0))
            x[:] = extract_upper_triangle_complex(_psd_project_hermitian(X), math.sqrt(2.0))
    elif isinstance(cone, ExponentialCone):
        _project_exp(x, cone)
    elif isinstance(cone, PowerCone):
        _project_pow(x, cone)
    elif isinstance(cone, (DualExponentialCone, DualPowerCone)):
        # Moreau: Proj_K*(v) = v + Proj_K(-v), convexset.jl:784-789
        v0 = x.copy()
        x *= -1.0
        if isinstance(cone, DualExponentialCone):
            _project_exp(x, ExponentialCone(3, cone.MAX_ITER, cone.EXP_TOL))
        else:
            _project_pow(x, PowerCone(cone.alpha, cone.MAX_ITER, cone.POW_TOL))
        x += v0
    else:
        raise TypeError("unsupported cone %r" % (cone,))


def project(s: np.ndarray, cones) -> None:
    """project!(::SplitVector, ::CompositeConvexSet), convexset.jl:885-891."""
    for rng, c in zip(row_ranges(cones), cones):
        project_cone(s[rng], c)


# ---- dual-cone / recession predicates ------------------------------------
def _is_pos_def(X: np.ndarray, tol: float) -> bool:
    """is_pos_def!, algebra.jl:226-233 (Cholesky success on X + tol I, upper)."""
    if np.iscomplexobj(X):   # Hermitian case: zpotrf on the upper triangle
        Xs = np.triu(X) + np.triu(X, 1).conj().T + tol * np.eye(X.shape[0])
        c, info = _lapack.zpotrf(Xs, lower=0)
        return info == 0
    Xs = np.triu(X) + np.triu(X, 1).T + tol * np.eye(X.shape[0])
    c, info = _lapack.dpotrf(Xs, lower=0)
    return info == 0


def _cone_matrix(x, cone):
    if isinstance(cone, ComplexPsdConeTriangle):
        return populate_upper_triangle_complex(x, cone.sqrt_dim, 1.0 / math.sqrt(2.0))
    if isinstance(cone, PsdCone):
        n = cone.sqrt_dim
        return x.reshape(n, n, order="F")
    return populate_upper_triangle(x, cone.sqrt_dim, 1.0 / math.sqrt(2.0))


def in_dual(x, cone, tol) -> bool:
    if isinstance(cone, ZeroSet):  # convexset.jl:30-32
        return True
    if isinstance(cone, Nonnegatives):  # convexset.jl:76-78
        return not np.any(x < -tol)
    if isinstance(cone, SecondOrderCone):  # convexset.jl:116-118
        return np.linalg.norm(x[1:]) <= (tol + x[0])
    if isinstance(cone, (PsdCone, PsdConeTriangle, ComplexPsdConeTriangle)):  # convexset.jl:324-329,415-419
        return _is_pos_def(_cone_matrix(x, cone), tol)
    if isinstance(cone, ExponentialCone):
        return _exp_in_dual(x, tol)
    if isinstance(cone, PowerCone):
        return _pow_in_dual(x, cone.alpha, tol)
    if isinstance(cone, DualExponentialCone):  # convexset.jl:780: dual of the dual = primal
        return _exp_in_cone(x, tol)
    if isinstance(cone, DualPowerCone):
        return _pow_in_cone(x, cone.alpha, tol)
    raise TypeError(cone)


def in_pol_recc(x, cone, tol) -> bool:
    if isinstance(cone, ZeroSet):  # convexset.jl:34-36
        return not np.any(np.abs(x) > tol)
    if isinstance(cone, Nonnegatives):  # convexset.jl:80-82
        return not np.any(x > tol)
    if isinstance(cone, SecondOrderCone):  # convexset.jl:120-122
        return np.linalg.norm(x[1:]) <= (tol - x[0])
    if isinstance(cone, Box):  # convexset.jl:858-860
        return (not np.any((cone.u == np.inf) & (x > tol))) and (not np.any((cone.l == -np.inf) & (x < -tol)))
    if isinstance(cone, (PsdCone, PsdConeTriangle, ComplexPsdConeTriangle)):  # convexset.jl:331-336,421-425 + algebra.jl:235-238
        return _is_pos_def(-_cone_matrix(x, cone), tol)
    if isinstance(cone, (ExponentialCone, PowerCone, DualExponentialCone, DualPowerCone)):
        return in_dual(-x, cone, tol)  # convexset.jl:616-618, 740-742, 781
    raise TypeError(cone)


def support_function(y, cone, tol) -> float:
    """support_function!(y, cone, tol): convexset.jl:850-856 (Box), :919-923 (cones)."""
    if isinstance(cone, Box):
        with np.errstate(invalid="ignore"):
            terms = np.where((np.abs(y) > tol) & (y > 0), y * cone.u, y * cone.l)
        s = 0.0
        for t in terms:  # sequential sum as in the reference loop
            s += t
        return s
    return 0.0 if in_dual(-y, cone, tol) else np.inf


# --------------------------------------------------------------------------
# algebra.jl KATs
# --------------------------------------------------------------------------
def clip(s, lo, hi, lo_new=None, hi_new=None):
    """algebra.jl:5-7."""
    lo_new = lo if lo_new is None else lo_new
    hi_new = hi if hi_new is None else hi_new
    return np.where(s < lo, lo_new, np.where(s > hi, hi_new, s))


def scaled_norm(E, v, p=2):
    """algebra.jl:9-47 (E = diagonal vector)."""
    if p == 2:
        return math.sqrt(float(np.sum((E * v) ** 2)))
    if p == np.inf:
        return float(np.max(np.abs(E * v))) if len(v) else 0.0
    if p == 1:
        return float(np.sum(np.abs(E * v)))
    raise ValueError("bad norm specified")


def col_norms(A: sp.csc_matrix, v=None):
    """col_norms!, algebra.jl:63-77 (inf-norm of columns, running max into v)."""
    A = sp.csc_matrix(A)
    out = np.zeros(A.shape[1]) if v is None else v
    if A.nnz:
        absA = abs(A)
        mx = np.asarray(absA.max(axis=0).todense()).ravel()
        np.maximum(out, mx, out=out)
    return out


def row_norms(A: sp.csc_matrix, v=None):
    """row_norms!, algebra.jl:93-107."""
    A = sp.csc_matrix(A)
    out = np.zeros(A.shape[0]) if v is None else v
    if A.nnz:
        absA = abs(A)
        mx = np.asarray(absA.max(axis=1).todense()).ravel()
        np.maximum(out, mx, out=out)
    return out


# --------------------------------------------------------------------------
# Settings (src/settings.jl:101-139)
# --------------------------------------------------------------------------
@dataclass
class Settings:
    rho: float = 0.1
    sigma: float = 1e-6
    alpha: float = 1.6
    eps_abs: float = 1e-5
    eps_rel: float = 1e-5
    eps_prim_inf: float = 1e-4
    eps_dual_inf: float = 1e-4
    max_iter: int = 5000
    check_termination: int = 25
    check_infeasibility: int = 40
    scaling: int = 10
    MIN_SCALING: float = 1e-4
    MAX_SCALING: float = 1e4
    adaptive_rho: bool = True
    adaptive_rho_interval: int = 40
    adaptive_rho_tolerance: float = 5.0
    adaptive_rho_max_adaptions: int = 2 ** 62
    RHO_MIN: float = 1e-6
    RHO_MAX: float = 1e6
    RHO_TOL: float = 1e-4
    RHO_EQ_OVER_RHO_INEQ: float = 1e3
    COSMO_INFTY: float = 1e20
    time_limit: float = 0.0
    obj_true: float = float("nan")   # settings.jl:128-129, residuals.jl:132-137
    obj_true_tol: float = 1e-3
    nearly_ratio: float = 100.0      # residuals.jl:119-125 (read by the MOI layer only)
    kkt_solver: str = "direct"  # "direct" (QDLDL stand-in) | "cg" | "minres" | "minres_reduced"
    tol_constant: float = 1.0   # kktsolver_indirect.jl:21
    tol_exponent: float = 1.5
    # accelerator (settings.jl:136-138).  The reference default is "anderson" (Type-II, QR, restarted memory,
    # mem = 15, safeguarded); the oracle defaults to "empty" (= EmptyAccelerator, docs/src/acceleration.md:9-12)
    # because only that configuration is pinned iterate by iterate.
    accelerator: str = "empty"
    accelerator_mem: int = 15
    accelerator_min_mem: int = 3
    safeguard: bool = True
    safeguard_tol: float = 2.0


# --------------------------------------------------------------------------
# Accelerator (COSMOAccelerators.jl ^0.1.0 -- NOT under /root/reference).
# PARITY UNPINNED: restated from the published method (Garstka, Cannon, Goulart,
# "Safeguarded Anderson acceleration for parametric nonexpansive operators", 2022,
# Alg. 2: type-II Anderson acceleration, least squares by an updated QR factorisation,
# memory restarted when full) and from the call sites in the reference
# (accelerator_interface.jl:58-130, solver.jl:143,157,274-292, setup.jl:10-14,44-50).
# The constants min_mem = 3 and the |eta|_2 > 1e4 rejection are the package defaults
# as remembered; the reference's own tests pin this path only behaviourally
# (AccelerationTests: status :Solved, #restarts == #rho adaptions).
# --------------------------------------------------------------------------
class AndersonAccelerator:
    """AndersonAccelerator{T, Type2{QRDecomp}, RestartedMemory, NoRegularizer}(dim; mem)."""

    def __init__(self, dim, mem=15, min_mem=3):
        if mem <= 2:
            raise ValueError("Memory has to be bigger than two.")
        self.dim = dim
        self.mem = min(mem, dim)
        self.min_mem = min_mem
        self.G = np.zeros((dim, self.mem))   # columns g_k - g_{k-1}
        self.Q = np.zeros((dim, self.mem))   # F = [f_k - f_{k-1}] = Q R
        self.R = np.zeros((self.mem, self.mem))
        self.g_last = np.zeros(dim)
        self.f = np.zeros(dim)
        self.f_last = np.zeros(dim)
        self.eta = np.zeros(self.mem)
        self.iter = 0
        self.init_phase = True
        self.success = False
        self.num_accelerated_steps = 0
        self.log = []   # (iteration, event) like CA.log!

    def empty_caches(self):
        self.G[:] = 0.0
        self.Q[:] = 0.0
        self.R[:] = 0.0
        self.iter = 0

    def restart(self):   # CA.restart!: forget the history; the next update! only stores (x, g, f)
        self.empty_caches()
        self.init_phase = True

    def update(self, g, x, num_iter):
        """CA.update!(aa, g, x, num_iter) with g = w = T(w_prev), x = w_prev."""
        self.f[:] = x - g
        if self.init_phase:
            self.g_last[:] = g
            self.f_last[:] = self.f
            self.init_phase = False
            return
        j = self.iter % self.mem          # 0-based column that receives the new differences
        if j == 0 and self.iter != 0:     # RestartedMemory: history full -> start again
            self.empty_caches()
            self.log.append((num_iter, "memory_full"))
        self.G[:, j] = g - self.g_last
        df = self.f - self.f_last
        self.g_last[:] = g
        self.f_last[:] = self.f
        # qr!: modified Gram-Schmidt update of F = Q R by the new column
        q = df
        for i in range(j):
            self.R[i, j] = self.Q[:, i] @ q
            q = q - self.R[i, j] * self.Q[:, i]
        self.R[j, j] = np.linalg.norm(q)
        with np.errstate(divide="ignore", invalid="ignore"):
            self.Q[:, j] = q / self.R[j, j]
        self.iter += 1

    def accelerate(self, g, x, num_iter):
        """CA.accelerate!(g, x, aa, num_iter): overwrites g with g - G eta, eta = argmin |f - F eta|_2."""
        l = min(self.iter, self.mem)
        if l < self.min_mem:
            self.success = False
            return
        eta = self.Q[:, :l].T @ self.f
        Rl = self.R[:l, :l]
        ok = bool(np.all(np.diag(Rl) != 0.0) and np.all(np.isfinite(Rl)))
        if ok:
            for i in range(l - 1, -1, -1):       # back substitution (LAPACK trtrs 'U','N','N')
                eta[i] = (eta[i] - Rl[i, i + 1:] @ eta[i + 1:]) / Rl[i, i]
            ok = bool(np.all(np.isfinite(eta)) and np.linalg.norm(eta) <= 1e4)
        if not ok:
            self.success = False
            self.log.append((num_iter, "acc_failed"))
            return
        self.eta[:l] = eta
        g -= self.G[:, :l] @ eta
        self.num_accelerated_steps += 1
        self.success = True


# --------------------------------------------------------------------------
# KKT solvers (src/linear_solver/*)
# --------------------------------------------------------------------------
class DirectKKT:
    """Stand-in for QdldlKKTSolver (kktsolver.jl:285-320): sparse LU of
    [P+sigma I, A'; A, -diag(1/rho)].  QDLDL itself is a third-party package."""

    def __init__(self, P, A, sigma, rho):
        self.P, self.A, self.sigma = sp.csc_matrix(P), sp.csc_matrix(A), sigma
        self.m, self.n = A.shape
        self.update_rho(rho)
        self.multiplications: List[int] = []

    def update_rho(self, rho):
        rho = np.broadcast_to(np.asarray(rho, dtype=float), (self.m,))
        K = sp.bmat([[self.P + self.sigma * sp.identity(self.n), self.A.T],
                     [self.A, -sp.diags(1.0 / rho)]], format="csc")
        self.lu = spla.splu(K)

    def solve(self, rhs):
        return self.lu.solve(rhs)


def cg_solve(x, L, b, abstol, reltol=0.0, maxiter=None):
    """IterativeSolvers.jl v0.9 ``cg!`` (src/cg.jl, not vendored) as called at
    kktsolver_indirect.jl:70: non-zero initial guess => r = b - A x costs one
    product; tol = max(reltol*|r0|, abstol); stop when |r| <= tol or
    iteration >= maxiter (default size(A,2)).  Returns (x, iterations, products).
    """
    n = b.shape[0]
    maxiter = n if maxiter is None else maxiter
    prods = 1
    r = b - L(x)
    u = np.zeros(n)
    residual = float(np.linalg.norm(r))
    prev_residual = 1.0
    tol = max(reltol * residual, abstol)
    it = 0
    while it < maxiter and not (residual <= tol):
        beta = residual ** 2 / prev_residual ** 2
        u = r + beta * u
        c = L(u)
        prods += 1
        alpha = residual ** 2 / float(u @ c)
        x += alpha * u
        r -= alpha * c
        prev_residual = residual
        residual = float(np.linalg.norm(r))
        it += 1
    return x, it, prods


def _givens(f, g):
    """LinearAlgebra.givensAlgorithm (real case): returns c, s, r with
    [c s; -s c] [f; g] = [r; 0]."""
    if g == 0.0:
        return 1.0, 0.0, f
    if f == 0.0:
        return 0.0, 1.0, g
    r = math.hypot(f, g)
    c, s = f / r, g / r
    if abs(f) > abs(g) and c < 0:
        c, s, r = -c, -s, -r
    return c, s, r


def minres_solve(x, L, b, abstol, reltol=0.0, maxiter=None):
    """IterativeSolvers.jl v0.9 ``minres!`` (src/minres.jl, not vendored) as
    called at kktsolver_indirect.jl:73,152.  Returns (x, iterations, products)."""
    n = b.shape[0]
    maxiter = n if maxiter is None else maxiter
    v_prev = np.zeros(n)
    v_curr = b - L(x)
    prods = 1
    resnorm = float(np.linalg.norm(v_curr))
    tol = max(reltol * resnorm, abstol)
    if resnorm > 0:
        v_curr = v_curr / resnorm
    w_prev, w_curr = np.zeros(n), np.zeros(n)
    H = [0.0, 0.0, 0.0, 0.0]
    rhs = [resnorm, 0.0]
    c_prev, s_prev, c_curr, s_curr = 1.0, 0.0, 1.0, 0.0
    it = 1
    while it <= maxiter and not (resnorm <= tol):
        v_next = L(v_curr)
        prods += 1
        if it > 1:
            v_next -= H[1] * v_prev
        proj = float(v_curr @ v_next)
        H[2] = proj
        v_next -= proj * v_curr
        H[3] = float(np.linalg.norm(v_next))
        v_next = v_next / H[3]
        if it > 2:
            H[0] = s_prev * H[1]
            H[1] = c_prev * H[1]
        if it > 1:
            tmp = -s_curr * H[1] + c_curr * H[2]
            H[1] = c_curr * H[1] + s_curr * H[2]
            H[2] = tmp
        c, s, H[2] = _givens(H[2], H[3])
        rhs[1] = -s * rhs[0]
        rhs[0] = c * rhs[0]
        w_next = v_curr.copy()
        if it > 1:
            w_next -= H[1] * w_curr
        if it > 2:
            w_next -= H[0] * w_prev
        w_next /= H[2]
        x += rhs[0] * w_next
        v_prev, v_curr = v_curr, v_next
        w_prev, w_curr = w_curr, w_next
        c_prev, s_prev, c_curr, s_curr = c_curr, s_curr, c, s
        rhs[0] = rhs[1]
        H[1] = H[3]
        resnorm = abs(rhs[1])
        it += 1
    return x, it - 1, prods


class IndirectReducedKKT:
    """IndirectReducedKKTSolver, kktsolver_indirect.jl:3-88 (CG or MINRES on
    (P + sigma I + A' rho A) y1 = x1 + A' rho x2 ;  y2 = rho (A y1 - x2))."""

    def __init__(self, P, A, sigma, rho, solver_type="CG", tol_constant=1.0, tol_exponent=1.5):
        self.P, self.A = sp.csc_matrix(P), sp.csc_matrix(A)
        self.At = self.A.T  # lazy adjoint (csr view of the csc arrays)
        self.m, self.n = A.shape
        self.sigma = sigma
        self.rho = np.array(np.broadcast_to(np.asarray(rho, dtype=float), (self.m,)))
        self.tol_constant, self.tol_exponent = tol_constant, tol_exponent
        self.solver_type = solver_type
        self.previous_solution = np.zeros(self.n)
        self.iteration_counter = 1
        self.multiplications: List[int] = []
        self.inner_iterations: List[int] = []

    def update_rho(self, rho):  # :164-166
        self.rho[:] = rho

    def get_tolerance(self):  # :168-170
        return self.tol_constant / self.iteration_counter ** self.tol_exponent

    def reduced_mul(self, x):  # :57-67
        tmp_m = self.A @ x
        tmp_m *= self.rho
        tmp_n = self.At @ tmp_m
        tmp_n += self.sigma * x
        y = self.P @ x
        y += tmp_n
        return y

    def solve(self, rhs):  # :36-88
        n, m = self.n, self.m
        x1, x2 = rhs[:n], rhs[n:]
        y2 = self.rho * x2
        y1 = self.At @ y2
        y1 += x1
        if self.solver_type == "CG":
            nrm = float(np.linalg.norm(y1))
            with np.errstate(divide="ignore"):
                abstol = self.get_tolerance() / nrm if nrm != 0 else np.inf
            _, it, prods = cg_solve(self.previous_solution, self.reduced_mul, y1, abstol=abstol, reltol=0.0)
        else:
            init_res = float(np.linalg.norm(self.reduced_mul(self.previous_solution) - y1))
            abstol = self.get_tolerance() / init_res if init_res != 0 else np.inf
            _, it, prods = minres_solve(self.previous_solution, self.reduced_mul, y1, abstol=abstol, reltol=0.0)
            prods += 1
        self.multiplications.append(prods)
        self.inner_iterations.append(it)
        y1 = self.previous_solution.copy()
        y2 = self.A @ y1
        y2 -= x2
        y2 *= self.rho
        self.iteration_counter += 1
        return np.concatenate([y1, y2])


class IndirectKKT:
    """IndirectKKTSolver, kktsolver_indirect.jl:90-162 (MINRES on the full KKT)."""

    def __init__(self, P, A, sigma, rho, tol_constant=1.0, tol_exponent=1.5):
        self.P, self.A = sp.csc_matrix(P), sp.csc_matrix(A)
        self.At = self.A.T
        self.m, self.n = A.shape
        self.sigma = sigma
        self.rho = np.array(np.broadcast_to(np.asarray(rho, dtype=float), (self.m,)))
        self.tol_constant, self.tol_exponent = tol_constant, tol_exponent
        self.previous_solution = np.zeros(self.n + self.m)
        self.iteration_counter = 1
        self.multiplications: List[int] = []
        self.inner_iterations: List[int] = []

    def update_rho(self, rho):
        self.rho[:] = rho

    def get_tolerance(self):
        return self.tol_constant / self.iteration_counter ** self.tol_exponent

    def kkt_mul(self, x):  # :130-148
        n = self.n
        x1, x2 = x[:n], x[n:]
        tmp_n = self.At @ x2
        tmp_n += self.sigma * x1
        y1 = self.P @ x1
        y1 += tmp_n
        y2 = -x2 / self.rho
        y2 += self.A @ x1
        return np.concatenate([y1, y2])

    def solve(self, rhs):  # :123-162
        init_res = float(np.linalg.norm(self.kkt_mul(self.previous_solution) - rhs))
        abstol = self.get_tolerance() / init_res if init_res != 0 else np.inf
        _, it, prods = minres_solve(self.previous_solution, self.kkt_mul, rhs, abstol=abstol, reltol=0.0)
        self.multiplications.append(prods + 1)
        self.inner_iterations.append(it)
        self.iteration_counter += 1
        return self.previous_solution.copy()


def make_kkt_solver(name, P, A, sigma, rho, settings: Settings):
    """_make_kkt_solver!, setup.jl:1-7."""
    if name == "direct":
        return DirectKKT(P, A, sigma, rho)
    if name == "cg":
        return IndirectReducedKKT(P, A, sigma, rho, "CG", settings.tol_constant, settings.tol_exponent)
    if name == "minres_reduced":
        return IndirectReducedKKT(P, A, sigma, rho, "MINRES", settings.tol_constant, settings.tol_exponent)
    if name == "minres":
        return IndirectKKT(P, A, sigma, rho, settings.tol_constant, settings.tol_exponent)
    raise ValueError(name)


# --------------------------------------------------------------------------
# Scaling (src/scaling.jl)
# --------------------------------------------------------------------------
@dataclass
class ScaleMatrices:  # types.jl:130-151
    D: np.ndarray
    E: np.ndarray
    c: float = 1.0

    @property
    def Dinv(self):
        return 1.0 / self.D

    @property
    def Einv(self):
        return 1.0 / self.E

    @property
    def cinv(self):
        return 1.0 / self.c


def _limit_scaling(s, st: Settings):  # scaling.jl:10-18
    return clip(s, st.MIN_SCALING, st.MAX_SCALING, 1.0, st.MAX_SCALING)


def scale_ruiz(P, q, A, b, cones, st: Settings):
    """scale_ruiz!, scaling.jl:21-116.  Returns scaled copies (P,q,A,b), scaled
    cones (Box bounds) and the ScaleMatrices."""
    P = sp.csc_matrix(P, dtype=float).copy()
    A = sp.csc_matrix(A, dtype=float).copy()
    q = np.array(q, dtype=float)
    b = np.array(b, dtype=float)
    m, n = A.shape
    D, E, c = np.ones(n), np.ones(m), 1.0

    def scale_data(Ds, Es):  # scaling.jl:157-168
        nonlocal P, A, q, b
        P = sp.csc_matrix(sp.diags(Ds) @ P @ sp.diags(Ds))
        A = sp.csc_matrix(sp.diags(Es) @ A @ sp.diags(Ds))
        q = Ds * q
        b = Es * b

    for _ in range(st.scaling):
        Dw = col_norms(P)                 # kkt_col_norms!, scaling.jl:3-8
        col_norms(A, Dw)
        Ew = row_norms(A)
        Dw = _limit_scaling(Dw, st)
        Ew = _limit_scaling(Ew, st)
        Dw = 1.0 / np.sqrt(Dw)            # inv_sqrt!, :125-127
        Ew = 1.0 / np.sqrt(Ew)
        scale_data(Dw, Ew)
        D *= Dw
        E *= Ew
        mean_col_norm_P = float(np.mean(col_norms(P))) if n else 0.0
        inf_norm_q = float(np.max(np.abs(q))) if n else 0.0
        if mean_col_norm_P != 0.0 and inf_norm_q != 0.0:
            inf_norm_q = float(_limit_scaling(inf_norm_q, st))
            scale_cost = max(inf_norm_q, mean_col_norm_P)
            scale_cost = float(_limit_scaling(scale_cost, st))
            ctmp = 1.0 / scale_cost
            P = P * ctmp
            q = q * ctmp
            c *= ctmp

    # rectify_set_scalings!, scaling.jl:129-142 + convexset.jl:905-958,978-982
    Ew = np.ones(m)
    changed = False
    for rng, cone in zip(row_ranges(cones), cones):
        if isinstance(cone, SCALAR_SCALED_CONES):
            tmp = np.mean(E[rng])
            Ew[rng] = tmp / E[rng]
            changed = True
    if changed:
        scale_data(np.ones(n), Ew)
        E *= Ew
    # issymmetric(P) || symmetrize_full!(P)  (scaling.jl:99) -- P stays symmetric here.
    # scale_sets!, scaling.jl:145-154 -> Box bounds, convexset.jl:863-867
    new_cones = []
    for rng, cone in zip(row_ranges(cones), cones):
        if isinstance(cone, Box):
            new_cones.append(Box(cone.l * E[rng], cone.u * E[rng]))
        else:
            new_cones.append(cone)
    return P, q, A, b, new_cones, ScaleMatrices(D, E, c)


# --------------------------------------------------------------------------
# rho vector (src/parameters.jl, src/setup.jl:75-85)
# --------------------------------------------------------------------------
def classify_constraints(cones, b, st: Settings):
    """classify_constraints!, setup.jl:75-85; convexset.jl:62-69, 831-842."""
    for rng, cone in zip(row_ranges(cones), cones):
        if isinstance(cone, Nonnegatives):
            cone.constr_type = b[rng] > st.COSMO_INFTY * st.MIN_SCALING
        elif isinstance(cone, Box):
            ct = np.zeros(cone.dim, dtype=np.int64)
            loose = (cone.l < -st.COSMO_INFTY * st.MIN_SCALING) & (cone.u > st.COSMO_INFTY * st.MIN_SCALING)
            with np.errstate(invalid="ignore"):
                eq = (~loose) & ((cone.u - cone.l) < st.RHO_TOL)
            ct[loose] = -1
            ct[eq] = 1
            cone.constr_type = ct


def apply_constraint_rho_scaling(rho_vec, cones, st: Settings):
    """parameters.jl:17-49."""
    for rng, cone in zip(row_ranges(cones), cones):
        if isinstance(cone, ZeroSet):
            rho_vec[rng] *= st.RHO_EQ_OVER_RHO_INEQ
        elif isinstance(cone, Nonnegatives):
            v = rho_vec[rng]
            v[cone.constr_type] = st.RHO_MIN
        elif isinstance(cone, Box):
            v = rho_vec[rng]
            v[cone.constr_type == -1] = st.RHO_MIN
            v[cone.constr_type == 1] *= st.RHO_EQ_OVER_RHO_INEQ


# --------------------------------------------------------------------------
# Result types (src/types.jl:65-112)
# --------------------------------------------------------------------------
@dataclass
class ResultInfo:
    r_prim: float = np.inf
    r_dual: float = np.inf
    max_norm_prim: float = 0.0
    max_norm_dual: float = 0.0
    rho_updates: List[float] = field(default_factory=list)


@dataclass
class Result:
    x: np.ndarray
    y: np.ndarray
    s: np.ndarray
    obj_val: float
    iter: int
    status: str
    info: ResultInfo
    times: dict
    # extras for parity checks (scaled internal state at exit)
    w: Optional[np.ndarray] = None
    rho_vec: Optional[np.ndarray] = None
    kkt: object = None
    history: Optional[list] = None
    safeguarding_iter: int = 0   # iter = ADMM iterations + safeguarding_iter (solver.jl:195-199)


class Workspace:
    """The slice of COSMO.Workspace (types.jl:348-391) the loop touches."""

    def __init__(self, P, q, A, b, cones, settings: Settings):
        self.st = settings
        self.P0 = sp.csc_matrix(P, dtype=float)
        self.A0 = sp.csc_matrix(A, dtype=float)
        self.q0 = np.array(q, dtype=float)
        self.b0 = np.array(b, dtype=float)
        self.cones0 = list(cones)
        self.m, self.n = self.A0.shape
        assert sum(c.dim for c in cones) == self.m
        self.x = np.zeros(self.n)
        self.s = np.zeros(self.m)
        self.mu = np.zeros(self.m)
        self.is_scaled = False
        self.is_optimized = False
        self.kkt = None
        self.accelerator = None
        self.accelerator_active = False
        self.rho_updates: List[float] = []

    # warm starts in *unscaled* coordinates, interface.jl:117-179
    def warm_start(self, x=None, s=None, y=None):
        if x is not None:
            self.x[:] = x
        if s is not None:
            self.s[:] = s
        if y is not None:
            self.mu[:] = -np.asarray(y)

    # ---- setup!, setup.jl:18-64 ------------------------------------------
    def setup(self):
        st = self.st
        if st.scaling != 0 and not self.is_scaled:
            self.P, self.q, self.A, self.b, self.cones, self.sm = scale_ruiz(
                self.P0, self.q0, self.A0, self.b0, self.cones0, st)
            self.is_scaled = True
            self._scale_variables()
        elif not self.is_scaled:
            self.P, self.q, self.A, self.b = self.P0, self.q0.copy(), self.A0, self.b0.copy()
            self.cones = self.cones0
            self.sm = ScaleMatrices(np.ones(self.n), np.ones(self.m), 1.0)
            self.is_scaled = True
        else:
            self._scale_variables()
        self.At = self.A.T
        classify_constraints(self.cones, self.b, st)
        if not self.is_optimized:  # set_rho_vec!, parameters.jl:3-13
            self.rho = st.rho
            self.rho_vec = self.rho * np.ones(self.m)
            apply_constraint_rho_scaling(self.rho_vec, self.cones, st)
            self.rho_updates.append(self.rho)
        # setup.jl:44-50: build the accelerator with the KKT solver, restart it on a re-solve
        if self.kkt is None:
            self.accelerator = (AndersonAccelerator(self.n + self.m, st.accelerator_mem, st.accelerator_min_mem)
                                if st.accelerator == "anderson" else None)
        elif self.accelerator is not None:
            self.accelerator.restart()
        self.accelerator_active = False
        if self.kkt is None:
            self.kkt = make_kkt_solver(st.kkt_solver, self.P, self.A, st.sigma, self.rho_vec, st)

    def _scale_variables(self):  # scale_variables!, scaling.jl:118-123
        self.x[:] = self.sm.Dinv * self.x
        self.mu[:] = self.sm.Einv * self.mu
        self.s[:] = self.sm.E * self.s
        self.mu *= self.sm.c

    # ---- residuals.jl ------------------------------------------------------
    def calculate_residuals(self, ignore_scaling=False):
        """residuals.jl:30-53."""
        r_prim = self.A @ self.xv + self.s - self.b
        r_dual = self.P @ self.xv + self.q - self.At @ self.mu
        if self.st.scaling != 0 and not ignore_scaling:
            r_prim = self.sm.Einv * r_prim
            r_dual = self.sm.cinv * (self.sm.Dinv * r_dual)
        return _ninf(r_prim), _ninf(r_dual)

    def max_res_component_norm(self, ignore_scaling=False):
        """residuals.jl:56-96."""
        unscale = self.st.scaling != 0 and not ignore_scaling
        Einv = self.sm.Einv if unscale else 1.0
        Dc = self.sm.Dinv * self.sm.cinv if unscale else 1.0
        mp = max(_ninf(Einv * (self.A @ self.xv)), _ninf(Einv * self.s), _ninf(Einv * self.b))
        md = max(_ninf(Dc * (self.P @ self.xv)), _ninf(Dc * self.q), _ninf(Dc * (self.At @ self.mu)))
        return mp, md

    def calculate_result_info(self):  # residuals.jl:149-153
        rp, rd = self.calculate_residuals()
        mp, md = self.max_res_component_norm()
        return ResultInfo(rp, rd, mp, md, self.rho_updates)

    def calculate_cost(self):  # residuals.jl:143-147
        return self.sm.cinv * (0.5 * float((self.P @ self.xv) @ self.xv) + float(self.q @ self.xv))

    def has_converged(self, r: ResultInfo):  # residuals.jl:98-140
        st = self.st
        obj_true_flag = True
        if not np.isnan(st.obj_true):   # a known optimal value must be met as well (residuals.jl:132-137)
            obj_true_flag = abs(st.obj_true - self.calculate_cost()) <= st.obj_true_tol
        return (r.r_prim < st.eps_abs + st.eps_rel * r.max_norm_prim) and \
               (r.r_dual < st.eps_abs + st.eps_rel * r.max_norm_dual) and obj_true_flag

    # ---- infeasibility.jl --------------------------------------------------
    def is_primal_infeasible(self, dy):  # infeasibility.jl:1-29
        st = self.st
        norm_dy = scaled_norm(self.sm.E, dy, np.inf)
        if norm_dy > st.eps_prim_inf:
            A_dy = self.sm.Dinv * (self.At @ dy)
            if _ninf(A_dy) <= st.eps_prim_inf * norm_dy:
                dy = dy * (-1.0 / norm_dy)
                dyt_b = float(dy @ self.b)
                sF = 0.0
                for rng, cone in zip(row_ranges(self.cones), self.cones):
                    sF += support_function(dy[rng], cone, st.eps_prim_inf)
                sF -= dyt_b
                if sF <= st.eps_prim_inf:
                    return True
        return False

    def is_dual_infeasible(self, dx):  # infeasibility.jl:32-68
        st = self.st
        norm_dx = scaled_norm(self.sm.D, dx, np.inf)
        if norm_dx > st.eps_dual_inf:
            if float(self.q @ dx) / (norm_dx * self.sm.c) < -st.eps_dual_inf:
                P_dx = self.sm.Dinv * (self.P @ dx)
                if _ninf(P_dx) / (norm_dx * self.sm.c) <= st.eps_dual_inf:
                    A_dx = self.sm.Einv * (self.A @ dx)
                    A_dx *= 1.0 / norm_dx
                    ok = all(in_pol_recc(A_dx[rng], cone, st.eps_dual_inf)
                             for rng, cone in zip(row_ranges(self.cones), self.cones))
                    if ok:
                        return True
        return False

    # ---- parameters.jl:53-92 ----------------------------------------------
    def adapt_rho_vec(self):
        st = self.st
        rp, rd = self.calculate_residuals(True)
        mp, md = self.max_res_component_norm(True)
        rp = rp / (mp + 1e-10)
        rd = rd / (md + 1e-10)
        new_rho = self.rho * math.sqrt(rp / (rd + 1e-10))
        new_rho = min(max(new_rho, st.RHO_MIN), st.RHO_MAX)
        if new_rho > st.adaptive_rho_tolerance * self.rho or new_rho < (1.0 / st.adaptive_rho_tolerance) * self.rho:
            self.rho = new_rho
            self.rho_vec[:] = new_rho
            apply_constraint_rho_scaling(self.rho_vec, self.cones, st)
            self.rho_updates.append(new_rho)
            self.kkt.update_rho(self.rho_vec)
            return True
        return False

    def recover_mu(self):  # solver.jl:24-26
        self.mu[:] = self.rho_vec * (self.w_prev[self.n:] - self.s)

    # ---- optimize!, solver.jl:78-203 ---------------------------------------
    def optimize(self, record_history=False, iter_callback=None) -> Result:
        st = self.st
        n, m = self.n, self.m
        t0 = time.perf_counter()
        self.setup()
        setup_time = time.perf_counter() - t0
        times = {"setup_time": setup_time, "proj_time": 0.0, "kkt_time": 0.0, "res_time": 0.0}
        status = "Undetermined"
        cost = np.inf
        res_info = ResultInfo(np.inf, np.inf, 0.0, 0.0, self.rho_updates)
        it = 0
        sigma, alpha = st.sigma, st.alpha
        rho_update_due = False
        infeasibility_check_due = False
        dy = np.zeros(m)
        history = [] if record_history else None

        # warm starting the operator variable, solver.jl:128-129
        self.w = np.concatenate([self.x, self.mu / self.rho_vec + self.s])
        self.w_prev = np.concatenate([self.x, np.zeros(m)])
        self.xv = self.w_prev[:n]  # x = view(w_prev, 1:n), types.jl:274
        self.is_optimized = True
        iter_start = time.perf_counter()

        def admm_x():  # solver.jl:32-56
            ls = np.concatenate([sigma * self.w[:n] - self.q, self.b - 2.0 * self.s + self.w[n:]])
            tk = time.perf_counter()
            sol = self.kkt.solve(ls)
            times["kkt_time"] += time.perf_counter() - tk
            s_tl = 2.0 * self.s - self.w[n:] - sol[n:] / self.rho_vec
            return sol[:n], s_tl

        def admm_w(x_tl, s_tl):  # solver.jl:62-65
            self.w[:n] = self.w[:n] + alpha * (x_tl - self.w[:n])
            self.w[n:] = self.w[n:] + alpha * (s_tl - self.s)

        x_tl, s_tl = admm_x()
        admm_w(x_tl, s_tl)

        aa = self.accelerator
        safeguarding_iter = 0

        def update_suggested(due):  # solver.jl:284-292: postponed to the next non-accelerated iteration
            return due and not (aa is not None and aa.success)

        def admm_z():  # admm_z!, :7-21
            tp = time.perf_counter()
            self.s[:] = self.w[n:]
            project(self.s, self.cones)
            times["proj_time"] += time.perf_counter() - tp

        while it + safeguarding_iter < st.max_iter:
            it += 1
            # acceleration_pre!, accelerator_interface.jl:58-75 (ImmediateActivation, :24-28)
            if aa is not None:
                if not self.accelerator_active and it >= 2:
                    self.accelerator_active = True
                if self.accelerator_active:
                    aa.update(self.w, self.w_prev, it)
                    aa.accelerate(self.w, self.w_prev, it)   # overwrites w
            if update_suggested(infeasibility_check_due):  # solver.jl:145-148
                self.recover_mu()
                dy[:] = self.mu
            self.w_prev[:] = self.w  # :151
            admm_z()
            # apply_rho_adaptation_rules!, :242-282 (interval > 0 only: deterministic)
            if st.adaptive_rho and st.adaptive_rho_interval > 0 and it % st.adaptive_rho_interval == 0 \
                    and (len(self.rho_updates) - 1) < st.adaptive_rho_max_adaptions:
                rho_update_due = True
            if update_suggested(rho_update_due):
                rho_update_due = False
                self.recover_mu()
                if self.adapt_rho_vec():
                    if aa is not None:   # the operator changed: restart the accelerator, :272-275
                        aa.restart()
                        aa.log.append((it, "rho_adapted"))
                    self.w[n:] = self.mu / self.rho_vec + self.s  # :278
            x_tl, s_tl = admm_x()
            admm_w(x_tl, s_tl)
            # acceleration_post!, accelerator_interface.jl:85-114: safeguarding
            if aa is not None and self.accelerator_active and aa.success and st.safeguard:
                nrm_tol = np.linalg.norm(aa.f) * st.safeguard_tol
                aa.f[:] = self.w_prev - self.w        # compute_accelerated_res_norm!, :120-123
                if np.linalg.norm(aa.f) > nrm_tol:
                    aa.log.append((it, "acc_guarded_declined"))
                    self.w_prev[:] = aa.g_last          # reset_accelerated_vector!, :126-130
                    self.w[:] = aa.g_last
                    admm_z()
                    x_tl, s_tl = admm_x()
                    admm_w(x_tl, s_tl)
                    safeguarding_iter += 1
                else:
                    aa.log.append((it, "acc_guarded_accepted"))
            if record_history:
                history.append(self.w.copy())
            if iter_callback is not None:
                iter_callback(it, self)

            # check_termination!, :303-356
            if it % st.check_termination == 0 or it == 1:
                tr = time.perf_counter()
                self.recover_mu()
                res_info = self.calculate_result_info()
                cost = self.calculate_cost()
                times["res_time"] += time.perf_counter() - tr
                if abs(cost) > 1e20:
                    status = "Unsolved"
                    break
                if self.has_converged(res_info):
                    status = "Solved"
                    break
            if it % st.check_infeasibility == 0:
                infeasibility_check_due = True
            elif update_suggested(infeasibility_check_due):
                infeasibility_check_due = False
                self.recover_mu()
                dy -= self.mu
                dx = self.w[:n] - self.w_prev[:n]
                if self.is_primal_infeasible(dy.copy()):
                    status, cost = "Primal_infeasible", np.inf
                    break
                if self.is_dual_infeasible(dx):
                    status, cost = "Dual_infeasible", -np.inf
                    break
            if st.time_limit != 0 and (time.perf_counter() - iter_start) > st.time_limit:
                res_info = self.calculate_result_info()
                status = "Time_limit_reached"
                break

        self.recover_mu()  # :167
        times["iter_time"] = time.perf_counter() - iter_start
        if it + safeguarding_iter == st.max_iter and status == "Undetermined":
            res_info = self.calculate_result_info()
            status = "Max_iter_reached"
        w_exit = self.w.copy()
        # reverse_scaling!, scaling.jl:170-179
        x = self.w_prev[:n].copy()
        s = self.s.copy()
        mu = self.mu.copy()
        if st.scaling != 0:
            x = self.sm.D * x
            s = self.sm.Einv * s
            mu = self.sm.E * mu * self.sm.cinv
            # keep the workspace variables unscaled too (they are re-scaled in setup! on re-solve)
        self.x[:] = x
        self.s[:] = s
        self.mu[:] = mu
        times["solver_time"] = time.perf_counter() - t0
        return Result(x, -mu, s, cost, it + safeguarding_iter, status, res_info, times, w=w_exit,
                      rho_vec=self.rho_vec.copy(), kkt=self.kkt, history=history, safeguarding_iter=safeguarding_iter)

    # update!(q, b), interface.jl:187-211
    def update(self, q=None, b=None):
        if q is not None:
            self.q0 = np.array(q, dtype=float)
            if self.is_scaled:
                self.q = (self.sm.D * self.q0) * self.sm.c
        if b is not None:
            self.b0 = np.array(b, dtype=float)
            if self.is_scaled:
                self.b = self.sm.E * self.b0


def _ninf(v):
    return float(np.max(np.abs(v))) if len(v) else 0.0


# --------------------------------------------------------------------------
# Model building (src/constraint.jl, src/interface.jl) -- input formatting only
# --------------------------------------------------------------------------
@dataclass
class Constraint:
    """COSMO.Constraint: A x + b in convex_set (constraint.jl:47-76)."""
    A: object
    b: object
    convex_set: object

    def __post_init__(self):
        A = self.A
        if not sp.issparse(A):
            A = np.atleast_2d(np.asarray(A, dtype=float))
        self.A = sp.csr_matrix(A, dtype=float)
        self.b = np.atleast_1d(np.asarray(self.b, dtype=float)).ravel()
        if self.A.shape[0] != self.b.shape[0]:
            raise ValueError("The dimensions of matrix A and vector b don't match.")
        if self.A.shape[0] != self.convex_set.dim:
            raise ValueError("The row dimension of A doesn't match the dimension of the constraint set.")


def _sort_key(c):  # sort_sets, interface.jl:466-475
    for k, T in enumerate((ZeroSet, Nonnegatives, Box, SecondOrderCone, PsdCone, PsdConeTriangle)):
        if isinstance(c, T):
            return k + 1
    return 6


def assemble(P, q, constraints: Sequence[Constraint]):
    """assemble!, interface.jl:30-77: merge Zero / Nonneg sets (:411-460, merged
    set pushed to the end), stable sort by set type, A_model = -A, b_model = b
    (process_constraint!, :478-485).  Returns (P, q, A, b, cones)."""
    cons = list(constraints)
    for T in (ZeroSet, Nonnegatives):
        idx = [i for i, c in enumerate(cons) if type(c.convex_set) is T]
        if len(idx) > 1:
            A = sp.vstack([cons[i].A for i in idx], format="csr")
            b = np.concatenate([cons[i].b for i in idx])
            merged = Constraint(A, b, T(A.shape[0]))
            cons = [c for i, c in enumerate(cons) if i not in idx] + [merged]
    cons.sort(key=lambda c: _sort_key(c.convex_set))  # Julia sort! is stable
    n = len(np.atleast_1d(q))
    A = sp.vstack([-c.A for c in cons], format="csc") if cons else sp.csc_matrix((0, n))
    b = np.concatenate([c.b for c in cons]) if cons else np.zeros(0)
    Pm = sp.csc_matrix(P, dtype=float)
    return Pm, np.asarray(q, dtype=float).ravel(), A, b, [c.convex_set for c in cons]


def solve(P, q, A, b, cones, settings: Optional[Settings] = None, x0=None, s0=None, y0=None,
          record_history=False) -> Result:
    ws = Workspace(P, q, A, b, cones, settings or Settings())
    ws.warm_start(x0, s0, y0)
    return ws.optimize(record_history=record_history)
