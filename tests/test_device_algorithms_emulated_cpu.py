"""NumPy models of two device code paths that have not run on a GPU yet, kept as CPU tests so
that the claims "control flow validated in NumPy" / "index maps checked by emulation" stay reproducible:

* csrc/psd_sign.cuh -- Pi_+(X) = (X + sign(X) X) / 2 by Newton-Schulz steps with the rigorous scaling
  (|X|_F, then beta = |S^2|_F^(1/2) whenever beta < 1) and the stopping rules of PsdSign::project;
* psd_small_kernel, d.triangle == 2 -- load / store index maps of the real embedding of a Hermitian matrix.

The functions below follow the CUDA code statement by statement (same scalars, same tests, same order)."""
import math

import numpy as np

from oracle import cosmo_oracle as O


def sign_project(X, tol=1e-7, rtol=5e-13, cap=64):
    """PsdSign<double>::project"""
    N = len(X)
    fro = math.sqrt((X * X).sum())
    S = X * (1.0 / fro if fro > 0 else 0.0)                     # sg_scale_kernel
    prev, it, gemms, next_check, resid = 1e300, 0, 0, 24, None
    while True:
        T = S @ S                                               # sym_gemm_kernel<SG_SQ>
        gemms += 1
        f, d = (T * T).sum(), ((np.eye(N) - T) ** 2).sum()
        if not (f > 0):                                         # sg_delta_kernel
            delta, ib, ib2 = (0.0 if f == 0 else f), 1.0, 1.0
        else:
            beta = math.sqrt(math.sqrt(f))
            if beta < 1:
                delta, ib, ib2 = 2.0, 1.0 / beta, 1.0 / (beta * beta)
            else:
                delta, ib, ib2 = math.sqrt(d / N), 1.0, 1.0
        S = 0.5 * ib * (3.0 * S - ib2 * (S @ T))                # sym_gemm_kernel<SG_UPD>
        gemms += 1
        it += 1
        if delta != delta:
            return None, it, gemms, resid
        if delta < tol:
            W = S @ X
            gemms += 1
            break
        if (it >= next_check and delta > 0.98 * prev) or it >= cap:
            W = S @ X                                           # SG_MUL
            R = S @ W                                           # SG_RES
            gemms += 2
            resid = math.sqrt(((R - X) ** 2).sum()) / (fro if fro > 0 else 1.0)
            if resid < rtol or (it >= cap and resid < 1e3 * rtol):
                break
            if it >= cap:
                return None, it, gemms, resid                   # caller falls back to the eigensolver
            next_check = it + 8
        prev = delta
    return 0.5 * (X + W), it, gemms, resid                      # sg_store_kernel


def _ref(X):
    w, V = np.linalg.eigh(X)
    return (V * np.maximum(w, 0)) @ V.T


def test_sign_function_projection_control_flow():
    rng = np.random.default_rng(0)
    B = rng.standard_normal((200, 200))
    v = rng.standard_normal(120)
    Q, _ = np.linalg.qr(rng.standard_normal((150, 150)))
    spec = lambda lam: (Q * lam) @ Q.T   # noqa: E731
    geo = 10.0 ** -np.arange(0, 15, 0.2)[:75]
    cases = {
        "wigner": ((B + B.T) / 2, 30, 2e-14),
        "zero": (np.zeros((50, 50)), 1, 0.0),
        "rank1_pos": (np.outer(v, v), 24, 2e-14),
        "rank1_neg": (-np.outer(v, v), 24, 2e-14),
        "zeros_in_spectrum": (spec(np.concatenate([np.linspace(1, 2, 50), np.zeros(50), -np.linspace(0.5, 3, 50)])), 24, 2e-14),
        "identity": (3.0 * np.eye(64), 10, 2e-14),
        "one_huge": (spec(np.concatenate([[1e6], rng.standard_normal(149)])), 60, 2e-14),
        "one_tiny": (spec(np.concatenate([np.linspace(1, 2, 75), -np.linspace(1, 2, 74), [1e-9]])), 64, 2e-14),
        "geometric_to_1e-15": (spec(np.concatenate([geo, -geo])), 64, 1e-11),
    }
    for name, (X, max_steps, bound) in cases.items():
        X = (X + X.T) / 2
        P, steps, gemms, resid = sign_project(X)
        assert P is not None, name
        nrm = np.linalg.norm(X) or 1.0
        assert np.linalg.norm(P - _ref(X)) / nrm <= bound, (name, np.linalg.norm(P - _ref(X)) / nrm)
        assert steps <= max_steps and gemms <= 2 * steps + 1 + 2 * max((steps - 24) // 8 + 2, 0), (name, steps, gemms)
        assert np.allclose(P, P.T, atol=1e-12 * nrm)


def _svec_pos(i, j):
    return j * (j + 1) // 2 + i


def kernel_load_complex(x, Nc):
    """psd_small_kernel, load branch d.triangle == 2"""
    N, inv = 2 * Nc, 1.0 / math.sqrt(2.0)
    M = np.zeros((N, N))
    for j in range(N):
        for i in range(N):
            I, bi, J, bj = i % Nc, i // Nc, j % Nc, j // Nc
            a, b = (I, J) if I < J else (J, I)
            if bi == bj:
                v = x[_svec_pos(a, b)]
                if a != b:
                    v *= inv
            elif I == J:
                v = 0.0
            else:
                im_ab = x[Nc * (Nc + 1) // 2 + b * (b - 1) // 2 + a] * inv
                b_IJ = im_ab if I < J else -im_ab
                v = b_IJ if bi == 1 else -b_IJ
            M[i, j] = v
    return M


def kernel_store_complex(V, Nc):
    """psd_small_kernel, store branch d.triangle == 2 (V already scaled by sqrt(max(lambda, 0)))"""
    N, tri, sqrt2 = 2 * Nc, Nc * (Nc + 1) // 2, math.sqrt(2.0)
    s = np.zeros(Nc * Nc)
    for e in range(Nc * Nc):
        imag = e >= tri
        ee = e - tri if imag else e
        if not imag:
            j = int((math.sqrt(8.0 * ee + 1.0) - 1.0) * 0.5)
            while (j + 1) * (j + 2) // 2 <= ee:
                j += 1
            while j * (j + 1) // 2 > ee:
                j -= 1
            i = ee - j * (j + 1) // 2
            acc = 0.5 * sum(V[i, k] * V[j, k] + V[Nc + i, k] * V[Nc + j, k] for k in range(N))
            s[e] = acc if i == j else sqrt2 * acc
        else:
            j = int((math.sqrt(8.0 * ee + 1.0) + 1.0) * 0.5)
            while j * (j + 1) // 2 <= ee:
                j += 1
            while j * (j - 1) // 2 > ee:
                j -= 1
            i = ee - j * (j - 1) // 2
            acc = sum(V[Nc + i, k] * V[j, k] - V[i, k] * V[Nc + j, k] for k in range(N))
            s[e] = sqrt2 * 0.5 * acc
    return s


def test_complex_psd_embedding_index_maps():
    rng = np.random.default_rng(4)
    for Nc in (2, 3, 7):
        Z = rng.standard_normal((Nc, Nc)) + 1j * rng.standard_normal((Nc, Nc))
        X = (Z + Z.conj().T) / 2
        x = O.extract_upper_triangle_complex(X, math.sqrt(2.0))
        M = kernel_load_complex(x, Nc)
        assert np.allclose(M, np.block([[X.real, -X.imag], [X.imag, X.real]]), atol=1e-15)
        w, Q = np.linalg.eigh(M)
        s = kernel_store_complex(Q * np.sqrt(np.maximum(w, 0)), Nc)
        ref = x.copy()
        O.project_cone(ref, O.ComplexPsdConeTriangle(Nc * Nc))
        assert np.max(np.abs(s - ref)) < 1e-13
