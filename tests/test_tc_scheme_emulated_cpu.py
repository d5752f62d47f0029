"""NumPy models of the arithmetic of the tensor-core PSD path (csrc/tc_gemm.cuh, csrc/psd_tc.cuh) -- CPU only.

They pin the SCHEME (what the kernels are supposed to compute), not the kernels: the GPU tests compare the kernels with
dgemm / LAPACK.  Covered: the int8 slicing is an exact digit expansion, every group sum fits int32, the (slices, groups)
variants have the accuracy DESIGN.md quotes, and the capped minimax Newton-Schulz schedule converges to sign(X) without
ever letting an eigenvalue change sign."""
import numpy as np
import pytest


def slice_rows(M, k):
    """slice_rows_kernel: per row 2^e > max|x|, t = x 2^(6-e) in (-64, 64); the k digits are the balanced base-128
    digits of v = rn(t 2^(7k-7)), read off as the 7-bit fields of v + 64 sum_j 128^j (slice_fixed / slice_pack4)."""
    mx = np.max(np.abs(M), axis=1)
    e = np.where(mx > 0, np.frexp(np.where(mx > 0, mx, 1.0))[1], 0)
    scale = np.ldexp(1.0, e - 6)
    t = M * np.ldexp(1.0, 6 - e)[:, None]
    v = np.rint(t * 2.0 ** (7 * k - 7)).astype(np.int64)
    u = v + sum(64 << (7 * j) for j in range(k))
    assert np.all(u >= 0)
    digits = []
    for p in range(k):
        f = u >> (7 * (k - 1 - p))
        if p > 0:
            f = f & 127
        digits.append((f - 64).astype(np.int64))
    rem = t * 2.0 ** (7 * k - 7) - v              # what the last digit does not hold, in its own units
    return digits, scale, rem


def sliced_product(A, B, k, g):
    """ozaki_gemm_kernel: groups s = p + q <= g + 1, exact integer group sums, fp64 accumulation from the least
    significant group, then the row / column scales."""
    da, sa, _ = slice_rows(A, k)
    db, sb, _ = slice_rows(B, k)
    acc = np.zeros((A.shape[0], B.shape[0]))
    gmax = 0
    for s in range(g + 1, 1, -1):
        G = np.zeros_like(acc, dtype=np.int64)
        for p in range(1, k + 1):
            q = s - p
            if 1 <= q <= k:
                G += da[p - 1] @ db[q - 1].T
        gmax = max(gmax, int(np.max(np.abs(G))))
        acc += G.astype(np.float64) * 2.0 ** (-7 * (s - 2))
    return acc * sa[:, None] * sb[None, :], gmax


def test_slicing_is_an_exact_digit_expansion():
    rng = np.random.default_rng(0)
    M = rng.standard_normal((40, 300)) * np.logspace(-6, 6, 40)[:, None]      # rows of very different scale
    M[3, :] = 0.0                                                             # a zero row
    M[5, 7] = 1e-300                                                          # an entry far below its row maximum
    for k in (4, 7, 8):
        digits, scale, rem = slice_rows(M, k)
        assert all(np.max(np.abs(d)) <= 64 for d in digits)                   # int8 with a margin
        recon = sum(d * 128.0 ** (-p) for p, d in enumerate(digits)) * scale[:, None]
        err = np.abs(recon - M)
        rowmax = np.max(np.abs(M), axis=1)
        assert np.all(err <= 2.0 ** (-7 * k) * np.maximum(rowmax, 1e-300)[:, None] * 2.0000001)   # truncation only: half a unit of the last digit
        assert np.all(np.abs(rem) <= 0.5)
        assert all(np.max(d) <= 63 for d in digits[1:])                        # only the leading digit reaches +64


@pytest.mark.parametrize("k,g,bound", [(8, 10, 2e-15), (8, 8, 5e-15), (7, 7, 1e-12), (6, 8, 1e-11), (4, 6, 2e-7)])
def test_sliced_product_accuracy_and_int32_range(k, g, bound):
    rng = np.random.default_rng(k * 16 + g)
    N = 384
    Gm = rng.standard_normal((N, N))
    A = (Gm + Gm.T) / np.sqrt(2.0 * N)
    A[0, :] *= 1e-3
    A[:, 0] *= 1e-3
    B = A @ A
    B = (B + B.T) / 2
    got, gmax = sliced_product(A, B, k, g)       # B symmetric: rows of B are its columns, as on the device
    assert gmax < 2 ** 31                         # TMEM accumulators are int32
    assert 8 * 65535 * 64 * 64 < 2 ** 31          # ... for every K < 65536 (8 pairs per group at most; HBM bounds N far below)
    ref = A @ B
    assert np.max(np.abs(got - ref) / (np.abs(A) @ np.abs(B))) < bound


def test_product_counts_of_the_passes():
    def total(k, g):
        return sum(1 for s in range(2, g + 2) for p in range(1, k + 1) if 1 <= s - p <= k)
    assert (total(8, 10), total(8, 8), total(7, 7), total(6, 8), total(4, 6)) == (49, 36, 28, 30, 15)
    # pass p covers the groups g + 1 - 4p ... (4 accumulators): (8, 10) -> 28 + 18 + 3 products per K step
    def per_pass(k, g):
        out, s_hi = [], g + 1
        while s_hi >= 2:
            out.append(sum(1 for s in range(max(2, s_hi - 3), s_hi + 1) for p in range(1, k + 1) if 1 <= s - p <= k))
            s_hi -= 4
        return out
    assert per_pass(8, 10) == [28, 18, 3] and per_pass(8, 8) == [26, 10] and per_pass(4, 6) == [12, 3]


def _g(y):
    return 0.5 * y * (3.0 - y * y)


@pytest.mark.parametrize("spectrum", ["wigner", "cluster_at_zero", "graded", "one_sided"])
def test_capped_minimax_newton_schulz_schedule(spectrum):
    """ns_coef_kernel + the update product, in exact (fp64 eigenvalue) arithmetic: every eigenvalue keeps its sign, stays
    in [-1, 1] after every step, and the iteration ends at sign(lambda) within the step count the driver expects."""
    rng = np.random.default_rng(3)
    N = 400
    if spectrum == "wigner":
        lam = np.linalg.eigvalsh((lambda G: (G + G.T) / 2)(rng.standard_normal((N, N))))
    elif spectrum == "cluster_at_zero":
        lam = np.concatenate([rng.standard_normal(N // 2), 1e-9 * rng.standard_normal(N // 2)])
    elif spectrum == "graded":
        lam = np.logspace(0, -10, N) * np.where(np.arange(N) % 2 == 0, 1.0, -1.0)
    else:
        lam = np.abs(rng.standard_normal(N)) + 0.1
    sgn = np.sign(lam)
    s = lam / np.linalg.norm(lam)                 # S0 = X / |X|_F
    l, alpha_max, steps = 1e-7, 1.5, 0
    while steps < 80:
        f = np.sum(s ** 4)                        # |Y|_F^2, Y = S^2
        beta = f ** 0.25
        u = min(1.0, beta)
        alpha = min(alpha_max, np.sqrt(3.0 / (1.0 + l + l * l)))
        gamma = alpha / u
        l = min(_g(alpha * l), _g(alpha), 1.0)
        delta = 2.0 if beta < 1.0 else np.sqrt(np.mean((1.0 - s ** 2) ** 2))
        s = _g(gamma * s)                         # S' = 1.5 gamma S - 0.5 gamma^3 S^3
        steps += 1
        assert np.all(np.abs(s) <= 1.0 + 1e-12) and np.all(np.sign(s) == sgn)
        if delta < 1e-7:
            break
    big = np.abs(lam) > 1e-6 * np.max(np.abs(lam))
    assert np.max(np.abs(s[big] - sgn[big])) < 1e-12
    # projection error bound of the weighted residual: |lambda| (1 - |s|) / 2
    assert np.max(np.abs(lam) * (1.0 - np.abs(s)) / 2) <= (1e-9 if spectrum in ("cluster_at_zero", "graded") else 1e-13) * np.max(np.abs(lam))
    assert steps <= (80 if spectrum in ("cluster_at_zero", "graded") else 32)


def test_device_digit_extraction_functions_on_the_host(tmp_path):
    """slice_fixed / slice_pack4 are __host__ __device__: tests/slice_probe.cu runs the functions the kernel calls on
    3.2 million values (edge cases, 40 binades) for 8, 7, 6 and 4 slices."""
    import os
    import shutil
    import subprocess
    nvcc = shutil.which("nvcc") or ("/usr/local/cuda/bin/nvcc" if os.path.exists("/usr/local/cuda/bin/nvcc") else None)
    if nvcc is None:
        pytest.skip("no nvcc")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "slice_probe")
    subprocess.run([nvcc, "-std=c++17", "-I", os.path.join(root, "cosmo.jl_b200", "csrc"), "-gencode", "arch=compute_100a,code=sm_100a",
                    "-o", exe, os.path.join(root, "tests", "slice_probe.cu"), "-lcuda"], check=True, cwd=str(tmp_path))
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0 and "bad 0" in out.stdout, out.stdout + out.stderr
