// psd_sign.cuh -- EXPERIMENTAL (off by default, COSMO_B200_PSD_SIGN=1): projection of a large symmetric matrix
// onto the PSD cone without an eigendecomposition,
//
//     Pi_+(X) = (X + sign(X) X) / 2 ,      sign(X) by Newton-Schulz:  S <- S (3 I - S^2) / 2 ,  S_0 = X / |X|_F
//
// The iteration keeps the sign of every eigenvalue only while rho(S) < sqrt 3, so the scaling must be a rigorous
// upper bound of the spectral radius: |X|_F at the start, and in every step rho(S)^2 <= |S^2|_F, which is available
// for free from the product the step needs anyway -- whenever beta = |S^2|_F^(1/2) < 1 the step uses S / beta
// (tightening the bound from a factor N^(1/2) to N^(1/4) and so on) without an extra product.
//
// (prototype profiles/notes/psd_sign_prototype.py: 21-33 steps give a relative error of
// 1e-15 against LAPACK on the iterates of config C4).  Every step is two N x N x N products of symmetric,
// commuting matrices -- GEMM-shaped work instead of the rotation sweeps of the block-Jacobi solver.  Eigenvalues
// with |lambda| < 1.5^-k |X| have not reached +-1 after k steps, but they enter the projection with weight
// |lambda| only, so the iteration is capped instead of waiting for them.
//
// Status: validated on B200 in round 2 (profiles/psd_sign_timing_r2_fp64fma.log).  Reachable only through the
// environment switch; falls back to the block-Jacobi path when the iteration misbehaves.
// COSMO_B200_PSD_SIGN=1 selects it: the same iteration as psd_tc.cuh with the products on the FP64 FMA pipe -- kept as
// the comparator of the tensor-core path (measured on B200, round 2: N = 2000, 24-32 steps, 42-54 ms, 5e-14).
//
// Included from psd.cuh (after PsdConeDesc, svec_pos and bj_load8, before PsdBatch).
#pragma once

namespace cosmo {

enum { SG_SQ = 0, SG_UPD = 1, SG_MUL = 2, SG_RES = 3 };

// tile index -> (bi, bj), bi <= bj, tiles of the upper triangle enumerated column by column
__device__ __forceinline__ void sg_tile(int t, int& bi, int& bj) {
  int j = (int)((sqrt(8.0 * (double)t + 1.0) - 1.0) / 2.0);
  while ((long long)(j + 1) * (j + 2) / 2 <= t) ++j;
  while ((long long)j * (j + 1) / 2 > t) --j;
  bj = j;
  bi = t - j * (j + 1) / 2;
}

// C = epi(A B) for symmetric, commuting N x N matrices A and B (column-major): the product is symmetric, so only
// the upper triangle is computed (128 x 128 CTA tiles, 8 x 8 register micro-tiles, K in chunks of 16 through shared
// memory) and mirrored on store -- C is exactly symmetric.  Both operand tiles are read "rows of a column block",
// which is contiguous for A as stored and, by symmetry, for B as well.
//   SG_SQ  : C = A B (= T),  partial[2 block] = sum_ij T_ij^2,  partial[2 block + 1] = sum_ij (delta_ij - T_ij)^2
//   SG_UPD : C = sc[2] (3 A - sc[3] A B) / 2        (A = S, B = T;  sc[2] = 1/beta, sc[3] = 1/beta^2, or 1, 1)
//   SG_MUL : C = A B
//   SG_RES : nothing stored,  partial[2 block] = sum_ij ((A B)_ij - X_ij)^2        (weighted residual |S^2 X - X|_F^2)
template <typename T, int EPI>
__global__ void __launch_bounds__(kBlock, 1) sym_gemm_kernel(int N, const T* __restrict__ A, const T* __restrict__ B,
                                                             const T* __restrict__ X, T* __restrict__ C, T* __restrict__ partial,
                                                             const T* __restrict__ sc) {
  constexpr int TM = 128, TK = 16, LDS_ = TM + 2;
  __shared__ __align__(16) T As[TK][LDS_];   // As[kk][i] = A[row0 + i][k0 + kk]
  __shared__ __align__(16) T Bs[TK][LDS_];   // Bs[kk][j] = B[k0 + kk][col0 + j] = B[col0 + j][k0 + kk]
  __shared__ T red[kWarpsPerBlock][2];
  int bi, bj;
  sg_tile((int)blockIdx.x, bi, bj);
  const T inv_b = (EPI == SG_UPD) ? sc[2] : T(1), inv_b2 = (EPI == SG_UPD) ? sc[3] : T(1);
  const int row0 = bi * TM, col0 = bj * TM;
  const int ti = (threadIdx.x % 16) * 8, tj = (threadIdx.x / 16) * 8;
  T acc[8][8];
#pragma unroll
  for (int a = 0; a < 8; ++a)
#pragma unroll
    for (int b = 0; b < 8; ++b) acc[a][b] = T(0);
  for (int k0 = 0; k0 < N; k0 += TK) {
    for (int e = threadIdx.x; e < TK * TM; e += blockDim.x) {
      const int i = e % TM, kk = e / TM;
      const int gk = k0 + kk;
      const int ga = row0 + i, gb = col0 + i;
      As[kk][i] = (ga < N && gk < N) ? A[ga + (long long)gk * N] : T(0);
      Bs[kk][i] = (gb < N && gk < N) ? B[gb + (long long)gk * N] : T(0);
    }
    __syncthreads();
#pragma unroll 4
    for (int kk = 0; kk < TK; ++kk) {
      T av[8], bv[8];
      bj_load8(&As[kk][ti], av);
      bj_load8(&Bs[kk][tj], bv);
#pragma unroll
      for (int a = 0; a < 8; ++a)
#pragma unroll
        for (int b = 0; b < 8; ++b) acc[a][b] += av[a] * bv[b];
    }
    __syncthreads();
  }
  T dsum = T(0), fsum = T(0);
#pragma unroll
  for (int b = 0; b < 8; ++b) {
    const int gj = col0 + tj + b;
#pragma unroll
    for (int a = 0; a < 8; ++a) {
      const int gi = row0 + ti + a;
      if (gi >= N || gj >= N || gi > gj) continue;     // upper triangle only; the mirror is written below
      const T p = acc[a][b];
      T out;
      if (EPI == SG_SQ) {
        const T d = ((gi == gj) ? T(1) : T(0)) - p;
        dsum += (gi == gj) ? d * d : T(2) * d * d;
        fsum += (gi == gj) ? p * p : T(2) * p * p;
        out = p;
      } else if (EPI == SG_UPD) {
        out = T(0.5) * inv_b * (T(3) * A[gi + (long long)gj * N] - inv_b2 * p);
      } else if (EPI == SG_MUL) {
        out = p;
      } else {
        const T d = p - X[gi + (long long)gj * N];
        fsum += (gi == gj) ? d * d : T(2) * d * d;
        continue;
      }
      C[gi + (long long)gj * N] = out;
      if (gi != gj) C[gj + (long long)gi * N] = out;
    }
  }
  if (EPI == SG_SQ || EPI == SG_RES) {
    dsum = warp_sum(dsum);
    fsum = warp_sum(fsum);
    if ((threadIdx.x & 31) == 0) { red[threadIdx.x >> 5][0] = fsum; red[threadIdx.x >> 5][1] = dsum; }
    __syncthreads();
    if (threadIdx.x == 0) {
      T f = T(0), dd = T(0);
      for (int w = 0; w < kWarpsPerBlock; ++w) { f += red[w][0]; dd += red[w][1]; }
      partial[2 * blockIdx.x] = f;
      partial[2 * blockIdx.x + 1] = dd;
    }
  }
}

// After T = S^2:  beta = |T|_F^(1/2) >= rho(S).  If beta < 1 the step rescales (sc[2] = 1/beta, sc[3] = 1/beta^2) and the
// convergence measure is void; otherwise sc[2] = sc[3] = 1 and sc[1] = delta = root mean square of the eigenvalues of
// I - S^2.  T = 0 (X = 0) counts as converged.
template <typename T>
__global__ void sg_delta_kernel(const T* __restrict__ partial, int nparts, int N, T* __restrict__ sc) {
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    T f = T(0), d = T(0);
    for (int i = 0; i < nparts; ++i) { f += partial[2 * i]; d += partial[2 * i + 1]; }
    const T beta = sqrt(sqrt(f));
    if (!(f > T(0))) { sc[1] = (f == T(0)) ? T(0) : f; sc[2] = T(1); sc[3] = T(1); }   // zero matrix, or NaN handed to the host
    else if (beta < T(1)) { sc[1] = T(2); sc[2] = T(1) / beta; sc[3] = T(1) / (beta * beta); }
    else { sc[1] = sqrt(d / (T)N); sc[2] = T(1); sc[3] = T(1); }
  }
}

// S = X / |X|_F with |X|_F^2 = sum of the partial sums the load kernel left behind (S = 0 when X = 0)
template <typename T>
__global__ void __launch_bounds__(kBlock) sg_scale_kernel(int N, const T* __restrict__ X, const T* __restrict__ fro_partials, int nparts,
                                                          T* __restrict__ S) {
  __shared__ T sc_s;
  if (threadIdx.x == 0) {
    T f = T(0);
    for (int i = 0; i < nparts; ++i) f += fro_partials[i];
    sc_s = (f > T(0)) ? T(1) / sqrt(f) : T(0);
  }
  __syncthreads();
  const T sc = sc_s;
  const long long total = (long long)N * N;
  for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < total; k += (long long)gridDim.x * blockDim.x) S[k] = X[k] * sc;
}

// r = |S^2 X - X|_F / |X|_F  -> sc[0]   (|X|_F^2 from the load kernel's partial sums)
template <typename T>
__global__ void sg_residual_kernel(const T* __restrict__ partial, int nparts, const T* __restrict__ fro_partials, int nfro,
                                   T* __restrict__ sc) {
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    T f = T(0), x2 = T(0);
    for (int i = 0; i < nparts; ++i) f += partial[2 * i];
    for (int i = 0; i < nfro; ++i) x2 += fro_partials[i];
    sc[0] = (x2 > T(0)) ? sqrt(f / x2) : T(0);
  }
}

// s[cone] = svec / square layout of P = (X + W) / 2, W = sign(X) X   (same store rule as psd_large_syrk_kernel)
template <typename T>
__global__ void __launch_bounds__(kBlock) sg_store_kernel(PsdConeDesc d, const T* __restrict__ X, const T* __restrict__ W,
                                                          T* __restrict__ s) {
  const int N = d.N;
  const T sqrt2 = T(1.41421356237309504880);
  const long long total = (long long)N * N;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(e % N), j = (int)(e / N);
    if (i > j) continue;
    const T v = T(0.5) * (X[i + (long long)j * N] + W[i + (long long)j * N]);
    if (d.triangle) {
      s[d.off + svec_pos(i, j)] = (i == j) ? v : sqrt2 * v;
    } else {
      s[d.off + (long long)j * N + i] = v;
      s[d.off + (long long)i * N + j] = v;   // mirror, convexset.jl:316-318
    }
  }
}

// Host driver.  X (N x N symmetric) and its Frobenius partial sums come from psd_large_load_kernel; S0 is a second N x N
// buffer of the caller.  Returns false when the iteration misbehaved (NaN) or memory ran out -- the caller then falls
// back to the eigensolver.
template <typename T>
struct PsdSign {
  T *S1_d = nullptr, *U_d = nullptr, *sc_d = nullptr, *part_d = nullptr;
  T* sc_h = nullptr;   // pinned mirror of sc_d: [1] = delta, [2] = 1/beta, [3] = 1/beta^2
  T* S1_scratch = nullptr;
  int capN = 0;
  int last_steps = 0;
  ~PsdSign() {
    cudaFree(S1_d); cudaFree(U_d); cudaFree(sc_d); cudaFree(part_d);
    if (sc_h) cudaFreeHost(sc_h);
  }
  static bool enabled() {
    const char* e = getenv("COSMO_B200_PSD_SIGN");
    return e && e[0] == '1';
  }
  bool ensure(int N) {
    if (N <= capN) return true;
    cudaFree(S1_d); cudaFree(U_d); cudaFree(part_d);
    S1_d = U_d = part_d = nullptr;
    const size_t nn = (size_t)N * N;
    const int nt = (N + 127) / 128;
    bool ok = cudaMalloc(&S1_d, nn * sizeof(T)) == cudaSuccess && cudaMalloc(&U_d, nn * sizeof(T)) == cudaSuccess &&
              cudaMalloc(&part_d, (size_t)std::max(nt * (nt + 1), 2 * kMaxGrid) * sizeof(T)) == cudaSuccess;
    if (ok && !sc_d) ok = cudaMalloc(&sc_d, 4 * sizeof(T)) == cudaSuccess && cudaMallocHost(&sc_h, 4 * sizeof(T)) == cudaSuccess;
    capN = ok ? N : 0;
    return ok;
  }

  bool project(const PsdConeDesc& d, const T* X_d, const T* fro_partials, int nfro, T* S0_d, T* s_out, cudaStream_t st,
               long long& launches) {
    const int N = d.N;
    if (!ensure(N)) return false;
    const int g = (int)std::min<long long>(((long long)N * N + kBlock - 1) / kBlock, kMaxGrid);
    sg_scale_kernel<T><<<g, kBlock, 0, st>>>(N, X_d, fro_partials, nfro, S0_d);
    ++launches;
    const int nt = (N + 127) / 128;
    const int ntiles = nt * (nt + 1) / 2;
    const int nparts = ntiles;
    bool ok = true;
    // C = epi(A B); X is the third operand of SG_RES, sc the scalars of SG_UPD
    auto product = [&](int epi, const T* A, const T* B, const T* Xo, T* C) {
      if (epi == SG_SQ) sym_gemm_kernel<T, SG_SQ><<<ntiles, kBlock, 0, st>>>(N, A, B, Xo, C, part_d, (const T*)nullptr);
      else if (epi == SG_UPD) sym_gemm_kernel<T, SG_UPD><<<ntiles, kBlock, 0, st>>>(N, A, B, Xo, C, (T*)nullptr, sc_d);
      else if (epi == SG_MUL) sym_gemm_kernel<T, SG_MUL><<<ntiles, kBlock, 0, st>>>(N, A, B, Xo, C, (T*)nullptr, (const T*)nullptr);
      else sym_gemm_kernel<T, SG_RES><<<ntiles, kBlock, 0, st>>>(N, A, B, Xo, (T*)nullptr, part_d, (const T*)nullptr);
      ++launches;
    };
    const double tol = sizeof(T) == 8 ? 1e-7 : 3e-4;   // quadratic convergence: the step after delta < tol reaches ~delta^2
    const double rtol = sizeof(T) == 8 ? 5e-13 : 2e-5;  // accepted weighted residual |S^2 X - X|_F / |X|_F
    T* S = S0_d;
    T* Sn = S1_d;
    double prev = 1e300, resid = -1.0;
    int it = 0, next_check = 24;
    const int cap = 64;
    bool have_W = false;
    for (;;) {
      S1_scratch = Sn;                                          // free until the update below overwrites it
      product(SG_SQ, S, S, (const T*)nullptr, U_d);
      sg_delta_kernel<T><<<1, 32, 0, st>>>(part_d, nparts, N, sc_d);
      product(SG_UPD, S, U_d, (const T*)nullptr, Sn);
      ++launches;
      if (!ok) return false;
      if (cudaMemcpyAsync(sc_h, sc_d, 4 * sizeof(T), cudaMemcpyDeviceToHost, st) != cudaSuccess) return false;
      if (cudaStreamSynchronize(st) != cudaSuccess) return false;
      std::swap(S, Sn);
      ++it;
      const double delta = (double)sc_h[1];
      if (!(delta == delta)) return false;                      // NaN in the input: let the eigensolver report it
      if (delta < tol) break;
      // Eigenvalues that are (numerically) zero never reach +-1, and they do not have to: they enter the projection
      // with weight |lambda|.  Once delta stalls, test the weighted residual of the candidate  W = S X:
      //   |S W - X|_F = |(S^2 - I) X|_F = (sum lambda_i^2 (1 - s_i^2)^2)^(1/2)   (twice an upper bound of the error)
      if ((it >= next_check && delta > 0.98 * prev) || it >= cap) {
        S1_scratch = Sn;                                        // S was just swapped: Sn is the stale iterate
        product(SG_MUL, S, X_d, (const T*)nullptr, U_d);
        product(SG_RES, S, U_d, X_d, (T*)nullptr);
        sg_residual_kernel<T><<<1, 32, 0, st>>>(part_d, nparts, fro_partials, nfro, sc_d);
        ++launches;
        if (!ok) return false;
        if (cudaMemcpyAsync(sc_h, sc_d, sizeof(T), cudaMemcpyDeviceToHost, st) != cudaSuccess) return false;
        if (cudaStreamSynchronize(st) != cudaSuccess) return false;
        resid = (double)sc_h[0];
        if (!(resid == resid)) return false;
        if (resid < rtol || (it >= cap && resid < 1e3 * rtol)) { have_W = true; break; }
        if (it >= cap) return false;                            // a cluster of eigenvalues at ~1e-10 |X|: use the eigensolver
        next_check = it + 8;
      }
      prev = delta;
    }
    last_steps = it;
    if (!have_W) {
      product(SG_MUL, S, X_d, (const T*)nullptr, U_d);
      if (!ok) return false;
    }
    sg_store_kernel<T><<<g, kBlock, 0, st>>>(d, X_d, U_d, s_out);
    ++launches;
    if (getenv("COSMO_B200_PSD_DEBUG")) fprintf(stderr, "[psd-sign] N=%d steps=%d delta=%g resid=%g\n", N, it, (double)sc_h[1], resid);
    return cudaGetLastError() == cudaSuccess;
  }
};

}  // namespace cosmo
