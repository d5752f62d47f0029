// aa.cuh -- safeguarded Anderson acceleration of the ADMM operator:
// AndersonAccelerator{T, Type2{QRDecomp}, RestartedMemory, NoRegularizer}, the reference's default
// accelerator (settings.jl:136-138), driven through acceleration_pre!/post! (accelerator_interface.jl:58-114).
//
// The arithmetic lives in COSMOAccelerators.jl (not part of the reference tree); it is restated here from the
// published method -- type-II Anderson acceleration, least squares min |f - F eta| by a QR factorisation of
// F = [f_k - f_{k-1}] that is extended one column per iteration (modified Gram-Schmidt), memory emptied when
// its `mem` columns are full.  Iterate-level parity with the package is therefore UNPINNED; the CPU oracle
// holds the same restatement and the reference's behavioural tests (AccelerationTests) are reproduced.
//
// All vectors have the length of the operator variable w = [w_x; w_s].  On a row-sharded run w_x is replicated:
// inner products run over [lo, dim) with lo = 0 on rank 0 and lo = n elsewhere, followed by an allreduce, so
// every rank holds the same R and eta and applies the same update to its copy of w_x.
#pragma once
#include "common.cuh"

namespace cosmo {

enum { AA_F2 = 0, AA_FACC2 = 1, AA_FLAG = 2, AA_NRM2 = 3, AA_SC_COUNT = 8 };

// CA.update!: f = x - g; first call after a restart only stores (g, f); otherwise the new columns
//   G[:, j] = g - g_last,  Q[:, j] = f - f_last (orthogonalised afterwards),  then g_last = g, f_last = f.
// out[0] = |f|^2 over [lo, dim)  (the safeguard's reference norm, accelerator_interface.jl:90).
template <typename T>
__global__ void __launch_bounds__(kBlock) aa_update_kernel(int dim, int lo, const T* __restrict__ g, const T* __restrict__ x,
                                                           T* __restrict__ f, T* __restrict__ f_last, T* __restrict__ g_last,
                                                           T* __restrict__ Gj, T* __restrict__ Qj, int init, RedBuf<T> rb) {
  T accS[1] = {0};
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < dim; i += gridDim.x * blockDim.x) {
    const T gi = g[i];
    const T fi = x[i] - gi;
    f[i] = fi;
    if (!init) {
      Gj[i] = gi - g_last[i];
      Qj[i] = fi - f_last[i];
    }
    g_last[i] = gi;
    f_last[i] = fi;
    if (i >= lo) accS[0] += fi * fi;
  }
  reduce_and_finalize<T, 1, 0>(accS, (const T*)nullptr, rb, NoFin());
}

// One modified Gram-Schmidt step on the new column q:
//   q -= r_prev * Qp   (skipped when Qp == nullptr),   out[0] = <Qi, q>  (or |q|^2 when Qi == nullptr)
template <typename T>
__global__ void __launch_bounds__(kBlock) aa_mgs_kernel(int dim, int lo, T* __restrict__ q, const T* __restrict__ Qp,
                                                        const T* __restrict__ r_prev, const T* __restrict__ Qi, RedBuf<T> rb) {
  T accS[1] = {0};
  const T r = Qp ? *r_prev : T(0);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < dim; i += gridDim.x * blockDim.x) {
    T qi = q[i];
    if (Qp) { qi -= r * Qp[i]; q[i] = qi; }
    if (i >= lo) accS[0] += (Qi ? Qi[i] : qi) * qi;
  }
  reduce_and_finalize<T, 1, 0>(accS, (const T*)nullptr, rb, NoFin());
}

// R[j,j] = sqrt(|q|^2) ; q /= R[j,j]
template <typename T>
__global__ void __launch_bounds__(kBlock) aa_normalize_kernel(int dim, T* __restrict__ q, const T* __restrict__ nrm2, T* __restrict__ rjj) {
  const T r = sqrt(*nrm2);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < dim; i += gridDim.x * blockDim.x) q[i] = q[i] / r;
  if (blockIdx.x == 0 && threadIdx.x == 0) *rjj = r;
}

// out[c] = <Q[:, c0 + c], f>, c < ncols <= 8
template <typename T>
__global__ void __launch_bounds__(kBlock) aa_qtf_kernel(int dim, int lo, const T* __restrict__ f, const T* __restrict__ Q, size_t ld,
                                                        int c0, int ncols, RedBuf<T> rb) {
  T accS[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int i = lo + blockIdx.x * blockDim.x + threadIdx.x; i < dim; i += gridDim.x * blockDim.x) {
    const T fi = f[i];
#pragma unroll
    for (int c = 0; c < 8; ++c)
      if (c < ncols) accS[c] += Q[(size_t)(c0 + c) * ld + i] * fi;
  }
  reduce_and_finalize<T, 8, 0>(accS, (const T*)nullptr, rb, NoFin());
}

// Back substitution R[0:l,0:l] eta = Q'f (in place; R column-major with leading dimension mem) and the
// acceptance test of the candidate: finite, non-singular, |eta|_2 <= 1e4.  flag[0] = 1 accepted / 0 rejected.
template <typename T>
__global__ void aa_solve_kernel(const T* __restrict__ R, int mem, int l, T* __restrict__ eta, T* __restrict__ flag) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  bool ok = true;
  for (int i = l - 1; i >= 0; --i) {
    T v = eta[i];
    for (int k = i + 1; k < l; ++k) v -= R[(size_t)k * mem + i] * eta[k];
    const T d = R[(size_t)i * mem + i];
    if (d == T(0) || !isfinite(d)) ok = false;
    eta[i] = v / d;
  }
  T nrm2 = 0;
  for (int i = 0; i < l; ++i) nrm2 += eta[i] * eta[i];
  const T nrm = sqrt(nrm2);
  if (!isfinite(nrm) || nrm > T(1e4)) ok = false;
  flag[0] = ok ? T(1) : T(0);
}

// CA.accelerate!: w -= G[:, 0:l] eta, only if the candidate was accepted
template <typename T>
__global__ void __launch_bounds__(kBlock) aa_apply_kernel(int dim, T* __restrict__ w, const T* __restrict__ G, size_t ld, int l,
                                                          const T* __restrict__ eta, const T* __restrict__ flag) {
  if (*flag == T(0)) return;
  __shared__ T e[32];
  if (threadIdx.x < l) e[threadIdx.x] = eta[threadIdx.x];
  __syncthreads();
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < dim; i += gridDim.x * blockDim.x) {
    T v = w[i];
    for (int c = 0; c < l; ++c) v -= G[(size_t)c * ld + i] * e[c];
    w[i] = v;
  }
}

// compute_accelerated_res_norm! (accelerator_interface.jl:120-123): f = w_prev - w, out[0] = |f|^2
template <typename T>
__global__ void __launch_bounds__(kBlock) aa_res_kernel(int dim, int lo, const T* __restrict__ w_prev, const T* __restrict__ w,
                                                        T* __restrict__ f, RedBuf<T> rb) {
  T accS[1] = {0};
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < dim; i += gridDim.x * blockDim.x) {
    const T fi = w_prev[i] - w[i];
    f[i] = fi;
    if (i >= lo) accS[0] += fi * fi;
  }
  reduce_and_finalize<T, 1, 0>(accS, (const T*)nullptr, rb, NoFin());
}

}  // namespace cosmo
