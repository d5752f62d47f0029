// cone3.cuh -- the 3-dimensional cones: ExponentialCone, PowerCone(alpha) and their duals
// (src/convexset.jl:497-789).  One thread per cone: each projection is a short scalar root
// search (bisection + Newton for K_exp, Newton for K_pow) that follows the reference's
// iteration step by step, so iterates agree with the CPU to the tolerance of the search
// (EXP_TOL / POW_TOL = 1e-8) -- exp/log/pow of the device maths library differ from libm in
// the last ulp only.  The dual cones use Moreau's identity Pi_K*(v) = v + Pi_K(-v)
// (convexset.jl:784-789).  Rows of these cones are written to s before proj_rhs_kernel runs
// (like the PSD rows), which then only builds the right-hand side from them.
#pragma once
#include "common.cuh"

namespace cosmo {

enum : unsigned char { C3_EXP = 0, C3_DUAL_EXP = 1, C3_POW = 2, C3_DUAL_POW = 3 };

template <typename T>
struct Cone3Table {
  int n;                       // number of 3-d cones on this rank
  const int* off;              // first row
  const unsigned char* kind;   // C3_*
  const T* alpha;              // power cones
  const int* max_iter;
  const T* tol;
};

// ---- membership (convexset.jl:600-612, 719-734) -----------------------------------
template <typename T>
__device__ __forceinline__ bool exp_in_cone(T x, T y, T z, T tol) {
  return (y > T(0) && y * exp(x / y) <= z + tol) || (x <= tol && y == T(0) && z >= -tol);
}
template <typename T>
__device__ __forceinline__ bool exp_in_dual(T x, T y, T z, T tol) {
  return (x < T(0) && -x * exp(y / x) - T(2.718281828459045) * z <= tol) ||
         (fabs(x) <= tol && y >= -tol && z >= -tol);
}
template <typename T>
__device__ __forceinline__ bool pow_in_cone(T x, T y, T z, T a, T tol) {
  return x >= T(0) && y >= T(0) && pow(x, a) * pow(y, T(1) - a) >= fabs(z) - tol;
}
template <typename T>
__device__ __forceinline__ bool pow_in_dual(T s, T t, T w, T a, T tol) {
  // a negative base (only reachable inside the tol band) is a DomainError in the reference: treated as "not in"
  if (!(s >= -tol && t >= -tol) || s < T(0) || t < T(0)) return false;
  return pow(s, a) * pow(t, T(1) - a) >= fabs(w) * pow(a, a) * pow(T(1) - a, T(1) - a) - tol;
}

// ---- K_exp projection (convexset.jl:510-597) ----------------------------------------
// Newton on dt = t - t0 for the minimiser over t at fixed multiplier lambda
template <typename T>
__device__ __forceinline__ T exp_find_min_t(T lam, T s0, T t0, T tol) {
  T dt = (-t0 > tol) ? -t0 : tol;
  for (int k = 0; k < 150; ++k) {
    const T f = dt * (dt + t0) / (lam * lam) - s0 / lam + log(dt / lam) + T(1);
    const T g = (T(2) * dt + t0) / (lam * lam) + T(1) / dt;
    dt = dt - f / g;
    if (dt <= -t0) { dt = -t0; break; }
    else if (dt <= T(0)) { dt = T(0); break; }
    else if (fabs(f) < tol) break;
  }
  return dt + t0;
}

// grad_dual! + find_minimizers!: updates v, returns the dual gradient
template <typename T>
__device__ __forceinline__ T exp_grad_dual(T lam, T* v, const T* v0, T tol) {
  v[2] = exp_find_min_t(lam, v0[1], v0[2], tol);
  v[1] = (T(1) / lam) * (v[2] - v0[2]) * v[2];
  v[0] = v0[0] - lam;
  return (v[1] == T(0)) ? v[0] : v[0] + v[1] * log(v[1] / v[2]);
}

template <typename T>
__device__ void project_exp(T* v, int max_iter, T tol) {
  if (exp_in_cone(v[0], v[1], v[2], T(0))) return;
  if (exp_in_dual(-v[0], -v[1], -v[2], T(0))) { v[0] = v[1] = v[2] = T(0); return; }
  if (v[0] < T(0) && v[1] < T(0)) { v[1] = T(0); v[2] = v[2] > T(0) ? v[2] : T(0); return; }
  const T v0[3] = {v[0], v[1], v[2]};
  // get_bisection_bounds
  T l = T(0), lam = T(0.125);
  T g = exp_grad_dual(lam, v, v0, tol);
  int guard = 0;
  while (g > T(0) && guard++ < 2000) {   // lambda doubles: the reference loop is unbounded, 2^2000 is not reachable
    l = lam;
    lam *= T(2);
    g = exp_grad_dual(lam, v, v0, tol);
  }
  T u = lam;
  for (int k = 0; k < max_iter; ++k) {
    lam = (u + l) / T(2);
    g = exp_grad_dual(lam, v, v0, tol);
    if (g > T(0)) l = lam; else u = lam;
    if (u - l < tol) break;
  }
}

// ---- K_pow projection (convexset.jl:646-713) ----------------------------------------
template <typename T>
__device__ __forceinline__ T pow_phic(T c0, T az, T r, T al) {
  const T v = T(0.5) * (c0 + sqrt(c0 * c0 + T(4) * al * r * (az - r)));
  return v > T(1e-10) ? v : T(1e-10);
}

template <typename T>
__device__ void project_pow(T* v, T a, int max_iter, T tol) {
  if (pow_in_cone(v[0], v[1], v[2], a, T(0))) return;
  if (pow_in_dual(-v[0], -v[1], -v[2], a, T(0))) { v[0] = v[1] = v[2] = T(0); return; }
  if (fabs(v[2]) <= tol) {
    v[0] = v[0] > T(0) ? v[0] : T(0);
    v[1] = v[1] > T(0) ? v[1] : T(0);
    return;
  }
  const T x0 = v[0], y0 = v[1], z0 = v[2], az = fabs(v[2]);
  T r = az / T(2), px = T(0), py = T(0);
  for (int k = 0; k < max_iter; ++k) {
    px = pow_phic(x0, az, r, a);
    py = pow_phic(y0, az, r, T(1) - a);
    const T pw = pow(px, a) * pow(py, T(1) - a);
    const T phi = pw - r;
    if (fabs(phi) < tol) break;
    const T dpx = a / (T(2) * px - x0) * (az - T(2) * r);
    const T dpy = (T(1) - a) / (T(2) * py - y0) * (az - T(2) * r);
    const T dphi = pw * (a * dpx / px + (T(1) - a) * dpy / py) - T(1);
    r = r - phi / dphi;
    r = r > T(0) ? r : T(0);
    r = r < az ? r : az;
  }
  v[0] = px; v[1] = py; v[2] = z0 * r / az;
}

template <typename T>
__device__ __forceinline__ void project_cone3(T* v, unsigned char kind, T a, int max_iter, T tol) {
  if (kind == C3_EXP) { project_exp(v, max_iter, tol); return; }
  if (kind == C3_POW) { project_pow(v, a, max_iter, tol); return; }
  const T v0[3] = {v[0], v[1], v[2]};
  v[0] = -v[0]; v[1] = -v[1]; v[2] = -v[2];
  if (kind == C3_DUAL_EXP) project_exp(v, max_iter, tol); else project_pow(v, a, max_iter, tol);
  v[0] += v0[0]; v[1] += v0[1]; v[2] += v0[2];
}

// s[rows of cone k] = Pi(w_s[rows of cone k]);  one thread per cone
template <typename T>
__global__ void __launch_bounds__(128) cone3_project_kernel(Cone3Table<T> t, const T* __restrict__ ws, T* __restrict__ s) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= t.n) return;
  const int o = t.off[k];
  T v[3] = {ws[o], ws[o + 1], ws[o + 2]};
  project_cone3(v, t.kind[k], t.alpha[k], t.max_iter[k], t.tol[k]);
  s[o] = v[0]; s[o + 1] = v[1]; s[o + 2] = v[2];
}

// Certificate shared by both infeasibility tests, as for the second-order cone: the primal test asks
// in_dual(-v) (support_function!, convexset.jl:933-936) and the dual test in_pol_recc(v) = in_dual(-v)
// (convexset.jl:616-618, 740-742, 781); for the dual cones in_dual is in_cone of the primal cone (:780).
// flag[0] = 1 if any cone fails.  Single block.
template <typename T>
__global__ void __launch_bounds__(kBlock) cone3_cert_kernel(Cone3Table<T> t, const T* __restrict__ v, T tol, T* __restrict__ flag) {
  int bad = 0;
  for (int k = threadIdx.x; k < t.n; k += blockDim.x) {
    const int o = t.off[k];
    const T x = -v[o], y = -v[o + 1], z = -v[o + 2];
    bool ok;
    switch (t.kind[k]) {
      case C3_EXP: ok = exp_in_dual(x, y, z, tol); break;
      case C3_DUAL_EXP: ok = exp_in_cone(x, y, z, tol); break;
      case C3_POW: ok = pow_in_dual(x, y, z, t.alpha[k], tol); break;
      default: ok = pow_in_cone(x, y, z, t.alpha[k], tol); break;
    }
    if (!ok) bad = 1;
  }
  bad = __syncthreads_or(bad);
  if (threadIdx.x == 0) flag[0] = bad ? T(1) : T(0);
}

}  // namespace cosmo
