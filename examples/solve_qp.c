/* The reference's examples/qp.jl (x* = [0.3, 0.7], objective 1.88) solved from plain C through the drop-in boundary
 * include/cosmo_b200.h -- what a host binding (the Julia shim of INTEGRATION.md, or any other FFI) does:
 *
 *     min 1/2 x'Px + q'x   s.t.  l <= A x <= u        P = [4 1; 1 2], q = [1 1], A = [1 1; 1 0; 0 1]
 *
 * in COSMO's model form  A_m x + s = b, s in Nonnegatives(6)  with  A_m = [A; -A],  b = [u; -l]
 * (interface.jl:478-485: the user's constraint is  [-A; A] x + [u; -l] >= 0).
 *
 *     gcc -std=c99 -Iinclude examples/solve_qp.c cosmo.jl_b200/libcosmo_b200.so -Wl,-rpath,$PWD/cosmo.jl_b200 -o solve_qp
 *
 * Exit code 0: solved and matches; 3: no usable GPU (the engine has no CPU fallback); 1: anything else. */
#include <math.h>
#include <stdio.h>
#include <string.h>

#include "cosmo_b200.h"

int main(void) {
  /* CSC, 0-based (index_base = 0) */
  int64_t P_colptr[] = {0, 2, 4}, P_rowval[] = {0, 1, 0, 1};
  double P_nz[] = {4, 1, 1, 2};
  /* A_m = [1 1; 1 0; 0 1; -1 -1; -1 0; 0 -1] */
  int64_t A_colptr[] = {0, 4, 8}, A_rowval[] = {0, 1, 3, 4, 0, 2, 3, 5};
  double A_nz[] = {1, 1, -1, -1, 1, 1, -1, -1};
  double q[] = {1, 1};
  double b[] = {1, 0.7, 0.7, -1, 0, 0};

  cosmo_b200_set set;
  memset(&set, 0, sizeof set);
  set.type = COSMO_B200_NONNEG;
  set.dim = 6;

  cosmo_b200_problem prob;
  memset(&prob, 0, sizeof prob);
  prob.dtype = COSMO_B200_F64;
  prob.index_base = 0;
  prob.device = 0;
  prob.m = 6;
  prob.n = 2;
  prob.P.nrows = 2; prob.P.ncols = 2; prob.P.colptr = P_colptr; prob.P.rowval = P_rowval; prob.P.nzval = P_nz;
  prob.A.nrows = 6; prob.A.ncols = 2; prob.A.colptr = A_colptr; prob.A.rowval = A_rowval; prob.A.nzval = A_nz;
  prob.q = q;
  prob.b = b;
  prob.n_sets = 1;
  prob.sets = &set;
  prob.c = 1.0; /* unscaled data: D = E = NULL */

  cosmo_b200_settings st;
  if (cosmo_b200_default_settings(&st) != COSMO_B200_OK) return 1;
  st.scaling = 0; /* the data above is not equilibrated */

  cosmo_b200_handle* h = NULL;
  int rc = cosmo_b200_create(&h, &prob, &st);
  if (rc == COSMO_B200_ERR_CUDA) {
    printf("no usable GPU: %s\n", cosmo_b200_last_error(NULL));
    return 3;
  }
  if (rc != COSMO_B200_OK) {
    printf("create failed (%d): %s\n", rc, cosmo_b200_last_error(NULL));
    return 1;
  }
  double x[2], s[6], mu[6], rho_updates[16];
  cosmo_b200_result res;
  memset(&res, 0, sizeof res);
  res.x = x; res.s = s; res.mu = mu;
  res.rho_updates = rho_updates; res.rho_updates_cap = 16;
  rc = cosmo_b200_solve(h, &res);
  if (rc != COSMO_B200_OK) {
    printf("solve failed (%d): %s\n", rc, cosmo_b200_last_error(h));
    cosmo_b200_destroy(h);
    return 1;
  }
  printf("status %d iter %lld obj %.6f x %.6f %.6f launches %lld\n", res.status, (long long)res.iter, res.obj_val, x[0], x[1],
         (long long)res.kernel_launches);
  cosmo_b200_destroy(h);
  int ok = res.status == COSMO_B200_SOLVED && fabs(res.obj_val - 1.88) < 1e-3 && fabs(x[0] - 0.3) < 1e-3 && fabs(x[1] - 0.7) < 1e-3;
  return ok ? 0 : 1;
}
