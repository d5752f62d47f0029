/* oracle/spmv_omp.c -- TEST / BENCH INFRASTRUCTURE, NOT THE PRODUCT.
 *
 * Row-parallel CSR mat-vec for the CPU reference arm of bench.py (`--impl reference`) and its `cpu_baseline` leg:
 * the oracle port (oracle/cosmo_oracle.py) spends > 99 % of an ADMM iteration of config C2 in the sparse products of
 * the reduced-KKT operator (kktsolver_indirect.jl:56-66: mul! with A, with A' and with P).  The reference's
 * SparseArrays.mul! is single-threaded; the bench contract asks for "all the host threads it can use", so the
 * reference arm swaps SciPy's single-threaded kernels for this one (same arithmetic per row, rows in parallel).
 * Only bench.py and tests/ may load it.
 *
 *   gcc -O3 -march=native -fopenmp -shared -fPIC -o oracle/_build/liboracle_spmv.so oracle/spmv_omp.c
 */
#include <stdint.h>

void oracle_csr_matvec(int64_t nrows, const int32_t* indptr, const int32_t* indices, const double* data,
                       const double* x, double* y) {
#pragma omp parallel for schedule(dynamic, 64)
  for (int64_t i = 0; i < nrows; ++i) {
    double acc = 0.0;
    for (int32_t k = indptr[i]; k < indptr[i + 1]; ++k) acc += data[k] * x[indices[k]];
    y[i] = acc;
  }
}

#ifdef _OPENMP
#include <omp.h>
#endif

int oracle_spmv_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* number of threads of the following mat-vecs (bench.py calibrates it: a container's CPU quota can be far below the
   number of cores it sees) */
void oracle_spmv_set_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}
