// Host-side check of the digit extraction used by slice_rows_kernel (csrc/tc_gemm.cuh): slice_fixed / slice_pack4 are
// __host__ __device__, so the very functions the kernel calls are exercised here without a GPU (compiled by
// tests/test_tc_scheme_emulated_cpu.py with nvcc; nothing is launched).
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include "tc_gemm.cuh"

using namespace cosmo::tc;

static unsigned long long rng_state = 0x9E3779B97F4A7C15ull;
static double next_unit() {   // xorshift, uniform in (-1, 1)
  rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17;
  return ((double)(rng_state >> 11) / 9007199254740992.0) * 2.0 - 1.0;
}

template <int K>
static long long check(long long* n_checked) {
  long long bad = 0;
  const double edge[] = {0.0, 0.5, -0.5, 1.0, -1.0, 63.5, -63.5, 64.0 - ldexp(1.0, -47), -(64.0 - ldexp(1.0, -47)),
                         ldexp(1.0, -60), -ldexp(1.0, -60), 1.0 / 3.0, -1.0 / 3.0, 32.0, -32.0, 0.49999999999999994};
  const int nedge = (int)(sizeof(edge) / sizeof(edge[0]));
  for (int it = 0; it < 200000; ++it) {
    unsigned long long u[4];
    double t[4];
    for (int j = 0; j < 4; ++j) {
      const int id = it * 4 + j;
      t[j] = id < nedge ? edge[id] : 64.0 * next_unit() * ldexp(1.0, -(int)(rng_state % 40));
      u[j] = slice_fixed<K>(t[j]);
    }
    int d[4][8];
    for (int p = 0; p < K; ++p) {
      const uint32_t w = slice_pack4<K>(u, p);
      for (int j = 0; j < 4; ++j) d[j][p] = (int)(int8_t)((w >> (8 * j)) & 0xffu);
    }
    for (int j = 0; j < 4; ++j) {
      long long v = 0;                                   // sum_p d_p 128^(K-1-p) must be rn(t 2^(7K-7)) exactly
      for (int p = 0; p < K; ++p) {
        v = v * 128 + d[j][p];
        const int lim_hi = p == 0 ? 64 : 63;
        if (d[j][p] < -64 || d[j][p] > lim_hi) ++bad;
      }
      if (v != llrint(t[j] * (double)(1ull << (7 * K - 7)))) ++bad;
      // against the sequential round-to-nearest expansion of the first version (rint per digit): the same number
      // whenever no remainder sits within one last-digit unit of a tie, digit by digit
      double r = t[j];
      bool near_tie = false, same = true;
      for (int p = 0; p < K; ++p) {
        const double frac = fabs(fabs(r - floor(r)) - 0.5);
        if (frac <= ldexp(1.0, -7 * (K - 1 - p))) near_tie = true;
        const double dd = rint(r);
        if ((int)dd != d[j][p]) same = false;
        r = (r - dd) * 128.0;
      }
      if (!same && !near_tie) ++bad;
      ++*n_checked;
    }
  }
  return bad;
}

int main() {
  long long n = 0;
  const long long bad = check<8>(&n) + check<7>(&n) + check<6>(&n) + check<4>(&n);
  printf("checked %lld bad %lld\n", n, bad);
  return bad == 0 ? 0 : 1;
}
