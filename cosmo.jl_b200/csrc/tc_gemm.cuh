// tc_gemm.cuh -- fp64-accurate products of symmetric matrices on the 5th-generation tensor cores (sm_100a).
//
// The PSD projection of a large cone (convexset.jl:219-263 in the reference: dsyevr + clamp + syrk) is computed here
// as  Pi_+(X) = (X + sign(X) X) / 2  with sign(X) from a Newton-Schulz iteration (psd_tc.cuh) -- nothing but
// N x N x N products of symmetric, commuting matrices.  tcgen05.mma has no fp64 kind, so every product is evaluated
// by error-free slicing (Ozaki scheme I):
//
//     M[r, :] = 2^(e_r - 6) * sum_p D_p[r, :] * 128^-(p-1),   D_p int8 digits in [-64, 64]   (slice_rows_kernel)
//     (A B)[m, n] = sA[m] sB[n] * sum_s 128^-(s-2) G_s[m, n],  G_s = sum_{p+q=s} A_p B_q^T     (int32, EXACT)
//
// G_s is accumulated by tcgen05.mma.kind::i8 in TMEM (|G_s| <= 8 * K * 64^2 < 2^31 for K <= 65536), the sum over s
// runs in fp64 registers of the epilogue warps (acc += 128^-(s-2) G_s, least significant groups first), so the only
// rounding errors are the truncation of the slices (2^-7k relative to the row maximum) and one fp64 rounding per group.
// Measured against numpy dgemm (N = 2000, profiles/tc_gemm_bringup_r2_*.log): k = 8 -> 1e-15, k = 7 -> 1e-13 relative
// to |A| |B|.
//
// Kernel anatomy (one persistent CTA per SM, 12 warps):
//   warp 0      TMA producer: cp.async.bulk.tensor.2d, 128-row x 128-byte boxes (= the tiles of 4 slices for one K step
//               of 32, see the layout note below), 128-byte hardware swizzle, into a ring of six 32 KB units
//               {A box, B box}, mbarrier complete_tx
//   warp 1      MMA issuer: one elected lane walks the pass list of the plan and issues tcgen05.mma.kind::i8 (M = N = 128,
//               K = 32) from shared-memory descriptors into four 128-column TMEM accumulators -- up to 26 slice-pair
//               products per stage; tcgen05.commit releases the stage / publishes the accumulators
//   warps 4..11 epilogue: tcgen05.ld 32x32b, exact int32 -> fp64 accumulation over the groups, scaling, fused
//               out = c0 * (A B) + c1 * D + c2 * I, mirrored store (the product of commuting symmetric matrices is
//               symmetric: only tiles of the upper triangle are computed) and two fused Frobenius reductions.
// Operand tiles are re-used by all slice pairs of four groups while they sit in shared memory: L2 -> SM traffic is
// what bounds an int8 product of this shape (the first version, one slice pair per stage, ran at 11 TB/s of L2 reads
// and 28 % of the int8 peak -- profiles/tc_gemm_bringup_r2_v1.log).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>
#include <algorithm>

namespace cosmo {
namespace tc {

constexpr int kTile = 128;         // output tile side = rows of one operand box = UMMA M = UMMA N
constexpr int kSlices = 8;         // slices stored per matrix (unused ones are zero)
constexpr int kKStep = 32;         // K per stage = K of one tcgen05.mma.kind::i8
constexpr int kBoxBytes = kTile * 128;   // one TMA box: 128 rows x (4 slices x 32 K-bytes), 128-byte swizzle
constexpr int kUnitBytes = 2 * kBoxBytes;    // ring unit: {A box, B box} of slices 0-3 (or of slices 4-7)
constexpr int kUnits = 6;                    // 192 KB ring: 3 K steps in flight when a pass needs both halves, 6 otherwise
constexpr int kThreads = 384;      // warpgroup 0: warp 0 TMA, warp 1 MMA (2, 3 idle); warpgroups 1, 2: epilogue
constexpr int kEpiWarps = 8;

// Layout of a sliced operand: int8 [row][K / 32][slice 0..7][32], i.e. a row of Np * 8 bytes in which the 8 slices of
// the same 32 K-values are adjacent.  One 128-byte TMA row therefore carries 4 slices of one K step, a 128 x 128-byte
// box is the operand tile of 4 slices at once, and the tile of slice j inside the box is addressed like the j-th
// K sub-step of an ordinary 128-byte-swizzled K-major tile (descriptor start address + 32 j bytes).
//
// The product list is static.  With K slices and G groups the groups s = p + q run from G + 1 down to 2; a PASS owns
// the four TMEM accumulators and covers the groups s_hi, s_hi - 1, s_hi - 2, s_hi - 3: pass 0 starts at G + 1, pass 1
// at G - 3, ...  Every slice pair of a pass is issued per K step from the same operand boxes, so a tile fetched once
// serves up to 28 products.  Everything below is resolved at compile time (template parameters K, G): the issuing
// thread executes two uniform adds and one UTCIMMA per product.
// G >= K is the number of groups kept: G = K is the classical truncation (error ~ K 2^-7K relative to the row maxima),
// every further group gains 7 bits; G = K + 2 leaves only the truncation of the operands themselves (2^-7K per
// element): with K = 8 the product is then more accurate than dgemm.
__host__ __device__ constexpr int pass_s_hi(int g, int pass) { return g + 1 - 4 * pass; }
__host__ __device__ constexpr int num_passes(int g) { return (g + 3) / 4; }
__host__ __device__ constexpr int pass_groups(int g, int pass) { return pass_s_hi(g, pass) - 1 < 4 ? pass_s_hi(g, pass) - 1 : 4; }
// bit 0: A slices 0-3, bit 1: A slices 4-7, bit 2: B slices 0-3, bit 3: B slices 4-7 (the same set for A and B by symmetry)
__host__ __device__ constexpr int pass_box_mask(int k, int g, int pass) {
  int m = 0;
  for (int gi = 0; gi < pass_groups(g, pass); ++gi)
    for (int p = 1; p <= k; ++p) {
      const int q = pass_s_hi(g, pass) - gi - p;
      if (q >= 1 && q <= k) m |= (1 << ((p - 1) / 4)) | (4 << ((q - 1) / 4));
    }
  return m;
}
__host__ __device__ constexpr int total_products(int k, int g) {
  int n = 0;
  for (int s = 2; s <= g + 1; ++s)
    for (int p = 1; p <= k; ++p)
      if (s - p >= 1 && s - p <= k) ++n;
  return n;
}

// ---------------------------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "LAB_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra LAB_DONE;\n\t"
      "bra LAB_WAIT;\n\t"
      "LAB_DONE:\n\t}\n" ::"r"(smem_u32(bar)), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_dst),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P1;\n\telect.sync _|P1, 0xffffffff;\n\tselp.b32 %0, 1, 0, P1;\n\t}\n" : "=r"(pred));
  return pred != 0;
}
// D[tmem] (+)= A[smem] * B[smem]^T, int8 x int8 -> int32, M = N = 128, K = 32
__device__ __forceinline__ void tc_mma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate));
}
// shared-memory matrix descriptor of a K-major operand tile inside a 128-row x 128-byte box written by TMA with the
// 128-byte swizzle (atoms of 8 rows x 128 bytes: stride between atoms along M / N = 1024 bytes); `saddr16` is the
// shared-memory byte address >> 4 of the first row + 32 * (slice within the box)
constexpr uint64_t kDescHi = ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);   // SBO, version 1, SWIZZLE_128B
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr16) { return kDescHi | (uint64_t)(saddr16 & 0x3FFFu); }
// instruction descriptor: S32 accumulate, signed 8-bit A and B, both K-major, N = 128, M = 128
constexpr uint32_t kIdescI8 = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kTile >> 3) << 17) | ((uint32_t)(kTile >> 4) << 24);

#define COSMO_TC_LD32(taddr, v)                                                                                          \
  asm volatile(                                                                                                          \
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                                                          \
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "                                          \
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"                          \
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),      \
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),           \
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),          \
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])                        \
      : "r"(taddr))

// (double)v for an int32 v without the conversion unit: 2^52 + 2^31 + v is exactly representable, its bit pattern is
// {0x43300000, v ^ 0x80000000}
__device__ __forceinline__ double biased_double(uint32_t v) { return __hiloint2double(0x43300000, (int)(v ^ 0x80000000u)); }
constexpr double kBias = 4503599627370496.0 + 2147483648.0;

// ---------------------------------------------------------------------------------------------------------------
// Slicing: one CTA per row.  scale[r] = 2^(e_r - 6) with 2^e_r > max |M[r, :]|, so t = x / scale is in (-64, 64).
// The K digits are the balanced base-128 digits of the fixed-point number v = rn(t 2^(7K-7)) (unit of the last kept
// digit = 1): with the bias B = 64 sum_{j<K} 128^j added, u = v + B >= 0 and the 7-bit FIELDS of u are digit + 64 --
// no carries, no per-digit rounding: one fp64 multiply and one conversion per element, integer field extraction per
// digit (the first version took rint / subtract / scale / convert per digit in fp64).  Digits are in [-64, 63], the
// leading one in [-64, 64]; all of it exact.  Output layout: see the note above.
// ---------------------------------------------------------------------------------------------------------------
template <int K>
struct SliceFix {
  static constexpr int kShift = 7 * K - 7;
  __host__ __device__ static constexpr unsigned long long bias() {
    unsigned long long b = 0;
    for (int j = 0; j < K; ++j) b += 64ull << (7 * j);
    return b;
  }
};

template <int K>
__host__ __device__ __forceinline__ unsigned long long slice_fixed(double t) {
  const double s = t * (double)(1ull << SliceFix<K>::kShift);
#ifdef __CUDA_ARCH__
  const long long v = __double2ll_rn(s);
#else
  const long long v = llrint(s);
#endif
  return (unsigned long long)(v + (long long)SliceFix<K>::bias());
}

// digit p (0 = most significant) of u[0..3] as the four int8 bytes of one word
template <int K>
__host__ __device__ __forceinline__ uint32_t slice_pack4(const unsigned long long* u, int p) {
  const int off = 7 * (K - 1 - p);
  uint32_t w = 0;
  if (p == 0) {   // the leading field runs over [0, 128]
#pragma unroll
    for (int j = 0; j < 4; ++j) w |= ((uint32_t)((int)(u[j] >> off) - 64) & 0xffu) << (8 * j);
    return w;
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) w |= ((uint32_t)(u[j] >> off) & 127u) << (8 * j);
  // field f -> int8(f - 64), four bytes at a time: f >= 64: f ^ 0x40;  f < 64: (f ^ 0x40) | 0x80 = f + 192
  const uint32_t x = w ^ 0x40404040u;
  return x | ((x & 0x40404040u) << 1);
}

// The row is read once, coalesced, for the maximum and parked in shared memory (element c at c + c / 8: the padding
// makes the 8-consecutive-columns-per-thread reads of the second pass conflict-free; read straight from global memory
// they cost 16 L1 wavefronts per load instruction).  Rows too long for shared memory (`staged` = 0) are re-read.
constexpr int kSliceStageMaxBytes = 200 * 1024;
__host__ __device__ constexpr size_t slice_stage_bytes(int N) { return ((size_t)N + (size_t)N / 8 + 8) * sizeof(double); }

template <typename T, int K>
__global__ void __launch_bounds__(256) slice_rows_kernel(const T* __restrict__ M, int N, int Np, int8_t* __restrict__ slices,
                                                        double* __restrict__ scale, int staged) {
  extern __shared__ __align__(16) double slice_srow[];
  const int r = blockIdx.x;
  const T* row = M + (size_t)r * N;   // symmetric: row r == column r of the column-major matrix
  __shared__ double red[8];
  double mx = 0.0;
  for (int c = threadIdx.x; c < N; c += blockDim.x) {
    const double v = (double)row[c];
    if (staged) slice_srow[c + (c >> 3)] = v;
    mx = fmax(mx, fabs(v));
  }
  for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
  __syncthreads();
  mx = red[0];
  for (int w = 1; w < 8; ++w) mx = fmax(mx, red[w]);
  int e = 0;
  if (mx > 0.0 && mx < 1.0e300) frexp(mx, &e);   // mx = f 2^e, f in [0.5, 1)  =>  |x| < 2^e
  if (threadIdx.x == 0) scale[r] = ldexp(1.0, e - 6);
  const double inv = ldexp(1.0, 6 - e);
  int8_t* out_row = slices + (size_t)r * Np * kSlices;
  for (int c0 = threadIdx.x * 8; c0 < N; c0 += blockDim.x * 8) {
    unsigned long long u[8];
    const double* sp = slice_srow + c0 + (c0 >> 3);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      double x = 0.0;
      if (c0 + j < N) x = staged ? sp[j] : (double)row[c0 + j];
      u[j] = slice_fixed<K>(x * inv);
    }
    int8_t* dst = out_row + (size_t)(c0 >> 5) * (kSlices * 32) + (c0 & 31);
#pragma unroll
    for (int p = 0; p < K; ++p)
      *reinterpret_cast<uint2*>(dst + p * 32) = make_uint2(slice_pack4<K>(u, p), slice_pack4<K>(u + 4, p));
  }
}

// ---------------------------------------------------------------------------------------------------------------
// The product kernel
// ---------------------------------------------------------------------------------------------------------------
template <typename T>
struct GemmArgs {
  int N, Np, ntiles, store;          // store: 1 = write `out`
  const int2* tiles;                 // (bi, bj), bi <= bj: tiles of the upper triangle
  const double* scaleA;              // Np row scales of the A operand
  const double* scaleB;              // Np row (= column) scales of the B operand
  T* out;                            // N x N, ld = N (symmetric, mirrored store)
  const T* D;                        // optional: out = c0 * (A B) + c1 * D + c2 * I
  const T* E;                        // optional reference of the second reduction
  int e_identity;                    // 1: second reduction against the identity
  const double* coef;                // device scalars c0, c1, c2
  double* partial;                   // 2 per tile: sum w out^2, sum w (E - out)^2   (w: 1 on the diagonal, 2 above it)
};

// all slice-pair products of one pass for one K step (compile-time list)
template <int K, int G, int PASS>
__device__ __forceinline__ void issue_pass_products(uint32_t sbase16, uint32_t tmem_base, int ks) {
  constexpr int S_HI = pass_s_hi(G, PASS);
  constexpr int NG = pass_groups(G, PASS);
#pragma unroll
  for (int p = 1; p <= K; ++p) {       // ordered by A slice: consecutive products share the A tile
#pragma unroll
    for (int g = 0; g < NG; ++g) {
      const int s = S_HI - g, q = s - p;
      if (q >= 1 && q <= K) {
        const uint32_t a_off = (uint32_t)((((p - 1) / 4) * kUnitBytes + ((p - 1) % 4) * 32) >> 4);
        const uint32_t b_off = (uint32_t)((((q - 1) / 4) * kUnitBytes + kBoxBytes + ((q - 1) % 4) * 32) >> 4);
        const bool first = (p == ((s - K) > 1 ? (s - K) : 1));      // first write of this accumulator in the pass
        tc_mma_i8(tmem_base + (uint32_t)(g * kTile), smem_desc(sbase16 + a_off), smem_desc(sbase16 + b_off), kIdescI8,
                  first ? (ks != 0 ? 1u : 0u) : 1u);
      }
    }
  }
}

struct PipeState { int unit; uint32_t phase; };   // position in the ring of kUnits units

// a pass needs the slices 4-7 of either operand <=> it occupies two ring units per K step (kUnits is even and the
// number of K steps is even, so a two-unit step never wraps inside)
__host__ __device__ constexpr int pass_units(int k, int g, int pass) { return (pass_box_mask(k, g, pass) & 0xA) ? 2 : 1; }

template <int K, int G, int PASS>
__device__ __forceinline__ void producer_pass(const CUtensorMap* tmapA, const CUtensorMap* tmapB, uint8_t* smem, uint64_t* full_bar,
                                              uint64_t* empty_bar, int nk, int m0, int n0, PipeState& st) {
  constexpr int NU = pass_units(K, G, PASS);
  for (int ks = 0; ks < nk; ++ks) {
#pragma unroll
    for (int h = 0; h < NU; ++h) {
      mbar_wait(&empty_bar[st.unit], st.phase ^ 1);
      if (elect_one()) {
        mbar_expect_tx(&full_bar[st.unit], (uint32_t)kUnitBytes);
        const uint32_t dst = smem_u32(smem) + (uint32_t)st.unit * kUnitBytes;
        const int c0 = ks * (kSlices * 32) + h * 128;
        tma_load_2d(dst, tmapA, &full_bar[st.unit], c0, m0);
        tma_load_2d(dst + kBoxBytes, tmapB, &full_bar[st.unit], c0, n0);
      }
      __syncwarp();
      if (++st.unit == kUnits) { st.unit = 0; st.phase ^= 1; }
    }
  }
}

template <int K, int G, int PASS>
__device__ __forceinline__ void mma_pass(uint8_t* smem, uint64_t* full_bar, uint64_t* empty_bar, uint64_t* tfull_bar, uint64_t* tempty_bar,
                                         uint32_t tmem_base, int nk, PipeState& st, uint32_t& tphase) {
  constexpr int NU = pass_units(K, G, PASS);
  mbar_wait(tempty_bar, tphase ^ 1);      // the epilogue has drained the accumulators of the previous pass
  tc_fence_after();
  for (int ks = 0; ks < nk; ++ks) {
    mbar_wait(&full_bar[st.unit], st.phase);
    if (NU == 2) mbar_wait(&full_bar[st.unit + 1], st.phase);
    tc_fence_after();
    if (elect_one()) {
      const uint32_t sbase16 = (smem_u32(smem) + (uint32_t)st.unit * kUnitBytes) >> 4;
      issue_pass_products<K, G, PASS>(sbase16, tmem_base, ks);
      tc_commit(&empty_bar[st.unit]);
      if (NU == 2) tc_commit(&empty_bar[st.unit + 1]);
      if (ks == nk - 1) tc_commit(tfull_bar);
    }
    __syncwarp();
    st.unit += NU;
    if (st.unit == kUnits) { st.unit = 0; st.phase ^= 1; }
  }
  tphase ^= 1;
}

// fold the accumulators of one pass into the fp64 registers: acc += 128^-(s-2) G_s, exactly converted
template <int G, int PASS>
__device__ __forceinline__ void epilogue_pass(double (&acc)[64], uint32_t tlane, uint64_t* tfull_bar, uint64_t* tempty_bar, uint32_t& tphase,
                                              int lane) {
  constexpr int S_HI = pass_s_hi(G, PASS);
  constexpr int NG = pass_groups(G, PASS);
  mbar_wait(tfull_bar, tphase);
  tphase ^= 1;
  tc_fence_after();
#pragma unroll
  for (int g = 0; g < NG; ++g) {
    const double sc = 1.0 / (double)(1ull << (7 * (S_HI - g - 2)));     // 128^-(s-2), a power of two
    const double nb = -kBias * sc;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      uint32_t v[32];
      COSMO_TC_LD32(tlane + (uint32_t)(g * kTile + c * 32), v);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int j = 0; j < 32; ++j) acc[c * 32 + j] += fma(biased_double(v[j]), sc, nb);   // fma is exact: v * sc
    }
  }
  tc_fence_before();
  __syncwarp();
  if (lane == 0) mbar_arrive(tempty_bar);
}

template <typename T, int K, int G>
__global__ void __launch_bounds__(kThreads, 1) ozaki_gemm_kernel(const __grid_constant__ CUtensorMap tmapA,
                                                                 const __grid_constant__ CUtensorMap tmapB, const GemmArgs<T> args) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ __align__(8) uint64_t full_bar[kUnits], empty_bar[kUnits], tfull_bar, tempty_bar;
  __shared__ uint32_t tmem_base_s;
  __shared__ double red[kEpiWarps][2];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    for (int i = 0; i < kUnits; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    mbar_init(&tfull_bar, 1);
    mbar_init(&tempty_bar, kEpiWarps);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {   // the allocating warp also frees
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  const int nk = args.Np / kKStep;

  // register re-distribution (setmaxnreg is per warpgroup): the TMA / MMA warps need few registers, the epilogue
  // threads hold a 64-element fp64 accumulator row each plus the 32 freshly loaded TMEM words
  if (warp == 0) {
    // ===== TMA producer (warp-uniform loop, one elected lane issues) =====
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
    PipeState st{0, 0};
    for (int t = blockIdx.x; t < args.ntiles; t += gridDim.x) {
      const int2 tl = args.tiles[t];
      const int m0 = tl.x * kTile, n0 = tl.y * kTile;
      producer_pass<K, G, 0>(&tmapA, &tmapB, smem, full_bar, empty_bar, nk, m0, n0, st);
      if (num_passes(G) > 1) producer_pass<K, G, (num_passes(G) > 1 ? 1 : 0)>(&tmapA, &tmapB, smem, full_bar, empty_bar, nk, m0, n0, st);
      if (num_passes(G) > 2) producer_pass<K, G, (num_passes(G) > 2 ? 2 : 0)>(&tmapA, &tmapB, smem, full_bar, empty_bar, nk, m0, n0, st);
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
    PipeState st{0, 0};
    uint32_t tphase = 0;
    for (int t = blockIdx.x; t < args.ntiles; t += gridDim.x) {
      mma_pass<K, G, 0>(smem, full_bar, empty_bar, &tfull_bar, &tempty_bar, tmem_base, nk, st, tphase);
      if (num_passes(G) > 1) mma_pass<K, G, (num_passes(G) > 1 ? 1 : 0)>(smem, full_bar, empty_bar, &tfull_bar, &tempty_bar, tmem_base, nk, st, tphase);
      if (num_passes(G) > 2) mma_pass<K, G, (num_passes(G) > 2 ? 2 : 0)>(smem, full_bar, empty_bar, &tfull_bar, &tempty_bar, tmem_base, nk, st, tphase);
    }
  } else if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
  } else {
    // ===== epilogue warps =====
    asm volatile("setmaxnreg.inc.sync.aligned.u32 224;");
    const int ew = warp - 4;
    const int quad = warp & 3;            // TMEM lanes [32 quad, 32 quad + 32) are the ones this warp may read
    const int half = ew >> 2;             // columns [64 half, 64 half + 64) of the tile
    const uint32_t tlane = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(half * 64);
    uint32_t tphase = 0;
    const double c0 = args.coef[0], c1 = args.coef[1], c2 = args.coef[2];
    for (int t = blockIdx.x; t < args.ntiles; t += gridDim.x) {
      const int2 tl = args.tiles[t];
      const int gm = tl.x * kTile + quad * 32 + lane;
      const int gn0 = tl.y * kTile + half * 64;
      double acc[64];
#pragma unroll
      for (int j = 0; j < 64; ++j) acc[j] = 0.0;
      epilogue_pass<G, 0>(acc, tlane, &tfull_bar, &tempty_bar, tphase, lane);
      if (num_passes(G) > 1) epilogue_pass<G, (num_passes(G) > 1 ? 1 : 0)>(acc, tlane, &tfull_bar, &tempty_bar, tphase, lane);
      if (num_passes(G) > 2) epilogue_pass<G, (num_passes(G) > 2 ? 2 : 0)>(acc, tlane, &tfull_bar, &tempty_bar, tphase, lane);
      // ---- final epilogue of the tile ----
      double r0 = 0.0, r1 = 0.0;
      const int N = args.N;
      if (gm < N) {
        const double sa = args.scaleA[gm];
#pragma unroll
        for (int j = 0; j < 64; ++j) {
          const int gn = gn0 + j;
          if (gn >= N || gm > gn) continue;               // upper triangle only; the mirror is written below
          double o = c0 * (acc[j] * sa * args.scaleB[gn]);
          const size_t tidx = (size_t)gn * N + gm;        // element (gm, gn) of a column-major matrix (coalesced over lanes)
          if (args.D) o += c1 * (double)args.D[tidx];
          if (gm == gn) o += c2;
          const double w = (gm == gn) ? 1.0 : 2.0;
          r0 += w * o * o;
          if (args.e_identity || args.E) {
            const double ref = args.E ? (double)args.E[tidx] : ((gm == gn) ? 1.0 : 0.0);
            r1 += w * (ref - o) * (ref - o);
          }
          if (args.store) {
            args.out[tidx] = (T)o;
            if (gm != gn) args.out[(size_t)gm * N + gn] = (T)o;
          }
        }
      }
      if (args.partial) {
        for (int o = 16; o > 0; o >>= 1) {
          r0 += __shfl_xor_sync(0xffffffffu, r0, o);
          r1 += __shfl_xor_sync(0xffffffffu, r1, o);
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");     // previous tile's reader is done with red[]
        if (lane == 0) { red[ew][0] = r0; red[ew][1] = r1; }
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (ew == 0 && lane == 0) {
          double a = 0.0, b = 0.0;
          for (int w = 0; w < kEpiWarps; ++w) { a += red[w][0]; b += red[w][1]; }
          args.partial[2 * t] = a;
          args.partial[2 * t + 1] = b;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Host side
// ---------------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// One sliced operand: int8 [Np][Np / 32][8][32] + Np scales, and its tensor map for the current Np.
struct Sliced {
  int8_t* d = nullptr;
  double* scale = nullptr;
  CUtensorMap map;
  int capNp = 0, mapNp = 0;
  ~Sliced() { cudaFree(d); cudaFree(scale); }
  bool ensure(int Np) {
    if (Np <= capNp) return true;
    cudaFree(d); cudaFree(scale);
    d = nullptr; scale = nullptr; capNp = 0; mapNp = 0;
    if (cudaMalloc(&d, (size_t)kSlices * Np * Np) != cudaSuccess || cudaMalloc(&scale, (size_t)Np * sizeof(double)) != cudaSuccess) return false;
    capNp = Np;
    return true;
  }
  // zero padding rows / columns and unused slices: called when the shape of the cone changes
  bool clear(int Np, cudaStream_t st) {
    mapNp = 0;
    return cudaMemsetAsync(d, 0, (size_t)kSlices * Np * Np, st) == cudaSuccess &&
           cudaMemsetAsync(scale, 0, (size_t)Np * sizeof(double), st) == cudaSuccess;
  }
  bool make_map(int Np) {
    if (mapNp == Np) return true;
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc) return false;
    const cuuint64_t dims[2] = {(cuuint64_t)Np * kSlices, (cuuint64_t)Np};
    const cuuint64_t strides[1] = {(cuuint64_t)Np * kSlices};
    const cuuint32_t box[2] = {128, (cuuint32_t)kTile};
    const cuuint32_t estr[2] = {1, 1};
    if (enc(&map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, d, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return false;
    mapNp = Np;
    return true;
  }
};

template <typename T>
struct OzakiGemm {
  int2* tiles_d = nullptr;
  int tilesNp = 0, ntiles = 0;
  int k = 8, g = 10, num_sms = 148;
  int N = 0, Np = 0;
  bool ready = false;
  std::string err;
  ~OzakiGemm() { cudaFree(tiles_d); }

  static constexpr int smem_bytes() { return kUnits * kUnitBytes + 1024; }
  // supported (slices, groups): (8, 10) exact-fp64 default, (8, 8) and (7, 7) classical truncations, (6, 8) and (4, 6)
  // for the fp32 model type
  template <int K, int G>
  static bool set_attr() {
    return cudaFuncSetAttribute(ozaki_gemm_kernel<T, K, G>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes()) == cudaSuccess;
  }
  template <int K>
  static bool set_slice_attr() {
    return cudaFuncSetAttribute(slice_rows_kernel<T, K>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSliceStageMaxBytes) == cudaSuccess;
  }
  static bool supported(int k_, int g_) {
    return (k_ == 8 && (g_ == 10 || g_ == 8)) || (k_ == 7 && g_ == 7) || (k_ == 6 && g_ == 8) || (k_ == 4 && g_ == 6);
  }
  bool configure(int k_, int g_, cudaStream_t st) {
    (void)st;
    if (!supported(k_, g_)) { err = "tc::OzakiGemm: unsupported (slices, groups)"; return false; }
    k = k_; g = g_;
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
    // per device, every time: function attributes are per device and cheap to set
    if (!(set_attr<8, 10>() && set_attr<8, 8>() && set_attr<7, 7>() && set_attr<6, 8>() && set_attr<4, 6>())) {
      err = "cudaFuncSetAttribute(ozaki_gemm_kernel)";
      return false;
    }
    if (!(set_slice_attr<8>() && set_slice_attr<7>() && set_slice_attr<6>() && set_slice_attr<4>())) {
      err = "cudaFuncSetAttribute(slice_rows_kernel)";
      return false;
    }
    ready = true;
    return true;
  }
  bool set_shape(int N_, cudaStream_t st) {
    N = N_;
    Np = (N + kTile - 1) / kTile * kTile;
    if (tilesNp != Np) {
      const int nt = Np / kTile;
      std::vector<int2> tl;
      for (int bj = 0; bj < nt; ++bj)
        for (int bi = 0; bi <= bj; ++bi) tl.push_back(make_int2(bi, bj));
      cudaFree(tiles_d);
      tiles_d = nullptr;
      if (cudaMalloc(&tiles_d, tl.size() * sizeof(int2)) != cudaSuccess) { err = "cudaMalloc tiles"; return false; }
      if (cudaMemcpyAsync(tiles_d, tl.data(), tl.size() * sizeof(int2), cudaMemcpyHostToDevice, st) != cudaSuccess) { err = "copy tiles"; return false; }
      cudaStreamSynchronize(st);
      ntiles = (int)tl.size();
      tilesNp = Np;
    }
    return true;
  }
  // slices of an N x N symmetric matrix (ld = N) into `sl` (padding rows / columns must have been cleared)
  bool slice(const T* M, Sliced& sl, cudaStream_t st) {
    const int staged = slice_stage_bytes(N) <= (size_t)kSliceStageMaxBytes ? 1 : 0;
    const size_t sm = staged ? slice_stage_bytes(N) : 0;
    if (k == 8) slice_rows_kernel<T, 8><<<N, 256, sm, st>>>(M, N, Np, sl.d, sl.scale, staged);
    else if (k == 7) slice_rows_kernel<T, 7><<<N, 256, sm, st>>>(M, N, Np, sl.d, sl.scale, staged);
    else if (k == 6) slice_rows_kernel<T, 6><<<N, 256, sm, st>>>(M, N, Np, sl.d, sl.scale, staged);
    else slice_rows_kernel<T, 4><<<N, 256, sm, st>>>(M, N, Np, sl.d, sl.scale, staged);
    return cudaGetLastError() == cudaSuccess;
  }
  // out = c0 (A B) + c1 D + c2 I   (+ reductions into partial[2 * ntiles])
  bool gemm(Sliced& A, Sliced& B, T* out, const T* D, const T* E, int e_identity, const double* coef_d, double* partial,
            cudaStream_t st) {
    if (!A.make_map(Np) || !B.make_map(Np)) { err = "cuTensorMapEncodeTiled failed"; return false; }
    GemmArgs<T> a;
    a.N = N; a.Np = Np; a.ntiles = ntiles; a.store = out ? 1 : 0;
    a.tiles = tiles_d; a.scaleA = A.scale; a.scaleB = B.scale; a.out = out; a.D = D; a.E = E; a.e_identity = e_identity;
    a.coef = coef_d; a.partial = partial;
    const int grid = std::min(ntiles, num_sms);
    if (k == 8 && g == 10) ozaki_gemm_kernel<T, 8, 10><<<grid, kThreads, smem_bytes(), st>>>(A.map, B.map, a);
    else if (k == 8) ozaki_gemm_kernel<T, 8, 8><<<grid, kThreads, smem_bytes(), st>>>(A.map, B.map, a);
    else if (k == 7) ozaki_gemm_kernel<T, 7, 7><<<grid, kThreads, smem_bytes(), st>>>(A.map, B.map, a);
    else if (k == 6) ozaki_gemm_kernel<T, 6, 8><<<grid, kThreads, smem_bytes(), st>>>(A.map, B.map, a);
    else ozaki_gemm_kernel<T, 4, 6><<<grid, kThreads, smem_bytes(), st>>>(A.map, B.map, a);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { err = std::string("ozaki_gemm_kernel launch: ") + cudaGetErrorString(e); return false; }
    return true;
  }
};

}  // namespace tc
}  // namespace cosmo
