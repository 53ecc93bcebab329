// Shared device helpers for libscgrhc (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "scgrhc.h"

namespace scgrhc {

constexpr int kWarp = 32;
constexpr unsigned kFull = 0xffffffffu;

// ---- mbarrier + 1-D bulk async copy (TMA engine, SASS: UBLKCP / SYNCS) ------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// try_wait suspends the thread for a hardware-defined time slice before it reports "not yet", so this loop is not a hot
// spin and needs no software back-off.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t"
      "}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// global -> shared bulk copy; dst/src 16-byte aligned, bytes a multiple of 16.
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
      "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

// ---- warp reductions over doubles ----------------------------------------------------------------
__device__ __forceinline__ double shfl_xor_d(double v, int m) { return __shfl_xor_sync(kFull, v, m); }

// Exact fp64 max over a warp with two 32-bit REDUX ops instead of five 64-bit shuffle rounds: map the
// double to an order-preserving unsigned 64-bit key, reduce the high words, then the low words among
// the lanes that hold the winning high word.  Inputs must not be NaN (callers keep NaN out of their
// running extrema); +-Inf and +-0 order correctly (-0 < +0 in key order, both compare equal as values).
__device__ __forceinline__ double warp_max(double v) {
  const uint32_t hi = (uint32_t)__double2hiint(v), lo = (uint32_t)__double2loint(v);
  const uint32_t m = (uint32_t)((int32_t)hi >> 31);          // all ones for negative values
  const uint32_t khi = hi ^ (m | 0x80000000u), klo = lo ^ m;
  const uint32_t mh = __reduce_max_sync(kFull, khi);
  const uint32_t ml = __reduce_max_sync(kFull, khi == mh ? klo : 0u);
  const uint32_t im = (mh & 0x80000000u) ? 0u : 0xffffffffu;  // winner was negative: undo the complement
  return __hiloint2double((int)(mh ^ (im | 0x80000000u)), (int)(ml ^ im));
}
__device__ __forceinline__ double warp_min(double v) { return -warp_max(-v); }

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int m = 16; m; m >>= 1) v = __dadd_rn(v, shfl_xor_d(v, m));
  return v;
}

// Three sums at once with a transposed butterfly: after the first two rounds every lane carries ONE of the three
// values, so the tree costs 12 shuffles and 6 additions instead of 30 and 15.  On return lane 0 holds the warp's sum of
// a, lane 8 that of b, lane 16 (and 24) that of c, in `a`; other lanes hold partial garbage.  Deterministic order.
__device__ __forceinline__ double warp_sum3(double a, double b, double c, int lane) {
  const bool up = (lane & 16) != 0, b3 = (lane & 8) != 0;
  // round 1 (xor 16): the lower half keeps (a, b) and gives away c, the upper half keeps c and gives away (a, b)
  const double r1 = shfl_xor_d(up ? a : c, 16), r2 = shfl_xor_d(b, 16);
  const double A = __dadd_rn(up ? c : a, r1), B = __dadd_rn(b, r2);   // lower: (a, b); upper: (c, junk)
  // round 2 (xor 8): lower half: bit 3 clear keeps a / gives b, bit 3 set keeps b / gives a; upper half: plain step on c
  const bool kb = !up && b3;
  const double keep = kb ? B : A, give = (up || b3) ? A : B;
  double v = __dadd_rn(keep, shfl_xor_d(give, 8));
#pragma unroll
  for (int m = 4; m; m >>= 1) v = __dadd_rn(v, shfl_xor_d(v, m));
  return v;
}

// Correctly rounded a/d from the correctly rounded reciprocal inv = RN(1/d) (Markstein): two FMA
// residual corrections.  Valid when no intermediate underflows; the caller guarantees that per window
// (see Normaliser::slow).  Validated against IEEE division in tests (scgrhc_selftest_div).
__device__ __forceinline__ double div_by_recip(double a, double d, double inv) {
  double q = __dmul_rn(a, inv);
  double r = __fma_rn(-d, q, a);
  q = __fma_rn(r, inv, q);
  r = __fma_rn(-d, q, a);
  q = __fma_rn(r, inv, q);
  return q;
}

// Tier 1 of the fp32 normalisation (window_kernel.cuh): q0 = RN(a * RN(1/d)) rounds to the same float as the exact
// quotient unless the low 29 bits of its mantissa lie within [0x0ffffff8, 0x10000008] (8 ulp64 around a float rounding
// boundary).  key = 8 * (distance of those bits from 0x0ffffff8), wrapping: risky <=> key <= kTier1Risky.  One shift-add.
constexpr uint32_t kTier1Risky = 128u;
__device__ __forceinline__ uint32_t tier1_key(double q0) { return ((uint32_t)__double2loint(q0) << 3) - 0x7fffffc0u; }

// streaming stores (outputs are written once and not re-read by this kernel)
__device__ __forceinline__ void st_cs(float* p, float v) { __stcs(p, v); }
__device__ __forceinline__ void st_cs(double* p, double v) { __stcs(p, v); }

}  // namespace scgrhc
