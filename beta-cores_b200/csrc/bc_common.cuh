// beta-cores B200: shared device helpers (sm_100a).
//   * mbarrier + cp.async.bulk (TMA engine, SASS UBLKCP) staging primitives
//   * FP64 tensor-core MMA (mma.sync m16n8k4 f64 -> SASS DMMA.8x8x4; tcgen05 has no f64 kind)
//   * double-double accumulation (order-insensitive S-vector sums)
//   * numpy-compatible arg-max ordering (NaN first, then larger value, then lower index)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "bc_npmean.h"

namespace bc {

// ------------------------------------------------------------- debug build --
// -DBC_DEBUG (tools/build_variants.py debug="-DBC_DEBUG"; the GPU suite is run against that library once per round, since
// compute-sanitizer is not available on this pool): device-side checks of the invariants a racy or mis-phased pipeline would
// break -- ring-slot and TMEM addressing bounds, tile bounds of the bulk copies -- and a watchdog on every mbarrier wait
// (a lost arrival becomes a printed trap instead of a hang).
#if defined(BC_DEBUG)
#include <stdio.h>
#define BC_DASSERT(cond)                                                                                      \
  do {                                                                                                        \
    if (!(cond)) {                                                                                            \
      printf("BC_DEBUG assert failed: %s  (%s:%d, block %d thread %d)\n", #cond, __FILE__, __LINE__, (int)blockIdx.x, \
             (int)threadIdx.x);                                                                               \
      __trap();                                                                                               \
    }                                                                                                         \
  } while (0)
#else
#define BC_DASSERT(cond) ((void)0)
#endif

// ---------------------------------------------------------------- mbarrier --
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
#if defined(BC_DEBUG)
  // watchdog form: ~2^26 polls (seconds) without the phase flipping = a lost arrival / wrong parity somewhere in the pipeline
  unsigned long long spins = 0;
  while (true) {
    uint32_t done = 0;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (done) return;
    if (++spins > (1ull << 26)) {
      printf("BC_DEBUG mbarrier watchdog: block %d thread %d waits on barrier at smem 0x%x parity %u\n", (int)blockIdx.x, (int)threadIdx.x,
             smem_u32(bar), parity);
      __trap();
    }
  }
#else
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra WAIT_DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
#endif
}
// 1-D bulk copy global -> shared through the TMA engine; completes `bytes` on `bar`.
// Requirements: dst, src 16-byte aligned, bytes a multiple of 16.
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// -------------------------------------------------------------------- DMMA --
// D(16x8) += A(16x4, row) * B(4x8, col).  Fragment layout (g = lane>>2, t = lane&3):
//   a0 = A[g][t], a1 = A[g+8][t];  b0 = B[k=t][n=g];
//   c0 = C[g][2t], c1 = C[g][2t+1], c2 = C[g+8][2t], c3 = C[g+8][2t+1].
__device__ __forceinline__ void dmma_m16n8k4(double (&c)[4], double a0, double a1, double b0) {
  asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
               : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
               : "d"(a0), "d"(a1), "d"(b0));
}

// ----------------------------------------------------------- double-double --
struct dd {
  double hi, lo;
};
__device__ __forceinline__ dd dd_add_d(dd a, double b) {  // Knuth TwoSum + renormalise
  double s = a.hi + b;
  double bb = s - a.hi;
  double e = (a.hi - (s - bb)) + (b - bb);
  e += a.lo;
  dd r;
  r.hi = s + e;
  r.lo = e - (r.hi - s);
  return r;
}
__device__ __forceinline__ dd dd_add(dd a, dd b) {
  double s = a.hi + b.hi;
  double bb = s - a.hi;
  double e = (a.hi - (s - bb)) + (b.hi - bb);
  e += a.lo + b.lo;
  dd r;
  r.hi = s + e;
  r.lo = e - (r.hi - s);
  return r;
}

// numpy's rounding of the mean of a constant row (np_sum_const / np_centred_const / np_score_const): bc_npmean.h

// ----------------------------------------------------------------- arg-max --
// numpy semantics: np.argmax returns the FIRST NaN if any NaN is present, else the first
// occurrence of the maximum; np.max returns NaN in the first case.
struct Best {
  double v;
  long long i;  // -1 = empty
};
__device__ __forceinline__ bool best_better(double va, long long ia, double vb, long long ib) {
  // is (va, ia) preferred over (vb, ib)?
  if (ib < 0) return ia >= 0;
  if (ia < 0) return false;
  bool na = isnan(va), nb = isnan(vb);
  if (na != nb) return na;
  if (na) return ia < ib;
  if (va != vb) return va > vb;
  return ia < ib;
}
__device__ __forceinline__ Best best_merge(Best a, Best b) { return best_better(b.v, b.i, a.v, a.i) ? b : a; }
__device__ __forceinline__ Best best_warp(Best x) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    Best y;
    y.v = __shfl_xor_sync(0xffffffffu, x.v, o);
    y.i = __shfl_xor_sync(0xffffffffu, x.i, o);
    x = best_merge(x, y);
  }
  return x;
}
__device__ __forceinline__ double warp_sum(double x) {  // fixed butterfly order: deterministic
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
  return x;
}
// np.max semantics for a running maximum (NaN sticks)
__device__ __forceinline__ double nanmax(double a, double b) {
  if (isnan(a) || isnan(b)) return a + b;
  return a > b ? a : b;
}

}  // namespace bc
