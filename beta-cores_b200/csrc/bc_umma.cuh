// beta-cores B200: tcgen05 / TMEM primitives (sm_100a inline PTX) used by the quantised projection kernel.
//   * TMEM allocation, tcgen05.mma.kind::i8 (int8 x int8 -> int32 accumulators in TMEM), tcgen05.commit -> mbarrier,
//     tcgen05.ld (TMEM -> registers), the tcgen05 fences
//   * shared-memory matrix descriptors for the K-major, 128-byte-swizzled operand image the quantisers write
//   * the exact int64 -> fp64 conversion used to recombine the digit diagonals
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace bc {

// ---- operand image geometry (shared by the quantisers and the projection kernel) ----
constexpr int kQSlices = 7;                            // signed base-256 digits per value (55 bits + sign)
constexpr int kQK = 128;                               // contraction length in bytes (= features, zero padded)
constexpr int kQTileRows = 128;                        // data rows per tile = UMMA M
constexpr int kQChunk = 32;                            // sample columns per chunk
constexpr int kQSliceA = kQTileRows * kQK;             // 16 KB: one digit plane of a row tile
constexpr int kQTileBytes = kQSlices * kQSliceA;       // 112 KB
constexpr int kQSliceB = kQChunk * kQK;                // 4 KB: one digit plane of a sample chunk
constexpr int kQChunkBytes = kQSlices * kQSliceB;      // 28 KB
constexpr int kQDiagCols = kQSlices * kQChunk;         // 224 TMEM columns per accumulator buffer

// byte offset of element (row r, contraction index c) inside one digit plane: 128-byte rows, the 16-byte chunk index
// XOR-ed with (r mod 8) -- the SWIZZLE_128B pattern tcgen05.mma expects for K-major operands (8-row x 128-byte atoms,
// atom base 1024-byte aligned).
__host__ __device__ __forceinline__ uint32_t q_swizzle_off(uint32_t r, uint32_t c) {
  return r * 128u + ((((c >> 4) ^ (r & 7u)) << 4) | (c & 15u));
}

#if defined(__CUDACC__)
// K-major SWIZZLE_128B shared-memory matrix descriptor: start address (>>4), leading byte offset (unused for swizzled
// K-major, 1), stride byte offset = 1024 B between 8-row groups, descriptor version 1 (sm_100), layout type 2.
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
  return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(1024u >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// instruction descriptor for kind::i8: D = s32, A = B = signed int8, both K-major, M = 128, N given.
__device__ __forceinline__ uint32_t umma_idesc_i8(int N) {
  return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
}
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {  // whole warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                   static_cast<uint32_t>(__cvta_generic_to_shared(dst_smem))),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {  // whole warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// D[tmem] (+)= A[smem] * B[smem]^T, one thread issues for the CTA.
__device__ __forceinline__ void umma_i8(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on an mbarrier once every tcgen05.mma issued so far by this thread has completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   static_cast<uint32_t>(__cvta_generic_to_shared(bar)))
               : "memory");
}
// 32 lanes x 8 consecutive 32-bit columns: thread t of the warp gets TMEM lane (lane base of taddr) + t
__device__ __forceinline__ void tmem_ld_x8(uint32_t taddr, uint32_t (&v)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_x4(uint32_t taddr, uint32_t (&v)[4]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3])
               : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_x1(uint32_t taddr, uint32_t& v) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(v) : "r"(taddr));
}
// one lane of a converged warp (elect.sync): the predicate ptxas recognises as "single thread, uniform operands"
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// conversion of an int64 to fp64, round to nearest (exact for |t| < 2^53) (I2F.F64.S64 on the conversion unit: measured a little faster here than
// the integer-add + FP64-add magic-number form, which occupies the contended FP64 pipe)
__device__ __forceinline__ double exact_ll2d(long long t) { return __ll2double_rn(t); }
// the NS int32 digit diagonals D_d = sum_{i+j=d} A_i B_j^T  ->  sum_d D_d 256^(NS-1-d)  as an fp64 (one rounding):
// the last three diagonals and the leading NS-3 are summed exactly in two int64 halves (integer pipe), each below 2^53.
template <int NS>
__device__ __forceinline__ double q_combine_n(const int (&d)[NS]) {
  static_assert(NS >= 4 && NS <= 7, "digit count");
  if (NS <= 6) {
    // |D_d| <= (d+1) 2^21, so the whole sum stays below 2^62: one int64 and ONE conversion (same single rounding as the
    // two-half form below, which the 7-digit split needs)
    // the four trailing digits by 32 x 32 -> 64 multiply-adds (IMAD.WIDE), the leading ones straight into the high word
    long long t = (long long)d[NS - 1];
    t += (long long)d[NS - 2] * 256;
    t += (long long)d[NS - 3] * 65536;
    t += (long long)d[NS - 4] * 16777216;
    if (NS >= 5) {
      int hi = (int)(t >> 32) + d[NS - 5];
      if (NS >= 6) hi += d[NS - 6] << 8;
      t = (long long)(((unsigned long long)(unsigned)hi << 32) | (unsigned long long)(unsigned)(int)t);
    }
    return exact_ll2d(t);
  }
  long long hi = 0;
#pragma unroll
  for (int i = 0; i < NS - 3; ++i) hi = (hi << 8) + (long long)d[i];
  const long long lo = ((long long)d[NS - 3] << 16) + ((long long)d[NS - 2] << 8) + (long long)d[NS - 1];
  return fma(exact_ll2d(hi), 16777216.0, exact_ll2d(lo));
}
__device__ __forceinline__ double q_combine(int d0, int d1, int d2, int d3, int d4, int d5, int d6) {
  const int d[7] = {d0, d1, d2, d3, d4, d5, d6};
  return q_combine_n<7>(d);
}
// register budget hand-over between warpgroups of one CTA (sm_90+): the service warps give registers back, the epilogue
// warps take them (the CTA's pool is launch registers x threads; both counts are multiples of 8)
template <int N>
__device__ __forceinline__ void reg_dealloc() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N>
__device__ __forceinline__ void reg_alloc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
#endif

}  // namespace bc
