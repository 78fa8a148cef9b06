// beta-cores B200: stage 1 (+2), tensor-core route -- the fused N x S projection with the contraction on tcgen05.
//
// Same contract as bc_project.cu (k_project): replaces, without materialising the N x S matrix,
//   BetaBlackBoxProjector.project_f / BlackBoxProjector.project   (bayesiancoresets/coreset/projector.py:51-55, :23-26)
//   the column sum / residual correlation of BetaCoreset._select   (bayesiancoresets/coreset/bcores.py:77-81)
//   the column sum of BetaCoreset._optimize's grd()                (bayesiancoresets/coreset/bcores.py:142-146)
//
// The reference contracts in fp64 (numpy dgemm).  tcgen05 has no f64 kind, and the legacy FP64 DMMA shares the FP64 pipe
// with the potential's polynomials, so here the contraction is done EXACTLY in integers (Ozaki splitting):
//   z_n = 2^(e_n) sum_i a_i 256^-(i+1),  theta_s = 2^(e_s) sum_j b_j 256^-(j+1),   a_i, b_j int8 digits (7 each, 55 bits)
//   z_n . theta_s = 2^(e_n + e_s) sum_d 256^-(d+2) D_d,     D_d = sum_{i+j=d} a_i . b_j   (int32, exact)
// keeping the diagonals d <= 6 (28 of the 49 digit pairs; the dropped ones are below 6*128*2^-58 of |z|max |theta|max).
// Every row carries its own exponent e_n, the samples share one, and both operands are first rescaled per FEATURE by a power
// of two (z_k 2^-c_k, theta_k 2^+c_k, c_k = ilogb max_n |z_nk|: exact, products unchanged) so that columns of very different
// magnitude keep their bits -- see k_feature_absmax below.
// tcgen05.mma.kind::i8 computes D_0..D_6 for a 128-row x 32-sample chunk into 7 x 32 TMEM columns: the MMA for row digit
// i multiplies against the sample digits 0..6-i STACKED along N (N = 32 (7 - i)), i.e. 7 x 4 instructions per chunk.
//
// The images always hold 7 digits; a launch may contract only the leading NS of them (NS = 5, 6, 7: 15 / 21 / 28 digit
// pairs, diagonals d < NS) -- the operands are then read as if rounded to 8 NS - 2 bits below the row / sample maximum
// (balanced digits: dropping the tail is a round-to-nearest), bc_set_contraction_digits.
//
// One persistent CTA per SM, 20 warps (five warpgroups: the register file is re-divided with setmaxnreg so that the sixteen
// epilogue warps run with 112 registers instead of the 96 a 640-thread launch gets, the four service warps with 32):
//   warps 0-15 : four epilogue groups of four warps (a warp may only read its own TMEM lane quarter).  Groups 0,1 take
//                the even chunks (accumulator buffer 0), groups 2,3 the odd chunks (buffer 1); within a pair each group
//                owns 16 of the chunk's 32 sample columns.  Thread = data row (TMEM lane): tcgen05.ld the 7 diagonals,
//                recombine in int64 (integer pipe), convert exactly (conversion unit), apply the model's potential four columns at a time
//                (FP64 pipe, four independent dependency chains), pivot shift, per-row statistics in registers, column
//                partials by a transposed warp butterfly -> shared-memory ring.  Four warps per SM sub-partition is what
//                keeps the FP64 pipe fed (8-cycle DFMA latency, measured tools/fp64_ipc.cu).
//   warp 16    : producer -- cp.async.bulk (TMA engine) of the row tile (112 KB, once per tile) and of the sample
//                chunks (28 KB, 3-stage ring).  Both images are stored in HBM already in the swizzled layout the tensor
//                core reads, so a copy is one contiguous burst.
//   warp 17    : MMA issuer (warp-uniform loop, one elected lane issues) + TMEM allocation.
//   warp 18    : reducer -- adds the four 32-row column partials of a chunk in fixed order and accumulates them per
//                column in double-double (order-insensitive S-vector).
#include "bc_common.cuh"
#include "bc_models.cuh"
#include "bc_kernels.h"
#include "bc_umma.cuh"

namespace bc {

constexpr int kQStagesB = 3;
constexpr int kQSlots = 4;
constexpr int kQGroups = 4;                       // epilogue groups of 4 warps (one warp per TMEM lane quarter)
constexpr int kQEpiWarps = 4 * kQGroups;          // 16
constexpr int kQThreads = (kQEpiWarps + 4) * 32;  // + producer, MMA issuer, reducer, one idle warp (setmaxnreg works on whole warpgroups)
#ifndef BC_Q_REGS_EPI          // experiment knobs (tools/build_variants.py); the defaults are what ships
#define BC_Q_REGS_EPI 112
#define BC_Q_REGS_SVC 32
#endif
constexpr int kQRegsEpi = BC_Q_REGS_EPI;          // 16 x 32 x 112 + 4 x 32 x 32 = 61440 = the CTA's pool at 96 registers x 640 threads
constexpr int kQRegsSvc = BC_Q_REGS_SVC;
constexpr int kQHalfCols = kQChunk / 2;           // columns of a chunk one group handles

struct QSmem {
  static constexpr size_t a = 0;
  static constexpr size_t b = a + kQTileBytes;
  static constexpr size_t ring = b + (size_t)kQStagesB * kQChunkBytes;    // [kQSlots][4 quarters][32 cols]
  static constexpr size_t pivs = ring + (size_t)kQSlots * 4 * kQChunk * 8;  // [2][128]
  static constexpr size_t stats = pivs + 2 * 128 * 8;                       // [3 groups][128][3]
  static constexpr size_t fin = stats + 3 * 128 * 3 * 8;                    // [4 warps][2]
  static constexpr size_t bars = fin + 4 * 2 * 8;
  static constexpr int nbars = 2 + 2 * kQStagesB + 4 + 2 * kQSlots + 2 + 2;
  static constexpr size_t tmem = bars + (size_t)nbars * 8;
  static constexpr size_t total = tmem + 16 + 1024;  // + slack to align the base to 1024 B
};
static_assert(QSmem::total <= kMaxSmem, "shared memory budget");

// mbarrier wait for the service warps (producer / MMA issuer / reducer): they are far off the critical path, and a
// spinning try_wait loop would steal issue slots from the epilogue warps that share their SM sub-partition.
__device__ __forceinline__ void mbar_wait_relaxed(uint64_t* bar, uint32_t parity) {
  uint32_t done = 0;
#if defined(BC_DEBUG)
  unsigned long long spins = 0;
#endif
  while (true) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (done) break;
    __nanosleep(64);
#if defined(BC_DEBUG)
    if (++spins > (1ull << 24)) {
      printf("BC_DEBUG mbarrier watchdog (service warp): block %d thread %d, barrier smem 0x%x parity %u\n", (int)blockIdx.x,
             (int)threadIdx.x, smem_u32(bar), parity);
      __trap();
    }
#endif
  }
}

// ------------------------------------------------------------------ quantisers --
// One warp per row; lane l owns contraction indices 4l .. 4l+3.  R = rows per image tile (128 data rows / 32 samples).
// Row r of the source is src + (r * ld); indices >= D are zero.  Writes the 7 digit planes of the row into the swizzled
// image and scale_out[r] = 2^(e - 32) (NaN if the row is not finite: poisons every contraction with it, like the
// reference's NaN propagation).  Rows in [n, tiles * R) are written as zeros.
template <int R>
__global__ void __launch_bounds__(256) k_quantise(const double* __restrict__ src, long long ld, long long n, int D,
                                                   unsigned char* __restrict__ image, double* __restrict__ scale_out,
                                                   double* __restrict__ aux_out, int aux_col, const unsigned long long* common_sc,
                                                   unsigned long long* common_sc_next, int* common_e_out, double* common_scale_out,
                                                   const int* __restrict__ fexp, int fsign) {
  const int lane = threadIdx.x & 31;
  // common_sc (the samples): every row shares ONE exponent, ilogb(max |x|) + 3 over the whole matrix -- their magnitudes are
  // alike, and a single scale keeps the projection epilogue free of per-column loads.  common_sc[0] = the bit pattern of
  // that maximum, common_sc[1] != 0 if any entry is not finite (gathered by k_prepare_rows / k_sample_absmax); block 0
  // publishes (exponent, scale = 2^(e - 32), NaN for a non-finite sample set: the reference's own result is NaN in every
  // quantity downstream of it) for the projection kernels and clears the OTHER scratch slot for the next sample set.
  int ce = 0;
  if (common_sc) {
    const double cmax = __longlong_as_double((long long)common_sc[0]);
    const bool cbad = common_sc[1] != 0;
    ce = (cmax > 0.0 && !cbad) ? ilogb(cmax) + 3 : 0;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
      *common_e_out = ce;
      *common_scale_out = cbad ? __longlong_as_double(0x7ff8000000000000LL) : scalbn(1.0, ce - 32);
      common_sc_next[0] = 0;
      common_sc_next[1] = 0;
    }
  }
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const long long rows_padded = ((n + R - 1) / R) * R;
  for (long long r = warp; r < rows_padded; r += nwarps) {
    double x[4] = {0.0, 0.0, 0.0, 0.0};
    if (r < n) {
      const double* p = src + r * ld;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int k = lane * 4 + i;
        if (k < D) x[i] = fexp ? scalbn(p[k], fsign * __ldg(fexp + k)) : p[k];   // feature exponents: see k_feature_exponents
      }
      if (aux_out && lane == 0) aux_out[r] = p[aux_col];
    }
    double amax = fmax(fmax(fabs(x[0]), fabs(x[1])), fmax(fabs(x[2]), fabs(x[3])));
    bool bad = !(isfinite(x[0]) && isfinite(x[1]) && isfinite(x[2]) && isfinite(x[3]));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      amax = fmax(amax, __shfl_xor_sync(0xffffffffu, amax, o));
      const int ob = __shfl_xor_sync(0xffffffffu, (int)bad, o);
      bad = bad || (ob != 0);
    }
    // |x| 2^-e < 1/4  =>  top digit within [-64, 64], the others in [-128, 127]
    const int e = common_sc ? ce : ((amax > 0.0 && !bad) ? ilogb(amax) + 3 : 0);
    uint32_t dig[kQSlices] = {0, 0, 0, 0, 0, 0, 0};
    if (!bad) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        long long Q = __double2ll_rn(scalbn(x[i], 56 - e));  // |Q| <= 2^54
#pragma unroll
        for (int s = kQSlices - 1; s >= 1; --s) {
          const long long d = (long long)(signed char)(Q & 0xff);
          dig[s] |= (uint32_t)(d & 0xff) << (8 * i);
          Q = (Q - d) >> 8;
        }
        dig[0] |= (uint32_t)(Q & 0xff) << (8 * i);
      }
    }
    const long long tile = r / R;
    const uint32_t rr = (uint32_t)(r % R);
    unsigned char* base = image + (size_t)tile * (kQSlices * R * kQK) + q_swizzle_off(rr, (uint32_t)lane * 4u);
#pragma unroll
    for (int s = 0; s < kQSlices; ++s) *reinterpret_cast<uint32_t*>(base + (size_t)s * (R * kQK)) = dig[s];
    if (scale_out && r < n && lane == 0) scale_out[r] = bad ? __longlong_as_double(0x7ff8000000000000LL) : scalbn(1.0, e - 32);
  }
}

// Feature exponents.  The digit split keeps 56 bits below the largest entry of a ROW, so a feature whose entries are all
// tiny next to another feature's (unstandardised data: an income column beside a 0/1 flag) would keep few significant
// bits, while an fp64 dot product rounds every product relative to itself.  Both operands are therefore rescaled per
// feature by a power of two -- x_k 2^-c_k with theta_k 2^+c_k, c_k = ilogb(max_n |x_nk|): exact, the products are
// unchanged, and every feature of the row image is O(1).  out[k] = c_k, or kFexpEmpty for a column with no finite
// non-zero entry (the caller maps it to 0 after taking the maximum over row shards).
constexpr int kFexpEmpty = -(1 << 30);
__global__ void __launch_bounds__(256) k_feature_absmax(const double* __restrict__ X, long long ld, long long n, int D,
                                                        unsigned long long* __restrict__ colmax_bits) {
  // blockDim = (32, 8): lane -> features lane + 32 j, 8 row groups per block; |x| compares like its bit pattern
  unsigned long long m[4] = {0, 0, 0, 0};
  for (long long r = (long long)blockIdx.x * blockDim.y + threadIdx.y; r < n; r += (long long)gridDim.x * blockDim.y) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = threadIdx.x + 32 * j;
      if (k < D) {
        const double v = fabs(X[r * ld + k]);
        const unsigned long long b = (unsigned long long)__double_as_longlong(v);
        if (isfinite(v) && b > m[j]) m[j] = b;
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int k = threadIdx.x + 32 * j;
    if (k < D && m[j]) atomicMax(colmax_bits + k, m[j]);
  }
}
__global__ void k_feature_exponents(const unsigned long long* __restrict__ colmax_bits, int D, int* __restrict__ out) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < D) {
    const double v = __longlong_as_double((long long)colmax_bits[k]);
    out[k] = (v > 0.0) ? ilogb(v) : kFexpEmpty;
  }
}

cudaError_t launch_feature_exponents(const double* X, long long ldx, long long n, int D, unsigned long long* scratch, int* out,
                                     cudaStream_t st) {
  cudaError_t e = cudaMemsetAsync(scratch, 0, sizeof(unsigned long long) * kQK, st);
  if (e != cudaSuccess) return e;
  if (n > 0) {
    long long blocks = (n + 63) / 64;
    if (blocks > 148 * 8) blocks = 148 * 8;
    k_feature_absmax<<<(int)blocks, dim3(32, 8), 0, st>>>(X, ldx, n, D, scratch);
  }
  k_feature_exponents<<<1, kQK, 0, st>>>(scratch, D, out);
  return cudaGetLastError();
}

cudaError_t launch_quantise_rows(const double* X, long long ldx, long long n, int D, unsigned char* image, double* rowscale,
                                 double* aux_out, int aux_col, const int* fexp, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  long long warps = ((n + kQTileRows - 1) / kQTileRows) * kQTileRows;
  long long blocks = (warps + 7) / 8;
  if (blocks > 148 * 16) blocks = 148 * 16;
  k_quantise<kQTileRows><<<(int)blocks, 256, 0, st>>>(X, ldx, n, D, image, rowscale, aux_out, aux_col, nullptr, nullptr, nullptr, nullptr, fexp, -1);
  return cudaGetLastError();
}

// Gathered row image: rows idx[0..n) of a quantised image (duplicates allowed, the reference sub-samples WITH replacement,
// bcores.py:53) re-packed into a compact image of ceil(n / 128) tiles, each row moved to the swizzle slot of its new
// position, with its scale (and optional per-row auxiliary value).  One warp per output row, lane l moves the 4 bytes of
// contraction indices 4l..4l+3 of every digit plane: 7 coalesced 128-byte reads and writes per row.  This is what puts
// sub-sampled passes (n_subsample_select / n_subsample_opt) and group passes on the tensor-core route.
__global__ void __launch_bounds__(256) k_gather_image(const unsigned char* __restrict__ src, const double* __restrict__ src_scale,
                                                      const double* __restrict__ src_aux, const long long* __restrict__ idx, long long n,
                                                      unsigned char* __restrict__ dst, double* __restrict__ dst_scale,
                                                      double* __restrict__ dst_aux) {
  const int lane = threadIdx.x & 31;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const long long rows_padded = ((n + kQTileRows - 1) / kQTileRows) * kQTileRows;
  for (long long r = warp; r < rows_padded; r += nwarps) {
    unsigned char* out = dst + (size_t)(r / kQTileRows) * kQTileBytes + q_swizzle_off((uint32_t)(r % kQTileRows), (uint32_t)lane * 4u);
    if (r < n) {
      const long long g = __ldg(idx + r);
      const unsigned char* in = src + (size_t)(g / kQTileRows) * kQTileBytes + q_swizzle_off((uint32_t)(g % kQTileRows), (uint32_t)lane * 4u);
#pragma unroll
      for (int sl = 0; sl < kQSlices; ++sl)
        *reinterpret_cast<uint32_t*>(out + (size_t)sl * kQSliceA) = __ldg(reinterpret_cast<const uint32_t*>(in + (size_t)sl * kQSliceA));
      if (lane == 0) {
        dst_scale[r] = __ldg(src_scale + g);
        if (dst_aux) dst_aux[r] = __ldg(src_aux + g);
      }
    } else {
#pragma unroll
      for (int sl = 0; sl < kQSlices; ++sl) *reinterpret_cast<uint32_t*>(out + (size_t)sl * kQSliceA) = 0u;
    }
  }
}

cudaError_t launch_gather_image(const unsigned char* src, const double* src_scale, const double* src_aux, const long long* idx, long long n,
                                unsigned char* dst, double* dst_scale, double* dst_aux, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  const long long rows_padded = ((n + kQTileRows - 1) / kQTileRows) * kQTileRows;
  long long blocks = (rows_padded + 7) / 8;
  if (blocks > 148 * 8) blocks = 148 * 8;
  k_gather_image<<<(int)blocks, 256, 0, st>>>(src, src_scale, src_aux, idx, n, dst, dst_scale, dst_aux);
  return cudaGetLastError();
}

// exponent shared by all S samples: ilogb(max |B'|) + 3 over the feature-scaled samples B'[s][k] = B[s][k] 2^fexp[k].
// A grid-wide max of the bit patterns (|x| orders like its bits; a non-finite entry sets the flag word) into one of two
// scratch slots; k_quantise turns it into (e, scale) and clears the other slot, which the next sample set will use.  The
// max is normally gathered by k_prepare_rows as it writes B (bc_small.cu); this kernel serves the callers that only have B.
__global__ void __launch_bounds__(256) k_sample_absmax(const double* __restrict__ B, int ldb, int S, int D, const int* __restrict__ fexp,
                                                       unsigned long long* __restrict__ scratch /* [0] max bits, [1] bad */) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  unsigned long long m = 0, bad = 0;
  for (int s = blockIdx.x * nw + wid; s < S; s += gridDim.x * nw) {
    for (int k = lane; k < D; k += 32) {
      const double b = B[(size_t)s * ldb + k];
      const double v = fabs(fexp ? scalbn(b, __ldg(fexp + k)) : b);
      if (!isfinite(v)) bad = 1;
      const unsigned long long bits = (unsigned long long)__double_as_longlong(v);
      if (isfinite(v) && bits > m) m = bits;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long om = __shfl_xor_sync(0xffffffffu, m, o);
    m = om > m ? om : m;
    bad |= __shfl_xor_sync(0xffffffffu, bad, o);
  }
  if (lane == 0) {
    if (m) atomicMax(scratch, m);
    if (bad) atomicOr(scratch + 1, 1ull);
  }
}

cudaError_t launch_quantise_samples(const double* B, int ldb, int S, int D, unsigned char* image, double* colscale, int* common_e,
                                    const int* fexp, unsigned long long* scratch4 /* two slots of two words */, int slot, bool have_max,
                                    cudaStream_t st) {
  unsigned long long* cur = scratch4 + 2 * (slot & 1);
  unsigned long long* nxt = scratch4 + 2 * ((slot & 1) ^ 1);
  if (!have_max) k_sample_absmax<<<(S + 7) / 8 < 148 ? (S + 7) / 8 : 148, 256, 0, st>>>(B, ldb, S, D, fexp, cur);
  const int rows = ((S + kQChunk - 1) / kQChunk) * kQChunk;
  BC_PREFER_MAX_SHARED(k_quantise<kQChunk>);
  k_quantise<kQChunk><<<(rows + 7) / 8, 256, 0, st>>>(B, ldb, S, D, image, nullptr, nullptr, 0, cur, nxt, common_e, colscale, fexp, +1);
  return cudaGetLastError();
}

// ------------------------------------------------------------------ projection --
// The MMAs of one 128-row x 32-sample chunk: row digit i against the sample digits 0..6-i stacked along N, KS blocks of
// 32 bytes along K.  +32 bytes along K inside the swizzle atom = +2 in the (>>4) start-address field of a descriptor.
template <int KS, int NS>
__device__ __forceinline__ void issue_chunk_mmas(uint32_t d0, uint64_t adesc0, uint64_t bdesc0, const uint32_t (&idesc)[kQSlices]) {
#pragma unroll
  for (int i = 0; i < NS; ++i) {
#pragma unroll
    for (int k = 0; k < KS; ++k) {
      umma_i8(d0 + (uint32_t)(i * kQChunk), adesc0 + (uint64_t)(i * (kQSliceA >> 4) + k * 2), bdesc0 + (uint64_t)(k * 2), idesc[i],
              (i | k) ? 1u : 0u);
    }
  }
}

template <class F, int MODE, int NS>
__global__ void __launch_bounds__(kQThreads, 1) k_project_q(const QProjArgs P) {
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  unsigned char* As = smem + QSmem::a;
  unsigned char* Bs = smem + QSmem::b;
  double* ring = reinterpret_cast<double*>(smem + QSmem::ring);
  double* pivs = reinterpret_cast<double*>(smem + QSmem::pivs);
  double* stats = reinterpret_cast<double*>(smem + QSmem::stats);
  double* fin = reinterpret_cast<double*>(smem + QSmem::fin);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + QSmem::bars);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + QSmem::tmem);
  uint64_t* full_a = bars + 0;
  uint64_t* empty_a = bars + 1;
  uint64_t* full_b = bars + 2;                    // [kQStagesB]
  uint64_t* empty_b = full_b + kQStagesB;         // [kQStagesB]
  uint64_t* tmem_full = empty_b + kQStagesB;      // [2]
  uint64_t* tmem_empty = tmem_full + 2;           // [2]
  uint64_t* ring_full = tmem_empty + 2;           // [kQSlots]
  uint64_t* ring_empty = ring_full + kQSlots;     // [kQSlots]
  uint64_t* piv_full = ring_empty + kQSlots;      // [2]
  uint64_t* stats_full = piv_full + 2;            // [1]
  uint64_t* stats_empty = stats_full + 1;         // [1]

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int S = P.S;
  const long long n = P.n;
  // The loop bounds are formed INSIDE each role (BC_Q_BOUNDS): computed here, ahead of the setmaxnreg branches, ptxas keeps
  // them in local memory for the 32-register service roles and reloads them once per chunk in every role (2.7 M local
  // loads per 1M-row launch in profiles/r02_ncu_k_project_q_lane_tables.txt).  The volatile moves pin the computation
  // behind the role's own register budget.
#define BC_Q_BOUNDS                                                     \
  long long n_role;                                                     \
  int S_role;                                                           \
  asm volatile("mov.u64 %0, %1;" : "=l"(n_role) : "l"(P.n));            \
  asm volatile("mov.u32 %0, %1;" : "=r"(S_role) : "r"(P.S));            \
  const long long ntiles = (n_role + kQTileRows - 1) / kQTileRows;      \
  const int nchunks = (S_role + kQChunk - 1) / kQChunk;
  const bool want_cols = (MODE == QMODE_COLSUM);

  if (tid == 0) {
    mbar_init(full_a, 1);
    mbar_init(empty_a, 1);
    for (int i = 0; i < kQStagesB; ++i) {
      mbar_init(full_b + i, 1);
      mbar_init(empty_b + i, 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(tmem_full + i, 1);
      mbar_init(tmem_empty + i, 8);
      mbar_init(piv_full + i, 4);
    }
    for (int i = 0; i < kQSlots; ++i) {
      mbar_init(ring_full + i, 8);
      mbar_init(ring_empty + i, 1);
    }
    mbar_init(stats_full, 12);
    mbar_init(stats_empty, 4);
    mbar_fence_init();
  }
  if (warp == kQEpiWarps + 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  // (the TMEM base address is read from shared memory by each role that needs it -- a value defined here would live in
  //  local memory across the setmaxnreg branches, like the loop bounds)
#define BC_Q_TMEM_BASE const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);

  // setmaxnreg sits at the head of each role's branch: ptxas allocates registers for the code a setmaxnreg dominates
  if (warp == kQEpiWarps) {
    // ======================= producer: TMA-engine bulk copies =======================
    reg_dealloc<kQRegsSvc>();
    BC_Q_BOUNDS
    if (lane == 0) {
      uint32_t itb = 0, tcount = 0;
      for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++tcount) {
        mbar_wait_relaxed(empty_a, (tcount & 1) ^ 1);
        mbar_arrive_expect_tx(full_a, (uint32_t)(NS * kQSliceA));
        const unsigned char* srcA = P.imgA + (size_t)tile * kQTileBytes;
        BC_DASSERT(tile >= 0 && tile < ntiles);
        BC_DASSERT((reinterpret_cast<uintptr_t>(srcA) & 15) == 0 && (smem_u32(As) & 1023) == 0);
#pragma unroll
        for (int s = 0; s < NS; ++s) bulk_g2s(As + (size_t)s * kQSliceA, srcA + (size_t)s * kQSliceA, kQSliceA, full_a);
        for (int c = 0; c < nchunks; ++c, ++itb) {
          const uint32_t st = itb % kQStagesB, ph = (itb / kQStagesB) & 1;
          mbar_wait_relaxed(empty_b + st, ph ^ 1);
          BC_DASSERT(st < (uint32_t)kQStagesB && c < nchunks);
          mbar_arrive_expect_tx(full_b + st, (uint32_t)(NS * kQSliceB));
          bulk_g2s(Bs + (size_t)st * kQChunkBytes, P.imgB + (size_t)c * kQChunkBytes, NS * kQSliceB, full_b + st);
        }
      }
    }
    __syncwarp();
  } else if (warp == kQEpiWarps + 1) {
    // ======================= MMA issuer =======================
    reg_dealloc<kQRegsSvc>();
    // All 32 lanes run the loop so that control flow and operands (descriptors, TMEM addresses) stay warp-uniform -- ptxas
    // then feeds tcgen05.mma from uniform registers directly; inside an `if (lane == 0)` region it wraps every MMA in an
    // ELECT / R2UR.BROADCAST loop (89 cycles per MMA measured, more than the MMA takes to execute).  One elected lane issues.
    {
      BC_Q_BOUNDS
      BC_Q_TMEM_BASE
      uint32_t idesc[kQSlices];
#pragma unroll
      for (int i = 0; i < kQSlices; ++i) idesc[i] = umma_idesc_i8(kQChunk * (i < NS ? NS - i : 1));
      const uint64_t adesc0 = umma_desc_sw128(smem_u32(As));
      uint32_t itb = 0, tcount = 0, use0 = 0, use1 = 0;
      for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++tcount) {
        mbar_wait_relaxed(full_a, tcount & 1);
        for (int c = 0; c < nchunks; ++c, ++itb) {
          const uint32_t st = itb % kQStagesB, ph = (itb / kQStagesB) & 1;
          const int buf = c & 1;
          const uint32_t use = buf ? use1 : use0;
          mbar_wait_relaxed(tmem_empty + buf, (use & 1) ^ 1);
          if (buf) ++use1; else ++use0;
          mbar_wait_relaxed(full_b + st, ph);
          tc_fence_after();
          const uint64_t bdesc0 = umma_desc_sw128(smem_u32(Bs + (size_t)st * kQChunkBytes));
          const uint32_t d0 = tmem_base + (uint32_t)(buf * kQDiagCols);
          BC_DASSERT((tmem_base & 0xffffu) + (uint32_t)(buf * kQDiagCols + NS * kQChunk) <= 512u);   // accumulator columns inside the allocation
          BC_DASSERT(P.ksteps >= 1 && P.ksteps <= 4);
          if (elect_one()) {
            // K blocks beyond the feature count hold only zero digits: not multiplied (D = 20 issues 7 MMAs per chunk
            // instead of 28).  Fully unrolled per block count: the issue rate of this one thread is on the critical path.
            switch (P.ksteps) {
              case 1: issue_chunk_mmas<1, NS>(d0, adesc0, bdesc0, idesc); break;
              case 2: issue_chunk_mmas<2, NS>(d0, adesc0, bdesc0, idesc); break;
              case 3: issue_chunk_mmas<3, NS>(d0, adesc0, bdesc0, idesc); break;
              default: issue_chunk_mmas<4, NS>(d0, adesc0, bdesc0, idesc); break;
            }
            umma_commit(empty_b + st);
            umma_commit(tmem_full + buf);
            if (c == nchunks - 1) umma_commit(empty_a);
          }
          __syncwarp();
        }
      }
    }
    __syncwarp();
  } else if (warp == kQEpiWarps + 3) {
    reg_dealloc<kQRegsSvc>();   // idle: pads the service warps to a full warpgroup
  } else if (warp == kQEpiWarps + 2) {
    // ============ reducer: 4 partials per chunk -> double-double column accumulators ============
    reg_dealloc<kQRegsSvc>();
    BC_Q_BOUNDS
    if (want_cols) {
      double* acc_hi = P.part_colsum + (size_t)blockIdx.x * 2 * P.Sld;
      double* acc_lo = acc_hi + P.Sld;
      for (int i = lane; i < 2 * P.Sld; i += 32) __stcg(acc_hi + i, 0.0);
      __syncwarp();
      uint32_t itb = 0;
      for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        for (int c = 0; c < nchunks; ++c, ++itb) {
          const uint32_t slot = itb % kQSlots, ph = (itb / kQSlots) & 1;
          mbar_wait_relaxed(ring_full + slot, ph);
          const double* rs = ring + (size_t)slot * 4 * kQChunk + lane;
          double v = rs[0];
          v += rs[kQChunk];
          v += rs[2 * kQChunk];
          v += rs[3 * kQChunk];
          __syncwarp();
          if (lane == 0) mbar_arrive(ring_empty + slot);
          const int col = c * kQChunk + lane;
          if (col < S) {
            dd a;
            a.hi = __ldcg(acc_hi + col);
            a.lo = __ldcg(acc_lo + col);
            a = dd_add_d(a, v);
            __stcg(acc_hi + col, a.hi);
            __stcg(acc_lo + col, a.lo);
          }
        }
      }
    }
  } else {
    // ================================ epilogue groups ================================
    reg_alloc<kQRegsEpi>();
    BC_Q_BOUNDS
    BC_Q_TMEM_BASE
    const int grp = warp >> 2;   // 0..3
    const int buf = grp >> 1;    // accumulator buffer = chunk parity this group serves
    const int half = grp & 1;    // which 16 of the chunk's 32 columns
    const int q = warp & 3;      // TMEM lane quarter this warp may read
    const int row = q * 32 + lane;
    const uint32_t tlane = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * kQDiagCols + half * kQHalfCols);
    const double dS = (double)S;
    const double rsum = (MODE == QMODE_SCORE) ? __ldg(P.resid + S) : 0.0;
    // one scale for all samples; the stored scales are those of the full 7-digit split: 2^-32 each, 256^-(NS+1) is needed
    const double cs = __ldg(P.colscale) * (double)(1ull << (8 * (kQSlices - NS)));
    const double kNaN = __longlong_as_double(0x7ff8000000000000LL);
    const typename F::Tabs tabs = F::tabs(P.pot_tabs, lane);   // lane-table forms: this lane's entries (bc_fastmath.cuh)
    constexpr int NB = kQHalfCols / 4;     // batches of 4 columns per chunk
    Best best = {0.0, -1};
    uint32_t use = 0, tcount = 0;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++tcount) {
      const long long p = tile * kQTileRows + row;
      const bool rv = p < n;
      const double rs = rv ? __ldg(P.rowscale + p) : 0.0;
      const double rsc = rs * cs;          // NaN for a non-finite row / sample set
      const double ra = (F::kRowAux && rv) ? __ldg(P.rowaux + p) : 0.0;
      const bool tail_rows = (tile + 1) * kQTileRows > n;   // warp-uniform
      double piv = 0.0;
      double s1 = 0.0, s2 = 0.0, sr = 0.0;
      bool have_piv = (MODE == QMODE_DOT) || grp == 0;
      for (int c = buf; c < nchunks; c += 2) {
        const uint32_t itb = tcount * (uint32_t)nchunks + (uint32_t)c;
        const bool need_mask = tail_rows || (c + 1) * kQChunk > S;   // warp-uniform: last row tile / ragged last chunk
        mbar_wait(tmem_full + buf, use & 1);
        ++use;
        tc_fence_after();
        int dg[4][NS];   // [column of the batch][diagonal]
        auto fetch = [&](int batch) {
          // this warp's lane quarter, this group's half of the chunk, inside the buffer the MMA warp has just filled
          BC_DASSERT(((tlane >> 16) & 0x7fu) == (uint32_t)(q * 32) + ((tmem_base >> 16) & 0x7fu));
          BC_DASSERT((tlane & 0xffffu) + (uint32_t)((NS - 1) * kQChunk + batch * 4 + 4) <= (tmem_base & 0xffffu) + (uint32_t)((buf + 1) * kQDiagCols));
#pragma unroll
          for (int d = 0; d < NS; ++d) {
            uint32_t v[4];
            tmem_ld_x4(tlane + (uint32_t)(d * kQChunk + batch * 4), v);
#pragma unroll
            for (int e = 0; e < 4; ++e) dg[e][d] = (int)v[e];
          }
        };
        // tcgen05.ld is asynchronous: its destination registers are defined only after tcgen05.wait::ld.  The empty asm
        // statements pin every use of the digits behind the wait (volatile asm statements keep their order).
        auto landed = [&]() {
          tmem_wait_ld();
#pragma unroll
          for (int e = 0; e < 4; ++e) {
#pragma unroll
            for (int d = 0; d < NS; ++d) asm volatile("" : "+r"(dg[e][d]));
          }
        };
        fetch(0);
        if (!have_piv) {
          // group 0 publishes the row pivots right after its first four columns of the tile
          mbar_wait(piv_full + (tcount & 1), (tcount >> 1) & 1);
          piv = pivs[(tcount & 1) * 128 + row];
          have_piv = true;
        }
        const uint32_t slot = itb % kQSlots, rph = (itb / kQSlots) & 1;
        double fv[8];
        // software pipeline inside the warp: while batch b is evaluated (FP64 pipe) the digits of batch b+1 -- fetched one
        // step earlier -- are recombined (integer pipe) in the same straight-line block, and the fetch of batch b+2 is
        // issued right behind the evaluation.  (With the 112 registers of the epilogue warps the fetch can also be issued
        // BEFORE the evaluation without spilling -- BC_Q_EARLY_FETCH -- but that measured 1 % slower.)
        double ccur[4], cnext[4];
        landed();
#pragma unroll
        for (int e = 0; e < 4; ++e) ccur[e] = rsc * q_combine_n<NS>(dg[e]);
        fetch(1);
#pragma unroll
        for (int b = 0; b < NB; ++b) {
          const int cb = c * kQChunk + half * kQHalfCols + b * 4;
          double cval[4], ca[4], fr[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) cval[e] = ccur[e];
          if (b + 1 < NB) {
            landed();   // digits of batch b+1
#pragma unroll
            for (int e = 0; e < 4; ++e) cnext[e] = rsc * q_combine_n<NS>(dg[e]);
            if (b + 2 < NB) {
#ifdef BC_Q_EARLY_FETCH   // measured (profiles/r02_q_tiers.txt): issuing the fetch here, before the evaluation, is no faster
              fetch(b + 2);
#endif
            } else {
              // the last digits of the chunk are in registers: hand the accumulator buffer back to the MMA issuer now,
              // with two batches of evaluation still to go -- the next chunk's MMAs then finish before this group needs them
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(tmem_empty + buf);
            }
          }
          if (MODE == QMODE_DOT) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              if (rv && cb + e < S) P.V[p * P.ldv + cb + e] = cval[e];
              fv[(b & 1) * 4 + e] = 0.0;
            }
          } else {
#pragma unroll
            for (int e = 0; e < 4; ++e) ca[e] = F::kColAux ? __ldg(P.colaux + min(cb + e, S - 1)) : 0.0;
            F::template evalv<4>(cval, ra, ca, P.mp, tabs, fr);
            if (grp == 0 && c == 0 && b == 0) {
              // pivot = the potential at the first sample: any per-row constant near the row mean keeps
              // sum f^2 - S mean^2 well conditioned; the other groups read it from shared memory.
              // A non-finite row (or sample set) has rsc = NaN: its pivot, hence every centred value of the row, is NaN
              // -- the reference's NaN propagation (evalv itself does not promise NaN in -> NaN out).
              piv = (rsc != rsc) ? kNaN : fr[0];
              pivs[(tcount & 1) * 128 + row] = piv;
              __syncwarp();
              if (lane == 0) mbar_arrive(piv_full + (tcount & 1));
            }
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              // __dsub_rn: never contracted -- a row whose potential is the same double for every sample must centre
              // to exactly 0 (-> 0/0 = NaN score, as in the reference, bcores.py:78)
              fr[e] = __dsub_rn(fr[e], piv);
            }
            if (need_mask) {
#pragma unroll
              for (int e = 0; e < 4; ++e) fr[e] = (rv && cb + e < S) ? fr[e] : 0.0;
            }
            if (MODE == QMODE_SCORE) {
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const double rr = __ldg(P.resid + min(cb + e, S - 1));
                s1 += fr[e];
                s2 = fma(fr[e], fr[e], s2);
                sr = fma(fr[e], rr, sr);
              }
            }
#pragma unroll
            for (int e = 0; e < 4; ++e) fv[(b & 1) * 4 + e] = fr[e];
          }
#ifndef BC_Q_EARLY_FETCH
          if (b + 2 < NB) fetch(b + 2);
#endif
#pragma unroll
          for (int e = 0; e < 4; ++e) ccur[e] = cnext[e];
          if (want_cols && (b & 1)) {
            // transposed butterfly: 32 rows x 8 columns -> lanes with (lane & 3) == 0 own one column total each
            int off = 0;
#pragma unroll
            for (int m = 16, hw = 4; m >= 4; m >>= 1, hw >>= 1) {
              const bool up = (lane & m) != 0;
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                if (i < hw) {
                  const double send = up ? fv[i] : fv[hw + i];
                  const double keep = up ? fv[hw + i] : fv[i];
                  fv[i] = keep + __shfl_xor_sync(0xffffffffu, send, m);
                }
              }
              if (up) off += hw;
            }
            fv[0] += __shfl_xor_sync(0xffffffffu, fv[0], 2);
            fv[0] += __shfl_xor_sync(0xffffffffu, fv[0], 1);
            if (b == 1) mbar_wait(ring_empty + slot, rph ^ 1);
            BC_DASSERT(slot < (uint32_t)kQSlots && off >= 0 && off < 8);
            if ((lane & 3) == 0) ring[(size_t)slot * 4 * kQChunk + q * kQChunk + half * kQHalfCols + (b >> 1) * 8 + off] = fv[0];
          }
        }
        if (want_cols) {
          __syncwarp();
          if (lane == 0) mbar_arrive(ring_full + slot);
        }
      }
      if (MODE == QMODE_SCORE) {
        if (grp != 0) {
          mbar_wait(stats_empty, (tcount & 1) ^ 1);
          double* st = stats + ((size_t)(grp - 1) * 128 + row) * 3;
          st[0] = s1;
          st[1] = s2;
          st[2] = sr;
          __syncwarp();
          if (lane == 0) mbar_arrive(stats_full);
        } else {
          mbar_wait(stats_full, tcount & 1);
#pragma unroll
          for (int g = 0; g < kQGroups - 1; ++g) {  // fixed order
            const double* st = stats + ((size_t)g * 128 + row) * 3;
            s1 += st[0];
            s2 += st[1];
            sr += st[2];
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(stats_empty);
          // centred quantities: v = g - gbar;  v.r = SR - gbar*sum(r);  |v|^2 = S2 - S*gbar^2
          const double gb = s1 / dS;
          double nrm2 = s2 - dS * gb * gb;
          if (nrm2 < 0.0) nrm2 = 0.0;
          const double dot = sr - gb * rsum;
          // bcores.py:78  corrs = vecs.dot(resid) / sqrt((vecs**2).sum(1)) / S
          double score = dot / sqrt(nrm2) / dS;
          // every sample gave exactly the pivot value: a constant row -- centred the way numpy rounds it (bc_common.cuh)
          if (s2 == 0.0) score = np_score_const(piv, S, rsum);
          if (rv) {
            if (P.scores) P.scores[p] = score;
            Best mine = {score, P.idx_offset + p};
            best = best_merge(best, mine);
          }
        }
      }
    }
    if (MODE == QMODE_SCORE && grp == 0) {
      best = best_warp(best);
      if (lane == 0) {
        fin[q * 2 + 0] = best.v;
        fin[q * 2 + 1] = __longlong_as_double(best.i);
      }
      named_bar_sync(1, 128);
      if (tid == 0) {
        Best b = {fin[0], __double_as_longlong(fin[1])};
        for (int w = 1; w < 4; ++w) {
          Best ob = {fin[w * 2 + 0], __double_as_longlong(fin[w * 2 + 1])};
          b = best_merge(b, ob);
        }
        P.part_misc[blockIdx.x * 4 + 2] = b.v;
        P.part_misc[blockIdx.x * 4 + 3] = __longlong_as_double(b.i);
      }
    }
  }

  // ---- teardown: every tcgen05 operation of this CTA is complete once all roles have left their loops ----
  tc_fence_before();
  __syncthreads();
  if (warp == kQEpiWarps + 1) {
    tc_fence_after();
    BC_Q_TMEM_BASE
    tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------ launch --
template <class F, int MODE, int NS>
static cudaError_t launch_q_one(const QProjArgs& P, int grid, cudaStream_t st) {
  auto kern = k_project_q<F, MODE, NS>;
  static DeviceOnce once;  // per instantiation, per device
  cudaError_t e = raise_dynamic_smem(kern, QSmem::total, once);
  if (e != cudaSuccess) return e;
  kern<<<grid, kQThreads, QSmem::total, st>>>(P);
  return cudaGetLastError();
}

template <class F, int MODE>
static cudaError_t launch_q_digits(const QProjArgs& P, int digits, int grid, cudaStream_t st) {
  switch (digits) {
    case 5: return launch_q_one<F, MODE, 5>(P, grid, st);
    case 6: return launch_q_one<F, MODE, 6>(P, grid, st);
    case 7: return launch_q_one<F, MODE, 7>(P, grid, st);
    default: return cudaErrorInvalidValue;
  }
}

template <class F>
static cudaError_t launch_q_mode(const QProjArgs& P, int mode, int digits, int grid, cudaStream_t st) {
  if (mode == QMODE_COLSUM) return launch_q_digits<F, QMODE_COLSUM>(P, digits, grid, st);
  return launch_q_digits<F, QMODE_SCORE>(P, digits, grid, st);
}

cudaError_t launch_project_q(const QProjArgs& P, int model, int kind, int poly, int mode, int digits, int grid, cudaStream_t st) {
  if (mode == QMODE_DOT) return launch_q_digits<LogisticF<KIND_LOGLIK, 0>, QMODE_DOT>(P, digits, grid, st);
  if (model == MODEL_LOGISTIC) {
    if (kind == KIND_LOGLIK && poly == kPowTab) return launch_q_mode<LogisticF<KIND_LOGLIK, kPowTab>>(P, mode, digits, grid, st);
    if (kind == KIND_LOGLIK) return launch_q_mode<LogisticF<KIND_LOGLIK, 0>>(P, mode, digits, grid, st);
    if (poly == kPowTab) return launch_q_mode<LogisticF<KIND_BETALIK, kPowTab>>(P, mode, digits, grid, st);
    if (poly == 20) return launch_q_mode<LogisticF<KIND_BETALIK, 20>>(P, mode, digits, grid, st);
    if (poly == kPowPolyMax) return launch_q_mode<LogisticF<KIND_BETALIK, kPowPolyMax>>(P, mode, digits, grid, st);
    return launch_q_mode<LogisticF<KIND_BETALIK, 0>>(P, mode, digits, grid, st);
  } else if (model == MODEL_GAUSSIAN) {
    if (kind == KIND_LOGLIK) return launch_q_mode<GaussianF<KIND_LOGLIK>>(P, mode, digits, grid, st);
    if (kind == KIND_BETALIK) return launch_q_mode<GaussianF<KIND_BETALIK>>(P, mode, digits, grid, st);
    return launch_q_mode<GaussianF<KIND_BETAGRAD>>(P, mode, digits, grid, st);
  } else {
    if (kind == KIND_LOGLIK) return launch_q_mode<NeurlinF<KIND_LOGLIK>>(P, mode, digits, grid, st);
    return launch_q_mode<NeurlinF<KIND_BETALIK>>(P, mode, digits, grid, st);
  }
}

}  // namespace bc
