// beta-cores B200: stage 1 (+2) -- the fused N x S projection.
//
// Replaces, without ever materialising the N x S matrix (unless asked to):
//   BetaBlackBoxProjector.project_f / BlackBoxProjector.project   (bayesiancoresets/coreset/projector.py:51-55, :23-26)
//   the column sum / residual correlation of BetaCoreset._select   (bayesiancoresets/coreset/bcores.py:77-81)
//   the column sum of BetaCoreset._optimize's grd()                (bayesiancoresets/coreset/bcores.py:142-146)
//
// One persistent CTA per SM, warp-specialised:
//   * CW = BM/16 compute warps.  Each owns 16 data rows of the CTA's BM-row tile and sweeps all S samples in chunks
//     of BN = 32: FP64 tensor-core MMAs (DMMA m16n8k4) for the contraction, then the model's potential on the
//     accumulator registers (FP64 pipe), minus a per-row pivot (the potential at the mean sample: keeps
//     sum f^2 - S mean^2 well conditioned).  Rows are warp-private, so the per-row statistics (sum f, sum f^2,
//     sum f r -> centred norm and correlation -> arg-max) never leave registers, and the compute warps never meet
//     at a CTA-wide barrier: while one warp of a scheduler is in its DMMA phase the other is in its FP64 epilogue,
//     which is what keeps both pipes busy.
//   * one producer warp: bulk copies through the TMA engine (cp.async.bulk, gather-capable) of the row tile (once
//     per tile) and of the sample chunks (NST-stage mbarrier ring).
//   * one reducer warp: the compute warps hand their 16-row column partials over through a small shared-memory
//     ring; the reducer adds the CW partials of a chunk in fixed order and accumulates them per column in
//     double-double (so the S-vector does not depend on the tile -> CTA -> GPU assignment).
// MODE_MATERIALISE additionally writes the centred rows (Hilbert / coreset-point path).
#include "bc_common.cuh"
#include "bc_models.cuh"
#include "bc_kernels.h"

namespace bc {

constexpr int kBN = 32;       // sample columns per chunk (4 DMMA column tiles per warp)
constexpr int kNT = kBN / 8;  // 8-column MMA tiles per warp
constexpr int kNV = 2 * kNT;  // column values per thread per chunk
constexpr int kSlots = 4;     // column-partial ring depth

struct SmemLayout {
  size_t a, b, ring, bbar, fin, bars, total;
};
__host__ __device__ inline SmemLayout smem_layout(int BM, int NST, int ss, int Dpad) {
  SmemLayout L;
  size_t o = 0;
  L.a = o; o += (size_t)BM * ss * 8;
  L.b = o; o += (size_t)NST * kBN * ss * 8;
  L.ring = o; o += (size_t)kSlots * (BM / 16) * kBN * 8;
  L.bbar = o; o += (size_t)Dpad * 8;
  L.fin = o; o += (size_t)(BM / 16) * 2 * 8;
  L.bars = o; o += (size_t)(2 + 2 * NST + 2 * kSlots) * 8;
  L.total = o;
  return L;
}

template <class F, int MODE, int BM, int NST>
__global__ void __launch_bounds__((BM / 16 + 2) * 32, 1) k_project(const ProjArgs P) {
  constexpr int CW = BM / 16;
  constexpr int BN = kBN, NT = kNT, NV = kNV;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int ss = P.ss, Dpad = P.Dpad, S = P.S;
  const SmemLayout L = smem_layout(BM, NST, ss, Dpad);
  double* As = reinterpret_cast<double*>(smem_raw + L.a);
  double* Bs = reinterpret_cast<double*>(smem_raw + L.b);
  double* ring = reinterpret_cast<double*>(smem_raw + L.ring);
  double* bbar_s = reinterpret_cast<double*>(smem_raw + L.bbar);
  double* fin = reinterpret_cast<double*>(smem_raw + L.fin);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + L.bars);
  uint64_t* full_a = bars + 0;
  uint64_t* empty_a = bars + 1;
  uint64_t* full_b = bars + 2;                // [NST]
  uint64_t* empty_b = full_b + NST;           // [NST]
  uint64_t* ring_full = empty_b + NST;        // [kSlots]
  uint64_t* ring_empty = ring_full + kSlots;  // [kSlots]

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long long n = P.n;
  const long long ntiles = (n + BM - 1) / BM;
  const int nchunks = (S + BN - 1) / BN;
  // sample-range split (raw materialise of few rows only): blockIdx.y owns the chunks [c_lo, c_hi) of every tile
  const int c_lo = (MODE == MODE_MATERIALISE) ? (int)(((long long)nchunks * blockIdx.y) / gridDim.y) : 0;
  const int c_hi = (MODE == MODE_MATERIALISE) ? (int)(((long long)nchunks * (blockIdx.y + 1)) / gridDim.y) : nchunks;
  const uint32_t a_row_bytes = (uint32_t)P.Dc * 8u;
  const uint32_t b_row_bytes = (uint32_t)Dpad * 8u;
  const bool want_cols = (MODE == MODE_COLSUM) || (MODE == MODE_MATERIALISE && P.want_colsum);

  if (tid == 0) {
    mbar_init(full_a, 1);
    mbar_init(empty_a, CW);
    for (int i = 0; i < NST; ++i) {
      mbar_init(full_b + i, 1);
      mbar_init(empty_b + i, CW);
    }
    for (int i = 0; i < kSlots; ++i) {
      mbar_init(ring_full + i, CW);
      mbar_init(ring_empty + i, 1);
    }
    mbar_fence_init();
  }
  // columns [Dc, ss) of the A tile are never written by the bulk copies: zero them once
  // (the k-loop reads up to Dpad).  The B operand is zero beyond Dk in global memory already.
  for (int i = tid; i < BM * (ss - P.Dc); i += blockDim.x) {
    int r = i / (ss - P.Dc), c = P.Dc + i % (ss - P.Dc);
    As[r * ss + c] = 0.0;
  }
  for (int i = tid; i < Dpad; i += blockDim.x) bbar_s[i] = P.bbar[i];
  __syncthreads();

  if (warp == CW) {
    // ======================= producer warp: TMA-engine bulk copies =======================
    uint32_t itb = 0, tcount = 0;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++tcount) {
      mbar_wait(empty_a, (tcount & 1) ^ 1);
      if (lane == 0) mbar_arrive_expect_tx(full_a, (uint32_t)BM * a_row_bytes);
      __syncwarp();
      for (int r = lane; r < BM; r += 32) {
        long long p = tile * BM + r;
        if (p > n - 1) p = n - 1;  // tail rows: duplicate a valid row, masked by the consumers
        const long long rid = P.rows ? P.rows[p] : p;
        bulk_g2s(As + (size_t)r * ss, P.A + rid * P.lda, a_row_bytes, full_a);
      }
      for (int c = c_lo; c < c_hi; ++c, ++itb) {
        const uint32_t st = itb % NST, ph = (itb / NST) & 1;
        mbar_wait(empty_b + st, ph ^ 1);
        if (lane == 0) mbar_arrive_expect_tx(full_b + st, (uint32_t)BN * b_row_bytes);
        __syncwarp();
        double* dst = Bs + (size_t)st * BN * ss;
        int s = c * BN + lane;
        if (s > S - 1) s = S - 1;  // tail columns: duplicate, masked by the consumers
        bulk_g2s(dst + (size_t)lane * ss, P.B + (size_t)s * P.ldb, b_row_bytes, full_b + st);
      }
    }
    return;
  }

  if (warp == CW + 1) {
    // ============ reducer warp: CW partials per chunk -> double-double column accumulators ============
    if (!want_cols) return;
    double* acc_hi = P.part_colsum + (size_t)blockIdx.x * 2 * P.Sld;
    double* acc_lo = acc_hi + P.Sld;
    for (int i = lane; i < 2 * P.Sld; i += 32) __stcg(acc_hi + i, 0.0);
    __syncwarp();
    uint32_t itb = 0;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      for (int c = c_lo; c < c_hi; ++c, ++itb) {
        const uint32_t slot = itb % kSlots, ph = (itb / kSlots) & 1;
        mbar_wait(ring_full + slot, ph);
        const double* rs = ring + (size_t)slot * CW * BN + lane;
        double v = rs[0];
#pragma unroll
        for (int w = 1; w < CW; ++w) v += rs[w * BN];
        __syncwarp();
        if (lane == 0) mbar_arrive(ring_empty + slot);
        const int col = c * BN + lane;
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
    return;
  }

  // ================================ compute warps ================================
  const int g = lane >> 2, t = lane & 3;
  const int wrow = warp * 16;
  const int r0 = wrow + g, r1 = r0 + 8;
  const double invS = 1.0 / (double)S;
  const double dS = (double)S;
  Best best = {0.0, -1};
  const double rsum = (MODE == MODE_SCORE) ? __ldg(P.resid + S) : 0.0;
  const double cabar = __ldg(P.bbar + Dpad);  // mean per-sample aux term (pivot)

  uint32_t itb = 0, tcount = 0;
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++tcount) {
    mbar_wait(full_a, tcount & 1);
    // ---- per-row pivot = potential at the mean sample (lanes 2r, 2r+1 hold row wrow + r) ----
    double piv0, piv1, ra0, ra1;
    {
      const int row = lane >> 1, half = lane & 1;
      const double* ar = As + (size_t)(wrow + row) * ss;
      double acc = 0.0;
      for (int k = 2 * half; k < Dpad; k += 4) {
        acc = fma(ar[k], bbar_s[k], acc);
        acc = fma(ar[k + 1], bbar_s[k + 1], acc);
      }
      acc += __shfl_xor_sync(0xffffffffu, acc, 1);
      long long p = tile * BM + wrow + row;
      if (p > n - 1) p = n - 1;
      double ra = 0.0;
      if (F::kRowAux) {
        if (P.rowaux) {
          const long long rid = P.rows ? P.rows[p] : p;
          ra = __ldg(P.rowaux + rid);
        } else {
          ra = ar[P.Dk];  // neural-linear: y is column Dk of the staged row
        }
      }
      const double piv = (MODE == MODE_MATERIALISE && P.raw) ? 0.0 : F::eval(acc, ra, cabar, P.mp);
      piv0 = __shfl_sync(0xffffffffu, piv, 2 * g);
      piv1 = __shfl_sync(0xffffffffu, piv, 2 * g + 16);
      ra0 = __shfl_sync(0xffffffffu, ra, 2 * g);
      ra1 = __shfl_sync(0xffffffffu, ra, 2 * g + 16);
    }
    const long long p0 = tile * BM + r0, p1 = tile * BM + r1;
    const bool v0 = p0 < n, v1 = p1 < n;
    double s1_0 = 0, s1_1 = 0, s2_0 = 0, s2_1 = 0, sr_0 = 0, sr_1 = 0;

    for (int c = c_lo; c < c_hi; ++c, ++itb) {
      const uint32_t st = itb % NST, ph = (itb / NST) & 1;
      double acc[NT][4];
#pragma unroll
      for (int j = 0; j < NT; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.0;
      mbar_wait(full_b + st, ph);
      {
        const double* a0p = As + (size_t)r0 * ss + t;
        const double* a1p = As + (size_t)r1 * ss + t;
        const double* bp = Bs + (size_t)st * BN * ss + (size_t)g * ss + t;
#pragma unroll 4
        for (int k = 0; k < Dpad; k += 4) {
          const double a0 = a0p[k], a1 = a1p[k];
#pragma unroll
          for (int j = 0; j < NT; ++j) dmma_m16n8k4(acc[j], a0, a1, bp[(size_t)j * 8 * ss + k]);
        }
      }
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(empty_b + st);
        if (c == c_hi - 1) mbar_arrive(empty_a);
      }

      // ---- epilogue: potential, pivot shift, row / column partial sums ----
      double colv[NV];
      const int cb = c * BN + 2 * t;
#pragma unroll
      for (int j = 0; j < NT; ++j) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int col = cb + j * 8 + e;
          const bool cv = col < S;
          const int colc = cv ? col : S - 1;
          const double ca = F::kColAux ? __ldg(P.colaux + colc) : 0.0;
          // __dsub_rn: never contracted into an FMA with the potential's last multiply -- a row whose
          // potential is the same double for every sample must centre to exactly 0 (-> 0/0 = NaN score,
          // as in the reference, bcores.py:78)
          double f0 = __dsub_rn(F::eval(acc[j][e], ra0, ca, P.mp), piv0);
          double f1 = __dsub_rn(F::eval(acc[j][2 + e], ra1, ca, P.mp), piv1);
          f0 = (v0 && cv) ? f0 : 0.0;
          f1 = (v1 && cv) ? f1 : 0.0;
          s1_0 += f0;
          s1_1 += f1;
          if (MODE == MODE_SCORE) {
            const double rr = __ldg(P.resid + colc);
            s2_0 = fma(f0, f0, s2_0);
            s2_1 = fma(f1, f1, s2_1);
            sr_0 = fma(f0, rr, sr_0);
            sr_1 = fma(f1, rr, sr_1);
          }
          if (MODE == MODE_MATERIALISE) {
            if (v0 && cv) P.V[p0 * P.ldv + col] = f0;
            if (v1 && cv) P.V[p1 * P.ldv + col] = f1;
          }
          colv[j * 2 + e] = f0 + f1;
        }
      }
      if (want_cols) {
        // transposed butterfly over the 8 row-groups (lane bits 4,3,2): 16 rows -> one lane per column
        int off = 0;
#pragma unroll
        for (int m = 16, half = NV / 2; m >= 4; m >>= 1, half >>= 1) {
          const bool up = (lane & m) != 0;
#pragma unroll
          for (int i = 0; i < NV / 2; ++i) {
            if (i < half) {
              const double send = up ? colv[i] : colv[half + i];
              const double keep = up ? colv[half + i] : colv[i];
              colv[i] = keep + __shfl_xor_sync(0xffffffffu, send, m);
            }
          }
          if (up) off += half;
        }
        // off = j*2 + e of the single column this lane now owns
        const uint32_t slot = itb % kSlots, rph = (itb / kSlots) & 1;
        mbar_wait(ring_empty + slot, rph ^ 1);
        ring[(size_t)slot * CW * BN + warp * BN + (off >> 1) * 8 + 2 * t + (off & 1)] = colv[0];
        __syncwarp();
        if (lane == 0) mbar_arrive(ring_full + slot);
      }
    }

    // ---- tile end: per-row statistics (the four lanes of a quad share rows r0, r1) ----
    s1_0 += __shfl_xor_sync(0xffffffffu, s1_0, 1);
    s1_0 += __shfl_xor_sync(0xffffffffu, s1_0, 2);
    s1_1 += __shfl_xor_sync(0xffffffffu, s1_1, 1);
    s1_1 += __shfl_xor_sync(0xffffffffu, s1_1, 2);
    const double gb0 = s1_0 * invS, gb1 = s1_1 * invS;
    if (MODE == MODE_SCORE) {
      s2_0 += __shfl_xor_sync(0xffffffffu, s2_0, 1);
      s2_0 += __shfl_xor_sync(0xffffffffu, s2_0, 2);
      s2_1 += __shfl_xor_sync(0xffffffffu, s2_1, 1);
      s2_1 += __shfl_xor_sync(0xffffffffu, s2_1, 2);
      sr_0 += __shfl_xor_sync(0xffffffffu, sr_0, 1);
      sr_0 += __shfl_xor_sync(0xffffffffu, sr_0, 2);
      sr_1 += __shfl_xor_sync(0xffffffffu, sr_1, 1);
      sr_1 += __shfl_xor_sync(0xffffffffu, sr_1, 2);
      if (t < 2) {  // lane t = 0 scores row r0, lane t = 1 scores row r1
        const double gb = t ? gb1 : gb0, S2 = t ? s2_1 : s2_0, SR = t ? sr_1 : sr_0;
        const long long p = t ? p1 : p0;
        // centred quantities: v = g - gbar;  v.r = SR - gbar*sum(r);  |v|^2 = S2 - S*gbar^2
        double nrm2 = S2 - dS * gb * gb;
        if (nrm2 < 0.0) nrm2 = 0.0;
        const double dot = SR - gb * rsum;
        // bcores.py:78  corrs = vecs.dot(resid) / sqrt((vecs**2).sum(1)) / S
        double score = dot / sqrt(nrm2) / dS;
        // every sample gave exactly the pivot value: a constant row -- centred the way numpy rounds it (bc_common.cuh)
        if (S2 == 0.0) score = np_score_const(t ? piv1 : piv0, S, rsum);
        if (p < n) {
          if (P.scores) P.scores[p] = score;
          Best mine = {score, P.idx_offset + p};
          best = best_merge(best, mine);
        }
      }
    }
    if (MODE == MODE_MATERIALISE && !P.raw) {
      // second pass over this thread's own output: subtract the row mean, exact two-pass norm
      double q0 = 0.0, q1 = 0.0;
      for (int c = 0; c < nchunks; ++c) {
        const int cb = c * BN + 2 * t;
#pragma unroll
        for (int j = 0; j < NT; ++j) {
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int col = cb + j * 8 + e;
            if (col < S) {
              if (v0) {
                const double x = P.V[p0 * P.ldv + col] - gb0;
                P.V[p0 * P.ldv + col] = x;
                q0 = fma(x, x, q0);
              }
              if (v1) {
                const double x = P.V[p1 * P.ldv + col] - gb1;
                P.V[p1 * P.ldv + col] = x;
                q1 = fma(x, x, q1);
              }
            }
          }
        }
      }
      q0 += __shfl_xor_sync(0xffffffffu, q0, 1);
      q0 += __shfl_xor_sync(0xffffffffu, q0, 2);
      q1 += __shfl_xor_sync(0xffffffffu, q1, 1);
      q1 += __shfl_xor_sync(0xffffffffu, q1, 2);
      // constant rows (every centred value exactly 0): the reference leaves numpy's rounding residue in every column
      // (usually 0, else one ulp) -- reproduced, since it decides zero-norm filters and NaN correlations downstream
      if (q0 == 0.0 && v0) {
        const double v = np_centred_const(piv0, S);
        if (v != 0.0) {
          for (int col = 2 * t; col < S; col += 8) {
            P.V[p0 * P.ldv + col] = v;
            if (col + 1 < S) P.V[p0 * P.ldv + col + 1] = v;
          }
          q0 = np_sum_const(v * v, S);
        }
      }
      if (q1 == 0.0 && v1) {
        const double v = np_centred_const(piv1, S);
        if (v != 0.0) {
          for (int col = 2 * t; col < S; col += 8) {
            P.V[p1 * P.ldv + col] = v;
            if (col + 1 < S) P.V[p1 * P.ldv + col + 1] = v;
          }
          q1 = np_sum_const(v * v, S);
        }
      }
      if (P.norms && t == 0) {
        if (v0) P.norms[p0] = sqrt(q0);
        if (v1) P.norms[p1] = sqrt(q1);
      }
    }
  }

  // ---- CTA result: the local arg-max ----
  if (MODE == MODE_SCORE) {
    best = best_warp(best);
    if (lane == 0) {
      fin[warp * 2 + 0] = best.v;
      fin[warp * 2 + 1] = __longlong_as_double(best.i);
    }
    named_bar_sync(1, CW * 32);
    if (tid == 0) {
      Best b = {fin[0], __double_as_longlong(fin[1])};
      for (int w = 1; w < CW; ++w) {
        Best ob = {fin[w * 2 + 0], __double_as_longlong(fin[w * 2 + 1])};
        b = best_merge(b, ob);
      }
      P.part_misc[blockIdx.x * 4 + 2] = b.v;
      P.part_misc[blockIdx.x * 4 + 3] = __longlong_as_double(b.i);
    }
  }
}

// Reduce the per-CTA partials in fixed order (one CTA).
//   out_dd : [2][Sld]  (hi plane, lo plane); column s = sum over rows of (f - pivot), element S = the sum over rows of
//            the row means = (sum over columns)/S, so that out[s] - out[S] is the column sum of the CENTRED matrix
//   out_best: {score, index-as-double-bits}
__global__ void __launch_bounds__(1024) k_project_finalize(const double* __restrict__ part_colsum, const double* __restrict__ part_misc,
                                                            int nctas, int S, int Sld, double* __restrict__ out_dd,
                                                            double* __restrict__ out_best, int mode, double* __restrict__ colsum_out) {
  __shared__ double sh[1024], sl[1024];
  __shared__ long long si[32];
  const int tid = threadIdx.x;
  if (mode != MODE_SCORE) {
    dd tot = {0.0, 0.0};
    for (int col = tid; col < S; col += blockDim.x) {
      dd a = {0.0, 0.0};
      // the partials are added in CTA order (the order is part of the result); their loads do not depend on the running
      // sum, so eight of them are put in flight together -- one L2 round trip per partial made this kernel 53 us at 148 CTAs
      int c = 0;
      for (; c + 8 <= nctas; c += 8) {
        double hi[8], lo[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          hi[q] = __ldcg(part_colsum + (size_t)(c + q) * 2 * Sld + col);
          lo[q] = __ldcg(part_colsum + (size_t)(c + q) * 2 * Sld + Sld + col);
        }
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          dd o = {hi[q], lo[q]};
          a = dd_add(a, o);
        }
      }
      for (; c < nctas; ++c) {
        dd o = {__ldcg(part_colsum + (size_t)c * 2 * Sld + col), __ldcg(part_colsum + (size_t)c * 2 * Sld + Sld + col)};
        a = dd_add(a, o);
      }
      out_dd[col] = a.hi;
      out_dd[Sld + col] = a.lo;
      tot = dd_add(tot, a);
    }
    sh[tid] = tot.hi;
    sl[tid] = tot.lo;
    __syncthreads();
    for (int o = blockDim.x >> 1; o > 0; o >>= 1) {  // fixed tree
      if (tid < o) {
        dd a = {sh[tid], sl[tid]}, b = {sh[tid + o], sl[tid + o]};
        a = dd_add(a, b);
        sh[tid] = a.hi;
        sl[tid] = a.lo;
      }
      __syncthreads();
    }
    if (tid == 0) {
      // (hi + lo) / S in double-double
      const double d = (double)S;
      const double qh = sh[0] / d;
      const double r = fma(-qh, d, sh[0]);
      const double ql = (r + sl[0]) / d;
      const double h = qh + ql;
      out_dd[S] = h;
      out_dd[Sld + S] = ql - (h - qh);
    }
    if (colsum_out) {
      // single-part jobs: the centred column sum right here, operation for operation what k_colsum_combine computes from
      // one part (bc_greedy_opt_step; saves the launch)
      __syncthreads();
      const dd om = {-out_dd[S], -out_dd[Sld + S]};
      for (int col = tid; col < S; col += blockDim.x) {
        dd a = {0.0, 0.0}, m = {0.0, 0.0};
        const dd o = {out_dd[col], out_dd[Sld + col]};
        a = dd_add(a, o);
        m = dd_add(m, om);
        a = dd_add(a, m);
        colsum_out[col] = a.hi;
      }
    }
  } else {
    Best b = {0.0, -1};
    for (int c = tid; c < nctas; c += blockDim.x) {
      Best o = {part_misc[c * 4 + 2], __double_as_longlong(part_misc[c * 4 + 3])};
      b = best_merge(b, o);
    }
    b = best_warp(b);
    if ((tid & 31) == 0) {
      sh[tid >> 5] = b.v;
      si[tid >> 5] = b.i;
    }
    __syncthreads();
    if (tid == 0) {
      for (int w = 1; w < (int)(blockDim.x >> 5); ++w) {
        Best o = {sh[w], si[w]};
        b = best_merge(b, o);
      }
      out_best[0] = b.v;
      out_best[1] = __longlong_as_double(b.i);
    }
  }
}

// ------------------------------------------------------------------ launch --
template <class F, int MODE, int BM, int NST>
static cudaError_t launch_one(const ProjArgs& P, int grid, size_t smem, cudaStream_t st, int csplit = 1) {
  auto kern = k_project<F, MODE, BM, NST>;
  static DeviceOnce once;  // per instantiation, per device
  cudaError_t e = raise_dynamic_smem(kern, kMaxSmem, once);
  if (e != cudaSuccess) return e;
  {
    static DeviceOnce carve_once;   // (per instantiation) same shared-memory split as the kernels around it: bc_kernels.h
    prefer_max_shared(kern, carve_once);
  }
  kern<<<dim3(grid, MODE == MODE_MATERIALISE ? csplit : 1), (BM / 16 + 2) * 32, smem, st>>>(P);
  return cudaGetLastError();
}

// tile configurations, tried in order: {rows per CTA tile, sample-chunk stages}
static const int kCfgs[3][2] = {{128, 2}, {64, 2}, {32, 2}};

template <class F, int MODE>
static cudaError_t launch_tile(const ProjArgs& P, int tile_cfg, int grid, size_t smem, cudaStream_t st, int csplit) {
  if (tile_cfg == 0) return launch_one<F, MODE, 128, 2>(P, grid, smem, st, csplit);
  if (tile_cfg == 1) return launch_one<F, MODE, 64, 2>(P, grid, smem, st, csplit);
  return launch_one<F, MODE, 32, 2>(P, grid, smem, st, csplit);
}

template <class F>
static cudaError_t launch_mode(const ProjArgs& P, int mode, int tile_cfg, int grid, size_t smem, cudaStream_t st, int csplit) {
  switch (mode) {
    case MODE_COLSUM: return launch_tile<F, MODE_COLSUM>(P, tile_cfg, grid, smem, st, 1);
    case MODE_SCORE: return launch_tile<F, MODE_SCORE>(P, tile_cfg, grid, smem, st, 1);
    default: return launch_tile<F, MODE_MATERIALISE>(P, tile_cfg, grid, smem, st, csplit);
  }
}

int project_tile_config(int Dpad, int* BM, int* BN, int* ss_out, size_t* smem_out) {
  const int ss = (Dpad % 8 == 4) ? Dpad : Dpad + 4;  // row stride = 4 mod 8 doubles: conflict-free fragment loads
  for (int i = 0; i < 3; ++i) {
    SmemLayout L = smem_layout(kCfgs[i][0], kCfgs[i][1], ss, Dpad);
    if (L.total <= kMaxSmem) {
      *BM = kCfgs[i][0];
      *BN = kBN;
      *ss_out = ss;
      *smem_out = L.total;
      return i;
    }
  }
  return -1;
}

// Few rows (the M coreset points, every optimiser step): a smaller row tile spreads them over more SMs and does not pay for
// 128 rows when 16 are valid.  Returns the configuration index >= base_cfg to use for n rows, with its shared memory size.
int project_tile_config_for_rows(int base_cfg, long long n, int sms, int Dpad, size_t* smem_out) {
  const int ss = (Dpad % 8 == 4) ? Dpad : Dpad + 4;
  int cfg = base_cfg;
  while (cfg < 2 && (n + kCfgs[cfg][0] - 1) / kCfgs[cfg][0] < sms) ++cfg;
  *smem_out = smem_layout(kCfgs[cfg][0], kCfgs[cfg][1], ss, Dpad).total;
  return cfg;
}
int project_tile_rows(int cfg) { return kCfgs[cfg][0]; }

int project_chunks(int S) { return (S + kBN - 1) / kBN; }

cudaError_t launch_project(const ProjArgs& P, int model, int kind, int poly, int mode, int tile_cfg, int grid, size_t smem,
                           cudaStream_t st, int csplit) {
  if (model == MODEL_LOGISTIC) {
    if (kind == KIND_LOGLIK) return launch_mode<LogisticF<KIND_LOGLIK, 0>>(P, mode, tile_cfg, grid, smem, st, csplit);
    if (poly == 20) return launch_mode<LogisticF<KIND_BETALIK, 20>>(P, mode, tile_cfg, grid, smem, st, csplit);
    if (poly == kPowPolyMax) return launch_mode<LogisticF<KIND_BETALIK, kPowPolyMax>>(P, mode, tile_cfg, grid, smem, st, csplit);
    return launch_mode<LogisticF<KIND_BETALIK, 0>>(P, mode, tile_cfg, grid, smem, st, csplit);
  } else if (model == MODEL_GAUSSIAN) {
    if (kind == KIND_LOGLIK) return launch_mode<GaussianF<KIND_LOGLIK>>(P, mode, tile_cfg, grid, smem, st, csplit);
    if (kind == KIND_BETALIK) return launch_mode<GaussianF<KIND_BETALIK>>(P, mode, tile_cfg, grid, smem, st, csplit);
    return launch_mode<GaussianF<KIND_BETAGRAD>>(P, mode, tile_cfg, grid, smem, st, csplit);
  } else {
    if (kind == KIND_LOGLIK) return launch_mode<NeurlinF<KIND_LOGLIK>>(P, mode, tile_cfg, grid, smem, st, csplit);
    return launch_mode<NeurlinF<KIND_BETALIK>>(P, mode, tile_cfg, grid, smem, st, csplit);
  }
}

cudaError_t launch_project_finalize(const double* part_colsum, const double* part_misc, int nctas, int S, int Sld,
                                    double* out_dd, double* out_best, int mode, cudaStream_t st, double* colsum_out) {
  const int threads = (mode == MODE_SCORE) ? 256 : (S >= 1024 ? 1024 : (S > 256 ? 512 : 256));
  BC_PREFER_MAX_SHARED(k_project_finalize);
  k_project_finalize<<<1, threads, 0, st>>>(part_colsum, part_misc, nctas, S, Sld, out_dd, out_best, mode, colsum_out);
  return cudaGetLastError();
}

}  // namespace bc
