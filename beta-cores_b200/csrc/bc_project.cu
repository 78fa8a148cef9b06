// beta-cores B200: stage 1 (+2) -- the fused N x S projection.
//
// Replaces, without ever materialising the N x S matrix (unless asked to):
//   BetaBlackBoxProjector.project_f / BlackBoxProjector.project   (bayesiancoresets/coreset/projector.py:51-55, :23-26)
//   the column sum / residual correlation of BetaCoreset._select   (bayesiancoresets/coreset/bcores.py:77-81)
//   the column sum of BetaCoreset._optimize's grd()                (bayesiancoresets/coreset/bcores.py:142-146)
//
// One persistent CTA per SM.  A CTA owns BM data rows at a time (bulk-copied once from HBM into
// shared memory by the TMA engine, gather-capable) and streams the S prepared samples through a
// two-stage mbarrier ring in chunks of BN.  Eight compute warps contract with FP64 tensor-core
// MMAs (DMMA), apply the model's potential in registers, subtract a per-row pivot (the potential
// at the mean sample: keeps sum f^2 - S mean^2 well conditioned), and reduce
//   * per column: sum over rows            (-> residual; double-double across tiles so the S-vector
//                                             does not depend on the tile->CTA->GPU assignment)
//   * per row:    sum f, sum f^2, sum f r   (-> centred norm and correlation -> arg-max)
// MODE_MATERIALISE additionally writes the centred rows (Hilbert / coreset-point path).
#include "bc_common.cuh"
#include "bc_models.cuh"
#include "bc_kernels.h"

namespace bc {

template <int BM_, int BN_>
struct Tile {
  static constexpr int BM = BM_, BN = BN_;
  static constexpr int WARPS_M = BM / 16;
  static constexpr int WARPS_N = kComputeWarps / WARPS_M;
  static constexpr int NT = BN / (8 * WARPS_N);   // 8-column MMA tiles per warp
  static constexpr int NV = 2 * NT;               // column values per thread per chunk
  static constexpr int TPR = kComputeThreads / BM;  // threads per row in the pivot prologue
  static_assert(WARPS_M * WARPS_N == kComputeWarps, "warp grid");
  static_assert(NT >= 1, "tile too narrow");
};

struct SmemLayout {
  size_t a, b, stage_cols, stage_rows, pivot, raux, gbar, bbar, bars, total;
};
__host__ __device__ inline SmemLayout smem_layout(int BM, int BN, int ss, int Dpad) {
  SmemLayout L;
  size_t o = 0;
  L.a = o; o += (size_t)BM * ss * 8;
  L.b = o; o += (size_t)2 * BN * ss * 8;
  L.stage_cols = o; o += (size_t)2 * 4 * BN * 8;
  L.stage_rows = o; o += (size_t)4 * BM * 3 * 8;
  L.pivot = o; o += (size_t)BM * 8;
  L.raux = o; o += (size_t)BM * 8;
  L.gbar = o; o += (size_t)BM * 8;
  L.bbar = o; o += (size_t)Dpad * 8;
  L.bars = o; o += 8 * 8;
  L.total = o;
  return L;
}

template <class F, int MODE, class T>
__global__ void __launch_bounds__(kThreads, 1) k_project(const ProjArgs P) {
  constexpr int BM = T::BM, BN = T::BN, NT = T::NT, NV = T::NV;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int ss = P.ss, Dpad = P.Dpad, S = P.S;
  const SmemLayout L = smem_layout(BM, BN, ss, Dpad);
  double* As = reinterpret_cast<double*>(smem_raw + L.a);
  double* Bs = reinterpret_cast<double*>(smem_raw + L.b);
  double* stage_cols = reinterpret_cast<double*>(smem_raw + L.stage_cols);
  double* stage_rows = reinterpret_cast<double*>(smem_raw + L.stage_rows);
  double* pivot_s = reinterpret_cast<double*>(smem_raw + L.pivot);
  double* raux_s = reinterpret_cast<double*>(smem_raw + L.raux);
  double* gbar_s = reinterpret_cast<double*>(smem_raw + L.gbar);
  double* bbar_s = reinterpret_cast<double*>(smem_raw + L.bbar);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + L.bars);
  uint64_t* full_a = bars + 0;
  uint64_t* empty_a = bars + 1;
  uint64_t* full_b = bars + 2;   // [2]
  uint64_t* empty_b = bars + 4;  // [2]

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long long n = P.n;
  const long long ntiles = (n + BM - 1) / BM;
  const int nchunks = (S + BN - 1) / BN;
  const uint32_t a_row_bytes = (uint32_t)P.Dc * 8u;
  const uint32_t b_row_bytes = (uint32_t)Dpad * 8u;

  if (tid == 0) {
    mbar_init(full_a, 1);
    mbar_init(empty_a, kComputeWarps);
    mbar_init(full_b + 0, 1);
    mbar_init(full_b + 1, 1);
    mbar_init(empty_b + 0, kComputeWarps);
    mbar_init(empty_b + 1, kComputeWarps);
    mbar_fence_init();
  }
  // columns [Dc, ss) of the A tile are never written by the bulk copies: zero them once
  // (the k-loop reads up to Dpad).  The B operand is zero beyond Dk in global memory already.
  for (int i = tid; i < BM * (ss - P.Dc); i += kThreads) {
    int r = i / (ss - P.Dc), c = P.Dc + i % (ss - P.Dc);
    As[r * ss + c] = 0.0;
  }
  for (int i = tid; i < Dpad; i += kThreads) bbar_s[i] = P.bbar[i];
  __syncthreads();

  if (warp == kComputeWarps) {
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
      for (int c = 0; c < nchunks; ++c, ++itb) {
        const uint32_t st = itb & 1, ph = (itb >> 1) & 1;
        mbar_wait(empty_b + st, ph ^ 1);
        if (lane == 0) mbar_arrive_expect_tx(full_b + st, (uint32_t)BN * b_row_bytes);
        __syncwarp();
        double* dst = Bs + (size_t)st * BN * ss;
        for (int r = lane; r < BN; r += 32) {
          int s = c * BN + r;
          if (s > S - 1) s = S - 1;  // tail columns: duplicate, masked by the consumers
          bulk_g2s(dst + (size_t)r * ss, P.B + (size_t)s * P.ldb, b_row_bytes, full_b + st);
        }
      }
    }
    return;
  }

  // ================================ compute warps ================================
  const int wm = warp % T::WARPS_M, wn = warp / T::WARPS_M;
  const int g = lane >> 2, t = lane & 3;
  const int r0 = wm * 16 + g, r1 = r0 + 8;
  const double invS = 1.0 / (double)S;
  const double dS = (double)S;

  // this CTA's double-double column accumulators live in its own slice of global memory
  double* acc_hi = P.part_colsum + (size_t)blockIdx.x * 2 * P.Sld;
  double* acc_lo = acc_hi + P.Sld;
  if (MODE != MODE_SCORE) {
    for (int i = tid; i < 2 * P.Sld; i += kComputeThreads) __stcg(acc_hi + i, 0.0);
  }
  dd sum_gbar = {0.0, 0.0};
  Best best = {0.0, -1};
  const double rsum = (MODE == MODE_SCORE) ? __ldg(P.resid + S) : 0.0;
  const double cabar = __ldg(P.bbar + Dpad);  // mean per-sample aux term (pivot)

  uint32_t itb = 0, tcount = 0;
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++tcount) {
    mbar_wait(full_a, tcount & 1);
    // ---- prologue: per-row pivot = potential at the mean sample ----
    {
      const int r = tid / T::TPR, q = tid % T::TPR;
      double acc = 0.0;
      for (int k = q; k < Dpad; k += T::TPR) acc = fma(As[r * ss + k], bbar_s[k], acc);
#pragma unroll
      for (int o = T::TPR / 2; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
      if (q == 0) {
        long long p = tile * BM + r;
        if (p > n - 1) p = n - 1;
        double ra = 0.0;
        if (F::kRowAux) {
          if (P.rowaux) {
            const long long rid = P.rows ? P.rows[p] : p;
            ra = __ldg(P.rowaux + rid);
          } else {
            ra = As[r * ss + P.Dk];  // neural-linear: y is column Dk of the staged row
          }
        }
        raux_s[r] = ra;
        pivot_s[r] = (MODE == MODE_MATERIALISE && P.raw) ? 0.0 : F::eval(acc, ra, cabar, P.mp);
      }
    }
    named_bar_sync(1, kComputeThreads);
    const long long p0 = tile * BM + r0, p1 = tile * BM + r1;
    const bool v0 = p0 < n, v1 = p1 < n;
    const double piv0 = pivot_s[r0], piv1 = pivot_s[r1];
    const double ra0 = raux_s[r0], ra1 = raux_s[r1];
    double s1_0 = 0, s1_1 = 0, s2_0 = 0, s2_1 = 0, sr_0 = 0, sr_1 = 0;

    for (int c = 0; c < nchunks; ++c, ++itb) {
      const uint32_t st = itb & 1, ph = (itb >> 1) & 1;
      double acc[NT][4];
#pragma unroll
      for (int j = 0; j < NT; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.0;
      mbar_wait(full_b + st, ph);
      {
        const double* a0p = As + (size_t)r0 * ss + t;
        const double* a1p = As + (size_t)r1 * ss + t;
        const double* bp = Bs + (size_t)st * BN * ss + (size_t)(wn * NT * 8 + g) * ss + t;
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
        if (c == nchunks - 1) mbar_arrive(empty_a);
      }

      // ---- epilogue: potential, pivot shift, row / column partial sums ----
      double colv[NV];
      const int cb = c * BN + wn * NT * 8 + 2 * t;
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
      if (MODE != MODE_SCORE) {
        // transposed butterfly over the 8 row-groups (lane bits 4,3,2): 16 rows -> one lane per column
        int cnt = NV, off = 0;
        bool writer = true;
#pragma unroll
        for (int m = 16; m >= 4; m >>= 1) {
          const bool up = (lane & m) != 0;
          if (cnt > 1) {
            const int half = cnt / 2;
#pragma unroll
            for (int i = 0; i < NV / 2; ++i) {
              if (i < half) {
                const double send = up ? colv[i] : colv[half + i];
                const double keep = up ? colv[half + i] : colv[i];
                colv[i] = keep + __shfl_xor_sync(0xffffffffu, send, m);
              }
            }
            if (up) off += half;
            cnt = half;
          } else {
            colv[0] += __shfl_xor_sync(0xffffffffu, colv[0], m);
            writer = writer && !up;
          }
        }
        double* sc = stage_cols + (size_t)(c & 1) * 4 * BN;
        if (writer) {
#pragma unroll
          for (int i = 0; i < NV; ++i) {
            if (i < cnt) {
              const int idx = off + i;  // = j*2 + e
              sc[wm * BN + wn * NT * 8 + (idx >> 1) * 8 + 2 * t + (idx & 1)] = colv[i];
            }
          }
        }
        named_bar_sync(2, kComputeThreads);
        if (tid < BN) {
          const int col = c * BN + tid;
          if (col < S) {
            double ps = sc[tid];
#pragma unroll
            for (int w = 1; w < T::WARPS_M; ++w) ps += sc[w * BN + tid];
            dd a;
            a.hi = __ldcg(acc_hi + col);
            a.lo = __ldcg(acc_lo + col);
            a = dd_add_d(a, ps);
            __stcg(acc_hi + col, a.hi);
            __stcg(acc_lo + col, a.lo);
          }
        }
      }
    }

    // ---- tile end: per-row statistics ----
    s1_0 += __shfl_xor_sync(0xffffffffu, s1_0, 1);
    s1_0 += __shfl_xor_sync(0xffffffffu, s1_0, 2);
    s1_1 += __shfl_xor_sync(0xffffffffu, s1_1, 1);
    s1_1 += __shfl_xor_sync(0xffffffffu, s1_1, 2);
    if (MODE == MODE_SCORE) {
      s2_0 += __shfl_xor_sync(0xffffffffu, s2_0, 1);
      s2_0 += __shfl_xor_sync(0xffffffffu, s2_0, 2);
      s2_1 += __shfl_xor_sync(0xffffffffu, s2_1, 1);
      s2_1 += __shfl_xor_sync(0xffffffffu, s2_1, 2);
      sr_0 += __shfl_xor_sync(0xffffffffu, sr_0, 1);
      sr_0 += __shfl_xor_sync(0xffffffffu, sr_0, 2);
      sr_1 += __shfl_xor_sync(0xffffffffu, sr_1, 1);
      sr_1 += __shfl_xor_sync(0xffffffffu, sr_1, 2);
    }
    if (t == 0) {
      double* sr = stage_rows + (size_t)wn * BM * 3;
      sr[r0 * 3 + 0] = s1_0;
      sr[r1 * 3 + 0] = s1_1;
      if (MODE == MODE_SCORE) {
        sr[r0 * 3 + 1] = s2_0;
        sr[r1 * 3 + 1] = s2_1;
        sr[r0 * 3 + 2] = sr_0;
        sr[r1 * 3 + 2] = sr_1;
      }
    }
    named_bar_sync(3, kComputeThreads);
    if (tid < BM) {  // whole warps: BM is a multiple of 32
      const long long p = tile * BM + tid;
      const bool valid = p < n;
      double S1 = stage_rows[tid * 3 + 0];
#pragma unroll
      for (int w = 1; w < T::WARPS_N; ++w) S1 += stage_rows[(size_t)w * BM * 3 + tid * 3 + 0];
      const double gb = S1 * invS;
      if (MODE == MODE_MATERIALISE) gbar_s[tid] = gb;
      const double wsum = warp_sum(valid ? gb : 0.0);
      if (lane == 0) sum_gbar = dd_add_d(sum_gbar, wsum);
      if (MODE == MODE_SCORE) {
        double S2 = stage_rows[tid * 3 + 1], SR = stage_rows[tid * 3 + 2];
#pragma unroll
        for (int w = 1; w < T::WARPS_N; ++w) {
          S2 += stage_rows[(size_t)w * BM * 3 + tid * 3 + 1];
          SR += stage_rows[(size_t)w * BM * 3 + tid * 3 + 2];
        }
        // centred quantities: v = g - gbar;  v.r = SR - gbar*sum(r);  |v|^2 = S2 - S*gbar^2
        double nrm2 = S2 - dS * gb * gb;
        if (nrm2 < 0.0) nrm2 = 0.0;
        const double dot = SR - gb * rsum;
        // bcores.py:78  corrs = vecs.dot(resid) / sqrt((vecs**2).sum(1)) / S
        const double score = dot / sqrt(nrm2) / dS;
        if (P.scores && valid) P.scores[p] = score;
        Best mine;
        mine.v = score;
        mine.i = valid ? (P.idx_offset + p) : -1;
        mine = best_warp(mine);
        best = best_merge(best, mine);
      }
    }
    if (MODE == MODE_MATERIALISE && !P.raw) {
      // second pass over this tile's own output: subtract the row mean, exact two-pass norm
      named_bar_sync(4, kComputeThreads);
      const double gb0 = gbar_s[r0], gb1 = gbar_s[r1];
      double q0 = 0.0, q1 = 0.0;
      for (int c = 0; c < nchunks; ++c) {
        const int cb = c * BN + wn * NT * 8 + 2 * t;
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
      if (t == 0) {
        stage_rows[(size_t)wn * BM * 3 + r0 * 3 + 1] = q0;
        stage_rows[(size_t)wn * BM * 3 + r1 * 3 + 1] = q1;
      }
      named_bar_sync(5, kComputeThreads);
      if (tid < BM && P.norms) {
        const long long p = tile * BM + tid;
        if (p < n) {
          double Q = stage_rows[tid * 3 + 1];
#pragma unroll
          for (int w = 1; w < T::WARPS_N; ++w) Q += stage_rows[(size_t)w * BM * 3 + tid * 3 + 1];
          P.norms[p] = sqrt(Q);
        }
      }
    }
  }

  // ---- CTA results: sum of row means (dd) and the local arg-max ----
  named_bar_sync(6, kComputeThreads);
  if (tid < BM && lane == 0) {
    stage_rows[warp * 4 + 0] = sum_gbar.hi;
    stage_rows[warp * 4 + 1] = sum_gbar.lo;
    stage_rows[warp * 4 + 2] = best.v;
    stage_rows[warp * 4 + 3] = __longlong_as_double(best.i);
  }
  named_bar_sync(7, kComputeThreads);
  if (tid == 0) {
    dd s = {stage_rows[0], stage_rows[1]};
    Best b = {stage_rows[2], __double_as_longlong(stage_rows[3])};
    for (int w = 1; w < BM / 32; ++w) {
      dd o = {stage_rows[w * 4 + 0], stage_rows[w * 4 + 1]};
      s = dd_add(s, o);
      Best ob = {stage_rows[w * 4 + 2], __double_as_longlong(stage_rows[w * 4 + 3])};
      b = best_merge(b, ob);
    }
    P.part_misc[blockIdx.x * 4 + 0] = s.hi;
    P.part_misc[blockIdx.x * 4 + 1] = s.lo;
    P.part_misc[blockIdx.x * 4 + 2] = b.v;
    P.part_misc[blockIdx.x * 4 + 3] = __longlong_as_double(b.i);
  }
}

// Reduce the per-CTA partials in fixed order.
//   out_dd : [2][Sld]  (hi plane, lo plane); element S = sum over rows of the row means
//   out_best: {score, index-as-double-bits}
__global__ void k_project_finalize(const double* __restrict__ part_colsum, const double* __restrict__ part_misc, int nctas,
                                   int S, int Sld, double* __restrict__ out_dd, double* __restrict__ out_best, int mode) {
  const int col = blockIdx.x * blockDim.x + threadIdx.x;
  if (mode != MODE_SCORE && col <= S) {
    dd a = {0.0, 0.0};
    if (col < S) {
      for (int c = 0; c < nctas; ++c) {
        dd o = {part_colsum[(size_t)c * 2 * Sld + col], part_colsum[(size_t)c * 2 * Sld + Sld + col]};
        a = dd_add(a, o);
      }
    } else {
      for (int c = 0; c < nctas; ++c) {
        dd o = {part_misc[c * 4 + 0], part_misc[c * 4 + 1]};
        a = dd_add(a, o);
      }
    }
    out_dd[col] = a.hi;
    out_dd[Sld + col] = a.lo;
  }
  if (mode == MODE_SCORE && blockIdx.x == 0) {
    Best b = {0.0, -1};
    for (int c = threadIdx.x; c < nctas; c += blockDim.x) {
      Best o = {part_misc[c * 4 + 2], __double_as_longlong(part_misc[c * 4 + 3])};
      b = best_merge(b, o);
    }
    __shared__ double sv[32];
    __shared__ long long si[32];
    b = best_warp(b);
    if ((threadIdx.x & 31) == 0) {
      sv[threadIdx.x >> 5] = b.v;
      si[threadIdx.x >> 5] = b.i;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int w = 1; w < (int)(blockDim.x >> 5); ++w) {
        Best o = {sv[w], si[w]};
        b = best_merge(b, o);
      }
      out_best[0] = b.v;
      out_best[1] = __longlong_as_double(b.i);
    }
  }
}

// ------------------------------------------------------------------ launch --
template <class F, int MODE, class T>
static cudaError_t launch_one(const ProjArgs& P, int grid, size_t smem, cudaStream_t st) {
  auto kern = k_project<F, MODE, T>;
  static bool attr_done = false;  // per instantiation
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxSmem);
    if (e != cudaSuccess) return e;
    attr_done = true;
  }
  kern<<<grid, kThreads, smem, st>>>(P);
  return cudaGetLastError();
}

template <class F, int MODE>
static cudaError_t launch_tile(const ProjArgs& P, int tile_cfg, int grid, size_t smem, cudaStream_t st) {
  if (tile_cfg == 0) return launch_one<F, MODE, Tile<64, 64>>(P, grid, smem, st);
  return launch_one<F, MODE, Tile<32, 32>>(P, grid, smem, st);
}

template <class F>
static cudaError_t launch_mode(const ProjArgs& P, int mode, int tile_cfg, int grid, size_t smem, cudaStream_t st) {
  switch (mode) {
    case MODE_COLSUM: return launch_tile<F, MODE_COLSUM>(P, tile_cfg, grid, smem, st);
    case MODE_SCORE: return launch_tile<F, MODE_SCORE>(P, tile_cfg, grid, smem, st);
    default: return launch_tile<F, MODE_MATERIALISE>(P, tile_cfg, grid, smem, st);
  }
}

int project_tile_config(int Dpad, int* BM, int* BN, int* ss_out, size_t* smem_out) {
  const int ss = (Dpad % 8 == 4) ? Dpad : Dpad + 4;
  const int cfgs[2][2] = {{64, 64}, {32, 32}};
  for (int i = 0; i < 2; ++i) {
    SmemLayout L = smem_layout(cfgs[i][0], cfgs[i][1], ss, Dpad);
    if (L.total <= kMaxSmem) {
      *BM = cfgs[i][0];
      *BN = cfgs[i][1];
      *ss_out = ss;
      *smem_out = L.total;
      return i;
    }
  }
  return -1;
}

cudaError_t launch_project(const ProjArgs& P, int model, int kind, int poly, int mode, int tile_cfg, int grid, size_t smem,
                           cudaStream_t st) {
  if (model == MODEL_LOGISTIC) {
    if (kind == KIND_LOGLIK) return launch_mode<LogisticF<KIND_LOGLIK, 0>>(P, mode, tile_cfg, grid, smem, st);
    if (poly == 20) return launch_mode<LogisticF<KIND_BETALIK, 20>>(P, mode, tile_cfg, grid, smem, st);
    if (poly == kPowPolyMax) return launch_mode<LogisticF<KIND_BETALIK, kPowPolyMax>>(P, mode, tile_cfg, grid, smem, st);
    return launch_mode<LogisticF<KIND_BETALIK, 0>>(P, mode, tile_cfg, grid, smem, st);
  } else if (model == MODEL_GAUSSIAN) {
    if (kind == KIND_LOGLIK) return launch_mode<GaussianF<KIND_LOGLIK>>(P, mode, tile_cfg, grid, smem, st);
    if (kind == KIND_BETALIK) return launch_mode<GaussianF<KIND_BETALIK>>(P, mode, tile_cfg, grid, smem, st);
    return launch_mode<GaussianF<KIND_BETAGRAD>>(P, mode, tile_cfg, grid, smem, st);
  } else {
    if (kind == KIND_LOGLIK) return launch_mode<NeurlinF<KIND_LOGLIK>>(P, mode, tile_cfg, grid, smem, st);
    return launch_mode<NeurlinF<KIND_BETALIK>>(P, mode, tile_cfg, grid, smem, st);
  }
}

cudaError_t launch_project_finalize(const double* part_colsum, const double* part_misc, int nctas, int S, int Sld,
                                    double* out_dd, double* out_best, int mode, cudaStream_t st) {
  const int threads = 128;
  const int blocks = (mode == MODE_SCORE) ? 1 : (S + 1 + threads - 1) / threads;
  k_project_finalize<<<blocks, (mode == MODE_SCORE) ? 256 : threads, 0, st>>>(part_colsum, part_misc, nctas, S, Sld, out_dd,
                                                                                out_best, mode);
  return cudaGetLastError();
}

}  // namespace bc
