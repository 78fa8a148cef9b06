// beta-cores B200: the small kernels around the projection.
//   * sample preparation (B operand, per-sample aux term, pivot sample)
//   * Gaussian per-row quadratic form  x Siginv x   (examples/common/gaussian.py:10)
//   * combination of double-double column-sum parts (one per rank) into the centred column sum
//   * coreset-side step: residual, correlations of the coreset points, weight gradient
//     (bayesiancoresets/coreset/bcores.py:77-79, :144-146)
//   * projected ADAM update (bayesiancoresets/util/opt.py:45-52)
// All of these touch S- or M-sized data: they are latency-bound, one or a few CTAs each.
#include "bc_kernels.h"

namespace bc {

// fixed-order block sum; result valid in every thread.  blockDim.x <= 1024, multiple of 32.
__device__ __forceinline__ double block_sum(double x, double* red /* >= 33 doubles */) {
  x = warp_sum(x);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = blockDim.x >> 5;
  __syncthreads();  // protect `red` from a previous call
  if (l == 0) red[w] = x;
  __syncthreads();
  if (w == 0) {
    double y = (l < nw) ? red[l] : 0.0;
    y = warp_sum(y);
    if (l == 0) red[32] = y;
  }
  __syncthreads();
  return red[32];
}

// ------------------------------------------------------- sample preparation --
// one block per sample s.  Gaussian: B[s][:] = Siginv theta_s and colaux[s] = theta_s Siginv theta_s -- two D-long dot
// products per thread, both summed over j in increasing order; the first walks a ROW of Siginv, so it reads the
// transposed copy siginvT (made once per bc_set_potential) to keep a warp's loads coalesced (the row walk of the original
// cost 32 sectors per load: 41 us per call at D = 100, now 3).
__global__ void __launch_bounds__(128) k_prepare_rows(int model, const double* __restrict__ theta, int S, int D, int ldt,
                                                      const double* __restrict__ siginv, const double* __restrict__ siginvT,
                                                      double* __restrict__ B, int ldb, double* __restrict__ colaux,
                                                      const int* __restrict__ fexp, unsigned long long* __restrict__ absmax_slot,
                                                      const int* __restrict__ siginv_diag) {
  __shared__ double red[33];
  __shared__ unsigned long long wm[4], wb[4];
  unsigned long long mx = 0, bad = 0;   // absmax_slot: max |B' entry| bit pattern / non-finite flag of the feature-scaled samples
  extern __shared__ double ths[];   // [D] this sample
  const int s = blockIdx.x;
  const double* th = theta + (size_t)s * ldt;
  bool diag = false;
  if (model == MODEL_GAUSSIAN) {
    int finite = 1;
    for (int k = threadIdx.x; k < D; k += blockDim.x) {
      ths[k] = th[k];
      finite &= isfinite(th[k]) ? 1 : 0;
    }
    // a DIAGONAL Siginv (k_offdiag_test; the reference's Gaussian example uses Sig = 500 I) and a finite sample: each dot
    // product below has one non-zero term -- the others add a signed zero -- so it is taken alone, bit for bit the same
    diag = __syncthreads_and(finite) && siginv_diag && *siginv_diag == 1;
  }
  double part = 0.0;
  for (int k = threadIdx.x; k < ldb; k += blockDim.x) {
    double b = 0.0;
    if (k < D) {
      if (model == MODEL_GAUSSIAN) {
        // B[s][k] = (Siginv theta_s)[k]      gaussian.py:12  x.dot(Siginv.dot(th.T))
        double acc = 0.0, acc2 = 0.0;
        int j = 0;
        if (diag) {
          const double skk = __ldg(siginv + (size_t)k * D + k);
          acc = fma(skk, ths[k], acc);
          acc2 = fma(ths[k], skk, acc2);
          j = D;
        }
        for (; j + 4 <= D; j += 4) {
          double u[4], v[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            u[q] = siginvT ? __ldg(siginvT + (size_t)(j + q) * D + k) : __ldg(siginv + (size_t)k * D + j + q);
            v[q] = __ldg(siginv + (size_t)(j + q) * D + k);
          }
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            acc = fma(u[q], ths[j + q], acc);
            acc2 = fma(ths[j + q], v[q], acc2);  // (th.dot(Siginv))[k]   gaussian.py:11
          }
        }
        for (; j < D; ++j) {
          acc = fma(siginvT ? __ldg(siginvT + (size_t)j * D + k) : __ldg(siginv + (size_t)k * D + j), ths[j], acc);
          acc2 = fma(ths[j], __ldg(siginv + (size_t)j * D + k), acc2);
        }
        b = acc;
        part = fma(ths[k], acc2, part);
      } else {
        b = th[k];
      }
    }
    B[(size_t)s * ldb + k] = b;
    if (absmax_slot && k < D) {          // what k_sample_absmax (bc_project_q.cu) gathers, while the entry is in a register
      const double v = fabs(fexp ? scalbn(b, __ldg(fexp + k)) : b);
      if (!isfinite(v)) bad = 1;
      const unsigned long long bits = (unsigned long long)__double_as_longlong(v);
      if (isfinite(v) && bits > mx) mx = bits;
    }
  }
  if (absmax_slot) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const unsigned long long om = __shfl_xor_sync(0xffffffffu, mx, o);
      mx = om > mx ? om : mx;
      bad |= __shfl_xor_sync(0xffffffffu, bad, o);
    }
    if ((threadIdx.x & 31) == 0) {
      wm[threadIdx.x >> 5] = mx;
      wb[threadIdx.x >> 5] = bad;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int w = 1; w < (int)(blockDim.x >> 5); ++w) {
        mx = wm[w] > mx ? wm[w] : mx;
        bad |= wb[w];
      }
      if (mx) atomicMax(absmax_slot, mx);
      if (bad) atomicOr(absmax_slot + 1, 1ull);
    }
  }
  if (model == MODEL_GAUSSIAN) {
    const double tot = block_sum(part, red);
    if (threadIdx.x == 0) colaux[s] = tot;
  }
}

// flag[0] = 1 if every off-diagonal entry of the D x D matrix A is (signed) zero, else 0.  One CTA.
__global__ void __launch_bounds__(256) k_offdiag_test(const double* __restrict__ A, int D, int* __restrict__ flag) {
  int nz = 0;
  for (int q = threadIdx.x; q < D * D; q += blockDim.x) {
    const int i = q / D, j = q - i * D;
    if (i != j && A[q] != 0.0) nz = 1;
  }
  nz = __syncthreads_or(nz);
  if (threadIdx.x == 0) flag[0] = nz ? 0 : 1;
}
cudaError_t launch_offdiag_test(const double* A, int D, int* flag, cudaStream_t st) {
  k_offdiag_test<<<1, 256, 0, st>>>(A, D, flag);
  return cudaGetLastError();
}

// bbar[k] = mean_s B[s][k] for k < ldb; bbar[ldb] = mean_s colaux[s] (0 if no colaux).
// bbar is the PIVOT sample of the FP64 route (its potential is subtracted from a row's before anything is accumulated): any
// value near the sample mean serves, but it must be the same on every launch configuration, so the summation order is
// fixed: kMeanChains interleaved chains per column (chain p adds the samples s = p, p + kMeanChains, ... in order), combined
// in chain order.  A CTA owns 32 columns; thread (chain, column) reads its samples straight from global memory -- the loads
// are independent of the running sum and coalesced over the columns -- so a column costs S / kMeanChains dependent adds
// (a single chain per column took 40 us at S = 1024).
constexpr int kMeanCols = 32, kMeanChains = 8, kMeanThreads = kMeanCols * kMeanChains;
__global__ void __launch_bounds__(kMeanThreads) k_prepare_mean(const double* __restrict__ B, int S, int ldb, const double* __restrict__ colaux,
                                                               double* __restrict__ bbar) {
  __shared__ double part[kMeanChains][kMeanCols];
  const int kk = threadIdx.x & (kMeanCols - 1), p = threadIdx.x / kMeanCols;
  const int k = blockIdx.x * kMeanCols + kk;
  double acc = 0.0;
  if (k <= ldb && (k < ldb || colaux)) {
    const double* src = (k < ldb) ? B + k : colaux;
    const size_t stride = (k < ldb) ? (size_t)ldb : 1;
    int s = p;
    for (; s + 3 * kMeanChains < S; s += 4 * kMeanChains) {
      const double v0 = src[(size_t)s * stride], v1 = src[(size_t)(s + kMeanChains) * stride];
      const double v2 = src[(size_t)(s + 2 * kMeanChains) * stride], v3 = src[(size_t)(s + 3 * kMeanChains) * stride];
      acc += v0;
      acc += v1;
      acc += v2;
      acc += v3;
    }
    for (; s < S; s += kMeanChains) acc += src[(size_t)s * stride];
  }
  part[p][kk] = acc;
  __syncthreads();
  if (p == 0 && k <= ldb) {
    double tot = part[0][kk];
#pragma unroll
    for (int q = 1; q < kMeanChains; ++q) tot += part[q][kk];
    bbar[k] = tot / (double)S;
  }
}

cudaError_t launch_prepare_samples(int model, const double* theta, int S, int D, int ldt, const double* siginv, const double* siginvT,
                                   double* B, int ldb, double* colaux, double* bbar, const int* fexp, unsigned long long* absmax_slot,
                                   const int* siginv_diag, cudaStream_t st) {
  BC_PREFER_MAX_SHARED(k_prepare_rows);
  k_prepare_rows<<<S, 128, (size_t)D * sizeof(double), st>>>(model, theta, S, D, ldt, siginv, siginvT, B, ldb, colaux, fexp, absmax_slot,
                                                             siginv_diag);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  BC_PREFER_MAX_SHARED(k_prepare_mean);
  k_prepare_mean<<<(ldb + 1 + kMeanCols - 1) / kMeanCols, kMeanThreads, 0, st>>>(B, S, ldb, model == MODEL_GAUSSIAN ? colaux : nullptr, bbar);
  return cudaGetLastError();
}

// --------------------------------------------------------------- x Siginv x --
// one warp per row; out[r] = sum_k x[k] * (sum_j x[j] Siginv[j][k])      gaussian.py:10
__global__ void k_rowquad(const double* __restrict__ X, long long n, int D, long long ldx, const double* __restrict__ siginv,
                          double* __restrict__ out) {
  extern __shared__ double xs[];  // [warps][D]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  double* x = xs + (size_t)warp * D;
  for (long long r = (long long)blockIdx.x * nw + warp; r < n; r += (long long)gridDim.x * nw) {
    for (int k = lane; k < D; k += 32) x[k] = X[r * ldx + k];
    __syncwarp();
    double acc = 0.0;
    for (int k = lane; k < D; k += 32) {
      double tk = 0.0;
      for (int j = 0; j < D; ++j) tk = fma(x[j], __ldg(siginv + (size_t)j * D + k), tk);
      acc = fma(x[k], tk, acc);
    }
    acc = warp_sum(acc);
    if (lane == 0) out[r] = acc;
    __syncwarp();
  }
}

cudaError_t launch_rowquad(const double* X, long long n, int D, long long ldx, const double* siginv, double* out, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  const int warps = 8;
  long long blocks = (n + warps - 1) / warps;
  if (blocks > 148 * 8) blocks = 148 * 8;
  k_rowquad<<<(int)blocks, warps * 32, (size_t)warps * D * sizeof(double), st>>>(X, n, D, ldx, siginv, out);
  return cudaGetLastError();
}

// ------------------------------------------------------- dd parts -> colsum --
// parts: [nparts][2][Sld]; element S of each part = sum of row means.
// out[s] = (sum_p part_p[s]) - (sum_p part_p[S])   i.e. the column sum of the CENTRED matrix.
__global__ void k_colsum_combine(const double* __restrict__ parts, int nparts, int S, int Sld, double* __restrict__ out) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= S) return;
  dd a = {0.0, 0.0}, m = {0.0, 0.0};
  for (int p = 0; p < nparts; ++p) {
    const double* q = parts + (size_t)p * 2 * Sld;
    dd o = {q[s], q[Sld + s]};
    a = dd_add(a, o);
    dd om = {-q[S], -q[Sld + S]};
    m = dd_add(m, om);
  }
  a = dd_add(a, m);
  out[s] = a.hi;
}

cudaError_t launch_colsum_combine(const double* parts, int nparts, int S, int Sld, double* out, cudaStream_t st) {
  BC_PREFER_MAX_SHARED(k_colsum_combine);
  k_colsum_combine<<<(S + 127) / 128, 128, 0, st>>>(parts, nparts, S, Sld, out);
  return cudaGetLastError();
}

// ------------------------------------------------------- coreset-side step --
// resid[s] = scaling*colsum[s] - sum_m w[m] Vc[m][s];  resid[S] = sum_s resid[s]      bcores.py:77 / :145
__global__ void k_core_resid(const double* __restrict__ colsum, double scaling, const double* __restrict__ Vc, int M, int S,
                             long long ldv, const double* __restrict__ w, double* __restrict__ resid) {
  __shared__ double red[33];
  double tot = 0.0;
  for (int s = threadIdx.x; s < S; s += blockDim.x) {
    double acc = 0.0;
    for (int m = 0; m < M; ++m) acc = fma(w[m], Vc[(size_t)m * ldv + s], acc);
    const double r = scaling * colsum[s] - acc;
    resid[s] = r;
    tot += r;
  }
  tot = block_sum(tot, red);
  if (threadIdx.x == 0) resid[S] = tot;
}

cudaError_t launch_core_resid(const double* colsum, double scaling, const double* Vc, int M, int S, long long ldv, const double* w,
                              double* resid, cudaStream_t st) {
  k_core_resid<<<1, 1024, 0, st>>>(colsum, scaling, Vc, M, S, ldv, w, resid);
  return cudaGetLastError();
}

// out[0] = max_{m >= skip} | Vc_m . r / |Vc_m| | / S   (np.max semantics: NaN propagates)      bcores.py:79
// grad[m] = -(Vc_m . r) / S                                                               bcores.py:146
template <bool GRAD>
__global__ void k_core_rows(const double* __restrict__ Vc, int M, int S, long long ldv, const double* __restrict__ resid, int skip,
                            double* __restrict__ out) {
  __shared__ double wmax[32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  double best = -INFINITY;
  for (int m = warp + (GRAD ? 0 : skip); m < M; m += nw) {
    const double* v = Vc + (size_t)m * ldv;
    double dot = 0.0, n2 = 0.0;
    for (int s = lane; s < S; s += 32) {
      const double x = v[s];
      dot = fma(x, resid[s], dot);
      if (!GRAD) n2 = fma(x, x, n2);
    }
    dot = warp_sum(dot);
    if (GRAD) {
      if (lane == 0) out[m] = -dot / (double)S;
    } else {
      n2 = warp_sum(n2);
      best = nanmax(best, fabs(dot / sqrt(n2)) / (double)S);
    }
  }
  if (!GRAD) {
    if (lane == 0) wmax[warp] = best;
    __syncthreads();
    if (threadIdx.x == 0) {
      double b = wmax[0];
      for (int i = 1; i < nw; ++i) b = nanmax(b, wmax[i]);
      out[0] = b;
    }
  }
}

cudaError_t launch_core_maxcorr(const double* Vc, int M, int S, long long ldv, const double* resid, int skip, double* out,
                                cudaStream_t st) {
  k_core_rows<false><<<1, 1024, 0, st>>>(Vc, M, S, ldv, resid, skip, out);
  return cudaGetLastError();
}

cudaError_t launch_core_grad(const double* Vc, int M, int S, long long ldv, const double* resid, double* grad, cudaStream_t st) {
  if (M <= 0) return cudaSuccess;
  k_core_rows<true><<<1, 1024, 0, st>>>(Vc, M, S, ldv, resid, 0, grad);
  return cudaGetLastError();
}

// resid + grad + ADAM of one optimiser step in ONE launch (bc_greedy_opt_step): the three kernels above back to back in a
// single CTA of the same shape -- the same loops, the same reduction orders, so the same bits -- separated by CTA barriers
// instead of launch boundaries (each of them runs for 4 us; the gaps between them cost as much again).
__device__ __forceinline__ double adam_update(double gi, double x, double* m1, double* m2, int i, double lr, double b1, double b2, double c1,
                                              double c2, double eps, const unsigned char* nn_mask);
__global__ void __launch_bounds__(1024) k_core_step(const double* __restrict__ colsum, double scaling, const double* __restrict__ Vc, int M,
                                                    int S, long long ldv, double* x, double* resid, double* grad, double* m1, double* m2,
                                                    double lr, double b1, double b2, double c1, double c2, double eps,
                                                    const unsigned char* __restrict__ nn_mask, const double* __restrict__ sched,
                                                    int* step_counter) {
  __shared__ double red[33];
  // sched / step_counter (CUDA-graph replays: the launch parameters are frozen, so what changes from step to step lives in
  // device memory): step i = *step_counter takes (lr, c1, c2) = sched[3 i ..] and leaves i + 1 behind
  int step = 0;
  if (sched) {
    step = *step_counter;
    lr = sched[3 * step];
    c1 = sched[3 * step + 1];
    c2 = sched[3 * step + 2];
  }
  {   // k_core_resid
    double tot = 0.0;
    for (int s = threadIdx.x; s < S; s += blockDim.x) {
      double acc = 0.0;
      for (int m = 0; m < M; ++m) acc = fma(x[m], Vc[(size_t)m * ldv + s], acc);
      const double r = scaling * colsum[s] - acc;
      resid[s] = r;
      tot += r;
    }
    tot = block_sum(tot, red);
    if (threadIdx.x == 0) resid[S] = tot;
  }
  __syncthreads();
  {   // k_core_rows<true>
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    for (int m = warp; m < M; m += nw) {
      const double* v = Vc + (size_t)m * ldv;
      double dot = 0.0;
      for (int s = lane; s < S; s += 32) dot = fma(v[s], resid[s], dot);
      dot = warp_sum(dot);
      if (lane == 0) grad[m] = -dot / (double)S;
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < M; i += blockDim.x) x[i] = adam_update(grad[i], x[i], m1, m2, i, lr, b1, b2, c1, c2, eps, nn_mask);   // k_adam
  if (sched && threadIdx.x == 0) *step_counter = step + 1;   // (every thread read it before the first barrier above)
}

cudaError_t launch_core_step(const double* colsum, double scaling, const double* Vc, int M, int S, long long ldv, double* x, double* resid,
                             double* grad, double* m1, double* m2, double lr, double b1, double b2, double c1, double c2, double eps,
                             const unsigned char* nn_mask, const double* sched, int* step_counter, cudaStream_t st) {
  BC_PREFER_MAX_SHARED(k_core_step);
  k_core_step<<<1, 1024, 0, st>>>(colsum, scaling, Vc, M, S, ldv, x, resid, grad, m1, m2, lr, b1, b2, c1, c2, eps, nn_mask, sched,
                                  step_counter);
  return cudaGetLastError();
}

// ---------------------------------------------------------- projected ADAM --
// util/opt.py:45-52, operation for operation (no FMA contraction: matches numpy's rounding):
//   m1 = b1*m1 + (1-b1)*g;  m2 = b2*m2 + (1-b2)*g**2
//   upd = lr*m1/c1/(eps + sqrt(m2/c2));  x -= upd;  x = max(x, 0)   [only where nn_mask, all if null]
// c1 = 1-b1**(i+1), c2 = 1-b2**(i+1) are computed by the host (python float pow, like the reference).
__device__ __forceinline__ double adam_update(double gi, double x, double* m1, double* m2, int i, double lr, double b1, double b2, double c1,
                                              double c2, double eps, const unsigned char* nn_mask) {
  const double a1 = __dadd_rn(__dmul_rn(b1, m1[i]), __dmul_rn(__dsub_rn(1.0, b1), gi));
  const double a2 = __dadd_rn(__dmul_rn(b2, m2[i]), __dmul_rn(__dsub_rn(1.0, b2), __dmul_rn(gi, gi)));
  m1[i] = a1;
  m2[i] = a2;
  const double upd = __ddiv_rn(__ddiv_rn(__dmul_rn(lr, a1), c1), __dadd_rn(eps, __dsqrt_rn(__ddiv_rn(a2, c2))));
  double xi = __dsub_rn(x, upd);
  if (!nn_mask || nn_mask[i]) xi = (xi > 0.0 || isnan(xi)) ? xi : 0.0;
  return xi;
}
__global__ void k_adam(const double* __restrict__ g, double* __restrict__ x, double* __restrict__ m1, double* __restrict__ m2, int n,
                       double lr, double b1, double b2, double c1, double c2, double eps, const unsigned char* __restrict__ nn_mask) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  x[i] = adam_update(g[i], x[i], m1, m2, i, lr, b1, b2, c1, c2, eps, nn_mask);
}

cudaError_t launch_adam(const double* g, double* x, double* m1, double* m2, int n, double lr, double b1, double b2, double c1,
                        double c2, double eps, const unsigned char* nn_mask, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  k_adam<<<(n + 127) / 128, 128, 0, st>>>(g, x, m1, m2, n, lr, b1, b2, c1, c2, eps, nn_mask);
  return cudaGetLastError();
}

// ------------------------------------------------- pseudo-point gradients ----
// BatchPSVICoreset (bayesiancoresets/coreset/bpsvi.py:49-55): gradient of the objective in the LOCATION of pseudo-point m,
//   ugrad[m][d] = -(w_m / S) sum_s resid_s * pg[m][s][d],
// pg = the model's d(log-likelihood)/d(point), centred over its LAST axis as the reference's projector does
// (coreset/projector.py:31).  The (M, S, D) tensor is never formed for the built-in models:
//   logistic (examples/common/model_lr.py:107-114):  pg = sigma(m_ms) theta_sd,  m = -p_m . theta_s
//        -> ugrad[m][d] = -(w_m/S) sum_s [resid_s sigma_ms] (theta_sd - mean_d theta_s)
//   Gaussian (examples/common/gaussian.py:17-20):    pg = (theta_s Siginv)_d - (p_m Siginv)_d
//        -> ugrad[m][d] = -(w_m/S) [ sum_s resid_s (A_sd - mean_d A_s) - (sum_s resid_s) (Bm_d - mean_d Bm) ],  A = Theta Siginv
// One CTA per pseudo-point; M, S, D are small (latency-bound).  B = the prepared samples (Theta, or Siginv Theta).
template <int MODEL>
__global__ void __launch_bounds__(256) k_pgrad_model(const double* __restrict__ P, int M, long long ldp, const double* __restrict__ B,
                                                     int S, int D, int ldb, const double* __restrict__ siginv,
                                                     const double* __restrict__ w, const double* __restrict__ resid,
                                                     double* __restrict__ out, long long ldo) {
  extern __shared__ double sm[];
  double* ps = sm;            // [D]   pseudo-point
  double* cs = ps + D;        // [S]   per-sample weight of the sample term
  double* tb = cs + S;        // [S]   mean over d of the sample term
  double* bm = tb + S;        // [D]   Gaussian: (p_m Siginv)_d
  __shared__ double red[33];
  const int m = blockIdx.x;
  for (int k = threadIdx.x; k < D; k += blockDim.x) ps[k] = P[m * ldp + k];
  __syncthreads();
  for (int s = threadIdx.x; s < S; s += blockDim.x) {
    const double* b = B + (size_t)s * ldb;
    double dot = 0.0, sum = 0.0;
    for (int k = 0; k < D; ++k) {
      dot = fma(ps[k], b[k], dot);
      sum += b[k];
    }
    double c = resid[s];
    if (MODEL == MODEL_LOGISTIC) {
      const double mm = -dot;
      const double sg = (mm < 100.0) ? exp(mm) / (1.0 + exp(mm)) : 1.0;   // model_lr.py:111-113
      c *= sg;
    }
    cs[s] = c;
    tb[s] = sum / (double)D;
  }
  double bmean = 0.0;
  if (MODEL == MODEL_GAUSSIAN) {
    double part = 0.0;
    for (int d = threadIdx.x; d < D; d += blockDim.x) {
      double acc = 0.0;
      for (int j = 0; j < D; ++j) acc = fma(ps[j], siginv[(size_t)j * D + d], acc);   // (p Siginv)_d   gaussian.py:20
      bm[d] = acc;
      part += acc;
    }
    bmean = block_sum(part, red) / (double)D;
  }
  __syncthreads();
  const double R = resid[S];
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    double acc = 0.0;
    for (int s = 0; s < S; ++s) acc = fma(cs[s], B[(size_t)s * ldb + d] - tb[s], acc);
    if (MODEL == MODEL_GAUSSIAN) acc -= R * (bm[d] - bmean);
    out[m * ldo + d] = -(w[m] * acc) / (double)S;
  }
}

// opaque gradient callbacks: G = (M, S, D) contiguous, already produced on the host by the user's function;
// centre != 0 subtracts the mean over the last axis first (projector.py:31)
__global__ void __launch_bounds__(256) k_pgrad_dense(const double* __restrict__ G, int M, int S, int D, const double* __restrict__ w,
                                                     const double* __restrict__ resid, int centre, double* __restrict__ out,
                                                     long long ldo) {
  extern __shared__ double sm[];
  double* cs = sm;       // [S] resid
  double* tb = cs + S;   // [S] row means
  const int m = blockIdx.x;
  const double* g = G + (size_t)m * S * D;
  for (int s = threadIdx.x; s < S; s += blockDim.x) {
    cs[s] = resid[s];
    double sum = 0.0;
    if (centre)
      for (int d = 0; d < D; ++d) sum += g[(size_t)s * D + d];
    tb[s] = sum / (double)D;
  }
  __syncthreads();
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    double acc = 0.0;
    for (int s = 0; s < S; ++s) acc = fma(cs[s], g[(size_t)s * D + d] - tb[s], acc);
    out[m * ldo + d] = -(w[m] * acc) / (double)S;
  }
}

cudaError_t launch_pgrad_model(int model, const double* P, int M, long long ldp, const double* B, int S, int D, int ldb,
                               const double* siginv, const double* w, const double* resid, double* out, long long ldo,
                               cudaStream_t st) {
  if (M <= 0) return cudaSuccess;
  const size_t smem = (size_t)(2 * D + 2 * S) * sizeof(double);
  if (model == MODEL_LOGISTIC) {
    k_pgrad_model<MODEL_LOGISTIC><<<M, 256, smem, st>>>(P, M, ldp, B, S, D, ldb, siginv, w, resid, out, ldo);
  } else {
    k_pgrad_model<MODEL_GAUSSIAN><<<M, 256, smem, st>>>(P, M, ldp, B, S, D, ldb, siginv, w, resid, out, ldo);
  }
  return cudaGetLastError();
}

cudaError_t launch_pgrad_dense(const double* G, int M, int S, int D, const double* w, const double* resid, int centre, double* out,
                               long long ldo, cudaStream_t st) {
  if (M <= 0) return cudaSuccess;
  k_pgrad_dense<<<M, 256, (size_t)2 * S * sizeof(double), st>>>(G, M, S, D, w, resid, centre, out, ldo);
  return cudaGetLastError();
}

}  // namespace bc
