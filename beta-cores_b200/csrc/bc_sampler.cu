// beta-cores B200: device-side posterior samplers.
//
// The coreset classes call the user's sampler once per optimiser step; in the reference it is host numpy
// (examples/zellner_logreg/main.py:86-111,139-144 -> bayesiancoresets/util/opt.py:10-33 get_laplace; examples/common/gaussian.py:28-32).
// With the projection pass at a few milliseconds per GPU that host call (mode search, D x D Cholesky, triangular
// inverse, S x D x D product, 1 MB upload) is the largest non-kernel item of a step.  These kernels do the same algebra
// on the device.  Everything is D x D with D <= 128 and M (coreset size) rows: one CTA, latency-bound.
//   k_laplace_logistic : mode of the weighted logistic log-joint with N(0, I) prior by damped Newton steps (the same
//                        iteration as examples/common/model_lr.py::_newton_mode), then L = inverse of the lower Cholesky
//                        factor of the negative Hessian at the mode -- get_laplace's (mu, LSig).
//   k_sample_affine    : Theta = mu + R L^T for S x D standard normals R drawn by the HOST stream (so that the random
//                        numbers are the reference's), sampler line `mu + randn(S, D).dot(LSig.T)`.
#include "bc_kernels.h"
#include "bc_common.cuh"

namespace bc {

constexpr int kLapThreads = 512;    // one CTA; every phase is latency-bound; 512 threads leave 128 registers each for the register-tiled Cholesky

__device__ __forceinline__ double blk_sum(double x, double* red) {  // all threads get the result
  x = warp_sum(x);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = blockDim.x >> 5;
  __syncthreads();
  if (l == 0) red[w] = x;
  __syncthreads();
  double y = (l < nw) ? red[l] : 0.0;
  y = warp_sum(y);
  return y;
}

// (Tried in round 2 for coresets of a handful of points, profiles/r02_laplace_parts.txt: H = I + U U^T with U = Z^T diag(r) of rank M
// factorises by an M x M recurrence, L_kk = sqrt(p), L_ik = u_i . G u_k / sqrt(p), G <- G - (G u_k)(G u_k)^T / p, p = 1 + u_k G u_k,
// without the D x D Hessian -- one warp, lane = row of G.  Its D sequential steps of shuffles + sqrt + reciprocal took 135 k cycles
// at D = 128, M = 6 against 19 k (Hessian) + 106 k for the blocked dense factorisation below: not kept.)
// In-place lower Cholesky of the D x D matrix H (row-major, leading dimension ld = D + 1, odd, so that walks down a column
// are free of shared-memory bank conflicts) in shared memory.  Blocked in panels of kPanel columns, three phases per panel:
//   A  warp 0 factorises the kPanel x kPanel diagonal block in REGISTERS (lane = row, shuffles broadcast the pivot row):
//      the serial part of the algorithm costs one sqrt + one reciprocal + two shuffle rounds per column and never touches
//      shared memory;
//   B  one thread per row below the block: x L11^T = A21 by forward substitution with the 16 row entries in registers;
//   C  every warp updates 32 x 8 blocks of the trailing triangle with 4 x 2 register tiles (0.75 shared-memory loads per FMA
//      instead of 2 for an entry-per-thread update).
// Every entry sees exactly the operations of the unblocked right-looking algorithm in the same order (subtract the
// contributions of columns 0, 1, ... in turn, then scale by the reciprocal pivot), so the factor does not depend on the
// blocking.  rd[j] = 1 / L[j][j].  Returns false if H is not positive definite.
constexpr int kPanel = 16;
__device__ bool chol_lower(double* H, int D, int ld, double* rd, int* flag) {
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nw = blockDim.x >> 5;
  if (tid == 0) *flag = 0;
  __syncthreads();
  for (int j0 = 0; j0 < D; j0 += kPanel) {
    const int nb = min(kPanel, D - j0);
    if (wid == 0) {                                   // ---- A
      double a[kPanel];
      const int r = lane;
#pragma unroll
      for (int c = 0; c < kPanel; ++c)                // rows / columns beyond the matrix: identity (factorises to itself)
        a[c] = (r < nb && c <= r) ? H[(j0 + r) * ld + j0 + c] : ((c == r) ? 1.0 : 0.0);
      bool bad = false;
      double myinv = 0.0;
#pragma unroll
      for (int c = 0; c < kPanel; ++c) {
        const double d = __shfl_sync(0xffffffffu, a[c], c);
        if (!(d > 0.0)) bad = true;
        const double s = sqrt(d), inv = 1.0 / s;
        a[c] = (r == c) ? s : a[c] * inv;
        if (r == c) myinv = inv;
#pragma unroll
        for (int k = 0; k < kPanel; ++k) {
          if (k <= c) continue;                      // (constant trip counts: the loops unroll completely, a[] stays in registers)
          const double lkc = __shfl_sync(0xffffffffu, a[c], k);
          if (r >= k) a[k] = fma(-a[c], lkc, a[k]);
        }
      }
      if (bad && lane == 0) *flag = 1;
      if (r < nb) {
#pragma unroll
        for (int c = 0; c < kPanel; ++c)
          if (c <= r) H[(j0 + r) * ld + j0 + c] = a[c];
        rd[j0 + r] = myinv;
      }
    }
    __syncthreads();
    if (*flag) return false;
    const int r0 = j0 + nb;
    if (r0 >= D) break;                               // (nb == kPanel from here on)
    if (tid < D - r0) {                               // ---- B
      const int i = r0 + tid;
      double x[kPanel];
#pragma unroll
      for (int c = 0; c < kPanel; ++c) x[c] = H[i * ld + j0 + c];
#pragma unroll
      for (int c = 0; c < kPanel; ++c) {
        x[c] *= rd[j0 + c];
#pragma unroll
        for (int k = 0; k < kPanel; ++k)
          if (k > c) x[k] = fma(-x[c], H[(j0 + k) * ld + j0 + c], x[k]);
      }
#pragma unroll
      for (int c = 0; c < kPanel; ++c) H[i * ld + j0 + c] = x[c];
    }
    __syncthreads();
    {                                                 // ---- C
      const int n = D - r0, li = lane >> 2, lk = lane & 3;
      int cnt = 0;
      for (int bi = 0; bi * 32 < n; ++bi) {
        for (int bk = 0; bk * 8 <= bi * 32 + 31 && bk * 8 < n; ++bk, ++cnt) {
          if (cnt % nw != wid) continue;
          int ia[4], kb[2];
          double acc[4][2];
#pragma unroll
          for (int q = 0; q < 4; ++q) ia[q] = r0 + bi * 32 + q * 8 + li;
#pragma unroll
          for (int q = 0; q < 2; ++q) kb[q] = r0 + bk * 8 + q * 4 + lk;
#pragma unroll
          for (int q = 0; q < 4; ++q)
#pragma unroll
            for (int p = 0; p < 2; ++p) acc[q][p] = (ia[q] < D && kb[p] <= ia[q]) ? H[ia[q] * ld + kb[p]] : 0.0;
          const int il[4] = {min(ia[0], D - 1), min(ia[1], D - 1), min(ia[2], D - 1), min(ia[3], D - 1)};
          const int kl[2] = {min(kb[0], D - 1), min(kb[1], D - 1)};
#pragma unroll 4
          for (int c = 0; c < kPanel; ++c) {
            double Li[4], Lk[2];
#pragma unroll
            for (int q = 0; q < 4; ++q) Li[q] = H[il[q] * ld + j0 + c];
#pragma unroll
            for (int p = 0; p < 2; ++p) Lk[p] = H[kl[p] * ld + j0 + c];
#pragma unroll
            for (int q = 0; q < 4; ++q)
#pragma unroll
              for (int p = 0; p < 2; ++p) acc[q][p] = fma(-Li[q], Lk[p], acc[q][p]);
          }
#pragma unroll
          for (int q = 0; q < 4; ++q)
#pragma unroll
            for (int p = 0; p < 2; ++p)
              if (ia[q] < D && kb[p] <= ia[q]) H[ia[q] * ld + kb[p]] = acc[q][p];
        }
      }
    }
    __syncthreads();
  }
  return true;
}

// Triangular solves by ONE warp with the vector in registers (lane owns entries 32 q + lane): per column one shuffle, one
// multiply by the reciprocal pivot and one FMA per owned entry; the factor is read down a column (forward, conflict-free
// with the odd leading dimension) or along a row (backward).
constexpr int kMaxQ = 6;   // D <= 192 (the shared-memory budget already limits D to 168)
__device__ __forceinline__ void tri_forward(const double* L, const double* rd, int D, int ld, double (&r)[kMaxQ]) {   // L y = x
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int q0 = 0; q0 < kMaxQ; ++q0) {
    if (q0 * 32 >= D) continue;
    for (int l = 0; l < 32; ++l) {
      const int j = q0 * 32 + l;
      if (j >= D) break;
      const double xj = __shfl_sync(0xffffffffu, r[q0], l) * rd[j];
      if (lane == l) r[q0] = xj;
#pragma unroll
      for (int q = 0; q < kMaxQ; ++q) {
        const int i = q * 32 + lane;
        if (q >= q0 && i > j && i < D) r[q] = fma(-L[i * ld + j], xj, r[q]);
      }
    }
  }
}
__device__ __forceinline__ void tri_backward(const double* L, const double* rd, int D, int ld, double (&r)[kMaxQ]) {  // L^T z = y
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int q0 = kMaxQ - 1; q0 >= 0; --q0) {
    if (q0 * 32 >= D) continue;
    for (int l = 31; l >= 0; --l) {
      const int j = q0 * 32 + l;
      if (j >= D) continue;
      const double xj = __shfl_sync(0xffffffffu, r[q0], l) * rd[j];
      if (lane == l) r[q0] = xj;
#pragma unroll
      for (int q = 0; q < kMaxQ; ++q) {
        const int i = q * 32 + lane;
        if (q <= q0 && i < j) r[q] = fma(-L[j * ld + i], xj, r[q]);
      }
    }
  }
}
__device__ __forceinline__ void tri_load(const double* x, int D, double (&r)[kMaxQ]) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int q = 0; q < kMaxQ; ++q) r[q] = (q * 32 + lane < D) ? x[q * 32 + lane] : 0.0;
}
__device__ __forceinline__ void tri_store(double* x, int D, const double (&r)[kMaxQ]) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int q = 0; q < kMaxQ; ++q)
    if (q * 32 + lane < D) x[q * 32 + lane] = r[q];
}

// x <- (L L^T)^-1 x by warp 0 alone (the other warps wait at the closing barrier)
__device__ void chol_solve(const double* L, const double* rd, int D, int ld, double* x) {
  if (threadIdx.x < 32) {
    double r[kMaxQ];
    tri_load(x, D, r);
    tri_forward(L, rd, D, ld, r);
    tri_backward(L, rd, D, ld, r);
    tri_store(x, D, r);
  }
  __syncthreads();
}

// out (D x D, row-major, global) = L^-1 for the lower-triangular L in shared memory: one warp per column c solves
// L x = e_c by column-oriented forward substitution with the residuals of rows 32 q + lane in registers.
__device__ void tri_inverse_lower(const double* L, const double* rd, int D, int ld, double* out) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int c = wid; c < D; c += nw) {
    double r[kMaxQ];
#pragma unroll
    for (int q = 0; q < kMaxQ; ++q) r[q] = (q * 32 + lane == c) ? 1.0 : 0.0;
#pragma unroll
    for (int q0 = 0; q0 < kMaxQ; ++q0) {
      if (q0 * 32 + 31 < c || q0 * 32 >= D) continue;
      for (int l = 0; l < 32; ++l) {
        const int i = q0 * 32 + l;
        if (i >= D) break;
        const double xi = __shfl_sync(0xffffffffu, r[q0], l) * rd[i];
        if (i < c) {
          if (lane == 0) out[i * D + c] = 0.0;
          continue;
        }
        if (lane == 0) out[i * D + c] = xi;
#pragma unroll
        for (int q = 0; q < kMaxQ; ++q) {
          const int k = q * 32 + lane;
          if (q >= q0 && k > i && k < D) r[q] = fma(-L[k * ld + i], xi, r[q]);
        }
      }
    }
    for (int i = lane; i < c && i < (c & ~31); i += 32) out[i * D + c] = 0.0;   // rows above the first visited block
  }
  __syncthreads();
}

// margins m_i = -z_i . th, one warp per row (coalesced)
__device__ void margins(const double* Z, long long ldz, int M, int D, const double* th, double* mrg) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int i = wid; i < M; i += nw) {
    double a = 0.0;
    for (int k = lane; k < D; k += 32) a = fma(Z[i * ldz + k], th[k], a);
    a = warp_sum(a);
    if (lane == 0) mrg[i] = -a;
  }
  __syncthreads();
}

// weighted log-joint (model_lr.py:88-93) from the margins at th: sum_i w_i ll(m_i) - D/2 log 2pi - |th|^2 / 2
__device__ double log_joint(const double* w, const double* mrg, int M, int D, const double* th, double* red) {
  double part = 0.0;
  for (int i = threadIdx.x; i < M; i += blockDim.x) {
    const double m = mrg[i];
    const double ll = (m < 100.0) ? -log1p(exp(m)) : -m;
    part = fma(w[i], ll, part);
  }
  for (int k = threadIdx.x; k < D; k += blockDim.x) part -= 0.5 * th[k] * th[k];
  return blk_sum(part, red) - 0.5 * (double)D * 1.8378770664093453;   // log(2 pi)
}


#if defined(BC_LAP_TRACE)   // phase stamps for tools/laplace_parts.py (tools/build_variants.py laptrace="-DBC_LAP_TRACE")
__device__ long long g_lap_trace[8];
extern "C" int bc_lap_trace_read(long long* out) { return (int)cudaMemcpyFromSymbol(out, g_lap_trace, sizeof(g_lap_trace)); }
#define LAP_TR(i) do { __syncthreads(); if (threadIdx.x == 0) g_lap_trace[i] = clock64(); } while (0)
#else
#define LAP_TR(i) do { } while (0)
#endif
__global__ void __launch_bounds__(kLapThreads) k_laplace_logistic(const double* __restrict__ Zg, long long ldzg, const double* __restrict__ w,
                                                                  int M, int D, double* __restrict__ mu_io, double* __restrict__ Lsig,
                                                                  int maxit, double tol, int* __restrict__ info, int flags, int stage_rows) {
  // flags bit 0: write the lower Cholesky factor C of the negative Hessian itself (get_laplace's LSigInv) instead of its inverse
  //       bit 1: Newton steps from the M x M dual system (Woodbury; needs M <= D): H = I + Z^T diag(d) Z is a rank-M update of
  //              the identity, H^-1 g = g - Z^T r (I + r K r)^-1 r Z g with r = sqrt(d), K = Z Z^T -- the same step as
  //              examples/common/model_lr.py::_newton_mode takes while the coreset is small, at M^3 instead of D^3 per iteration
  extern __shared__ double sm[];
  const int ld = D + 1, Mp = (M + 1) & ~1;
  const bool dual = (flags & 2) && M <= D;
  double* H = sm;                 // [D*ld]  (the dual system A, M x (M+1), lives here during the mode search)
  double* th = H + D * ld;        // [D]
  double* tn = th + D;            // [D] trial point
  double* g = tn + D;             // [D] gradient, then Newton step
  double* sw = g + D;             // [Mp] w_i s_i, then w_i c_i
  double* mrg = sw + Mp;          // [Mp] margins at th
  double* mtr = mrg + Mp;         // [Mp] margins at the trial point
  double* red = mtr + Mp;         // [64]
  double* rd = red + 64;          // [D] reciprocal diagonal of the Cholesky factor
  double* rr = rd + D;            // [Mp] r_i = sqrt(w_i c_i)            (dual steps)
  double* rhs = rr + Mp;          // [Mp] r_i z_i.g, then the dual solution (dual steps)
  __shared__ int flag;
  const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, wid = tid >> 5, nw = nt >> 5;
  // stage_rows: the M coreset rows are read a dozen times (margins, gradient, Hessian or dual system per Newton step, the
  // final Hessian) by loops whose trip count is M: from global memory every trip is an L2 round trip.  When they fit beside
  // the D x D workspace they are copied into shared memory once.
  LAP_TR(0);
  const double* Z = Zg;
  long long ldz = ldzg;
  if (stage_rows) {
    double* Zs = rhs + Mp;        // [M][D]
    for (int q = tid; q < M * D; q += nt) {
      const int i = q / D, kk = q - i * D;
      Zs[q] = Zg[i * ldzg + kk];
    }
    Z = Zs;
    ldz = D;
  }
  for (int k = tid; k < D; k += nt) th[k] = mu_io[k];
  __syncthreads();
  margins(Z, ldz, M, D, th, mrg);
  double f = log_joint(w, mrg, M, D, th, red);
  LAP_TR(1);
  int it = 0, status = 0;
  auto hessian = [&]() {   // H = I + Z^T diag(w c) Z (lower triangle) from the margins at th
    for (int i = tid; i < M; i += nt) {
      const double m = mrg[i];
      double c = 0.0;
      if (m < 100.0) {
        const double e = exp(m), s = e / (1.0 + e);
        c = s * (1.0 - s);
      }
      sw[i] = w[i] * c;
    }
    __syncthreads();
    for (int a = wid; a < D; a += nw) {
      for (int b = lane; b <= a; b += 32) {
        double a0 = (a == b) ? 1.0 : 0.0, a1 = 0.0;
        int i = 0;
        for (; i + 1 < M; i += 2) {
          a0 = fma(sw[i] * Z[i * ldz + a], Z[i * ldz + b], a0);
          a1 = fma(sw[i + 1] * Z[(i + 1) * ldz + a], Z[(i + 1) * ldz + b], a1);
        }
        if (i < M) a0 = fma(sw[i] * Z[i * ldz + a], Z[i * ldz + b], a0);
        H[a * ld + b] = a0 + a1;
      }
    }
    __syncthreads();
  };
  for (; it < maxit; ++it) {
    // gradient -th + Z^T (w s)
    for (int i = tid; i < M; i += nt) {
      const double m = mrg[i];
      double s = 1.0;
      if (m < 100.0) {
        const double e = exp(m);
        s = e / (1.0 + e);
      }
      sw[i] = w[i] * s;
    }
    __syncthreads();
    for (int k = tid; k < D; k += nt) {
      double acc = -th[k];
      for (int i = 0; i < M; ++i) acc = fma(sw[i], Z[i * ldz + k], acc);
      g[k] = acc;
    }
    __syncthreads();
    if (dual) {
      const int lda = M + 1;
      for (int i = tid; i < M; i += nt) {
        const double m = mrg[i];
        double c = 0.0;
        if (m < 100.0) {
          const double e = exp(m), sg = e / (1.0 + e);
          c = sg * (1.0 - sg);
        }
        rr[i] = sqrt(w[i] * c);
      }
      __syncthreads();
      // A = I + r_i r_j z_i.z_j (lower triangle), one warp per entry; rhs_i = r_i z_i.g
      const int npairs = M * (M + 1) / 2;
      for (int p = wid; p < npairs + M; p += nw) {
        if (p < npairs) {
          int i = (int)((sqrt(8.0 * (double)p + 1.0) - 1.0) * 0.5);
          while (i * (i + 1) / 2 > p) --i;
          while ((i + 1) * (i + 2) / 2 <= p) ++i;
          const int j = p - i * (i + 1) / 2;
          double a = 0.0;
          for (int k = lane; k < D; k += 32) a = fma(Z[i * ldz + k], Z[j * ldz + k], a);
          a = warp_sum(a);
          if (lane == 0) H[i * lda + j] = ((i == j) ? 1.0 : 0.0) + rr[i] * rr[j] * a;
        } else {
          const int i = p - npairs;
          double a = 0.0;
          for (int k = lane; k < D; k += 32) a = fma(Z[i * ldz + k], g[k], a);
          a = warp_sum(a);
          if (lane == 0) rhs[i] = rr[i] * a;
        }
      }
      __syncthreads();
      if (!chol_lower(H, M, lda, rd, &flag)) {
        status = 2;
        break;
      }
      chol_solve(H, rd, M, lda, rhs);
      for (int i = tid; i < M; i += nt) rhs[i] *= rr[i];
      __syncthreads();
      for (int k = tid; k < D; k += nt) {
        double acc = g[k];
        for (int i = 0; i < M; ++i) acc = fma(-rhs[i], Z[i * ldz + k], acc);
        g[k] = acc;                // g <- Newton step
      }
      __syncthreads();
    } else {
      hessian();
      if (!chol_lower(H, D, ld, rd, &flag)) {
        status = 2;
        break;
      }
      chol_solve(H, rd, D, ld, g);   // g <- Newton step
    }
    // damped step: halve until the log-joint does not decrease (it is strictly concave)
    double t = 1.0, fn = f;
    for (;;) {
      for (int k = tid; k < D; k += nt) tn[k] = fma(t, g[k], th[k]);
      __syncthreads();
      margins(Z, ldz, M, D, tn, mtr);
      fn = log_joint(w, mtr, M, D, tn, red);
      if (fn >= f - 1e-15 * fabs(f) || t < 1e-8) break;
      t *= 0.5;
    }
    double dm = 0.0, am = 0.0;
    for (int k = tid; k < D; k += nt) {
      dm = fmax(dm, fabs(tn[k] - th[k]));
      am = fmax(am, fabs(tn[k]));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      dm = fmax(dm, __shfl_xor_sync(0xffffffffu, dm, o));
      am = fmax(am, __shfl_xor_sync(0xffffffffu, am, o));
    }
    __syncthreads();
    if (lane == 0) {
      red[wid] = dm;
      red[32 + wid] = am;
    }
    __syncthreads();
    dm = 0.0;
    am = 0.0;
    for (int q = 0; q < nw; ++q) {
      dm = fmax(dm, red[q]);
      am = fmax(am, red[32 + q]);
    }
    __syncthreads();
    for (int k = tid; k < D; k += nt) th[k] = tn[k];
    for (int i = tid; i < M; i += nt) mrg[i] = mtr[i];
    f = fn;
    __syncthreads();
    // a full step this small is in the quadratic regime: the point it lands on is within ~|z| dm^2 of the mode
    if (dm <= tol * (1.0 + am) || (t == 1.0 && dm <= 1e-7 * (1.0 + am))) {
      ++it;
      break;
    }
  }
  LAP_TR(2);
  if (status == 0) {
    hessian();
    LAP_TR(3);
    if (!chol_lower(H, D, ld, rd, &flag)) status = 2;
    LAP_TR(4);
  }
  if (status == 0) {
    if (flags & 1) {
      for (int q = tid; q < D * D; q += nt) {
        const int i = q / D, j = q - i * D;
        Lsig[q] = (j <= i) ? H[i * ld + j] : 0.0;   // get_laplace's LSigInv
      }
    } else {
      tri_inverse_lower(H, rd, D, ld, Lsig);   // get_laplace's LSig
    }
    for (int k = tid; k < D; k += nt) mu_io[k] = th[k];
  }
  LAP_TR(5);
  if (tid == 0) {
    info[0] = status;
    info[1] = it;
  }
}

// Conjugate weighted posteriors on the device (the host-free optimiser loop): precision H, its lower Cholesky factor C and
// the mean AS THE REFERENCE COMPUTES IT, mu = LSigp LSigp^T v with LSigp = C^-1, i.e. C^-1 (C^-T v) (gaussian.py:28-32,
// model_neurlinr.py:115-122 -- the Gram matrix of the inverse FACTOR, not H^-1).
//   model 1 (Gaussian mean, known covariance): H = A0 + (sum_i w_i) A1,            v = v0 + A1 (sum_i w_i x_i)
//   model 2 (neural-linear head):              H = A0 + X^T diag(w) X / sigsq,     v = v0 + X^T (w y) / sigsq   (rows z = [x, y])
// with A0 = Sig0inv, A1 = Siginv, v0 = Sig0inv mu0.  One CTA.
__global__ void __launch_bounds__(kLapThreads) k_conjugate_factor(int model, const double* __restrict__ Z, long long ldz,
                                                                  const double* __restrict__ w, int M, int D, const double* __restrict__ A0,
                                                                  const double* __restrict__ A1, const double* __restrict__ v0, double sigsq,
                                                                  double* __restrict__ mu_out, double* __restrict__ C_out, int* __restrict__ info) {
  extern __shared__ double sm[];
  const int ld = D + 1;
  double* H = sm;            // [D*ld]
  double* v = H + D * ld;    // [D]
  double* xs = v + D;        // [D] weighted row sum
  double* rd = xs + D;       // [D]
  double* red = rd + D;      // [64]
  __shared__ int flag;
  const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, wid = tid >> 5, nw = nt >> 5;
  if (model == 1) {
    double part = 0.0;
    for (int i = tid; i < M; i += nt) part += w[i];
    const double sw = blk_sum(part, red);
    for (int k = tid; k < D; k += nt) {
      double acc = 0.0;
      for (int i = 0; i < M; ++i) acc = fma(w[i], Z[i * ldz + k], acc);
      xs[k] = acc;
    }
    __syncthreads();
    for (int a = wid; a < D; a += nw)
      for (int b = lane; b <= a; b += 32) H[a * ld + b] = fma(sw, A1[a * D + b], A0[a * D + b]);
    for (int a = wid; a < D; a += nw) {
      double acc = 0.0;
      for (int k = lane; k < D; k += 32) acc = fma(A1[a * D + k], xs[k], acc);
      acc = warp_sum(acc);
      if (lane == 0) v[a] = v0[a] + acc;
    }
  } else {
    const double is2 = 1.0 / sigsq;
    for (int a = wid; a < D; a += nw) {
      for (int b = lane; b <= a; b += 32) {
        double acc = 0.0;
        for (int i = 0; i < M; ++i) acc = fma(w[i] * Z[i * ldz + a], Z[i * ldz + b], acc);
        H[a * ld + b] = fma(acc, is2, A0[a * D + b]);
      }
    }
    for (int k = tid; k < D; k += nt) {
      double acc = 0.0;
      for (int i = 0; i < M; ++i) acc = fma(w[i] * Z[i * ldz + D], Z[i * ldz + k], acc);
      v[k] = fma(acc, is2, v0[k]);
    }
  }
  // Structure check: a precision whose off-diagonal entries are all exactly zero (isotropic or diagonal prior and noise
  // covariances: the reference's Gaussian example, Sig0 = I, Sig = 500 I) factorises entry by entry.  The general algorithm
  // below computes exactly that -- every off-diagonal operation adds a signed zero -- so the shortcut changes no bit, only
  // the time (D sequential pivots of a sqrt and a reciprocal each are most of this kernel).
  __syncthreads();
  int offdiag = 0;
  for (int a = wid; a < D; a += nw)
    for (int b = lane; b < a; b += 32) offdiag |= (H[a * ld + b] != 0.0) ? 1 : 0;
  const int dense = __syncthreads_or(offdiag);
  if (!dense) {
    int bad = 0;
    for (int k = tid; k < D; k += nt) {
      const double d = H[k * ld + k];
      if (!(d > 0.0)) bad = 1;
      const double sq = sqrt(d), inv = 1.0 / sq;
      H[k * ld + k] = sq;
      rd[k] = inv;
      v[k] = (v[k] * inv) * inv;           // C^T y = v, then C mu = y
    }
    if (__syncthreads_or(bad)) {
      if (tid == 0) info[0] = 2;
      return;
    }
  } else {
    if (!chol_lower(H, D, ld, rd, &flag)) {
      if (tid == 0) info[0] = 2;
      return;
    }
    if (tid < 32) {
      double r[kMaxQ];
      tri_load(v, D, r);
      tri_backward(H, rd, D, ld, r);         // C^T y = v
      tri_forward(H, rd, D, ld, r);          // C mu = y
      tri_store(v, D, r);
    }
    __syncthreads();
  }
  for (int k = tid; k < D; k += nt) mu_out[k] = v[k];
  for (int i = wid; i < D; i += nw)
    for (int j = lane; j < D; j += 32) C_out[i * D + j] = (j <= i) ? H[i * ld + j] : 0.0;
  if (tid == 0) {
    info[0] = 0;
    info[1] = dense ? 0 : 1;               // 1: the factor is diagonal (bc_sample_solve_hinted skips the substitution)
  }
}

cudaError_t launch_conjugate_factor(int model, const double* Z, long long ldz, const double* w, int M, int D, const double* A0, const double* A1,
                                    const double* v0, double sigsq, double* mu, double* C, int* info, cudaStream_t st) {
  const size_t smem = ((size_t)D * (D + 1) + 3 * (size_t)D + 64) * sizeof(double);
  const size_t cap = kMaxSmem - 1024;
  static DeviceOnce once;
  cudaError_t e = raise_dynamic_smem(k_conjugate_factor, cap, once);
  if (e != cudaSuccess) return e;
  if (smem > cap) return cudaErrorInvalidValue;
  BC_PREFER_MAX_SHARED(k_conjugate_factor);
  k_conjugate_factor<<<1, kLapThreads, smem, st>>>(model, Z, ldz, w, M, D, A0, A1, v0, sigsq, mu, C, info);
  return cudaGetLastError();
}

// Theta[s][d] = mu[d] + sum_k R[s][k] L[d][k]      (mu + randn(S, D).dot(L.T))
__global__ void __launch_bounds__(128) k_sample_affine(const double* __restrict__ mu, const double* __restrict__ L, const double* __restrict__ R,
                                                       int S, int D, double* __restrict__ out, int ldo) {
  extern __shared__ double rs[];   // this sample's normals
  const int s = blockIdx.x;
  for (int k = threadIdx.x; k < D; k += blockDim.x) rs[k] = R[(size_t)s * D + k];
  __syncthreads();
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    double acc = 0.0;
    for (int k = 0; k <= d; ++k) acc = fma(rs[k], L[(size_t)d * D + k], acc);   // L is lower triangular
    out[(size_t)s * ldo + d] = mu[d] + acc;
  }
}

// Theta[s][:] = mu + C^-1 R[s][:]  for the LOWER Cholesky factor C of the negative Hessian: the same samples as
// `mu + randn(S, D).dot(LSig.T)` with LSig = C^-1 (get_laplace, util/opt.py:27-33), without forming the inverse -- the
// host then only factors (dpotrf) and never inverts (dtrtri is the slower of the two at D = 128).  One WARP per sample:
// column-oriented forward substitution with the right-hand side in registers (tri_forward); the CTA first stages the lower
// triangle of C into shared memory (row-major, odd leading dimension: walks down a column are conflict-free) with the
// reciprocal pivots.  A step costs a shuffle, a multiply and one FMA per owned entry: 2 us per sample at D = 128, all
// samples in parallel (the thread-per-sample form it replaces took 225 us at D = 100).
constexpr int kSolveWarps = 16;
__global__ void __launch_bounds__(kSolveWarps * 32) k_sample_solve(const double* __restrict__ mu, const double* __restrict__ C,
                                                                   const double* __restrict__ R, int S, int D, double* __restrict__ out,
                                                                   int ldo, const int* __restrict__ diag_hint) {
  extern __shared__ double sm[];
  const int ld = D | 1;
  double* L = sm;              // [D][ld]
  double* rd = L + D * ld;     // [D]
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nw = blockDim.x >> 5;
  // diag_hint (bc_conjugate_factor's info[1]): the factor is diagonal -- the substitution reduces to its first operation per
  // entry, x_j = r_j / C_jj (every other one adds a signed zero), and the triangle need not be staged at all
  const bool diag = diag_hint && *diag_hint == 1;
  if (!diag) {
    for (int i = wid; i < D; i += nw)
      for (int j = lane; j <= i; j += 32) L[i * ld + j] = __ldg(C + (size_t)i * D + j);
  }
  for (int i = tid; i < D; i += blockDim.x) rd[i] = 1.0 / __ldg(C + (size_t)i * D + i);
  __syncthreads();
  const int s = blockIdx.x * kSolveWarps + wid;
  if (s >= S) return;
  double r[kMaxQ];
  tri_load(R + (size_t)s * D, D, r);
  if (diag) {
#pragma unroll
    for (int q = 0; q < kMaxQ; ++q)
      if (q * 32 + lane < D) r[q] *= rd[q * 32 + lane];
  } else {
    tri_forward(L, rd, D, ld, r);
  }
#pragma unroll
  for (int q = 0; q < kMaxQ; ++q)
    if (q * 32 + lane < D) out[(size_t)s * ldo + q * 32 + lane] = __ldg(mu + q * 32 + lane) + r[q];
}

cudaError_t launch_sample_solve(const double* mu, const double* C, const double* R, int S, int D, double* out, int ldo, const int* diag_hint,
                                cudaStream_t st) {
  if (S <= 0) return cudaSuccess;
  const size_t smem = ((size_t)D * (D | 1) + D) * sizeof(double);
  const size_t cap = kMaxSmem - 1024;
  static DeviceOnce once;
  cudaError_t e = raise_dynamic_smem(k_sample_solve, cap, once);
  if (e != cudaSuccess) return e;
  if (smem > cap) return cudaErrorInvalidValue;
  BC_PREFER_MAX_SHARED(k_sample_solve);
  k_sample_solve<<<(S + kSolveWarps - 1) / kSolveWarps, kSolveWarps * 32, smem, st>>>(mu, C, R, S, D, out, ldo, diag_hint);
  return cudaGetLastError();
}

// ------------------------------------------------ non-negative least squares --
// min |A x - b|, x >= 0 for the S x m block of cached datapoint rows (A's column j = row pos[j] of the cache):
// OrthoPursuit._reweight / SparseNNLS.optimize (orthopursuit.py:37-42, snnls.py:82-97), where the reference calls
// scipy.optimize.nnls.  Lawson-Hanson active-set iteration, one CTA:
//   * the least-squares problems on the passive set P are solved through the Cholesky factor of the Gram block G_PP
//     (G = A^T A, m x m, computed once), appended to row by row as columns enter P and rebuilt when columns leave;
//   * every solve is refined twice by the corrected semi-normal equations -- residual rho = b - A_P s formed in the data
//     space, G_PP d = A_P^T rho, s += d -- which brings the normal-equation solve to the accuracy of a QR-based one
//     (Bjorck) for the conditioning coreset columns have;
//   * the dual w = A^T (b - A x) is likewise formed from the residual, not from G;
//   * a WARM start is taken from x0 (the current weights: P starts as their support), so the usual call -- one new column
//     against an already optimal support -- takes one or two outer iterations instead of m.
// The minimiser is unique for independent columns, so the result agrees with scipy's to rounding (tests); info[0] = 0 ok,
// 1 iteration cap, 2 factorisation failed (numerically dependent columns: the caller falls back to the host routine).
__device__ void tri_forward_only(const double* L, const double* rd, int D, int ld, double* x) {   // L y = x by warp 0
  if (threadIdx.x < 32) {
    double r[kMaxQ];
    tri_load(x, D, r);
    tri_forward(L, rd, D, ld, r);
    tri_store(x, D, r);
  }
  __syncthreads();
}

__global__ void __launch_bounds__(kLapThreads) k_nnls_gram(const double* __restrict__ Vact, int S, const long long* __restrict__ pos, int m,
                                                           const double* __restrict__ b, const double* __restrict__ x0,
                                                           double* __restrict__ x_out, double* __restrict__ rho /* S scratch */,
                                                           int maxit, int* __restrict__ info) {
  extern __shared__ double sm[];
  const int ldg = m | 1;
  double* G = sm;                  // [m][ldg]
  double* L = G + (size_t)m * ldg; // [m][ldg]  factor of G_PP, rows / columns in the order of plist
  double* c = L + (size_t)m * ldg; // [m] A^T b
  double* x = c + m;               // [m]
  double* sv = x + m;              // [m] LS solution, P order
  double* w = sv + m;              // [m] dual
  double* gv = w + m;              // [m] right-hand sides, P order
  double* rd = gv + m;             // [m]
  double* red = rd + m;            // [64]
  int* plist = reinterpret_cast<int*>(red + 64);   // [m]
  int* inP = plist + m;                            // [m]
  __shared__ int flag, s_p, s_k, s_it;
  __shared__ double s_val;
  const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, wid = tid >> 5, nw = nt >> 5;

  // Gram matrix and A^T b: one warp per entry of the lower triangle
  const int npairs = m * (m + 1) / 2;
  for (int q = wid; q < npairs + m; q += nw) {
    int i, j;
    const double* vj;
    if (q < npairs) {
      i = (int)((sqrt(8.0 * (double)q + 1.0) - 1.0) * 0.5);
      while (i * (i + 1) / 2 > q) --i;
      while ((i + 1) * (i + 2) / 2 <= q) ++i;
      j = q - i * (i + 1) / 2;
      vj = Vact + (size_t)pos[j] * S;
    } else {
      i = q - npairs;
      j = -1;
      vj = b;
    }
    const double* vi = Vact + (size_t)pos[i] * S;
    double a0 = 0.0, a1 = 0.0;
    int t = lane;
    for (; t + 32 < S; t += 64) {
      a0 = fma(vi[t], vj[t], a0);
      a1 = fma(vi[t + 32], vj[t + 32], a1);
    }
    if (t < S) a0 = fma(vi[t], vj[t], a0);
    const double d = warp_sum(a0 + a1);
    if (lane == 0) {
      if (j >= 0) {
        G[i * ldg + j] = d;
        G[j * ldg + i] = d;
      } else {
        c[i] = d;
      }
    }
  }
  for (int j = tid; j < m; j += nt) {
    const double v = x0 ? x0[j] : 0.0;
    x[j] = (v > 0.0) ? v : 0.0;
    inP[j] = (v > 0.0) ? 1 : 0;
  }
  if (tid == 0) {
    int p = 0;
    for (int j = 0; j < m; ++j)
      if (x0 && x0[j] > 0.0) plist[p++] = j;
    s_p = p;
    s_it = 0;
  }
  __syncthreads();
  double cmax = 0.0;
  for (int j = 0; j < m; ++j) cmax = fmax(cmax, fabs(c[j]));
  const double tol = 10.0 * (double)max(S, m) * 2.220446049250313e-16 * cmax;

  // residual rho = b - sum_{a in P} val_a * column(plist[a]) into the global scratch
  auto residual = [&](const double* val_P, int p) {
    for (int t = tid; t < S; t += nt) {
      double acc = b[t];
      for (int a = 0; a < p; ++a) acc = fma(-val_P[a], Vact[(size_t)pos[plist[a]] * S + t], acc);
      rho[t] = acc;
    }
    __syncthreads();
  };
  // out[q] = column(cols ? cols[q] : q) . rho for q < n
  auto project = [&](double* out, const int* cols, int n) {
    for (int q = wid; q < n; q += nw) {
      const double* v = Vact + (size_t)pos[cols ? cols[q] : q] * S;
      double a0 = 0.0;
      for (int t = lane; t < S; t += 32) a0 = fma(v[t], rho[t], a0);
      a0 = warp_sum(a0);
      if (lane == 0) out[q] = a0;
    }
    __syncthreads();
  };
  auto rebuild = [&](int p) -> bool {    // L = chol(G_PP) from scratch
    for (int q = tid; q < p * p; q += nt) {
      const int a = q / p, bb = q - a * p;
      if (bb <= a) L[a * ldg + bb] = G[plist[a] * ldg + plist[bb]];
    }
    __syncthreads();
    if (p == 0) return true;
    return chol_lower(L, p, ldg, rd, &flag);
  };
  auto solve = [&](int p) {              // sv = argmin |A_P s - b| with two CSNE refinements
    for (int a = tid; a < p; a += nt) sv[a] = c[plist[a]];
    __syncthreads();
    chol_solve(L, rd, p, ldg, sv);
    for (int r = 0; r < 2; ++r) {
      residual(sv, p);
      project(gv, plist, p);
      chol_solve(L, rd, p, ldg, gv);
      for (int a = tid; a < p; a += nt) sv[a] += gv[a];
      __syncthreads();
    }
  };

  int status = 0;
  if (!rebuild(s_p)) status = 2;
  bool warm = s_p > 0;
  while (status == 0) {
    int p = s_p;
    if (!warm) {
      // dual of the current x; stop when no inactive column can still reduce the residual
      for (int a = tid; a < p; a += nt) gv[a] = x[plist[a]];
      __syncthreads();
      residual(gv, p);
      project(w, nullptr, m);
      if (tid == 0) {
        int k = -1;
        double best = tol;
        for (int j = 0; j < m; ++j)
          if (!inP[j] && w[j] > best) {
            best = w[j];
            k = j;
          }
        s_k = k;
      }
      __syncthreads();
      const int k = s_k;
      if (k < 0 || p == m) break;
      // append column k to the factor: L l = G[P, k], d^2 = G_kk - |l|^2
      for (int a = tid; a < p; a += nt) gv[a] = G[plist[a] * ldg + k];
      __syncthreads();
      if (p > 0) tri_forward_only(L, rd, p, ldg, gv);
      if (tid == 0) {
        double d2 = G[k * ldg + k];
        for (int a = 0; a < p; ++a) d2 -= gv[a] * gv[a];
        s_val = d2;
      }
      __syncthreads();
      const double d2 = s_val;
      if (!(d2 > 1e-13 * G[k * ldg + k])) {       // numerically dependent on the passive columns: cannot enter (LH's test)
        if (tid == 0) inP[k] = 2;                  // excluded from the candidates of this call
        __syncthreads();
        continue;
      }
      for (int a = tid; a < p; a += nt) L[p * ldg + a] = gv[a];
      if (tid == 0) {
        const double d = sqrt(d2);
        L[p * ldg + p] = d;
        rd[p] = 1.0 / d;
        plist[p] = k;
        inP[k] = 1;
        s_p = p + 1;
      }
      __syncthreads();
      p += 1;
    }
    // inner loop: bring the least-squares solution on P back into the positive orthant
    bool fresh = !warm;                  // the last entry of plist has just entered with x = 0
    for (;;) {
      solve(p);
      if (tid == 0) {
        int neg = 0;
        for (int a = 0; a < p; ++a)
          if (!(sv[a] > 0.0)) neg = 1;
        s_k = neg;
        s_it += 1;
      }
      __syncthreads();
      if (fresh && !(sv[p - 1] > 0.0)) {
        // the column that just entered would leave again at once (Lawson-Hanson's sign test on the new coordinate):
        // take it out, bar it for this round, and look for another one
        if (tid == 0) {
          inP[plist[p - 1]] = 2;
          s_p = p - 1;
        }
        __syncthreads();
        break;
      }
      fresh = false;
      if (!s_k) {
        for (int j = tid; j < m; j += nt) {
          x[j] = 0.0;
          if (inP[j] == 2) inP[j] = 0;   // x moves: the barred columns may be tried again
        }
        __syncthreads();
        for (int a = tid; a < p; a += nt) x[plist[a]] = sv[a];
        __syncthreads();
        break;
      }
      if (s_it > maxit) {
        status = 1;
        break;
      }
      if (tid == 0) {
        // step from x towards s until the first coordinate reaches zero; the coordinates that block the step leave P
        double alpha = 1.0;
        for (int a = 0; a < p; ++a)
          if (!(sv[a] > 0.0)) {
            const double xa = x[plist[a]], den = xa - sv[a];
            const double r = (den > 0.0) ? xa / den : 0.0;
            if (r < alpha) alpha = r;
          }
        int q = 0;
        for (int a = 0; a < p; ++a) {
          const int j = plist[a];
          const double xa = x[j], den = xa - sv[a];
          const bool blocks = !(sv[a] > 0.0) && (((den > 0.0) ? xa / den : 0.0) <= alpha);
          const double xn = xa + alpha * (sv[a] - xa);
          if (!blocks && xn > 0.0) {
            x[j] = xn;
            plist[q++] = j;
          } else {
            x[j] = 0.0;
            inP[j] = 0;
          }
        }
        s_p = q;
      }
      __syncthreads();
      p = s_p;
      if (!rebuild(p)) {
        status = 2;
        break;
      }
      if (p == 0) break;
    }
    warm = false;
  }
  for (int j = tid; j < m; j += nt) x_out[j] = x[j];
  if (tid == 0) {
    info[0] = status;
    info[1] = s_it;
  }
}

cudaError_t launch_nnls_gram(const double* Vact, int S, const long long* pos, int m, const double* b, const double* x0, double* x_out,
                             double* rho, int maxit, int* info, cudaStream_t st) {
  const size_t smem = (2 * (size_t)m * (m | 1) + 6 * (size_t)m + 64) * sizeof(double) + 2 * (size_t)m * sizeof(int);
  const size_t cap = kMaxSmem - 1024;
  static DeviceOnce once;
  cudaError_t e = raise_dynamic_smem(k_nnls_gram, cap, once);
  if (e != cudaSuccess) return e;
  if (smem > cap) return cudaErrorInvalidValue;
  k_nnls_gram<<<1, kLapThreads, smem, st>>>(Vact, S, pos, m, b, x0, x_out, rho, maxit, info);
  return cudaGetLastError();
}

cudaError_t launch_laplace_logistic(const double* Z, long long ldz, const double* w, int M, int D, double* mu_io, double* Lsig, int maxit,
                                    double tol, int* info, int flags, cudaStream_t st) {
  size_t smem = ((size_t)D * (D + 1) + 4 * (size_t)D + 5 * (size_t)((M + 1) & ~1) + 64) * sizeof(double);
  const size_t cap = kMaxSmem - 1024;   // the kernel also has a few bytes of static shared memory
  static DeviceOnce once;
  cudaError_t e = raise_dynamic_smem(k_laplace_logistic, cap, once);
  if (e != cudaSuccess) return e;
  if (smem > cap) return cudaErrorInvalidValue;
  const size_t rows = (size_t)M * D * sizeof(double);
  const int stage_rows = smem + rows <= cap ? 1 : 0;
  if (stage_rows) smem += rows;
  BC_PREFER_MAX_SHARED(k_laplace_logistic);
  k_laplace_logistic<<<1, kLapThreads, smem, st>>>(Z, ldz, w, M, D, mu_io, Lsig, maxit, tol, info, flags, stage_rows);
  return cudaGetLastError();
}

cudaError_t launch_sample_affine(const double* mu, const double* L, const double* R, int S, int D, double* out, int ldo, cudaStream_t st) {
  if (S <= 0) return cudaSuccess;
  BC_PREFER_MAX_SHARED(k_sample_affine);
  k_sample_affine<<<S, 128, (size_t)D * sizeof(double), st>>>(mu, L, R, S, D, out, ldo);
  return cudaGetLastError();
}

}  // namespace bc
