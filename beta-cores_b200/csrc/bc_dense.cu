// beta-cores B200: stage 2 on a MATERIALISED n x S matrix V (row n = datapoint n's centred
// projection, contiguous) -- the layout HilbertCoreset hands to the snnls solvers
// (bayesiancoresets/coreset/hilbert.py:17: A = vecs.T) -- plus the S-length vector steps of
// GIGA / Frank-Wolfe / OrthoPursuit (bayesiancoresets/snnls/{giga,frankwolfe,orthopursuit}.py).
// The n x S passes are HBM-bound: one warp per row, 16-byte coalesced loads, the S-vector(s) in
// shared memory, warp-shuffle + block arg-max, two-phase grid reduction (no float atomics).
#include "bc_kernels.h"

namespace bc {

__device__ __forceinline__ double2 ld_stream2(const double* p) {
  double2 r;
  asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
  return r;
}

__device__ __forceinline__ double block_sum_d(double x, double* red) {
  x = warp_sum(x);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = blockDim.x >> 5;
  __syncthreads();
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

// per-warp row reduction: up to three accumulators (v.u0, v.u1, v.v)
template <int NU, bool NORM>
__device__ __forceinline__ void row_dots(const double* __restrict__ v, int S, const double* __restrict__ u_s, int lane, bool vec_ok,
                                         double& d0, double& d1, double& n2) {
  d0 = d1 = n2 = 0.0;
  if (vec_ok) {
    // S even and row 16-byte aligned.  Four 16-byte loads per lane are issued before the first is consumed: with one
    // warp per row the memory-level parallelism has to come from here (HBM latency x bandwidth needs ~40 KB in flight per SM)
    for (int c0 = lane * 2; c0 < S; c0 += 256) {
      double2 x[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int c = c0 + 64 * k;
        x[k] = (c < S) ? ld_stream2(v + c) : make_double2(0.0, 0.0);
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int c = c0 + 64 * k;
        if (c < S) {
          const double2 ua = *reinterpret_cast<const double2*>(u_s + c);
          d0 = fma(x[k].x, ua.x, d0);
          d0 = fma(x[k].y, ua.y, d0);
          if (NU > 1) {
            const double2 ub = *reinterpret_cast<const double2*>(u_s + S + c);
            d1 = fma(x[k].x, ub.x, d1);
            d1 = fma(x[k].y, ub.y, d1);
          }
          if (NORM) {
            n2 = fma(x[k].x, x[k].x, n2);
            n2 = fma(x[k].y, x[k].y, n2);
          }
        }
      }
    }
  } else {
    for (int c = lane; c < S; c += 32) {
      const double x = v[c];
      d0 = fma(x, u_s[c], d0);
      if (NU > 1) d1 = fma(x, u_s[S + c], d1);
      if (NORM) n2 = fma(x, x, n2);
    }
  }
  d0 = warp_sum(d0);
  if (NU > 1) d1 = warp_sum(d1);
  if (NORM) n2 = warp_sum(n2);
}

// ------------------------------------------------------------- row norms ----
// norms[r] = sqrt(sum_s V[r][s]^2)        giga.py:10, frankwolfe.py:10, hilbert.py:16
__global__ void k_dense_rownorms(const double* __restrict__ V, long long n, int S, long long ldv, double* __restrict__ norms) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  const bool vec_ok = (S % 2 == 0) && (ldv % 2 == 0) && ((reinterpret_cast<uintptr_t>(V) & 15) == 0);
  for (long long r = (long long)blockIdx.x * nw + warp; r < n; r += (long long)gridDim.x * nw) {
    const double* v = V + r * ldv;
    double n2 = 0.0;
    if (vec_ok) {
      for (int c = lane * 2; c < S; c += 64) {
        const double2 x = ld_stream2(v + c);
        n2 = fma(x.x, x.x, n2);
        n2 = fma(x.y, x.y, n2);
      }
    } else {
      for (int c = lane; c < S; c += 32) n2 = fma(v[c], v[c], n2);
    }
    n2 = warp_sum(n2);
    if (lane == 0) norms[r] = sqrt(n2);
  }
}

cudaError_t launch_dense_rowstats(const double* V, long long n, int S, long long ldv, const double*, int, double* norms, double*,
                                  cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  long long blocks = (n + 7) / 8;
  if (blocks > 148 * 8) blocks = 148 * 8;
  k_dense_rownorms<<<(int)blocks, 256, 0, st>>>(V, n, S, ldv, norms);
  return cudaGetLastError();
}

// ------------------------------------------------------------ column sums ---
// part[b][s] = sum over block b's rows; then dd-combine over blocks in fixed order.
__global__ void k_dense_colsum_part(const double* __restrict__ V, long long n, int S, long long ldv, double* __restrict__ part) {
  const long long rows_per = (n + gridDim.x - 1) / gridDim.x;
  const long long r0 = (long long)blockIdx.x * rows_per;
  long long r1 = r0 + rows_per;
  if (r1 > n) r1 = n;
  for (int s = threadIdx.x; s < S; s += blockDim.x) {
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    long long r = r0;
    for (; r + 3 < r1; r += 4) {
      a0 += V[r * ldv + s];
      a1 += V[(r + 1) * ldv + s];
      a2 += V[(r + 2) * ldv + s];
      a3 += V[(r + 3) * ldv + s];
    }
    for (; r < r1; ++r) a0 += V[r * ldv + s];
    part[(size_t)blockIdx.x * S + s] = (a0 + a1) + (a2 + a3);
  }
}
// one warp per column: the lanes add strided shares of the per-block partials in double-double, then merge -- the sum is
// order-insensitive at double-double precision (a single thread per column adding ~1200 partials in sequence cost 0.1-0.2 ms)
__global__ void __launch_bounds__(256) k_dense_colsum_fin(const double* __restrict__ part, int nparts, int S, int Sld,
                                                          double* __restrict__ out_dd) {
  const int lane = threadIdx.x & 31;
  const int s = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (s > S) return;
  dd a = {0.0, 0.0};
  if (s < S)
    for (int p = lane; p < nparts; p += 32) a = dd_add_d(a, part[(size_t)p * S + s]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    dd b;
    b.hi = __shfl_xor_sync(0xffffffffu, a.hi, o);
    b.lo = __shfl_xor_sync(0xffffffffu, a.lo, o);
    a = dd_add(a, b);
  }
  if (lane == 0) {
    out_dd[s] = a.hi;  // element S (sum of row means) is 0: V is already centred
    out_dd[Sld + s] = a.lo;
  }
}

cudaError_t launch_dense_colsum(const double* V, long long n, int S, long long ldv, double* part, int nparts, double* out_dd,
                                int Sld, cudaStream_t st) {
  if (n < nparts) nparts = n > 0 ? (int)n : 1;
  k_dense_colsum_part<<<nparts, 256, 0, st>>>(V, n, S, ldv, part);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  k_dense_colsum_fin<<<(S + 1 + 7) / 8, 256, 0, st>>>(part, nparts, S, Sld, out_dd);
  return cudaGetLastError();
}

// ------------------------------------------------------- score + arg-max ----
// mode 0 (FW)  : score = (V_r . u0) / norm_r                                   frankwolfe.py:15-17
// mode 1 (GIGA): s0 = V_r.u0/norm_r, s1 = V_r.u1/norm_r;
//                ok = s1 > -1+1e-14 and 1-s1^2 > 0; score = s0 / (ok ? sqrt(1-s1^2) : inf)   giga.py:31-38
// mode 2 (CORR): score = (V_r . u0) / sqrt(sum V_r^2) / S                      bcores.py:78 (black-box path)
// mode 3 (OMP) : as FW, plus best of (-score) over rows with active[r] != 0    orthopursuit.py:17-35
// part: [nblocks][4] = {best v, best i, neg-best v, neg-best i}
template <int MODE>
__global__ void __launch_bounds__(256) k_dense_score(const double* __restrict__ V, long long n, int S, long long ldv,
                                                     const double* __restrict__ norms, const double* __restrict__ u,
                                                     const unsigned char* __restrict__ active, long long idx_offset,
                                                     double* __restrict__ part, double* __restrict__ scores, const double* stop) {
  extern __shared__ double u_s[];  // NU * S
  if (stop && stop[1] != 0.0) return;   // a device-resident solver run has stopped (k_solver_step): nothing to score
  constexpr int NU = (MODE == 1) ? 2 : 1;
  __shared__ double bv[8][2];
  __shared__ long long bi[8][2];
  for (int i = threadIdx.x; i < NU * S; i += blockDim.x) u_s[i] = u[i];
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  const bool vec_ok = (S % 2 == 0) && (ldv % 2 == 0) && ((reinterpret_cast<uintptr_t>(V) & 15) == 0);
  Best best = {0.0, -1}, nbest = {0.0, -1};
  for (long long r = (long long)blockIdx.x * nw + warp; r < n; r += (long long)gridDim.x * nw) {
    double d0, d1, n2;
    row_dots<NU, MODE == 2>(V + r * ldv, S, u_s, lane, vec_ok, d0, d1, n2);
    double score;
    if (MODE == 2) {
      score = d0 / sqrt(n2) / (double)S;
    } else {
      const double nr = norms[r];
      const double s0 = d0 / nr;
      if (MODE == 1) {
        const double s1 = d1 / nr;
        const bool ok = (s1 > -1.0 + 1e-14) && (1.0 - s1 * s1 > 0.0);
        score = s0 / (ok ? sqrt(1.0 - s1 * s1) : INFINITY);
      } else {
        score = s0;
      }
    }
    if (lane == 0) {
      if (scores) scores[r] = score;
      Best c = {score, idx_offset + r};
      best = best_merge(best, c);
      if (MODE == 3 && active[r]) {
        Best cn = {-score, idx_offset + r};
        nbest = best_merge(nbest, cn);
      }
    }
  }
  if (lane == 0) {
    bv[warp][0] = best.v; bi[warp][0] = best.i;
    bv[warp][1] = nbest.v; bi[warp][1] = nbest.i;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < nw; ++w) {
      Best o = {bv[w][0], bi[w][0]}, on = {bv[w][1], bi[w][1]};
      best = best_merge(best, o);
      nbest = best_merge(nbest, on);
    }
    part[blockIdx.x * 4 + 0] = best.v;
    part[blockIdx.x * 4 + 1] = __longlong_as_double(best.i);
    part[blockIdx.x * 4 + 2] = nbest.v;
    part[blockIdx.x * 4 + 3] = __longlong_as_double(nbest.i);
  }
}
// merge of the per-CTA candidates: the order (NaN first, larger value, lower index) is total, so any merge tree gives the
// sequential result; one block, each thread takes a strided share (a single thread walking several hundred partials cost
// 0.1 ms per scoring pass)
__global__ void __launch_bounds__(256) k_dense_score_fin(const double* __restrict__ part, int nparts, double* __restrict__ out,
                                                         const double* stop) {
  __shared__ double sv[2][8];
  __shared__ long long si[2][8];
  if (stop && stop[1] != 0.0) return;
  Best b = {0.0, -1}, nb = {0.0, -1};
  for (int p = threadIdx.x; p < nparts; p += blockDim.x) {
    Best o = {part[p * 4 + 0], __double_as_longlong(part[p * 4 + 1])};
    Best on = {part[p * 4 + 2], __double_as_longlong(part[p * 4 + 3])};
    b = best_merge(b, o);
    nb = best_merge(nb, on);
  }
  b = best_warp(b);
  nb = best_warp(nb);
  const int w = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0) {
    sv[0][w] = b.v;
    si[0][w] = b.i;
    sv[1][w] = nb.v;
    si[1][w] = nb.i;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int k = 1; k < (int)(blockDim.x >> 5); ++k) {
      Best o = {sv[0][k], si[0][k]}, on = {sv[1][k], si[1][k]};
      b = best_merge(b, o);
      nb = best_merge(nb, on);
    }
    out[0] = b.v;
    out[1] = __longlong_as_double(b.i);
    out[2] = nb.v;
    out[3] = __longlong_as_double(nb.i);
  }
}

cudaError_t launch_dense_score(const double* V, long long n, int S, long long ldv, const double* norms, const double* u, int mode,
                               const unsigned char* active, long long idx_offset, double* part, int nparts, double* out,
                               double* scores, cudaStream_t st, const double* stop) {
  long long want = (n + 7) / 8;
  if (want < 1) want = 1;
  const size_t smem = (size_t)((mode == 1) ? 2 : 1) * S * sizeof(double);
  // one wave of resident blocks: the rows are dealt out grid-stride, so a partial second wave would only add a tail.
  // SM count, the opt-in shared-memory limit (S > 6144 samples need more than the default 48 KB) and the occupancy are
  // per device and per instantiation: looked up once each.
  constexpr size_t kScoreSmemMax = 200 * 1024;
  static DeviceOnce once[4];
  static int sms_of[64];
  const int dev = current_device();
  if (smem > kScoreSmemMax) return cudaErrorInvalidValue;
  int o = 0;
  {
    cudaError_t e;
    switch (mode) {
      case 0:
        e = raise_dynamic_smem(k_dense_score<0>, kScoreSmemMax, once[0]);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, k_dense_score<0>, 256, smem);
        break;
      case 1:
        e = raise_dynamic_smem(k_dense_score<1>, kScoreSmemMax, once[1]);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, k_dense_score<1>, 256, smem);
        break;
      case 2:
        e = raise_dynamic_smem(k_dense_score<2>, kScoreSmemMax, once[2]);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, k_dense_score<2>, 256, smem);
        break;
      default:
        e = raise_dynamic_smem(k_dense_score<3>, kScoreSmemMax, once[3]);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, k_dense_score<3>, 256, smem);
        break;
    }
    if (e != cudaSuccess) return e;
  }
  int sms = (dev >= 0 && dev < 64) ? sms_of[dev] : 0;
  if (!sms) {
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (dev >= 0 && dev < 64) sms_of[dev] = sms;
  }
  if (o < 1) o = 1;
  int resident = sms * o;
  if (resident < nparts) nparts = resident;
  if (want < nparts) nparts = (int)want;
  switch (mode) {
    case 0: k_dense_score<0><<<nparts, 256, smem, st>>>(V, n, S, ldv, norms, u, active, idx_offset, part, scores, stop); break;
    case 1: k_dense_score<1><<<nparts, 256, smem, st>>>(V, n, S, ldv, norms, u, active, idx_offset, part, scores, stop); break;
    case 2: k_dense_score<2><<<nparts, 256, smem, st>>>(V, n, S, ldv, norms, u, active, idx_offset, part, scores, stop); break;
    default: k_dense_score<3><<<nparts, 256, smem, st>>>(V, n, S, ldv, norms, u, active, idx_offset, part, scores, stop); break;
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  k_dense_score_fin<<<1, 256, 0, st>>>(part, nparts, out, stop);
  return cudaGetLastError();
}

// ------------------------------------------------------- small row helpers --
// out[s] = sum_m w[m] V[idx[m]][s]      (= A.dot(w) restricted to the selected columns; giga.py:21)
__global__ void k_dense_combine(const double* __restrict__ V, long long ldv, int S, const long long* __restrict__ idx,
                                const double* __restrict__ w, int m, double* __restrict__ out) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= S) return;
  double acc = 0.0;
  for (int i = 0; i < m; ++i) acc = fma(w[i], V[idx[i] * ldv + s], acc);
  out[s] = acc;
}
cudaError_t launch_dense_combine(const double* V, long long ldv, int S, const long long* idx, const double* w, int m, double* out,
                                 cudaStream_t st) {
  k_dense_combine<<<(S + 127) / 128, 128, 0, st>>>(V, ldv, S, idx, w, m, out);
  return cudaGetLastError();
}

// V[r][:] -= mean(V[r][:])      projector.py:26 / :55 for black-box (host callback) potentials
__global__ void k_dense_center(double* __restrict__ V, long long n, int S, long long ldv) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  for (long long r = (long long)blockIdx.x * nw + warp; r < n; r += (long long)gridDim.x * nw) {
    double* v = V + r * ldv;
    double a = 0.0;
    const double first = v[0];
    bool same = true;
    for (int c = lane; c < S; c += 32) {
      a += v[c];
      same = same && (v[c] == first);
    }
    a = warp_sum(a) / (double)S;
    // a row of S identical values is centred the way numpy rounds its mean (bc_common.cuh, np_sum_const)
    if (__all_sync(0xffffffffu, same)) a = np_sum_const(first, S) / (double)S;
    for (int c = lane; c < S; c += 32) v[c] -= a;
  }
}
cudaError_t launch_dense_center(double* V, long long n, int S, long long ldv, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  long long blocks = (n + 7) / 8;
  if (blocks > 148 * 8) blocks = 148 * 8;
  BC_PREFER_MAX_SHARED(k_dense_center);
  k_dense_center<<<(int)blocks, 256, 0, st>>>(V, n, S, ldv);
  return cudaGetLastError();
}

__global__ void k_dense_gather(const double* __restrict__ V, long long ldv, int S, const long long* __restrict__ idx, long long m,
                               double* __restrict__ out, long long ldo) {
  for (long long r = blockIdx.x; r < m; r += gridDim.x) {
    const double* src = V + idx[r] * ldv;
    double* dst = out + r * ldo;
    for (int s = threadIdx.x; s < S; s += blockDim.x) dst[s] = src[s];
  }
}
cudaError_t launch_dense_gather(const double* V, long long ldv, int S, const long long* idx, long long m, double* out, long long ldo,
                                cudaStream_t st) {
  if (m <= 0) return cudaSuccess;
  long long blocks = m < 148 * 16 ? m : 148 * 16;
  k_dense_gather<<<(int)blocks, 128, 0, st>>>(V, ldv, S, idx, m, out, ldo);
  return cudaGetLastError();
}

// out[c][r] = A[r][c]   (user-supplied (S, N) C-contiguous A -> datapoint-major N x S)
__global__ void k_transpose(const double* __restrict__ A, long long rows, long long cols, long long lda, double* __restrict__ out,
                            long long ldo) {
  __shared__ double tile[32][33];
  const long long bx = (long long)blockIdx.x * 32, by = (long long)blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const long long r = by + j, c = bx + threadIdx.x;
    if (r < rows && c < cols) tile[j][threadIdx.x] = A[r * lda + c];
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const long long c = bx + j, r = by + threadIdx.x;
    if (r < rows && c < cols) out[c * ldo + r] = tile[threadIdx.x][j];
  }
}
cudaError_t launch_transpose(const double* A, long long rows, long long cols, long long lda, double* out, long long ldo,
                             cudaStream_t st) {
  if (rows <= 0 || cols <= 0) return cudaSuccess;
  dim3 grid((unsigned)((cols + 31) / 32), (unsigned)((rows + 31) / 32));
  k_transpose<<<grid, dim3(32, 8), 0, st>>>(A, rows, cols, lda, out, ldo);
  return cudaGetLastError();
}

// -------------------------------------------------- S-length solver steps ---
// One CTA.  out layout (doubles):
//  op 0 GIGA direction (giga.py:21-30): in xw (raw A w), bn -> u[0:S] = cdir/|cdir|, u[S:2S] = xw/nw;
//        out[0] = cdirnrm, out[1] = nw
//  op 1 GIGA step (giga.py:42-62): in xw (raw), xf, bn, bnorm=aux -> out[0]=gA, out[1]=gB, out[2]=alpha, out[3]=beta
//  op 2 residual (frankwolfe.py:16, orthopursuit.py:18): u = b - xw; out[0] = sqrt(sum (xw-b)^2)  (snnls.py:28-29)
//  op 3 FW step (frankwolfe.py:30-31): in xw, xf, b, aux = nsum/nf -> out[0]=gammanum, out[1]=gammadenom
// The bodies of ops 1-3 as device functions (every thread of the CTA returns the same scalars): k_vec_step and the
// device-resident solver iteration k_solver_step run exactly the same arithmetic.
__device__ __forceinline__ void vec_giga_step(const double* xw, const double* xf, const double* b, int S, double aux, double* red, double& gA,
                                              double& gB, double& alpha, double& beta) {
  const int tid = threadIdx.x, nt = blockDim.x;
  double a = 0.0, f2 = 0.0;
  for (int s = tid; s < S; s += nt) {
    a = fma(xw[s], xw[s], a);
    f2 = fma(xf[s], xf[s], f2);
  }
  double nw = sqrt(block_sum_d(a, red));
  nw = (nw == 0.0) ? 1.0 : nw;
  const double nf = sqrt(block_sum_d(f2, red));
  double dbf = 0.0, dbw = 0.0, dwf = 0.0;
  for (int s = tid; s < S; s += nt) {
    const double wn = xw[s] / nw, fn = xf[s] / nf;
    dbf = fma(b[s], fn, dbf);
    dbw = fma(b[s], wn, dbw);
    dwf = fma(wn, fn, dwf);
  }
  dbf = block_sum_d(dbf, red);
  dbw = block_sum_d(dbw, red);
  dwf = block_sum_d(dwf, red);
  gA = dbf - dbw * dwf;
  gB = dbw - dbf * dwf;
  const double ca = gB / (gA + gB) / nw;
  const double cb = gA / (gA + gB) / nf;
  double x2 = 0.0;
  for (int s = tid; s < S; s += nt) {
    const double x = ca * xw[s] + cb * xf[s];
    x2 = fma(x, x, x2);
  }
  const double nx = sqrt(block_sum_d(x2, red));
  double xb = 0.0;
  for (int s = tid; s < S; s += nt) {
    const double x = ca * xw[s] + cb * xf[s];
    xb = fma(x / nx, b[s], xb);
  }
  xb = block_sum_d(xb, red);
  const double scale = aux / nx * xb;  // aux = bnorm
  alpha = ca * scale;
  beta = cb * scale;
}
__device__ __forceinline__ double vec_resid(const double* xw, const double* b, int S, double* u, double* red) {
  const int tid = threadIdx.x, nt = blockDim.x;
  double e2 = 0.0;
  for (int s = tid; s < S; s += nt) {
    const double d = xw[s] - b[s];
    if (u) u[s] = b[s] - xw[s];
    e2 = fma(d, d, e2);
  }
  e2 = block_sum_d(e2, red);
  return sqrt(e2);
}
__device__ __forceinline__ void vec_fw_step(const double* xw, const double* xf, const double* b, int S, double aux, double* red, double& num,
                                            double& den) {
  const int tid = threadIdx.x, nt = blockDim.x;
  num = 0.0;
  den = 0.0;
  for (int s = tid; s < S; s += nt) {
    const double dlt = aux * xf[s] - xw[s];
    num = fma(dlt, b[s] - xw[s], num);
    den = fma(dlt, dlt, den);
  }
  num = block_sum_d(num, red);
  den = block_sum_d(den, red);
}

// `stop` (optional): the control block of a device-resident solver run (k_solver_step); stop[1] != 0 = a guard has
// tripped, the speculatively queued iterations behind it do nothing
__global__ void k_vec_step(int op, const double* __restrict__ xw, const double* __restrict__ xf, const double* __restrict__ b, int S,
                           double aux, double* __restrict__ u, double* __restrict__ out, const double* stop) {
  __shared__ double red[33];
  const int tid = threadIdx.x, nt = blockDim.x;
  if (stop && stop[1] != 0.0) return;
  if (op == 0) {
    double a = 0.0;
    for (int s = tid; s < S; s += nt) a = fma(xw[s], xw[s], a);
    double nw = sqrt(block_sum_d(a, red));
    nw = (nw == 0.0) ? 1.0 : nw;
    double d = 0.0;
    for (int s = tid; s < S; s += nt) d = fma(b[s], xw[s] / nw, d);
    d = block_sum_d(d, red);
    double c2 = 0.0;
    for (int s = tid; s < S; s += nt) {
      const double xn = xw[s] / nw;
      const double c = b[s] - d * xn;
      u[s] = c;
      u[S + s] = xn;
      c2 = fma(c, c, c2);
    }
    const double cn = sqrt(block_sum_d(c2, red));
    for (int s = tid; s < S; s += nt) u[s] = u[s] / cn;
    if (tid == 0) {
      out[0] = cn;
      out[1] = nw;
    }
  } else if (op == 1) {
    double gA, gB, alpha, beta;
    vec_giga_step(xw, xf, b, S, aux, red, gA, gB, alpha, beta);
    if (tid == 0) {
      out[0] = gA;
      out[1] = gB;
      out[2] = alpha;
      out[3] = beta;
    }
  } else if (op == 2) {
    const double e = vec_resid(xw, b, S, u, red);
    if (tid == 0) out[0] = e;
  } else {
    double num, den;
    vec_fw_step(xw, xf, b, S, aux, red, num, den);
    if (tid == 0) {
      out[0] = num;
      out[1] = den;
    }
  }
}

cudaError_t launch_vec_step(int op, const double* xw, const double* xf, const double* b, int S, double aux, double* u, double* out,
                            cudaStream_t st, const double* stop) {
  k_vec_step<<<1, 256, 0, st>>>(op, xw, xf, b, S, aux, u, out, stop);
  return cudaGetLastError();
}

// ---------------------------------- device-resident solver iteration (GIGA / Frank-Wolfe) ----
// What SparseNNLS.build does on the HOST between the score pass of an iteration and the next one (snnls.py:50-62 with
// giga.py:40-64 / frankwolfe.py:19-40), as one single-CTA kernel working on device-resident solver state, so that a run of
// iterations can be queued without a host round trip per iteration:
//   guards of _select (GIGA: cdirnrm < TOL) -> activation of the selected datapoint f in the row cache -> line search ->
//   its guards -> w = alpha w, w_f = max(0, w_f + beta) -> A w re-formed from the cached rows -> new error -> the strict
//   monotone check (error > previous error: restore the weights, stop).
// Any guard that trips sets ctl[1] and leaves the state as it was BEFORE the iteration; the iterations queued behind it see
// the flag and do nothing; the host then takes that one iteration through its own code path, which raises and handles the
// reference's NumericalPrecisionError exactly as before.  The arithmetic is the host path's, operation for operation
// (same kernels' code, explicit roundings where Python would round twice).
//   ctl (doubles): [0] m = cached datapoints, [1] status (0 ok, 1 select guard, 2 line-search guard, 3 error not monotone),
//                  [2] iterations completed, [3] 1 once a completed iteration passed the monotone check (the reference
//                  resets its retry latch there), [4] error of the current weights (valid whenever some weight is positive),
//                  [5] check_error_monotone, [6] capacity of the cache
//   aw[cap], act[cap] (int64 bit patterns), aw_prev[cap]: weights / global indices of the cached datapoints, weight backup
__global__ void __launch_bounds__(256) k_solver_step(int algo, const double* __restrict__ V, long long ldv, int S,
                                                     const double* __restrict__ norms, const double* __restrict__ b,
                                                     const double* __restrict__ b_err, double aux, double tol,
                                                     double* __restrict__ Vact, double* ctl, double* aw, double* aw_prev, long long* act,
                                                     double* xw, const double* sel /* [0..3] vec scalars, [4..7] score result */) {
  // b: the right-hand side of the line search (GIGA: b / |b|, with aux = |b|; Frank-Wolfe: b, with aux = sum of norms);
  // b_err: the right-hand side of error() -- always the caller's b (snnls.py:28-29)
  __shared__ double red[33];
  __shared__ int s_k, s_pos;
  const int tid = threadIdx.x, nt = blockDim.x;
  if (ctl[1] != 0.0) return;
  int m = (int)ctl[0];
  const bool check = ctl[5] != 0.0;
  const long long f = __double_as_longlong(sel[5]);
  if (algo == 0 && sel[0] < tol) {          // giga.py:28-29  cdirnrm < TOL
    if (tid == 0) ctl[1] = 1.0;
    return;
  }
  if (f < 0 || m >= (int)ctl[6]) {          // no candidate / cache full: let the host handle it
    if (tid == 0) ctl[1] = 4.0;
    return;
  }
  // size() > 0 before the step, position of f in the cache
  if (tid == 0) { s_k = -1; s_pos = 0; }
  __syncthreads();
  int pos = 0, k = -1;
  for (int j = tid; j < m; j += nt) {
    if (aw[j] > 0.0) pos = 1;
    if (act[j] == f) k = j;
  }
  if (pos) s_pos = 1;
  if (k >= 0) s_k = k;
  __syncthreads();
  const bool size_nonzero = s_pos != 0;
  k = s_k;
  const double prev_error = ctl[4];
  if (k < 0) {                               // first selection of f: its row joins the cache with weight 0
    k = m;
    for (int s = tid; s < S; s += nt) Vact[(size_t)k * S + s] = V[f * ldv + s];
    if (tid == 0) {
      act[k] = f;
      aw[k] = 0.0;
      ctl[0] = (double)(m + 1);
    }
    m += 1;
    __syncthreads();
  }
  const double* xf = Vact + (size_t)k * S;
  double alpha, beta;
  if (algo == 0) {
    double gA, gB;
    vec_giga_step(xw, xf, b, S, aux, red, gA, gB, alpha, beta);
    if (gA <= 0.0 || gB < 0.0) {             // giga.py:48-49
      if (tid == 0) ctl[1] = 2.0;
      return;
    }
  } else {
    const double nf = norms[f];
    if (!size_nonzero) {                     // frankwolfe.py:21-23
      alpha = 0.0;
      beta = __ddiv_rn(aux, nf);
    } else {
      double num, den;
      vec_fw_step(xw, xf, b, S, __ddiv_rn(aux, nf), red, num, den);
      if (num < 0.0 || den == 0.0 || num > den) {   // frankwolfe.py:32-33
        if (tid == 0) ctl[1] = 2.0;
        return;
      }
      alpha = __dsub_rn(1.0, __ddiv_rn(num, den));
      beta = __ddiv_rn(__dmul_rn(__ddiv_rn(aux, nf), num), den);
    }
  }
  // w = alpha w ; w_f = max(0, w_f + beta)      (giga.py:63-64, frankwolfe.py:39-40; two roundings, like the host's floats)
  for (int j = tid; j < m; j += nt) {
    const double old = aw[j];
    aw_prev[j] = old;
    double v = __dmul_rn(alpha, old);
    if (j == k) {
      v = __dadd_rn(v, beta);
      v = (v > 0.0) ? v : 0.0;
    }
    aw[j] = v;
  }
  __syncthreads();
  // A w from the cached rows, in cache order (k_dense_combine), then the error (k_vec_step op 2)
  for (int s = tid; s < S; s += nt) {
    double acc = 0.0;
    for (int j = 0; j < m; ++j) acc = fma(aw[j], Vact[(size_t)j * S + s], acc);
    xw[s] = acc;
  }
  __syncthreads();
  const double err = vec_resid(xw, b_err, S, nullptr, red);
  if (check && size_nonzero && err > prev_error) {   // snnls.py:56-62
    for (int j = tid; j < m; j += nt) aw[j] = aw_prev[j];
    __syncthreads();
    for (int s = tid; s < S; s += nt) {
      double acc = 0.0;
      for (int j = 0; j < m; ++j) acc = fma(aw[j], Vact[(size_t)j * S + s], acc);
      xw[s] = acc;
    }
    if (tid == 0) ctl[1] = 3.0;
    return;
  }
  if (tid == 0) {
    ctl[4] = err;
    ctl[2] += 1.0;
    if (check && size_nonzero) ctl[3] = 1.0;
  }
}

cudaError_t launch_solver_step(int algo, const double* V, long long ldv, int S, const double* norms, const double* b, const double* b_err,
                               double aux, double tol, double* Vact, double* ctl, double* aw, double* aw_prev, long long* act, double* xw,
                               const double* sel, cudaStream_t st) {
  k_solver_step<<<1, 256, 0, st>>>(algo, V, ldv, S, norms, b, b_err, aux, tol, Vact, ctl, aw, aw_prev, act, xw, sel);
  return cudaGetLastError();
}

}  // namespace bc
