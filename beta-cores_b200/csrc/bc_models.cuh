// beta-cores B200: per-element potentials f(x_n, theta_s), fused into the epilogue of the
// projection contraction.  Each functor restates one reference function; `c` is the
// contraction value  sum_k A[n][k] * B[s][k]  for the operands bc_set_samples() prepares.
// The transcendental parts use the branch-free restricted-domain functions of bc_fastmath.cuh.
#pragma once
#if defined(__CUDACC__)
#include "bc_common.cuh"
#endif
#include "bc_fastmath.cuh"

namespace bc {

enum : int { MODEL_LOGISTIC = 0, MODEL_GAUSSIAN = 1, MODEL_NEURLIN = 2 };
enum : int { KIND_LOGLIK = 0, KIND_BETALIK = 1, KIND_BETAGRAD = 2 };

constexpr int kPowPolyMax = 24;  // highest degree of the (1+t)^-beta polynomial
constexpr int kPowTab = -1;      // POLY value of the lane-table form of (1+t)^-beta (tensor-core kernel only)
constexpr int kPowTabDeg = 6;    // degree of Q in (1+w)^-beta = 1 + w Q(w)
struct ModelParams {
  double p[8];
  double q[kPowPolyMax + 1];  // logistic beta-likelihood: (1+t)^-beta on t in [0,1] as a polynomial in 2t-1, highest degree
                              // first, fitted by bc_set_potential for the current beta (bc_api.cu: fit_pow_poly)
  double w[kPowTabDeg + 1];   // lane-table form: Q's coefficients, highest degree first (bc_api.cu: fit_pow_tab)
};

// per-thread tables of a potential's lane-table form (bc_fastmath.cuh: LaneTab32); empty for the potentials without one.
//   dev: the 64 doubles bc_set_potential uploads for the logistic beta-likelihood, [0..32) = 1/s_j, [32..64) = s_j^-beta
//        with s_j = 1 + (j + 1/2)/32 the centres of the 32 intervals of s = 1 + t in [1, 2]
struct NoTabs {};
struct ExpTabs {
  LaneTab32 e2;
};
struct LogisticTabs {
  LaneTab32 e2, rs, us;
};

// reference: examples/common/model_lr.py:72-79 (log_likelihood), :81-86 (beta_likelihood).
//   A = Z (rows y_n x_n), B = Theta, m = -c.
//   p[0] = beta, p[1] = (beta+1)/beta
// POLY = degree of the (1+t)^-beta polynomial (20 or 24), or 0: evaluate it as exp(-beta log1p t) (any beta).
//
// beta-likelihood, overflow-free: a = |m|, t = e^-a in (0,1], big = 1/(1+t), small = t big,
//   E = big^beta = (1+t)^-beta, G = e^(-beta a) so that small^beta = G E.  Then
//   (1+e^m)^-beta     = (m >= 0) ? G E : E
//   (1+e^m)^(-beta-1) + (1+e^-m)^(-beta-1) = big^(beta+1) + small^(beta+1) = E big (1 + G t)
//   f = -(k1 (1+e^m)^-beta - (...)) = E * ( big (1 + G t) - k1 ((m >= 0) ? G : 1) )
template <int KIND, int POLY = 20>
struct LogisticF {
  static constexpr bool kRowAux = false, kColAux = false;
  static constexpr bool kTab = (POLY == kPowTab);
  struct Tabs : LogisticTabs {};
#if defined(__CUDACC__)
  __device__ __forceinline__ static Tabs tabs(const double* dev, int lane) {
    Tabs T = {};
    if (kTab) {
      T.e2.mine = kExp2Tab32[lane];
      if (KIND == KIND_BETALIK) {
        T.rs.mine = __ldg(dev + lane);
        T.us.mine = __ldg(dev + 32 + lane);
      } else {               // log-likelihood: 1/s_j and log(s_j), no model constant in them
        T.rs.mine = kRcpTab32[lane];
        T.us.mine = kLogTab32[lane];
      }
      // keep the entries in registers: without the barrier ptxas re-reads a __constant__ entry at every use, a load with
      // 32 different addresses per warp (seen as LDC.64 c[0x3][R] once per batch in the SASS of the evaluation loop)
      asm volatile("" : "+d"(T.e2.mine), "+d"(T.rs.mine), "+d"(T.us.mine));
    }
    return T;
  }
#endif
  BC_HD static double eval(double c, double, double, const ModelParams& mp) {
    const double m = -c;
    const double a = fabs(m);
    const double t = exp_nonpos(-a);
    if (KIND == KIND_LOGLIK) {
      // m < 100: -log1p(e^m); else -m   ==   -(max(m, 0) + log1p(e^-|m|))   (for m >= 100 the log1p term is below ulp(m)/2)
      const double r = -(fmax(m, 0.0) + log1p_unit(t));
      return (m != m) ? m : r;
    } else {
      const double beta = mp.p[0], k1 = mp.p[1];
      const double G = exp_nonpos(-beta * a);
      const double big = rcp_1to2(1.0 + t);
      double E;
      if (POLY <= 0) {
        E = exp_nonpos(-beta * log1p_unit(t));
      } else {
        E = horner<(POLY > 0 ? POLY : 1)>(mp.q + (kPowPolyMax - POLY), fm_fma(t, 2.0, -1.0));
      }
      const double sel = (m >= 0.0) ? k1 * G : k1;
      const double r = E * fm_fma(big, fm_fma(G, t, 1.0), -sel);
      return (m != m) ? m : r;
    }
  }
  // W independent elements, stage by stage (instruction-level parallelism for the FP64 pipe).  Same arithmetic as eval()
  // for finite arguments.  Used by the tensor-core kernel, which supplies its own NaN handling (a non-finite row poisons
  // the row's pivot) -- a NaN argument here gives an unspecified result.
  //   p[2] = 700, p[3] = 700 / beta (low word cleared): the largest |m| passed to e^-a and to e^(-beta a); beyond them the
  //          terms are below 1e-300.  (One clamp at 700 / max(1, beta) for both, as shipped in round 1, froze e^(-beta a) at
  //          e^(-700 beta) for |m| > 700: wrong by k1 e^-7 at beta = 0.01.)
  template <int W>
  BC_HD static void evalv(const double (&c)[W], double, const double (&)[W], const ModelParams& mp, const Tabs& T, double (&out)[W]) {
    double a[W], x[W], t[W];
    if (kTab && KIND == KIND_LOGLIK) {
      // lane-table form of -(max(m, 0) + log1p(e^-|m|)), m = -c: t by exp_tab_v, log(1 + t) = log(s_j) + log1p(w) with the
      // interval tables of bc_fastmath.cuh (T.us holds log(s_j) here): 22 FP64 instructions instead of 42
      double s[W];
      BC_UNROLL for (int i = 0; i < W; ++i) {
        const int hi = fm_hi(c[i]) & 0x7fffffff;
        x[i] = -fm_hilo2d(hi < 0x4085e000 ? hi : 0x4085e000, fm_lo(c[i]));   // |c| capped at 700 (and a hair: the low word stays)
      }
      exp_tab_v<W, W, 3>(x, T.e2, t);   // degree-5 form: t enters through log(1 + t), absolute error below 2e-16
      BC_UNROLL for (int i = 0; i < W; ++i) s[i] = 1.0 + t[i];
      double d[W], R[W], L[W], w[W], w2[W], P[W];
      BC_UNROLL for (int i = 0; i < W; ++i) {
        const int hs = fm_hi(s[i]) < 0x3fffffff ? fm_hi(s[i]) : 0x3fffffff;     // (as in the beta-likelihood form below)
        d[i] = s[i] - fm_hilo2d((hs & 0xffff8000) | 0x4000, 0);
        R[i] = T.rs.at(hs >> 15);
        L[i] = T.us.at(hs >> 15);
      }
      BC_UNROLL for (int i = 0; i < W; ++i) w[i] = d[i] * R[i];
      BC_UNROLL for (int i = 0; i < W; ++i) w2[i] = w[i] * w[i];
      horner_v<5, W>(kLog1pTabPoly, w, P);
      BC_UNROLL for (int i = 0; i < W; ++i) P[i] = fm_fma(w2[i], P[i], w[i]) + L[i];
      BC_UNROLL for (int i = 0; i < W; ++i) {
        const int hi = fm_hi(c[i]);
        const double mx = fm_hilo2d(hi < 0 ? (hi & 0x7fffffff) : 0, hi < 0 ? fm_lo(c[i]) : 0);   // max(-c, 0)
        out[i] = -(mx + P[i]);
      }
    } else if (kTab) {
      // lane-table form: 35 FP64 instructions per element instead of 65.
      //   |c| and the clamp on the integer pipe (p[2] has a zero low word, so comparing high words is the exact compare);
      //   t = e^-a, G = e^(-beta a) by exp_tab_v, both with the one-step reduction: the error K 1.7e-18 e^x (K = 32 x / ln2) is at
      //   most 3e-17 absolute, and G enters the result with weight k1: measured against 50-digit arithmetic the worst error of the
      //   potential is the same with the two-step reduction (6.5e-16 vs 7.2e-16 at beta = 0.1, 7.5e-16 vs 5.9e-16 at 0.01), and
      //   6.8e-16 / 9.8e-16 with the degree-5 form of the table exponential used here (test bound: 5e-16 k1, k1 = 11 / 101);
      //   big = 1/(1+t) by one cubic step;
      //   E = (1+t)^-beta: s = 1 + t in [1,2] falls into interval j = its top five mantissa bits, s = s_j + d, |d| <= 1/64,
      //   E = s_j^-beta (1 + w)^-beta with w = d / s_j, (1 + w)^-beta = 1 + w Q(w), Q of degree 6: 10 FP64 instructions
      //   against the 21 of the degree-20 polynomial.
      const double beta = mp.p[0], k1 = mp.p[1];
      const int amax_hi = fm_hi(mp.p[2]), bmax_hi = fm_hi(mp.p[3]);
      double G[W], big[W], E[W], s[W];
      {
        double xx[2 * W], yy[2 * W];
        BC_UNROLL for (int i = 0; i < W; ++i) {
          // the high word alone is capped (one integer min each): beyond the cap the value keeps its low word, i.e. exceeds
          // the cap by less than 2^-20 of it -- still inside the range exp_tab_v takes
          const int hi = fm_hi(c[i]) & 0x7fffffff, lo = fm_lo(c[i]);
          a[i] = fm_hilo2d(hi < amax_hi ? hi : amax_hi, lo);
          xx[i] = -a[i];
          xx[W + i] = -beta * fm_hilo2d(hi < bmax_hi ? hi : bmax_hi, lo);
        }
        exp_tab_v<2 * W, 2 * W, 3>(xx, T.e2, yy);
        BC_UNROLL for (int i = 0; i < W; ++i) {
          t[i] = yy[i];
          G[i] = yy[W + i];
        }
      }
      BC_UNROLL for (int i = 0; i < W; ++i) s[i] = 1.0 + t[i];
      rcp_1to2_cubic_v<W>(s, big);
      {
        double d[W], R[W], U[W], w[W], Q[W];
        BC_UNROLL for (int i = 0; i < W; ++i) {
          // interval = the top five mantissa bits of s; its centre s_j = those bits with a 1 appended.  s = 2 exactly (t = 1) is
          // capped into the last interval.  The lane tables take the index modulo 32.
          const int hs = fm_hi(s[i]) < 0x3fffffff ? fm_hi(s[i]) : 0x3fffffff;
          d[i] = s[i] - fm_hilo2d((hs & 0xffff8000) | 0x4000, 0);
          R[i] = T.rs.at(hs >> 15);
          U[i] = T.us.at(hs >> 15);
        }
        BC_UNROLL for (int i = 0; i < W; ++i) w[i] = d[i] * R[i];
        horner_v<kPowTabDeg, W>(mp.w, w, Q);
        BC_UNROLL for (int i = 0; i < W; ++i) Q[i] = w[i] * Q[i];
        BC_UNROLL for (int i = 0; i < W; ++i) E[i] = fm_fma(U[i], Q[i], U[i]);
      }
      BC_UNROLL for (int i = 0; i < W; ++i) {
        const double sel = (fm_hi(c[i]) < 0) ? k1 * G[i] : k1;
        out[i] = E[i] * fm_fma(big[i], fm_fma(G[i], t[i], 1.0), -sel);
      }
    } else if (KIND == KIND_LOGLIK) {
      BC_UNROLL for (int i = 0; i < W; ++i) a[i] = fabs(c[i]);
      BC_UNROLL for (int i = 0; i < W; ++i) x[i] = (a[i] > 700.0) ? -700.0 : -a[i];
      exp_core_v<W>(x, t);
      double u[W], l[W];
      BC_UNROLL for (int i = 0; i < W; ++i) u[i] = fm_fma(t[i], 2.0, -1.0);
      horner_v<22, W>(kLog1pPoly, u, l);
      BC_UNROLL for (int i = 0; i < W; ++i) out[i] = -(fmax(-c[i], 0.0) + l[i]);
    } else {
      const double beta = mp.p[0], k1 = mp.p[1], amax = mp.p[2], bmax = mp.p[3];
      double G[W], big[W], E[W], u[W], b[W];
      BC_UNROLL for (int i = 0; i < W; ++i) {
        const double ai = fabs(c[i]);
        a[i] = (ai > amax) ? amax : ai;
        b[i] = (ai > bmax) ? bmax : ai;
      }
      {
        // both exponentials of all W elements advance together: 2W independent dependency chains
        double xx[2 * W], yy[2 * W];
        BC_UNROLL for (int i = 0; i < W; ++i) {
          xx[i] = -a[i];
          xx[W + i] = -beta * b[i];
        }
        exp_core_v<2 * W>(xx, yy);
        BC_UNROLL for (int i = 0; i < W; ++i) {
          t[i] = yy[i];
          G[i] = yy[W + i];
        }
      }
      BC_UNROLL for (int i = 0; i < W; ++i) u[i] = 1.0 + t[i];
      rcp_1to2_v<W>(u, big);
      if (POLY <= 0) {
        double l[W];
        BC_UNROLL for (int i = 0; i < W; ++i) u[i] = fm_fma(t[i], 2.0, -1.0);
        horner_v<22, W>(kLog1pPoly, u, l);
        BC_UNROLL for (int i = 0; i < W; ++i) x[i] = -beta * l[i];
        exp_core_v<W>(x, E);
      } else {
        BC_UNROLL for (int i = 0; i < W; ++i) u[i] = fm_fma(t[i], 2.0, -1.0);
        horner_v<(POLY > 0 ? POLY : 1), W>(mp.q + (kPowPolyMax - POLY), u, E);
      }
      BC_UNROLL for (int i = 0; i < W; ++i) {
        // m = -c >= 0  <=>  sign bit of c set (c = +0 gives G = 1 either way): an integer test, not an FP64 compare
        const double sel = (fm_hi(c[i]) < 0) ? k1 * G[i] : k1;
        out[i] = E[i] * fm_fma(big[i], fm_fma(G[i], t[i], 1.0), -sel);
      }
    }
  }
};

// reference: examples/common/gaussian.py:7-15 (loglik), :34-44 (beta-lik), :46-62 (d/dbeta).
//   A = X, B = Theta Siginv^T (so c = x Siginv theta), rowaux = x Siginv x, colaux = th Siginv th,
//   q = rowaux + colaux - 2c.
//   loglik : p[0] - 0.5 q                       p[0] = -d/2 log 2pi - 1/2 logdetSig
//   betalik: p[1] exp(p[2] q) - p[3]            p[1] = 1/beta, p[2] = -beta/2, p[3] = (1+beta)^(-d/2-1)
//   betagrad: p[4](p[1] e - p[3]) - p[5] e - p[6] q e - p[7],  e = exp(p[2] q)
//            p[4] = logcnst, p[5] = 1/beta^2, p[6] = 1/(2 beta), p[7] = p[3] log(1+beta)
template <int KIND>
struct GaussianF {
  static constexpr bool kRowAux = true, kColAux = true;
  // the kinds with an exponential take it from the 2^(j/32) lane table in the tensor-core kernel (exp_tab_v: 10 FP64
  // instructions instead of 14); evalv<W>'s last-but-one argument.  The host build passes a plain array.
  struct Tabs : ExpTabs {};
#if defined(__CUDACC__)
  __device__ __forceinline__ static Tabs tabs(const double*, int lane) {
    Tabs T = {};
    if (KIND != KIND_LOGLIK) {
      T.e2.mine = kExp2Tab32[lane];
      asm volatile("" : "+d"(T.e2.mine));   // stays in registers (see LogisticF::tabs)
    }
    return T;
  }
#endif
  BC_HD static double eval(double c, double ra, double ca, const ModelParams& mp) {
    const double q = ra + ca - 2.0 * c;
    if (KIND == KIND_LOGLIK) {
      return mp.p[0] - 0.5 * q;
    } else if (KIND == KIND_BETALIK) {
      return mp.p[1] * exp_clamped(mp.p[2] * q) - mp.p[3];
    } else {
      const double e = exp_clamped(mp.p[2] * q);
      const double t1 = mp.p[4] * (mp.p[1] * e - mp.p[3]);
      return t1 - mp.p[5] * e - mp.p[6] * q * e - mp.p[7];
    }
  }
  template <int W>
  BC_HD static void evalv(const double (&c)[W], double ra, const double (&ca)[W], const ModelParams& mp, const Tabs& T, double (&out)[W]) {
    double q[W];
    BC_UNROLL for (int i = 0; i < W; ++i) q[i] = ra + ca[i] - 2.0 * c[i];
    if (KIND == KIND_LOGLIK) {
      BC_UNROLL for (int i = 0; i < W; ++i) out[i] = mp.p[0] - 0.5 * q[i];
    } else {
      double x[W], e[W];
      BC_UNROLL for (int i = 0; i < W; ++i) {
        const double y = mp.p[2] * q[i];
        const double yc = (y < -700.0) ? -700.0 : y;
        x[i] = (yc > 700.0) ? 700.0 : yc;
      }
      exp_tab_v<W, W, 3>(x, T.e2, e);   // one-step reduction, degree-5 form: absolute error below 3e-16 of the largest term
      BC_UNROLL for (int i = 0; i < W; ++i) {
        const double y = mp.p[2] * q[i];
        const double ee = (y != y) ? y : e[i];
        if (KIND == KIND_BETALIK) {
          out[i] = mp.p[1] * ee - mp.p[3];
        } else {
          const double t1 = mp.p[4] * (mp.p[1] * ee - mp.p[3]);
          out[i] = t1 - mp.p[5] * ee - mp.p[6] * q[i] * ee - mp.p[7];
        }
      }
    }
  }
};

// reference: examples/common/model_neurlinr.py:90-97 (loglik), :102-110 (beta-lik).
//   A = Phi (first D columns of z), B = Theta, rowaux = y (column D of z), u = c,
//   r2 = y^2 - 2 u y + u^2.
//   loglik : p[0] - p[1] r2                     p[0] = -1/2 log(2 pi s2), p[1] = 1/(2 s2)
//   betalik: p[2] (p[3] exp(p[4] r2) + p[5])    p[2] = (2 pi s2)^(-beta/2), p[3] = -(beta+1)/beta,
//                                               p[4] = -beta/(2 s2), p[5] = 1/sqrt(1+beta)
template <int KIND>
struct NeurlinF {
  static constexpr bool kRowAux = true, kColAux = false;
  // the kinds with an exponential take it from the 2^(j/32) lane table in the tensor-core kernel (exp_tab_v: 10 FP64
  // instructions instead of 14); evalv<W>'s last-but-one argument.  The host build passes a plain array.
  struct Tabs : ExpTabs {};
#if defined(__CUDACC__)
  __device__ __forceinline__ static Tabs tabs(const double*, int lane) {
    Tabs T = {};
    if (KIND != KIND_LOGLIK) {
      T.e2.mine = kExp2Tab32[lane];
      asm volatile("" : "+d"(T.e2.mine));   // stays in registers (see LogisticF::tabs)
    }
    return T;
  }
#endif
  BC_HD static double eval(double c, double y, double, const ModelParams& mp) {
    const double r2 = y * y - 2.0 * c * y + c * c;
    if (KIND == KIND_LOGLIK) {
      return mp.p[0] - mp.p[1] * r2;
    } else {
      return mp.p[2] * (mp.p[3] * exp_clamped(mp.p[4] * r2) + mp.p[5]);
    }
  }
  template <int W>
  BC_HD static void evalv(const double (&c)[W], double y, const double (&)[W], const ModelParams& mp, const Tabs& T, double (&out)[W]) {
    double r2[W];
    BC_UNROLL for (int i = 0; i < W; ++i) r2[i] = y * y - 2.0 * c[i] * y + c[i] * c[i];
    if (KIND == KIND_LOGLIK) {
      BC_UNROLL for (int i = 0; i < W; ++i) out[i] = mp.p[0] - mp.p[1] * r2[i];
    } else {
      double x[W], e[W];
      BC_UNROLL for (int i = 0; i < W; ++i) {
        const double z = mp.p[4] * r2[i];
        const double zc = (z < -700.0) ? -700.0 : z;
        x[i] = (zc > 700.0) ? 700.0 : zc;
      }
      exp_tab_v<W, W, 3>(x, T.e2, e);   // one-step reduction, degree-5 form: absolute error below 3e-16 of the largest term
      BC_UNROLL for (int i = 0; i < W; ++i) {
        const double z = mp.p[4] * r2[i];
        const double ee = (z != z) ? z : e[i];
        out[i] = mp.p[2] * (mp.p[3] * ee + mp.p[5]);
      }
    }
  }
};

}  // namespace bc
