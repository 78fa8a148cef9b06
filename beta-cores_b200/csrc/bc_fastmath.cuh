// beta-cores B200: branch-free fp64 elementary functions for the projection epilogue.
//
// The epilogue evaluates N*S potentials per pass (1e10 at the north-star size) on the FP64 pipe; libm's
// exp/log1p/pow carry slow-path branches and ~2x the instructions these restricted-domain versions need:
//   exp_clamped(x) : e^x for x clamped to [-700, 700]      (Cody-Waite reduction + degree-11 polynomial, <1 ulp)
//   rcp_1to2(x)    : 1/x for x in [1, 2]                    (MUFU.RCP64H seed + two Newton steps)
//   horner<N>      : polynomial with coefficients in the kernel-parameter constant bank / __constant__ memory
//   log1p_unit(t)  : log(1+t) for t in [0, 1]               (degree-22 Chebyshev-economised polynomial, 3e-19 abs)
// Every function is a pure function of its argument bits: equal inputs give equal outputs (the row-constant
// -> exact-zero centring property of the reference, bcores.py:78, depends on it).
// The file also compiles as plain C++ (tests/native builds it with g++ to check the polynomials on the CPU).
#pragma once
#include <stdint.h>
#if defined(__CUDACC__)
#define BC_HD __device__ __forceinline__
#else
#include <cmath>
#include <cstring>
#define BC_HD inline
#endif

namespace bc {

BC_HD double fm_fma(double a, double b, double c) {
#if defined(__CUDACC__)
  return fma(a, b, c);
#else
  return std::fma(a, b, c);
#endif
}
BC_HD int fm_lo(double x) {
#if defined(__CUDACC__)
  return __double2loint(x);
#else
  uint64_t u;
  memcpy(&u, &x, 8);
  return (int)(uint32_t)u;
#endif
}
BC_HD int fm_hi(double x) {
#if defined(__CUDACC__)
  return __double2hiint(x);
#else
  uint64_t u;
  memcpy(&u, &x, 8);
  return (int)(uint32_t)(u >> 32);
#endif
}
BC_HD double fm_i2d(int k) {
#if defined(__CUDACC__)
  return __int2double_rn(k);
#else
  return (double)k;
#endif
}
BC_HD double fm_add_exponent(double p, int k) {  // p * 2^k for results that stay normal
#if defined(__CUDACC__)
  return __hiloint2double(__double2hiint(p) + (k << 20), __double2loint(p));
#else
  uint64_t u;
  memcpy(&u, &p, 8);
  u += (uint64_t)(int64_t)k << 52;
  memcpy(&p, &u, 8);
  return p;
#endif
}

// e^x for |x| <= 700 (no range or NaN handling: see the wrappers below).  |error| < 1 ulp.
// W independent arguments advance through the dependent FMA chain in lock step: the source order IS the instruction
// level parallelism the FP64 pipe needs (one chain per thread leaves it idle for the length of its latency).
#if defined(__CUDACC__)
#define BC_UNROLL _Pragma("unroll")
#else
#define BC_UNROLL
#endif
template <int W>
BC_HD void exp_core_v(const double (&x)[W], double (&y)[W]) {
  const double kMagic = 6755399441055744.0;  // 1.5 * 2^52: the integer nearest to kf lands in the low word
  const double c[11] = {2.76327915826623365e-07, 2.75572245781815778e-06, 2.48014850709743224e-05, 1.98412699096287991e-04,
                        1.38888889526998121e-03, 8.33333333330933505e-03, 4.16666666664866486e-02, 1.66666666666667018e-01,
                        5.00000000000001887e-01, 1.0,                     1.0};
  double kf[W], r[W], p[W];
  int k[W];
  BC_UNROLL for (int i = 0; i < W; ++i) kf[i] = fm_fma(x[i], 1.4426950408889634, kMagic);
  BC_UNROLL for (int i = 0; i < W; ++i) {
    k[i] = fm_lo(kf[i]);
    kf[i] = fm_i2d(k[i]);      // the same value as kf - kMagic, from the conversion unit instead of the FP64 pipe
  }
  BC_UNROLL for (int i = 0; i < W; ++i) r[i] = fm_fma(kf[i], -6.93147180369123816490e-01, x[i]);
  BC_UNROLL for (int i = 0; i < W; ++i) r[i] = fm_fma(kf[i], -1.90821492927058770002e-10, r[i]);
  BC_UNROLL for (int i = 0; i < W; ++i) p[i] = 2.51149959513019969e-08;
  BC_UNROLL for (int j = 0; j < 11; ++j) {
    BC_UNROLL for (int i = 0; i < W; ++i) p[i] = fm_fma(p[i], r[i], c[j]);
  }
  BC_UNROLL for (int i = 0; i < W; ++i) y[i] = fm_add_exponent(p[i], k[i]);
}
BC_HD double exp_core(double x) {
  const double a[1] = {x};
  double y[1];
  exp_core_v<1>(a, y);
  return y[0];
}
// e^x with x clamped to [-700, 700]; NaN propagates.
BC_HD double exp_clamped(double x) {
  double xc = (x < -700.0) ? -700.0 : x;
  xc = (xc > 700.0) ? 700.0 : xc;
  const double y = exp_core(xc);
  return (x != x) ? x : y;
}
// e^x for x <= 0 (or NaN: result unspecified, the caller tests its own input), clamped below at -700
BC_HD double exp_nonpos(double x) { return exp_core((x < -700.0) ? -700.0 : x); }

// 1/x for x in [1, 2]
BC_HD double rcp_1to2(double x) {
#if defined(__CUDACC__)
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  double e = fm_fma(-x, r, 1.0);
  r = fm_fma(r, e, r);
  e = fm_fma(-x, r, 1.0);
  r = fm_fma(r, e, r);
  return r;
#else
  return 1.0 / x;
#endif
}

template <int W>
BC_HD void rcp_1to2_v(const double (&x)[W], double (&r)[W]) {
#if defined(__CUDACC__)
  double e[W];
  BC_UNROLL for (int i = 0; i < W; ++i) asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r[i]) : "d"(x[i]));
  BC_UNROLL for (int i = 0; i < W; ++i) e[i] = fm_fma(-x[i], r[i], 1.0);
  BC_UNROLL for (int i = 0; i < W; ++i) r[i] = fm_fma(r[i], e[i], r[i]);
  BC_UNROLL for (int i = 0; i < W; ++i) e[i] = fm_fma(-x[i], r[i], 1.0);
  BC_UNROLL for (int i = 0; i < W; ++i) r[i] = fm_fma(r[i], e[i], r[i]);
#else
  for (int i = 0; i < W; ++i) r[i] = 1.0 / x[i];
#endif
}

template <int N, int W>
BC_HD void horner_v(const double* c, const double (&x)[W], double (&p)[W]) {
  BC_UNROLL for (int i = 0; i < W; ++i) p[i] = c[0];
  BC_UNROLL for (int j = 1; j <= N; ++j) {
    BC_UNROLL for (int i = 0; i < W; ++i) p[i] = fm_fma(p[i], x[i], c[j]);
  }
}

// coefficients highest degree first: c[0] x^N + ... + c[N]
template <int N>
BC_HD double horner(const double* c, double x) {
  double p = c[0];
#pragma unroll
  for (int i = 1; i <= N; ++i) p = fm_fma(p, x, c[i]);
  return p;
}

// log(1+t), t in [0, 1]: polynomial in 2t-1 (Chebyshev interpolant at 64 nodes truncated to degree 22, monomial form)
#if defined(__CUDACC__)
static __constant__
#else
static const
#endif
    double kLog1pPoly[23] = {
        -2.74225015005545668e-12, 8.37205032055466135e-12,  -1.05354638931769186e-11, 3.46318473395092910e-11,
        -1.49468230558556527e-10, 4.71527752610951278e-10,  -1.44590556232740489e-09, 4.63129693464390621e-09,
        -1.49378205823659422e-08, 4.82569499167816174e-08,  -1.56804706454277560e-07, 5.13181046263311709e-07,
        -1.69350924755814828e-06, 5.64503012252609974e-06,  -1.90519737016393414e-05, 6.53210528462085902e-05,
        -2.28623685422467925e-04, 8.23045267500557409e-04,  -3.08641975308595294e-03, 1.23456790123452585e-02,
        -5.55555555555555663e-02, 3.33333333333333315e-01,  4.05465108108164385e-01};

BC_HD double log1p_unit(double t) { return horner<22>(kLog1pPoly, fm_fma(t, 2.0, -1.0)); }

}  // namespace bc
