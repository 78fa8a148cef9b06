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
BC_HD double fm_hilo2d(int hi, int lo) {
#if defined(__CUDACC__)
  return __hiloint2double(hi, lo);
#else
  const uint64_t u = ((uint64_t)(uint32_t)hi << 32) | (uint64_t)(uint32_t)lo;
  double x;
  memcpy(&x, &u, 8);
  return x;
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


// ------------------------------------------------------------------------------------------------ lane tables --
// A 32-entry table of doubles held ACROSS THE LANES OF A WARP: lane l keeps entry l in one register pair and a lookup with a
// per-thread index is one SHFL.IDX pair.  Unlike a gather from shared memory (LDS.64 with a divergent address: 2-3 bank-
// conflict wavefronts each, tried in round 1 and no faster than the FP64 instructions it replaced) the shuffle has no
// conflicts and no address arithmetic.  Only for code every lane of the warp executes together (the tensor-core kernel's
// epilogue).  The host build indexes a plain array.
struct LaneTab32 {
#if defined(__CUDACC__)
  double mine;   // entry [lane]
  // (shfl.idx takes the low five bits of the index: callers pass j unmasked, also a negative one)
  __device__ __forceinline__ double at(int j) const { return __shfl_sync(0xffffffffu, mine, j); }
#else
  const double* t;
  double at(int j) const { return t[j & 31]; }
#endif
};

// 2^(j/32), j = 0..31, correctly rounded
#if defined(__CUDACC__)
static __constant__
#else
static const
#endif
    double kExp2Tab32[32] = {1.0,
                             1.0218971486541166,
                             1.0442737824274138,
                             1.0671404006768237,
                             1.0905077326652577,
                             1.1143867425958924,
                             1.1387886347566916,
                             1.1637248587775775,
                             1.189207115002721,
                             1.215247359980469,
                             1.241857812073484,
                             1.2690509571917332,
                             1.2968395546510096,
                             1.3252366431597413,
                             1.3542555469368927,
                             1.383909881963832,
                             1.4142135623730951,
                             1.4451808069770467,
                             1.4768261459394993,
                             1.5091644275934228,
                             1.5422108254079407,
                             1.5759808451078865,
                             1.6104903319492543,
                             1.6457554781539649,
                             1.681792830507429,
                             1.7186192981224779,
                             1.7562521603732995,
                             1.7947090750031072,
                             1.8340080864093424,
                             1.8741676341103,
                             1.9152065613971474,
                             1.9571441241754002};

// e^x for -700 <= x <= 700 with the table: x = (32 k + j) ln2/32 + r, |r| <= ln2/64, e^x = 2^k T[j] (1 + q(r)),
// q(r) = r + r^2 P(r) with P of degree 4 (|1 + q - e^r| < 2.3e-19 e^r): 1 + (1 or 2) + 7 FP64 instructions against the
// 14 of exp_core_v.  Error: the rounding of T[j] and of the final FMA, < 1 ulp together.
// NLO of the W arguments (the first ones) take the one-step reduction r = x - K fl(ln2/32): its error K 1.7e-18 matters only
// relative to e^x itself and is for the caller to allow (see LogisticF: t = e^-a enters the result with weight O(t)).
template <int W, int NLO, int PDEG = 4>
BC_HD void exp_tab_v(const double (&x)[W], const LaneTab32& T, double (&y)[W]) {
  const double kMagic = 6755399441055744.0;
  // q(r) = r + r^2 P(r).  PDEG = 4: interpolant, |1 + q - e^r| < 2.3e-19 e^r.  PDEG = 3: weighted minimax fit (Lawson
  // iteration on r^2 (P - g) / e^r), < 8.8e-17 e^r -- 0.4 ulp for one FP64 instruction less.
  const double c4[5] = {1.38889253992791387136e-03, 8.33336254165508924507e-03, 4.16666666665591739482e-02, 1.66666666665806706416e-01,
                        5.00000000000000000000e-01};
  const double c3[4] = {8.33332307384487304402e-03, 4.16668942132359595987e-02, 1.66666666669932850287e-01, 4.99999999991721566506e-01};
  double kf[W], r[W], r2[W], p[W], t[W];
  int K[W];
  BC_UNROLL for (int i = 0; i < W; ++i) kf[i] = fm_fma(x[i], 4.61662413084468283841e+01, kMagic);
  BC_UNROLL for (int i = 0; i < W; ++i) {
    K[i] = fm_lo(kf[i]);
    kf[i] = fm_i2d(K[i]);
  }
  BC_UNROLL for (int i = 0; i < W; ++i) t[i] = T.at(K[i]);   // entry K mod 32
  BC_UNROLL for (int i = 0; i < W; ++i)
    r[i] = (i < NLO) ? fm_fma(kf[i], -2.16608493924982901946e-02, x[i]) : fm_fma(kf[i], -2.16608493865351192653e-02, x[i]);
  BC_UNROLL for (int i = 0; i < W; ++i) if (i >= NLO) r[i] = fm_fma(kf[i], -5.96317165397058656257e-12, r[i]);
  BC_UNROLL for (int i = 0; i < W; ++i) r2[i] = r[i] * r[i];
  BC_UNROLL for (int i = 0; i < W; ++i) p[i] = (PDEG == 4) ? c4[0] : c3[0];
  BC_UNROLL for (int j = 1; j <= PDEG; ++j) {
    BC_UNROLL for (int i = 0; i < W; ++i) p[i] = fm_fma(p[i], r[i], (PDEG == 4) ? c4[j] : c3[j < 4 ? j : 3]);
  }
  BC_UNROLL for (int i = 0; i < W; ++i) p[i] = fm_fma(r2[i], p[i], r[i]);
  // 2^k: (K >> 5) << 20 added to the high word = ((K & ~31) << 15), one logic + one multiply-add on the integer pipe
  BC_UNROLL for (int i = 0; i < W; ++i) {
    const double v = fm_fma(t[i], p[i], t[i]);
#if defined(__CUDACC__)
    int hi;   // one IMAD: left to itself ptxas turns the multiply into a shift and spends a separate add
    asm("mad.lo.s32 %0, %1, 32768, %2;" : "=r"(hi) : "r"(K[i] & ~31), "r"(fm_hi(v)));
    y[i] = fm_hilo2d(hi, fm_lo(v));
#else
    y[i] = fm_hilo2d(fm_hi(v) + (K[i] & ~31) * 32768, fm_lo(v));
#endif
  }
}

// 1/x for x in [1, 2]: the 2^-23 seed of rcp.approx.ftz.f64 and ONE cubic step r (1 + e + e^2), e = 1 - x r: 3 FP64
// instructions instead of the 4 of two Newton steps, error e^3 < 2^-60 before the final rounding.
template <int W>
BC_HD void rcp_1to2_cubic_v(const double (&x)[W], double (&r)[W]) {
  double e[W];
#if defined(__CUDACC__)
  BC_UNROLL for (int i = 0; i < W; ++i) asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r[i]) : "d"(x[i]));
#else
  // host stand-in for the seed: 1/x cut to the 20 mantissa bits of the high word (coarser than the hardware's 2^-23)
  for (int i = 0; i < W; ++i) r[i] = fm_hilo2d(fm_hi(1.0 / x[i]), 0);
#endif
  BC_UNROLL for (int i = 0; i < W; ++i) e[i] = fm_fma(-x[i], r[i], 1.0);
  BC_UNROLL for (int i = 0; i < W; ++i) e[i] = fm_fma(e[i], e[i], e[i]);
  BC_UNROLL for (int i = 0; i < W; ++i) r[i] = fm_fma(r[i], e[i], r[i]);
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

// 1/s_j and log(s_j) for the 32 interval centres s_j = 1 + (j + 1/2)/32 of [1, 2], correctly rounded (the lane-table form of
// log1p: log(s) = log(s_j) + log1p(w), w = (s - s_j)/s_j, |w| <= 1/65), and log1p(w) = w + w^2 P(w), P of degree 5
// (|error| < 1.3e-17 on the interval)
#if defined(__CUDACC__)
static __constant__
#else
static const
#endif
    double kRcpTab32[32] = {0.98461538461538467, 0.95522388059701491, 0.92753623188405798, 0.90140845070422537, 0.87671232876712324,
                            0.85333333333333339, 0.83116883116883122, 0.810126582278481,   0.79012345679012341, 0.77108433734939763,
                            0.75294117647058822, 0.73563218390804597, 0.7191011235955056,  0.70329670329670335, 0.68817204301075274,
                            0.67368421052631577, 0.65979381443298968, 0.64646464646464652, 0.63366336633663367, 0.62135922330097082,
                            0.60952380952380958, 0.59813084112149528, 0.58715596330275233, 0.57657657657657657, 0.5663716814159292,
                            0.55652173913043479, 0.54700854700854706, 0.53781512605042014, 0.52892561983471076, 0.52032520325203258,
                            0.51200000000000001, 0.50393700787401574};
#if defined(__CUDACC__)
static __constant__
#else
static const
#endif
    double kLogTab32[32] = {0.015504186535965254, 0.045809536031294201, 0.075223421237587532, 0.10379679368164356, 0.13157635778871926,
                            0.15860503017663857,  0.18492233849401199,  0.21056476910734964,  0.23556607131276691, 0.25995752443692605,
                            0.28376817313064462,  0.30702503529491187,  0.32975328637246798,  0.3519764231571782,  0.37371640979358406,
                            0.39499380824086899,  0.41582789514371099,  0.43623676677491807,  0.45623743348158757, 0.47584590486996392,
                            0.49507726679785152,  0.51394575110223428,  0.53246479886947184,  0.5506471179526623,  0.56850473535266877,
                            0.58604904500357824,  0.60329085143808425,  0.62024040975185757,  0.63690746223706918, 0.65330127201274568,
                            0.66943065394262924,  0.68530400309891937};
#if defined(__CUDACC__)
static __constant__
#else
static const
#endif
    double kLog1pTabPoly[6] = {1.42896638659497532409e-01,  -1.66711099228961212582e-01, 1.99999996490731690724e-01,
                               -2.49999996052097717136e-01, 3.33333333333379500107e-01,  -5.00000000000051958438e-01};

BC_HD double log1p_unit(double t) { return horner<22>(kLog1pPoly, fm_fma(t, 2.0, -1.0)); }

}  // namespace bc
