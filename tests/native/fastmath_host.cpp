// Host build of the projection epilogue's scalar functions (beta-cores_b200/csrc/bc_fastmath.cuh, bc_models.cuh) so that
// the CPU test-suite can check the polynomials and the algebraic rewrites against high-precision references.
// TEST INFRASTRUCTURE: compiled by tests/test_fastmath_cpu.py with g++ into tests/native/_build/; not part of the product.
#include "../../beta-cores_b200/csrc/bc_models.cuh"

extern "C" {
double fm_exp(double x) { return bc::exp_clamped(x); }
double fm_log1p_unit(double t) { return bc::log1p_unit(t); }
// q: kPowPolyMax+1 coefficients laid out as bc_set_potential stores them (right-aligned)
double fm_logistic(int kind, int poly, double c, double beta, const double* q) {
  bc::ModelParams mp;
  for (int i = 0; i < 8; ++i) mp.p[i] = 0.0;
  mp.p[0] = beta;
  mp.p[1] = (beta + 1.) / beta;
  for (int i = 0; i <= bc::kPowPolyMax; ++i) mp.q[i] = q ? q[i] : 0.0;
  if (kind == bc::KIND_LOGLIK) return bc::LogisticF<bc::KIND_LOGLIK, 0>::eval(c, 0, 0, mp);
  if (poly == 20) return bc::LogisticF<bc::KIND_BETALIK, 20>::eval(c, 0, 0, mp);
  if (poly == bc::kPowPolyMax) return bc::LogisticF<bc::KIND_BETALIK, bc::kPowPolyMax>::eval(c, 0, 0, mp);
  return bc::LogisticF<bc::KIND_BETALIK, 0>::eval(c, 0, 0, mp);
}
// the 4-wide stage-interleaved form the tensor-core kernel uses (bc_project_q.cu)
void fm_logistic_v4(int kind, int poly, const double* c4, double beta, const double* q, double* out4) {
  bc::ModelParams mp;
  for (int i = 0; i < 8; ++i) mp.p[i] = 0.0;
  mp.p[0] = beta;
  mp.p[1] = (beta + 1.) / beta;
  mp.p[2] = 700.0;
  mp.p[3] = bc::fm_hilo2d(bc::fm_hi(700.0 / beta), 0);
  for (int i = 0; i <= bc::kPowPolyMax; ++i) mp.q[i] = q ? q[i] : 0.0;
  double c[4] = {c4[0], c4[1], c4[2], c4[3]}, ca[4] = {0, 0, 0, 0}, o[4];
  if (kind == bc::KIND_LOGLIK) bc::LogisticF<bc::KIND_LOGLIK, 0>::evalv<4>(c, 0, ca, mp, bc::LogisticF<bc::KIND_LOGLIK, 0>::Tabs(), o);
  else if (poly == 20) bc::LogisticF<bc::KIND_BETALIK, 20>::evalv<4>(c, 0, ca, mp, bc::LogisticF<bc::KIND_BETALIK, 20>::Tabs(), o);
  else if (poly == bc::kPowPolyMax) bc::LogisticF<bc::KIND_BETALIK, bc::kPowPolyMax>::evalv<4>(c, 0, ca, mp, bc::LogisticF<bc::KIND_BETALIK, bc::kPowPolyMax>::Tabs(), o);
  else bc::LogisticF<bc::KIND_BETALIK, 0>::evalv<4>(c, 0, ca, mp, bc::LogisticF<bc::KIND_BETALIK, 0>::Tabs(), o);
  for (int i = 0; i < 4; ++i) out4[i] = o[i];
}
// the lane-table form (LogisticF<KIND_BETALIK, kPowTab>): w = Q's coefficients, rs / us = the 32-entry tables of bc_fit_pow_tab
void fm_logistic_v4_tab(const double* c4, double beta, const double* w, const double* rs, const double* us, double* out4) {
  bc::ModelParams mp;
  for (int i = 0; i < 8; ++i) mp.p[i] = 0.0;
  mp.p[0] = beta;
  mp.p[1] = (beta + 1.) / beta;
  mp.p[2] = 700.0;
  mp.p[3] = bc::fm_hilo2d(bc::fm_hi(700.0 / beta), 0);
  for (int i = 0; i <= bc::kPowTabDeg; ++i) mp.w[i] = w[i];
  typedef bc::LogisticF<bc::KIND_BETALIK, bc::kPowTab> F;
  F::Tabs T;
  T.e2.t = bc::kExp2Tab32;
  T.rs.t = rs;
  T.us.t = us;
  double c[4] = {c4[0], c4[1], c4[2], c4[3]}, ca[4] = {0, 0, 0, 0}, o[4];
  F::evalv<4>(c, 0, ca, mp, T, o);
  for (int i = 0; i < 4; ++i) out4[i] = o[i];
}
// lane-table form of the logistic log-likelihood (static tables)
void fm_logistic_loglik_v4_tab(const double* c4, double* out4) {
  bc::ModelParams mp;
  for (int i = 0; i < 8; ++i) mp.p[i] = 0.0;
  typedef bc::LogisticF<bc::KIND_LOGLIK, bc::kPowTab> F;
  F::Tabs T;
  T.e2.t = bc::kExp2Tab32;
  T.rs.t = bc::kRcpTab32;
  T.us.t = bc::kLogTab32;
  double c[4] = {c4[0], c4[1], c4[2], c4[3]}, ca[4] = {0, 0, 0, 0}, o[4];
  F::evalv<4>(c, 0, ca, mp, T, o);
  for (int i = 0; i < 4; ++i) out4[i] = o[i];
}
// the degree-5 form (weighted minimax q)
double fm_exp_tab5(double x, int lo) {
  bc::LaneTab32 T;
  T.t = bc::kExp2Tab32;
  const double a[1] = {x};
  double y[1];
  if (lo) bc::exp_tab_v<1, 0, 3>(a, T, y);
  else bc::exp_tab_v<1, 1, 3>(a, T, y);
  return y[0];
}
// exp_tab_v: e^x for -700 <= x <= 700, one-step (lo = 0) or two-step (lo = 1) reduction
double fm_exp_tab(double x, int lo) {
  bc::LaneTab32 T;
  T.t = bc::kExp2Tab32;
  const double a[1] = {x};
  double y[1];
  if (lo) bc::exp_tab_v<1, 0>(a, T, y);
  else bc::exp_tab_v<1, 1>(a, T, y);
  return y[0];
}
// evalv<4> of the Gaussian / neural-linear potentials (table exponential), element 0 returned
double fm_gaussian_v(int kind, double c, double ra, double ca, const double* p8) {
  bc::ModelParams mp;
  for (int i = 0; i < 8; ++i) mp.p[i] = p8[i];
  const double cc[4] = {c, c, c, c}, caa[4] = {ca, ca, ca, ca};
  double o[4];
  if (kind == bc::KIND_LOGLIK) {
    bc::GaussianF<bc::KIND_LOGLIK>::Tabs T; T.e2.t = bc::kExp2Tab32;
    bc::GaussianF<bc::KIND_LOGLIK>::evalv<4>(cc, ra, caa, mp, T, o);
  } else if (kind == bc::KIND_BETALIK) {
    bc::GaussianF<bc::KIND_BETALIK>::Tabs T; T.e2.t = bc::kExp2Tab32;
    bc::GaussianF<bc::KIND_BETALIK>::evalv<4>(cc, ra, caa, mp, T, o);
  } else {
    bc::GaussianF<bc::KIND_BETAGRAD>::Tabs T; T.e2.t = bc::kExp2Tab32;
    bc::GaussianF<bc::KIND_BETAGRAD>::evalv<4>(cc, ra, caa, mp, T, o);
  }
  return o[0];
}
double fm_neurlin_v(int kind, double c, double y, const double* p8) {
  bc::ModelParams mp;
  for (int i = 0; i < 8; ++i) mp.p[i] = p8[i];
  const double cc[4] = {c, c, c, c}, caa[4] = {0, 0, 0, 0};
  double o[4];
  if (kind == bc::KIND_LOGLIK) {
    bc::NeurlinF<bc::KIND_LOGLIK>::Tabs T; T.e2.t = bc::kExp2Tab32;
    bc::NeurlinF<bc::KIND_LOGLIK>::evalv<4>(cc, y, caa, mp, T, o);
  } else {
    bc::NeurlinF<bc::KIND_BETALIK>::Tabs T; T.e2.t = bc::kExp2Tab32;
    bc::NeurlinF<bc::KIND_BETALIK>::evalv<4>(cc, y, caa, mp, T, o);
  }
  return o[0];
}
double fm_gaussian(int kind, double c, double ra, double ca, const double* p8) {
  bc::ModelParams mp;
  for (int i = 0; i < 8; ++i) mp.p[i] = p8[i];
  if (kind == bc::KIND_LOGLIK) return bc::GaussianF<bc::KIND_LOGLIK>::eval(c, ra, ca, mp);
  if (kind == bc::KIND_BETALIK) return bc::GaussianF<bc::KIND_BETALIK>::eval(c, ra, ca, mp);
  return bc::GaussianF<bc::KIND_BETAGRAD>::eval(c, ra, ca, mp);
}
double fm_neurlin(int kind, double c, double y, const double* p8) {
  bc::ModelParams mp;
  for (int i = 0; i < 8; ++i) mp.p[i] = p8[i];
  if (kind == bc::KIND_LOGLIK) return bc::NeurlinF<bc::KIND_LOGLIK>::eval(c, y, 0, mp);
  return bc::NeurlinF<bc::KIND_BETALIK>::eval(c, y, 0, mp);
}
}
