// beta-cores B200: extern "C" entry points (see include/betacores.h for the contract).
#include <math.h>
#include <string.h>
#include <new>
#include "../../include/betacores.h"
#include "bc_kernels.h"
#include "bc_umma.cuh"

using namespace bc;

#include <atomic>
static thread_local int g_last_cuda = 0;
static std::atomic<long long> g_launches{0};   // kernels launched by this library (bench.py reports it)
#define BC_LAUNCHED(n) g_launches.fetch_add((n), std::memory_order_relaxed)

struct bc_ctx {
  int device = 0, sms = 0;
  // potential
  bool potential_set = false;
  int model = 0, kind = 0, Dk = 0, Dc = 0, Dpad = 0, ss = 0, tile_cfg = 0, BM = 0, BN = 0;
  int poly = 0;  // logistic beta-likelihood: degree of the (1+t)^-beta polynomial in use (0 = exp/log1p form)
  size_t smem = 0;
  ModelParams mp{};
  // lane-table form of the logistic beta-likelihood (tensor-core route): 64 doubles per beta, written once into a slot of
  // `tab_dev` and never modified afterwards (a kernel in flight keeps reading its slot whatever bc_set_potential does next)
  static constexpr int kTabSlots = 16;
  double* tab_dev = nullptr;
  double tab_beta[kTabSlots] = {0};
  int tab_count = 0;
  const double* tab_cur = nullptr;   // slot of the current potential's tables (logistic beta-likelihood), or null
  bool tab_form = false;             // the tensor-core kernels run the current potential in its lane-table form
  bool tab_off = false;              // more distinct betas than slots have come by: polynomial forms from here on
  int pot_form = 0;                  // bc_set_potential_form: 0 = fastest form available, 1 = polynomial forms only
  const double* d_siginv = nullptr;
  double* siginvT = nullptr;   // transposed copy of d_siginv (coalesced row walks in k_prepare_rows), refreshed after every bc_set_potential
  size_t cap_sT = 0;
  bool siginvT_ready = false;
  double* nnls_rho = nullptr;  // S-length residual scratch of bc_nnls
  size_t cap_rho = 0;
  int* siginv_diag = nullptr;  // device flag: d_siginv is a diagonal matrix (k_prepare_rows then takes one term per dot product)
  // samples
  bool samples_set = false;
  int S = 0;
  double* B = nullptr;  // S x Dpad
  double* colaux = nullptr;
  double* bbar = nullptr;
  size_t capB = 0, capS = 0, capD = 0;
  // per-CTA partials
  double* part_colsum = nullptr;
  size_t cap_part = 0;
  double* part_misc = nullptr;
  double* dense_part = nullptr;
  size_t cap_dense = 0;
  // tensor-core route: quantised sample image + per-sample scale (valid when q_ready)
  bool q_ready = false;
  double* qB = nullptr;  // raw bytes, allocated in doubles
  size_t cap_qB = 0;
  double* colscale = nullptr;
  size_t cap_cs = 0;
  // feature exponents of the row image in use (bc_set_feature_exponents); the sample image is built with their negatives
  bool fexp_on = false;
  int* fexp = nullptr;                     // kQK ints
  unsigned long long* fscratch = nullptr;  // kQK column maxima (bc_feature_exponents)
  unsigned long long* sscratch = nullptr;  // two slots of 2 words: max bits / non-finite flag of a sample set; a call uses slot
  int sslot = 0;                           // `sslot` and its quantise kernel clears the other one for the next call
  int q_digits = 6;                        // leading digits of the 7-digit images a launch contracts (bc_set_contraction_digits)
  // bc_greedy_opt_step, single-part jobs: where the next column-sum pass also leaves the centred column sum (its finalize
  // kernel then does the work of bc_colsum_combine); consumed by that pass
  double* fuse_colsum_out = nullptr;
};

static int cuda_fail(cudaError_t e) {
  g_last_cuda = (int)e;
  return BC_ERR_CUDA;
}
#define BC_CUDA(x)                                \
  do {                                            \
    cudaError_t e__ = (x);                        \
    if (e__ != cudaSuccess) return cuda_fail(e__); \
  } while (0)

static int grow(double** p, size_t* cap, size_t need) {
  if (need <= *cap) return BC_OK;
  if (*p) BC_CUDA(cudaFree(*p));
  *p = nullptr;
  *cap = 0;
  BC_CUDA(cudaMalloc((void**)p, need * sizeof(double)));
  *cap = need;
  return BC_OK;
}

// (1+t)^-beta on t in [0,1] as a polynomial in x = 2t-1: Chebyshev interpolant at 64 nodes, truncated to `deg`,
// converted to the monomial basis (coefficients decay like 3^-k, so Horner in x is well conditioned); all in
// long double.  out: deg+1 coefficients, highest degree first.  Returns the truncation bound sum_{j>deg} |c_j|.
static double fit_pow_poly(double beta, int deg, double* out) {
  const int n = 64;
  const long double pi = 3.14159265358979323846264338327950288L;
  long double fs[n], c[n];
  for (int k = 0; k < n; ++k) {
    const long double x = cosl(pi * (k + 0.5L) / n);
    fs[k] = powl(1.0L + 0.5L * (x + 1.0L), -(long double)beta);
  }
  for (int j = 0; j < n; ++j) {
    long double a = 0.0L;
    for (int k = 0; k < n; ++k) a += fs[k] * cosl(j * pi * (k + 0.5L) / n);
    c[j] = 2.0L * a / n;
  }
  c[0] *= 0.5L;
  long double err = 0.0L;
  for (int j = deg + 1; j < n; ++j) err += fabsl(c[j]);
  long double Tm[kPowPolyMax + 1] = {0}, Tc[kPowPolyMax + 1] = {0}, Tn[kPowPolyMax + 1], mono[kPowPolyMax + 1] = {0};
  Tm[0] = 1.0L;  // T_0
  Tc[1] = 1.0L;  // T_1
  mono[0] += c[0];
  if (deg >= 1) mono[1] += c[1];
  for (int j = 2; j <= deg; ++j) {
    for (int i = 0; i <= deg; ++i) Tn[i] = (i > 0 ? 2.0L * Tc[i - 1] : 0.0L) - Tm[i];
    for (int i = 0; i <= deg; ++i) {
      mono[i] += c[j] * Tn[i];
      Tm[i] = Tc[i];
      Tc[i] = Tn[i];
    }
  }
  for (int i = 0; i <= deg; ++i) out[i] = (double)mono[deg - i];
  return (double)err;
}

// Lane-table form of (1+t)^-beta (bc_models.cuh, LogisticF<KIND_BETALIK, kPowTab>): s = 1 + t in [1, 2] is cut into 32
// intervals with centres s_j = 1 + (j + 1/2)/32;  s^-beta = s_j^-beta (1 + w)^-beta,  w = (s - s_j)/s_j,  |w| <= 1/65,
// (1 + w)^-beta = 1 + w Q(w).  rs[j] = 1/s_j and us[j] = s_j^-beta correctly rounded; Q = the interpolant of
// ((1+w)^-beta - 1)/w at kPowTabDeg+1 Chebyshev nodes, monomial coefficients highest degree first; all in long double.
// Returns max |w Q(w) - ((1+w)^-beta - 1)| over the interval for the ROUNDED coefficients.
static double fit_pow_tab(double beta, double* w_out, double* rs, double* us) {
  const long double b = (long double)beta;
  for (int j = 0; j < 32; ++j) {
    const long double sj = 1.0L + ((long double)j + 0.5L) / 32.0L;
    rs[j] = (double)(1.0L / sj);
    us[j] = (double)powl(sj, -b);
  }
  constexpr int n = kPowTabDeg + 1;
  const long double pi = 3.14159265358979323846264338327950288L;
  const long double wmax = (1.0L / 65.0L) * 1.0005L;
  auto Qx = [&](long double w) -> long double {
    if (fabsl(w) < 1e-9L) return -b + b * (b + 1.0L) * 0.5L * w;
    return expm1l(-b * log1pl(w)) / w;
  };
  long double A[n][n + 1];   // Vandermonde in x = w / wmax
  for (int i = 0; i < n; ++i) {
    const long double x = cosl(pi * (i + 0.5L) / n);
    long double pw = 1.0L;
    for (int k = 0; k < n; ++k, pw *= x) A[i][k] = pw;
    A[i][n] = Qx(x * wmax);
  }
  for (int col = 0; col < n; ++col) {   // Gauss-Jordan with partial pivoting
    int piv = col;
    for (int r = col + 1; r < n; ++r)
      if (fabsl(A[r][col]) > fabsl(A[piv][col])) piv = r;
    for (int k = 0; k <= n; ++k) {
      const long double t = A[col][k];
      A[col][k] = A[piv][k];
      A[piv][k] = t;
    }
    for (int r = 0; r < n; ++r) {
      if (r == col) continue;
      const long double f = A[r][col] / A[col][col];
      for (int k = col; k <= n; ++k) A[r][k] -= f * A[col][k];
    }
  }
  long double sc = 1.0L;
  for (int k = 0; k < n; ++k, sc *= wmax) w_out[kPowTabDeg - k] = (double)(A[k][n] / A[k][k] / sc);
  long double err = 0.0L;
  for (int i = 0; i <= 1024; ++i) {
    const long double w = -wmax + 2.0L * wmax * i / 1024.0L;
    long double q = 0.0L;
    for (int k = 0; k < n; ++k) q = q * w + (long double)w_out[k];
    const long double e = fabsl(w * q - expm1l(-b * log1pl(w)));
    if (e > err) err = e;
  }
  return (double)err;
}

extern "C" {

int bc_version(void) { return 102; }

const char* bc_error_string(int code) {
  switch (code) {
    case BC_OK: return "ok";
    case BC_ERR_ARG: return "bad argument";
    case BC_ERR_ALIGN: return "alignment requirement violated (16-byte base pointer, even leading dimension)";
    case BC_ERR_UNSUPPORTED: return "shape not supported by the kernels (feature dimension too large for one shared-memory tile)";
    case BC_ERR_STATE: return "call order: bc_set_potential / bc_set_samples first";
    case BC_ERR_CUDA: return "CUDA runtime error (see bc_last_cuda_error)";
    default: return "unknown error";
  }
}

int bc_last_cuda_error(void) { return g_last_cuda; }

int bc_fit_pow_poly(double beta, int degree, double* h_coef, double* h_err) {
  if (!(beta > 0.0) || degree < 1 || degree > kPowPolyMax || !h_coef) return BC_ERR_ARG;
  const double e = fit_pow_poly(beta, degree, h_coef);
  if (h_err) *h_err = e;
  return BC_OK;
}
int bc_fit_pow_tab(double beta, double* h_w, double* h_rs, double* h_us, double* h_err) {
  if (!(beta > 0.0) || !h_w || !h_rs || !h_us) return BC_ERR_ARG;
  const double e = fit_pow_tab(beta, h_w, h_rs, h_us);
  if (h_err) *h_err = e;
  return BC_OK;
}
int bc_set_potential_form(bc_ctx* c, int form) {
  if (!c || form < 0 || form > 1) return BC_ERR_ARG;
  c->pot_form = form;
  c->potential_set = false;   // takes effect with the next bc_set_potential
  return BC_OK;
}
int bc_potential_form(const bc_ctx* c) { return (c && c->tab_form) ? 2 : 1; }
int64_t bc_launch_count(void) { return (int64_t)g_launches.load(std::memory_order_relaxed); }

int bc_create(int device, bc_ctx** out) {
  if (!out) return BC_ERR_ARG;
  BC_CUDA(cudaSetDevice(device));
  bc_ctx* c = new (std::nothrow) bc_ctx();
  if (!c) return BC_ERR_ARG;
  c->device = device;
  cudaDeviceProp prop;
  BC_CUDA(cudaGetDeviceProperties(&prop, device));
  c->sms = prop.multiProcessorCount;
  BC_CUDA(cudaMalloc((void**)&c->part_misc, (size_t)4 * 1024 * sizeof(double)));
  BC_CUDA(cudaMalloc((void**)&c->fexp, kQK * sizeof(int)));
  BC_CUDA(cudaMemset(c->fexp, 0, kQK * sizeof(int)));
  BC_CUDA(cudaMalloc((void**)&c->fscratch, kQK * sizeof(unsigned long long)));
  BC_CUDA(cudaMalloc((void**)&c->sscratch, 4 * sizeof(unsigned long long)));
  BC_CUDA(cudaMemset(c->sscratch, 0, 4 * sizeof(unsigned long long)));
  BC_CUDA(cudaMalloc((void**)&c->tab_dev, (size_t)bc_ctx::kTabSlots * 64 * sizeof(double)));
  *out = c;
  return BC_OK;
}

int bc_destroy(bc_ctx* c) {
  if (!c) return BC_OK;
  cudaFree(c->B);
  cudaFree(c->siginvT);
  cudaFree(c->siginv_diag);
  cudaFree(c->nnls_rho);
  cudaFree(c->colaux);
  cudaFree(c->bbar);
  cudaFree(c->part_colsum);
  cudaFree(c->part_misc);
  cudaFree(c->dense_part);
  cudaFree(c->qB);
  cudaFree(c->colscale);
  cudaFree(c->fexp);
  cudaFree(c->fscratch);
  cudaFree(c->sscratch);
  cudaFree(c->tab_dev);
  delete c;
  return BC_OK;
}

int bc_sm_count(const bc_ctx* c) { return c ? c->sms : 0; }
int bc_colsum_ld(int S) { return S + 1; }

int bc_set_potential(bc_ctx* c, int model, int kind, int D, const double* h_params, const double* d_siginv) {
  if (!c || D <= 0 || !h_params) return BC_ERR_ARG;
  if (model < 0 || model > 2 || kind < 0 || kind > 2) return BC_ERR_ARG;
  if (kind == BC_KIND_BETAGRAD && model != BC_MODEL_GAUSSIAN) return BC_ERR_UNSUPPORTED;
  if (model == BC_MODEL_GAUSSIAN && !d_siginv) return BC_ERR_ARG;
  const int extra = (model == BC_MODEL_NEURLIN) ? 1 : 0;   // y rides along as column D
  const int Dc = ((D + extra + 1) / 2) * 2;                 // bulk-copy width (16-byte multiple)
  const int Dpad = ((Dc + 3) / 4) * 4;                      // contraction length (DMMA k = 4)
  int BM, BN, ss;
  size_t smem;
  const int cfg = project_tile_config(Dpad, &BM, &BN, &ss, &smem);
  if (cfg < 0) return BC_ERR_UNSUPPORTED;
  c->model = model;
  c->kind = kind;
  c->Dk = D;
  c->Dc = Dc;
  c->Dpad = Dpad;
  c->ss = ss;
  c->tile_cfg = cfg;
  c->BM = BM;
  c->BN = BN;
  c->smem = smem;
  for (int i = 0; i < 8; ++i) c->mp.p[i] = h_params[i];
  for (int i = 0; i <= kPowPolyMax; ++i) c->mp.q[i] = 0.0;
  c->poly = 0;
  if (model == BC_MODEL_LOGISTIC && kind == BC_KIND_BETALIK) {
    const double beta = h_params[0];
    if (!(beta > 0.0)) return BC_ERR_ARG;
    // clamps of |m| ahead of the two exponentials (LogisticF::evalv): 700 for e^-a, 700 / beta for e^(-beta a).  The low
    // words are zero so that the lane-table form can compare high words on the integer pipe.
    {
      const double bmax = 700.0 / beta;
      unsigned long long bits;
      memcpy(&bits, &bmax, 8);
      bits &= 0xffffffff00000000ull;
      memcpy(&c->mp.p[3], &bits, 8);
      c->mp.p[2] = 700.0;
    }
    const int degs[2] = {20, kPowPolyMax};
    for (int i = 0; i < 2 && c->poly == 0; ++i) {
      double q[kPowPolyMax + 1];
      if (fit_pow_poly(beta, degs[i], q) < 2.5e-17) {  // below a quarter ulp of (1+t)^-beta in [2^-beta, 1]
        for (int k = 0; k <= degs[i]; ++k) c->mp.q[kPowPolyMax - degs[i] + k] = q[k];
        c->poly = degs[i];
      }
    }
    c->tab_cur = nullptr;
    double w[kPowTabDeg + 1], tabs[64];
    if (c->pot_form == 0 && !c->tab_off && fit_pow_tab(beta, w, tabs, tabs + 32) < 5e-17) {   // under half an ulp of (1+t)^-beta in [1/2, 1]: beta up to about 1
      int slot = -1;
      for (int i = 0; i < c->tab_count; ++i)
        if (c->tab_beta[i] == beta) slot = i;
      if (slot < 0) {
        if (c->tab_count == bc_ctx::kTabSlots) {
          // every slot taken: this workspace sees a new beta all the time (learn_beta = True moves it every optimiser step).
          // Fitting and uploading tables per step, and waiting for the kernels that still read old slots, costs more than the
          // tables save: from here on the workspace stays with the polynomial forms.
          c->tab_off = true;
        }
      }
      if (c->tab_off) {
        slot = -1;
      } else if (slot < 0) {
        slot = c->tab_count++;
        c->tab_beta[slot] = beta;
        BC_CUDA(cudaMemcpy(c->tab_dev + (size_t)slot * 64, tabs, sizeof(tabs), cudaMemcpyHostToDevice));
      }
      if (slot >= 0) {
        for (int k = 0; k <= kPowTabDeg; ++k) c->mp.w[k] = w[k];
        c->tab_cur = c->tab_dev + (size_t)slot * 64;
      }
    }
    c->tab_form = c->tab_cur != nullptr;
  } else {
    c->tab_cur = nullptr;
    // the logistic log-likelihood's tables hold no model constant (bc_fastmath.cuh); Gaussian / neural-linear potentials take
    // their exponential from the lane table in every form
    c->tab_form = (model == BC_MODEL_LOGISTIC && kind == BC_KIND_LOGLIK && c->pot_form == 0);
  }
  c->d_siginv = d_siginv;
  c->siginvT_ready = false;
  c->potential_set = true;
  c->samples_set = false;
  return BC_OK;
}

int bc_set_samples(bc_ctx* c, const double* d_theta, int S, int ldt, void* stream) {
  if (!c || !d_theta || S <= 0 || ldt < c->Dk) return BC_ERR_ARG;
  if (!c->potential_set) return BC_ERR_STATE;
  int rc;
  if ((rc = grow(&c->B, &c->capB, (size_t)S * c->Dpad))) return rc;
  if ((rc = grow(&c->colaux, &c->capS, (size_t)S))) return rc;
  if ((rc = grow(&c->bbar, &c->capD, (size_t)c->Dpad + 1))) return rc;
  if ((rc = grow(&c->part_colsum, &c->cap_part, (size_t)c->sms * 2 * bc_colsum_ld(S)))) return rc;
  c->S = S;
  if (c->model == BC_MODEL_GAUSSIAN && !c->siginvT_ready) {
    if ((rc = grow(&c->siginvT, &c->cap_sT, (size_t)c->Dk * c->Dk))) return rc;
    BC_CUDA(launch_transpose(c->d_siginv, c->Dk, c->Dk, c->Dk, c->siginvT, c->Dk, (cudaStream_t)stream));
    if (!c->siginv_diag) BC_CUDA(cudaMalloc((void**)&c->siginv_diag, sizeof(int)));
    BC_CUDA(launch_offdiag_test(c->d_siginv, c->Dk, c->siginv_diag, (cudaStream_t)stream));
    BC_LAUNCHED(2);
    c->siginvT_ready = true;
  }
  const bool q_route = c->Dk <= kQK;
  if (q_route) c->sslot ^= 1;
  BC_CUDA(launch_prepare_samples(c->model, d_theta, S, c->Dk, ldt, c->d_siginv, c->model == BC_MODEL_GAUSSIAN ? c->siginvT : nullptr, c->B,
                                 c->Dpad, c->colaux, c->bbar, c->fexp_on ? c->fexp : nullptr, q_route ? c->sscratch + 2 * c->sslot : nullptr,
                                 c->model == BC_MODEL_GAUSSIAN ? c->siginv_diag : nullptr, (cudaStream_t)stream));
  BC_LAUNCHED(2);
  c->q_ready = false;
  if (c->Dk <= kQK) {
    const size_t chunks = (size_t)(S + kQChunk - 1) / kQChunk;
    if ((rc = grow(&c->qB, &c->cap_qB, chunks * kQChunkBytes / sizeof(double)))) return rc;
    if ((rc = grow(&c->colscale, &c->cap_cs, 2))) return rc;   // [0] = scale, [1] = the shared exponent (int)
    BC_CUDA(launch_quantise_samples(c->B, c->Dpad, S, c->Dk, reinterpret_cast<unsigned char*>(c->qB), c->colscale,
                                    reinterpret_cast<int*>(c->colscale + 1), c->fexp_on ? c->fexp : nullptr, c->sscratch, c->sslot, true,
                                    (cudaStream_t)stream));
    BC_LAUNCHED(1);
    c->q_ready = true;
  }
  c->samples_set = true;
  return BC_OK;
}

int bc_feature_exponents(bc_ctx* c, const double* d_X, int64_t ldx, int64_t n, int D, int32_t* d_fexp_out, void* stream) {
  if (!c || !d_X || !d_fexp_out || n < 0 || D <= 0 || ldx < D) return BC_ERR_ARG;
  if (D > kQK) return BC_ERR_UNSUPPORTED;
  BC_CUDA(launch_feature_exponents(d_X, ldx, n, D, c->fscratch, d_fexp_out, (cudaStream_t)stream));
  BC_LAUNCHED(n > 0 ? 2 : 1);
  return BC_OK;
}

int bc_set_feature_exponents(bc_ctx* c, const int32_t* d_fexp, int D, void* stream) {
  if (!c || D < 0 || D > kQK || (d_fexp && D == 0)) return BC_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  BC_CUDA(cudaMemsetAsync(c->fexp, 0, kQK * sizeof(int), st));
  if (d_fexp) BC_CUDA(cudaMemcpyAsync(c->fexp, d_fexp, (size_t)D * sizeof(int), cudaMemcpyDeviceToDevice, st));
  c->fexp_on = d_fexp != nullptr;
  if (c->samples_set && c->q_ready) {   // the sample image in place was built for other exponents
    c->sslot ^= 1;
    BC_CUDA(launch_quantise_samples(c->B, c->Dpad, c->S, c->Dk, reinterpret_cast<unsigned char*>(c->qB), c->colscale,
                                    reinterpret_cast<int*>(c->colscale + 1), c->fexp_on ? c->fexp : nullptr, c->sscratch, c->sslot, false, st));
    BC_LAUNCHED(2);
  }
  return BC_OK;
}

int bc_rowquad(bc_ctx* c, const double* d_X, int64_t n, int64_t ldx, double* d_out, void* stream) {
  if (!c || !d_X || !d_out || n < 0) return BC_ERR_ARG;
  if (!c->potential_set || c->model != BC_MODEL_GAUSSIAN) return BC_ERR_STATE;
  BC_CUDA(launch_rowquad(d_X, n, c->Dk, ldx, c->d_siginv, d_out, (cudaStream_t)stream));
  BC_LAUNCHED(1);
  return BC_OK;
}

static int project_common(bc_ctx* c, int mode, const double* d_X, int64_t ldx, const int64_t* d_rows, int64_t n,
                          const double* d_rowaux, ProjArgs& P, int* grid, int* cfg, size_t* smem) {
  if (!c || !d_X || n < 0) return BC_ERR_ARG;
  if (!c->potential_set || !c->samples_set) return BC_ERR_STATE;
  if ((reinterpret_cast<uintptr_t>(d_X) & 15) || (ldx & 1) || ldx < c->Dc) return BC_ERR_ALIGN;
  if (c->model == BC_MODEL_GAUSSIAN && !d_rowaux) return BC_ERR_ARG;
  P.A = d_X;
  P.lda = ldx;
  P.rows = reinterpret_cast<const long long*>(d_rows);
  P.n = n;
  P.idx_offset = 0;
  P.B = c->B;
  P.ldb = c->Dpad;
  P.S = c->S;
  P.Dk = c->Dk;
  P.Dc = c->Dc;
  P.Dpad = c->Dpad;
  P.ss = c->ss;
  P.colaux = (c->model == BC_MODEL_GAUSSIAN) ? c->colaux : nullptr;
  P.rowaux = (c->model == BC_MODEL_GAUSSIAN) ? d_rowaux : nullptr;
  P.bbar = c->bbar;
  P.mp = c->mp;
  P.part_colsum = c->part_colsum;
  P.part_misc = c->part_misc;
  P.Sld = bc_colsum_ld(c->S);
  P.resid = nullptr;
  P.scores = nullptr;
  P.V = nullptr;
  P.ldv = 0;
  P.norms = nullptr;
  P.raw = 0;
  P.want_colsum = 0;
  *cfg = project_tile_config_for_rows(c->tile_cfg, n, c->sms, c->Dpad, smem);
  const int bm = project_tile_rows(*cfg);
  const int64_t tiles = (n + bm - 1) / bm;
  *grid = (int)(tiles < c->sms ? tiles : c->sms);
  (void)mode;
  return BC_OK;
}

int bc_project_colsum(bc_ctx* c, const double* d_X, int64_t ldx, const int64_t* d_rows, int64_t n, const double* d_rowaux,
                      double* d_out_dd, void* stream) {
  ProjArgs P;
  int grid, rc, cfg;
  size_t smem;
  if (!d_out_dd) return BC_ERR_ARG;
  if ((rc = project_common(c, MODE_COLSUM, d_X, ldx, d_rows, n, d_rowaux, P, &grid, &cfg, &smem))) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  if (n == 0) {
    BC_CUDA(cudaMemsetAsync(d_out_dd, 0, sizeof(double) * 2 * P.Sld, st));
    return BC_OK;
  }
  BC_CUDA(launch_project(P, c->model, c->kind, c->poly, MODE_COLSUM, cfg, grid, smem, st));
  double* fused = c->fuse_colsum_out;
  c->fuse_colsum_out = nullptr;
  BC_CUDA(launch_project_finalize(c->part_colsum, c->part_misc, grid, c->S, P.Sld, d_out_dd, nullptr, MODE_COLSUM, st, fused));
  BC_LAUNCHED(2);
  return BC_OK;
}

int bc_project_score(bc_ctx* c, const double* d_X, int64_t ldx, const int64_t* d_rows, int64_t n, const double* d_rowaux,
                     const double* d_resid, int64_t idx_offset, double* d_best, double* d_scores, void* stream) {
  ProjArgs P;
  int grid, rc, cfg;
  size_t smem;
  if (!d_resid || !d_best || n <= 0) return BC_ERR_ARG;
  if ((rc = project_common(c, MODE_SCORE, d_X, ldx, d_rows, n, d_rowaux, P, &grid, &cfg, &smem))) return rc;
  P.resid = d_resid;
  P.scores = d_scores;
  P.idx_offset = idx_offset;
  cudaStream_t st = (cudaStream_t)stream;
  BC_CUDA(launch_project(P, c->model, c->kind, c->poly, MODE_SCORE, cfg, grid, smem, st));
  BC_CUDA(launch_project_finalize(c->part_colsum, c->part_misc, grid, c->S, P.Sld, nullptr, d_best, MODE_SCORE, st));
  BC_LAUNCHED(2);
  return BC_OK;
}

int bc_project_materialise(bc_ctx* c, const double* d_X, int64_t ldx, const int64_t* d_rows, int64_t n, const double* d_rowaux,
                           double* d_V, int64_t ldv, double* d_norms, double* d_out_dd, int raw, void* stream) {
  ProjArgs P;
  int grid, rc, cfg;
  size_t smem;
  if (!d_V) return BC_ERR_ARG;
  if ((rc = project_common(c, MODE_MATERIALISE, d_X, ldx, d_rows, n, d_rowaux, P, &grid, &cfg, &smem))) return rc;
  if (ldv < c->S) return BC_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  if (n == 0) {
    if (d_out_dd) BC_CUDA(cudaMemsetAsync(d_out_dd, 0, sizeof(double) * 2 * P.Sld, st));
    return BC_OK;
  }
  P.V = d_V;
  P.ldv = ldv;
  P.norms = raw ? nullptr : d_norms;
  P.raw = raw ? 1 : 0;
  P.want_colsum = d_out_dd ? 1 : 0;
  // Few rows (the M coreset points of every optimiser step): one 16-row warp would walk all S samples alone.  Split the
  // sample range over CTAs instead -- each writes raw potentials for its chunks -- and centre the rows in a second
  // small kernel, the reference's own arithmetic (projector.py:26,55: bls -= bls.mean(axis=1)).
  const int64_t tiles = (n + project_tile_rows(cfg) - 1) / project_tile_rows(cfg);
  if (tiles * 4 <= c->sms && !d_out_dd && (raw || !d_norms)) {
    int csplit = (int)(c->sms / tiles);
    if (csplit > project_chunks(c->S)) csplit = project_chunks(c->S);
    if (csplit > 1) {
      P.raw = 1;
      P.norms = nullptr;
      BC_CUDA(launch_project(P, c->model, c->kind, c->poly, MODE_MATERIALISE, cfg, grid, smem, st, csplit));
      BC_LAUNCHED(1);
      if (!raw) {
        BC_CUDA(launch_dense_center(d_V, n, c->S, ldv, st));
        BC_LAUNCHED(1);
      }
      return BC_OK;
    }
  }
  BC_CUDA(launch_project(P, c->model, c->kind, c->poly, MODE_MATERIALISE, cfg, grid, smem, st));
  BC_LAUNCHED(1);
  if (d_out_dd) {
    BC_CUDA(launch_project_finalize(c->part_colsum, c->part_misc, grid, c->S, P.Sld, d_out_dd, nullptr, MODE_MATERIALISE, st));
    BC_LAUNCHED(1);
  }
  return BC_OK;
}

// ---- tensor-core route ------------------------------------------------------------------------
int bc_q_max_features(void) { return kQK; }

int bc_set_contraction_digits(bc_ctx* c, int digits) {
  if (!c || digits < 5 || digits > kQSlices) return BC_ERR_ARG;
  c->q_digits = digits;
  return BC_OK;
}
int bc_contraction_digits(const bc_ctx* c) { return c ? c->q_digits : 0; }

int bc_q_image_bytes(int64_t n, int64_t* bytes) {
  if (n < 0 || !bytes) return BC_ERR_ARG;
  *bytes = ((n + kQTileRows - 1) / kQTileRows) * (int64_t)kQTileBytes;
  return BC_OK;
}

int bc_quantise_rows(bc_ctx* c, const double* d_X, int64_t ldx, int64_t n, int D, int aux_col, void* d_image, double* d_rowscale,
                     double* d_aux_out, const int32_t* d_fexp, void* stream) {
  if (!c || !d_X || !d_image || !d_rowscale || n < 0 || D <= 0 || ldx < D) return BC_ERR_ARG;
  if (D > kQK) return BC_ERR_UNSUPPORTED;
  if (d_aux_out && (aux_col < 0 || aux_col >= ldx)) return BC_ERR_ARG;
  if (reinterpret_cast<uintptr_t>(d_image) & 15) return BC_ERR_ALIGN;
  BC_CUDA(launch_quantise_rows(d_X, ldx, n, D, reinterpret_cast<unsigned char*>(d_image), d_rowscale, d_aux_out, aux_col, d_fexp,
                               (cudaStream_t)stream));
  BC_LAUNCHED(1);
  return BC_OK;
}

int bc_q_gather_rows(bc_ctx* c, const void* d_image, const double* d_rowscale, const double* d_rowaux, const int64_t* d_idx, int64_t n,
                     void* d_image_out, double* d_rowscale_out, double* d_rowaux_out, void* stream) {
  if (!c || !d_image || !d_rowscale || !d_image_out || !d_rowscale_out || n < 0 || (n > 0 && !d_idx)) return BC_ERR_ARG;
  if ((d_rowaux_out != nullptr) != (d_rowaux != nullptr)) return BC_ERR_ARG;
  if ((reinterpret_cast<uintptr_t>(d_image) & 15) || (reinterpret_cast<uintptr_t>(d_image_out) & 15)) return BC_ERR_ALIGN;
  if (n == 0) return BC_OK;
  BC_CUDA(launch_gather_image(reinterpret_cast<const unsigned char*>(d_image), d_rowscale, d_rowaux, reinterpret_cast<const long long*>(d_idx), n,
                              reinterpret_cast<unsigned char*>(d_image_out), d_rowscale_out, d_rowaux_out, (cudaStream_t)stream));
  BC_LAUNCHED(1);
  return BC_OK;
}

static int q_common(bc_ctx* c, const void* d_image, const double* d_rowscale, int64_t n, const double* d_rowaux, QProjArgs& P,
                    int* grid) {
  if (!c || !d_image || !d_rowscale || n < 0) return BC_ERR_ARG;
  if (!c->potential_set || !c->samples_set) return BC_ERR_STATE;
  if (!c->q_ready) return BC_ERR_UNSUPPORTED;
  if (reinterpret_cast<uintptr_t>(d_image) & 15) return BC_ERR_ALIGN;
  if (c->model != BC_MODEL_LOGISTIC && !d_rowaux) return BC_ERR_ARG;
  P.imgA = reinterpret_cast<const unsigned char*>(d_image);
  P.rowscale = d_rowscale;
  P.imgB = reinterpret_cast<const unsigned char*>(c->qB);
  P.colscale = c->colscale;
  P.n = n;
  P.idx_offset = 0;
  P.S = c->S;
  P.ksteps = (c->Dk + 31) / 32;
  P.colaux = (c->model == BC_MODEL_GAUSSIAN) ? c->colaux : nullptr;
  P.rowaux = (c->model != BC_MODEL_LOGISTIC) ? d_rowaux : nullptr;
  P.mp = c->mp;
  P.pot_tabs = c->tab_cur;
  P.part_colsum = c->part_colsum;
  P.part_misc = c->part_misc;
  P.Sld = bc_colsum_ld(c->S);
  P.resid = nullptr;
  P.scores = nullptr;
  P.V = nullptr;
  P.ldv = 0;
  const int64_t tiles = (n + kQTileRows - 1) / kQTileRows;
  *grid = (int)(tiles < c->sms ? tiles : c->sms);
  return BC_OK;
}

int bc_project_colsum_q(bc_ctx* c, const void* d_image, const double* d_rowscale, int64_t n, const double* d_rowaux,
                        double* d_out_dd, void* stream) {
  QProjArgs P;
  int grid, rc, cfg;
  size_t smem;
  if (!d_out_dd) return BC_ERR_ARG;
  if ((rc = q_common(c, d_image, d_rowscale, n, d_rowaux, P, &grid))) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  if (n == 0) {
    BC_CUDA(cudaMemsetAsync(d_out_dd, 0, sizeof(double) * 2 * P.Sld, st));
    return BC_OK;
  }
  BC_CUDA(launch_project_q(P, c->model, c->kind, c->tab_form ? kPowTab : c->poly, QMODE_COLSUM, c->q_digits, grid, st));
  double* fused = c->fuse_colsum_out;
  c->fuse_colsum_out = nullptr;
  BC_CUDA(launch_project_finalize(c->part_colsum, c->part_misc, grid, c->S, P.Sld, d_out_dd, nullptr, MODE_COLSUM, st, fused));
  BC_LAUNCHED(2);
  return BC_OK;
}

int bc_project_score_q(bc_ctx* c, const void* d_image, const double* d_rowscale, int64_t n, const double* d_rowaux,
                       const double* d_resid, int64_t idx_offset, double* d_best, double* d_scores, void* stream) {
  QProjArgs P;
  int grid, rc, cfg;
  size_t smem;
  if (!d_resid || !d_best || n <= 0) return BC_ERR_ARG;
  if ((rc = q_common(c, d_image, d_rowscale, n, d_rowaux, P, &grid))) return rc;
  P.resid = d_resid;
  P.scores = d_scores;
  P.idx_offset = idx_offset;
  cudaStream_t st = (cudaStream_t)stream;
  BC_CUDA(launch_project_q(P, c->model, c->kind, c->tab_form ? kPowTab : c->poly, QMODE_SCORE, c->q_digits, grid, st));
  BC_CUDA(launch_project_finalize(c->part_colsum, c->part_misc, grid, c->S, P.Sld, nullptr, d_best, MODE_SCORE, st));
  BC_LAUNCHED(2);
  return BC_OK;
}

int bc_contraction_q(bc_ctx* c, const void* d_image, const double* d_rowscale, int64_t n, double* d_V, int64_t ldv, void* stream) {
  QProjArgs P;
  int grid, rc, cfg;
  size_t smem;
  static const double dummy = 0.0;
  if (!d_V || n <= 0) return BC_ERR_ARG;
  if ((rc = q_common(c, d_image, d_rowscale, n, &dummy, P, &grid))) return rc;
  if (ldv < c->S) return BC_ERR_ARG;
  P.rowaux = nullptr;
  P.V = d_V;
  P.ldv = ldv;
  BC_CUDA(launch_project_q(P, c->model, c->kind, c->tab_form ? kPowTab : c->poly, QMODE_DOT, c->q_digits, grid, (cudaStream_t)stream));
  BC_LAUNCHED(1);
  return BC_OK;
}

int bc_colsum_combine(bc_ctx* c, const double* d_parts, int nparts, int S, double* d_out, void* stream) {
  if (!c || !d_parts || !d_out || nparts <= 0 || S <= 0) return BC_ERR_ARG;
  BC_CUDA(launch_colsum_combine(d_parts, nparts, S, bc_colsum_ld(S), d_out, (cudaStream_t)stream));
  BC_LAUNCHED(1);
  return BC_OK;
}

int bc_core_resid(bc_ctx* c, const double* d_colsum, double scaling, const double* d_Vc, int M, int S, int64_t ldv,
                  const double* d_w, double* d_resid, void* stream) {
  if (!c || !d_colsum || !d_resid || M < 0 || S <= 0 || (M > 0 && (!d_Vc || !d_w))) return BC_ERR_ARG;
  BC_CUDA(launch_core_resid(d_colsum, scaling, d_Vc, M, S, ldv, d_w, d_resid, (cudaStream_t)stream));
  BC_LAUNCHED(1);
  return BC_OK;
}

int bc_core_maxcorr(bc_ctx* c, const double* d_Vc, int M, int S, int64_t ldv, const double* d_resid, int skip, double* d_out,
                    void* stream) {
  if (!c || !d_Vc || !d_resid || !d_out || M <= 0 || S <= 0) return BC_ERR_ARG;
  BC_CUDA(launch_core_maxcorr(d_Vc, M, S, ldv, d_resid, skip, d_out, (cudaStream_t)stream));
  BC_LAUNCHED(1);
  return BC_OK;
}

int bc_core_grad(bc_ctx* c, const double* d_Vc, int M, int S, int64_t ldv, const double* d_resid, double* d_grad, void* stream) {
  if (!c || !d_Vc || !d_resid || !d_grad || M <= 0 || S <= 0) return BC_ERR_ARG;
  BC_CUDA(launch_core_grad(d_Vc, M, S, ldv, d_resid, d_grad, (cudaStream_t)stream));
  BC_LAUNCHED(1);
  return BC_OK;
}

int bc_adam_step(bc_ctx* c, const double* d_g, double* d_x, double* d_m1, double* d_m2, int n, double lr, double b1, double b2,
                 double c1, double c2, double eps, const unsigned char* d_nn_mask, void* stream) {
  if (!c || !d_g || !d_x || !d_m1 || !d_m2 || n < 0) return BC_ERR_ARG;
  BC_CUDA(launch_adam(d_g, d_x, d_m1, d_m2, n, lr, b1, b2, c1, c2, eps, d_nn_mask, (cudaStream_t)stream));
  BC_LAUNCHED(1);
  return BC_OK;
}

int bc_greedy_opt_step(bc_ctx* c, const bc_step_args* a, void* stream) {
  if (!c || !a || a->phase < 0 || a->phase > 2) return BC_ERR_ARG;
  if (!a->d_theta || a->M <= 0 || !a->d_pts || !a->d_Vc || !a->d_parts || !a->d_colsum || !a->d_resid || !a->d_grad || !a->d_w ||
      !a->d_m1 || !a->d_m2 || a->n < 0)
    return BC_ERR_ARG;
  if (a->phase == 2 && (!a->d_parts_all || a->nparts < 1)) return BC_ERR_ARG;
  int rc;
  bool combined = false;
  if (a->phase != 2) {
    if ((rc = bc_set_samples(c, a->d_theta, a->S, a->ldt, stream))) return rc;
    if (a->ev_pass_begin) BC_CUDA(cudaEventRecord((cudaEvent_t)a->ev_pass_begin, (cudaStream_t)stream));
    // a single-part job (phase 0) with rows to pass over: the pass's finalize kernel also writes the centred column sum
    c->fuse_colsum_out = (a->phase == 0 && a->n > 0) ? a->d_colsum : nullptr;
    combined = c->fuse_colsum_out != nullptr;
    if (a->d_image && a->n > 0) {
      const void* img = a->d_image;
      const double* rs = a->d_rowscale;
      const double* ra = a->d_rowaux_q;
      if (a->d_rows) {
        if (!a->d_gimage || !a->d_growscale) return BC_ERR_ARG;
        if ((rc = bc_q_gather_rows(c, a->d_image, a->d_rowscale, a->d_rowaux_q, a->d_rows, a->n, a->d_gimage, a->d_growscale,
                                   a->d_rowaux_q ? a->d_growaux : nullptr, stream)))
          return rc;
        img = a->d_gimage;
        rs = a->d_growscale;
        ra = a->d_rowaux_q ? a->d_growaux : nullptr;
      }
      rc = bc_project_colsum_q(c, img, rs, a->n, ra, a->d_parts, stream);
    } else {
      if (a->n > 0 && !a->d_X) return BC_ERR_ARG;
      rc = bc_project_colsum(c, a->d_X ? a->d_X : a->d_pts, a->d_X ? a->ldx : a->ldp, a->d_rows, a->n, a->d_rowaux, a->d_parts, stream);
    }
    c->fuse_colsum_out = nullptr;
    if (rc) return rc;
    if (a->ev_pass_end) BC_CUDA(cudaEventRecord((cudaEvent_t)a->ev_pass_end, (cudaStream_t)stream));
    if (a->phase == 1) return BC_OK;
  }
  if (a->phase == 2) {
    if ((rc = bc_colsum_combine(c, a->d_parts_all, a->nparts, a->S, a->d_colsum, stream))) return rc;
  } else if (!combined) {
    if ((rc = bc_colsum_combine(c, a->d_parts, 1, a->S, a->d_colsum, stream))) return rc;
  }
  if ((rc = bc_project_materialise(c, a->d_pts, a->ldp, nullptr, a->M, a->d_pts_rowaux, a->d_Vc, a->ldv, nullptr, nullptr, 0, stream))) return rc;
  // residual, gradient and the projected ADAM update in one launch (the kernels of bc_core_resid / bc_core_grad / bc_adam_step
  // back to back in one CTA)
  BC_CUDA(launch_core_step(a->d_colsum, a->scaling, a->d_Vc, a->M, a->S, a->ldv, a->d_w, a->d_resid, a->d_grad, a->d_m1, a->d_m2, a->lr,
                           a->b1, a->b2, a->c1, a->c2, a->eps, a->d_nn_mask, a->d_sched, a->d_step_counter, (cudaStream_t)stream));
  BC_LAUNCHED(1);
  return BC_OK;
}

// ---- support for replaying an optimiser step as a CUDA graph (coreset/_greedy.py) ----
int bc_sample_slot(const bc_ctx* c) { return c ? c->sslot : BC_ERR_ARG; }
int bc_set_sample_slot(bc_ctx* c, int slot) {
  if (!c || (slot != 0 && slot != 1)) return BC_ERR_ARG;
  c->sslot = slot;
  return BC_OK;
}
int64_t bc_add_launch_count(int64_t n) {
  g_launches.fetch_add((long long)n, std::memory_order_relaxed);
  return (int64_t)g_launches.load(std::memory_order_relaxed);
}

int bc_core_pgrad(bc_ctx* c, const double* d_P, int M, int64_t ldp, const double* d_w, const double* d_resid, double* d_out,
                  int64_t ldo, void* stream) {
  if (!c || !d_P || !d_w || !d_resid || !d_out || M < 0 || ldp < 1 || ldo < 1) return BC_ERR_ARG;
  if (!c->potential_set || !c->samples_set) return BC_ERR_STATE;
  if (c->model == BC_MODEL_NEURLIN) return BC_ERR_UNSUPPORTED;   // the reference defines no gradient for it (model_neurlinr.py:99-100)
  if (ldp < c->Dk || ldo < c->Dk) return BC_ERR_ARG;
  if ((size_t)(2 * c->Dk + 2 * c->S) * sizeof(double) > 48 * 1024) return BC_ERR_UNSUPPORTED;
  BC_CUDA(launch_pgrad_model(c->model, d_P, M, ldp, c->B, c->S, c->Dk, c->Dpad, c->d_siginv, d_w, d_resid, d_out, ldo,
                             (cudaStream_t)stream));
  BC_LAUNCHED(1);
  return BC_OK;
}

int bc_dense_pgrad(bc_ctx* c, const double* d_G, int M, int S, int D, const double* d_w, const double* d_resid, int centre,
                   double* d_out, int64_t ldo, void* stream) {
  if (!c || !d_G || !d_w || !d_resid || !d_out || M < 0 || S <= 0 || D <= 0 || ldo < D) return BC_ERR_ARG;
  if ((size_t)2 * S * sizeof(double) > 48 * 1024) return BC_ERR_UNSUPPORTED;
  BC_CUDA(launch_pgrad_dense(d_G, M, S, D, d_w, d_resid, centre, d_out, ldo, (cudaStream_t)stream));
  BC_LAUNCHED(1);
  return BC_OK;
}

static int laplace_common(bc_ctx* c, const double* d_Z, int64_t ldz, const double* d_w, int M, int D, double* d_mu, double* d_L, int maxit,
                          double tol, int* d_info, int flags, void* stream) {
  if (!c || !d_Z || !d_w || !d_mu || !d_L || !d_info || M < 1 || D < 1 || D > 160 || ldz < D || maxit < 1) return BC_ERR_ARG;
  if (((size_t)D * (D + 1) + 4 * (size_t)D + 5 * (size_t)M + 80) * sizeof(double) > kMaxSmem - 1024) return BC_ERR_UNSUPPORTED;
  BC_CUDA(launch_laplace_logistic(d_Z, ldz, d_w, M, D, d_mu, d_L, maxit, tol, d_info, flags, (cudaStream_t)stream));
  BC_LAUNCHED(1);
  return BC_OK;
}

int bc_laplace_logistic(bc_ctx* c, const double* d_Z, int64_t ldz, const double* d_w, int M, int D, double* d_mu, double* d_L, int maxit,
                        double tol, int* d_info, void* stream) {
  return laplace_common(c, d_Z, ldz, d_w, M, D, d_mu, d_L, maxit, tol, d_info, 0, stream);
}

int bc_laplace_logistic_factor(bc_ctx* c, const double* d_Z, int64_t ldz, const double* d_w, int M, int D, double* d_mu, double* d_C, int maxit,
                               double tol, int* d_info, void* stream) {
  return laplace_common(c, d_Z, ldz, d_w, M, D, d_mu, d_C, maxit, tol, d_info, 3, stream);
}

int bc_conjugate_factor(bc_ctx* c, int model, const double* d_Z, int64_t ldz, const double* d_w, int M, int D, const double* d_A0,
                        const double* d_A1, const double* d_v0, double sigsq, double* d_mu, double* d_C, int* d_info, void* stream) {
  if (!c || !d_Z || !d_w || !d_A0 || !d_v0 || !d_mu || !d_C || !d_info || M < 1 || D < 1 || D > 160) return BC_ERR_ARG;
  if (model == BC_MODEL_GAUSSIAN) {
    if (!d_A1 || ldz < D) return BC_ERR_ARG;
  } else if (model == BC_MODEL_NEURLIN) {
    if (!(sigsq > 0.0) || ldz < D + 1) return BC_ERR_ARG;
  } else {
    return BC_ERR_UNSUPPORTED;
  }
  BC_CUDA(launch_conjugate_factor(model, d_Z, ldz, d_w, M, D, d_A0, d_A1, d_v0, sigsq, d_mu, d_C, d_info, (cudaStream_t)stream));
  BC_LAUNCHED(1);
  return BC_OK;
}

int bc_sample_affine(bc_ctx* c, const double* d_mu, const double* d_L, const double* d_R, int S, int D, double* d_theta, int ldt,
                     void* stream) {
  if (!c || !d_mu || !d_L || !d_R || !d_theta || S < 0 || D < 1 || ldt < D) return BC_ERR_ARG;
  BC_CUDA(launch_sample_affine(d_mu, d_L, d_R, S, D, d_theta, ldt, (cudaStream_t)stream));
  BC_LAUNCHED(1);
  return BC_OK;
}

int bc_sample_solve_hinted(bc_ctx* c, const double* d_mu, const double* d_C, const double* d_R, int S, int D, double* d_theta, int ldt,
                           const int* d_factor_info, void* stream) {
  if (!c || !d_mu || !d_C || !d_R || !d_theta || S < 0 || D < 1 || D > 160 || ldt < D) return BC_ERR_ARG;
  BC_CUDA(launch_sample_solve(d_mu, d_C, d_R, S, D, d_theta, ldt, d_factor_info ? d_factor_info + 1 : nullptr, (cudaStream_t)stream));
  BC_LAUNCHED(1);
  return BC_OK;
}

int bc_sample_solve(bc_ctx* c, const double* d_mu, const double* d_C, const double* d_R, int S, int D, double* d_theta, int ldt,
                    void* stream) {
  return bc_sample_solve_hinted(c, d_mu, d_C, d_R, S, D, d_theta, ldt, nullptr, stream);
}

int bc_dense_rownorms(bc_ctx* c, const double* d_V, int64_t n, int S, int64_t ldv, double* d_norms, void* stream) {
  if (!c || !d_V || !d_norms || n < 0 || S <= 0) return BC_ERR_ARG;
  BC_CUDA(launch_dense_rowstats(d_V, n, S, ldv, nullptr, 0, d_norms, nullptr, (cudaStream_t)stream));
  BC_LAUNCHED(1);
  return BC_OK;
}

static int dense_ws(bc_ctx* c, int S) {
  const size_t need = (size_t)c->sms * 8 * (size_t)(S > 4 ? S : 4);
  return grow(&c->dense_part, &c->cap_dense, need);
}

int bc_dense_center(bc_ctx* c, double* d_V, int64_t n, int S, int64_t ldv, void* stream) {
  if (!c || !d_V || n < 0 || S <= 0) return BC_ERR_ARG;
  BC_CUDA(launch_dense_center(d_V, n, S, ldv, (cudaStream_t)stream));
  BC_LAUNCHED(1);
  return BC_OK;
}

int bc_dense_colsum(bc_ctx* c, const double* d_V, int64_t n, int S, int64_t ldv, double* d_out_dd, void* stream) {
  if (!c || !d_V || !d_out_dd || n < 0 || S <= 0) return BC_ERR_ARG;
  int rc;
  if ((rc = dense_ws(c, S))) return rc;
  BC_CUDA(launch_dense_colsum(d_V, n, S, ldv, c->dense_part, c->sms * 8, d_out_dd, bc_colsum_ld(S), (cudaStream_t)stream));
  BC_LAUNCHED(2);
  return BC_OK;
}

int bc_dense_score(bc_ctx* c, int mode, const double* d_V, int64_t n, int S, int64_t ldv, const double* d_norms,
                   const double* d_u, const unsigned char* d_active, int64_t idx_offset, double* d_out, double* d_scores,
                   void* stream) {
  if (!c || !d_V || !d_u || !d_out || n <= 0 || S <= 0 || mode < 0 || mode > 3) return BC_ERR_ARG;
  if (mode != BC_SCORE_CORR && !d_norms) return BC_ERR_ARG;
  if (mode == BC_SCORE_OMP && !d_active) return BC_ERR_ARG;
  if ((size_t)S * 2 * sizeof(double) > 200 * 1024) return BC_ERR_UNSUPPORTED;
  int rc;
  if ((rc = dense_ws(c, S))) return rc;
  BC_CUDA(launch_dense_score(d_V, n, S, ldv, d_norms, d_u, mode, d_active, idx_offset, c->dense_part, c->sms * 8, d_out, d_scores,
                             (cudaStream_t)stream));
  BC_LAUNCHED(2);
  return BC_OK;
}

int bc_dense_combine(bc_ctx* c, const double* d_V, int64_t ldv, int S, const int64_t* d_idx, const double* d_w, int m,
                     double* d_out, void* stream) {
  if (!c || !d_V || !d_out || S <= 0 || m < 0 || (m > 0 && (!d_idx || !d_w))) return BC_ERR_ARG;
  BC_CUDA(launch_dense_combine(d_V, ldv, S, reinterpret_cast<const long long*>(d_idx), d_w, m, d_out, (cudaStream_t)stream));
  BC_LAUNCHED(1);
  return BC_OK;
}

int bc_dense_gather(bc_ctx* c, const double* d_V, int64_t ldv, int S, const int64_t* d_idx, int64_t m, double* d_out, int64_t ldo,
                    void* stream) {
  if (!c || !d_V || !d_out || S <= 0 || m < 0 || (m > 0 && !d_idx)) return BC_ERR_ARG;
  BC_CUDA(launch_dense_gather(d_V, ldv, S, reinterpret_cast<const long long*>(d_idx), m, d_out, ldo, (cudaStream_t)stream));
  BC_LAUNCHED(1);
  return BC_OK;
}

int bc_transpose(bc_ctx* c, const double* d_A, int64_t rows, int64_t cols, int64_t lda, double* d_out, int64_t ldo, void* stream) {
  if (!c || !d_A || !d_out || rows < 0 || cols < 0) return BC_ERR_ARG;
  BC_CUDA(launch_transpose(d_A, rows, cols, lda, d_out, ldo, (cudaStream_t)stream));
  BC_LAUNCHED(1);
  return BC_OK;
}

int bc_vec_step(bc_ctx* c, int op, const double* d_xw, const double* d_xf, const double* d_b, int S, double aux, double* d_u,
                double* d_out, void* stream) {
  if (!c || !d_xw || !d_b || !d_out || S <= 0 || op < 0 || op > 3) return BC_ERR_ARG;
  if ((op == BC_VEC_GIGA_STEP || op == BC_VEC_FW_STEP) && !d_xf) return BC_ERR_ARG;
  if (op == BC_VEC_GIGA_DIR && !d_u) return BC_ERR_ARG;
  BC_CUDA(launch_vec_step(op, d_xw, d_xf, d_b, S, aux, d_u, d_out, (cudaStream_t)stream));
  BC_LAUNCHED(1);
  return BC_OK;
}

int bc_nnls_max_columns(void) { return 112; }

int bc_nnls(bc_ctx* c, const double* d_rows, int S, const int64_t* d_pos, int m, const double* d_b, const double* d_x0, double* d_x, int maxit,
            int* d_info, void* stream) {
  if (!c || !d_rows || !d_pos || !d_b || !d_x || !d_info || S <= 0 || m < 1) return BC_ERR_ARG;
  if (m > bc_nnls_max_columns()) return BC_ERR_UNSUPPORTED;
  int rc;
  if ((rc = grow(&c->nnls_rho, &c->cap_rho, (size_t)S))) return rc;
  BC_CUDA(launch_nnls_gram(d_rows, S, reinterpret_cast<const long long*>(d_pos), m, d_b, d_x0, d_x, c->nnls_rho, maxit > 0 ? maxit : 3 * m + 10,
                           d_info, (cudaStream_t)stream));
  BC_LAUNCHED(1);
  return BC_OK;
}

int bc_solver_iterations(bc_ctx* c, int algo, int iterations, const double* d_V, int64_t n, int S, int64_t ldv, const double* d_norms,
                         const double* d_b_search, const double* d_b, double aux, double tol, double* d_Vact, double* d_ctl, double* d_aw,
                         double* d_aw_prev, int64_t* d_act, double* d_xw, double* d_u, double* d_scratch8, void* stream) {
  if (!c || (algo != 0 && algo != 1) || iterations < 0 || !d_V || n <= 0 || S <= 0 || !d_norms || !d_b_search || !d_b || !d_Vact || !d_ctl ||
      !d_aw || !d_aw_prev || !d_act || !d_xw || !d_u || !d_scratch8)
    return BC_ERR_ARG;
  if ((size_t)S * 2 * sizeof(double) > 200 * 1024) return BC_ERR_UNSUPPORTED;
  int rc;
  if ((rc = dense_ws(c, S))) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  for (int i = 0; i < iterations; ++i) {
    // _select: GIGA direction / Frank-Wolfe residual, then the score pass with its arg-max (giga.py:20-38, frankwolfe.py:15-17)
    BC_CUDA(launch_vec_step(algo == 0 ? BC_VEC_GIGA_DIR : BC_VEC_RESID, d_xw, nullptr, algo == 0 ? d_b_search : d_b, S, 0.0, d_u, d_scratch8,
                            st, d_ctl));
    BC_CUDA(launch_dense_score(d_V, n, S, ldv, d_norms, d_u, algo == 0 ? BC_SCORE_GIGA : BC_SCORE_FW, nullptr, 0, c->dense_part, c->sms * 8,
                               d_scratch8 + 4, nullptr, st, d_ctl));
    // _reweight + error + monotone check
    BC_CUDA(launch_solver_step(algo, d_V, ldv, S, d_norms, d_b_search, d_b, aux, tol, d_Vact, d_ctl, d_aw, d_aw_prev,
                               reinterpret_cast<long long*>(d_act), d_xw, d_scratch8, st));
    BC_LAUNCHED(4);
  }
  return BC_OK;
}

int bc_host_project(int device, int model, int kind, int D, const double* h_params, const double* h_siginv, const double* h_X,
                    int64_t n, int64_t ldx_h, const double* h_theta, int S, double* h_V, int centred) {
  if (!h_X || !h_theta || !h_V || n < 0 || S <= 0 || D <= 0) return BC_ERR_ARG;
  bc_ctx* c = nullptr;
  int rc = bc_create(device, &c);
  if (rc) return rc;
  const int extra = (model == BC_MODEL_NEURLIN) ? 1 : 0;
  if (ldx_h < D + extra) {
    bc_destroy(c);
    return BC_ERR_ARG;
  }
  const int64_t ldx = ((D + extra + 3) / 4) * 4;  // padded device layout
  double *dX = nullptr, *dT = nullptr, *dV = nullptr, *dSig = nullptr, *dRa = nullptr;
  cudaError_t e = cudaSuccess;
  auto fail = [&](int code) {
    cudaFree(dX); cudaFree(dT); cudaFree(dV); cudaFree(dSig); cudaFree(dRa);
    bc_destroy(c);
    return code;
  };
  const size_t nn = (size_t)(n > 0 ? n : 1);
  if ((e = cudaMalloc((void**)&dX, nn * ldx * 8)) != cudaSuccess) return fail(cuda_fail(e));
  if ((e = cudaMalloc((void**)&dT, (size_t)S * D * 8)) != cudaSuccess) return fail(cuda_fail(e));
  if ((e = cudaMalloc((void**)&dV, nn * S * 8)) != cudaSuccess) return fail(cuda_fail(e));
  if ((e = cudaMemset(dX, 0, nn * ldx * 8)) != cudaSuccess) return fail(cuda_fail(e));
  if (n > 0 && (e = cudaMemcpy2D(dX, ldx * 8, h_X, ldx_h * 8, (size_t)(D + extra) * 8, n, cudaMemcpyHostToDevice)) != cudaSuccess)
    return fail(cuda_fail(e));
  if ((e = cudaMemcpy(dT, h_theta, (size_t)S * D * 8, cudaMemcpyHostToDevice)) != cudaSuccess) return fail(cuda_fail(e));
  if (model == BC_MODEL_GAUSSIAN) {
    if (!h_siginv) return fail(BC_ERR_ARG);
    if ((e = cudaMalloc((void**)&dSig, (size_t)D * D * 8)) != cudaSuccess) return fail(cuda_fail(e));
    if ((e = cudaMemcpy(dSig, h_siginv, (size_t)D * D * 8, cudaMemcpyHostToDevice)) != cudaSuccess) return fail(cuda_fail(e));
    if ((e = cudaMalloc((void**)&dRa, nn * 8)) != cudaSuccess) return fail(cuda_fail(e));
  }
  if ((rc = bc_set_potential(c, model, kind, D, h_params, dSig))) return fail(rc);
  if ((rc = bc_set_samples(c, dT, S, D, nullptr))) return fail(rc);
  if (model == BC_MODEL_GAUSSIAN && (rc = bc_rowquad(c, dX, n, ldx, dRa, nullptr))) return fail(rc);
  if ((rc = bc_project_materialise(c, dX, ldx, nullptr, n, dRa, dV, S, nullptr, nullptr, centred ? 0 : 1, nullptr))) return fail(rc);
  if ((e = cudaDeviceSynchronize()) != cudaSuccess) return fail(cuda_fail(e));
  if (n > 0 && (e = cudaMemcpy(h_V, dV, (size_t)n * S * 8, cudaMemcpyDeviceToHost)) != cudaSuccess) return fail(cuda_fail(e));
  return fail(BC_OK);
}

}  // extern "C"
