/* betacores.h -- C ABI of libbetacores.so, the B200 (sm_100a) implementation of the
 * beta-Cores coreset-construction hot path.
 *
 * The reference (dionman/beta-cores) is pure Python/numpy and has no FFI of its own: its
 * boundary is the `bayesiancoresets` Python API.  This library sits UNDER a drop-in copy of that
 * API (beta-cores_b200/bayesiancoresets, loaded with ctypes); every entry point below names the
 * reference code it replaces (paths relative to the reference root).
 *
 * Conventions
 *   - every function returns 0 (BC_OK) or a negative bc error code; nothing throws;
 *   - `d_` pointers are DEVICE pointers owned by the caller, `h_` pointers are HOST pointers;
 *   - all matrices are fp64, row-major; `ld*` are leading dimensions in elements;
 *   - device work is enqueued on `stream` (a cudaStream_t passed as void*; NULL = default stream)
 *     and is asynchronous; results are valid after the stream is synchronised;
 *   - a bc_ctx owns a per-device workspace; calls on one ctx must be stream-ordered with each
 *     other (use one ctx per concurrent stream).  No other global state.
 *   - data-row operands: base pointer 16-byte aligned, ld even (rows are staged with 16-byte
 *     bulk copies through the TMA engine).
 */
#ifndef BETACORES_H_
#define BETACORES_H_
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct bc_ctx bc_ctx;

enum { BC_OK = 0, BC_ERR_ARG = -1, BC_ERR_ALIGN = -2, BC_ERR_UNSUPPORTED = -3, BC_ERR_STATE = -4, BC_ERR_CUDA = -5 };
/* models: examples/common/{model_lr,gaussian,model_neurlinr}.py */
enum { BC_MODEL_LOGISTIC = 0, BC_MODEL_GAUSSIAN = 1, BC_MODEL_NEURLIN = 2 };
/* potentials: log-likelihood, beta-likelihood, d(beta-likelihood)/d(beta) (Gaussian only) */
enum { BC_KIND_LOGLIK = 0, BC_KIND_BETALIK = 1, BC_KIND_BETAGRAD = 2 };
/* bc_dense_score modes */
enum { BC_SCORE_FW = 0, BC_SCORE_GIGA = 1, BC_SCORE_CORR = 2, BC_SCORE_OMP = 3 };
/* bc_vec_step ops */
enum { BC_VEC_GIGA_DIR = 0, BC_VEC_GIGA_STEP = 1, BC_VEC_RESID = 2, BC_VEC_FW_STEP = 3 };

int bc_version(void);
const char* bc_error_string(int code);
int bc_last_cuda_error(void);           /* cudaError_t of the most recent BC_ERR_CUDA */
int64_t bc_launch_count(void);          /* kernels this library has launched in this process (instrumentation) */
/* Host-only diagnostic: the coefficients (highest degree first, variable 2t-1) bc_set_potential fits for
 * (1+t)^-beta on t in [0,1] -- the pow() of examples/common/model_lr.py:85 -- and their truncation bound.
 * degree <= 24.  Used by the CPU tests to check the polynomial against a high-precision reference. */
int bc_fit_pow_poly(double beta, int degree, double* h_coef, double* h_err);
/* Host-only diagnostic: the lane-table form of the same pow() that the tensor-core kernels use when it is accurate enough
 * for the given beta (csrc/bc_models.cuh, LogisticF<KIND_BETALIK, kPowTab>):  h_w = the 7 coefficients of Q (highest degree
 * first) in (1+w)^-beta = 1 + w Q(w), |w| <= 1/65;  h_rs[j] = 1/s_j and h_us[j] = s_j^-beta for the 32 interval centres
 * s_j = 1 + (j + 1/2)/32 of s = 1 + t;  *h_err = max |w Q(w) - ((1+w)^-beta - 1)|. */
int bc_fit_pow_tab(double beta, double* h_w, double* h_rs, double* h_us, double* h_err);
/* Which evaluation form of a potential the tensor-core kernels may use: 0 = the fastest one that meets the quarter-ulp
 * bound (default), 1 = the polynomial forms only (what the FP64-DMMA kernels always use).  Takes effect with the next
 * bc_set_potential.  bc_potential_form: 2 if the current potential runs in its lane-table form, else 1. */
int bc_set_potential_form(bc_ctx* ctx, int form);
int bc_potential_form(const bc_ctx* ctx);
int bc_create(int device, bc_ctx** ctx);
int bc_destroy(bc_ctx* ctx);
int bc_sm_count(const bc_ctx* ctx);
int bc_colsum_ld(int S);                /* leading dimension of the (hi, lo) planes below: S + 1 */

/* ---- stage 1: potentials and posterior samples ------------------------------------------- */
/* Select the potential f(x_n, theta_s).  D = contraction length (feature count; the neural-linear
 * target y is column D of each data row).  h_params: 8 host doubles, see bc_models.cuh / the
 * Python model modules (they hold beta, sigma^2, normalisers ...).  d_siginv: device D x D (Gaussian
 * only); it must stay allocated and UNCHANGED until the next bc_set_potential (later calls read it and a transposed copy).  Replaces the likelihood callbacks of
 * examples/common/model_lr.py:72-86, gaussian.py:7-15,34-62, model_neurlinr.py:90-110. */
int bc_set_potential(bc_ctx* ctx, int model, int kind, int D, const double* h_params, const double* d_siginv);
/* Install the S current posterior samples (device, S x ldt): what Projector.update() stores in
 * self.samples (bayesiancoresets/coreset/projector.py:36-37, :65-66). */
int bc_set_samples(bc_ctx* ctx, const double* d_theta, int S, int ldt, void* stream);
/* Gaussian per-row term x Siginv x (gaussian.py:10), computed once per dataset. */
int bc_rowquad(bc_ctx* ctx, const double* d_X, int64_t n, int64_t ldx, double* d_out, void* stream);

/* Fused projection passes over n data rows (row r = d_X + (d_rows ? d_rows[r] : r) * ldx).
 * d_rowaux: Gaussian x Siginv x indexed by absolute row id, else NULL.
 *
 * colsum : column sum over rows of the row-centred projection, as double-double planes
 *          d_out_dd[0..S) hi, d_out_dd[ld..ld+S) lo, element S = sum of row means (subtract it:
 *          bc_colsum_combine).  Replaces `vecs.sum(axis=0)` of project_f()/project() output
 *          (projector.py:51-55 + bcores.py:77,145 / sparsevi.py:76,132) without materialising it.
 * score  : corrs = vecs.dot(resid)/sqrt((vecs**2).sum(1))/S and its np.argmax (bcores.py:78,81):
 *          d_best[0] = best score, d_best[1] = int64 position (bit pattern) + idx_offset;
 *          NaN ordering as numpy.  d_resid has S+1 entries (residual, then its sum).
 * materialise: writes the centred rows (what project_f()/project() return) and optionally their
 *          2-norms (hilbert.py:16) and the column sum planes; raw != 0 writes the un-centred
 *          potential itself (what the likelihood callbacks return) and no norms. */
int bc_project_colsum(bc_ctx* ctx, const double* d_X, int64_t ldx, const int64_t* d_rows, int64_t n, const double* d_rowaux,
                      double* d_out_dd, void* stream);
int bc_project_score(bc_ctx* ctx, const double* d_X, int64_t ldx, const int64_t* d_rows, int64_t n, const double* d_rowaux,
                     const double* d_resid, int64_t idx_offset, double* d_best, double* d_scores, void* stream);
int bc_project_materialise(bc_ctx* ctx, const double* d_X, int64_t ldx, const int64_t* d_rows, int64_t n, const double* d_rowaux,
                           double* d_V, int64_t ldv, double* d_norms, double* d_out_dd, int raw, void* stream);
/* ---- stage 1 (+2), tensor-core route --------------------------------------------------------
 * Same results as bc_project_colsum / bc_project_score (same reference lines), with the contraction on the
 * 5th-generation tensor cores: rows and samples are split error-free into 7 signed base-256 digits and multiplied
 * exactly in int8 x int8 -> int32 (tcgen05.mma.kind::i8, accumulators in TMEM); see csrc/bc_project_q.cu.
 * Feature count D <= bc_q_max_features() (128).  Gather lists: through bc_q_gather_rows.
 * The row image is built ONCE per dataset; bc_set_samples() builds the sample image. */
int bc_q_max_features(void);
/* Precision tier of the tensor-core contraction.  The images always hold 7 digits (55 bits + sign below the row / sample
 * maximum: the product of two fp64 operands to the accuracy of a dgemm, the reference's numpy `dot`,
 * examples/common/model_lr.py:83); a launch contracts the leading `digits` of them: 7 (default) = 28 digit pairs,
 * 6 = 21 pairs (operands as if rounded to 46 bits, |error| <= 2^-44 max|x_n| max|theta| sqrt(D) typical), 5 = 15 pairs
 * (38 bits).  BASELINE.json's north star allows "FP64 DMMA or 3xTF32" accuracy (3xTF32 keeps about 30 bits); the tier that
 * keeps the selected indices exact on every parity case is reported in DESIGN.md. */
int bc_set_contraction_digits(bc_ctx* ctx, int digits);   /* 5, 6 or 7 */
int bc_contraction_digits(const bc_ctx* ctx);
int bc_q_image_bytes(int64_t n, int64_t* bytes);     /* device bytes of the quantised image of n rows */
/* Sub-sampled / group passes on the tensor-core route (bcores.py:52-55 `np.random.randint(N, size=n)` rows, :46-51 group rows):
 * rows d_idx[0..n) of a quantised image (local row numbers, duplicates allowed) are re-packed, with their scales and optional
 * per-row auxiliary values, into a compact image of bc_q_image_bytes(n) bytes that bc_project_colsum_q / _score_q take like
 * any other.  d_rowaux / d_rowaux_out: both NULL or both given. */
int bc_q_gather_rows(bc_ctx* ctx, const void* d_image, const double* d_rowscale, const double* d_rowaux, const int64_t* d_idx, int64_t n,
                     void* d_image_out, double* d_rowscale_out, double* d_rowaux_out, void* stream);
/* Feature exponents (optional, recommended for data whose columns differ much in magnitude).  The digit split keeps 56
 * bits below the largest entry of a row; with x_k 2^-c_k in the row image and theta_k 2^+c_k in the sample image
 * (c_k = ilogb max_n |x_nk|, powers of two: the products are unchanged) every feature is O(1) in its row and the
 * truncation is relative to the largest PRODUCT, as in an fp64 dot product, not to the largest entries.
 *   bc_feature_exponents: d_fexp_out[k] = c_k of the n rows given, or -2^30 for a column with no finite non-zero entry
 *                         (row shards: take the element-wise maximum over the shards, then replace -2^30 by 0);
 *   bc_set_feature_exponents: the exponents of the row image the next _q passes will use (NULL: none).  They stay in
 *                         force across bc_set_samples / bc_set_potential; a sample image already built is rebuilt. */
int bc_feature_exponents(bc_ctx* ctx, const double* d_X, int64_t ldx, int64_t n, int D, int32_t* d_fexp_out, void* stream);
int bc_set_feature_exponents(bc_ctx* ctx, const int32_t* d_fexp, int D, void* stream);
/* d_image (16-byte aligned, bc_q_image_bytes(n)), d_rowscale[n]; optional d_aux_out[n] = column aux_col of every row
 * (the neural-linear target y, model_neurlinr.py:104); d_fexp: D feature exponents applied to the rows (NULL: none) --
 * the passes over this image then need bc_set_feature_exponents(ctx, d_fexp, D). */
int bc_quantise_rows(bc_ctx* ctx, const double* d_X, int64_t ldx, int64_t n, int D, int aux_col, void* d_image, double* d_rowscale,
                     double* d_aux_out, const int32_t* d_fexp, void* stream);
/* d_rowaux[n]: Gaussian x Siginv x (bc_rowquad) / neural-linear y (d_aux_out above); NULL for the logistic model. */
int bc_project_colsum_q(bc_ctx* ctx, const void* d_image, const double* d_rowscale, int64_t n, const double* d_rowaux,
                        double* d_out_dd, void* stream);
int bc_project_score_q(bc_ctx* ctx, const void* d_image, const double* d_rowscale, int64_t n, const double* d_rowaux,
                       const double* d_resid, int64_t idx_offset, double* d_best, double* d_scores, void* stream);
/* diagnostic: d_V[n x ldv] = the contraction X B^T itself (B = the prepared samples), as the tensor-core route computes it */
int bc_contraction_q(bc_ctx* ctx, const void* d_image, const double* d_rowscale, int64_t n, double* d_V, int64_t ldv, void* stream);

/* out[s] = sum_parts (colsum_s) - sum_parts (sum of row means); parts = [nparts][2][ld] (one per rank). */
int bc_colsum_combine(bc_ctx* ctx, const double* d_parts, int nparts, int S, double* d_out, void* stream);

/* ---- stage 3: coreset-side step ------------------------------------------------------------ */
/* resid = scaling*colsum - w . Vc ; resid[S] = sum(resid)          (bcores.py:77, :145) */
int bc_core_resid(bc_ctx* ctx, const double* d_colsum, double scaling, const double* d_Vc, int M, int S, int64_t ldv,
                  const double* d_w, double* d_resid, void* stream);
/* out[0] = max_{m>=skip} |Vc_m.resid/|Vc_m||/S, np.max NaN semantics    (bcores.py:79-80) */
int bc_core_maxcorr(bc_ctx* ctx, const double* d_Vc, int M, int S, int64_t ldv, const double* d_resid, int skip, double* d_out,
                    void* stream);
/* grad[m] = -(Vc_m . resid)/S                                        (bcores.py:146) */
int bc_core_grad(bc_ctx* ctx, const double* d_Vc, int M, int S, int64_t ldv, const double* d_resid, double* d_grad, void* stream);
/* Pseudo-point location gradients of BatchPSVICoreset (bpsvi.py:49-55):
 *   out[m][d] = -(w[m]/S) sum_s resid[s] * pg[m][s][d],  pg = d(log-likelihood)/d(point) centred over its last axis (projector.py:31).
 * bc_core_pgrad: pg of the current potential's model (logistic: model_lr.py:107-114; Gaussian: gaussian.py:17-20) against the
 *   installed samples, without forming the (M, S, D) tensor.  d_P: M x ldp pseudo-points; d_resid has S+1 entries.
 * bc_dense_pgrad: pg given as a contiguous (M, S, D) device tensor (opaque user callbacks); centre != 0 centres it first. */
int bc_core_pgrad(bc_ctx* ctx, const double* d_P, int M, int64_t ldp, const double* d_w, const double* d_resid, double* d_out,
                  int64_t ldo, void* stream);
int bc_dense_pgrad(bc_ctx* ctx, const double* d_G, int M, int S, int D, const double* d_w, const double* d_resid, int centre,
                   double* d_out, int64_t ldo, void* stream);
/* one projected-ADAM update of x given g (util/opt.py:45-52; c1 = 1-b1**(i+1), c2 = 1-b2**(i+1));
 * d_nn_mask NULL = clamp every coordinate (nn_opt), else clamp where mask != 0 (partial_nn_opt). */
int bc_adam_step(bc_ctx* ctx, const double* d_g, double* d_x, double* d_m1, double* d_m2, int n, double lr, double b1, double b2,
                 double c1, double c2, double eps, const unsigned char* d_nn_mask, void* stream);

/* ---- one optimiser step in one call ------------------------------------------------------------------------------
 * bc_greedy_opt_step enqueues, on `stream`, everything one projected-ADAM step of BetaCoreset._optimize / SparseVICoreset._optimize
 * does after the sampler has produced Theta (bcores.py:142-146, util/opt.py:45-52):
 *     bc_set_samples -> [bc_q_gather_rows] -> bc_project_colsum(_q) -> bc_colsum_combine -> bc_project_materialise (the M coreset
 *     rows) -> bc_core_resid -> bc_core_grad -> bc_adam_step
 * -- the same arithmetic in the same order as the individual entry points, so the results are bit-identical; what it saves is
 * the host's per-call overhead (a dozen FFI crossings per step matter when a step is a few hundred microseconds: the sub-sampled
 * Gaussian example runs 1000 of them per coreset point) and launches: a single-part job's pass leaves the combined column sum
 * itself, and residual + gradient + ADAM run as one kernel (k_core_step).  A sharded job exchanges the column-sum parts between the pass and the
 * combine: it calls the two halves (`phase`).  All pointers are device pointers. */
typedef struct bc_step_args {
  const double* d_theta; int S; int ldt;                         /* posterior samples of this step (S x ldt) */
  /* data rows: the tensor-core route when d_image is set (with its scratch image for gathered passes), else the FP64 route */
  const void* d_image; const double* d_rowscale; const double* d_rowaux_q;
  void* d_gimage; double* d_growscale; double* d_growaux;
  const double* d_X; int64_t ldx; const double* d_rowaux;
  const int64_t* d_rows; int64_t n;                              /* gather list (NULL: the first n rows) and row count */
  double scaling;                                                /* N / n for a sub-sampled pass, 1 otherwise (bcores.py:55) */
  /* coreset side */
  const double* d_pts; int64_t ldp; int M; const double* d_pts_rowaux;
  double* d_Vc; int64_t ldv;                                     /* M x S workspace: the centred projection of the coreset points */
  double* d_parts; double* d_colsum; double* d_resid; double* d_grad;   /* 2 (S+1), S, S+1, M doubles */
  /* projected ADAM (util/opt.py:45-52) */
  double* d_w; double* d_m1; double* d_m2; double lr, b1, b2, c1, c2, eps; const unsigned char* d_nn_mask;
  /* optional instrumentation: two cudaEvent_t recorded on `stream` right before / after the data-row pass (NULL: none) */
  void* ev_pass_begin; void* ev_pass_end;
  /* row-sharded jobs run the step in two halves around their exchange of the column-sum parts:
   * phase 0 = the whole step (single rank); 1 = samples .. this rank's part in d_parts; 2 = combine the nparts parts at
   * d_parts_all (rank order) .. ADAM */
  int phase; int nparts; const double* d_parts_all;
  /* CUDA-graph replays (launch parameters frozen at capture): when d_sched is set, step i = *d_step_counter takes
   * (lr, c1, c2) = d_sched[3 i ..] instead of the values above, and the step leaves i + 1 in *d_step_counter */
  const double* d_sched; int* d_step_counter;
} bc_step_args;
int bc_greedy_opt_step(bc_ctx* ctx, const bc_step_args* args, void* stream);
/* Replaying captured steps leaves two pieces of HOST state behind the device's: which of the two sample-maximum scratch slots
 * the last bc_set_samples used (it alternates per call), and the launch counter.  A caller that captures bc_greedy_opt_step /
 * bc_set_samples into a CUDA graph reads the slot right after capturing (bc_sample_slot), restores the slot of the graph it
 * replayed last (bc_set_sample_slot) and accounts for the replayed kernels (bc_add_launch_count). */
int bc_sample_slot(const bc_ctx* ctx);
int bc_set_sample_slot(bc_ctx* ctx, int slot);
int64_t bc_add_launch_count(int64_t n);

/* ---- device-side posterior samplers (the reference's samplers are host callbacks) ----------------
 * bc_laplace_logistic: Laplace approximation of the weighted logistic posterior with N(0, I) prior -- what
 *   bayesiancoresets/util/opt.py:10-33 get_laplace / examples/zellner_logreg/main.py:86-111 compute on the host: d_mu holds
 *   the start on entry and the mode on return (damped Newton, |step|_inf <= tol (1 + |mu|_inf)), d_L (D x D) the inverse of
 *   the lower Cholesky factor of the negative Hessian there (get_laplace's LSig).  d_Z: M x ldz coreset rows, d_w their
 *   weights (rows with weight 0 do not contribute).  d_info[0] = 0 ok / 2 Hessian not positive definite, d_info[1] = steps.
 * bc_sample_affine: d_theta[s][:] = d_mu + d_R[s][:] . d_L^T for S x D standard normals d_R -- the sampler line
 *   `mu + np.random.randn(S, D).dot(LSig.T)` (main.py:144), with the normals drawn by the caller's own stream. */
int bc_laplace_logistic(bc_ctx* ctx, const double* d_Z, int64_t ldz, const double* d_w, int M, int D, double* d_mu, double* d_L,
                        int maxit, double tol, int* d_info, void* stream);
/* Host-free forms for the optimiser loop (weights and coreset rows already on the device, nothing returns to the host):
 * bc_laplace_logistic_factor: as bc_laplace_logistic, with the Newton steps taken from the M x M dual system while M <= D
 *   (Woodbury: the negative Hessian is a rank-M update of the identity) and d_C = the lower Cholesky FACTOR of the negative
 *   Hessian at the mode (get_laplace's LSigInv), to be used with bc_sample_solve.
 * bc_conjugate_factor: the conjugate weighted posteriors of examples/common/gaussian.py:28-32 (model = BC_MODEL_GAUSSIAN:
 *   H = A0 + (sum w) A1, v = v0 + A1 sum_i w_i x_i) and model_neurlinr.py:115-122 (BC_MODEL_NEURLIN: H = A0 + X^T diag(w) X / sigsq,
 *   v = v0 + X^T (w y) / sigsq, rows [x, y]); A0 = Sig0inv, A1 = Siginv, v0 = Sig0inv mu0 (D x D row-major / D).  d_C = chol(H),
 *   d_mu = C^-1 C^-T v -- the reference's `LSigp.dot(LSigp.T)` applied to v.  D <= 160.  d_info[0] = 0 ok / 2 not positive
 *   definite; d_info[1] = 1 if H (hence C) is diagonal. */
int bc_laplace_logistic_factor(bc_ctx* ctx, const double* d_Z, int64_t ldz, const double* d_w, int M, int D, double* d_mu, double* d_C,
                               int maxit, double tol, int* d_info, void* stream);
int bc_conjugate_factor(bc_ctx* ctx, int model, const double* d_Z, int64_t ldz, const double* d_w, int M, int D, const double* d_A0,
                        const double* d_A1, const double* d_v0, double sigsq, double* d_mu, double* d_C, int* d_info, void* stream);
int bc_sample_affine(bc_ctx* ctx, const double* d_mu, const double* d_L, const double* d_R, int S, int D, double* d_theta, int ldt,
                     void* stream);
/* bc_sample_solve: the same samples from the Cholesky factor itself, d_theta[s][:] = d_mu + C^-1 d_R[s][:] (d_C: D x D lower
 *   factor of the negative Hessian, row-major, get_laplace's LSigInv; util/opt.py:27-33 inverts it and multiplies) -- the
 *   host side then only factors.  D <= 160. */
int bc_sample_solve(bc_ctx* ctx, const double* d_mu, const double* d_C, const double* d_R, int S, int D, double* d_theta, int ldt,
                    void* stream);
/* ... with the d_info words of the bc_conjugate_factor call that produced d_C (device pointer, or NULL): d_info[1] == 1 marks a
 *   DIAGONAL factor (diagonal prior and noise precisions: examples/zellner_gaussian/main.py:37-38, Sig0 = I, Sig = 500 I), for
 *   which bc_conjugate_factor skips the D sequential pivots and this call the substitution -- the general code computes the
 *   same bits there, every off-diagonal operation adding a signed zero. */
int bc_sample_solve_hinted(bc_ctx* ctx, const double* d_mu, const double* d_C, const double* d_R, int S, int D, double* d_theta, int ldt,
                           const int* d_factor_info, void* stream);

/* ---- stage 2 on a materialised n x S matrix (snnls solvers, black-box projections) --------- */
int bc_dense_rownorms(bc_ctx* ctx, const double* d_V, int64_t n, int S, int64_t ldv, double* d_norms, void* stream);
/* V[r][:] -= mean(V[r][:])  (projector.py:26, :55) for matrices produced by host callbacks */
int bc_dense_center(bc_ctx* ctx, double* d_V, int64_t n, int S, int64_t ldv, void* stream);
int bc_dense_colsum(bc_ctx* ctx, const double* d_V, int64_t n, int S, int64_t ldv, double* d_out_dd, void* stream);
/* d_out[0..4) = {best score, best index bits, best of -score over active rows, its index bits}
 * (giga.py:20-38, frankwolfe.py:15-17, orthopursuit.py:17-35, bcores.py:78-81). */
int bc_dense_score(bc_ctx* ctx, int mode, const double* d_V, int64_t n, int S, int64_t ldv, const double* d_norms,
                   const double* d_u, const unsigned char* d_active, int64_t idx_offset, double* d_out, double* d_scores,
                   void* stream);
/* out[s] = sum_m w[m] V[idx[m]][s]   (A.dot(w) over the selected columns; giga.py:21) */
int bc_dense_combine(bc_ctx* ctx, const double* d_V, int64_t ldv, int S, const int64_t* d_idx, const double* d_w, int m,
                     double* d_out, void* stream);
int bc_dense_gather(bc_ctx* ctx, const double* d_V, int64_t ldv, int S, const int64_t* d_idx, int64_t m, double* d_out, int64_t ldo,
                    void* stream);
int bc_transpose(bc_ctx* ctx, const double* d_A, int64_t rows, int64_t cols, int64_t lda, double* d_out, int64_t ldo, void* stream);
/* S-length line-search / direction steps of the solvers, one CTA (giga.py:21-30,42-62;
 * frankwolfe.py:16,30-31; snnls.py:28-29).  See bc_dense.cu for the per-op operand/output layout. */
int bc_vec_step(bc_ctx* ctx, int op, const double* d_xw, const double* d_xf, const double* d_b, int S, double aux, double* d_u,
                double* d_out, void* stream);

/* Non-negative least squares  min |A x - b|, x >= 0  on the device, for OrthoPursuit._reweight / SparseNNLS.optimize
 * (orthopursuit.py:37-42, snnls.py:82-97: the reference calls scipy.optimize.nnls).  A is S x m, given by its COLUMNS: column j
 * = row d_pos[j] (S contiguous doubles) of the row-major matrix d_rows (the solvers' cache of selected datapoints).  d_x0
 * (optional, m doubles >= 0): warm start -- the iteration begins on its support; NULL starts from x = 0 like scipy.
 * Lawson-Hanson active set on the Gram matrix with corrected-semi-normal-equation refinement, one CTA (csrc/bc_sampler.cu).
 * d_info[0] = 0 converged / 1 iteration cap / 2 numerically dependent columns (use a host routine); d_info[1] = solves taken.
 * m <= bc_nnls_max_columns().  The reference pins no version of scipy's routine for this path: agreement is to rounding. */
int bc_nnls_max_columns(void);
int bc_nnls(bc_ctx* ctx, const double* d_rows, int S, const int64_t* d_pos, int m, const double* d_b, const double* d_x0, double* d_x,
            int maxit, int* d_info, void* stream);

/* Device-resident solver iterations (GIGA: algo 0, Frank-Wolfe: algo 1).  Queues `iterations` iterations of SparseNNLS.build
 * (snnls.py:44-62 with giga.py:20-64 / frankwolfe.py:15-40) on `stream` WITHOUT a host round trip between them: per iteration
 * the direction / residual step, the score pass over the n datapoints with its arg-max, and one single-CTA kernel that does
 * what the host does between two score passes -- the guards, activation of the selected datapoint in the row cache, the line
 * search, w = alpha w, w_f = max(0, w_f + beta), A w re-formed from the cached rows, the new error and the strict monotone
 * check.  A guard that trips (cdirnrm < tol; gA <= 0 or gB < 0; gammanum / gammadenom out of range; error not monotone) sets
 * d_ctl[1], leaves the state as before that iteration and turns the iterations queued behind it into no-ops; the caller takes
 * that iteration through the per-iteration entry points, which is where the reference's NumericalPrecisionError handling lives.
 *   d_ctl (8 doubles): [0] cached datapoints m, [1] status (0 ok, 1 select guard, 2 line-search guard, 3 error not monotone,
 *     4 cache full / no candidate), [2] iterations completed (in/out), [3] set to 1 once a completed iteration has passed the
 *     monotone check, [4] error of the current weights (in: valid if any weight is positive), [5] check_error_monotone, [6]
 *     cache capacity in rows;   d_Vact: capacity x S row cache;  d_aw / d_act / d_aw_prev: capacity weights, global indices, backup;
 *   d_b_search: b / |b| with aux = |b| (GIGA) or b with aux = sum of the datapoint norms (Frank-Wolfe);  d_b: the caller's b;
 *   d_xw: A w of the current weights (in/out);  d_u: 2 S scratch;  d_scratch8: 8 doubles. */
int bc_solver_iterations(bc_ctx* ctx, int algo, int iterations, const double* d_V, int64_t n, int S, int64_t ldv, const double* d_norms,
                         const double* d_b_search, const double* d_b, double aux, double tol, double* d_Vact, double* d_ctl, double* d_aw,
                         double* d_aw_prev, int64_t* d_act, double* d_xw, double* d_u, double* d_scratch8, void* stream);

/* ---- host-buffer convenience (what a foreign-language binding would call first) ------------- */
/* h_V[n x S] = centred potential matrix = BetaBlackBoxProjector.project_f(pts, beta) /
 * BlackBoxProjector.project(pts) (projector.py:51-55, :23-26) for HOST inputs; copies in, runs the
 * materialise pass on `device`, copies out, synchronises.  h_X is n x ldx_h, h_theta is S x D.
 * centred = 0 returns the un-centred potential (the likelihood callback's own return value). */
int bc_host_project(int device, int model, int kind, int D, const double* h_params, const double* h_siginv, const double* h_X,
                    int64_t n, int64_t ldx_h, const double* h_theta, int S, double* h_V, int centred);

/* ---- host side: numpy's legacy global random stream, natively (csrc/bc_hostrng.cpp) ----------------------------------
 * The reference's samplers end in `np.random.randn(S, D)` (examples/zellner_gaussian/main.py:87-92, zellner_logreg/main.py:
 * 139-144) and its sub-sampled modes draw `np.random.randint(N, size=n)` (bayesiancoresets/coreset/bcores.py:53), all from
 * numpy's one global RandomState.  These two functions continue that stream bit for bit (MT19937 words, polar Box-Muller
 * with the cached second value, masked-rejection bounded integers) from a state in numpy's own representation
 * (`np.random.get_state()` -> key, pos, has_gauss, cached_gaussian), several times faster than numpy's generator; `threads`
 * worker threads share the per-pair log / sqrt.  No device work.  randint: high - 1 < 2^32. */
typedef struct bc_mt_state { uint32_t key[624]; int32_t pos; int32_t has_gauss; double gauss; } bc_mt_state;
int bc_mt_randn(bc_mt_state* st, double* h_out, int64_t n, int threads);
int bc_mt_randint(bc_mt_state* st, int64_t high, int64_t* h_out, int64_t n);

#ifdef __cplusplus
}
#endif
#endif /* BETACORES_H_ */
