// beta-cores B200: internal launch interface between the kernel translation units and bc_api.cu.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>
#include "bc_models.cuh"

#include <atomic>

namespace bc {

// One flag per CUDA device ordinal: function attributes (the opt-in dynamic shared-memory limit) and SM counts belong to a
// DEVICE, and one process may drive several (Engine.get(device)); a per-process `static bool` would leave the second
// device's kernels without their attribute.  Thread-safe; ordinals >= 64 simply repeat the (idempotent) call.
struct DeviceOnce {
  std::atomic<unsigned long long> mask{0};
  bool done(int dev) const { return dev >= 0 && dev < 64 && ((mask.load(std::memory_order_acquire) >> dev) & 1ull); }
  void set(int dev) {
    if (dev >= 0 && dev < 64) mask.fetch_or(1ull << dev, std::memory_order_release);
  }
};
inline int current_device() {
  int d = 0;
  cudaGetDevice(&d);
  return d;
}
template <class K>
inline cudaError_t raise_dynamic_smem(K kern, size_t bytes, DeviceOnce& once) {
  const int dev = current_device();
  if (once.done(dev)) return cudaSuccess;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e == cudaSuccess) once.set(dev);
  return e;
}

// The small kernels of an optimiser step run between launches of the projection kernels, which take nearly all of an SM's
// shared memory.  A kernel that leaves the shared-memory / L1 split at its default makes the SM re-partition on the way in
// and again on the way out; asking for the maximum carve-out everywhere keeps one configuration for the whole step.
template <class K>
inline void prefer_max_shared(K kern, DeviceOnce& once) {
  const int dev = current_device();
  if (once.done(dev)) return;
  cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared);
  once.set(dev);
}
#define BC_PREFER_MAX_SHARED(kern)          \
  do {                                      \
    static DeviceOnce carve_once__;         \
    prefer_max_shared(kern, carve_once__);  \
  } while (0)

constexpr int kComputeWarps = 8;
constexpr int kComputeThreads = kComputeWarps * 32;
constexpr int kThreads = kComputeThreads + 32;  // + one TMA producer warp
constexpr size_t kMaxSmem = 227 * 1024;

enum : int { MODE_COLSUM = 0, MODE_SCORE = 1, MODE_MATERIALISE = 2 };

struct ProjArgs {
  // data rows (A operand): row r of the processed list is A + (rows ? rows[r] : r) * lda
  const double* A;
  long long lda;
  const long long* rows;  // optional gather list (device), length n
  long long n;            // rows to process
  long long idx_offset;   // added to the reported arg-max position (row-sharded multi-GPU)
  // prepared samples (B operand): S x ldb, zero beyond Dk, ldb == Dpad
  const double* B;
  int ldb;
  int S, Dk, Dc, Dpad, ss;
  const double* colaux;  // [S] or null
  const double* rowaux;  // indexed by absolute row id, or null (neural-linear reads column Dk of the row)
  const double* bbar;    // [Dpad+1] mean prepared sample (pivot); element Dpad = mean of colaux
  ModelParams mp;
  // per-CTA partials
  double* part_colsum;  // [grid][2][Sld]
  double* part_misc;    // [grid][4]
  int Sld;
  // MODE_SCORE
  const double* resid;  // [S+1]: residual, then its sum
  double* scores;       // optional [n]
  // MODE_MATERIALISE
  double* V;
  long long ldv;
  double* norms;  // optional [n]
  int raw;        // 1: write the un-centred potential f itself (no pivot, no mean subtraction)
  int want_colsum;  // MODE_MATERIALISE: also accumulate the column sums
};

int project_tile_config(int Dpad, int* BM, int* BN, int* ss, size_t* smem);
int project_tile_config_for_rows(int base_cfg, long long n, int sms, int Dpad, size_t* smem_out);
int project_tile_rows(int cfg);
cudaError_t launch_project(const ProjArgs& P, int model, int kind, int poly, int mode, int tile_cfg, int grid, size_t smem, cudaStream_t st, int csplit = 1);
int project_chunks(int S);
cudaError_t launch_project_finalize(const double* part_colsum, const double* part_misc, int nctas, int S, int Sld, double* out_dd,
                                    double* out_best, int mode, cudaStream_t st, double* colsum_out = nullptr);


// ---- bc_project_q.cu: tensor-core (tcgen05 int8 Ozaki) route of the fused projection ----
enum : int { QMODE_COLSUM = MODE_COLSUM, QMODE_SCORE = MODE_SCORE, QMODE_DOT = 3 };
struct QProjArgs {
  const unsigned char* imgA;  // [tiles][7][128][128] swizzled
  const double* rowscale;     // [n]
  const unsigned char* imgB;  // [chunks][7][32][128] swizzled
  const double* colscale;     // [1]: the samples share one power-of-two scale
  long long n;
  long long idx_offset;
  int S;
  int ksteps;            // 32-byte K blocks that hold features: ceil(D / 32), 1..4
  const double* colaux;  // [S] or null
  const double* rowaux;  // [n] or null
  ModelParams mp;
  const double* pot_tabs;  // the potential's lane tables (LogisticTabs: 64 doubles) or null
  double* part_colsum;  // [grid][2][Sld]
  double* part_misc;    // [grid][4]
  int Sld;
  const double* resid;  // [S+1]
  double* scores;       // optional [n]
  double* V;            // QMODE_DOT: the contraction itself, n x ldv
  long long ldv;
};
cudaError_t launch_feature_exponents(const double* X, long long ldx, long long n, int D, unsigned long long* scratch, int* out,
                                     cudaStream_t st);
cudaError_t launch_quantise_rows(const double* X, long long ldx, long long n, int D, unsigned char* image, double* rowscale,
                                 double* aux_out, int aux_col, const int* fexp, cudaStream_t st);
cudaError_t launch_gather_image(const unsigned char* src, const double* src_scale, const double* src_aux, const long long* idx, long long n,
                                unsigned char* dst, double* dst_scale, double* dst_aux, cudaStream_t st);
cudaError_t launch_quantise_samples(const double* B, int ldb, int S, int D, unsigned char* image, double* colscale, int* common_e,
                                    const int* fexp, unsigned long long* scratch4, int slot, bool have_max, cudaStream_t st);
cudaError_t launch_project_q(const QProjArgs& P, int model, int kind, int poly, int mode, int digits, int grid, cudaStream_t st);

// ---- bc_small.cu: sample preparation, coreset-side step, ADAM ----
cudaError_t launch_prepare_samples(int model, const double* theta, int S, int D, int ldt, const double* siginv, const double* siginvT,
                                   double* B, int ldb, double* colaux, double* bbar, const int* fexp, unsigned long long* absmax_slot,
                                   const int* siginv_diag, cudaStream_t st);
cudaError_t launch_offdiag_test(const double* A, int D, int* flag, cudaStream_t st);
cudaError_t launch_rowquad(const double* X, long long n, int D, long long ldx, const double* siginv, double* out, cudaStream_t st);
cudaError_t launch_colsum_combine(const double* parts, int nparts, int S, int Sld, double* out, cudaStream_t st);
cudaError_t launch_core_resid(const double* colsum, double scaling, const double* Vc, int M, int S, long long ldv, const double* w,
                              double* resid, cudaStream_t st);
cudaError_t launch_core_step(const double* colsum, double scaling, const double* Vc, int M, int S, long long ldv, double* x, double* resid,
                             double* grad, double* m1, double* m2, double lr, double b1, double b2, double c1, double c2, double eps,
                             const unsigned char* nn_mask, const double* sched, int* step_counter, cudaStream_t st);
cudaError_t launch_core_maxcorr(const double* Vc, int M, int S, long long ldv, const double* resid, int skip, double* out,
                                cudaStream_t st);
cudaError_t launch_core_grad(const double* Vc, int M, int S, long long ldv, const double* resid, double* grad, cudaStream_t st);
cudaError_t launch_adam(const double* g, double* x, double* m1, double* m2, int n, double lr, double b1, double b2, double c1,
                        double c2, double eps, const unsigned char* nn_mask, cudaStream_t st);

cudaError_t launch_pgrad_model(int model, const double* P, int M, long long ldp, const double* B, int S, int D, int ldb,
                               const double* siginv, const double* w, const double* resid, double* out, long long ldo,
                               cudaStream_t st);
cudaError_t launch_pgrad_dense(const double* G, int M, int S, int D, const double* w, const double* resid, int centre, double* out,
                               long long ldo, cudaStream_t st);

// ---- bc_sampler.cu: device-side posterior samplers ----
cudaError_t launch_laplace_logistic(const double* Z, long long ldz, const double* w, int M, int D, double* mu_io, double* Lsig, int maxit,
                                    double tol, int* info, int flags, cudaStream_t st);
cudaError_t launch_conjugate_factor(int model, const double* Z, long long ldz, const double* w, int M, int D, const double* A0, const double* A1,
                                    const double* v0, double sigsq, double* mu, double* C, int* info, cudaStream_t st);
cudaError_t launch_sample_solve(const double* mu, const double* C, const double* R, int S, int D, double* out, int ldo, const int* diag_hint,
                                cudaStream_t st);
cudaError_t launch_nnls_gram(const double* Vact, int S, const long long* pos, int m, const double* b, const double* x0, double* x_out,
                             double* rho, int maxit, int* info, cudaStream_t st);
cudaError_t launch_sample_affine(const double* mu, const double* L, const double* R, int S, int D, double* out, int ldo, cudaStream_t st);

// ---- bc_dense.cu: materialised (n x S) matrix kernels for the snnls solvers ----
cudaError_t launch_dense_rowstats(const double* V, long long n, int S, long long ldv, const double* u, int nu, double* norms,
                                  double* dots, cudaStream_t st);
cudaError_t launch_dense_colsum(const double* V, long long n, int S, long long ldv, double* part, int nparts, double* out_dd,
                                int Sld, cudaStream_t st);
cudaError_t launch_dense_score(const double* V, long long n, int S, long long ldv, const double* norms, const double* u, int mode,
                               const unsigned char* active, long long idx_offset, double* part, int nparts, double* out,
                               double* scores, cudaStream_t st, const double* stop = nullptr);
cudaError_t launch_dense_combine(const double* V, long long ldv, int S, const long long* idx, const double* w, int m, double* out,
                                 cudaStream_t st);
cudaError_t launch_dense_center(double* V, long long n, int S, long long ldv, cudaStream_t st);
cudaError_t launch_dense_gather(const double* V, long long ldv, int S, const long long* idx, long long m, double* out, long long ldo,
                                cudaStream_t st);
cudaError_t launch_transpose(const double* A, long long rows, long long cols, long long lda, double* out, long long ldo,
                             cudaStream_t st);
cudaError_t launch_vec_step(int op, const double* xw, const double* xf, const double* b, int S, double aux, double* u, double* out,
                            cudaStream_t st, const double* stop = nullptr);
cudaError_t launch_solver_step(int algo, const double* V, long long ldv, int S, const double* norms, const double* b, const double* b_err,
                               double aux, double tol, double* Vact, double* ctl, double* aw, double* aw_prev, long long* act, double* xw,
                               const double* sel, cudaStream_t st);

}  // namespace bc
