/* tbnn.h -- C ABI of the B200-native HMC hot path behind the TensorBNN Python API.
 *
 * The reference (alpha-davidson/TensorBNN) has no FFI: its boundary is the Python
 * class API (network.py, predictor.py, paramAdapter.py) and all arithmetic is
 * delegated to TensorFlow(-Probability).  Each entry point below replaces the
 * reference call site cited next to it; tensorbnn_b200/_lib.py binds them with
 * ctypes (see INTEGRATION.md for the stub a reference maintainer would add).
 *
 * Conventions: every function returns 0 on success, non-zero on error, with
 * tbnn_last_error() giving a thread-local message.  Device pointers are BORROWED
 * (they must outlive the call and any work queued on `stream`); the library owns
 * only the opaque handle and its workspace.  A handle is bound to one CUDA device
 * and is not thread-safe.  `stream` is a cudaStream_t passed as void*.
 * Element type of every real-valued buffer is the handle's dtype (float or double).
 *
 * Flat layouts (must equal network.states / network.hyperStates order,
 * network.py:173-191, :542-543):
 *   theta[C][P]: per dense layer W row-major [out][in], then b[out]; per
 *                prelu/squareprelu layer slopes[width].
 *   hyper[C][H]: 4 scalars per dense layer, 1 per prelu, 2 per squareprelu,
 *                then 1 for the Gaussian likelihood.
 */
#ifndef TBNN_H
#define TBNN_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct tbnn_handle tbnn_handle;

/* layer kinds: names of the reference layer classes (predictor.py:30-34) */
enum {
  TBNN_DENSE_CAUCHY = 0,   /* layer.py:101  CauchyDenseLayer == DenseLayer ("dense") */
  TBNN_DENSE_GAUSSIAN = 1, /* layer.py:282  GaussianDenseLayer ("denseGaussian")      */
  TBNN_ACT_RELU = 10,      /* activationFunctions.py:27  */
  TBNN_ACT_TANH = 11,      /* :53  */
  TBNN_ACT_SIGMOID = 12,   /* :40  */
  TBNN_ACT_EXP = 13,       /* :14  */
  TBNN_ACT_ELU = 14,       /* :66  */
  TBNN_ACT_LEAKYRELU = 15, /* :92  (constant slope alpha) */
  TBNN_ACT_PRELU = 16,     /* :117 */
  TBNN_ACT_SQUAREPRELU = 17 /* :274 */
};

enum { TBNN_LIK_GAUSSIAN = 0,       /* likelihood.py:63  sigma = hyper[-1]^2 */
       TBNN_LIK_FIXED_GAUSSIAN = 1, /* likelihood.py:136 sigma = fixed_sd    */
       TBNN_LIK_BERNOULLI = 2 };    /* likelihood.py:205 */

enum { TBNN_F32 = 0, TBNN_F64 = 1 };

/* tbnn_desc.flags: kernel-selection overrides used by the parity tests (results are the same
 * target either way; only the kernel that computes it changes). */
enum { TBNN_FLAG_NO_WIDE = 1,  /* never use the wide-first-layer row sweeps */
       TBNN_FLAG_NO_UMMA = 2,  /* never use the tcgen05 (tensor-core) kernels */
       TBNN_FLAG_NO_WIDE2 = 4, /* use the phase-serial wide sweep instead of the warp-specialised one */
       TBNN_FLAG_UMMA_SWEEP = 8, /* wide-first-layer row sweep on tcgen05 (3xTF32, k_sweep_umma) instead of FP32 FFMA2:
                                    correct to ~1e-6 but not yet faster (profiles/r1d_summary.md), so opt-in */
       TBNN_FLAG_NO_PERSISTENT = 16, /* never run a trajectory as one persistent launch (k_traj_small / k_traj_narrow) */
       TBNN_FLAG_NO_NARROW = 32, /* persistent trajectories on the tile engine only (k_traj_small), not k_traj_narrow */
       TBNN_FLAG_NO_UMMA_TRAIN = 64 /* hidden-layer GEMMs of the training sweep on FP32 FFMA (k_partial) instead of
                                       tcgen05 3xTF32 (k_train_umma) */ };

typedef struct {
  int32_t kind;    /* TBNN_DENSE_* or TBNN_ACT_* */
  int32_t in_dim;  /* dense: inputs; activations with slopes: width; else 0 */
  int32_t out_dim; /* dense: outputs; else 0 */
  double alpha;    /* leakyrelu slope */
} tbnn_layer_desc;

typedef struct {
  int32_t n_layers;
  const tbnn_layer_desc* layers;
  int32_t likelihood; /* TBNN_LIK_* */
  double fixed_sd;    /* FixedGaussianLikelihood(sd=) */
  int32_t dtype;      /* TBNN_F32 / TBNN_F64 */
  int32_t chains;     /* C independent chains batched per launch (reference: 1) */
  int32_t device;     /* CUDA ordinal */
  int32_t flags;      /* TBNN_FLAG_* (0 = defaults) */
} tbnn_desc;

const char* tbnn_last_error(void);
int tbnn_version(void);

/* network.__init__ + network.add (network.py:19-58, :173-191) */
int tbnn_create(const tbnn_desc* desc, tbnn_handle** out);
int tbnn_destroy(tbnn_handle* h);
int tbnn_num_params(const tbnn_handle* h);  /* P */
int tbnn_num_hypers(const tbnn_handle* h);  /* H */
/* kernels launched by this handle since creation (bench.py's gpu_launches) */
int64_t tbnn_launch_count(const tbnn_handle* h);

/* Which row-sweep kernel the handle planned: kernel_kind 0 = generic tile engine (k_partial),
 * 1 = wide-first-layer FFMA sweep (k_sweep_wide), 2 = warp-specialised wide sweep (k_sweep_wide2; forward-only
 * sweeps still use kind 1); CTAs per chain (valid after set_data), rows per
 * tile / pass, dynamic shared memory per CTA.  Any out pointer may be NULL. */
int tbnn_sweep_info(const tbnn_handle* h, int* kernel_kind, int* ctas_per_chain, int* rows_per_tile,
                    int* smem_bytes);

/* Developer aid: clock64 phase marks of CTA 0 of one (warm) wide-sweep launch; clocks_host64[0] =
 * number of marks, followed by the marks.  Fails unless tbnn_sweep_info reports kernel_kind 1. */
int tbnn_wide_profile(tbnn_handle* h, const void* theta, long long* clocks_host64, void* stream);

/* Training set resident in HBM (network.py:41-45: tf.constant).  X[N][D] row-major,
 * Y[N][out].  set_data borrows device pointers; set_data_host copies HOST buffers
 * into library-owned device memory on `stream` (the end-to-end path of bench.py).
 * With TBNN_FLAG_UMMA_SWEEP the call also re-lays X into the tcgen05 sweep's core-matrix tiles (a snapshot taken at
 * this call, on the legacy default stream for tbnn_set_data); the borrowed pointers must stay valid regardless.
 * tbnn_sweep_info: kernel_kind 0 = tile engine, 1 = phase-serial wide sweep, 2 = FFMA2 wide sweep, 3 = tcgen05 sweep. */
int tbnn_set_data(tbnn_handle* h, const void* X_dev, const void* Y_dev, int64_t n_rows);
int tbnn_set_data_host(tbnn_handle* h, const void* X_host, const void* Y_host, int64_t n_rows,
                       void* stream);

/* Main target and gradient: the closure of network.py:370-392 differentiated by TFP.
 * theta[C][P], hyper[C][H] -> logp[C], grad[C][P]; lik_stat[C] (may be NULL) receives
 * the sum of squared residuals (Gaussian kinds) or the Bernoulli log-likelihood. */
int tbnn_logp_grad(tbnn_handle* h, const void* theta, const void* hyper, void* logp, void* grad,
                   void* lik_stat, void* stream);

/* Hyper target and gradient: closure of network.py:417-440.  sse[C] may be NULL
 * (recomputed by a forward sweep when the likelihood is Gaussian). */
int tbnn_hyper_logp_grad(tbnn_handle* h, const void* theta, const void* hyper, const void* sse,
                         void* logp_h, void* grad_h, void* stream);

/* Deterministic L-step leapfrog in TFP's operation order (SimpleLeapfrogIntegrator as
 * invoked from network.py:394-408): p+=e/2 g; L x {theta+=e p; g=grad; p+=e g}; p-=e/2 g.
 * eps[C] are HOST doubles.  Outputs may alias inputs. */
int tbnn_trajectory(tbnn_handle* h, const void* theta, const void* hyper, const void* momentum,
                    const double* eps_host, int L, void* theta_out, void* mom_out, void* logp_out,
                    void* grad_out, void* stream);

/* One HMC transition of the main chain = bootstrap_results + one_step + MH of
 * sample_chain(num_results=1) (network.py:394-411).  theta[C][P] is updated in place.
 * Momentum ~ N(0,I) and u ~ U[0,1) come from Philox4x32-10 keyed by (seed, counter, chain)
 * unless injected (momentum_in[C][P], u_in[C] non-NULL; used by the parity tests).
 * stats[C][4] (dtype) = {log_accept_ratio, accept prob min(1,e^lar) (network.py:410-411),
 * accepted (0/1), squared jump distance |theta_new-theta_old|^2 (paramAdapter.py:222)}. */
int tbnn_hmc_step(tbnn_handle* h, void* theta, const void* hyper, uint64_t seed, uint64_t counter,
                  const double* eps_host, int L, const void* momentum_in, const void* u_in,
                  void* stats, void* stream);

/* The momentum tbnn_hmc_step would draw for (seed, counter): momentum_out[C][P] (dtype),
 * ke_out[C] (dtype, may be NULL) = 0.5*|p|^2.  Used by tests to pin the Philox stream. */
int tbnn_draw_momentum(tbnn_handle* h, uint64_t seed, uint64_t counter, void* momentum_out, void* ke_out,
                       void* stream);

/* Measurement hook for bench.py's roofline: launches the likelihood row-sweep kernel (the dominant
 * kernel of a leapfrog step) `iters` times on `stream` for theta[C][P], each launch bracketed by CUDA
 * events on that stream, and returns the average and minimum launch duration in milliseconds. */
int tbnn_time_sweep(tbnn_handle* h, const void* theta, int iters, float* avg_ms, float* min_ms,
                    void* stream);

/* Measurement hook for the row-sharded path (after tbnn_comm_init and one gradient evaluation): times the
 * per-gradient-evaluation exchange -- local reduction of the CTA partials + ncclAllReduce of C*(Ppad+4) values -- `iters`
 * times with CUDA events on `stream`; average / minimum in milliseconds.  Collective: every rank must call it. */
int tbnn_time_allreduce(tbnn_handle* h, int iters, float* avg_ms, float* min_ms, void* stream);

/* One HMC transition of the hyper chain + the hand-rolled dual averaging of
 * network.py:442-471 (constants :241-248).  hyper[C][H] updated in place.
 * da_state[C][3] (dtype) = {h, logEpsilonBar, hyper_step_size}, updated in place.
 * stats[C][2] = {log_accept_ratio, accept prob}.  epoch = 0-based iteration. */
int tbnn_hyper_step(tbnn_handle* h, const void* theta, void* hyper, uint64_t seed, uint64_t counter,
                    int hyperL, double epoch, double burnin, double hyper_step0, void* da_state,
                    const void* momentum_in, const void* u_in, void* stats, void* stream);

/* paramAdapter.gridSearch (paramAdapter.py:158-196): exhaustive first-maximum arg-max of
 * the UCB of calcUCB (:113-141) over eGrid x lGrid, float32 like the reference (:60).
 * All pointers are HOST float32; n_hist <= 64.  prev[n_hist][2] = (e, L) history.
 * Writes the chosen (e, L) to out_eL[2] and, if non-NULL, the best ucb to out_ucb. */
int tbnn_adapter_ucb(int device, const float* eGrid, int eNumber, const float* lGrid, int lNumber,
                     const float* prev, int n_hist, const float* Kinv, const float* KinvR, float s,
                     float p, float rootbeta, float el, float eu, float Ll, float Lu,
                     const float* sigma2x2, float* out_eL, float* out_ucb);

/* predictor.predict (predictor.py:132-155): samples[S][P] x Xtest[M][D].
 * out (may be NULL) receives [S][out][M]; moments (may be NULL) receives
 * [3][out][M] = {count, mean, M2} over the S samples (posterior predictive mean / sd). */
int tbnn_predict(tbnn_handle* h, const void* samples, int64_t S, const void* Xtest, int64_t M,
                 void* out, void* moments, void* stream);

/* Which predictor kernel tbnn_predict uses: 0 = FFMA tile engine (k_predict), 1 = tcgen05 3xTF32
 * tensor-core sweep (k_predict_umma: fp32, <= 8 inputs, <= 4 outputs, hidden widths <= 80). */
int tbnn_predict_info(const tbnn_handle* h, int* kernel_kind);

/* Row-sharded sampling (BASELINE config 4): every rank holds rows [r*N/G,(r+1)*N/G) and
 * all-reduces likelihood partials once per gradient evaluation.  unique_id is the 128-byte
 * ncclUniqueId produced on rank 0 by tbnn_comm_unique_id and distributed by host code. */
int tbnn_comm_unique_id(void* unique_id_128);
int tbnn_comm_init(tbnn_handle* h, const void* unique_id_128, int rank, int world);

#ifdef __cplusplus
}
#endif
#endif /* TBNN_H */
