// kernels.h -- host-side launchers of the CUDA kernels (one explicit instantiation per dtype).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "plan.h"

namespace tbnn {

// p = a1*eps*g applied as: p += m1*eps*g ; p -= m2*eps*g ; theta += m3*eps*p   (per chain eps)
struct StepCoef { double m1, m2, m3; };

// finalize uses the S-split variant (CTA = 32 parameters x 8 S-slices) from this many partials on;
// prior_part must hold C * ceil(Ppad/32) doubles.
constexpr int FINALIZE_SPLIT_MIN_S = 16;

template <typename T> struct Launch {
  // flat <-> padded
  // w1p (optional): pair-interleaved copy of W_1 kept in step with `padded` (k_wide2.cu)
  static void pad(const ModelPlan& mp, int C, const T* flat, T* padded, cudaStream_t st, T* w1p = nullptr);
  static void unpad(const ModelPlan& mp, int C, const T* padded, T* flat, cudaStream_t st);
  // likelihood partial gradients: grid (S, C)
  static void partial(const ModelPlan& mp, int C, int S, bool backward, const T* theta_pad,
                      const T* X, const T* Y, long long N, T* partial, double* stat_part,
                      cudaStream_t st);
  // local reduction of partials into gsum[C][Ppad+4] (stat at [Ppad]); used before an all-reduce
  static void reduce_partials(const ModelPlan& mp, int C, int S, const T* partial,
                              const double* stat_part, T* gsum, cudaStream_t st);
  // gradient assembly (+ priors) and leapfrog update.  If gsum != nullptr it is used
  // instead of partial/stat_part (S ignored).
  static void finalize(const ModelPlan& mp, int C, int S, const T* partial, const double* stat_part,
                       const T* gsum, const T* hyper, long long N_total, T* theta_pad, T* mom_pad,
                       T* grad_pad, const T* eps_dev, StepCoef cf, double* logp, double* stat_out,
                       double* prior_part, unsigned* ticket, cudaStream_t st, T* w1p = nullptr);
  // whole L-step trajectory in ONE launch (one CTA per chain) when the training set is a single tile and the
  // parameters fit the shared-memory plan; same arithmetic as partial (S = 1) + finalize
  static bool traj_small_ok(const ModelPlan& mp, long long N, int S);
  static void traj_small(const ModelPlan& mp, int C, const T* X, const T* Y, long long N, const T* hyper,
                         long long N_total, T* theta_pad, T* mom_pad, T* grad_pad, const T* eps_dev, int L,
                         double* logp_first, double* stat_first, double* logp_last, double* stat_last,
                         cudaStream_t st);
  // the same for narrow networks (every width <= 32, <= 64 rows): one row per half-warp (narrow.cuh)
  static bool plan_traj_narrow(const ModelPlan& mp, ModelPlan& np, size_t smem_limit);
  static void traj_narrow(const ModelPlan& np, int C, const T* X, const T* Y, long long N, const T* hyper,
                          long long N_total, T* theta_pad, T* mom_pad, T* grad_pad, const T* eps_dev, int L,
                          double* logp_first, double* stat_first, double* logp_last, double* stat_last,
                          cudaStream_t st);
  // momentum ~ N(0, I) (Philox) or copy of injected flat momentum; ke[c] = 0.5*sum p^2
  static void momentum(const ModelPlan& mp, int C, uint64_t seed, uint64_t call, const T* injected_flat,
                       T* mom_pad, double* ke, cudaStream_t st);
  // Metropolis-Hastings select; writes flat theta and stats[C][4]
  static void mh(const ModelPlan& mp, int C, uint64_t seed, uint64_t call, const T* u_in,
                 const T* theta0_pad, const T* theta1_pad, const T* mom1_pad, const double* logp0,
                 const double* logp1, const double* ke0, const double* stat0, const double* stat1,
                 double* stat_cur, T* theta_flat, T* stats, cudaStream_t st);
  // hyper chain: one CTA per chain
  static void hyper_eval(const ModelPlan& mp, int C, const T* theta_flat, const T* hyper,
                         const double* sse, long long N_total, T* logp, T* grad, cudaStream_t st);
  static void hyper_step(const ModelPlan& mp, int C, const T* theta_flat, T* hyper, const double* sse,
                         long long N_total, uint64_t seed, uint64_t call, int L, double epoch,
                         double burnin, double hyper_step0, T* da_state, const T* mom_in,
                         const T* u_in, T* stats, cudaStream_t st);
  // predictor: samples already padded [S][Ppad]
  static void predict(const ModelPlan& mp, const T* samples_pad, long long s0, long long S_chunk,
                      long long S_total, const T* X, long long M, int rows_per_cta, T* out,
                      T* moments, cudaStream_t st);
};

// wide-first-layer row sweep (k_wide.cu), fp32 only; `wp` is the wide plan (TR = wide_rows_per_pass()).
bool wide_supported(const ModelPlan& mp);
int wide_rows_per_pass();
int wide_scratch_elems(const ModelPlan& mp);
void launch_sweep_wide(const ModelPlan& wp, int C, int S, bool backward, const float* theta_pad,
                       const float* X, const float* Y, long long N, float* partial, double* stat_part,
                       cudaStream_t st, long long* prof = nullptr);

// warp-specialised wide-first-layer sweep (k_wide2.cu), fp32, forward + backward only; `wp` from plan_wide2.
// It reads W_1 from a pair-interleaved copy w1p[C][w1p_elems] ([k quad][output pair][k in quad][2]) that
// k_pad / k_finalize keep in step with theta_pad.
__host__ __device__ inline int w1p_quad_stride(int out_p) { return 4 * out_p + 4; }
__host__ __device__ inline int w1p_elems(const ModelPlan& mp) { return (mp.b[0].in_p >> 2) * w1p_quad_stride(mp.b[0].out_p); }
bool wide2_supported(const ModelPlan& mp);
bool plan_wide2(const ModelPlan& mp, ModelPlan& wp, size_t smem_limit);
void launch_sweep_wide2(const ModelPlan& wp, int C, int S, const float* theta_pad, const float* w1p,
                        const float* X, const float* Y, long long N, float* partial, double* stat_part,
                        cudaStream_t st);

// tcgen05 (3xTF32) wide-first-layer row sweep (k_sweep_umma.cu), fp32, forward + backward.  The training matrix is
// re-laid once per tbnn_set_data into core-matrix tiles (launch_tile_x); the plan depends on N and S.
struct USweepPlan {
  int FC, nch;         // features per chunk (multiple of 8, <= 128); chunks = ceil((D + 1) / FC)
  int TRc, ntiles;     // row capacity of a tile; tiles in total (tile t = rows [N t / ntiles, N (t+1) / ntiles))
  int NP, NM;          // block-0 outputs staged per half (multiple of 8) / rows of a stacked [hi; lo] operand (64)
  int xbytes, wbytes;  // bytes of one X chunk (hi or lo) / of one stacked W1 chunk
  int stage_bytes;     // X hi, X lo, [W hi; W lo]
  int off_stage;       // byte offsets into dynamic shared memory
  int off_dz, dzbytes, dz_cg;   // stacked dZ1 operand, its column-group stride
  int off_wtt, wtt_ofs[MAXB];   // transposed weights of blocks >= 1 (floats from off_wtt)
  int off_wt, off_g, off_red, off_dwx;   // tail parameters / accumulators, reduction scratch + barriers, drain buffer
  int smem_bytes;
};
bool usweep_supported(const ModelPlan& mp);
bool plan_usweep(const ModelPlan& mp, long long N, int S, size_t smem_limit, ModelPlan& wp, USweepPlan& up);
size_t usweep_xt_bytes(const USweepPlan& up);
void launch_tile_x(const USweepPlan& up, int D, const float* X, long long N, float* Xt, cudaStream_t st);
void launch_sweep_umma(const ModelPlan& wp, const USweepPlan& up, int C, int S, const float* theta_pad,
                       const float* Xt, const float* Y, long long N, float* partial, double* stat_part,
                       cudaStream_t st);

// tcgen05 (3xTF32) training row sweep for networks whose hidden blocks are GEMM-shaped (k_train_umma.cu), fp32,
// forward + backward; same partial / stat_part contract as Launch<float>::partial.
struct TrainUmmaPlan {
  int G;                        // GEMM blocks 0..G-1 (the hidden blocks); block G = last block, on the CUDA cores
  int HW;                       // hidden width: out of blocks 0..G-1 and in of blocks 1..G (64 or 128)
  int K0p, nK0;                 // padded input width (multiple of 8) and its 32-deep chunks
  int N0w;                      // N of the block-0 weight-gradient GEMM: pad16(D + 1)
  int act;                      // activation shared by the hidden blocks
  int NTmax;                    // tile slots the shared-memory plan has room for (64-wide: 3 or 2)
  int NT;                       // tiles a CTA keeps in flight: 1 (128-wide), 2 or 3 (64-wide; chosen with the work-item split)
  int TPR;                      // threads per training row in the row-worker warps (2, or 4 for 64-wide networks)
  int fx_stride;                // floats per (thread, row) of the last-block exchange buffer: 1, 2 or 4 (>= OUT)
  int wimg_chain;               // bytes of one chain's weight operand images
  int fimg[MAXB], bimg[MAXB];   // byte offsets of the forward / backward operand image of block l
  int a_stage, b_stage;         // bytes of one ring stage
  int off_a, off_b, off_par, off_bar;                          // shared-memory byte offsets
  int par_bias, par_slope, par_sraw, par_wl, par_accl, par_accb, par_fx;   // float offsets inside the parameter region
  int smem_bytes;
  int scratch_cta;              // floats of pre-activation scratch per CTA
};
bool plan_train_umma(const ModelPlan& mp, TrainUmmaPlan& tp, size_t smem_limit);
size_t train_umma_wimg_bytes(const TrainUmmaPlan& tp, int C);
int train_umma_tiles_in_flight(const TrainUmmaPlan& tp);   // most tiles a CTA can keep in flight (tp.NT = the number in use)
size_t train_umma_scratch_bytes(const TrainUmmaPlan& tp, int num_sms);
void launch_train_umma(const ModelPlan& mp, const TrainUmmaPlan& tp, int num_sms, int C, int S, const float* theta_pad,
                       unsigned char* wimg, float* scratch, const float* X, const float* Y, long long N,
                       float* partial, double* stat_part, cudaStream_t st);

// tcgen05 (3xTF32) posterior-predictive sweep (k_predict_umma.cu), fp32 only; samples are FLAT [S][P].
bool predict_umma_supported(const ModelPlan& mp);
bool launch_predict_umma(const ModelPlan& mp, int num_sms, const float* samples, long long s0, long long S_chunk,
                         const float* X, long long M, float* out, float* moments, cudaStream_t st);

void launch_adapter_ucb(const float* eGrid, int eNumber, const float* lGrid, int lNumber,
                        const float* prev, int n_hist, const float* Kinv, const float* KinvR, float s,
                        float p, float rootbeta, float el, float eu, float Ll, float Lu,
                        const float* sigma, float* out /*[3] e, L, ucb*/, void* workspace,
                        cudaStream_t st);
size_t adapter_workspace_bytes(int eNumber, int lNumber);

}  // namespace tbnn
