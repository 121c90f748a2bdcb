// api.cu -- C ABI (include/tbnn.h): handle, host-side planner, launch sequencing, NCCL glue.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <mutex>
#include <vector>

#include "../../include/tbnn.h"
#include "kernels.h"

using namespace tbnn;

// ------------------------------------------------------------------ errors
static thread_local std::string g_err;
static int fail(const std::string& m) { g_err = m; return 1; }
#define CU(x)                                                                       \
  do {                                                                              \
    cudaError_t e_ = (x);                                                           \
    if (e_ != cudaSuccess)                                                          \
      return fail(std::string(#x) + ": " + cudaGetErrorString(e_));                 \
  } while (0)
#define CK(x) do { int r_ = (x); if (r_) return r_; } while (0)

// ------------------------------------------------------------------ NCCL through dlopen
struct NcclId { char b[128]; };
struct NcclApi {
  void* lib = nullptr;
  int (*GetUniqueId)(NcclId*) = nullptr;
  int (*CommInitRank)(void**, int, NcclId, int) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*CommDestroy)(void*) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};
static NcclApi g_nccl;
static int load_nccl() {
  if (g_nccl.lib) return 0;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* n : names) {
    g_nccl.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (g_nccl.lib) break;
  }
  if (!g_nccl.lib) return fail("dlopen(libnccl.so.2) failed; import torch first or set LD_LIBRARY_PATH");
  g_nccl.GetUniqueId = (int (*)(NcclId*))dlsym(g_nccl.lib, "ncclGetUniqueId");
  g_nccl.CommInitRank = (int (*)(void**, int, NcclId, int))dlsym(g_nccl.lib, "ncclCommInitRank");
  g_nccl.AllReduce = (int (*)(const void*, void*, size_t, int, int, void*, cudaStream_t))dlsym(
      g_nccl.lib, "ncclAllReduce");
  g_nccl.CommDestroy = (int (*)(void*))dlsym(g_nccl.lib, "ncclCommDestroy");
  g_nccl.GetErrorString = (const char* (*)(int))dlsym(g_nccl.lib, "ncclGetErrorString");
  if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllReduce)
    return fail("libnccl is missing required symbols");
  return 0;
}
#define NC(x)                                                                                     \
  do {                                                                                            \
    int r_ = (x);                                                                                 \
    if (r_ != 0)                                                                                  \
      return fail(std::string(#x) + ": " + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r_) : "nccl error")); \
  } while (0)

// ------------------------------------------------------------------ handle
struct tbnn_handle {
  int dtype = TBNN_F32, C = 1, device = 0, num_sms = 148;
  size_t esz = 4;
  ModelPlan mp;        // training plan
  ModelPlan wp;        // wide-first-layer sweep plan (fp32 only)
  ModelPlan w2;        // warp-specialised wide sweep plan (fp32, forward + backward)
  bool use_wide = false;
  bool use_wide2 = false;
  bool use_umma_predict = false;   // tcgen05 predictor (fp32, GEMM-shaped hidden layers)
  bool want_usweep = false;        // tcgen05 wide-first-layer sweep allowed for this network / dtype / flags
  bool use_usweep = false;         // ... and planned for the current data (after_data)
  ModelPlan uw;                    // its tail plan
  USweepPlan up;
  bool use_tu = false;             // tcgen05 training sweep (k_train_umma) planned for this network / dtype / flags
  TrainUmmaPlan tu;
  void* tu_wimg = nullptr;         // per-chain weight operand images (owned)
  void* tu_scratch = nullptr;      // per-CTA pre-activation scratch (owned)
  int S_tu = 1;                    // CTAs per chain of that sweep (128-row tiles)
  void* Xt = nullptr;              // training matrix in core-matrix tiles (owned)
  size_t Xt_cap = 0;
  // CUDA graphs of 2^k interior leapfrog steps (sweep + finalize each), replayed instead of 2 * 2^k launches;
  // valid for one (hyper pointer, data epoch), rebuilt when either changes
  static constexpr int GRAPH_LOG2_MAX = 5;
  cudaGraphExec_t step_graph[GRAPH_LOG2_MAX + 1] = {};
  const void* graph_hyper = nullptr;
  uint64_t graph_epoch = 0, data_epoch = 1;
  cudaStream_t cap_stream = nullptr;
  bool graphs_off = false;
  ModelPlan np;                    // plan of the narrow-network persistent trajectory kernel
  bool has_narrow = false;
  bool no_persistent = false;      // TBNN_FLAG_NO_PERSISTENT: never use the one-launch trajectory kernel
  ModelPlan pp;        // predictor (forward-only) plan
  int pp_rows = 0;     // rows per CTA of the predictor
  // data
  const void* X = nullptr;
  const void* Y = nullptr;
  void* X_own = nullptr;
  void* Y_own = nullptr;
  size_t X_cap = 0, Y_cap = 0;
  long long N = 0, N_total = 0;
  int S = 1;
  // workspace
  void *theta_pad = nullptr, *theta0_pad = nullptr, *mom_pad = nullptr, *grad_pad = nullptr;
  void* w1p = nullptr;   // pair-interleaved W_1 copy of every chain (warp-specialised wide sweep)
  void *partial = nullptr, *gsum = nullptr, *eps_dev = nullptr, *flat_tmp = nullptr, *small_T = nullptr;
  double *stat_part = nullptr, *prior_part = nullptr, *dbl = nullptr;  // dbl: [8][C] doubles
  unsigned* ticket = nullptr;
  size_t partial_cap = 0;
  int nblkF = 1;
  void* pred_ws = nullptr;
  size_t pred_ws_bytes = 0;
  // NCCL
  void* comm = nullptr;
  int rank = 0, world = 1;
  int64_t launches = 0;
  double* logp0() { return dbl; }
  double* logp1() { return dbl + C; }
  double* ke0() { return dbl + 2 * C; }
  double* stat0() { return dbl + 3 * C; }
  double* stat1() { return dbl + 4 * C; }
  double* stat_cur() { return dbl + 5 * C; }
  double* sse_tmp() { return dbl + 6 * C; }
};

static inline int pad4(int x) { return (x + 3) & ~3; }
static inline int lead(int w) { int v = pad4(w); if (((v >> 2) & 1) == 0) v += 4; return v; }
constexpr size_t SMEM_LIMIT = 227 * 1024;

static int map_act(int kind) {
  switch (kind) {
    case TBNN_ACT_RELU: return ACT_RELU;
    case TBNN_ACT_TANH: return ACT_TANH;
    case TBNN_ACT_SIGMOID: return ACT_SIGMOID;
    case TBNN_ACT_EXP: return ACT_EXP;
    case TBNN_ACT_ELU: return ACT_ELU;
    case TBNN_ACT_LEAKYRELU: return ACT_LEAKY;
    case TBNN_ACT_PRELU: return ACT_PRELU;
    case TBNN_ACT_SQUAREPRELU: return ACT_SQPRELU;
  }
  return -1;
}

// Structure of the network: blocks, flat / padded / hyper offsets.
static int plan_structure(const tbnn_desc* d, ModelPlan& mp) {
  memset(&mp, 0, sizeof(mp));
  int nb = 0, pcur = 0, fcur = 0, hcur = 0;
  for (int i = 0; i < d->n_layers; ++i) {
    const tbnn_layer_desc& L = d->layers[i];
    if (L.kind == TBNN_DENSE_CAUCHY || L.kind == TBNN_DENSE_GAUSSIAN) {
      if (nb == MAXB) return fail("too many dense layers (max 8)");
      if (L.in_dim <= 0 || L.out_dim <= 0) return fail("dense layer with non-positive dims");
      BlockPlan& b = mp.b[nb];
      if (nb > 0 && mp.b[nb - 1].out != L.in_dim) return fail("dense layer input width mismatch");
      b.in = L.in_dim; b.out = L.out_dim; b.in_p = pad4(b.in); b.out_p = pad4(b.out);
      b.ld_in = nb == 0 ? lead(b.in) : mp.b[nb - 1].ld_out;
      b.ld_out = lead(b.out);
      b.prior = L.kind == TBNN_DENSE_CAUCHY ? PRIOR_CAUCHY : PRIOR_GAUSS;
      b.act = ACT_NONE; b.alpha = 0.0;
      b.pw = pcur; pcur += b.out_p * b.ld_in;
      b.pb = pcur; pcur += b.out_p;
      b.ps = -1;
      b.fw = fcur; fcur += b.out * b.in;
      b.fb = fcur; fcur += b.out;
      b.fs = -1;
      b.hw = hcur; hcur += 4;
      b.ha = -1;
      ++nb;
    } else {
      const int a = map_act(L.kind);
      if (a < 0) return fail("unknown layer kind " + std::to_string(L.kind));
      if (nb == 0) return fail("an activation cannot precede the first dense layer");
      BlockPlan& b = mp.b[nb - 1];
      if (b.act != ACT_NONE) return fail("at most one activation per dense layer is supported");
      b.act = a; b.alpha = L.alpha;
      if (act_has_slopes(a)) {
        if (L.in_dim != b.out) return fail("prelu/squareprelu width must equal the dense output width");
        b.ps = pcur; pcur += b.out_p;
        b.fs = fcur; fcur += b.out;
        b.ha = hcur; hcur += (a == ACT_PRELU ? 1 : 2);
      }
    }
  }
  if (nb == 0) return fail("network has no dense layer");
  mp.nb = nb;
  mp.D = mp.b[0].in; mp.OUT = mp.b[nb - 1].out;
  mp.D_p = pad4(mp.D); mp.ld0 = mp.b[0].ld_in;
  mp.P = fcur; mp.Ppad = pcur;
  mp.lik = d->likelihood == TBNN_LIK_GAUSSIAN ? LIK_GAUSS
           : d->likelihood == TBNN_LIK_FIXED_GAUSSIAN ? LIK_FIXED : LIK_BERN;
  mp.fixed_sd = (double)(float)d->fixed_sd;   // tf.cast(self.sd, dtype), likelihood.py:161 (Q14)
  mp.lik_h = mp.lik == LIK_GAUSS ? hcur++ : -1;
  mp.H = hcur;
  if (mp.H > 8 * MAXB) return fail("too many hyper parameters");
  return 0;
}

// Shared-memory layout for a given tile height.  train: with z / dZ buffers and gradient
// accumulators; predict: forward only with `acc_elems` accumulator elements.
static size_t plan_smem(ModelPlan& mp, int TR, bool train, bool w_in_smem, int acc_elems, size_t esz,
                        bool g_in_smem = true, int x_bufs = 1, int g_skip = 0, int scr_min = 0) {
  mp.TR = TR;
  int cur = 0, ldmax = 4;
  mp.offX = cur; cur += x_bufs * TR * mp.ld0;
  int scr = scr_min;
  for (int l = 0; l < mp.nb; ++l) {
    BlockPlan& b = mp.b[l];
    b.offS = cur; cur += TR * b.ld_out;
    if (train && act_keeps_z(b.act)) { b.offZ = cur; cur += TR * b.ld_out; } else b.offZ = -1;
    ldmax = std::max(ldmax, b.ld_out);
    const int ntile = (TR / 4) * (b.out_p / 4);
    int ks = 1;
    if (ntile < NT) ks = std::min(std::min(NT / ntile, b.in_p / 4), 16);
    if (ks < 2) ks = 1;
    b.ksplit = ks;
    if (ks > 1) scr = std::max(scr, ks * 16 * ntile);
  }
  mp.ldmax = ldmax;
  mp.offDa = cur; cur += TR * ldmax;       // also used by the forward-only statistic sweep
  mp.offDb = cur; if (train) cur += TR * ldmax;
  mp.offScr = cur; cur += scr;
  if (w_in_smem) { mp.offW = cur; cur += mp.Ppad; } else mp.offW = -1;
  if (train && !g_in_smem) mp.offG = -1;
  else { mp.offG = cur; cur += train ? mp.Ppad - g_skip : pad4(acc_elems); }
  const int per8 = (int)(8 / esz);
  cur = (cur + per8 - 1) / per8 * per8;
  mp.offRed = cur; cur += 64 * per8;
  mp.smem_elems = cur;
  return (size_t)cur * esz;
}

static int plan_train(ModelPlan& mp, size_t esz) {
  // preference: weights + accumulators in smem; weights streamed from global (L1/L2);
  // accumulators in the CTA's global partial slice as well (very large layers / fp64)
  const bool opts[3][2] = {{true, true}, {false, true}, {false, false}};
  for (int o = 0; o < 3; ++o)
    for (int TR = 64; TR >= 4; TR -= 4)
      if (plan_smem(mp, TR, true, opts[o][0], 0, esz, opts[o][1]) <= SMEM_LIMIT) {
        if (o < 2 && TR < 8) continue;   // prefer the next option over degenerate tiles
        return 0;
      }
  return fail("network too large for the shared-memory tile engine");
}

// Wide-first-layer sweep (k_wide.cu): 16-row passes with two X buffers; W1 accumulators in
// registers; blocks >= 1 are a narrow tail with [RB]-row batch buffers (wp.TR = RB).
static bool plan_wide(const ModelPlan& mp, ModelPlan& wp) {
  if (!wide_supported(mp)) return false;
  const int WTR = wide_rows_per_pass();
  for (int RB = 4 * WTR; RB >= WTR; RB -= WTR) {
    wp = mp;
    wp.TR = RB;
    int cur = 0;
    wp.offX = cur; cur += 2 * WTR * wp.ld0;
    for (int l = 0; l < wp.nb; ++l) {
      BlockPlan& b = wp.b[l];
      b.ksplit = 1;
      b.offS = cur; cur += RB * b.ld_out;
      if (act_keeps_z(b.act)) { b.offZ = cur; cur += RB * b.ld_out; } else b.offZ = -1;
      if (l >= 1) { b.offD = cur; cur += RB * b.ld_out; } else b.offD = -1;
    }
    wp.ldmax = wp.b[0].ld_out;
    wp.offDa = cur; cur += WTR * wp.b[0].ld_out;
    wp.offDb = wp.offDa;
    wp.offScr = cur; cur += wide_scratch_elems(mp);
    wp.offW = cur; cur += wp.Ppad;
    wp.offG = cur; cur += wp.Ppad - wp.b[0].pb;
    cur = (cur + 1) / 2 * 2;
    wp.offRed = cur; cur += 64 * 2;
    wp.smem_elems = cur;
    if ((size_t)cur * 4 <= SMEM_LIMIT) return true;
  }
  return false;
}

static int plan_predict(tbnn_handle* h) {
  ModelPlan& pp = h->pp;
  pp = h->mp;
  // rows per CTA: as many as the accumulators allow, tile height up to 128
  for (int TR = 128; TR >= 4; TR -= 4) {
    size_t base = plan_smem(pp, TR, false, true, 0, h->esz);
    if (base > SMEM_LIMIT) continue;
    const size_t room = SMEM_LIMIT - base;
    int rows = (int)(room / (2 * pp.OUT * h->esz));
    rows = rows / TR * TR;
    if (rows < TR) continue;
    rows = std::min(rows, 8192);
    if (plan_smem(pp, TR, false, true, rows * pp.OUT * 2, h->esz) <= SMEM_LIMIT) {
      h->pp_rows = rows;
      return 0;
    }
  }
  return fail("network too large for the predictor kernel");
}

// ------------------------------------------------------------------ small kernels
__global__ void k_widen(const float* src, double* dst, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = (double)src[i];
}
template <typename T> __global__ void k_cast_out(const double* src, T* dst, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = (T)src[i];
}
__global__ void k_sum_stat(const double* stat_part, int S, double* out, int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double s = 0.0;
  for (int j = 0; j < S; ++j) s += stat_part[(size_t)c * S + j];
  out[c] = s;
}
template <typename T> __global__ void k_read_stat(const T* gsum, int W, int off, double* out, int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < C) out[c] = (double)gsum[(size_t)c * W + off];
}

// ------------------------------------------------------------------ API
extern "C" const char* tbnn_last_error(void) { return g_err.c_str(); }
extern "C" int tbnn_version(void) { return 100; }

extern "C" int tbnn_create(const tbnn_desc* d, tbnn_handle** out) {
  if (!d || !out) return fail("null argument");
  if (d->chains < 1) return fail("chains must be >= 1");
  if (d->dtype != TBNN_F32 && d->dtype != TBNN_F64) return fail("dtype must be TBNN_F32 or TBNN_F64");
  int ndev = 0;
  CU(cudaGetDeviceCount(&ndev));
  if (ndev == 0) return fail("no CUDA device: tensorbnn_b200 has no CPU fallback");
  if (d->device < 0 || d->device >= ndev) return fail("bad device ordinal");
  CU(cudaSetDevice(d->device));
  tbnn_handle* h = new tbnn_handle();
  h->dtype = d->dtype; h->C = d->chains; h->device = d->device;
  h->esz = d->dtype == TBNN_F32 ? 4 : 8;
  cudaDeviceProp prop;
  CU(cudaGetDeviceProperties(&prop, d->device));
  h->num_sms = prop.multiProcessorCount;
  if (plan_structure(d, h->mp) || plan_train(h->mp, h->esz)) { delete h; return 1; }
  h->use_wide = d->dtype == TBNN_F32 && !(d->flags & TBNN_FLAG_NO_WIDE) && plan_wide(h->mp, h->wp);
  h->use_wide2 = h->use_wide && !(d->flags & TBNN_FLAG_NO_WIDE2) && plan_wide2(h->mp, h->w2, SMEM_LIMIT);
  h->no_persistent = (d->flags & TBNN_FLAG_NO_PERSISTENT) != 0;
  h->has_narrow = !(d->flags & TBNN_FLAG_NO_NARROW) &&
                  (d->dtype == TBNN_F32 ? Launch<float>::plan_traj_narrow(h->mp, h->np, SMEM_LIMIT)
                                        : Launch<double>::plan_traj_narrow(h->mp, h->np, SMEM_LIMIT));
  h->use_umma_predict = d->dtype == TBNN_F32 && !(d->flags & TBNN_FLAG_NO_UMMA) && predict_umma_supported(h->mp);
  h->want_usweep = d->dtype == TBNN_F32 && (d->flags & TBNN_FLAG_UMMA_SWEEP) && !(d->flags & (TBNN_FLAG_NO_WIDE | TBNN_FLAG_NO_UMMA)) &&
                   usweep_supported(h->mp);
  h->use_tu = d->dtype == TBNN_F32 && !(d->flags & (TBNN_FLAG_NO_UMMA | TBNN_FLAG_NO_UMMA_TRAIN)) && !h->use_wide &&
              plan_train_umma(h->mp, h->tu, SMEM_LIMIT);
  if (plan_predict(h)) h->pp_rows = 0;   // predictor unavailable for this network/dtype; tbnn_predict reports it
  const ModelPlan& mp = h->mp;
  const size_t C = h->C, e = h->esz, pp = (size_t)mp.Ppad;
  h->nblkF = (mp.Ppad + 31) / 32;   // enough for both finalize variants
  if (h->use_wide2) CU(cudaMalloc(&h->w1p, C * (size_t)w1p_elems(h->mp) * sizeof(float)));
  if (h->use_tu) {
    CU(cudaMalloc(&h->tu_wimg, train_umma_wimg_bytes(h->tu, h->C)));
    CU(cudaMalloc(&h->tu_scratch, train_umma_scratch_bytes(h->tu, h->num_sms)));
  }
  CU(cudaMalloc(&h->theta_pad, C * pp * e));
  CU(cudaMalloc(&h->theta0_pad, C * pp * e));
  CU(cudaMalloc(&h->mom_pad, C * pp * e));
  CU(cudaMalloc(&h->grad_pad, C * pp * e));
  CU(cudaMalloc(&h->gsum, C * (pp + 4) * e));
  CU(cudaMalloc(&h->eps_dev, C * e));
  CU(cudaMalloc(&h->flat_tmp, C * (size_t)mp.P * e));
  CU(cudaMalloc(&h->small_T, C * 16 * e));
  CU(cudaMalloc(&h->prior_part, C * (size_t)h->nblkF * sizeof(double)));
  CU(cudaMalloc(&h->dbl, 8 * C * sizeof(double)));
  CU(cudaMalloc(&h->ticket, C * sizeof(unsigned)));
  CU(cudaMemset(h->ticket, 0, C * sizeof(unsigned)));
  CU(cudaMemset(h->dbl, 0, 8 * C * sizeof(double)));
  *out = h;
  return 0;
}

extern "C" int tbnn_destroy(tbnn_handle* h) {
  if (!h) return 0;
  cudaSetDevice(h->device);
  void* ptrs[] = {h->theta_pad, h->theta0_pad, h->mom_pad, h->grad_pad, h->gsum, h->eps_dev, h->flat_tmp,
                  h->small_T, h->prior_part, h->dbl, h->ticket, h->partial, h->stat_part, h->X_own,
                  h->Y_own, h->pred_ws, h->w1p, h->Xt, h->tu_wimg, h->tu_scratch};
  for (void* p : ptrs) if (p) cudaFree(p);
  for (auto& g : h->step_graph) if (g) cudaGraphExecDestroy(g);
  if (h->cap_stream) cudaStreamDestroy(h->cap_stream);
  if (h->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(h->comm);
  delete h;
  return 0;
}

extern "C" int tbnn_num_params(const tbnn_handle* h) { return h ? h->mp.P : -1; }
extern "C" int tbnn_num_hypers(const tbnn_handle* h) { return h ? h->mp.H : -1; }
extern "C" int64_t tbnn_launch_count(const tbnn_handle* h) { return h ? h->launches : -1; }
extern "C" int tbnn_predict_info(const tbnn_handle* h, int* kernel_kind) {
  if (!h) return fail("null handle");
  if (kernel_kind) *kernel_kind = h->use_umma_predict ? 1 : 0;
  return 0;
}
extern "C" int tbnn_sweep_info(const tbnn_handle* h, int* kernel_kind, int* ctas_per_chain, int* rows_per_tile,
                               int* smem_bytes) {
  if (!h) return fail("null handle");
  const ModelPlan& p = h->use_usweep ? h->uw : (h->use_wide2 ? h->w2 : (h->use_wide ? h->wp : h->mp));
  if (kernel_kind) *kernel_kind = h->use_tu ? 4 : (h->use_usweep ? 3 : (h->use_wide2 ? 2 : (h->use_wide ? 1 : 0)));
  if (ctas_per_chain) *ctas_per_chain = h->use_tu ? h->S_tu : h->S;
  if (rows_per_tile) *rows_per_tile = h->use_tu ? 128 : p.TR;
  if (smem_bytes) *smem_bytes = h->use_tu ? h->tu.smem_bytes : (int)((size_t)p.smem_elems * h->esz);
  return 0;
}

static int sync_n_total(tbnn_handle* h, cudaStream_t st) {
  h->N_total = h->N;
  if (!h->comm) return 0;
  long long* d = reinterpret_cast<long long*>(h->dbl + 7 * h->C);
  CU(cudaMemcpyAsync(d, &h->N, sizeof(long long), cudaMemcpyHostToDevice, st));
  NC(g_nccl.AllReduce(d, d, 1, 4 /*ncclInt64*/, 0 /*ncclSum*/, h->comm, st));
  CU(cudaMemcpyAsync(&h->N_total, d, sizeof(long long), cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  return 0;
}

static int after_data(tbnn_handle* h, long long n_rows, cudaStream_t st = 0) {
  if (n_rows <= 0) return fail("n_rows must be positive");
  h->N = n_rows;
  h->data_epoch++;                 // cached leapfrog graphs point at the old data / partial buffers
  {
    const long long smax = std::max(1, h->num_sms / h->C);
    // row-granular balanced split (wide-first-layer sweeps): every CTA gets N/S (+-1) rows, at least 8
    const int s_row = (int)std::max<long long>(1, std::min<long long>(smax, (n_rows + 7) / 8));
    // tile-granular split (generic tile engine)
    const long long ntile = (n_rows + h->mp.TR - 1) / h->mp.TR;
    const long long sm2 = std::min(smax, ntile);
    const long long q = (ntile + sm2 - 1) / sm2;
    const int s_tile = (int)((ntile + q - 1) / q);
    h->use_usweep = h->want_usweep && plan_usweep(h->mp, n_rows, s_row, SMEM_LIMIT, h->uw, h->up);
    h->S = (h->use_wide || h->use_usweep) ? s_row : s_tile;
    {
      // persistent grid of num_sms CTAs over C * S work items of ceil(ntile / S) 128-row tiles each: pick the S with the
      // smallest makespan (a quarter tile of per-item overhead: parameters, zeroing and writing the partial slice)
      const long long nt128 = (n_rows + 127) / 128;
      double best = 1e300;
      int bestS = 1;
      // A CTA keeps NT tiles in flight: a trailing group of fewer tiles costs almost as much as a full one (the tiles of
      // a group overlap each other's GEMMs and epilogues), so item lengths are counted in groups.  The 64-wide kernel
      // exists with 3 and with 2 tiles in flight (3 is 2 % faster per tile); the split decides which one fits.
      const int nt_hi = h->use_tu ? train_umma_tiles_in_flight(h->tu) : 1, nt_lo = nt_hi == 3 ? 2 : nt_hi;
      int bestNT = nt_hi;
      for (int NT = nt_hi; NT >= nt_lo; --NT) {
        for (long long S = 1; S <= std::min<long long>(nt128, h->num_sms); ++S) {
          const long long waves = ((long long)h->C * S + h->num_sms - 1) / h->num_sms;
          const long long tiles = (nt128 + S - 1) / S, rem = tiles % NT;
          const double len = (double)(tiles - rem) + (rem ? std::max((double)rem, 0.7 * NT) : 0.0);
          const double cost = (double)waves * (len + 0.25) * (NT == 2 && nt_hi == 3 ? 1.02 : 1.0);
          if (cost < best * (1.0 - 1e-9)) { best = cost; bestS = (int)S; bestNT = NT; }
        }
      }
      if (h->use_tu) h->tu.NT = bestNT;
      h->S_tu = bestS;
    }
    if (h->use_tu) h->S = h->S_tu;   // one partial count for the sweep, the forward-only statistic sweep and finalize
  }
  const size_t need = (size_t)h->C * h->S * h->mp.Ppad * h->esz;
  if (need > h->partial_cap) {
    if (h->partial) cudaFree(h->partial);
    if (h->stat_part) cudaFree(h->stat_part);
    CU(cudaMalloc(&h->partial, need));
    CU(cudaMalloc(&h->stat_part, (size_t)h->C * h->num_sms * sizeof(double) + 64));
    h->partial_cap = need;
  }
  // tcgen05 sweep: plan for this N / S and re-lay the training matrix into core-matrix tiles
  if (h->use_usweep) {
    const size_t xb = usweep_xt_bytes(h->up);
    if (xb > h->Xt_cap) {
      if (h->Xt) cudaFree(h->Xt);
      h->Xt = nullptr; h->Xt_cap = 0;
      CU(cudaMalloc(&h->Xt, xb));
      h->Xt_cap = xb;
    }
    launch_tile_x(h->up, h->mp.D, (const float*)h->X, n_rows, (float*)h->Xt, st);
    h->launches++;
    CU(cudaGetLastError());
  }
  return sync_n_total(h, 0);
}

extern "C" int tbnn_set_data(tbnn_handle* h, const void* X, const void* Y, int64_t n_rows) {
  if (!h || !X || !Y) return fail("null argument");
  CU(cudaSetDevice(h->device));
  h->X = X; h->Y = Y;
  return after_data(h, n_rows);
}

extern "C" int tbnn_set_data_host(tbnn_handle* h, const void* X, const void* Y, int64_t n_rows, void* stream) {
  if (!h || !X || !Y) return fail("null argument");
  CU(cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)stream;
  const size_t xb = (size_t)n_rows * h->mp.D * h->esz, yb = (size_t)n_rows * h->mp.OUT * h->esz;
  if (xb > h->X_cap) { if (h->X_own) cudaFree(h->X_own); CU(cudaMalloc(&h->X_own, xb)); h->X_cap = xb; }
  if (yb > h->Y_cap) { if (h->Y_own) cudaFree(h->Y_own); CU(cudaMalloc(&h->Y_own, yb)); h->Y_cap = yb; }
  CU(cudaMemcpyAsync(h->X_own, X, xb, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(h->Y_own, Y, yb, cudaMemcpyHostToDevice, st));
  h->X = h->X_own; h->Y = h->Y_own;
  return after_data(h, n_rows, st);
}

// the row sweep: wide-first-layer kernel when planned (fp32), else the generic tile engine
template <typename T>
static void sweep(tbnn_handle* h, bool backward, cudaStream_t st) {
  if (h->use_tu && backward) {
    launch_train_umma(h->mp, h->tu, h->num_sms, h->C, h->S_tu, (const float*)h->theta_pad, (unsigned char*)h->tu_wimg,
                      (float*)h->tu_scratch, (const float*)h->X, (const float*)h->Y, h->N, (float*)h->partial,
                      h->stat_part, st);
    h->launches++;
  } else if (h->use_usweep && backward) {
    launch_sweep_umma(h->uw, h->up, h->C, h->S, (const float*)h->theta_pad, (const float*)h->Xt, (const float*)h->Y,
                      h->N, (float*)h->partial, h->stat_part, st);
  } else if (h->use_wide2 && backward) {
    launch_sweep_wide2(h->w2, h->C, h->S, (const float*)h->theta_pad, (const float*)h->w1p, (const float*)h->X,
                       (const float*)h->Y, h->N, (float*)h->partial, h->stat_part, st);
  } else if (h->use_wide) {
    launch_sweep_wide(h->wp, h->C, h->S, backward, (const float*)h->theta_pad, (const float*)h->X,
                      (const float*)h->Y, h->N, (float*)h->partial, h->stat_part, st);
  } else {
    Launch<T>::partial(h->mp, h->C, h->S, backward, (const T*)h->theta_pad, (const T*)h->X, (const T*)h->Y,
                       h->N, (T*)h->partial, h->stat_part, st);
  }
  h->launches++;
}

// one likelihood sweep + gradient assembly / leapfrog update
template <typename T>
static int eval_step(tbnn_handle* h, const T* hyper, StepCoef cf, double* logp, double* stat_out,
                     cudaStream_t st) {
  const ModelPlan& mp = h->mp;
  sweep<T>(h, true, st);
  const T* gsum = nullptr;
  if (h->comm) {
    Launch<T>::reduce_partials(mp, h->C, h->S, (const T*)h->partial, h->stat_part, (T*)h->gsum, st);
    h->launches++;
    NC(g_nccl.AllReduce(h->gsum, h->gsum, (size_t)h->C * (mp.Ppad + 4), h->dtype == TBNN_F32 ? 7 : 8, 0,
                        h->comm, st));
    gsum = (const T*)h->gsum;
  }
  Launch<T>::finalize(mp, h->C, h->S, (const T*)h->partial, h->stat_part, gsum, hyper, h->N_total,
                      (T*)h->theta_pad, (T*)h->mom_pad, (T*)h->grad_pad, (const T*)h->eps_dev, cf, logp,
                      stat_out, h->prior_part, h->ticket, st, (T*)h->w1p);
  h->launches++;
  CU(cudaGetLastError());
  return 0;
}

template <typename T>
static int upload_eps(tbnn_handle* h, const double* eps_host, cudaStream_t st) {
  std::vector<T> e(h->C);
  for (int c = 0; c < h->C; ++c) e[c] = (T)eps_host[c];
  // pageable source: the runtime stages the bytes before returning, so `e` may die here
  CU(cudaMemcpyAsync(h->eps_dev, e.data(), h->C * sizeof(T), cudaMemcpyHostToDevice, st));
  return 0;
}

template <typename T>
static int cast_out(tbnn_handle* h, const double* src, void* dst, int n, cudaStream_t st) {
  k_cast_out<T><<<(n + 127) / 128, 128, 0, st>>>(src, (T*)dst, n);
  h->launches++;
  return 0;
}

static int check_ready(tbnn_handle* h) {
  if (!h) return fail("null handle");
  if (!h->X) return fail("tbnn_set_data has not been called");
  CU(cudaSetDevice(h->device));
  return 0;
}

template <typename T>
static int logp_grad_impl(tbnn_handle* h, const void* theta, const void* hyper, void* logp, void* grad,
                          void* lik_stat, cudaStream_t st) {
  const ModelPlan& mp = h->mp;
  Launch<T>::pad(mp, h->C, (const T*)theta, (T*)h->theta_pad, st, (T*)h->w1p);
  h->launches++;
  CK(eval_step<T>(h, (const T*)hyper, StepCoef{0, 0, 0}, h->logp1(), h->stat1(), st));
  if (grad) { Launch<T>::unpad(mp, h->C, (const T*)h->grad_pad, (T*)grad, st); h->launches++; }
  if (logp) CK(cast_out<T>(h, h->logp1(), logp, h->C, st));
  if (lik_stat) CK(cast_out<T>(h, h->stat1(), lik_stat, h->C, st));
  CU(cudaGetLastError());
  return 0;
}

extern "C" int tbnn_logp_grad(tbnn_handle* h, const void* theta, const void* hyper, void* logp, void* grad,
                              void* lik_stat, void* stream) {
  CK(check_ready(h));
  if (!theta || !hyper) return fail("null argument");
  return h->dtype == TBNN_F32 ? logp_grad_impl<float>(h, theta, hyper, logp, grad, lik_stat, (cudaStream_t)stream)
                              : logp_grad_impl<double>(h, theta, hyper, logp, grad, lik_stat, (cudaStream_t)stream);
}

// CUDA graph of n = 2^k interior leapfrog steps (coefficients {1, 0, 1}, no log-posterior output), captured on a
// private stream and replayed on the caller's; nullptr when graphs are unavailable (NCCL row sharding, a failed
// capture), in which case the caller launches the kernels directly.
template <typename T>
static cudaGraphExec_t interior_graph(tbnn_handle* h, const T* hyper, int k) {
  if (h->graphs_off || h->comm) return nullptr;
  if (h->graph_hyper != (const void*)hyper || h->graph_epoch != h->data_epoch) {
    for (auto& g : h->step_graph) if (g) { cudaGraphExecDestroy(g); g = nullptr; }
    h->graph_hyper = hyper; h->graph_epoch = h->data_epoch;
  }
  if (h->step_graph[k]) return h->step_graph[k];
  if (!h->cap_stream && cudaStreamCreateWithFlags(&h->cap_stream, cudaStreamNonBlocking) != cudaSuccess) {
    h->graphs_off = true; cudaGetLastError(); return nullptr;
  }
  cudaGraph_t graph = nullptr;
  const int64_t launches0 = h->launches;
  bool ok = cudaStreamBeginCapture(h->cap_stream, cudaStreamCaptureModeRelaxed) == cudaSuccess;
  if (ok) {
    for (int j = 0; j < (1 << k) && ok; ++j)
      ok = eval_step<T>(h, hyper, StepCoef{1.0, 0.0, 1.0}, nullptr, nullptr, h->cap_stream) == 0;
    ok = (cudaStreamEndCapture(h->cap_stream, &graph) == cudaSuccess) && ok && graph;
  }
  h->launches = launches0;             // nothing ran yet: replays are counted where they are launched
  if (ok) ok = cudaGraphInstantiate(&h->step_graph[k], graph, 0) == cudaSuccess;
  if (graph) cudaGraphDestroy(graph);
  if (!ok) { h->graphs_off = true; h->step_graph[k] = nullptr; cudaGetLastError(); return nullptr; }
  return h->step_graph[k];
}

// leapfrog on the padded state held in the handle (TFP order)
template <typename T>
static int leapfrog_impl(tbnn_handle* h, const T* hyper, int L, double* logp_first, double* stat_first,
                         double* logp_last, double* stat_last, cudaStream_t st) {
  if (!h->comm && !h->no_persistent && h->has_narrow && h->N <= 64) {
    // narrow networks on a few rows: the whole trajectory in one launch, one row per half-warp (k_traj_narrow)
    Launch<T>::traj_narrow(h->np, h->C, (const T*)h->X, (const T*)h->Y, h->N, hyper, h->N_total, (T*)h->theta_pad,
                           (T*)h->mom_pad, (T*)h->grad_pad, (const T*)h->eps_dev, L, logp_first, stat_first,
                           logp_last, stat_last, st);
    h->launches++;
    CU(cudaGetLastError());
    return 0;
  }
  if (!h->comm && !h->no_persistent && !h->use_wide && Launch<T>::traj_small_ok(h->mp, h->N, h->S)) {
    // small problems: the whole trajectory in one persistent launch (k_traj_small)
    Launch<T>::traj_small(h->mp, h->C, (const T*)h->X, (const T*)h->Y, h->N, hyper, h->N_total, (T*)h->theta_pad,
                          (T*)h->mom_pad, (T*)h->grad_pad, (const T*)h->eps_dev, L, logp_first, stat_first,
                          logp_last, stat_last, st);
    h->launches++;
    CU(cudaGetLastError());
    return 0;
  }
  CK(eval_step<T>(h, hyper, StepCoef{0.5, 0.0, 1.0}, logp_first, stat_first, st));
  int left = L - 1;                    // interior steps: graphs of 32, 16, ... steps, then single launches
  for (int k = tbnn_handle::GRAPH_LOG2_MAX; k >= 1 && left > 0; --k) {
    while (left >= (1 << k)) {
      cudaGraphExec_t g = interior_graph<T>(h, hyper, k);
      if (!g) break;
      CU(cudaGraphLaunch(g, st));
      h->launches += 2 << k;
      left -= 1 << k;
    }
  }
  for (; left > 0; --left) CK(eval_step<T>(h, hyper, StepCoef{1.0, 0.0, 1.0}, nullptr, nullptr, st));
  CK(eval_step<T>(h, hyper, StepCoef{1.0, 0.5, 0.0}, logp_last, stat_last, st));
  return 0;
}

template <typename T>
static int trajectory_impl(tbnn_handle* h, const void* theta, const void* hyper, const void* momentum,
                           const double* eps_host, int L, void* theta_out, void* mom_out, void* logp_out,
                           void* grad_out, cudaStream_t st) {
  const ModelPlan& mp = h->mp;
  CK(upload_eps<T>(h, eps_host, st));
  Launch<T>::pad(mp, h->C, (const T*)theta, (T*)h->theta_pad, st, (T*)h->w1p);
  Launch<T>::pad(mp, h->C, (const T*)momentum, (T*)h->mom_pad, st);
  h->launches += 2;
  CK(leapfrog_impl<T>(h, (const T*)hyper, L, nullptr, nullptr, h->logp1(), h->stat1(), st));
  if (theta_out) { Launch<T>::unpad(mp, h->C, (const T*)h->theta_pad, (T*)theta_out, st); h->launches++; }
  if (mom_out) { Launch<T>::unpad(mp, h->C, (const T*)h->mom_pad, (T*)mom_out, st); h->launches++; }
  if (grad_out) { Launch<T>::unpad(mp, h->C, (const T*)h->grad_pad, (T*)grad_out, st); h->launches++; }
  if (logp_out) CK(cast_out<T>(h, h->logp1(), logp_out, h->C, st));
  CU(cudaGetLastError());
  return 0;
}

extern "C" int tbnn_trajectory(tbnn_handle* h, const void* theta, const void* hyper, const void* momentum,
                               const double* eps_host, int L, void* theta_out, void* mom_out, void* logp_out,
                               void* grad_out, void* stream) {
  CK(check_ready(h));
  if (!theta || !hyper || !momentum || !eps_host) return fail("null argument");
  if (L < 1) return fail("L must be >= 1");
  cudaStream_t st = (cudaStream_t)stream;
  return h->dtype == TBNN_F32
             ? trajectory_impl<float>(h, theta, hyper, momentum, eps_host, L, theta_out, mom_out, logp_out, grad_out, st)
             : trajectory_impl<double>(h, theta, hyper, momentum, eps_host, L, theta_out, mom_out, logp_out, grad_out, st);
}

template <typename T>
static int hmc_step_impl(tbnn_handle* h, void* theta, const void* hyper, uint64_t seed, uint64_t counter,
                         const double* eps_host, int L, const void* momentum_in, const void* u_in,
                         void* stats, cudaStream_t st) {
  const ModelPlan& mp = h->mp;
  const size_t bytes = (size_t)h->C * mp.Ppad * sizeof(T);
  CK(upload_eps<T>(h, eps_host, st));
  Launch<T>::pad(mp, h->C, (const T*)theta, (T*)h->theta_pad, st, (T*)h->w1p);
  CU(cudaMemcpyAsync(h->theta0_pad, h->theta_pad, bytes, cudaMemcpyDeviceToDevice, st));
  Launch<T>::momentum(mp, h->C, seed, counter, (const T*)momentum_in, (T*)h->mom_pad, h->ke0(), st);
  h->launches += 2;
  CK(leapfrog_impl<T>(h, (const T*)hyper, L, h->logp0(), h->stat0(), h->logp1(), h->stat1(), st));
  Launch<T>::mh(mp, h->C, seed, counter, (const T*)u_in, (const T*)h->theta0_pad, (const T*)h->theta_pad,
                (const T*)h->mom_pad, h->logp0(), h->logp1(), h->ke0(), h->stat0(), h->stat1(),
                h->stat_cur(), (T*)theta, (T*)stats, st);
  h->launches++;
  CU(cudaGetLastError());
  return 0;
}

extern "C" int tbnn_hmc_step(tbnn_handle* h, void* theta, const void* hyper, uint64_t seed, uint64_t counter,
                             const double* eps_host, int L, const void* momentum_in, const void* u_in,
                             void* stats, void* stream) {
  CK(check_ready(h));
  if (!theta || !hyper || !eps_host) return fail("null argument");
  if (L < 1) return fail("L must be >= 1");
  cudaStream_t st = (cudaStream_t)stream;
  return h->dtype == TBNN_F32
             ? hmc_step_impl<float>(h, theta, hyper, seed, counter, eps_host, L, momentum_in, u_in, stats, st)
             : hmc_step_impl<double>(h, theta, hyper, seed, counter, eps_host, L, momentum_in, u_in, stats, st);
}

extern "C" int tbnn_draw_momentum(tbnn_handle* h, uint64_t seed, uint64_t counter, void* momentum_out,
                                  void* ke_out, void* stream) {
  if (!h || !momentum_out) return fail("null argument");
  CU(cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)stream;
  if (h->dtype == TBNN_F32) {
    Launch<float>::momentum(h->mp, h->C, seed, counter, nullptr, (float*)h->mom_pad, h->ke0(), st);
    Launch<float>::unpad(h->mp, h->C, (const float*)h->mom_pad, (float*)momentum_out, st);
    if (ke_out) CK(cast_out<float>(h, h->ke0(), ke_out, h->C, st));
  } else {
    Launch<double>::momentum(h->mp, h->C, seed, counter, nullptr, (double*)h->mom_pad, h->ke0(), st);
    Launch<double>::unpad(h->mp, h->C, (const double*)h->mom_pad, (double*)momentum_out, st);
    if (ke_out) CK(cast_out<double>(h, h->ke0(), ke_out, h->C, st));
  }
  h->launches += 2;
  CU(cudaGetLastError());
  return 0;
}

template <typename T>
static int time_sweep_impl(tbnn_handle* h, const void* theta, int iters, float* avg_ms, float* min_ms,
                           cudaStream_t st) {
  const ModelPlan& mp = h->mp;
  Launch<T>::pad(mp, h->C, (const T*)theta, (T*)h->theta_pad, st, (T*)h->w1p);
  h->launches++;
  std::vector<cudaEvent_t> ev(2 * iters);
  for (auto& e : ev) CU(cudaEventCreate(&e));
  for (int i = 0; i < iters; ++i) {
    CU(cudaEventRecord(ev[2 * i], st));
    sweep<T>(h, true, st);
    CU(cudaEventRecord(ev[2 * i + 1], st));
  }
  CU(cudaStreamSynchronize(st));
  CU(cudaGetLastError());
  double tot = 0.0;
  float mn = 1e30f;
  for (int i = 0; i < iters; ++i) {
    float ms = 0.f;
    CU(cudaEventElapsedTime(&ms, ev[2 * i], ev[2 * i + 1]));
    tot += ms;
    mn = std::min(mn, ms);
  }
  for (auto& e : ev) cudaEventDestroy(e);
  if (avg_ms) *avg_ms = (float)(tot / iters);
  if (min_ms) *min_ms = mn;
  return 0;
}

extern "C" int tbnn_time_sweep(tbnn_handle* h, const void* theta, int iters, float* avg_ms, float* min_ms,
                               void* stream) {
  CK(check_ready(h));
  if (!theta || iters < 1) return fail("bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  return h->dtype == TBNN_F32 ? time_sweep_impl<float>(h, theta, iters, avg_ms, min_ms, st)
                              : time_sweep_impl<double>(h, theta, iters, avg_ms, min_ms, st);
}

// Measurement hook for the row-sharded path: the per-gradient-evaluation exchange (k_reduce_partials + ncclAllReduce of
// C * (Ppad + 4) values), `iters` times, each bracketed by CUDA events on `stream`.
extern "C" int tbnn_time_allreduce(tbnn_handle* h, int iters, float* avg_ms, float* min_ms, void* stream) {
  CK(check_ready(h));
  if (!h->comm) return fail("tbnn_time_allreduce needs a communicator (tbnn_comm_init)");
  if (iters < 1 || !avg_ms || !min_ms) return fail("bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  cudaEvent_t e0, e1;
  CU(cudaEventCreate(&e0));
  CU(cudaEventCreate(&e1));
  double tot = 0.0;
  float mn = 1e30f;
  for (int i = 0; i < iters + 2; ++i) {
    CU(cudaEventRecord(e0, st));
    if (h->dtype == TBNN_F32)
      Launch<float>::reduce_partials(h->mp, h->C, h->S, (const float*)h->partial, h->stat_part, (float*)h->gsum, st);
    else
      Launch<double>::reduce_partials(h->mp, h->C, h->S, (const double*)h->partial, h->stat_part, (double*)h->gsum, st);
    NC(g_nccl.AllReduce(h->gsum, h->gsum, (size_t)h->C * (h->mp.Ppad + 4), h->dtype == TBNN_F32 ? 7 : 8, 0, h->comm, st));
    CU(cudaEventRecord(e1, st));
    CU(cudaEventSynchronize(e1));
    float ms = 0.f;
    CU(cudaEventElapsedTime(&ms, e0, e1));
    if (i >= 2) { tot += ms; mn = std::min(mn, ms); }
    h->launches += 1;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  *avg_ms = (float)(tot / iters);
  *min_ms = mn;
  return 0;
}

// Developer aid: phase clocks (clock64 of CTA 0) of one wide-sweep launch; clocks[0] = number of marks.
extern "C" int tbnn_wide_profile(tbnn_handle* h, const void* theta, long long* clocks_host64, void* stream) {
  CK(check_ready(h));
  if (!theta || !clocks_host64) return fail("null argument");
  if (!h->use_wide) return fail("the handle does not use the wide-first-layer sweep");
  cudaStream_t st = (cudaStream_t)stream;
  long long* d = nullptr;
  CU(cudaMalloc(&d, 64 * sizeof(long long)));
  CU(cudaMemsetAsync(d, 0, 64 * sizeof(long long), st));
  Launch<float>::pad(h->mp, h->C, (const float*)theta, (float*)h->theta_pad, st, (float*)h->w1p);
  for (int rep = 0; rep < 3; ++rep)   // the last (warm) launch is the one reported
    launch_sweep_wide(h->wp, h->C, h->S, true, (const float*)h->theta_pad, (const float*)h->X, (const float*)h->Y,
                      h->N, (float*)h->partial, h->stat_part, st, d);
  h->launches += 4;
  CU(cudaMemcpyAsync(clocks_host64, d, 64 * sizeof(long long), cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  cudaFree(d);
  return 0;
}

// SSE of the current theta (forward-only sweep) for the Gaussian likelihood's hyper term
template <typename T>
static int compute_sse(tbnn_handle* h, const void* theta, cudaStream_t st) {
  const ModelPlan& mp = h->mp;
  Launch<T>::pad(mp, h->C, (const T*)theta, (T*)h->theta_pad, st, (T*)h->w1p);
  sweep<T>(h, false, st);
  k_sum_stat<<<(h->C + 127) / 128, 128, 0, st>>>(h->stat_part, h->S, h->sse_tmp(), h->C);
  h->launches += 2;
  if (h->comm)
    NC(g_nccl.AllReduce(h->sse_tmp(), h->sse_tmp(), (size_t)h->C, 8 /*ncclDouble*/, 0, h->comm, st));
  return 0;
}

extern "C" int tbnn_hyper_logp_grad(tbnn_handle* h, const void* theta, const void* hyper, const void* sse,
                                    void* logp_h, void* grad_h, void* stream) {
  CK(check_ready(h));
  if (!theta || !hyper || !logp_h || !grad_h) return fail("null argument");
  cudaStream_t st = (cudaStream_t)stream;
  const double* sse_d = nullptr;
  if (h->mp.lik == LIK_GAUSS) {
    if (sse) {
      // caller-supplied statistic in dtype: widen to double on the device (no host round trip)
      if (h->dtype == TBNN_F32) {
        k_widen<<<(h->C + 127) / 128, 128, 0, st>>>((const float*)sse, h->sse_tmp(), h->C);
        h->launches++;
      } else {
        CU(cudaMemcpyAsync(h->sse_tmp(), sse, h->C * 8, cudaMemcpyDeviceToDevice, st));
      }
    } else {
      CK(h->dtype == TBNN_F32 ? compute_sse<float>(h, theta, st) : compute_sse<double>(h, theta, st));
    }
    sse_d = h->sse_tmp();
  }
  if (h->dtype == TBNN_F32)
    Launch<float>::hyper_eval(h->mp, h->C, (const float*)theta, (const float*)hyper, sse_d, h->N_total,
                              (float*)logp_h, (float*)grad_h, st);
  else
    Launch<double>::hyper_eval(h->mp, h->C, (const double*)theta, (const double*)hyper, sse_d, h->N_total,
                               (double*)logp_h, (double*)grad_h, st);
  h->launches++;
  CU(cudaGetLastError());
  return 0;
}

extern "C" int tbnn_hyper_step(tbnn_handle* h, const void* theta, void* hyper, uint64_t seed, uint64_t counter,
                               int hyperL, double epoch, double burnin, double hyper_step0, void* da_state,
                               const void* momentum_in, const void* u_in, void* stats, void* stream) {
  CK(check_ready(h));
  if (!theta || !hyper || !da_state) return fail("null argument");
  if (hyperL < 1) return fail("hyperL must be >= 1");
  cudaStream_t st = (cudaStream_t)stream;
  const double* sse_d = nullptr;
  if (h->mp.lik == LIK_GAUSS) {
    CK(h->dtype == TBNN_F32 ? compute_sse<float>(h, theta, st) : compute_sse<double>(h, theta, st));
    sse_d = h->sse_tmp();
  }
  if (h->dtype == TBNN_F32)
    Launch<float>::hyper_step(h->mp, h->C, (const float*)theta, (float*)hyper, sse_d, h->N_total, seed, counter,
                              hyperL, epoch, burnin, hyper_step0, (float*)da_state, (const float*)momentum_in,
                              (const float*)u_in, (float*)stats, st);
  else
    Launch<double>::hyper_step(h->mp, h->C, (const double*)theta, (double*)hyper, sse_d, h->N_total, seed,
                               counter, hyperL, epoch, burnin, hyper_step0, (double*)da_state,
                               (const double*)momentum_in, (const double*)u_in, (double*)stats, st);
  h->launches++;
  CU(cudaGetLastError());
  return 0;
}

extern "C" int tbnn_adapter_ucb(int device, const float* eGrid, int eNumber, const float* lGrid, int lNumber,
                                const float* prev, int n_hist, const float* Kinv, const float* KinvR, float s,
                                float p, float rootbeta, float el, float eu, float Ll, float Lu,
                                const float* sigma2x2, float* out_eL, float* out_ucb) {
  if (!eGrid || !lGrid || !prev || !Kinv || !KinvR || !sigma2x2 || !out_eL) return fail("null argument");
  if (n_hist < 1 || n_hist > 64) return fail("n_hist must be in [1, 64]");
  if (eNumber < 1 || lNumber < 1) return fail("empty grid");
  CU(cudaSetDevice(device));
  const size_t fl = (size_t)eNumber + lNumber + 2 * n_hist + (size_t)n_hist * n_hist + n_hist + 4;
  const size_t wsb = adapter_workspace_bytes(eNumber, lNumber);
  // one device buffer per device, kept between calls (the adapter decides every few epochs: no cudaMalloc / cudaFree
  // on the sampling path); a call is still synchronous -- its result steers the next epoch
  static std::mutex mu;
  static char* cache[64] = {};
  static size_t cache_bytes[64] = {};
  std::lock_guard<std::mutex> lock(mu);
  const size_t need = fl * sizeof(float) + wsb + 256;
  if (device < 0 || device >= 64) return fail("bad device ordinal");
  if (cache_bytes[device] < need) {
    if (cache[device]) cudaFree(cache[device]);
    cache[device] = nullptr; cache_bytes[device] = 0;
    CU(cudaMalloc(&cache[device], need));
    cache_bytes[device] = need;
  }
  char* buf = cache[device];
  std::vector<float> host(fl);
  size_t o = 0;
  const size_t oE = o; memcpy(&host[o], eGrid, eNumber * 4); o += eNumber;
  const size_t oL = o; memcpy(&host[o], lGrid, lNumber * 4); o += lNumber;
  const size_t oP = o; memcpy(&host[o], prev, 2 * n_hist * 4); o += 2 * n_hist;
  const size_t oK = o; memcpy(&host[o], Kinv, (size_t)n_hist * n_hist * 4); o += (size_t)n_hist * n_hist;
  const size_t oR = o; memcpy(&host[o], KinvR, n_hist * 4); o += n_hist;
  const size_t oO = o;
  float* d = reinterpret_cast<float*>(buf);
  cudaError_t e = cudaMemcpy(d, host.data(), fl * sizeof(float), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) {
    void* ws = buf + ((fl * sizeof(float) + 255) / 256) * 256;
    launch_adapter_ucb(d + oE, eNumber, d + oL, lNumber, d + oP, n_hist, d + oK, d + oR, s, p, rootbeta, el, eu,
                       Ll, Lu, sigma2x2, d + oO, ws, 0);
    e = cudaGetLastError();
  }
  float res[3] = {0, 0, 0};
  if (e == cudaSuccess) e = cudaMemcpy(res, d + oO, 3 * sizeof(float), cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) return fail(std::string("adapter_ucb: ") + cudaGetErrorString(e));
  out_eL[0] = res[0]; out_eL[1] = res[1];
  if (out_ucb) *out_ucb = res[2];
  return 0;
}

template <typename T>
static int predict_impl(tbnn_handle* h, const void* samples, int64_t S, const void* Xtest, int64_t M, void* out,
                        void* moments, cudaStream_t st) {
  const ModelPlan& pp = h->pp;
  // samples are padded chunk by chunk into a bounded workspace
  const size_t per = (size_t)pp.Ppad * sizeof(T);
  const int64_t chunk = std::max<int64_t>(1, std::min<int64_t>(S, (int64_t)((256ull << 20) / per)));
  if (chunk * per > h->pred_ws_bytes) {
    if (h->pred_ws) cudaFree(h->pred_ws);
    CU(cudaMalloc(&h->pred_ws, chunk * per));
    h->pred_ws_bytes = chunk * per;
  }
  // rows per CTA: fill the SMs when M is small
  int rows = h->pp_rows;
  const int64_t want = (M + h->num_sms - 1) / h->num_sms;
  if (want < rows) rows = (int)std::max<int64_t>(pp.TR, (want + pp.TR - 1) / pp.TR * pp.TR);
  for (int64_t s0 = 0; s0 < S; s0 += chunk) {
    const int64_t sc = std::min<int64_t>(chunk, S - s0);
    // pad uses the training plan's offsets (identical parameter layout) with C = sc "chains"
    ModelPlan tmp = h->mp;
    const int64_t maxy = 32768;
    for (int64_t b0 = 0; b0 < sc; b0 += maxy) {
      const int nb = (int)std::min<int64_t>(maxy, sc - b0);
      Launch<T>::pad(tmp, nb, (const T*)samples + (size_t)(s0 + b0) * pp.P, (T*)h->pred_ws + (size_t)b0 * pp.Ppad, st);
      h->launches++;
    }
    Launch<T>::predict(pp, (const T*)h->pred_ws, s0, sc, S, (const T*)Xtest, M, rows, (T*)out, (T*)moments, st);
    h->launches++;
  }
  CU(cudaGetLastError());
  return 0;
}

extern "C" int tbnn_predict(tbnn_handle* h, const void* samples, int64_t S, const void* Xtest, int64_t M,
                            void* out, void* moments, void* stream) {
  if (!h || !samples || !Xtest) return fail("null argument");
  if (S < 1 || M < 1) return fail("S and M must be positive");
  if (!out && !moments) return fail("one of out / moments must be given");
  CU(cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)stream;
  if (h->use_umma_predict) {
    if (!launch_predict_umma(h->mp, h->num_sms, (const float*)samples, 0, S, (const float*)Xtest, M, (float*)out,
                             (float*)moments, st))
      return fail("internal: tcgen05 predictor plan rejected at launch");
    h->launches++;
    CU(cudaGetLastError());
    return 0;
  }
  if (h->pp_rows <= 0) return fail("network too large for the predictor kernel (weights must fit in shared memory)");
  return h->dtype == TBNN_F32 ? predict_impl<float>(h, samples, S, Xtest, M, out, moments, st)
                              : predict_impl<double>(h, samples, S, Xtest, M, out, moments, st);
}

extern "C" int tbnn_comm_unique_id(void* unique_id_128) {
  if (!unique_id_128) return fail("null argument");
  CK(load_nccl());
  NcclId id;
  NC(g_nccl.GetUniqueId(&id));
  memcpy(unique_id_128, &id, sizeof(id));
  return 0;
}

extern "C" int tbnn_comm_init(tbnn_handle* h, const void* unique_id_128, int rank, int world) {
  if (!h || !unique_id_128) return fail("null argument");
  CK(load_nccl());
  CU(cudaSetDevice(h->device));
  NcclId id;
  memcpy(&id, unique_id_128, sizeof(id));
  NC(g_nccl.CommInitRank(&h->comm, world, id, rank));
  h->rank = rank; h->world = world;
  if (h->X) return sync_n_total(h, 0);
  return 0;
}
