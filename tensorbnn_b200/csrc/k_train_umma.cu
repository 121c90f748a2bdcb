// k_train_umma.cu -- the row sweep (log-likelihood + full parameter gradient of one leapfrog step) with every
// hidden-layer contraction on the 5th-generation tensor cores (tcgen05.mma kind::tf32, error-compensated 3xTF32,
// fp32 accumulators in tensor memory).  Same contract as k_partial (k_main.cu): per (chain, CTA slot) a partial
// gradient in the padded layout plus a likelihood statistic; k_finalize adds priors and does the leapfrog update.
//
// Reference arithmetic replaced: tf.matmul(W, A) of layer.predict (layer.py:278) for the forward pass and the
// two GEMMs per layer TF's autodiff adds for the backward pass (dA = dZ W, dW = dZ^T A), invoked from TFP's
// leapfrog through network.py:370-408; activations (activationFunctions.py), likelihood residuals
// (likelihood.py:88-94,162-167,225-236) and the last (<= 4 output) layer stay on the CUDA cores.
//
// Shapes: networks whose dense blocks 0..G-1 all have the same output width HW in {64, 128} and the same
// activation, followed by one last block with <= 4 outputs (C3: 1-64-64-64-1 SquarePrelu, C4: 32-128-128-128-1 ReLU).
//
// One persistent CTA per SM, three roles over two operand rings in shared memory (4 stages for the 64-wide network, 3
// for the 128-wide one):
//   4 TPR row-worker warps: TPR threads share a training row of the 128-row tile (TPR = 4 for the 64-wide network, 2 for
//              the 128-wide one), each owning HW / TPR columns.  They read accumulators from tensor memory
//              (tcgen05.ld), apply bias / activation / derivatives, and WRITE MMA OPERANDS: K-major chunks of 32 features
//              for the forward (A_l) and data-gradient (dZ_l) GEMMs, and transposed (K = rows) chunks for the
//              weight-gradient GEMMs -- chunk q of those holds the 32 rows of lane quarter q.  hi/lo TF32 split, plain
//              SWIZZLE_NONE core matrices whose column-group stride is padded by 16 bytes so the transposed 4-byte
//              stores are bank-conflict free.
//   1 warp     MMA issuer (one thread): per 32-deep chunk and 8-deep k step lo*hi + hi*lo + hi*hi.  Its descriptors are
//              32-bit words on the uniform datapath (umma::mma_tf32_ss32) -- see the comment there.
//   1 warp     TMA producer (one thread): streams pre-split, pre-tiled weight operands (k_train_prep) from L2 with
//              cp.async.bulk; a 128-wide network does not fit its weights (2 layers x 2 orientations x hi/lo x 64 KB)
//              in shared memory, so they are streamed per tile (~0.5 MB / tile, L2 resident).
// Two or three tiles are in flight per CTA (64-wide: 3 or 2, 128-wide: 2), alternating segment by segment, so one
// tile's epilogue overlaps the others' GEMMs.
// GEMM order per tile: F_0 .. F_{G-1} | B_{G-1}, W_{G-1}, .., B_1, W_1, W_0   (F: Z_l = A_{l-1} W_l^T, B: dA_{l-1} =
// dZ_l W_l, W: [dW_l | db_l] = dZ_l^T [A_{l-1} | 1]).  The bias gradient falls out of a constant-one operand row; with
// HW = 64 the free upper half of the M = 128 operand carries z*dA so the slope gradients (Prelu / SquarePrelu) fall
// out of the same column.  (128-wide network with two tiles in flight: tensor memory has no room for that column, the
// bias gradients of the hidden blocks are warp column sums.)  Weight gradients are drained per tile from tensor memory and added to this CTA's slice
// of the partial buffer with vector reductions (red.global.add.v4.f32): the accumulation chain inside the tensor
// core stays 48 MMAs long (its accumulator rounds toward zero, profiles/r1d_summary.md).
// Pre-activations of blocks 0..G-2 are parked in a per-CTA global scratch (L2 resident) between forward and backward.
#include "engine.cuh"
#include "kernels.h"
#include "umma.cuh"

namespace tbnn {
#ifndef TU_SHFL128
#define TU_SHFL128 false
#endif
#ifndef TU_ROLL_ALL
#define TU_ROLL_ALL false
#endif

// TPR threads share a training row (each owns HW / TPR columns): 4 TPR row-worker warps, then the MMA issuer warp
// (+ TMEM alloc) and the TMA producer warp.  TPR = 2: 320 threads, 168 registers.  TPR = 4 (64-wide networks): 576
// threads, 16 columns per thread, 112 registers -- four row-worker warps per scheduler instead of two hide the
// dependent-issue latency this kernel is bound by (profiles/r2h_summary.md).
__host__ __device__ constexpr int tu_threads(int tpr) { return 32 * (4 * tpr + 2); }
constexpr int TU_NS_MAX = 4;                   // ring stages: 4 for the 64-wide network (all four chunks of a weight-gradient GEMM
                                               // in flight at once), 3 for the 128-wide one (shared memory)
__host__ __device__ constexpr int tu_ns(int hw) { return hw == 64 ? 4 : 3; }
constexpr int TU_CGA = 128 * 16 + 16;          // column-group stride (bytes) of a 128-row operand chunk
constexpr int TU_HALFA = 8 * TU_CGA;           // bytes of its hi (or lo) half
constexpr int TU_ASTAGE = 2 * TU_HALFA;

__host__ __device__ constexpr int tu_cgs(int rows) { return 128 * (rows / 8) + 16; }
__host__ __device__ constexpr int tu_half(int rows) { return 8 * tu_cgs(rows); }

template <int NTHR> __device__ __forceinline__ void ew_barrier() { asm volatile("bar.sync 1, %0;\n" ::"n"(NTHR) : "memory"); }
__device__ __forceinline__ void mbar_arrive_n(uint64_t* bar, uint32_t n) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(n) : "memory");
}
__device__ __forceinline__ void red_add4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};\n" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void red_add1(float* p, float a) {
  asm volatile("red.global.add.f32 [%0], %1;\n" ::"l"(p), "f"(a) : "memory");
}
__device__ __forceinline__ void sts32(uint32_t saddr, float v) {
  asm volatile("st.shared.f32 [%0], %1;\n" ::"r"(saddr), "f"(v) : "memory");
}
__device__ __forceinline__ void sts128(uint32_t saddr, float a, float b, float c, float d) {
  asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};\n" ::"r"(saddr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
// 16 accumulator columns
__device__ __forceinline__ void tmem_ld16w(uint32_t taddr, float (&v)[16]) {
  umma::tmem_ld16(taddr, v);
  umma::tmem_ld_wait();
}
// column sums of 16 values per lane over the 32 lanes of a warp: on return every lane holds the sum of column
// colsum16_col(lane) (16 shuffles; lanes 2i and 2i + 1 hold the same column)
__device__ __forceinline__ float colsum16(float (&v)[16], int lane) {
#pragma unroll
  for (int s = 8; s >= 1; s >>= 1) {
    const bool up = (lane & (2 * s)) != 0;
#pragma unroll
    for (int i = 0; i < s; ++i) {
      const float send = up ? v[i] : v[i + s];
      const float keep = up ? v[i + s] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 2 * s);
    }
  }
  return v[0] + __shfl_xor_sync(0xffffffffu, v[0], 1);
}
__device__ __forceinline__ int colsum16_col(int lane) { return lane >> 1; }

// column sums over the 32 lanes of a warp: on return lane i holds sum over lanes of v[i] (31 shuffles)
__device__ __forceinline__ float colsum32(float (&v)[32], int lane) {
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) {
    const bool up = (lane & s) != 0;
#pragma unroll
    for (int i = 0; i < s; ++i) {
      const float send = up ? v[i] : v[i + s];
      const float keep = up ? v[i + s] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, s);
    }
  }
  return v[0];
}

// ------------------------------------------------------------------ weight operand images
// Per chain: forward image of block l (B operand [N = HW][K = in_l], K-major) for l < G and backward image
// (B operand [N = in_l][K = HW] = W_l^T) for 1 <= l < G, in chunks of 32 k: hi half then lo half, each 8 column
// groups of tu_cgs(HW) bytes; 16-byte group (n, kq) at (n / 8) * 128 + kq * cgs + (n % 8) * 16.
__global__ void k_train_prep(const __grid_constant__ ModelPlan mp, const __grid_constant__ TrainUmmaPlan tp,
                             const float* __restrict__ theta_pad, unsigned char* __restrict__ wimg) {
  const int c = blockIdx.y;
  const int gid = blockIdx.x * blockDim.x + threadIdx.x;
  const int HW = tp.HW, nH = HW / 32, gper = HW * 8;      // groups per chunk
  // image list: F_0 (nK0 chunks), F_1..F_{G-1} (nH chunks each), B_1..B_{G-1}
  int idx = gid, l = -1, bwd = 0, nch = 0;
  for (int i = 0; i < 2 * tp.G - 1; ++i) {
    const int li = i < tp.G ? i : i - tp.G + 1;
    const int n = (i == 0 ? tp.nK0 : nH) * gper;
    if (idx < n) { l = li; bwd = i >= tp.G; nch = (i == 0 ? tp.nK0 : nH); break; }
    idx -= n;
  }
  if (l < 0) return;
  (void)nch;
  const BlockPlan& b = mp.b[l];
  const int chunk = idx / gper, rem = idx - chunk * gper, n = rem >> 3, kq = rem & 7, k0 = chunk * 32 + kq * 4;
  const float* th = theta_pad + (size_t)c * mp.Ppad + b.pw;
  float v[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int k = k0 + i;
    if (!bwd) v[i] = (n < b.out && k < b.in) ? th[n * b.ld_in + k] : 0.f;     // W_l[n][k]
    else v[i] = (k < b.out && n < b.in) ? th[k * b.ld_in + n] : 0.f;          // W_l[k][n]
  }
  float hi[4], lo[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) umma::split_tf32_fast(v[i], hi[i], lo[i]);
  const int cgs = tu_cgs(HW);
  unsigned char* base = wimg + (size_t)c * tp.wimg_chain + (bwd ? tp.bimg[l] : tp.fimg[l]) + (size_t)chunk * 2 * tu_half(HW);
  const int off = (n >> 3) * 128 + kq * cgs + (n & 7) * 16;
  *reinterpret_cast<float4*>(base + off) = make_float4(hi[0], hi[1], hi[2], hi[3]);
  *reinterpret_cast<float4*>(base + tu_half(HW) + off) = make_float4(lo[0], lo[1], lo[2], lo[3]);
}

// ------------------------------------------------------------------ activation helpers (compile-time kind)
// ACTK: ACT_RELU | ACT_SQPRELU (any slope-type activation: z is kept, `slope` is the effective negative-side
// slope) | -1 (parameter-free activation evaluated through act_fwd / act_deriv_from_out)
template <int ACTK> __device__ __forceinline__ float tu_keep(int act, float z, float a) {   // what the backward pass needs
  return ACTK == ACT_SQPRELU ? z : a;
}
template <int ACTK> __device__ __forceinline__ float tu_act(int act, float z, float slope) {
  if (ACTK == ACT_RELU) return fmaxf(z, 0.f);
  if (ACTK == ACT_SQPRELU) return z < 0.f ? slope * z : z;
  return act_fwd<float>(act, z, 0.f);
}
template <int ACTK> __device__ __forceinline__ float tu_from_keep(int act, float s, float slope) {   // activation output from the kept value
  if (ACTK == ACT_SQPRELU) return s < 0.f ? slope * s : s;
  return s;
}
template <int ACTK> __device__ __forceinline__ float tu_deriv(int act, float s, float slope) {
  if (ACTK == ACT_RELU) return s > 0.f ? 1.f : 0.f;
  if (ACTK == ACT_SQPRELU) return s < 0.f ? slope : 1.f;
  return act_deriv_from_out<float>(act, s);
}

// NCG column groups (4 features each, hi and lo halves), starting at group kq0, of this thread's row of a K-major
// operand chunk of 32 features
template <int NCG>
__device__ __forceinline__ void put_row_part(uint32_t stage, int r, int kq0, const float* v) {
#pragma unroll
  for (int kq = 0; kq < NCG; ++kq) {
    float h[4], l[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) umma::split_tf32_rn_exact(v[4 * kq + i], h[i], l[i]);
    const uint32_t a = stage + (kq0 + kq) * TU_CGA + r * 16;
    sts128(a, h[0], h[1], h[2], h[3]);
    sts128(a + TU_HALFA, l[0], l[1], l[2], l[3]);
  }
}
// one value of the transposed (K = rows) operand: operand row n, k = lane, column-group stride cgs
__device__ __forceinline__ void put_t(uint32_t stage, int half, int cgs, int n, int lane, float v) {
  float h, l;
  umma::split_tf32_rn_exact(v, h, l);
  const uint32_t a = stage + (n >> 3) * 128 + (lane >> 2) * cgs + (n & 7) * 16 + (lane & 3) * 4;
  sts32(a, h);
  sts32(a + half, l);
}
// The same with the lane- and column-base-dependent part of the address hoisted: tb = put_t_base(stage, cgs, n0, lane)
// for a column base n0 that is a multiple of 8; operand row n0 + j with a COMPILE-TIME j is then a store at an
// immediate offset (the address arithmetic of put_t was 11 % of all executed instructions of the 64-wide kernel)
__device__ __forceinline__ uint32_t put_t_base(uint32_t stage, int cgs, int n0, int lane) {
  return stage + (n0 >> 3) * 128 + (lane >> 2) * cgs + (lane & 3) * 4;
}
__device__ __forceinline__ void put_t_at(uint32_t tb_hi, uint32_t tb_lo, int j, float v) {
  float h, l;
  umma::split_tf32_rn_exact(v, h, l);
  const uint32_t off = (uint32_t)((j >> 3) * 128 + (j & 7) * 16);
  sts32(tb_hi + off, h);
  sts32(tb_lo + off, l);
}

// Optional clock64 timeline of CTA 0 (build with -DTBNN_TU_PROFILE; read with tbnn_tu_profile): (tag, clock) pairs of
// row-worker warp 0 (role 0) and of the MMA issuer (role 1).  tag = role << 28 | k << 20 | ti << 16 | phase.
#ifdef TBNN_TU_PROFILE
__device__ long long g_tu_prof[2][2][4096];
__device__ int g_tu_prof_n[2];
#define TU_MARK(role, k, ti, phase)                                                                         \
  do {                                                                                                      \
    if (blockIdx.x == 0 && g_tu_prof_n[role] < 4096) {                                                      \
      const int i_ = g_tu_prof_n[role]++;                                                                   \
      g_tu_prof[role][0][i_] = ((long long)(role) << 28) | ((long long)(k) << 20) | ((long long)(ti) << 16) | (phase); \
      g_tu_prof[role][1][i_] = clock64();                                                                   \
    }                                                                                                       \
  } while (0)
#else
#define TU_MARK(role, k, ti, phase) do { } while (0)
#endif

struct TuBars {
  uint64_t fullA[TU_NS_MAX], emptyA[TU_NS_MAX], fullB[TU_NS_MAX], emptyB[TU_NS_MAX], accfull[8], accfree[8];   // acc*: [tile slot][F/B, W]
};

// Tiles in flight per CTA.  A tile is a serial chain (F_0 -> epilogue -> F_1 -> ... -> B_1 -> epilogue -> W_0); with two
// tiles in flight the row workers run one tile's epilogue while the tensor core runs the other tile's GEMM.  Tensor
// memory holds two tiles' accumulators only for the 64-wide network.
constexpr int TU_NT64 = 3;                     // most tile slots of the 64-wide network (tensor memory: 3 x 144 columns)

// TPR = 2: 168 registers per thread (register allocation rounds the 10 warps up to 12; a launch with 200 is refused);
// TPR = 4: 112
template <int HW, int ACTK, int TPR, int NT>
__global__ void __launch_bounds__(tu_threads(TPR), 1)
k_train_umma(const __grid_constant__ ModelPlan mp, const __grid_constant__ TrainUmmaPlan tp, int C, int S,
             const float* __restrict__ theta_pad, const unsigned char* __restrict__ wimg,
             const float* __restrict__ X, const float* __restrict__ Y, long long N, float* __restrict__ partial,
             double* __restrict__ stat_part, float* __restrict__ scratch) {
  extern __shared__ __align__(128) unsigned char smraw[];
  __shared__ uint32_t tmem_slot;
  static_assert(NT >= 1 && NT <= (HW == 64 ? TU_NT64 : 2), "tiles in flight");
  // 128-wide network with two tiles in flight: tensor memory is 2 x (128 + 128) columns, so the weight-gradient GEMMs
  // have no room for the constant-one column -- the bias gradients are column sums on the CUDA cores instead
  constexpr bool BIASCS = HW == 128 && NT == 2;
  constexpr int TU_NS = tu_ns(HW);              // ring stages
  constexpr int nH = HW / 32;                   // chunks of a hidden-width contraction
  constexpr int RWT = 128 * TPR;                // row-worker threads
  constexpr int TU_MMA_WARP = 4 * TPR, TU_TMA_WARP = 4 * TPR + 1;
  constexpr int HH = HW / TPR;                  // columns per row worker (TPR threads share a row)
  constexpr int nP = HH / 16;                   // 16-column pieces per row worker
  static_assert(HH == 16 || HH == 32 || HH == 64, "columns per row worker");
  constexpr int NWH = BIASCS ? HW : HW + 16;    // N of a hidden weight-gradient GEMM: [A | 1 | pad] (BIASCS: [A])
  constexpr bool SLOPES = ACTK == ACT_SQPRELU;
  constexpr bool STACKQ = SLOPES && HW == 64;   // slope gradients ride in operand rows 64..127
  // warp index through a shuffle: the compiler then treats everything derived from it (roles, column groups, tensor-
  // memory columns) as warp-uniform -- uniform branches and registers instead of per-thread ones
  // (measured per width: C3 4.75 -> 4.30 ms; the 128-wide kernel lost 2 % with it and keeps the plain form)
  const int tid = threadIdx.x, warp = TU_SHFL128 || HW == 64 ? __shfl_sync(0xffffffffu, tid >> 5, 0) : (tid >> 5), lane = tid & 31;
  const int G = tp.G, D = mp.D, OUT = mp.OUT, hact = tp.act;
  TuBars* bars = reinterpret_cast<TuBars*>(smraw + tp.off_bar);
  float* par = reinterpret_cast<float*>(smraw + tp.off_par);
  const uint32_t ringA = smem_u32(smraw + tp.off_a), ringB = smem_u32(smraw + tp.off_b);
  const long long ntile = (N + 127) >> 7;
  // Tiles of work item s of a chain: [s * tbase + min(s, trem), +tbase + (s < trem)) with tbase = ntile / S, trem = ntile % S
  // computed ONCE -- a 64-bit division per item is a subroutine call whose result the compiler treats as thread-varying,
  // which drags the MMA issuer's ring counter and every descriptor derived from it off the uniform datapath.
  const long long tile_base = ntile / S;
  const int tile_rem = (int)(ntile - tile_base * S);
  const int nitem = C * S;

  if (tid == 0) {
    for (int i = 0; i < TU_NS; ++i) {
      mbar_init(&bars->fullA[i], 4 * TPR);
      mbar_init(&bars->emptyA[i], 1);
      mbar_init(&bars->fullB[i], TPR);          // a weight-gradient chunk is written by the TPR warps of one lane quarter
      mbar_init(&bars->emptyB[i], 1);
    }
    for (int i = 0; i < 8; ++i) {
      mbar_init(&bars->accfull[i], 1);
      mbar_init(&bars->accfree[i], 4 * TPR);
    }
    mbar_fence_init();
  }
  if (warp == TU_MMA_WARP) umma::tmem_alloc(&tmem_slot, 512);
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tbase = tmem_slot;
  // Tensor-memory columns of tile slot ti.  F / B accumulator "big" (F GEMMs: the hi*hi products), the small products of
  // the F GEMMs "small" -- two short accumulation chains instead of one long one, the accumulator rounds toward zero --
  // and the accumulator of the W GEMMs.
  // 64-wide network, NT tile slots of 144 columns: [small / W: 80][big: 64] -- the small-product accumulator is only
  // live during the forward pass of a tile and the weight-gradient accumulator only during its backward pass, so they
  // share columns (the hand-over is ordered by the row workers: W operands exist only after F_{G-1} was read, and the next
  // tile's first operand only after W_0 was drained).  128-wide network: [big 128][small 128][W 144].
  constexpr int SLOTW = HW == 64 ? 144 : 256, SMALLW = HW == 64 ? 80 : 128;   // columns of a tile slot / of its shared part
  auto col_big = [](int ti) -> uint32_t { return (uint32_t)(NT > 1 ? SLOTW * ti + SMALLW : 0); };
  auto col_small = [](int ti) -> uint32_t { return (uint32_t)(NT > 1 ? SLOTW * ti : 128); };
  auto col_w = [](int ti) -> uint32_t { return (uint32_t)(NT > 1 ? SLOTW * ti : 256); };
  // Segments of a tile, in order (k = 0 .. 2G): 0: X -> F_0 operands | 1..G-1: epilogue of F_{k-1} -> F_k operands |
  // G: epilogue of F_{G-1}, last block, likelihood, dZ_{G-1} -> B_{G-1}, W_{G-1} operands | G+j: epilogue of B_{G-j} ->
  // B_{G-1-j}, W_{G-1-j} operands | 2G: drain W_0.  The tiles in flight alternate segment by segment; every role walks
  // through (tile group, segment, tile slot) in the same order, so the operand rings see one fixed chunk sequence.

  if (warp == TU_TMA_WARP) {
    // ================================================================ TMA producer
    if (lane == 0) {
      uint32_t cc = 0;
      const uint32_t cbytes = 2u * tu_half(HW);
      // The four chunks of a weight-gradient GEMM are filled by the row workers, but this thread still has to SEE
      // each stage's release: a parity wait is only unambiguous while the waiter is at most one phase behind.
      auto skip_w = [&]() {
        for (int q = 0; q < 4; ++q, ++cc) {
          const uint32_t st = cc % TU_NS, use = cc / TU_NS;
          mbar_wait_parked(&bars->emptyB[st], (use & 1u) ^ 1u);
        }
      };
      auto load = [&](const unsigned char* src) {
        const uint32_t st = cc % TU_NS, use = cc / TU_NS;
        mbar_wait_parked(&bars->emptyB[st], (use & 1u) ^ 1u);
        mbar_expect_tx(&bars->fullB[st], cbytes);        // one arrival ...
        mbar_arrive_n(&bars->fullB[st], TPR - 1);        // ... and the rest of the count
        bulk_g2s(smraw + tp.off_b + st * tp.b_stage, src, cbytes, &bars->fullB[st]);
        ++cc;
      };
      for (int item = blockIdx.x; item < nitem; item += gridDim.x) {
        const int c = item / S, s = item - c * S;
        const long long t0 = s * tile_base + (s < tile_rem ? s : tile_rem), t1 = t0 + tile_base + (s < tile_rem ? 1 : 0);
        const unsigned char* img = wimg + (size_t)c * tp.wimg_chain;
        for (long long tb = t0; tb < t1; tb += NT) {
          const int ntl = NT == 1 ? 1 : (int)((t1 - tb) < NT ? (t1 - tb) : NT);
          for (int k = 0; k < 2 * G; ++k) {
            for (int ti = 0; ti < ntl; ++ti) {
              if (k < G) {                                   // F_k
                const int nch = k == 0 ? tp.nK0 : nH;
                for (int ch = 0; ch < nch; ++ch) load(img + tp.fimg[k] + (size_t)ch * cbytes);
              } else {                                       // B_l (l >= 1), W_l
                const int l = 2 * G - 1 - k;
                if (l >= 1)
                  for (int ch = 0; ch < nH; ++ch) load(img + tp.bimg[l] + (size_t)ch * cbytes);
                skip_w();
              }
            }
          }
        }
      }
    }
  } else if (warp == TU_MMA_WARP) {
    // ================================================================ MMA issuer
    if (lane == 0) {
      // accpar: bit bi = parity of the number of uses of accumulator region bi.  (A dynamically indexed counter array
      // lives in local memory; the spin wait on a parity loaded from there made the compiler treat everything after
      // it as thread-varying -- ring counter, descriptors -- and wrap every MMA in an R2UR waterfall.)
      uint32_t cc = 0, accpar = 0u;
      const uint32_t idF = umma::idesc_tf32(128, HW, false, false);
      const uint32_t idWh = umma::idesc_tf32(128, NWH, false, false);
      const uint32_t idW0 = umma::idesc_tf32(128, tp.N0w, false, false);
      // one GEMM of `nch` chunks (last chunk: ks_last k steps) of tile slot ti into accumulator region `acc` (0: F / B,
      // 1: W); split: the lo*hi and hi*lo products go to their own accumulator (forward GEMMs)
      int prof_k = 0;
      auto gemm = [&](int ti, int nch, int ks_last, uint32_t idesc, int cgsB, int acc, bool split) {
        const int bi = 2 * ti + acc;
        TU_MARK(1, prof_k, ti, 0 + 8 * acc);
        mbar_wait(&bars->accfree[bi], ((accpar >> bi) & 1u) ^ 1u);
        accpar ^= 1u << bi;
        umma::fence_after_sync();
        const uint32_t dbig = umma::tmem_addr(tbase, 0, acc == 0 ? col_big(ti) : col_w(ti));
        const uint32_t dsml = split ? umma::tmem_addr(tbase, 0, col_small(ti)) : dbig;
        const uint32_t halfB = 8u * cgsB;
        auto chunk = [&](int ch) {
          const uint32_t st = cc % TU_NS, use = cc / TU_NS;
          mbar_wait(&bars->fullA[st], use & 1u);         // one thread: spinning is cheap and wakes up at once
          mbar_wait(&bars->fullB[st], use & 1u);
          umma::fence_after_sync();
          if (ch == 0) TU_MARK(1, prof_k, ti, 1 + 8 * acc);
          const uint32_t a0 = ringA + st * TU_ASTAGE, b0 = ringB + st * tp.b_stage;
          const int ks = ch == nch - 1 ? ks_last : 4;
          // descriptors as 32-bit words (umma::mma_tf32_ss32: they stay in uniform registers); a k step advances the
          // start-address field (bytes >> 4) of the lo word by two column groups
          const uint32_t hiw = umma::desc_hi(128u);
          const uint32_t aH0 = umma::desc_lo(a0, TU_CGA), aL0 = umma::desc_lo(a0 + TU_HALFA, TU_CGA);
          const uint32_t bH0 = umma::desc_lo(b0, cgsB), bL0 = umma::desc_lo(b0 + halfB, cgsB);
          const uint32_t stepA = (uint32_t)((2 * TU_CGA) >> 4), stepB = (uint32_t)((2 * cgsB) >> 4);
          for (int k = 0; k < ks; ++k) {
            const uint32_t aH = aH0 + k * stepA, aL = aL0 + k * stepA, bH = bH0 + k * stepB, bL = bL0 + k * stepB;
            const bool first = ch == 0 && k == 0;
            umma::mma_tf32_ss32(dsml, aL, hiw, bH, hiw, idesc, !first);
            umma::mma_tf32_ss32(dsml, aH, hiw, bL, hiw, idesc, true);
            umma::mma_tf32_ss32(dbig, aH, hiw, bH, hiw, idesc, split ? !first : true);
          }
          umma::commit(&bars->emptyA[st]);
          umma::commit(&bars->emptyB[st]);
        };
        if (TU_ROLL_ALL || HW == 128) {
#pragma unroll 1
          for (int ch = 0; ch < nch; ++ch, ++cc) chunk(ch);
        } else {
          for (int ch = 0; ch < nch; ++ch, ++cc) chunk(ch);
        }
        umma::commit(&bars->accfull[bi]);
        TU_MARK(1, prof_k, ti, 2 + 8 * acc);
      };
      const int ks0 = (tp.K0p - 32 * (tp.nK0 - 1)) / 8;
      for (int item = blockIdx.x; item < nitem; item += gridDim.x) {
        const int s = item % S;
        const long long t0 = s * tile_base + (s < tile_rem ? s : tile_rem), t1 = t0 + tile_base + (s < tile_rem ? 1 : 0);
        for (long long tb = t0; tb < t1; tb += NT) {
          const int ntl = NT == 1 ? 1 : (int)((t1 - tb) < NT ? (t1 - tb) : NT);
          for (int k = 0; k < 2 * G; ++k) {
            for (int ti = 0; ti < ntl; ++ti) {
              prof_k = k;
              if (TU_ROLL_ALL || HW == 128) {
                // One call site per accumulator and a rolled chunk loop: the 128-wide kernel is held back by instruction
                // fetch (210 KB of SASS per variant, instruction-cache hit rate 81 %): 11.87 -> 11.40 ms at C4.  The
                // 64-wide kernel is bound by how fast this one thread issues: there the unrolled form below is 8 % faster
                // (5.58 vs 6.03 ms at C3), and running the issuer warp-wide with an elected lane is slower still
                // (profiles/r2h_summary.md).
                const int l = 2 * G - 1 - k;                 // backward segments: block whose dZ was just produced
                if (k < G || l >= 1) gemm(ti, k == 0 ? tp.nK0 : nH, k == 0 ? ks0 : 4, idF, tu_cgs(HW), 0, k < G);   // F_k or B_l
                if (k >= G) gemm(ti, 4, 4, l >= 1 ? idWh : idW0, l >= 1 ? tu_cgs(NWH) : tu_cgs(tp.N0w), 1, false);  // W_l
              } else if (k == 0) gemm(ti, tp.nK0, ks0, idF, tu_cgs(HW), 0, true);
              else if (k < G) gemm(ti, nH, 4, idF, tu_cgs(HW), 0, true);
              else {
                const int l = 2 * G - 1 - k;
                if (l >= 1) gemm(ti, nH, 4, idF, tu_cgs(HW), 0, false);                                  // B_l: dA_{l-1}
                gemm(ti, 4, 4, l >= 1 ? idWh : idW0, l >= 1 ? tu_cgs(NWH) : tu_cgs(tp.N0w), 1, false);    // W_l
              }
            }
          }
        }
      }
    }
  } else {
    // ================================================================ row workers: two threads per row, each
    // owning HH = HW / 2 of the columns (warps 0-3: columns [0, HH), warps 4-7: [HH, HW))
    const int grp = warp >> 2, wq = warp & 3;
    const int r = 32 * wq + lane;                        // row of the tile = TMEM lane
    const int cb = grp * HH;                             // first column of this thread
    const uint32_t lane_t = (uint32_t)(32 * wq) << 16;
    uint32_t cc = 0, accpar = 0u;                        // bit bi = parity of the uses of accumulator region bi
    float* bias_s = par + tp.par_bias;                   // [G][HW]
    float* slope_s = par + tp.par_slope;                 // [G][HW] effective negative-side slope
    float* sfac_s = par + tp.par_sraw;                   // [G][HW] d(effective slope)/d(parameter): 2 s or 1
    float* wl_s = par + tp.par_wl;                       // [OUT][HW], then bias [4]
    float* accl_s = par + tp.par_accl + wq * (OUT * HW + 4);   // per warp quarter [OUT][HW] + [4]: gradient of the last
                                                         // block (summed in fixed order at the end: reruns are bit-identical)
    float* accb_s = par + tp.par_accb + wq * (G * HW);   // BIASCS: per lane quarter [G][HW] bias-gradient column sums
    const BlockPlan& bL = mp.b[G];
    double stat = 0.0;
    float* gout = nullptr;

    auto wait_acc = [&](int ti, int a) {
      const int bi = 2 * ti + a;
      mbar_wait_parked(&bars->accfull[bi], (accpar >> bi) & 1u);
      accpar ^= 1u << bi;
      umma::fence_after_sync();
    };
    auto free_acc = [&](int ti, int a) {                  // this warp has read its share of that accumulator region
      umma::fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->accfree[2 * ti + a]);
    };
    // Ring A protocol: EVERY row-worker warp takes part in EVERY chunk, in order -- slot() waits for the slot's release,
    // the owner(s) write, done() arrives (8 arrivals complete a chunk).  A warp that skipped the chunks it does not
    // fill would fall two phases behind (or ahead of) a barrier, where a parity wait is ambiguous.  The release of a
    // ring-B slot is the same event (the MMA issuer commits both), so ring B needs no wait of its own here.
    auto slot = [&](uint32_t cq) -> uint32_t {
      const uint32_t st = cq % TU_NS, use = cq / TU_NS;
      mbar_wait_parked(&bars->emptyA[st], (use & 1u) ^ 1u);
      return st;
    };
    auto done = [&](uint32_t st, bool wrote) {
      if (wrote) fence_proxy_async();                    // only the warps that wrote operand bytes need the proxy fence
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->fullA[st]);
    };
    // K-major chunks of this thread's columns (values v[0..HH)) as the A operand of a hidden-width GEMM
    auto put_chunks = [&](const float* v) {
#pragma unroll
      for (int jj = 0; jj < nH; ++jj) {
        const uint32_t st = slot(cc + jj);
        bool own;
        if (HH >= 32) {
          own = jj / (HH / 32) == grp;
          if (own) put_row_part<8>(ringA + st * TU_ASTAGE, r, 0, v + (HH >= 32 ? 32 * (jj % (HH / 32)) : 0));
        } else {                                           // 16 columns: half a chunk
          own = jj == (grp >> 1);
          if (own) put_row_part<4>(ringA + st * TU_ASTAGE, r, (grp & 1) * 4, v);
        }
        done(st, own);
      }
      cc += nH;
    };
    // drain a weight-gradient GEMM of block `lb` (>= 1): operand row m = TMEM lane = output feature; the two column
    // groups split the input features, group 1 also takes the constant-one column (bias, slopes)
    auto drain_hidden = [&](int ti, int lb) {
      const BlockPlan& bp = mp.b[lb];
      wait_acc(ti, 1);
      if (r < HW) {
        float* gw = gout + bp.pw + r * bp.ld_in + cb;
#pragma unroll 1
        for (int n0 = 0; n0 < HH; n0 += 16) {
          float v[16];
          umma::tmem_ld16(tbase + lane_t + col_w(ti) + cb + n0, v);
          umma::tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; i += 4) red_add4(gw + n0 + i, v[i], v[i + 1], v[i + 2], v[i + 3]);
        }
      }
      if (!BIASCS && grp == TPR - 1) {
        float v[8];
        umma::tmem_ld8(tbase + lane_t + col_w(ti) + HW, v);
        umma::tmem_ld_wait();
        if (r < HW) red_add1(gout + bp.pb + r, v[0]);
        else if (STACKQ && act_has_slopes(bp.act)) red_add1(gout + bp.ps + (r - HW), v[0] * sfac_s[lb * HW + (r - HW)]);
      }
      free_acc(ti, 1);
    };
    // operands produced once dZ_l of a tile is known: B_l (l >= 1) and W_l
    auto bwd_produce = [&](int l, int ti, bool valid, const float* xrow, const float* dz, const float* qv) {
      float* scr = scratch + ((size_t)blockIdx.x * NT + ti) * tp.scratch_cta;
      if (l >= 1) put_chunks(dz);                          // dA_{l-1} = dZ_l W_l: critical path first
      if (tid == 0) TU_MARK(0, 60 + l, ti, 4);
      if (l < G - 1) drain_hidden(ti, l + 1);              // W_{l+1} of this tile, before its accumulator is reused
      if (tid == 0) TU_MARK(0, 60 + l, ti, 5);
#pragma unroll 1
      for (int q = 0; q < 4; ++q) {                        // W_l: chunk q = the 32 rows of lane quarter q
        const uint32_t st = slot(cc + q);
        if (q == wq) {
          const uint32_t sa = ringA + st * TU_ASTAGE, sb = ringB + st * tp.b_stage;
          const uint32_t ta = put_t_base(sa, TU_CGA, cb, lane);
#pragma unroll
          for (int j = 0; j < HH; ++j) put_t_at(ta, ta + TU_HALFA, j, dz[j]);
          if (HW == 64) {
#pragma unroll
            for (int j = 0; j < HH; ++j) put_t_at(ta + 8 * 128, ta + 8 * 128 + TU_HALFA, j, STACKQ ? qv[STACKQ ? j : 0] : 0.f);
          }
          if (l >= 1) {
            const int cgs = tu_cgs(NWH), half = tu_half(NWH);
            const float* sc = scr + (size_t)(l - 1) * HW * 128;
            const float* sl = slope_s + (l - 1) * HW + cb;
            const uint32_t tbb = put_t_base(sb, cgs, cb, lane);
            float4 kv[HH / 4];
#pragma unroll
            for (int g4 = 0; g4 < HH / 4; ++g4) kv[g4] = *reinterpret_cast<const float4*>(sc + ((size_t)(cb / 4 + g4) * 128 + r) * 4);
#pragma unroll
            for (int g4 = 0; g4 < HH / 4; ++g4) {
              const float k4[4] = {kv[g4].x, kv[g4].y, kv[g4].z, kv[g4].w};
#pragma unroll
              for (int i = 0; i < 4; ++i)
                put_t_at(tbb, tbb + half, 4 * g4 + i, tu_from_keep<ACTK>(hact, k4[i], SLOPES ? sl[4 * g4 + i] : 0.f));
            }
            if (!BIASCS && grp == TPR - 1) put_t(sb, half, cgs, HW, lane, 1.f);
          } else if (grp == 0) {
            const int cgs = tu_cgs(tp.N0w), half = tu_half(tp.N0w);
            for (int k = 0; k < D; ++k) put_t(sb, half, cgs, k, lane, valid ? xrow[k] : 0.f);
            put_t(sb, half, cgs, D, lane, 1.f);
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) mbar_arrive(&bars->fullB[st]);
        }
        done(st, false);                                   // the owner fenced above
      }
      cc += 4;
      if (BIASCS && l >= 1) {                              // bias gradient of block l: column sums of dZ_l over this warp's rows
#pragma unroll
        for (int g = 0; g < (HH >= 32 ? HH / 32 : 1); ++g) {
          float pr[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) pr[i] = dz[(HH >= 32 ? 32 * g : 0) + i];
          const float cs = colsum32(pr, lane);
          accb_s[l * HW + cb + 32 * g + lane] += cs;
        }
      }
    };

    for (int item = blockIdx.x; item < nitem; item += gridDim.x) {
      const int c = item / S, s = item - c * S;
      const long long t0 = s * tile_base + (s < tile_rem ? s : tile_rem), t1 = t0 + tile_base + (s < tile_rem ? 1 : 0);
      const float* th = theta_pad + (size_t)c * mp.Ppad;
      gout = partial + ((size_t)c * S + s) * mp.Ppad;
      ew_barrier<RWT>();                                      // previous item's parameters are dead
      for (int e = tid; e < G * HW; e += RWT) {
        const int l = e / HW, j = e - l * HW;
        const BlockPlan& b = mp.b[l];
        bias_s[e] = th[b.pb + j];
        float sl = 0.f, fac = 0.f;
        if (b.act == ACT_PRELU) { sl = th[b.ps + j]; fac = 1.f; }
        else if (b.act == ACT_SQPRELU) { const float t = th[b.ps + j]; sl = t * t; fac = 2.f * t; }
        else if (b.act == ACT_LEAKY) sl = (float)b.alpha;
        slope_s[e] = sl;
        sfac_s[e] = fac;
      }
      for (int e = tid; e < OUT * HW + 4; e += RWT) {
        float v = 0.f;
        if (e < OUT * HW) { const int o = e / HW, k = e - o * HW; v = th[bL.pw + o * bL.ld_in + k]; }
        else if (e - OUT * HW < OUT) v = th[bL.pb + e - OUT * HW];
        wl_s[e] = v;
#pragma unroll
        for (int w = 0; w < 4; ++w) par[tp.par_accl + w * (OUT * HW + 4) + e] = 0.f;
      }
      if (BIASCS)
        for (int e = tid; e < 4 * G * HW; e += RWT) par[tp.par_accb + e] = 0.f;
      for (int i = 4 * tid; i < mp.Ppad; i += 4 * RWT) *reinterpret_cast<float4*>(gout + i) = make_float4(0.f, 0.f, 0.f, 0.f);
      __threadfence();
      ew_barrier<RWT>();
      stat = 0.0;

      for (long long tb = t0; tb < t1; tb += NT) {
        const int ntl = NT == 1 ? 1 : (int)((t1 - tb) < NT ? (t1 - tb) : NT);
        for (int k = 0; k <= 2 * G; ++k) {
#pragma unroll 1
          for (int ti_ = 0; ti_ < ntl; ++ti_) {
            const int ti = NT == 1 ? 0 : ti_;              // a compile-time constant when one tile is in flight
            const long long row = (tb + ti) * 128 + r;
            const bool valid = row < N;
            const float* xrow = X + (valid ? row : 0) * (long long)D;
            float* scr = scratch + ((size_t)blockIdx.x * NT + ti) * tp.scratch_cta;
            if (tid == 0) TU_MARK(0, k, ti, 0);
            float dz[HH];                                     // values of this thread's columns (z / kept value / dZ)
            float qv[STACKQ ? HH : 1];
            int bl = -1;                                      // block whose dZ this segment produced (-1: none)
            if (k == 0) {
              // ------------------------------------------------ F_0 operand: this row of X (column group 0)
              for (int ch = 0; ch < tp.nK0; ++ch) {
                const uint32_t st = slot(cc + ch);
                if (grp == 0) {
                  const uint32_t stage = ringA + st * TU_ASTAGE;
                  const int ngr = min(8, (tp.K0p - 32 * ch) >> 2);
                  for (int kq = 0; kq < ngr; ++kq) {
                    float h[4], l4[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                      const int kk = 32 * ch + 4 * kq + i;
                      const float x = (valid && kk < D) ? xrow[kk] : 0.f;
                      umma::split_tf32_rn_exact(x, h[i], l4[i]);
                    }
                    const uint32_t a = stage + kq * TU_CGA + r * 16;
                    sts128(a, h[0], h[1], h[2], h[3]);
                    sts128(a + TU_HALFA, l4[0], l4[1], l4[2], l4[3]);
                  }
                }
                done(st, grp == 0);
              }
              cc += tp.nK0;
            } else if (k <= G) {
              // ------------------------------------------------ epilogue of F_l
              const int l = k - 1;
              float yv[4] = {0.f, 0.f, 0.f, 0.f};
              if (l == G - 1 && valid) {                       // targets requested before the accumulator wait
#pragma unroll
                for (int o = 0; o < 4; ++o)
                  if (o < OUT) yv[o] = Y[row * (long long)OUT + o];
              }
              wait_acc(ti, 0);
              if (tid == 0) TU_MARK(0, k, ti, 1);
#pragma unroll
              for (int j = 0; j < nP; ++j) {
                float vb[16], vs[16];
                umma::tmem_ld16(tbase + lane_t + col_big(ti) + cb + 16 * j, vb);
                umma::tmem_ld16(tbase + lane_t + col_small(ti) + cb + 16 * j, vs);
                umma::tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 16; ++i) dz[16 * j + i] = vb[i] + vs[i];
              }
              free_acc(ti, 0);
              const float* bz = bias_s + l * HW + cb;
              const float* sl = slope_s + l * HW + cb;
              if (l < G - 1) {
                float* sc = scr + (size_t)l * HW * 128;
#pragma unroll
                for (int q4 = 0; q4 < HH / 4; ++q4) {
                  float keep[4];
#pragma unroll
                  for (int i = 0; i < 4; ++i) {
                    const int col = 4 * q4 + i;
                    const float z = dz[col] + bz[col];
                    const float a = tu_act<ACTK>(hact, z, SLOPES ? sl[col] : 0.f);
                    keep[i] = tu_keep<ACTK>(hact, z, a);
                    dz[col] = a;
                  }
                  *reinterpret_cast<float4*>(sc + ((size_t)(cb / 4 + q4) * 128 + r) * 4) = make_float4(keep[0], keep[1], keep[2], keep[3]);
                }
                put_chunks(dz);
              } else {
                // ---------------------------------------------- last hidden block, last block, likelihood
                const int FXS = tp.fx_stride;
                float* fx_s = par + tp.par_fx + ti * (TPR * 128 * FXS);
                float f[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int col = 0; col < HH; ++col) {
                  const float z = dz[col] + bz[col];
                  const float a = tu_act<ACTK>(hact, z, SLOPES ? sl[col] : 0.f);
                  dz[col] = tu_keep<ACTK>(hact, z, a);
#pragma unroll
                  for (int o = 0; o < 4; ++o)
                    if (o < OUT) f[o] = fmaf(wl_s[o * HW + cb + col], a, f[o]);
                }
#pragma unroll
                for (int o = 0; o < 4; ++o)
                  if (o < OUT) fx_s[(grp * 128 + r) * FXS + o] = f[o];
                ew_barrier<RWT>();
                // fixed order (group 0 + group 1 + ...) so all threads of a row get identical bits
#pragma unroll
                for (int o = 0; o < 4; ++o) {
                  if (o < OUT) {
                    float acc = fx_s[r * FXS + o];
#pragma unroll
                    for (int g = 1; g < TPR; ++g) acc += fx_s[(g * 128 + r) * FXS + o];
                    f[o] = acc;
                  }
                }
                float dfl[4] = {0.f, 0.f, 0.f, 0.f};
                const float lo = 1e-8f, hi = (float)(1 - 1e-7);          // likelihood.py:229-230
#pragma unroll
                for (int o = 0; o < 4; ++o) {
                  if (o < OUT) {
                    const float fo = act_fwd<float>(bL.act, f[o] + wl_s[OUT * HW + o], 0.f);
                    if (valid) {
                      const float y = yv[o];
                      float df;
                      if (mp.lik == LIK_BERN) {
                        const float p = fo < lo ? lo : (fo > hi ? hi : fo);
                        if (grp == 0) stat += (double)((1.f - y) * log1pf(-p) + y * logf(p));
                        df = (fo < lo || fo > hi) ? 0.f : (y / p - (1.f - y) / (1.f - p));
                      } else {
                        const float res = y - fo;
                        if (grp == 0) stat += (double)res * (double)res;
                        df = res;
                      }
                      dfl[o] = df * act_deriv_from_out<float>(bL.act, fo);
                    }
                  }
                }
                // gradient of the last block (column sums over the tile's rows), dA_{G-1}, dZ_{G-1}
#pragma unroll
                for (int g = 0; g < (HH >= 32 ? HH / 32 : 1); ++g) {
#pragma unroll
                  for (int o = 0; o < 4; ++o) {
                    if (o < OUT) {
                      if (HH >= 32) {
                        float pr[32];
#pragma unroll
                        for (int i = 0; i < 32; ++i)
                          pr[i] = dfl[o] * tu_from_keep<ACTK>(hact, dz[(HH >= 32 ? 32 * g : 0) + i], SLOPES ? sl[(HH >= 32 ? 32 * g : 0) + i] : 0.f);
                        const float cs = colsum32(pr, lane);
                        accl_s[o * HW + cb + 32 * g + lane] += cs;
                      } else {
                        float pr[16];
#pragma unroll
                        for (int i = 0; i < 16; ++i) pr[i] = dfl[o] * tu_from_keep<ACTK>(hact, dz[i], SLOPES ? sl[i] : 0.f);
                        const float cs = colsum16(pr, lane);
                        if ((lane & 1) == 0) accl_s[o * HW + cb + colsum16_col(lane)] += cs;
                      }
                    }
                  }
                }
                if (grp == 0) {
                  float bsum[4];
#pragma unroll
                  for (int o = 0; o < 4; ++o) {
                    float v = dfl[o];
#pragma unroll
                    for (int sft = 16; sft > 0; sft >>= 1) v += __shfl_xor_sync(0xffffffffu, v, sft);
                    bsum[o] = v;
                  }
                  if (lane == 0)
                    for (int o = 0; o < OUT; ++o) accl_s[OUT * HW + o] += bsum[o];
                }
#pragma unroll
                for (int kk = 0; kk < HH; ++kk) {
                  float dA = 0.f;
#pragma unroll
                  for (int o = 0; o < 4; ++o)
                    if (o < OUT) dA = fmaf(dfl[o], wl_s[o * HW + cb + kk], dA);
                  const float sk = dz[kk];
                  if (STACKQ) qv[STACKQ ? kk : 0] = sk < 0.f ? sk * dA : 0.f;
                  dz[kk] = dA * tu_deriv<ACTK>(hact, sk, SLOPES ? sl[kk] : 0.f);
                }
                bl = G - 1;
              }
            } else if (k < 2 * G) {
              // ------------------------------------------------ epilogue of B_lp: dA_{lp-1} -> dZ_{lp-1}
              const int lp = 2 * G - k;
              const float* sc = scr + (size_t)(lp - 1) * HW * 128;
              const float* sl = slope_s + (lp - 1) * HW + cb;
              float4 kv[HH / 4];
#pragma unroll
              for (int g4 = 0; g4 < HH / 4; ++g4) kv[g4] = *reinterpret_cast<const float4*>(sc + ((size_t)(cb / 4 + g4) * 128 + r) * 4);
              wait_acc(ti, 0);
              if (tid == 0) TU_MARK(0, k, ti, 1);
#pragma unroll
              for (int j = 0; j < nP; ++j) {
                float v[16];
                tmem_ld16w(tbase + lane_t + col_big(ti) + cb + 16 * j, v);
#pragma unroll
                for (int i = 0; i < 16; ++i) dz[16 * j + i] = v[i];
              }
              free_acc(ti, 0);
#pragma unroll
              for (int g4 = 0; g4 < HH / 4; ++g4) {
                const float k4[4] = {kv[g4].x, kv[g4].y, kv[g4].z, kv[g4].w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  const int col = 4 * g4 + i;
                  const float dA = dz[col];
                  if (STACKQ) qv[STACKQ ? col : 0] = k4[i] < 0.f ? k4[i] * dA : 0.f;
                  dz[col] = dA * tu_deriv<ACTK>(hact, k4[i], SLOPES ? sl[col] : 0.f);
                }
              }
              bl = lp - 1;
            } else {
              // ------------------------------------------------ drain W_0
              const BlockPlan& b0 = mp.b[0];
              wait_acc(ti, 1);
              if (grp == 0) {
                if (r < HW) {
                  float* gw = gout + b0.pw + r * b0.ld_in;
                  for (int n0 = 0; n0 < tp.N0w; n0 += 8) {
                    float v[8];
                    umma::tmem_ld8(tbase + lane_t + col_w(ti) + n0, v);
                    umma::tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                      const int n = n0 + i;
                      if (n < D) red_add1(gw + n, v[i]);
                      else if (n == D) red_add1(gout + b0.pb + r, v[i]);
                    }
                  }
                } else if (STACKQ && act_has_slopes(b0.act)) {
                  float v[8];
                  umma::tmem_ld8(tbase + lane_t + col_w(ti) + (D & ~7), v);
                  umma::tmem_ld_wait();
                  float pick = 0.f;
#pragma unroll
                  for (int i = 0; i < 8; ++i)
                    if (i == (D & 7)) pick = v[i];
                  red_add1(gout + b0.ps + (r - HW), pick * sfac_s[r - HW]);
                }
              }
              free_acc(ti, 1);
            }
            if (bl >= 0) bwd_produce(bl, ti, valid, xrow, dz, qv);   // one inlined copy for both kinds of segment
            if (tid == 0) TU_MARK(0, k, ti, 3);
          }
        }
      }
      // ---------------------------------------------------- item epilogue: last block gradient, statistic
      ew_barrier<RWT>();
      {
        const float* a0 = par + tp.par_accl;
        const int ws = OUT * HW + 4;
        for (int e = tid; e < OUT * HW; e += RWT) {
          const int o = e / HW, k = e - o * HW;
          gout[bL.pw + o * bL.ld_in + k] = ((a0[e] + a0[ws + e]) + a0[2 * ws + e]) + a0[3 * ws + e];
        }
        if (tid < OUT) {
          const int e = OUT * HW + tid;
          gout[bL.pb + tid] = ((a0[e] + a0[ws + e]) + a0[2 * ws + e]) + a0[3 * ws + e];
        }
        if (BIASCS) {                                      // bias gradients of blocks 1 .. G-1 (fixed order over the lane quarters)
          const float* b0s = par + tp.par_accb;
          const int wsb = G * HW;
          for (int e = HW + tid; e < G * HW; e += RWT) {
            const int l = e / HW, j = e - l * HW;
            gout[mp.b[l].pb + j] = ((b0s[e] + b0s[wsb + e]) + b0s[2 * wsb + e]) + b0s[3 * wsb + e];
          }
        }
      }
      {
        double* red = reinterpret_cast<double*>(par + tp.par_accl + 4 * (OUT * HW + 4));
        const double w = warp_sum(stat);
        if (lane == 0 && grp == 0) red[wq] = w;
        ew_barrier<RWT>();
        if (tid == 0) stat_part[(size_t)c * S + s] = ((red[0] + red[1]) + red[2]) + red[3];
      }
    }
  }
  __syncwarp();
  umma::fence_before_sync();
  __syncthreads();
  if (warp == TU_MMA_WARP) umma::tmem_dealloc(tbase, 512);
}

// ------------------------------------------------------------------ host side
static inline int tu_pad(int x, int m) { return (x + m - 1) / m * m; }

bool plan_train_umma(const ModelPlan& mp, TrainUmmaPlan& tp, size_t smem_limit) {
  const int G = mp.nb - 1;
  if (G < 2 || G > MAXB - 1) return false;
  const int HW = mp.b[0].out;
  if (HW != 64 && HW != 128) return false;
  if (mp.OUT > 4 || mp.D > 128 || mp.D < 1) return false;
  const int act = mp.b[0].act;
  for (int l = 0; l < G; ++l) {
    if (mp.b[l].out != HW || mp.b[l].act != act) return false;
    if (l >= 1 && mp.b[l].in != HW) return false;
  }
  if (act_has_slopes(act) && HW != 64) return false;      // slope gradients ride in the free upper operand half
  const BlockPlan& bL = mp.b[G];
  if (bL.in != HW || act_keeps_z(bL.act)) return false;
  tp.G = G; tp.HW = HW; tp.act = act;
  tp.K0p = tu_pad(mp.D, 8);
  tp.nK0 = (tp.K0p + 31) / 32;
  tp.N0w = tu_pad(mp.D + 1, 16);
  if (HW == 64 && tp.N0w > 80) return false;              // the weight-gradient accumulator of a tile slot is 80 columns
  const int chunk = 2 * tu_half(HW);
  int cur = 0;
  for (int l = 0; l < G; ++l) { tp.fimg[l] = cur; cur += (l == 0 ? tp.nK0 : HW / 32) * chunk; }
  tp.bimg[0] = 0;
  for (int l = 1; l < G; ++l) { tp.bimg[l] = cur; cur += (HW / 32) * chunk; }
  tp.wimg_chain = cur;
  tp.a_stage = TU_ASTAGE;
  tp.b_stage = std::max(chunk, std::max(2 * tu_half(HW + 16), 2 * tu_half(tp.N0w)));
  tp.b_stage = tu_pad(tp.b_stage, 128);
  int off = 0;
  tp.off_a = off; off += tu_ns(HW) * tp.a_stage;
  off = tu_pad(off, 128);
  tp.off_b = off; off += tu_ns(HW) * tp.b_stage;
  tp.off_par = off;
  int pf = 0;
  tp.par_bias = pf; pf += G * HW;
  tp.par_slope = pf; pf += G * HW;
  tp.par_sraw = pf; pf += G * HW;
  tp.par_wl = pf; pf += mp.OUT * HW + 4;
  tp.par_accl = pf; pf += 4 * (mp.OUT * HW + 4) + 16;    // one per lane quarter, + 8 doubles of reduction scratch
  tp.par_accb = pf; pf += HW == 128 ? 4 * G * HW : 0;    // 128-wide, two tiles in flight: bias-gradient column sums per lane quarter
  // 64-wide network: four threads per row and three tile slots when the exchange buffer of the last block
  // ([tile slot][thread of the row][row][fx_stride] floats) still fits, else fewer; TBNN_TU_TPR=2 in the environment keeps
  // two threads per row (A/B measurements, tests).
  tp.fx_stride = mp.OUT == 1 ? 1 : (mp.OUT == 2 ? 2 : 4);
  tp.par_fx = pf;
  const char* env = getenv("TBNN_TU_TPR");
  const int want = env ? atoi(env) : 0;
  const int rest = tu_pad((int)sizeof(TuBars), 16) + 16;
  tp.TPR = 2;
  tp.NTmax = HW == 64 ? 2 : 1;
  if (HW == 64) {
    const int cand[4][2] = {{4, 3}, {4, 2}, {2, 3}, {2, 2}};
    for (auto& c : cand) {
      if (want == 2 && c[0] != 2) continue;
      if ((size_t)(off + (pf + c[1] * c[0] * 128 * tp.fx_stride) * 4 + rest) <= smem_limit) { tp.TPR = c[0]; tp.NTmax = c[1]; break; }
    }
  }
  if (HW == 128) {                                       // two tiles in flight when the second exchange buffer fits
    const char* e128 = getenv("TBNN_TU_NT128");          // TBNN_TU_NT128=1: one tile in flight (A/B measurements, tests)
    const int want128 = e128 ? atoi(e128) : 0;
    if (want128 != 1 && tp.N0w <= 128 && (size_t)(off + (pf + 2 * tp.TPR * 128 * tp.fx_stride) * 4 + rest) <= smem_limit) tp.NTmax = 2;
  }
  pf += tp.NTmax * tp.TPR * 128 * tp.fx_stride;
  off += pf * 4;
  off = tu_pad(off, 16);
  tp.off_bar = off; off += (int)sizeof(TuBars);
  tp.smem_bytes = off;
  tp.scratch_cta = (G - 1) * HW * 128;                   // per tile slot
  tp.NT = tp.NTmax;                                      // the caller may lower it to 2 (work-item planner, api.cu)
  return (size_t)off <= smem_limit;
}

#ifdef TBNN_TU_PROFILE
extern "C" int tbnn_tu_profile(long long* out, int cap) {
  // out: [role][tag | clock][cap]; returns entries of role 0 in the low 16 bits, role 1 in the high 16 bits
  static long long buf[2][2][4096];
  int n[2];
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(buf, g_tu_prof, sizeof(buf));
  cudaMemcpyFromSymbol(n, g_tu_prof_n, sizeof(n));
  for (int r = 0; r < 2; ++r)
    for (int q = 0; q < 2; ++q)
      for (int i = 0; i < cap && i < 4096; ++i) out[(r * 2 + q) * cap + i] = buf[r][q][i];
  int zero[2] = {0, 0};
  cudaMemcpyToSymbol(g_tu_prof_n, zero, sizeof(zero));
  return n[0] | (n[1] << 16);
}
#endif

int train_umma_tiles_in_flight(const TrainUmmaPlan& tp) { return tp.NTmax; }   // the most; tp.NT may be lower
size_t train_umma_wimg_bytes(const TrainUmmaPlan& tp, int C) { return (size_t)C * tp.wimg_chain; }
size_t train_umma_scratch_bytes(const TrainUmmaPlan& tp, int num_sms) { return (size_t)num_sms * 4 * tp.scratch_cta * 4; }   // up to four tile slots

void launch_train_umma(const ModelPlan& mp, const TrainUmmaPlan& tp, int num_sms, int C, int S, const float* theta_pad,
                       unsigned char* wimg, float* scratch, const float* X, const float* Y, long long N,
                       float* partial, double* stat_part, cudaStream_t st) {
  const int HW = tp.HW;
  const int groups = (tp.nK0 + (2 * tp.G - 2) * (HW / 32)) * HW * 8;
  k_train_prep<<<dim3((groups + 255) / 256, C), 256, 0, st>>>(mp, tp, theta_pad, wimg);
  const int grid = std::min(num_sms, C * S);
  const int actk = tp.act == ACT_RELU ? ACT_RELU : (act_keeps_z(tp.act) ? ACT_SQPRELU : -1);
#define TU_LAUNCH(HWV, AK, TP, NTV)                                                                              \
  do {                                                                                                           \
    cudaFuncSetAttribute(k_train_umma<HWV, AK, TP, NTV>, cudaFuncAttributeMaxDynamicSharedMemorySize,            \
                         tp.smem_bytes);                                                                         \
    k_train_umma<HWV, AK, TP, NTV><<<grid, tu_threads(TP), tp.smem_bytes, st>>>(                                 \
        mp, tp, C, S, theta_pad, wimg, X, Y, N, partial, stat_part, scratch);                                    \
  } while (0)
#define TU_LAUNCH_ACT(HWV, TP, NTV)                                                                              \
  do {                                                                                                           \
    if (actk == ACT_RELU) TU_LAUNCH(HWV, ACT_RELU, TP, NTV);                                                     \
    else if (actk == ACT_SQPRELU) TU_LAUNCH(HWV, ACT_SQPRELU, TP, NTV);                                          \
    else TU_LAUNCH(HWV, -1, TP, NTV);                                                                            \
  } while (0)
  if (HW == 64) {
    if (tp.TPR == 4 && tp.NT == 3) TU_LAUNCH_ACT(64, 4, 3);
    else if (tp.TPR == 4) TU_LAUNCH_ACT(64, 4, 2);
    else if (tp.NT == 3) TU_LAUNCH_ACT(64, 2, 3);
    else TU_LAUNCH_ACT(64, 2, 2);
  } else {
    if (tp.NT == 2) TU_LAUNCH_ACT(128, 2, 2);
    else TU_LAUNCH_ACT(128, 2, 1);
  }
#undef TU_LAUNCH_ACT
#undef TU_LAUNCH
}

}  // namespace tbnn
