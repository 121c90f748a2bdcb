// k_wide2.cu -- warp-specialised row sweep for a WIDE first dense layer (e.g. 784 -> 20, the docs
// ClassificationExample shape), fp32, forward + likelihood + backward in one pass over the rows.
//
// Same contract as k_partial / k_sweep_wide: one CTA = a contiguous block of training rows of one chain;
// output = this CTA's partial gradient (padded layout) + likelihood statistic.  It replaces, for those
// rows, network.predict (network.py:141-171), layer.predict (layer.py:266-279), the activations, the
// likelihood residuals (likelihood.py:88-94,162-167,225-236) and TF's reverse-mode autodiff of them.
//
// The phase-serial predecessor (k_wide.cu) is latency bound (profiles/r1b_summary.md).  Here the stages of
// a pass (<= 14 rows) run CONCURRENTLY on different passes, coupled only by mbarriers:
//   producer warp : TMA bulk copies (cp.async.bulk -> mbarrier) of X rows into a ring of three tiles;
//   F warps (3)   : block-0 forward z1 = X W1^T, split-K across warps and across 4 lane groups, register
//                   tile 4 rows x 2*NO outputs; packed fma.rn.f32x2 (SASS FFMA2) on OUTPUT pairs
//                   (w[o][k], w[o+1][k]) read from a pair-interleaved copy of W1 against a broadcast x;
//   T warps (2x2) : two groups of two warps; group g takes the passes p = g (mod 2), so two passes are inside the
//                   tail at any time (it is the latency-critical stage: one group alone held the sweep at 51.7 us,
//                   two give 45.4 us at C2).  Per pass: cross-warp reduction + bias + activation of block 0, the narrow tail (blocks >= 1,
//                   widths <= 32: lane = (row pair, output quad), 2 x 4 register tiles), likelihood, data
//                   gradient back to dz1 -- the latency-critical chain between F and B of a pass;
//   A warp        : weight / bias / slope gradients of everything except W1, from double-buffered batch
//                   buffers, one pass behind the T warps;
//   B warps (7)   : dW1[o][k] += dz1[r][o] X[r][k] with the accumulators in REGISTERS for the CTA's whole
//                   row range (thread = one 4-wide k chunk x all outputs, FFMA2 on output pairs).
// A tile stays in shared memory from its load until B is done with it, so X is read from L2/HBM exactly once.
#include "async.cuh"
#include "engine.cuh"
#include "kernels.h"
#include "narrow.cuh"

#include <cstdio>
#include <cstdlib>

namespace tbnn {

constexpr int W2_NF = 3;                 // forward warps
constexpr int W2_NB = 7;                 // backward warps (224 k-quads)
constexpr int W2_NT = 2;                 // tail warps per group (8 rows each)
constexpr int W2_NTG = 2;                // tail groups: group g takes the passes p = g (mod W2_NTG) -- the tail is the
                                         // latency-critical stage, two passes are in flight in it at any time
constexpr int W2_THREADS = 32 * (W2_NF + W2_NB + W2_NT * W2_NTG + 2);   // + accumulate warp + producer warp = 512
constexpr int W2_TROWS = 16;             // rows of the forward register tiling (4 row groups x 4)
constexpr int W2_NBUF = 3;               // X tiles in the ring
constexpr int W2_MAXNO = 5;              // block-0 outputs <= 20 (register budget of the F / B tiles)
constexpr int W2_MAXNB = 4;              // dense blocks (the layer loops are unrolled at compile time)

typedef unsigned long long u64;

__device__ __forceinline__ void ffma2(u64& acc, u64 a, u64 b) {
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(a), "l"(b));
}
__device__ __forceinline__ u64 pack2(float lo, float hi) {
  u64 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ float2 unpack2(u64 v) {
  float2 r;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
  return r;
}
__device__ __forceinline__ void t_barrier() { asm volatile("bar.sync 1, %0;\n" ::"n"(32 * W2_NT * W2_NTG) : "memory"); }
// waiting with back-off: the waiter is not on the critical path, leave the issue slots to the others
__device__ __forceinline__ void mbar_wait_relaxed(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) __nanosleep(64);
}

// `rows` rows split into `npass` balanced passes: pass p covers [lo(p), lo(p) + size(p))
struct PassSplit {
  int base, rem;
  __device__ __forceinline__ int lo(int p) const { return p * base + (p < rem ? p : rem); }
  __device__ __forceinline__ int size(int p) const { return base + (p < rem ? 1 : 0); }
};

// ---- quad-wide activation helpers (the activation kind is warp-uniform: one branch per quad)
__device__ __forceinline__ void slopes_quad(const BlockPlan& b, const float* Wt, int o0, float (&sl)[4]) {
  if (b.act == ACT_PRELU) {
    ld4(Wt + b.ps + o0, sl);
  } else if (b.act == ACT_SQPRELU) {
    ld4(Wt + b.ps + o0, sl);
#pragma unroll
    for (int e = 0; e < 4; ++e) sl[e] *= sl[e];
  } else {
#pragma unroll
    for (int e = 0; e < 4; ++e) sl[e] = (float)b.alpha;
  }
}
// a = act(z) for the valid outputs o0+e < out; a = z = 0 for padding (fwd_store semantics, engine.cuh)
__device__ __forceinline__ void act_quad(const BlockPlan& b, const float* Wt, int o0, float (&z)[4], float (&a)[4]) {
  switch (b.act) {
    case ACT_NONE:
#pragma unroll
      for (int e = 0; e < 4; ++e) a[e] = z[e];
      break;
    case ACT_RELU:
#pragma unroll
      for (int e = 0; e < 4; ++e) a[e] = z[e] > 0.f ? z[e] : 0.f;
      break;
    case ACT_TANH:
#pragma unroll
      for (int e = 0; e < 4; ++e) a[e] = tanhf(z[e]);
      break;
    case ACT_SIGMOID:
#pragma unroll
      for (int e = 0; e < 4; ++e) a[e] = 1.f / (1.f + expf(-z[e]));
      break;
    case ACT_EXP:
#pragma unroll
      for (int e = 0; e < 4; ++e) a[e] = expf(z[e]);
      break;
    case ACT_ELU:
#pragma unroll
      for (int e = 0; e < 4; ++e) a[e] = z[e] > 0.f ? z[e] : expm1f(z[e]);
      break;
    default: {
      float sl[4];
      slopes_quad(b, Wt, o0, sl);
#pragma unroll
      for (int e = 0; e < 4; ++e) a[e] = z[e] < 0.f ? sl[e] * z[e] : z[e];
    }
  }
#pragma unroll
  for (int e = 0; e < 4; ++e)
    if (o0 + e >= b.out) { a[e] = 0.f; z[e] = 0.f; }
}
// bias + activation + store of one output quad of one row
__device__ __forceinline__ void store_quad(const BlockPlan& b, const float* Wt, float* bs, int row, int o0,
                                           const float (&acc)[4]) {
  float bias[4], z[4], a[4];
  ld4(Wt + b.pb + o0, bias);
#pragma unroll
  for (int e = 0; e < 4; ++e) z[e] = acc[e] + bias[e];
  act_quad(b, Wt, o0, z, a);
  st4(bs + b.offS + row * b.ld_out + o0, a);
  if (b.offZ >= 0) st4(bs + b.offZ + row * b.ld_out + o0, z);
}
// dz[k] = da[k] * act'(block pb) for the quad k0..k0+3 of one row; leaves the slope contribution in Z
__device__ __forceinline__ void dact_store_quad(const BlockPlan& pb, const float* Wt, float* bs, int row, int k0,
                                                const float (&da)[4], float* dst) {
  float dz[4];
  if (act_keeps_z(pb.act)) {
    float zz[4], sl[4], cp[4];
    ld4(bs + pb.offZ + row * pb.ld_out + k0, zz);
    slopes_quad(pb, Wt, k0, sl);
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const bool neg = zz[t] < 0.f;
      dz[t] = neg ? da[t] * sl[t] : da[t];
      cp[t] = neg ? zz[t] * da[t] : 0.f;
    }
    if (act_has_slopes(pb.act)) {
#pragma unroll
      for (int t = 0; t < 4; ++t)
        if (k0 + t >= pb.out) cp[t] = 0.f;
      st4(bs + pb.offZ + row * pb.ld_out + k0, cp);
    }
  } else {
    float a[4];
    ld4(bs + pb.offS + row * pb.ld_out + k0, a);
#pragma unroll
    for (int t = 0; t < 4; ++t) dz[t] = da[t] * act_deriv_from_out<float>(pb.act, a[t]);
  }
#pragma unroll
  for (int t = 0; t < 4; ++t)
    if (k0 + t >= pb.out) dz[t] = 0.f;
  st4(dst, dz);
}

template <int NO, int NB>
__global__ void __launch_bounds__(W2_THREADS, 1)
k_sweep_wide2(const __grid_constant__ ModelPlan mp, int S, const float* __restrict__ theta_pad,
              const float* __restrict__ w1p, const float* __restrict__ X, const float* __restrict__ Y,
              long long N, float* __restrict__ partial, double* __restrict__ stat_part, long long* prof) {
  extern __shared__ __align__(16) unsigned char smraw[];
  float* sm = reinterpret_cast<float*>(smraw);
  // developer aid: clock64 marks of CTA 0, one row of 32 slots per role.  Compiled in only with -DTBNN_W2_PROFILE
  // (the marks cost ~2 us per sweep); then TBNN_W2_PROF=1 prints the timeline of the fourth launch to stderr.
#ifdef TBNN_W2_PROFILE
  if (blockIdx.x != 0 || blockIdx.y != 0) prof = nullptr;
#define W2_MARK(role, idx) do { if (prof && (idx) < 32) prof[(role) * 32 + (idx)] = clock64(); } while (0)
#else
#define W2_MARK(role, idx) do { } while (0)
#endif
  constexpr int OP = 4 * NO;               // padded outputs of block 0
  constexpr int HO = 2 * NO;               // outputs per forward half
  constexpr int QS = 4 * OP + 4;           // floats per k quad of the pair-interleaved W1 (w1p_quad_stride)
  const int c = blockIdx.y, s = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const BlockPlan& b0 = mp.b[0];
  const int ld0 = mp.ld0, D = mp.D, nch = mp.D_p >> 2, RB = mp.TR;
  float* W1s = sm + mp.offW;               // pair-interleaved W1: [k quad][output pair][k in quad][2]
  float* Wt = sm + mp.offDb - b0.pb;       // every other parameter, indexed like the padded theta (>= b0.pb)
  float* zs = sm + mp.offScr;              // [2][W2_NF][W2_TROWS][OP]
  float* dzring = sm + mp.offDa;           // [W2_NBUF][W2_TROWS][OP]: dz of block 0, one slot per X tile
  double* red = reinterpret_cast<double*>(sm + mp.offRed);
  uint64_t* bars = reinterpret_cast<uint64_t*>(red + 8);
  uint64_t* full = bars;                   // [3] X tile landed (tx bytes)
  uint64_t* empty = bars + 3;              // [3] B warps are done with the tile
  uint64_t* zready = bars + 6;             // [2] F warps stored their partial sums
  uint64_t* zfree = bars + 8;              // [2] T warps consumed them
  uint64_t* dzready = bars + 10;           // [3] T warps published dz1
  uint64_t* wbar = bars + 13;              // parameters landed
  uint64_t* tdone = bars + 14;             // [2] T warps finished a pass (batch buffer set p & 1 complete)
  uint64_t* afree = bars + 16;             // [2] A warp consumed the set
  const long long r_begin = N * s / S, r_end = N * (s + 1) / S;
  const int rows = (int)(r_end - r_begin);
  const int npass = (rows + RB - 1) / RB;
  const int setsz = mp.ldmax;              // floats per batch-buffer set (plan_wide2)
  PassSplit ps;
  ps.base = npass > 0 ? rows / npass : 0;
  ps.rem = npass > 0 ? rows - ps.base * npass : 0;

  if (tid == 0) {
    for (int i = 0; i < 3; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], W2_NB); mbar_init(&dzready[i], W2_NT); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&zready[i], W2_NF); mbar_init(&zfree[i], W2_NT);
      mbar_init(&tdone[i], W2_NT); mbar_init(&afree[i], 1);
    }
    mbar_init(wbar, 1);
    mbar_fence_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) W2_MARK(7, 0);

  if (warp == W2_NF + W2_NB + W2_NT * W2_NTG + 1) {
    // ================================================================= producer
    if (lane == 0) {
      const uint32_t b1 = (uint32_t)(nch * QS * 4), b2 = (uint32_t)((mp.Ppad - b0.pb) * 4);
      fence_proxy_async();
      mbar_expect_tx(wbar, b1 + b2);
      bulk_g2s(W1s, w1p + (size_t)c * nch * QS, b1, wbar);
      bulk_g2s(sm + mp.offDb, theta_pad + (size_t)c * mp.Ppad + b0.pb, b2, wbar);
    }
    for (int p = 0; p < npass; ++p) {
      const int b = p % W2_NBUF, n = p / W2_NBUF;
      if (n > 0) mbar_wait_relaxed(&empty[b], (uint32_t)((n - 1) & 1));
      const int lo = ps.lo(p), R = ps.size(p);
      float* dst = sm + mp.offX + b * RB * ld0;
      if (lane == 0) {
        fence_proxy_async();
        W2_MARK(0, p);
        mbar_expect_tx(&full[b], (uint32_t)(R * D * 4));
      }
      __syncwarp();
      if (lane < R)
        bulk_g2s(dst + lane * ld0, X + (r_begin + lo + lane) * (long long)D, (uint32_t)(D * 4), &full[b]);
    }
  } else if (warp < W2_NF) {
    // ================================================================= F: block-0 forward
    // lane = (k quad group kq, output half og, row group rg): rows 4rg..4rg+3, output pairs og*NO..og*NO+NO-1
    const int kq = lane >> 3, rg = lane & 3, og = (lane >> 2) & 1;
    const int nsteps = (nch + 3) >> 2;
    const float* w0 = W1s + og * (NO * 8);
    const bool bit0 = (kq & 1) != 0, bit1 = (kq & 2) != 0;
    const int row_out = rg * 4 + (bit0 ? 2 : 0) + (bit1 ? 1 : 0);
    mbar_wait(wbar, 0u);
    for (int p = 0; p < npass; ++p) {
      const int b = p % W2_NBUF;
      mbar_wait(&full[b], (uint32_t)((p / W2_NBUF) & 1));
      if (warp == 0 && lane == 0) W2_MARK(1, p);
      const float* Xs = sm + mp.offX + b * RB * ld0 + (rg * 4) * ld0;
      u64 acc[4][NO];                        // [row][output pair]: (z[2jp], z[2jp+1]) partial sums
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < NO; ++j) acc[i][j] = 0ull;
      for (int st = warp; st < nsteps; st += W2_NF) {
        const int ch = 4 * st + kq;
        if (ch < nch) {
          float4 xv[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) xv[i] = *reinterpret_cast<const float4*>(Xs + i * ld0 + 4 * ch);
          const float* wq = w0 + ch * QS;
#pragma unroll
          for (int j = 0; j < NO; ++j) {
            const ulonglong2 wa = *reinterpret_cast<const ulonglong2*>(wq + 8 * j);       // k0, k1
            const ulonglong2 wb = *reinterpret_cast<const ulonglong2*>(wq + 8 * j + 4);   // k2, k3
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              ffma2(acc[i][j], wa.x, pack2(xv[i].x, xv[i].x));
              ffma2(acc[i][j], wa.y, pack2(xv[i].y, xv[i].y));
              ffma2(acc[i][j], wb.x, pack2(xv[i].z, xv[i].z));
              ffma2(acc[i][j], wb.y, pack2(xv[i].w, xv[i].w));
            }
          }
        }
      }
      // transposing reduction over the 4 k-quad lane groups: this lane keeps row `row_out`
      float h1[2][HO];
#pragma unroll
      for (int i2 = 0; i2 < 2; ++i2)
#pragma unroll
        for (int j = 0; j < NO; ++j) {
          const float2 a2 = unpack2(acc[i2][j]), b2 = unpack2(acc[i2 + 2][j]);
          const float s0 = bit0 ? a2.x : b2.x, k0 = bit0 ? b2.x : a2.x;
          const float s1 = bit0 ? a2.y : b2.y, k1 = bit0 ? b2.y : a2.y;
          h1[i2][2 * j] = k0 + __shfl_xor_sync(0xffffffffu, s0, 8);
          h1[i2][2 * j + 1] = k1 + __shfl_xor_sync(0xffffffffu, s1, 8);
        }
      float v[HO];
#pragma unroll
      for (int j = 0; j < HO; ++j) {
        const float a = h1[0][j], bb = h1[1][j];
        const float send = bit1 ? a : bb, keep = bit1 ? bb : a;
        v[j] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
      }
      const int zb = p & 1;
      if (p >= 2) mbar_wait(&zfree[zb], (uint32_t)(((p >> 1) - 1) & 1));
      float* zo = zs + ((zb * W2_NF + warp) * W2_TROWS + row_out) * OP + og * HO;
#pragma unroll
      for (int j = 0; j < HO; j += 2) *reinterpret_cast<float2*>(zo + j) = make_float2(v[j], v[j + 1]);
      __syncwarp();
      if (lane == 0) mbar_arrive(&zready[zb]);
      if (warp == 0 && lane == 0) W2_MARK(2, p);
    }
  } else if (warp < W2_NF + W2_NB) {
    // ================================================================= B: dW1 in registers
    const int bt = tid - 32 * W2_NF;
    const bool has = bt < nch;
    u64 acc[OP / 2][4];                    // [output pair][k]: (dW1[2jo][k], dW1[2jo+1][k])
#pragma unroll
    for (int jo = 0; jo < OP / 2; ++jo)
#pragma unroll
      for (int k = 0; k < 4; ++k) acc[jo][k] = 0ull;
    for (int p = 0; p < npass; ++p) {
      const int b = p % W2_NBUF;
      const uint32_t par = (uint32_t)((p / W2_NBUF) & 1);
      mbar_wait_relaxed(&dzready[b], par);
      mbar_wait(&full[b], par);
      if (warp == W2_NF && lane == 0) W2_MARK(5, p);
      const int R = ps.size(p);
      if (has) {
        const float* xr = sm + mp.offX + b * RB * ld0 + 4 * bt;
        const float* dzr = dzring + b * W2_TROWS * OP;
#pragma unroll 2
        for (int r = 0; r < R; ++r) {
          const float4 x = *reinterpret_cast<const float4*>(xr + r * ld0);
          const u64 xd[4] = {pack2(x.x, x.x), pack2(x.y, x.y), pack2(x.z, x.z), pack2(x.w, x.w)};
#pragma unroll
          for (int jo = 0; jo < OP / 2; jo += 2) {
            const ulonglong2 dz = *reinterpret_cast<const ulonglong2*>(dzr + r * OP + 2 * jo);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              ffma2(acc[jo][k], dz.x, xd[k]);
              ffma2(acc[jo + 1][k], dz.y, xd[k]);
            }
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[b]);
      if (warp == W2_NF && lane == 0) W2_MARK(6, p);
    }
    float* out = partial + ((size_t)c * S + s) * mp.Ppad;
    if (bt < (ld0 >> 2)) {
#pragma unroll
      for (int o = 0; o < OP; ++o) {
        float v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float2 f2 = unpack2(acc[o >> 1][k]);
          v[k] = has ? ((o & 1) ? f2.y : f2.x) : 0.f;
        }
        st4(out + b0.pw + o * ld0 + 4 * bt, v);
      }
    }
  } else if (warp < W2_NF + W2_NB + W2_NT * W2_NTG) {
    // ================================================================= T: tail, likelihood, dz1
    // lane = (row pair rp, output quad oq): rows row0, row0+1 of the pass, outputs 4oq..4oq+3 of every block
    const int twa = warp - (W2_NF + W2_NB), tg = twa / W2_NT, tw = twa % W2_NT;
    const int rp = lane & 3, oq = lane >> 2, row0 = tw * 8 + 2 * rp, o0 = 4 * oq;
    const int OUT = mp.OUT;
    float stat = 0.f;
    const BlockPlan& bl = mp.b[NB - 1];
    mbar_wait(wbar, 0u);
    for (int p = tg; p < npass; p += W2_NTG) {
      const int b = p % W2_NBUF, zb = p & 1;
      const int lo = ps.lo(p), R = ps.size(p);
      const bool act[2] = {row0 < R, row0 + 1 < R};
      float* bs = sm + zb * setsz;           // this pass' batch-buffer set
      float* dz0 = dzring + (b * W2_TROWS + row0) * OP;
      // labels of the outputs this lane owns in the last block (global loads issued early)
      float yv[2][4];
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int e = 0; e < 4; ++e)
          yv[i][e] = (act[i] && o0 + e < OUT) ? Y[(r_begin + lo + row0 + i) * (long long)OUT + o0 + e] : 0.f;
      mbar_wait(&zready[zb], (uint32_t)((p >> 1) & 1));
      if (tw == 0 && lane == 0) W2_MARK(3, p);
      if (p >= 2) mbar_wait(&afree[zb], (uint32_t)(((p >> 1) - 1) & 1));
      // ---- block 0: sum of the F warps' partials, bias, activation
      if (oq < NO && act[0]) {
        float sacc[2][4];
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const float* zp = zs + ((zb * W2_NF) * W2_TROWS + row0 + i) * OP + o0;
          float t4[W2_NF][4];
#pragma unroll
          for (int w = 0; w < W2_NF; ++w) ld4(zp + w * W2_TROWS * OP, t4[w]);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            float sum = t4[0][e];
#pragma unroll
            for (int w = 1; w < W2_NF; ++w) sum += t4[w][e];
            sacc[i][e] = sum;
          }
        }
        store_quad(b0, Wt, bs, row0, o0, sacc[0]);
        if (act[1]) store_quad(b0, Wt, bs, row0 + 1, o0, sacc[1]);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&zfree[zb]);
      // ---- forward through blocks 1..NB-1 (layer.py:276-279 + activation)
#pragma unroll
      for (int l = 1; l < NB; ++l) {
        const BlockPlan& bk = mp.b[l];
        if (act[0] && oq < (bk.out_p >> 2)) {
          const float* ap = bs + mp.b[l - 1].offS + row0 * bk.ld_in;
          const float* wq = Wt + bk.pw + o0 * bk.ld_in;
          const int kch = bk.in_p >> 2;
          float a4[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
#pragma unroll 2
          for (int kc = 0; kc < kch; ++kc) {
            float av[2][4], wv[4][4];
            ld4(ap + 4 * kc, av[0]);
            ld4(ap + bk.ld_in + 4 * kc, av[1]);
#pragma unroll
            for (int e = 0; e < 4; ++e) ld4(wq + e * bk.ld_in + 4 * kc, wv[e]);
#pragma unroll
            for (int t = 0; t < 4; ++t)
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                a4[0][e] = fmaf(av[0][t], wv[e][t], a4[0][e]);
                a4[1][e] = fmaf(av[1][t], wv[e][t], a4[1][e]);
              }
          }
          store_quad(bk, Wt, bs, row0, o0, a4[0]);
          if (act[1]) store_quad(bk, Wt, bs, row0 + 1, o0, a4[1]);
        }
        __syncwarp();
      }
      // ---- likelihood residual -> dz of the last block (same arithmetic as lik_phase, engine.cuh)
      if (oq < (bl.out_p >> 2)) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          if (!act[i]) continue;
          const int row = row0 + i;
          const float* Sl = bs + bl.offS + row * bl.ld_out;
          float* Zl = bl.offZ >= 0 ? bs + bl.offZ + row * bl.ld_out : nullptr;
          float* Dl = NB == 1 ? dz0 + i * OP : bs + bl.offD + row * bl.ld_out;
          float dzq[4], ccq[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int o = o0 + e;
            float dz = 0.f, cc = 0.f;
            if (o < OUT) {
              const float f = Sl[o], y = yv[i][e];
              float df;
              if (mp.lik == LIK_BERN) {
                const float lo_ = 1e-8f, hi_ = (float)(1 - 1e-7);
                const float pp = f < lo_ ? lo_ : (f > hi_ ? hi_ : f);
                stat += (1.f - y) * log1pf(-pp) + y * logf(pp);
                df = (f < lo_ || f > hi_) ? 0.f : (y / pp - (1.f - y) / (1.f - pp));
              } else {
                const float res = y - f;
                stat = fmaf(res, res, stat);
                df = res;
              }
              if (act_keeps_z(bl.act)) {
                const float z = Zl[o];
                const bool neg = z < 0.f;
                const float sl = eff_slope<float>(bl.act, Wt + (bl.ps >= 0 ? bl.ps : b0.pb), o, (float)bl.alpha);
                dz = neg ? df * sl : df;
                cc = neg ? z * df : 0.f;
              } else {
                dz = df * act_deriv_from_out<float>(bl.act, f);
              }
            }
            dzq[e] = dz;
            ccq[e] = cc;
          }
          st4(Dl + o0, dzq);
          if (act_has_slopes(bl.act)) st4(Zl + o0, ccq);
        }
      }
      __syncwarp();
      // ---- data gradient: dz_{l-1}[k] = (sum_o dz_l[o] W_l[o][k]) * act'_{l-1}
#pragma unroll
      for (int l = NB - 1; l >= 1; --l) {
        const BlockPlan& bk = mp.b[l];
        const BlockPlan& pb = mp.b[l - 1];
        if (act[0] && oq < (bk.in_p >> 2)) {
          const int och = bk.out_p >> 2, ld = bk.ld_in;
          const float* dzr = bs + bk.offD + row0 * bk.ld_out;
          const float* wk = Wt + bk.pw + o0;
          float da[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
#pragma unroll 2
          for (int oc = 0; oc < och; ++oc) {
            float dv[2][4], wv[4][4];
            ld4(dzr + 4 * oc, dv[0]);
            ld4(dzr + bk.ld_out + 4 * oc, dv[1]);
#pragma unroll
            for (int e = 0; e < 4; ++e) ld4(wk + (4 * oc + e) * ld, wv[e]);
#pragma unroll
            for (int e = 0; e < 4; ++e)
#pragma unroll
              for (int t = 0; t < 4; ++t) {
                da[0][t] = fmaf(dv[0][e], wv[e][t], da[0][t]);
                da[1][t] = fmaf(dv[1][e], wv[e][t], da[1][t]);
              }
          }
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            if (!act[i]) continue;
            float* dst = l == 1 ? dz0 + i * OP + o0 : bs + pb.offD + (row0 + i) * pb.ld_out + o0;
            dact_store_quad(pb, Wt, bs, row0 + i, o0, da[i], dst);
          }
        }
        __syncwarp();
      }
      if (lane == 0) {
        mbar_arrive(&dzready[b]);
        if (tw == 0) W2_MARK(4, p);
        mbar_arrive(&tdone[zb]);
      }
    }
    const double ws = warp_sum((double)stat);
    if (lane == 0) red[twa] = ws;
    t_barrier();
    if (twa == 0 && lane == 0) {
      double tot = 0.0;
      for (int w = 0; w < W2_NT * W2_NTG; ++w) tot += red[w];
      stat_part[(size_t)c * S + s] = tot;
    }
  } else {
    // ================================================================= A: gradients of everything except W1
    float* G = sm + mp.offG - b0.pb;
    const int ldz0 = b0.ld_out;
    for (int i = b0.pb + lane; i < mp.Ppad; i += 32) G[i] = 0.f;
    mbar_wait(wbar, 0u);
    __syncwarp();
    for (int p = 0; p < npass; ++p) {
      const int b = p % W2_NBUF, zb = p & 1;
      const int R = ps.size(p);
      const float* bs = sm + zb * setsz;
      mbar_wait_relaxed(&tdone[zb], (uint32_t)((p >> 1) & 1));
      if (NB > 1) narrow_accum<float>(mp, Wt, G, bs, R, lane, 32);
      for (int o = lane; o < b0.out_p; o += 32) {
        float sb = 0.f;
        for (int r = 0; r < R; ++r) sb += dzring[(b * W2_TROWS + r) * OP + o];
        G[b0.pb + o] += sb;
        if (act_has_slopes(b0.act)) {
          const float* Z0 = bs + b0.offZ;
          float sc = 0.f;
          for (int r = 0; r < R; ++r) sc += Z0[r * ldz0 + o];
          const float f = b0.act == ACT_SQPRELU ? 2.f * Wt[b0.ps + o] : 1.f;
          G[b0.ps + o] += f * sc;
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&afree[zb]);
    }
    float* out = partial + ((size_t)c * S + s) * mp.Ppad;
    for (int i = b0.pb + lane; i < mp.Ppad; i += 32) out[i] = G[i];
    if (lane == 0) W2_MARK(7, 1);
  }
#undef W2_MARK
}

// ------------------------------------------------------------------ host side
bool wide2_supported(const ModelPlan& mp) {
  const BlockPlan& b0 = mp.b[0];
  for (int l = 1; l < mp.nb; ++l)
    if (mp.b[l].in_p > 32 || mp.b[l].out_p > 32) return false;
  return mp.D % 4 == 0 && mp.D_p >= 64 && (mp.ld0 >> 2) <= 32 * W2_NB && (b0.out_p >> 2) <= W2_MAXNO &&
         mp.OUT <= 32 && mp.nb <= W2_MAXNB;
}

// Shared-memory plan (offsets in floats): wp.TR = rows per pass (as many as fit, <= 16), wp.offW = the
// pair-interleaved W1, wp.offDb = the other parameters (padded theta from b[0].pb on), wp.ldmax = size of one
// batch-buffer set (S_l / Z_l / D_l of every block for 16 rows; two sets), wp.offDa = dz ring of block 0.
bool plan_wide2(const ModelPlan& mp, ModelPlan& wp, size_t smem_limit) {
  if (!wide2_supported(mp)) return false;
  for (int RB = W2_TROWS; RB >= 8; --RB) {
    wp = mp;
    wp.TR = RB;
    const int OP = wp.b[0].out_p;
    int cur = 0;
    wp.offW = cur; cur += w1p_elems(wp);
    wp.offDb = cur; cur += wp.Ppad - wp.b[0].pb;
    wp.offX = cur; cur += W2_NBUF * RB * wp.ld0;
    const int after_x = cur;
    wp.offScr = cur; cur += 2 * W2_NF * W2_TROWS * OP;
    wp.offDa = cur; cur += W2_NBUF * W2_TROWS * OP;
    const int set0 = cur;
    for (int l = 0; l < wp.nb; ++l) {
      BlockPlan& b = wp.b[l];
      b.ksplit = 1;
      b.offS = cur; cur += W2_TROWS * b.ld_out;
      if (act_keeps_z(b.act)) { b.offZ = cur; cur += W2_TROWS * b.ld_out; } else b.offZ = -1;
      if (l >= 1) { b.offD = cur; cur += W2_TROWS * b.ld_out; } else b.offD = -1;
    }
    wp.ldmax = cur - set0;            // second set follows the first
    cur += wp.ldmax;
    wp.offG = cur; cur += wp.Ppad - wp.b[0].pb;
    cur = (cur + 3) / 4 * 4;
    wp.offRed = cur; cur += 2 * 8 + 2 * 20;     // 8 doubles + 18 mbarriers (20 reserved)
    wp.smem_elems = cur;
    // rows RB..15 of the forward register tiling read past the last tile: those reads must stay inside
    // the allocation (they are never used)
    if (cur - after_x < (W2_TROWS - RB) * wp.ld0) continue;
    if ((size_t)cur * 4 <= smem_limit) return true;
  }
  return false;
}

template <int NO, int NB>
static void launch_w2(const ModelPlan& wp, dim3 g, size_t smem, const float* theta_pad, const float* w1p,
                      const float* X, const float* Y, long long N, float* partial, double* stat_part,
                      cudaStream_t st) {
  cudaFuncSetAttribute(k_sweep_wide2<NO, NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  static const bool want_prof = getenv("TBNN_W2_PROF") != nullptr;
  static long long* dprof = nullptr;
  if (want_prof && !dprof) cudaMalloc(&dprof, 8 * 32 * sizeof(long long));
  if (want_prof) cudaMemsetAsync(dprof, 0, 8 * 32 * sizeof(long long), st);
  k_sweep_wide2<NO, NB><<<g, W2_THREADS, smem, st>>>(wp, (int)g.x, theta_pad, w1p, X, Y, N, partial, stat_part,
                                                    want_prof ? dprof : nullptr);
  if (want_prof) {
    static int shown = 0;
    long long h[8 * 32];
    cudaStreamSynchronize(st);
    cudaMemcpy(h, dprof, sizeof(h), cudaMemcpyDeviceToHost);
    if (shown++ == 3) {
      const char* names[8] = {"tma issue", "F tile landed", "F published", "T start", "T done", "B start", "B done", "begin/end A"};
      const long long t0 = h[7 * 32];
      for (int r = 0; r < 8; ++r) {
        fprintf(stderr, "[w2_prof] %-14s", names[r]);
        for (int i = 0; i < 32; ++i) if (h[r * 32 + i]) fprintf(stderr, " %d:%lld", i, h[r * 32 + i] - t0);
        fprintf(stderr, "\n");
      }
    }
  }
}

template <int NO>
static void launch_no2(const ModelPlan& wp, dim3 g, size_t smem, const float* theta_pad, const float* w1p,
                       const float* X, const float* Y, long long N, float* partial, double* stat_part,
                       cudaStream_t st) {
  switch (wp.nb) {
    case 1: launch_w2<NO, 1>(wp, g, smem, theta_pad, w1p, X, Y, N, partial, stat_part, st); break;
    case 2: launch_w2<NO, 2>(wp, g, smem, theta_pad, w1p, X, Y, N, partial, stat_part, st); break;
    case 3: launch_w2<NO, 3>(wp, g, smem, theta_pad, w1p, X, Y, N, partial, stat_part, st); break;
    default: launch_w2<NO, 4>(wp, g, smem, theta_pad, w1p, X, Y, N, partial, stat_part, st); break;
  }
}

void launch_sweep_wide2(const ModelPlan& wp, int C, int S, const float* theta_pad, const float* w1p,
                        const float* X, const float* Y, long long N, float* partial, double* stat_part,
                        cudaStream_t st) {
  dim3 g(S, C);
  const size_t smem = (size_t)wp.smem_elems * sizeof(float);
  switch (wp.b[0].out_p >> 2) {
    case 1: launch_no2<1>(wp, g, smem, theta_pad, w1p, X, Y, N, partial, stat_part, st); break;
    case 2: launch_no2<2>(wp, g, smem, theta_pad, w1p, X, Y, N, partial, stat_part, st); break;
    case 3: launch_no2<3>(wp, g, smem, theta_pad, w1p, X, Y, N, partial, stat_part, st); break;
    case 4: launch_no2<4>(wp, g, smem, theta_pad, w1p, X, Y, N, partial, stat_part, st); break;
    default: launch_no2<5>(wp, g, smem, theta_pad, w1p, X, Y, N, partial, stat_part, st); break;
  }
}

}  // namespace tbnn
