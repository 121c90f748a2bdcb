// k_wide2.cu -- warp-specialised row sweep for a WIDE first dense layer (e.g. 784 -> 20, the docs
// ClassificationExample shape), fp32, forward + likelihood + backward in one pass over the rows.
//
// Same contract as k_partial / k_sweep_wide: one CTA = a contiguous block of training rows of one chain;
// output = this CTA's partial gradient (padded layout) + likelihood statistic.  It replaces, for those
// rows, network.predict (network.py:141-171), layer.predict (layer.py:266-279), the activations, the
// likelihood residuals (likelihood.py:88-94,162-167,225-236) and TF's reverse-mode autodiff of them.
//
// The phase-serial predecessor (k_wide.cu) is latency bound (profiles/r1b_summary.md).  Here the four
// stages of a pass (<= 14 rows) run CONCURRENTLY on different passes, coupled only by mbarriers:
//   producer warp : TMA bulk copies (cp.async.bulk -> mbarrier) of X rows into a ring of three tiles;
//   F warps (6)   : block-0 forward z1 = X W1^T, split-K across warps and across 4 lane groups, register
//                   tile 4 rows x 2*NO outputs, packed fma.rn.f32x2 (SASS FFMA2) on (even k, odd k) pairs
//                   that come straight out of the 128-bit shared-memory loads; partial sums to scratch;
//   T warps (2)   : cross-warp reduction + bias + activation of block 0, the narrow tail (blocks >= 1,
//                   widths <= 32: lane = (row % 8, output quad)), likelihood, data gradient back to dz1,
//                   then the weight / bias / slope gradients of everything except W1;
//   B warps (7)   : dW1[o][k] += dz1[r][o] X[r][k] with the accumulators in REGISTERS for the CTA's whole
//                   row range (thread = one 4-wide k chunk x all outputs, FFMA2 on output pairs).
// A tile stays in shared memory from its load until B is done with it, so X is read from L2/HBM exactly once.
#include "async.cuh"
#include "engine.cuh"
#include "kernels.h"
#include "narrow.cuh"

namespace tbnn {

constexpr int W2_NF = 6;                 // forward warps
constexpr int W2_NB = 7;                 // backward warps (224 k-quads)
constexpr int W2_NT = 2;                 // tail warps (8 rows each)
constexpr int W2_THREADS = 32 * (W2_NF + W2_NB + W2_NT + 1);   // + producer warp = 512
constexpr int W2_TROWS = 16;             // rows of the forward register tiling (4 row groups x 4)
constexpr int W2_NBUF = 3;               // X tiles in the ring
constexpr int W2_MAXNO = 5;              // block-0 outputs <= 20 (register budget of the F / B tiles)

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
__device__ __forceinline__ void t_barrier() { asm volatile("bar.sync 1, %0;\n" ::"n"(32 * W2_NT) : "memory"); }

// rows [lo, hi) of pass p when `rows` rows are split into `npass` balanced passes
__device__ __forceinline__ int pass_lo(int rows, int npass, int p) { return (int)(((long long)rows * p) / npass); }

template <int NO>
__global__ void __launch_bounds__(W2_THREADS, 1)
k_sweep_wide2(const __grid_constant__ ModelPlan mp, int S, const float* __restrict__ theta_pad,
              const float* __restrict__ X, const float* __restrict__ Y, long long N,
              float* __restrict__ partial, double* __restrict__ stat_part) {
  extern __shared__ __align__(16) unsigned char smraw[];
  float* sm = reinterpret_cast<float*>(smraw);
  constexpr int OP = 4 * NO;               // padded outputs of block 0
  constexpr int HO = 2 * NO;               // outputs per forward half
  const int c = blockIdx.y, s = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const BlockPlan& b0 = mp.b[0];
  const int ld0 = mp.ld0, D = mp.D, nch = mp.D_p >> 2, RB = mp.TR;
  const float* thg = theta_pad + (size_t)c * mp.Ppad;
  float* Ws = sm + mp.offW;
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
  const long long r_begin = N * s / S, r_end = N * (s + 1) / S;
  const int rows = (int)(r_end - r_begin);
  const int npass = (rows + RB - 1) / RB;

  if (tid == 0) {
    for (int i = 0; i < 3; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], W2_NB); mbar_init(&dzready[i], W2_NT); }
    for (int i = 0; i < 2; ++i) { mbar_init(&zready[i], W2_NF); mbar_init(&zfree[i], W2_NT); }
    mbar_init(wbar, 1);
    mbar_fence_init();
  }
  __syncthreads();

  if (warp == W2_NF + W2_NB + W2_NT) {
    // ================================================================= producer
    if (lane == 0) {
      fence_proxy_async();
      mbar_expect_tx(wbar, (uint32_t)(mp.Ppad * 4));
      bulk_g2s(Ws, thg, (uint32_t)(mp.Ppad * 4), wbar);
    }
    for (int p = 0; p < npass; ++p) {
      const int b = p % W2_NBUF, n = p / W2_NBUF;
      if (n > 0) mbar_wait(&empty[b], (uint32_t)((n - 1) & 1));
      const int lo = pass_lo(rows, npass, p), R = pass_lo(rows, npass, p + 1) - lo;
      float* dst = sm + mp.offX + b * RB * ld0;
      if (lane == 0) {
        fence_proxy_async();
        mbar_expect_tx(&full[b], (uint32_t)(R * D * 4));
      }
      __syncwarp();
      if (lane < R)
        bulk_g2s(dst + lane * ld0, X + (r_begin + lo + lane) * (long long)D, (uint32_t)(D * 4), &full[b]);
    }
  } else if (warp < W2_NF) {
    // ================================================================= F: block-0 forward
    const int kq = lane >> 3, rg = lane & 3, og = (lane >> 2) & 1;
    const int nsteps = (nch + 3) >> 2;
    const float* w0 = Ws + b0.pw + (og * HO) * ld0;
    const bool bit0 = (kq & 1) != 0, bit1 = (kq & 2) != 0;
    const int row_out = rg * 4 + (bit0 ? 2 : 0) + (bit1 ? 1 : 0);
    mbar_wait(wbar, 0u);
    for (int p = 0; p < npass; ++p) {
      const int b = p % W2_NBUF;
      mbar_wait(&full[b], (uint32_t)((p / W2_NBUF) & 1));
      const float* Xs = sm + mp.offX + b * RB * ld0 + (rg * 4) * ld0;
      u64 acc[4][HO];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < HO; ++j) acc[i][j] = 0ull;
      for (int st = warp; st < nsteps; st += W2_NF) {
        const int ch = 4 * st + kq;
        if (ch < nch) {
          ulonglong2 xv[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) xv[i] = *reinterpret_cast<const ulonglong2*>(Xs + i * ld0 + 4 * ch);
#pragma unroll
          for (int j = 0; j < HO; ++j) {
            const ulonglong2 wv = *reinterpret_cast<const ulonglong2*>(w0 + j * ld0 + 4 * ch);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              ffma2(acc[i][j], xv[i].x, wv.x);
              ffma2(acc[i][j], xv[i].y, wv.y);
            }
          }
        }
      }
      // (even k) + (odd k), then the transposing reduction over the 4 k-quad lane groups
      float h1[2][HO];
#pragma unroll
      for (int i2 = 0; i2 < 2; ++i2)
#pragma unroll
        for (int j = 0; j < HO; ++j) {
          const float2 a2 = unpack2(acc[i2][j]), b2 = unpack2(acc[i2 + 2][j]);
          const float a = a2.x + a2.y, bb = b2.x + b2.y;
          const float send = bit0 ? a : bb, keep = bit0 ? bb : a;
          h1[i2][j] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
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
      mbar_wait(&dzready[b], par);
      mbar_wait(&full[b], par);
      const int R = pass_lo(rows, npass, p + 1) - pass_lo(rows, npass, p);
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
  } else {
    // ================================================================= T: tail, likelihood, small gradients
    const int tw = warp - (W2_NF + W2_NB), ttid = tid - 32 * (W2_NF + W2_NB);
    const int r8 = lane & 7, oq = lane >> 3, row = tw * 8 + r8;
    const int nb = mp.nb, OUT = mp.OUT;
    float* G = sm + mp.offG - b0.pb;       // accumulators of everything except W1
    float* S0 = sm + b0.offS;
    float* Z0 = b0.offZ >= 0 ? sm + b0.offZ : nullptr;
    const int ldz0 = b0.ld_out;
    for (int i = b0.pb + ttid; i < mp.Ppad; i += 32 * W2_NT) G[i] = 0.f;
    mbar_wait(wbar, 0u);
    t_barrier();
    float stat = 0.f;
    const BlockPlan& bl = mp.b[nb - 1];
    for (int p = 0; p < npass; ++p) {
      const int b = p % W2_NBUF, zb = p & 1;
      const int lo = pass_lo(rows, npass, p), R = pass_lo(rows, npass, p + 1) - lo;
      const bool active = row < R;
      const long long grow = r_begin + lo + row;
      float* dz0row = dzring + (b * W2_TROWS + row) * OP;    // this row's dz of block 0
      // labels of the outputs this lane owns in the last block (global loads issued early)
      float yv[2][4];
#pragma unroll
      for (int qi = 0; qi < 2; ++qi)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int o = 4 * (oq + 4 * qi) + e;
          yv[qi][e] = (active && o < OUT) ? Y[grow * (long long)OUT + o] : 0.f;
        }
      mbar_wait(&zready[zb], (uint32_t)((p >> 1) & 1));
      // ---- block 0: sum of the F warps' partials, bias, activation
      if (active) {
#pragma unroll
        for (int qi = 0; qi < 2; ++qi) {
          const int q = oq + 4 * qi;
          if (q < NO) {
            const float* zp = zs + ((zb * W2_NF) * W2_TROWS + row) * OP + 4 * q;
            float sacc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int w = 0; w < W2_NF; ++w) {
              float t4[4];
              ld4(zp + w * W2_TROWS * OP, t4);
#pragma unroll
              for (int e = 0; e < 4; ++e) sacc[e] += t4[e];
            }
#pragma unroll
            for (int e = 0; e < 4; ++e) fwd_store<float>(b0, Ws, S0, Z0, row, 4 * q + e, sacc[e]);
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&zfree[zb]);
      // ---- forward through blocks 1..nb-1 (layer.py:276-279 + activation)
      for (int l = 1; l < nb; ++l) {
        const BlockPlan& bk = mp.b[l];
        if (active) {
          const float* ap = sm + mp.b[l - 1].offS + row * bk.ld_in;
          const int kch = bk.in_p >> 2, nq = bk.out_p >> 2;
#pragma unroll
          for (int qi = 0; qi < 2; ++qi) {
            const int q = oq + 4 * qi;
            if (q < nq) {
              const float* wq = Ws + bk.pw + (4 * q) * bk.ld_in;
              float a4[4] = {0.f, 0.f, 0.f, 0.f};
              for (int kc = 0; kc < kch; ++kc) {
                float av[4];
                ld4(ap + 4 * kc, av);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  float wv[4];
                  ld4(wq + e * bk.ld_in + 4 * kc, wv);
#pragma unroll
                  for (int t = 0; t < 4; ++t) a4[e] = fmaf(av[t], wv[t], a4[e]);
                }
              }
#pragma unroll
              for (int e = 0; e < 4; ++e)
                fwd_store<float>(bk, Ws, sm + bk.offS, bk.offZ >= 0 ? sm + bk.offZ : nullptr, row, 4 * q + e, a4[e]);
            }
          }
        }
        __syncwarp();
      }
      // ---- likelihood residual -> dz of the last block (same arithmetic as lik_phase, engine.cuh)
      if (active) {
        const float* Sl = sm + bl.offS + row * bl.ld_out;
        float* Zl = bl.offZ >= 0 ? sm + bl.offZ + row * bl.ld_out : nullptr;
        float* Dl = nb == 1 ? dz0row : sm + bl.offD + row * bl.ld_out;
#pragma unroll
        for (int qi = 0; qi < 2; ++qi) {
          const int q = oq + 4 * qi;
          if (q < (bl.out_p >> 2)) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int o = 4 * q + e;
              float dz = 0.f, cc = 0.f;
              if (o < OUT) {
                const float f = Sl[o], y = yv[qi][e];
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
                  const float sl = eff_slope<float>(bl.act, Ws + (bl.ps >= 0 ? bl.ps : 0), o, (float)bl.alpha);
                  dz = neg ? df * sl : df;
                  cc = neg ? z * df : 0.f;
                } else {
                  dz = df * act_deriv_from_out<float>(bl.act, f);
                }
              }
              Dl[o] = dz;
              if (act_has_slopes(bl.act)) Zl[o] = cc;
            }
          }
        }
      }
      __syncwarp();
      // ---- data gradient: dz_{l-1}[k] = (sum_o dz_l[o] W_l[o][k]) * act'_{l-1}
      for (int l = nb - 1; l >= 1; --l) {
        const BlockPlan& bk = mp.b[l];
        const BlockPlan& pb = mp.b[l - 1];
        if (active) {
          const int kch = bk.in_p >> 2, och = bk.out_p >> 2, ld = bk.ld_in;
          const float* dzr = sm + bk.offD + row * bk.ld_out;
          const float* Sp = sm + pb.offS + row * pb.ld_out;
          float* Zp = pb.offZ >= 0 ? sm + pb.offZ + row * pb.ld_out : nullptr;
          float* Dp = l == 1 ? dz0row : sm + pb.offD + row * pb.ld_out;
#pragma unroll
          for (int qi = 0; qi < 2; ++qi) {
            const int kq4 = oq + 4 * qi;
            if (kq4 < kch) {
              float da[4] = {0.f, 0.f, 0.f, 0.f};
              for (int oc = 0; oc < och; ++oc) {
                float dv[4];
                ld4(dzr + 4 * oc, dv);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  float wv[4];
                  ld4(Ws + bk.pw + (4 * oc + e) * ld + 4 * kq4, wv);
#pragma unroll
                  for (int t = 0; t < 4; ++t) da[t] = fmaf(dv[e], wv[t], da[t]);
                }
              }
#pragma unroll
              for (int t = 0; t < 4; ++t) {
                const int k = 4 * kq4 + t;
                float dzp = 0.f, cp = 0.f;
                if (k < pb.out) {
                  if (act_keeps_z(pb.act)) {
                    const float zz = Zp[k];
                    const bool neg = zz < 0.f;
                    const float sl = eff_slope<float>(pb.act, Ws + (pb.ps >= 0 ? pb.ps : 0), k, (float)pb.alpha);
                    dzp = neg ? da[t] * sl : da[t];
                    cp = neg ? zz * da[t] : 0.f;
                  } else {
                    dzp = da[t] * act_deriv_from_out<float>(pb.act, Sp[k]);
                  }
                }
                Dp[k] = dzp;
                if (act_has_slopes(pb.act)) Zp[k] = cp;
              }
            }
          }
        }
        __syncwarp();
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&dzready[b]);
      // ---- gradients of everything except W1 over this pass' rows (both T warps together)
      t_barrier();
      if (nb > 1) narrow_accum<float>(mp, Ws, G, sm, R, ttid, 32 * W2_NT);
      for (int o = ttid; o < b0.out_p; o += 32 * W2_NT) {
        float sb = 0.f;
        for (int r = 0; r < R; ++r) sb += dzring[(b * W2_TROWS + r) * OP + o];
        G[b0.pb + o] += sb;
        if (act_has_slopes(b0.act)) {
          float sc = 0.f;
          for (int r = 0; r < R; ++r) sc += Z0[r * ldz0 + o];
          const float f = b0.act == ACT_SQPRELU ? 2.f * Ws[b0.ps + o] : 1.f;
          G[b0.ps + o] += f * sc;
        }
      }
      t_barrier();
    }
    float* out = partial + ((size_t)c * S + s) * mp.Ppad;
    for (int i = b0.pb + ttid; i < mp.Ppad; i += 32 * W2_NT) out[i] = G[i];
    const double ws = warp_sum((double)stat);
    if (lane == 0) red[tw] = ws;
    t_barrier();
    if (ttid == 0) {
      double tot = 0.0;
      for (int w = 0; w < W2_NT; ++w) tot += red[w];
      stat_part[(size_t)c * S + s] = tot;
    }
  }
}

// ------------------------------------------------------------------ host side
bool wide2_supported(const ModelPlan& mp) {
  const BlockPlan& b0 = mp.b[0];
  for (int l = 1; l < mp.nb; ++l)
    if (mp.b[l].in_p > 32 || mp.b[l].out_p > 32) return false;
  return mp.D % 4 == 0 && mp.D_p >= 64 && (mp.ld0 >> 2) <= 32 * W2_NB && (b0.out_p >> 2) <= W2_MAXNO &&
         mp.OUT <= 32;
}

// Shared-memory plan: wp.TR = rows per pass (as many as fit, <= 14 ... 16), offsets in floats.
bool plan_wide2(const ModelPlan& mp, ModelPlan& wp, size_t smem_limit) {
  if (!wide2_supported(mp)) return false;
  for (int RB = W2_TROWS; RB >= 4; --RB) {
    wp = mp;
    wp.TR = RB;
    const int OP = wp.b[0].out_p;
    int cur = 0;
    wp.offW = cur; cur += wp.Ppad;
    wp.offX = cur; cur += W2_NBUF * RB * wp.ld0;
    // rows RB..15 of the forward register tiling read past a tile: keep those reads inside the allocation
    cur += (W2_TROWS - RB) * wp.ld0;
    wp.offScr = cur; cur += 2 * W2_NF * W2_TROWS * OP;
    wp.offDa = cur; cur += W2_NBUF * W2_TROWS * OP;
    wp.offDb = wp.offDa;
    for (int l = 0; l < wp.nb; ++l) {
      BlockPlan& b = wp.b[l];
      b.ksplit = 1;
      b.offS = cur; cur += W2_TROWS * b.ld_out;
      if (act_keeps_z(b.act)) { b.offZ = cur; cur += W2_TROWS * b.ld_out; } else b.offZ = -1;
      if (l >= 1) { b.offD = cur; cur += W2_TROWS * b.ld_out; } else b.offD = -1;
    }
    wp.ldmax = wp.b[0].ld_out;
    wp.offG = cur; cur += wp.Ppad - wp.b[0].pb;
    cur = (cur + 3) / 4 * 4;
    wp.offRed = cur; cur += 2 * 8 + 2 * 16;     // 8 doubles + 14 mbarriers (16 reserved)
    wp.smem_elems = cur;
    if ((size_t)cur * 4 <= smem_limit) return RB >= 8;
  }
  return false;
}

template <int NO>
static void launch_no2(const ModelPlan& wp, dim3 g, size_t smem, const float* theta_pad, const float* X,
                       const float* Y, long long N, float* partial, double* stat_part, cudaStream_t st) {
  cudaFuncSetAttribute(k_sweep_wide2<NO>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  k_sweep_wide2<NO><<<g, W2_THREADS, smem, st>>>(wp, (int)g.x, theta_pad, X, Y, N, partial, stat_part);
}

void launch_sweep_wide2(const ModelPlan& wp, int C, int S, const float* theta_pad, const float* X,
                        const float* Y, long long N, float* partial, double* stat_part, cudaStream_t st) {
  dim3 g(S, C);
  const size_t smem = (size_t)wp.smem_elems * sizeof(float);
  switch (wp.b[0].out_p >> 2) {
    case 1: launch_no2<1>(wp, g, smem, theta_pad, X, Y, N, partial, stat_part, st); break;
    case 2: launch_no2<2>(wp, g, smem, theta_pad, X, Y, N, partial, stat_part, st); break;
    case 3: launch_no2<3>(wp, g, smem, theta_pad, X, Y, N, partial, stat_part, st); break;
    case 4: launch_no2<4>(wp, g, smem, theta_pad, X, Y, N, partial, stat_part, st); break;
    default: launch_no2<5>(wp, g, smem, theta_pad, X, Y, N, partial, stat_part, st); break;
  }
}

}  // namespace tbnn
