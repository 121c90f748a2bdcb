// k_wide.cu -- row sweep specialised for a WIDE first dense layer (e.g. 784 -> 20, the
// docs ClassificationExample shape): fp32 FFMA with register-resident tiles.
//
// Same contract as k_partial (k_main.cu): one CTA = a contiguous block of training rows of one
// chain; output = this CTA's partial gradient (padded layout) + likelihood statistic.  What is
// specialised is block 0 (layer.py:266-279 forward and TF's autodiff of it):
//   * X rows arrive by TMA bulk copies (cp.async.bulk -> mbarrier), double buffered per 16-row pass;
//   * forward  z1[r][o] = sum_k X[r][k] W1[o][k]: lane = (k-quad, row%8), thread tile = 2 rows x all
//     outputs, W1 read from shared memory as warp-wide broadcasts, split-K across the 8 warps,
//     transposing shuffle reduction + one shared-memory pass across warps;
//   * backward dW1[o][k] += dz1[r][o] X[r][k]: thread = one 4-wide k chunk x all outputs, the
//     accumulators stay in REGISTERS for the CTA's whole row range (no shared-memory traffic for dW1);
//   * blocks >= 1 must be narrow (widths <= 32): they run one row per warp with lane = neuron
//     (narrow.cuh), three CTA barriers per pass; their weight gradients are accumulated from
//     shared-memory batch buffers once every ModelPlan::TR rows.
#include "async.cuh"
#include "engine.cuh"
#include "kernels.h"
#include "narrow.cuh"

namespace tbnn {

constexpr int NTW = 256;   // threads per CTA
constexpr int WTR = 16;    // rows per pass (= ModelPlan::TR of the wide plan)
constexpr int WRG = 2;     // 8-row groups per pass
static_assert(WTR == 8 * WRG && WTR == 2 * (NTW / 32), "one row per half-warp in the narrow tail");

template <int NO, bool BWD>
__global__ void __launch_bounds__(NTW, 1)
k_sweep_wide(const __grid_constant__ ModelPlan mp, int S, const float* __restrict__ theta_pad,
             const float* __restrict__ X, const float* __restrict__ Y, long long N,
             float* __restrict__ partial, double* __restrict__ stat_part, long long* __restrict__ prof) {
  extern __shared__ __align__(16) unsigned char smraw[];
  float* sm = reinterpret_cast<float*>(smraw);
  constexpr int OP = 4 * NO;                 // padded outputs of block 0
  // phase clocks of CTA (0,0) (developer aid, tbnn_wide_profile): prof[i] = clock64 at mark i
  int nprof = 0;
#define WIDE_MARK()                                                                    \
  do {                                                                                 \
    if (prof != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0 && nprof < 63) \
      prof[1 + nprof++] = clock64();                                                   \
  } while (0)
  const int c = blockIdx.y, s = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const BlockPlan& b0 = mp.b[0];
  const int ld0 = mp.ld0, D = mp.D, nch = mp.D_p >> 2, ldz = b0.ld_out;
  const float* thg = theta_pad + (size_t)c * mp.Ppad;
  float* Ws = sm + mp.offW;
  TileCtx<float> cx;
  cx.sm = sm;
  cx.Wp = Ws;
  cx.G = sm + mp.offG - b0.pb;               // accumulators of everything except W1
  double* red = reinterpret_cast<double*>(sm + mp.offRed);
  uint64_t* bars = reinterpret_cast<uint64_t*>(red + 40);
  const long long r_begin = N * s / S, r_end = N * (s + 1) / S;
  const int npass = (int)((r_end - r_begin + WTR - 1) / WTR);

  const int PB = mp.TR / WTR;                 // passes per batch of the narrow tail
  float* S0 = sm + b0.offS;
  float* Z0 = b0.offZ >= 0 ? sm + b0.offZ : nullptr;
  float* dZ0 = sm + mp.offDa;                 // dz of block 0 for the current pass [WTR][ldz]

  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    mbar_init(&bars[2], 1);
    mbar_fence_init();
  }
  if (BWD)
    for (int i = b0.pb + tid; i < mp.Ppad; i += NTW) cx.G[i] = 0.f;
  // padding columns of both X buffers (TMA writes exactly D floats per row)
  for (int e = tid; e < 2 * WTR * (ld0 - D); e += NTW) {
    const int r = e / (ld0 - D), j = e - r * (ld0 - D);
    sm[mp.offX + r * ld0 + D + j] = 0.f;
  }
  __syncthreads();

  // Rows past this CTA's range are loaded too when they exist (they belong to the next CTA and are
  // masked out below), so every pass except the data set's very last one is a full 16-row copy.
  // issued by the first 16 lanes of warp 0, one row each (lane 0 arms the barrier first)
  auto issue = [&](int p) {
    const long long row0 = r_begin + (long long)p * WTR;
    const int nload = (int)((N - row0) < WTR ? (N - row0) : WTR);
    float* dst = sm + mp.offX + (p & 1) * WTR * ld0;
    if (lane == 0) {
      fence_proxy_async();
      mbar_expect_tx(&bars[p & 1], (uint32_t)(nload * D * 4));
    }
    __syncwarp();
    if (lane < nload)
      bulk_g2s(dst + lane * ld0, X + (row0 + lane) * (long long)D, (uint32_t)(D * 4), &bars[p & 1]);
  };
  if (warp == 0) {
    if (lane == 0) {
      // the chain's padded parameters: one bulk copy (Ppad is a multiple of 4 floats)
      fence_proxy_async();
      mbar_expect_tx(&bars[2], (uint32_t)(mp.Ppad * 4));
      bulk_g2s(Ws, thg, (uint32_t)(mp.Ppad * 4), &bars[2]);
    }
    if (npass > 0) issue(0);
    if (npass > 1) issue(1);
  }
  WIDE_MARK();   // 0: prologue issued
  float accW[BWD ? OP : 1][4];
#pragma unroll
  for (int o = 0; o < (BWD ? OP : 1); ++o)
#pragma unroll
    for (int q = 0; q < 4; ++q) accW[o][q] = 0.f;
  float stat = 0.f;

  for (int p = 0; p < npass; ++p) {
    float* Xs = sm + mp.offX + (p & 1) * WTR * ld0;
    const long long row0 = r_begin + (long long)p * WTR;
    const int nr = (int)((r_end - row0) < WTR ? (r_end - row0) : WTR);
    const int rb0 = (p % PB) * WTR;           // first batch-buffer row of this pass
    if (N - row0 < WTR) {                      // the data set's last, short tile: rows that do not exist
      for (int e = tid; e < (WTR - (int)(N - row0)) * mp.D_p; e += NTW) {
        const int r = (int)(N - row0) + e / mp.D_p, j = e % mp.D_p;
        Xs[r * ld0 + j] = 0.f;
      }
      __syncthreads();
    }
    WIDE_MARK();   // pass start
    if (p == 0) mbar_wait(&bars[2], 0u);
    mbar_wait(&bars[p & 1], (uint32_t)((p >> 1) & 1));
    WIDE_MARK();   // data arrived

    // ---------------- block 0 forward: split-K over warps, lane = (k quad, row % 8)
    {
      const int kq = lane >> 3, r8 = lane & 7;
      const int rg_eff = (nr + 7) >> 3;
      float acc[WRG][OP];
#pragma unroll
      for (int g = 0; g < WRG; ++g)
#pragma unroll
        for (int o = 0; o < OP; ++o) acc[g][o] = 0.f;
      const int nsc = (nch + 3) >> 2;
      const float* w0 = Ws + b0.pw;
      for (int sc = warp; sc < nsc; sc += NTW / 32) {
        const int ch = 4 * sc + kq;
        if (ch < nch) {
          float xv[WRG][4];
#pragma unroll
          for (int g = 0; g < WRG; ++g) {
            if (g < rg_eff) ld4(Xs + (r8 + 8 * g) * ld0 + 4 * ch, xv[g]);
            else { xv[g][0] = xv[g][1] = xv[g][2] = xv[g][3] = 0.f; }
          }
#pragma unroll
          for (int o = 0; o < OP; ++o) {
            float wv[4];
            ld4(w0 + o * ld0 + 4 * ch, wv);
#pragma unroll
            for (int g = 0; g < WRG; ++g)
              if (g < rg_eff) {
#pragma unroll
                for (int q = 0; q < 4; ++q) acc[g][o] = fmaf(xv[g][q], wv[q], acc[g][o]);
              }
          }
        }
      }
      // transposing reduction over the 4 k-quads: lane keeps the NO outputs o = b0*2NO + b1*NO + j
      const bool bit0 = (kq & 1) != 0, bit1 = (kq & 2) != 0;
      float* scr = sm + mp.offScr;
#pragma unroll
      for (int g = 0; g < WRG; ++g) {
        float h1[2 * NO];
#pragma unroll
        for (int j = 0; j < 2 * NO; ++j) {
          const float a = acc[g][j], b = acc[g][j + 2 * NO];
          const float send = bit0 ? a : b, keep = bit0 ? b : a;
          h1[j] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
        }
#pragma unroll
        for (int j = 0; j < NO; ++j) {
          const float a = h1[j], b = h1[j + NO];
          const float send = bit1 ? a : b, keep = bit1 ? b : a;
          const float v = keep + __shfl_xor_sync(0xffffffffu, send, 16);
          const int o = (bit0 ? 2 * NO : 0) + (bit1 ? NO : 0) + j;
          scr[(warp * WTR + r8 + 8 * g) * OP + o] = v;
        }
      }
      WIDE_MARK();   // fwd FFMA + shuffle done (this warp)
      __syncthreads();                         // (1) every warp is also past block 0 backward of pass p-1
      WIDE_MARK();   // barrier 1 passed
      if (warp == 0 && p >= 1 && p + 1 < npass) issue(p + 1);
      for (int e = tid; e < WTR * OP; e += NTW) {
        const int r = e / OP, o = e - r * OP;
        float v = 0.f;
#pragma unroll
        for (int w = 0; w < NTW / 32; ++w) v += scr[(w * WTR + r) * OP + o];
        fwd_store<float>(b0, Ws, S0 + rb0 * ldz, Z0 ? Z0 + rb0 * ldz : nullptr, r, o, v);
      }
      __syncthreads();                         // (2)
      WIDE_MARK();   // reduce + act done
    }

    // ---------------- narrow tail: one row per half-warp (forward, likelihood, data gradient to dz0)
    {
      const int r = 2 * warp + (lane >> 4);
      stat += narrow_row<float>(mp, Ws, sm, rb0 + r, lane & 15, r < nr, Y, row0 + r, dZ0 + r * ldz);
    }
    WIDE_MARK();   // narrow done (this warp)
    __syncthreads();                           // (3)
    WIDE_MARK();   // barrier 3 passed

    if (BWD) {
      // ---------------- block 0 backward: dW1 in registers, db1 / slope gradients by the last warp
      if (tid < nch) {
        const float* xr = Xs + 4 * tid;
#pragma unroll 2
        for (int r = 0; r < nr; ++r) {
          float xv[4];
          ld4(xr + r * ld0, xv);
#pragma unroll
          for (int jo = 0; jo < NO; ++jo) {
            float dv[4];
            ld4(dZ0 + r * ldz + 4 * jo, dv);
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
              for (int q = 0; q < 4; ++q)
                accW[BWD ? 4 * jo + i : 0][q] = fmaf(dv[i], xv[q], accW[BWD ? 4 * jo + i : 0][q]);
          }
        }
      } else if (tid >= NTW - 32) {
        for (int o = lane; o < b0.out_p; o += 32) {
          float sb = 0.f;
          for (int r = 0; r < nr; ++r) sb += dZ0[r * ldz + o];
          cx.G[b0.pb + o] += sb;
          if (act_has_slopes(b0.act)) {
            const float* Zc = Z0 + rb0 * ldz;
            float sc2 = 0.f;
            for (int r = 0; r < nr; ++r) sc2 += Zc[r * ldz + o];
            const float f = b0.act == ACT_SQPRELU ? 2.f * Ws[b0.ps + o] : 1.f;
            cx.G[b0.ps + o] += f * sc2;
          }
        }
      }
      WIDE_MARK();   // bwd FFMA done (this warp)
      // ---------------- blocks >= 1: gradient accumulation once per batch of PB passes
      if (mp.nb > 1 && ((p + 1) % PB == 0 || p + 1 == npass)) {
        narrow_accum<float>(mp, Ws, cx.G, sm, rb0 + nr);
        __syncthreads();                       // batch buffers are reused by the next pass
      }
    }
  }
  if (npass == 0) mbar_wait(&bars[2], 0u);    // never exit with a bulk copy in flight
  __syncthreads();
  WIDE_MARK();   // all passes done

  if (BWD) {
    float* out = partial + ((size_t)c * S + s) * mp.Ppad;
    if (tid < nch) {
#pragma unroll
      for (int o = 0; o < OP; ++o) st4(out + b0.pw + o * ld0 + 4 * tid, accW[BWD ? o : 0]);
    } else if (tid < (ld0 >> 2)) {
      const float z[4] = {0.f, 0.f, 0.f, 0.f};
      for (int o = 0; o < OP; ++o) st4(out + b0.pw + o * ld0 + 4 * tid, z);
    }
    for (int i = b0.pb + tid; i < mp.Ppad; i += NTW) out[i] = cx.G[i];
  }
  const double tot = block_sum((double)stat, red);
  if (tid == 0) stat_part[(size_t)c * S + s] = tot;
  WIDE_MARK();   // epilogue done
  if (prof != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) prof[0] = nprof;
#undef WIDE_MARK
}

template <int NO>
static void launch_no(const ModelPlan& wp, dim3 g, size_t smem, bool backward, const float* theta_pad,
                      const float* X, const float* Y, long long N, float* partial, double* stat_part,
                      long long* prof, cudaStream_t st) {
  if (backward) {
    cudaFuncSetAttribute(k_sweep_wide<NO, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k_sweep_wide<NO, true><<<g, NTW, smem, st>>>(wp, (int)g.x, theta_pad, X, Y, N, partial, stat_part, prof);
  } else {
    cudaFuncSetAttribute(k_sweep_wide<NO, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k_sweep_wide<NO, false><<<g, NTW, smem, st>>>(wp, (int)g.x, theta_pad, X, Y, N, partial, stat_part, prof);
  }
}

bool wide_supported(const ModelPlan& mp) {
  const BlockPlan& b0 = mp.b[0];
  for (int l = 1; l < mp.nb; ++l)
    if (mp.b[l].in_p > 32 || mp.b[l].out_p > 32) return false;
  return mp.D % 4 == 0 && mp.D_p >= 64 && (mp.D_p >> 2) <= NTW - 32 && (mp.ld0 >> 2) <= NTW &&
         b0.out_p <= 32 && mp.OUT <= 32;
}
int wide_rows_per_pass() { return WTR; }
int wide_scratch_elems(const ModelPlan& mp) { return (NTW / 32) * WTR * mp.b[0].out_p; }

void launch_sweep_wide(const ModelPlan& wp, int C, int S, bool backward, const float* theta_pad,
                       const float* X, const float* Y, long long N, float* partial, double* stat_part,
                       cudaStream_t st, long long* prof) {
  dim3 g(S, C);
  const size_t smem = (size_t)wp.smem_elems * sizeof(float);
  switch (wp.b[0].out_p >> 2) {
    case 1: launch_no<1>(wp, g, smem, backward, theta_pad, X, Y, N, partial, stat_part, prof, st); break;
    case 2: launch_no<2>(wp, g, smem, backward, theta_pad, X, Y, N, partial, stat_part, prof, st); break;
    case 3: launch_no<3>(wp, g, smem, backward, theta_pad, X, Y, N, partial, stat_part, prof, st); break;
    case 4: launch_no<4>(wp, g, smem, backward, theta_pad, X, Y, N, partial, stat_part, prof, st); break;
    case 5: launch_no<5>(wp, g, smem, backward, theta_pad, X, Y, N, partial, stat_part, prof, st); break;
    case 6: launch_no<6>(wp, g, smem, backward, theta_pad, X, Y, N, partial, stat_part, prof, st); break;
    case 7: launch_no<7>(wp, g, smem, backward, theta_pad, X, Y, N, partial, stat_part, prof, st); break;
    default: launch_no<8>(wp, g, smem, backward, theta_pad, X, Y, N, partial, stat_part, prof, st); break;
  }
}

}  // namespace tbnn
