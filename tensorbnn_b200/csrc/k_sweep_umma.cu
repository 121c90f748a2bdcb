// k_sweep_umma.cu -- row sweep for a WIDE first dense layer (e.g. 784 -> 20, the docs ClassificationExample
// shape) with the two big contractions on the 5th-generation tensor cores (tcgen05, kind::tf32, fp32 accumulators
// in tensor memory), error-compensated: every operand is split x = hi + lo (both TF32-exact, round to nearest) and
// ONE MMA per K step produces all partial products by stacking hi / lo along M and N:
//     forward   [Whi; Wlo] (M) x [Xhi | Xlo]^T (N = 2 rows)      -> Z1^T pieces   (lanes = outputs, columns = rows)
//     backward  [dZhi; dZlo]^T (M) x [XThi | XTlo]^T (N = 2 FC)   -> dW1 pieces    (lanes = outputs, columns = features)
// A tf32 MMA costs ~90 cycles whatever N <= 128 is (tools/umma_rate.cu), so the instruction count is what matters.
// Same contract as k_partial / k_sweep_wide2: one CTA = a contiguous block of training rows of one chain;
// output = this CTA's partial gradient (padded layout) + likelihood statistic.  It replaces, for those rows,
// network.predict (network.py:141-171), layer.predict (layer.py:266-279), the activations, the likelihood
// residuals (likelihood.py:88-94,162-167,225-236) and TF's reverse-mode autodiff of them.
//
// Data: the training matrix is re-laid ONCE per tbnn_set_data into tiles of <= TRc rows x chunks of FC features in
// the tcgen05 SWIZZLE_NONE core-matrix order (8 rows x 16 bytes per core), K-major for both passes (a row-major
// copy for the forward pass, a feature-major copy for the backward pass; MN-major SWIZZLE_NONE operands do not
// work, tools/umma_test.cu), with one extra constant feature (value 1 for real rows) so that the bias and the bias
// gradient fall out of the same MMAs.  A chunk is one contiguous block in HBM/L2 and arrives with a single TMA bulk
// copy (cp.async.bulk -> mbarrier).
//
// Tensor-core accumulation rounds toward zero (tools/umma_test.cu: the error grows by half an ulp per accumulated
// MMA), so chains are kept short: hi*hi and the small cross terms live in different TMEM columns / lanes and are
// summed on the CUDA cores, and dW1 is drained from tensor memory after every chunk instead of accumulating over
// tiles.
//
// Warp roles (384 threads, one CTA per SM):
//   warps 0-3  : tail -- Z1^T from TMEM transposed through shared memory, then thread = row: activation, the narrow
//                blocks >= 1 in registers, likelihood, data gradient back to dZ1 (written as the stacked hi / lo A
//                operand of the backward MMAs); afterwards warps 0-1 drain the dW1 chunks (TMEM -> partial) while
//                warps 2-3 accumulate the gradients of every parameter except W1 / b1;
//   warp 4     : producer -- one TMA bulk copy per chunk into a ring of stages (both passes over X);
//   warp 5     : one thread issues every tcgen05.mma and commits stage releases / phase completions;
//   warps 6-11 : converters -- split the landed fp32 chunk into TF32 hi (in place) and lo, and (forward pass)
//                stage the matching stacked W1 chunk [Whi; Wlo] from the padded theta.
#include "async.cuh"
#include "engine.cuh"
#include "kernels.h"
#include "narrow.cuh"
#include "umma.cuh"

namespace tbnn {

constexpr int US_THREADS = 384;
constexpr int US_CONV_WARP0 = 6;
constexpr int US_NCONV = 6;
constexpr int US_CONV_THREADS = 32 * US_NCONV;
constexpr int US_NSTAGE = 2;
constexpr int US_MAXNB = 4;
constexpr int US_WREG = 6;     // W-chunk float4 per converter thread held in registers while the X chunk flies
constexpr int US_DZ_CG = 8 * 128 + 16;   // bytes between K column groups of the dZ operand (64 rows; +16: no bank conflicts)

__device__ __forceinline__ void tail_barrier() { asm volatile("bar.sync 1, 128;\n" ::: "memory"); }
__device__ __forceinline__ void drain_barrier() { asm volatile("bar.sync 2, 64;\n" ::: "memory"); }
__device__ __forceinline__ void accum_barrier() { asm volatile("bar.sync 3, 64;\n" ::: "memory"); }

// ------------------------------------------------------------------ X re-tiling (once per set_data)
// Two copies, both K-major for their MMA (tcgen05 SWIZZLE_NONE cores of 8 x 16 bytes); feature D is the constant 1:
//   forward  Xt [tile][chunk][(r/8)*RGx + (f/4)*128 + (r%8)*16 + (f%4)*4 bytes], RGx = FC*32    (M = rows,  K = features)
//   backward XtT[tile][chunk][(f/8)*RGt + (r/4)*128 + (f%8)*16 + (r%4)*4 bytes], RGt = TRc*32   (M = features, K = rows)
__global__ void __launch_bounds__(256)
k_tile_x(const __grid_constant__ USweepPlan up, int D, const float* __restrict__ X, long long N,
         float* __restrict__ Xt) {
  const int tile = blockIdx.x, c = blockIdx.y;
  const long long r0 = N * tile / up.ntiles, r1 = N * (tile + 1) / up.ntiles;
  const int R = (int)(r1 - r0);
  const int fq = up.FC >> 2;
  const size_t chunk = ((size_t)tile * up.nch + c) * up.xbytes;
  unsigned char* dst = reinterpret_cast<unsigned char*>(Xt) + chunk;
  unsigned char* dstT = dst + (size_t)up.ntiles * up.nch * up.xbytes;
  const uint32_t RGx = (uint32_t)up.FC * 32u, RGt = (uint32_t)up.TRc * 32u;
  for (int e = threadIdx.x; e < up.TRc * fq; e += blockDim.x) {
    const int r = e / fq, q = e - r * fq;
    float v[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int f = c * up.FC + 4 * q + i;
      v[i] = 0.f;
      if (r < R) v[i] = f < D ? X[(r0 + r) * (long long)D + f] : (f == D ? 1.f : 0.f);
    }
    *reinterpret_cast<float4*>(dst + (uint32_t)(r >> 3) * RGx + (uint32_t)q * 128u + (uint32_t)(r & 7) * 16u) =
        make_float4(v[0], v[1], v[2], v[3]);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int fl = 4 * q + i;
      *reinterpret_cast<float*>(dstT + (uint32_t)(fl >> 3) * RGt + (uint32_t)(r >> 2) * 128u + (uint32_t)(fl & 7) * 16u +
                                (uint32_t)(r & 3) * 4u) = v[i];
    }
  }
}

// ---- per-row helpers of the tail, deliberately NOT inlined and looping at run time: the activation switch and the
// transcendental code exist once instead of once per unrolled element (instruction-cache footprint)
// in: z in srow[0..out_p); out: a in srow, z in zrow (when the block keeps z)
__device__ __noinline__ void us_act_row(const BlockPlan& b, const float* Wt, float* srow, float* zrow) {
  for (int o = 0; o < b.out_p; ++o) {
    float z = 0.f, a = 0.f;
    if (o < b.out) {
      z = srow[o];
      float slope = 0.f;
      if (act_keeps_z(b.act)) slope = eff_slope<float>(b.act, Wt + (b.ps >= 0 ? b.ps : 0), o, (float)b.alpha);
      a = act_fwd<float>(b.act, z, slope);
    }
    srow[o] = a;
    if (zrow) zrow[o] = z;
  }
}
// in: da in drow[0..out_p); out: dz in drow, slope contribution c in zrow (when the block has slopes)
__device__ __noinline__ void us_dact_row(const BlockPlan& pb, const float* Wt, float* drow, const float* srow,
                                         float* zrow) {
  for (int k = 0; k < pb.out_p; ++k) {
    float dzp = 0.f, cp = 0.f;
    if (k < pb.out) {
      const float da = drow[k];
      if (act_keeps_z(pb.act)) {
        const float zz = zrow[k];
        const bool neg = zz < 0.f;
        const float sl = eff_slope<float>(pb.act, Wt + (pb.ps >= 0 ? pb.ps : 0), k, (float)pb.alpha);
        dzp = neg ? da * sl : da;
        cp = neg ? zz * da : 0.f;
      } else {
        dzp = da * act_deriv_from_out<float>(pb.act, srow[k]);
      }
    }
    drow[k] = dzp;
    if (act_has_slopes(pb.act)) zrow[k] = cp;
  }
}
// likelihood residual of one row -> dz of the last block in drow (same arithmetic as narrow_row); returns the
// row's contribution to the likelihood statistic
__device__ __noinline__ float us_lik_row(const ModelPlan& mp, const float* Wt, const float* srow, float* zrow,
                                         float* drow, const float* __restrict__ yrow) {
  const BlockPlan& b = mp.b[mp.nb - 1];
  float stat = 0.f;
  for (int o = 0; o < b.out_p; ++o) {
    float dzo = 0.f, cc = 0.f;
    if (o < mp.OUT) {
      const float f = srow[o], y = yrow[o];
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
      if (act_keeps_z(b.act)) {
        const float zz = zrow[o];
        const bool neg = zz < 0.f;
        const float sl = eff_slope<float>(b.act, Wt + (b.ps >= 0 ? b.ps : 0), o, (float)b.alpha);
        dzo = neg ? df * sl : df;
        cc = neg ? zz * df : 0.f;
      } else {
        dzo = df * act_deriv_from_out<float>(b.act, f);
      }
    }
    drow[o] = dzo;
    if (act_has_slopes(b.act)) zrow[o] = cc;
  }
  return stat;
}

// ------------------------------------------------------------------ the sweep
template <int WMAX>
__global__ void __launch_bounds__(US_THREADS, 1)
k_sweep_umma(const __grid_constant__ ModelPlan mp, const __grid_constant__ USweepPlan up, int S,
             const float* __restrict__ theta_pad, const float* __restrict__ Xt, const float* __restrict__ Y,
             long long N, float* __restrict__ partial, double* __restrict__ stat_part) {
  extern __shared__ __align__(128) unsigned char smraw[];
  __shared__ uint32_t tmem_slot;
  float* sm = reinterpret_cast<float*>(smraw);
  const int c = blockIdx.y, s = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const BlockPlan& b0 = mp.b[0];
  const int D = mp.D, nb = mp.nb, FC = up.FC, nch = up.nch, TRc = up.TRc, NP = up.NP;
  const int tpc = up.ntiles / S;                  // tiles per CTA
  const uint32_t RGx = (uint32_t)FC * 32u;        // bytes between 8-row groups of a forward X chunk / W chunk
  const uint32_t RGt = (uint32_t)TRc * 32u;       // bytes between 8-feature groups of a backward (transposed) X chunk
  const size_t xt_copy = (size_t)up.ntiles * nch * up.xbytes;   // bytes of one tiled copy of X
  double* red = reinterpret_cast<double*>(smraw + up.off_red);
  uint64_t* bars = reinterpret_cast<uint64_t*>(red + 8);
  uint64_t* full = bars;                          // [NSTAGE] X chunk landed (tx bytes)
  uint64_t* conv = bars + US_NSTAGE;              // [NSTAGE] converters done with the stage
  uint64_t* freeb = bars + 2 * US_NSTAGE;         // [NSTAGE] MMAs reading the stage completed
  uint64_t* zdone = bars + 3 * US_NSTAGE;         // forward MMAs of the tile completed
  uint64_t* dzready = zdone + 1;                  // tail published dZ1 (stacked hi / lo operand)
  uint64_t* dwdone = zdone + 2;                   // [2] backward MMAs of a chunk completed (TMEM buffer b)
  uint64_t* dwfree = zdone + 4;                   // [2] the drain warps emptied TMEM buffer b
  const float* th = theta_pad + (size_t)c * mp.Ppad;
  float* Wt = sm + (up.off_wt >> 2) - b0.pb;      // tail parameters, indexed like the padded theta (>= b0.pb)
  float* G = sm + (up.off_g >> 2) - b0.pb;        // their gradient accumulators
  const uint32_t ZC = 2u * (uint32_t)TRc;         // TMEM: Z1^T pieces in columns [0, ZC), dW1 buffers behind
  const uint32_t DWC = 2u * (uint32_t)FC;         // columns of one dW1 buffer

  if (warp == 0) umma::tmem_alloc(&tmem_slot, 512);
  if (tid == 0) {
    for (int i = 0; i < US_NSTAGE; ++i) { mbar_init(&full[i], 1); mbar_init(&conv[i], US_NCONV); mbar_init(&freeb[i], 1); }
    mbar_init(zdone, 1); mbar_init(dzready, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(&dwdone[i], 1); mbar_init(&dwfree[i], 1); }
    mbar_fence_init();
  }
  // the stacked W1 / dZ1 operands have rows that are never written (outputs >= out, rows of the lo half): zero once
  for (int i = tid; i < (up.off_dz + up.dzbytes - up.off_stage) / 16; i += US_THREADS)
    reinterpret_cast<float4*>(smraw + up.off_stage)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  fence_proxy_async();
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tbase = tmem_slot;

  if (warp == 4) {
    // ================================================================= producer
    if (lane == 0) {
      int q = 0;
      for (int ti = 0; ti < tpc; ++ti) {
        const int tile = s * tpc + ti;
        for (int pass = 0; pass < 2; ++pass)
          for (int ch = 0; ch < nch; ++ch, ++q) {
            const int st = q % US_NSTAGE, n = q / US_NSTAGE;
            if (n > 0) mbar_wait(&freeb[st], (uint32_t)((n - 1) & 1));
            unsigned char* dst = smraw + up.off_stage + st * up.stage_bytes;
            mbar_expect_tx(&full[st], (uint32_t)up.xbytes);
            bulk_g2s(dst, reinterpret_cast<const unsigned char*>(Xt) + (size_t)pass * xt_copy +
                              ((size_t)tile * nch + ch) * up.xbytes,
                     (uint32_t)up.xbytes, &full[st]);
          }
      }
    }
  } else if (warp == 5) {
    // ================================================================= MMA issue
    if (lane == 0) {
      // descriptors are built once per stage; the K loop only bumps the start-address field (bytes >> 4)
      const uint32_t idF = umma::idesc_tf32(128, (int)ZC, false, false);    // [Whi; Wlo] x [Xhi | Xlo]
      const uint32_t idB = umma::idesc_tf32(128, (int)DWC, false, false);   // [dZhi; dZlo] x [XThi | XTlo]
      uint64_t dXf[US_NSTAGE], dXb[US_NSTAGE], dW[US_NSTAGE];
      for (int st = 0; st < US_NSTAGE; ++st) {
        const uint32_t x = smem_u32(smraw + up.off_stage + st * up.stage_bytes);
        dXf[st] = umma::smem_desc(x, 128u, RGx);
        dXb[st] = umma::smem_desc(x, 128u, RGt);
        dW[st] = umma::smem_desc(x + 2u * (uint32_t)up.xbytes, 128u, RGx);
      }
      const uint64_t dDZ = umma::smem_desc(smem_u32(smraw + up.off_dz), (uint32_t)US_DZ_CG, 128u);
      int q = 0, g = 0;                              // g: running index of backward chunks (TMEM buffer g & 1)
      for (int ti = 0; ti < tpc; ++ti) {
        const int tile = s * tpc + ti;
        const int R = (int)(N * (tile + 1) / up.ntiles - N * tile / up.ntiles);
        // ---- forward: Z1^T pieces (TMEM columns [0, ZC)) = sum over chunks and K steps
        for (int ch = 0; ch < nch; ++ch, ++q) {
          const int st = q % US_NSTAGE, n = q / US_NSTAGE;
          mbar_wait(&conv[st], (uint32_t)(n & 1));
          umma::fence_after_sync();
          const int feats = min(FC, D + 1 - ch * FC);
          const int nks = (feats + 7) >> 3;
          uint64_t a = dW[st], b = dXf[st];
          for (int ks = 0; ks < nks; ++ks, a += 16, b += 16) umma::mma_tf32_ss(tbase, a, b, idF, ch > 0 || ks > 0);
          umma::commit(&freeb[st]);
        }
        umma::commit(zdone);
        mbar_wait(dzready, (uint32_t)(ti & 1));
        umma::fence_after_sync();
        // ---- backward: dW1 pieces of chunk ch (TMEM buffer g & 1), K = rows of the tile
        const int nkb = (R + 7) >> 3;
        for (int ch = 0; ch < nch; ++ch, ++q, ++g) {
          const int st = q % US_NSTAGE, n = q / US_NSTAGE;
          const int bsel = g & 1, u = g >> 1;
          if (u > 0) { mbar_wait(&dwfree[bsel], (uint32_t)((u - 1) & 1)); umma::fence_after_sync(); }
          mbar_wait(&conv[st], (uint32_t)(n & 1));
          umma::fence_after_sync();
          const uint32_t d = tbase + ZC + (uint32_t)bsel * DWC;
          uint64_t a = dDZ, b = dXb[st];
          for (int ks = 0; ks < nkb; ++ks, a += (2 * US_DZ_CG) >> 4, b += 16) umma::mma_tf32_ss(d, a, b, idB, ks > 0);
          umma::commit(&freeb[st]);
          umma::commit(&dwdone[bsel]);
        }
      }
    }
  } else if (warp >= US_CONV_WARP0) {
    // ================================================================= converters
    const int ct = tid - 32 * US_CONV_WARP0;
    const int xq = up.xbytes >> 4;                 // float4 per X chunk
    const int wq = NP * (FC >> 2);                 // float4 per W chunk (rows < NP)
    int q = 0;
    for (int ti = 0; ti < tpc; ++ti) {
      for (int pass = 0; pass < 2; ++pass)
        for (int ch = 0; ch < nch; ++ch, ++q) {
          const int st = q % US_NSTAGE, n = q / US_NSTAGE;
          unsigned char* sb = smraw + up.off_stage + st * up.stage_bytes;
          // W chunk values into registers first (L2 latency overlaps the wait for the X chunk)
          float4 wv[US_WREG];
          if (pass == 0) {
#pragma unroll
            for (int j = 0; j < US_WREG; ++j) {
              const int e = ct + j * US_CONV_THREADS;
              wv[j] = make_float4(0.f, 0.f, 0.f, 0.f);
              if (e < wq) {
                const int o = e % NP, kq = e / NP, f = ch * FC + 4 * kq;
                if (o < b0.out) {
                  float v[4] = {0.f, 0.f, 0.f, 0.f};
                  if (f + 3 < b0.ld_in) ld4(th + b0.pw + o * b0.ld_in + f, v);
                  else
                    for (int i = 0; i < 4; ++i) if (f + i < b0.ld_in) v[i] = th[b0.pw + o * b0.ld_in + f + i];
#pragma unroll
                  for (int i = 0; i < 4; ++i) {
                    if (f + i == D) v[i] = th[b0.pb + o];
                    else if (f + i > D) v[i] = 0.f;
                  }
                  wv[j] = make_float4(v[0], v[1], v[2], v[3]);
                }
              }
            }
          }
          mbar_wait(&full[st], (uint32_t)(n & 1));
          float4* xh = reinterpret_cast<float4*>(sb);
          float4* xl = reinterpret_cast<float4*>(sb + up.xbytes);
          for (int i = ct; i < xq; i += US_CONV_THREADS) {
            const float4 v = xh[i];
            float4 h, l;
            umma::split_tf32(v.x, h.x, l.x); umma::split_tf32(v.y, h.y, l.y);
            umma::split_tf32(v.z, h.z, l.z); umma::split_tf32(v.w, h.w, l.w);
            xh[i] = h; xl[i] = l;
          }
          if (pass == 0) {
            unsigned char* wh = sb + 2 * up.xbytes;          // rows 0..31: hi, rows 32..63: lo
            unsigned char* wl = wh + 4u * RGx;
#pragma unroll
            for (int j = 0; j < US_WREG; ++j) {
              const int e = ct + j * US_CONV_THREADS;
              if (e < wq) {
                const int o = e % NP, kq = e / NP;
                float4 h, l;
                umma::split_tf32(wv[j].x, h.x, l.x); umma::split_tf32(wv[j].y, h.y, l.y);
                umma::split_tf32(wv[j].z, h.z, l.z); umma::split_tf32(wv[j].w, h.w, l.w);
                const uint32_t off = (uint32_t)(o >> 3) * RGx + (uint32_t)kq * 128u + (uint32_t)(o & 7) * 16u;
                *reinterpret_cast<float4*>(wh + off) = h;
                *reinterpret_cast<float4*>(wl + off) = l;
              }
            }
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) mbar_arrive(&conv[st]);
        }
    }
  } else {
    // ================================================================= tail (warps 0-3)
    const int t = tid;                              // row slot
    const uint32_t lane_base = (uint32_t)(32 * warp);
    for (int i = b0.pb + t; i < mp.Ppad; i += 128) { Wt[i] = th[i]; G[i] = 0.f; }
    tail_barrier();
    double stat = 0.0;
    unsigned char* dzop = smraw + up.off_dz;
    float* dwx = reinterpret_cast<float*>(smraw + up.off_dwx);    // [32][FC]: warp 1's dW1 piece of the chunk
    float* out = partial + ((size_t)c * S + s) * mp.Ppad;
    int g = 0;
    for (int ti = 0; ti < tpc; ++ti) {
      const int tile = s * tpc + ti;
      const long long r0 = N * tile / up.ntiles;
      const int R = (int)(N * (tile + 1) / up.ntiles - r0);
      const bool active = t < R;
      mbar_wait(zdone, (uint32_t)(ti & 1));
      umma::fence_after_sync();
      // ---- Z1^T (lanes = outputs) -> rows in shared memory: warp 0 stores hi*hi + hi*lo into the S_0 rows,
      //      warp 1 stores lo*hi into the dz rows of block 0; the row threads add the two
      if (warp < 2) {
        float* dstb = sm + (warp == 0 ? b0.offS : b0.offD);
        for (int rr = 0; rr < TRc; rr += 8) {
          float v1[8], v2[8];
          umma::tmem_ld8(umma::tmem_addr(tbase, lane_base, rr), v1);
          if (warp == 0) umma::tmem_ld8(umma::tmem_addr(tbase, lane_base, TRc + rr), v2);
          umma::tmem_ld_wait();
          if (lane < b0.out_p) {
#pragma unroll
            for (int i = 0; i < 8; ++i) dstb[(rr + i) * b0.ld_out + lane] = warp == 0 ? v1[i] + v2[i] : v1[i];
          }
        }
      }
      umma::fence_before_sync();
      tail_barrier();
      float a[WMAX], dz[WMAX];
      if (active) {
        float* S0 = sm + b0.offS + t * b0.ld_out;
        const float* L0 = sm + b0.offD + t * b0.ld_out;
        // ---- block 0: z1 = (lo*hi) + (hi*hi + hi*lo); the bias came through the constant feature
#pragma unroll
        for (int o4 = 0; o4 < WMAX; o4 += 4)
          if (o4 < b0.out_p) {
            float q[4], l4[4];
            ld4(S0 + o4, q); ld4(L0 + o4, l4);
#pragma unroll
            for (int e = 0; e < 4; ++e) q[e] += l4[e];
            st4(S0 + o4, q);
          }
        us_act_row(b0, Wt, S0, b0.offZ >= 0 ? sm + b0.offZ + t * b0.ld_out : nullptr);
#pragma unroll
        for (int o4 = 0; o4 < WMAX; o4 += 4) {
          float q[4] = {0.f, 0.f, 0.f, 0.f};
          if (o4 < b0.out_p) ld4(S0 + o4, q);
          a[o4] = q[0]; a[o4 + 1] = q[1]; a[o4 + 2] = q[2]; a[o4 + 3] = q[3];
        }
        // ---- forward through blocks 1..nb-1 (widths <= WMAX), weights broadcast from shared memory
        for (int l = 1; l < nb; ++l) {
          const BlockPlan& b = mp.b[l];
          float* Sl = sm + b.offS + t * b.ld_out;
#pragma unroll
          for (int o4 = 0; o4 < WMAX; o4 += 4) {
            if (o4 < b.out_p) {
              float acc[4];
              ld4(Wt + b.pb + o4, acc);
#pragma unroll
              for (int k4 = 0; k4 < WMAX; k4 += 4) {
                if (k4 < b.in_p) {
#pragma unroll
                  for (int e = 0; e < 4; ++e) {
                    float w[4];
                    ld4(Wt + b.pw + (o4 + e) * b.ld_in + k4, w);
                    acc[e] = fmaf(w[0], a[k4], acc[e]); acc[e] = fmaf(w[1], a[k4 + 1], acc[e]);
                    acc[e] = fmaf(w[2], a[k4 + 2], acc[e]); acc[e] = fmaf(w[3], a[k4 + 3], acc[e]);
                  }
                }
              }
              st4(Sl + o4, acc);
            }
          }
          us_act_row(b, Wt, Sl, b.offZ >= 0 ? sm + b.offZ + t * b.ld_out : nullptr);
          if (l < nb - 1) {
#pragma unroll
            for (int o4 = 0; o4 < WMAX; o4 += 4)
              if (o4 < b.out_p) { float q[4]; ld4(Sl + o4, q); a[o4] = q[0]; a[o4 + 1] = q[1]; a[o4 + 2] = q[2]; a[o4 + 3] = q[3]; }
          }
        }
        // ---- likelihood residual -> dz of the last block
        {
          const BlockPlan& b = mp.b[nb - 1];
          stat += (double)us_lik_row(mp, Wt, sm + b.offS + t * b.ld_out,
                                     b.offZ >= 0 ? sm + b.offZ + t * b.ld_out : nullptr,
                                     sm + b.offD + t * b.ld_out, Y + (r0 + t) * (long long)mp.OUT);
        }
        // ---- data gradient back to block 0: dz_{l-1}[k] = (sum_o W_l[o][k] dz_l[o]) * act'_{l-1}
        for (int l = nb - 1; l >= 1; --l) {
          const BlockPlan& b = mp.b[l];
          const BlockPlan& pb = mp.b[l - 1];
          const float* Dl = sm + b.offD + t * b.ld_out;
          float* Dp = sm + pb.offD + t * pb.ld_out;
          float da[WMAX];
#pragma unroll
          for (int k = 0; k < WMAX; ++k) da[k] = 0.f;
#pragma unroll
          for (int o4 = 0; o4 < WMAX; o4 += 4) {
            if (o4 < b.out_p) {
              float dq[4];
              ld4(Dl + o4, dq);
#pragma unroll
              for (int e = 0; e < 4; ++e) {
#pragma unroll
                for (int k4 = 0; k4 < WMAX; k4 += 4) {
                  if (k4 < b.in_p) {
                    float w[4];
                    ld4(Wt + b.pw + (o4 + e) * b.ld_in + k4, w);
#pragma unroll
                    for (int i = 0; i < 4; ++i) da[k4 + i] = fmaf(w[i], dq[e], da[k4 + i]);
                  }
                }
              }
            }
          }
#pragma unroll
          for (int k4 = 0; k4 < WMAX; k4 += 4)
            if (k4 < pb.out_p) { const float q[4] = {da[k4], da[k4 + 1], da[k4 + 2], da[k4 + 3]}; st4(Dp + k4, q); }
          us_dact_row(pb, Wt, Dp, sm + pb.offS + t * pb.ld_out, pb.offZ >= 0 ? sm + pb.offZ + t * pb.ld_out : nullptr);
        }
        // dz of block 0 back into registers for the operand store
        {
          const float* D0 = sm + b0.offD + t * b0.ld_out;
#pragma unroll
          for (int o4 = 0; o4 < WMAX; o4 += 4) {
            float q[4] = {0.f, 0.f, 0.f, 0.f};
            if (o4 < b0.out_p) ld4(D0 + o4, q);
            dz[o4] = q[0]; dz[o4 + 1] = q[1]; dz[o4 + 2] = q[2]; dz[o4 + 3] = q[3];
          }
        }
      } else {
#pragma unroll
        for (int o = 0; o < WMAX; ++o) dz[o] = 0.f;
      }
      // ---- dZ1 as the stacked K-major A operand of the backward MMAs: rows o (hi) and 32 + o (lo), column = row t
      if (t < TRc) {
        const uint32_t cb = (uint32_t)(t >> 2) * (uint32_t)US_DZ_CG + (uint32_t)(t & 3) * 4u;
#pragma unroll
        for (int o = 0; o < WMAX; ++o) {
          if (o < NP) {
            float h, l;
            umma::split_tf32(o < b0.out ? dz[o] : 0.f, h, l);
            const uint32_t off = cb + (uint32_t)(o >> 3) * 128u + (uint32_t)(o & 7) * 16u;
            *reinterpret_cast<float*>(dzop + off) = h;
            *reinterpret_cast<float*>(dzop + off + 512u) = l;      // row 32 + o: four row groups further
          }
        }
      }
      fence_proxy_async();
      tail_barrier();
      if (t == 0) mbar_arrive(dzready);
      if (warp >= 2) {
        // ---- gradients of everything except W1 / b1 from the batch buffers of this tile (64 threads)
        const int t2 = t - 64;
        if (nb > 1) narrow_accum<float>(mp, Wt, G, sm, R, t2, 64);
        if (act_has_slopes(b0.act)) {
          for (int o = t2; o < b0.out; o += 64) {
            float sc = 0.f;
            for (int r = 0; r < R; ++r) sc += sm[b0.offZ + r * b0.ld_out + o];
            const float f = b0.act == ACT_SQPRELU ? 2.f * Wt[b0.ps + o] : 1.f;
            G[b0.ps + o] += f * sc;
          }
        }
      } else {
        // ---- drain the dW1 chunks: lanes 0..31 = dZhi rows (x XThi in columns [0, FC), x XTlo in [FC, 2 FC)),
        //      lanes 32..63 = dZlo rows (x XThi).  dW1[o][f] = (hi*lo + lo*hi) + hi*hi, written to the partial
        for (int ch = 0; ch < nch; ++ch, ++g) {
          const int bsel = g & 1, u = g >> 1;
          mbar_wait(&dwdone[bsel], (uint32_t)(u & 1));
          umma::fence_after_sync();
          const uint32_t col0 = ZC + (uint32_t)bsel * DWC;
          if (warp == 1) {
            for (int f0 = 0; f0 < FC; f0 += 8) {
              float v[8];
              umma::tmem_ld8(umma::tmem_addr(tbase, lane_base, col0 + f0), v);
              umma::tmem_ld_wait();
              if (lane < b0.out_p) {
                *reinterpret_cast<float4*>(dwx + lane * FC + f0) = make_float4(v[0], v[1], v[2], v[3]);
                *reinterpret_cast<float4*>(dwx + lane * FC + f0 + 4) = make_float4(v[4], v[5], v[6], v[7]);
              }
            }
          }
          umma::fence_before_sync();
          drain_barrier();
          if (warp == 0) {
            for (int f0 = 0; f0 < FC; f0 += 8) {
              float v1[8], v2[8];
              umma::tmem_ld8(umma::tmem_addr(tbase, lane_base, col0 + f0), v1);
              umma::tmem_ld8(umma::tmem_addr(tbase, lane_base, col0 + FC + f0), v2);
              umma::tmem_ld_wait();
              if (lane < b0.out_p) {
                const float4 xa = *reinterpret_cast<const float4*>(dwx + lane * FC + f0);
                const float4 xb = *reinterpret_cast<const float4*>(dwx + lane * FC + f0 + 4);
                const float v3[8] = {xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z, xb.w};
                const int f = ch * FC + f0;
                float* dst = out + b0.pw + lane * b0.ld_in + f;
#pragma unroll
                for (int i = 0; i < 8; ++i) v1[i] = (v2[i] + v3[i]) + v1[i];
                if (f + 7 < D) {
                  if (ti > 0) {
                    const float4 o1 = *reinterpret_cast<const float4*>(dst), o2 = *reinterpret_cast<const float4*>(dst + 4);
                    v1[0] += o1.x; v1[1] += o1.y; v1[2] += o1.z; v1[3] += o1.w;
                    v1[4] += o2.x; v1[5] += o2.y; v1[6] += o2.z; v1[7] += o2.w;
                  }
                  *reinterpret_cast<float4*>(dst) = make_float4(v1[0], v1[1], v1[2], v1[3]);
                  *reinterpret_cast<float4*>(dst + 4) = make_float4(v1[4], v1[5], v1[6], v1[7]);
                } else {
#pragma unroll
                  for (int i = 0; i < 8; ++i) {
                    if (f + i < D) dst[i] = ti > 0 ? dst[i] + v1[i] : v1[i];
                    else if (f + i == D) out[b0.pb + lane] = ti > 0 ? out[b0.pb + lane] + v1[i] : v1[i];
                  }
                }
              }
            }
          }
          umma::fence_before_sync();
          drain_barrier();
          if (t == 0) mbar_arrive(&dwfree[bsel]);
        }
      }
      tail_barrier();
    }
    // ---- likelihood statistic of the CTA, the other gradients, zero padding of the W1 rows
    const double ws = warp_sum(stat);
    if (lane == 0) red[warp] = ws;
    tail_barrier();
    if (t == 0) stat_part[(size_t)c * S + s] = (red[0] + red[1]) + (red[2] + red[3]);
    for (int e = t; e < b0.out_p * (b0.ld_in - D); e += 128) {
      const int o = e / (b0.ld_in - D), k = D + e - o * (b0.ld_in - D);
      out[b0.pw + o * b0.ld_in + k] = 0.f;
    }
    for (int i = b0.pb + b0.out_p + t; i < mp.Ppad; i += 128) out[i] = G[i];
  }
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tbase, 512);
}

// ------------------------------------------------------------------ host side
static inline int up8(int x) { return (x + 7) & ~7; }

bool usweep_supported(const ModelPlan& mp) {
  const BlockPlan& b0 = mp.b[0];
  if (mp.nb > US_MAXNB || b0.out_p > 32 || mp.OUT > 32 || mp.D < 64) return false;
  for (int l = 1; l < mp.nb; ++l)
    if (mp.b[l].in_p > 32 || mp.b[l].out_p > 32) return false;
  return true;
}

// Plan for N rows split over S CTAs per chain.  wp = mp + batch-buffer offsets (floats) of the tail.
bool plan_usweep(const ModelPlan& mp, long long N, int S, size_t smem_limit, ModelPlan& wp, USweepPlan& up) {
  if (!usweep_supported(mp) || N <= 0 || S < 1) return false;
  memset(&up, 0, sizeof(up));
  const BlockPlan& b0 = mp.b[0];
  long long per = (N + S - 1) / S;
  int TRc = (int)std::min<long long>(128, up8((int)std::min<long long>(per, 128)));
  const long long need = (N + TRc - 1) / TRc;
  const long long tpc = (need + S - 1) / S;
  if (tpc * S > 0x7fffffffLL) return false;
  up.TRc = TRc;
  up.ntiles = (int)(tpc * S);
  up.NP = up8(b0.out_p);
  up.NM = 64;                                            // stacked operand rows: 32 hi + 32 lo
  up.dz_cg = US_DZ_CG;
  up.dzbytes = (TRc / 4) * US_DZ_CG + 128 * 16;          // + slack: the M = 128 MMA reads 16 row groups
  // tail: batch buffers (S_l, Z_l, dz_l per row) at the start of shared memory
  wp = mp;
  wp.TR = TRc;
  int cur = 0;
  for (int l = 0; l < wp.nb; ++l) {
    BlockPlan& b = wp.b[l];
    b.ksplit = 1;
    b.offS = cur; cur += TRc * b.ld_out;
    if (act_keeps_z(b.act)) { b.offZ = cur; cur += TRc * b.ld_out; } else b.offZ = -1;
    b.offD = cur; cur += TRc * b.ld_out;
  }
  const int buf_floats = (cur + 3) & ~3;
  const int tailp = (mp.Ppad - b0.pb + 3) & ~3;
  const int fixed = buf_floats * 4 + 2 * tailp * 4 + up.dzbytes + 8 * 8 + 16 * 8 + 512;
  // chunk width: as wide as two stages, the drain buffer and the 512 TMEM columns allow, then balanced
  const int feats = mp.D + 1;
  int best = 0;
  for (int FC = 128; FC >= 8; FC -= 8) {
    const long long stage = 2LL * TRc * FC * 4 + 64LL * FC * 4;
    if (US_NSTAGE * stage + 32LL * FC * 4 + fixed <= (long long)smem_limit && 2 * TRc + 4 * FC <= 512) { best = FC; break; }
  }
  if (!best) return false;
  up.nch = (feats + best - 1) / best;
  up.FC = up8((feats + up.nch - 1) / up.nch);
  up.xbytes = TRc * up.FC * 4;
  up.wbytes = 64 * up.FC * 4;
  up.stage_bytes = 2 * up.xbytes + up.wbytes;
  int off = buf_floats * 4;
  up.off_wt = off; off += tailp * 4;
  up.off_g = off; off += tailp * 4;
  up.off_dwx = off; off += 32 * up.FC * 4;
  off = (off + 127) & ~127;
  up.off_stage = off; off += US_NSTAGE * up.stage_bytes;
  up.off_dz = off; off += up.dzbytes;
  off = (off + 15) & ~15;
  // the M = 128 MMAs read 16 row groups of the stacked W1 operand (only 8 exist; rows >= 64 are never used):
  // keep those reads inside the allocation
  const int last_w = up.off_stage + (US_NSTAGE - 1) * up.stage_bytes + 2 * up.xbytes;
  off = std::max(off, last_w + 16 * up.FC * 32);
  up.off_red = off; off += 8 * 8 + 16 * 8;
  up.smem_bytes = (off + 15) & ~15;
  wp.smem_elems = up.smem_bytes / 4;
  if ((size_t)up.smem_bytes > smem_limit) return false;
  if (up.NP * (up.FC / 4) > US_WREG * US_CONV_THREADS) return false;
  return true;
}

// two tiled copies: row-major cores (forward pass) and feature-major cores (backward pass)
size_t usweep_xt_bytes(const USweepPlan& up) { return 2 * (size_t)up.ntiles * up.nch * up.xbytes; }

void launch_tile_x(const USweepPlan& up, int D, const float* X, long long N, float* Xt, cudaStream_t st) {
  dim3 g(up.ntiles, up.nch);
  k_tile_x<<<g, 256, 0, st>>>(up, D, X, N, Xt);
}

template <int WMAX>
static void launch_us(const ModelPlan& wp, const USweepPlan& up, dim3 g, const float* theta_pad, const float* Xt,
                      const float* Y, long long N, float* partial, double* stat_part, cudaStream_t st) {
  cudaFuncSetAttribute(k_sweep_umma<WMAX>, cudaFuncAttributeMaxDynamicSharedMemorySize, up.smem_bytes);
  k_sweep_umma<WMAX><<<g, US_THREADS, up.smem_bytes, st>>>(wp, up, (int)g.x, theta_pad, Xt, Y, N, partial, stat_part);
}

void launch_sweep_umma(const ModelPlan& wp, const USweepPlan& up, int C, int S, const float* theta_pad,
                       const float* Xt, const float* Y, long long N, float* partial, double* stat_part,
                       cudaStream_t st) {
  dim3 g(S, C);
  int w = up.NP;
  for (int l = 0; l < wp.nb; ++l) w = std::max(w, up8(wp.b[l].out_p));
  if (w <= 8) launch_us<8>(wp, up, g, theta_pad, Xt, Y, N, partial, stat_part, st);
  else if (w <= 16) launch_us<16>(wp, up, g, theta_pad, Xt, Y, N, partial, stat_part, st);
  else if (w <= 24) launch_us<24>(wp, up, g, theta_pad, Xt, Y, N, partial, stat_part, st);
  else launch_us<32>(wp, up, g, theta_pad, Xt, Y, N, partial, stat_part, st);
}

}  // namespace tbnn
