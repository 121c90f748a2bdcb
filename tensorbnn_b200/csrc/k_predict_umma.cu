// k_predict_umma.cu -- posterior-predictive sweep (predictor.predict, predictor.py:132-155) on the
// 5th-generation tensor cores: stored samples x test rows for MLPs whose hidden layers are GEMM-shaped.
//
// One CTA owns a block of test rows and marches through the samples (like k_predict.cu); the per-row
// running mean / M2 of the fused-moments mode live in shared memory.  Per sample the hidden weight
// matrices are split into TF32 hi / lo parts and laid out as K-major SWIZZLE_NONE core matrices in
// shared memory.  Per 128-row tile:
//   block 0      (D <= 8 inputs)  : CUDA cores, thread = row; a1 is split hi / lo and written straight
//                                   into TENSOR MEMORY as the A operand of the first GEMM (tcgen05.st);
//   blocks 1..nb-2                : Z[128 x N] = A[128 x K] W^T with tcgen05.mma kind::tf32, A from TMEM,
//                                   B from shared memory, error-compensated 3xTF32 (lo*hi + hi*lo + hi*hi),
//                                   fp32 accumulators in TMEM; the epilogue (tcgen05.ld, bias, activation,
//                                   split) writes the next A operand back to TMEM -- activations never
//                                   touch shared or global memory;
//   last block   (<= 4 outputs)   : CUDA cores, folded into the last epilogue.
// The two warpgroups of the CTA work on different tiles, so one warpgroup's epilogue overlaps the
// other's MMAs.
#include "engine.cuh"
#include "kernels.h"
#include "umma.cuh"

namespace tbnn {

constexpr int PU_THREADS = 256;
constexpr int PU_MAXD = 8;
constexpr int PU_MAXOUT = 4;

struct PredUmmaPlan {
  int NW;                    // TMEM column stride of one operand / accumulator region (multiple of 16)
  int Kp[MAXB], Np[MAXB];    // padded K (multiple of 8) and N (multiple of 16) of GEMM blocks 1..nb-2
  int wofs[MAXB];            // shared-memory byte offset of block l's hi weights; lo at + wbytes[l]
  int wbytes[MAXB];
  // per-sample small parameters, staged as zero-padded fp32 arrays of NW entries (byte offsets):
  int bias_ofs[MAXB];        // bias of block l (l <= nb-2)
  int slope_ofs[MAXB];       // effective negative-side slope of block l (s^2 / s / alpha), else unused
  int w0_ofs;                // block 0 weights transposed: [D][NW]
  int wl_ofs;                // last block weights: [OUT][NW], followed by its bias [4]
  int acc_ofs;               // byte offset of the moment accumulators [rows_per_cta][OUT][2]
  int smem_bytes;            // without the accumulators
  int hidden_act;            // activation shared by blocks 0..nb-2, or -1 when they differ
};

__device__ __forceinline__ void wg_barrier(int wg) {
  asm volatile("bar.sync %0, 128;\n" ::"r"(1 + wg) : "memory");
}

// hidden activation, specialised at compile time (A = -1: per-block runtime switch)
template <int A> __device__ __forceinline__ float act_hidden(int act_rt, float z, float slope) {
  if (A == ACT_RELU) return fmaxf(z, 0.f);
  if (A == ACT_TANH) return tanhf(z);
  if (A == ACT_SIGMOID) return 1.f / (1.f + expf(-z));
  if (A == ACT_LEAKY || A == ACT_PRELU || A == ACT_SQPRELU) return z < 0.f ? slope * z : z;
  if (A == ACT_NONE) return z;
  return act_fwd<float>(act_rt, z, slope);
}

template <int A>
__global__ void __launch_bounds__(PU_THREADS, 1)
k_predict_umma(const __grid_constant__ ModelPlan mp, const __grid_constant__ PredUmmaPlan pu,
               const float* __restrict__ samples, long long s0, long long S_chunk,
               const float* __restrict__ X, long long M, int rows_per_cta, float* __restrict__ out,
               float* __restrict__ moments) {
  extern __shared__ __align__(128) unsigned char smraw[];
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(8) uint64_t bars[2];
  // warp index through a shuffle: the compiler then treats everything derived from it (roles, column groups, tensor-
  // memory columns) as warp-uniform -- uniform branches and registers instead of per-thread ones
  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
  const int wg = warp >> 2, wq = warp & 3, trow = 32 * wq + lane;
  float* acc = reinterpret_cast<float*>(smraw + pu.acc_ofs);
  const long long r_begin = (long long)blockIdx.x * rows_per_cta;
  if (r_begin >= M) return;
  const int nrows = (int)((M - r_begin) < rows_per_cta ? (M - r_begin) : rows_per_cta);
  const int ntiles = (nrows + 127) >> 7;
  const int nb = mp.nb, OUT = mp.OUT, D = mp.D, NW = pu.NW;
  const BlockPlan& b0 = mp.b[0];
  const BlockPlan& bl = mp.b[nb - 1];
  const float* w0t = reinterpret_cast<const float*>(smraw + pu.w0_ofs);
  const float* wlp = reinterpret_cast<const float*>(smraw + pu.wl_ofs);

  if (warp == 0) umma::tmem_alloc(&tmem_slot, 512);
  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    mbar_fence_init();
  }
  if (moments) {
    for (int e = tid; e < nrows * OUT; e += PU_THREADS) {
      const int lr = e / OUT, o = e - lr * OUT;
      const long long gi = (long long)o * M + r_begin + lr;
      acc[2 * e] = s0 > 0 ? moments[(long long)OUT * M + gi] : 0.f;
      acc[2 * e + 1] = s0 > 0 ? moments[2 * (long long)OUT * M + gi] : 0.f;
    }
  }
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tbase = tmem_slot;
  const uint32_t colZ = (uint32_t)(wg * 256), colAh = colZ + NW, colAl = colZ + 2 * NW;
  const uint32_t lane_base = (uint32_t)(32 * wq);
  uint32_t phase = 0;

  for (long long s = 0; s < S_chunk; ++s) {
    __syncthreads();   // every warpgroup is done with the previous sample's parameters
    const float* src = samples + s * (long long)mp.P;
    // ---- small parameters: zero-padded arrays (bias, effective slope, block 0 / last block weights)
    for (int l = 0; l <= nb - 2; ++l) {
      const BlockPlan& b = mp.b[l];
      float* bp = reinterpret_cast<float*>(smraw + pu.bias_ofs[l]);
      float* sp = reinterpret_cast<float*>(smraw + pu.slope_ofs[l]);
      for (int n = tid; n < NW; n += PU_THREADS) {
        bp[n] = n < b.out ? src[b.fb + n] : 0.f;
        float sl = 0.f;
        if (n < b.out) {
          if (b.act == ACT_PRELU) sl = src[b.fs + n];
          else if (b.act == ACT_SQPRELU) { const float t = src[b.fs + n]; sl = t * t; }
          else if (b.act == ACT_LEAKY) sl = (float)b.alpha;
        }
        sp[n] = sl;
      }
    }
    for (int e = tid; e < D * NW; e += PU_THREADS) {
      const int d = e / NW, n = e - d * NW;
      reinterpret_cast<float*>(smraw + pu.w0_ofs)[e] = n < b0.out ? src[b0.fw + n * D + d] : 0.f;
    }
    for (int e = tid; e < OUT * NW + 4; e += PU_THREADS) {
      float v = 0.f;
      if (e < OUT * NW) {
        const int o = e / NW, n = e - o * NW;
        if (n < bl.in) v = src[bl.fw + o * bl.in + n];
      } else if (e - OUT * NW < OUT) {
        v = src[bl.fb + e - OUT * NW];
      }
      reinterpret_cast<float*>(smraw + pu.wl_ofs)[e] = v;
    }
    // ---- GEMM weights: TF32 hi / lo split, K-major core-matrix layout
    for (int l = 1; l <= nb - 2; ++l) {
      const BlockPlan& b = mp.b[l];
      const int Kp = pu.Kp[l], Np = pu.Np[l];
      unsigned char* wh = smraw + pu.wofs[l];
      unsigned char* wl = wh + pu.wbytes[l];
      const uint32_t cg = 128u * (uint32_t)(Np >> 3);
      const int kq = Kp >> 2;
      for (int e = tid; e < Np * kq; e += PU_THREADS) {   // one 16-byte chunk (4 consecutive k) per step
        const int n = e / kq, k = 4 * (e - n * kq);
        float hi[4], lo[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float v = (n < b.out && k + i < b.in) ? src[b.fw + n * b.in + k + i] : 0.f;
          umma::split_tf32_trunc(v, hi[i], lo[i]);
        }
        const uint32_t o = umma::core_off(n, k, 128u, cg);
        st4(reinterpret_cast<float*>(wh + o), hi);
        st4(reinterpret_cast<float*>(wl + o), lo);
      }
    }
    fence_proxy_async();
    __syncthreads();
    const float cnt = (float)(s0 + s + 1);

    for (int tile = wg; tile < ntiles; tile += 2) {
      const int lr = tile * 128 + trow;
      const bool valid = lr < nrows;
      const long long row = r_begin + lr;
      // ---------------- block 0 on CUDA cores -> A operand (hi / lo) in TMEM
      float xv[PU_MAXD];
#pragma unroll
      for (int d = 0; d < PU_MAXD; ++d) xv[d] = (valid && d < D) ? X[row * (long long)D + d] : 0.f;
      {
        const int K1 = pu.Kp[1];
        const float* bp = reinterpret_cast<const float*>(smraw + pu.bias_ofs[0]);
        const float* sp = reinterpret_cast<const float*>(smraw + pu.slope_ofs[0]);
        for (int c0 = 0; c0 < K1; c0 += 8) {
          float z[8], sl[8], h[8], l[8];
          ld4(bp + c0, *reinterpret_cast<float(*)[4]>(&z[0]));
          ld4(bp + c0 + 4, *reinterpret_cast<float(*)[4]>(&z[4]));
          ld4(sp + c0, *reinterpret_cast<float(*)[4]>(&sl[0]));
          ld4(sp + c0 + 4, *reinterpret_cast<float(*)[4]>(&sl[4]));
#pragma unroll
          for (int d = 0; d < PU_MAXD; ++d) {
            if (d < D) {
              float w[8];
              ld4(w0t + d * NW + c0, *reinterpret_cast<float(*)[4]>(&w[0]));
              ld4(w0t + d * NW + c0 + 4, *reinterpret_cast<float(*)[4]>(&w[4]));
#pragma unroll
              for (int i = 0; i < 8; ++i) z[i] = fmaf(w[i], xv[d], z[i]);
            }
          }
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            float a = act_hidden<A>(b0.act, z[i], sl[i]);
            if (A == -1 && c0 + i >= b0.out) a = 0.f;
            umma::split_tf32_trunc(a, h[i], l[i]);
          }
          umma::tmem_st8(umma::tmem_addr(tbase, lane_base, colAh + c0), h);
          umma::tmem_st8(umma::tmem_addr(tbase, lane_base, colAl + c0), l);
        }
        umma::tmem_st_wait();
      }
      float f[PU_MAXOUT];
#pragma unroll
      for (int o = 0; o < PU_MAXOUT; ++o) f[o] = o < OUT ? wlp[OUT * NW + o] : 0.f;

      for (int l = 1; l <= nb - 2; ++l) {
        const BlockPlan& b = mp.b[l];
        const int Kp = pu.Kp[l], Np = pu.Np[l];
        umma::fence_before_sync();
        wg_barrier(wg);
        if (wq == 0 && lane == 0) {
          umma::fence_after_sync();
          const uint32_t id = umma::idesc_tf32(128, Np, false, false);
          const uint32_t bh = smem_u32(smraw + pu.wofs[l]), blo = bh + (uint32_t)pu.wbytes[l];
          const uint32_t cg = 128u * (uint32_t)(Np >> 3);
          const uint32_t d = umma::tmem_addr(tbase, 0, colZ);
          // B descriptors as 32-bit words: they stay in uniform registers and advance with one add per k step (with
          // 64-bit descriptors every MMA was wrapped in an R2UR waterfall: 91 cycles per MMA, profiles/r2h_summary.md)
          const uint32_t hiw = umma::desc_hi(128u), stepB = (2u * cg) >> 4;
          const uint32_t bH0 = umma::desc_lo(bh, cg), bL0 = umma::desc_lo(blo, cg);
          for (int ks = 0; ks < (Kp >> 3); ++ks) {
            const uint32_t tAh = umma::tmem_addr(tbase, 0, colAh + 8 * ks);
            const uint32_t tAl = umma::tmem_addr(tbase, 0, colAl + 8 * ks);
            umma::mma_tf32_ts32(d, tAl, bH0 + ks * stepB, hiw, id, ks > 0);
            umma::mma_tf32_ts32(d, tAh, bL0 + ks * stepB, hiw, id, true);
            umma::mma_tf32_ts32(d, tAh, bH0 + ks * stepB, hiw, id, true);
          }
          umma::commit(&bars[wg]);
        }
        mbar_wait(&bars[wg], phase);
        phase ^= 1u;
        umma::fence_after_sync();
        // ---------------- epilogue: bias + activation; next A operand to TMEM, or the last block on CUDA cores
        const bool last_gemm = (l == nb - 2);
        const float* bp = reinterpret_cast<const float*>(smraw + pu.bias_ofs[l]);
        const float* sp = reinterpret_cast<const float*>(smraw + pu.slope_ofs[l]);
        for (int c0 = 0; c0 < Np; c0 += 8) {
          float v[8], bz[8], sl[8], h[8], lo8[8];
          umma::tmem_ld8(umma::tmem_addr(tbase, lane_base, colZ + c0), v);
          ld4(bp + c0, *reinterpret_cast<float(*)[4]>(&bz[0]));
          ld4(bp + c0 + 4, *reinterpret_cast<float(*)[4]>(&bz[4]));
          ld4(sp + c0, *reinterpret_cast<float(*)[4]>(&sl[0]));
          ld4(sp + c0 + 4, *reinterpret_cast<float(*)[4]>(&sl[4]));
          umma::tmem_ld_wait();
          float a[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            a[i] = act_hidden<A>(b.act, v[i] + bz[i], sl[i]);
            if (A == -1 && c0 + i >= b.out) a[i] = 0.f;
          }
          if (last_gemm) {
#pragma unroll
            for (int o = 0; o < PU_MAXOUT; ++o) {
              if (o < OUT) {
                float w[8];
                ld4(wlp + o * NW + c0, *reinterpret_cast<float(*)[4]>(&w[0]));
                ld4(wlp + o * NW + c0 + 4, *reinterpret_cast<float(*)[4]>(&w[4]));
#pragma unroll
                for (int i = 0; i < 8; ++i) f[o] = fmaf(w[i], a[i], f[o]);
              }
            }
          } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) umma::split_tf32_trunc(a[i], h[i], lo8[i]);
            umma::tmem_st8(umma::tmem_addr(tbase, lane_base, colAh + c0), h);
            umma::tmem_st8(umma::tmem_addr(tbase, lane_base, colAl + c0), lo8);
          }
        }
        if (!last_gemm) umma::tmem_st_wait();
      }
      // ---------------- output activation, store / moments
      if (valid) {
#pragma unroll
        for (int o = 0; o < PU_MAXOUT; ++o) {
          if (o < OUT) {
            float slope = 0.f;
            if (act_keeps_z(bl.act)) slope = eff_slope<float>(bl.act, src + (bl.fs >= 0 ? bl.fs : 0), o, (float)bl.alpha);
            const float y = act_fwd<float>(bl.act, f[o], slope);
            if (out) out[((s0 + s) * OUT + o) * M + row] = y;
            if (moments) {
              const int ai = 2 * (lr * OUT + o);
              const float mean = acc[ai], dlt = y - mean;
              const float mnew = mean + dlt / cnt;
              acc[ai] = mnew;
              acc[ai + 1] += dlt * (y - mnew);
            }
          }
        }
      }
    }
  }
  umma::fence_before_sync();
  __syncthreads();
  if (moments) {
    for (int e = tid; e < nrows * OUT; e += PU_THREADS) {
      const int lr = e / OUT, o = e - lr * OUT;
      const long long gi = (long long)o * M + r_begin + lr;
      moments[gi] = (float)(s0 + S_chunk);
      moments[(long long)OUT * M + gi] = acc[2 * e];
      moments[2 * (long long)OUT * M + gi] = acc[2 * e + 1];
    }
  }
  if (warp == 0) umma::tmem_dealloc(tbase, 512);
}

// ------------------------------------------------------------------ host side
static inline int padto(int x, int m) { return (x + m - 1) / m * m; }

static bool make_plan(const ModelPlan& mp, PredUmmaPlan& pu) {
  if (mp.nb < 3 || mp.D > PU_MAXD || mp.OUT > PU_MAXOUT) return false;
  int NW = 16, cur = 0;
  for (int l = 1; l <= mp.nb - 2; ++l) {
    const BlockPlan& b = mp.b[l];
    if (b.in > 128 || b.out > 128) return false;
    pu.Kp[l] = padto(b.in, 8);
    pu.Np[l] = padto(b.out, 16);
    NW = std::max(NW, std::max(padto(b.in, 16), pu.Np[l]));
    pu.wofs[l] = cur;
    pu.wbytes[l] = pu.Kp[l] * pu.Np[l] * 4;
    cur += 2 * pu.wbytes[l];
  }
  // A operand of GEMM l+1 has K = Kp[l+1] <= Np[l] columns; everything fits in NW columns
  for (int l = 1; l < mp.nb - 2; ++l)
    if (pu.Kp[l + 1] > pu.Np[l]) return false;
  if (3 * NW > 256) return false;            // two warpgroups x (Z, A hi, A lo) in 512 TMEM columns
  pu.NW = NW;
  for (int l = 0; l <= mp.nb - 2; ++l) {
    pu.bias_ofs[l] = cur; cur += NW * 4;
    pu.slope_ofs[l] = cur; cur += NW * 4;
  }
  pu.w0_ofs = cur; cur += mp.D * NW * 4;
  pu.wl_ofs = cur; cur += padto((mp.OUT * NW + 4) * 4, 128);
  pu.hidden_act = mp.b[0].act;
  for (int l = 1; l <= mp.nb - 2; ++l)
    if (mp.b[l].act != pu.hidden_act) pu.hidden_act = -1;
  if (mp.b[0].out > NW) return false;
  pu.acc_ofs = cur;
  pu.smem_bytes = cur;
  return cur + 128 * mp.OUT * 8 <= 220 * 1024;
}

bool predict_umma_supported(const ModelPlan& mp) {
  PredUmmaPlan pu;
  return make_plan(mp, pu);
}

// samples: flat [S_chunk][P] (network.states order).  Returns false when the network is not eligible.
bool launch_predict_umma(const ModelPlan& mp, int num_sms, const float* samples, long long s0, long long S_chunk,
                         const float* X, long long M, float* out, float* moments, cudaStream_t st) {
  PredUmmaPlan pu;
  if (!make_plan(mp, pu)) return false;
  // rows per CTA: a multiple of 128, one CTA per SM when M is large, bounded by the accumulators
  const int room = (220 * 1024 - pu.smem_bytes) / (mp.OUT * 8);
  long long rows = (M + num_sms - 1) / num_sms;
  rows = (rows + 127) / 128 * 128;
  const long long cap = std::max(128, room / 128 * 128);
  if (rows > cap) rows = cap;
  const int grid = (int)((M + rows - 1) / rows);
  const size_t smem = (size_t)pu.smem_bytes + (size_t)rows * mp.OUT * 8;
#define PU_LAUNCH(ACTV)                                                                                   \
  do {                                                                                                    \
    cudaFuncSetAttribute(k_predict_umma<ACTV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);   \
    k_predict_umma<ACTV><<<grid, PU_THREADS, smem, st>>>(mp, pu, samples, s0, S_chunk, X, M, (int)rows, out, \
                                                         moments);                                        \
  } while (0)
  switch (pu.hidden_act) {
    case ACT_RELU: PU_LAUNCH(ACT_RELU); break;
    case ACT_TANH: PU_LAUNCH(ACT_TANH); break;
    case ACT_SIGMOID: PU_LAUNCH(ACT_SIGMOID); break;
    case ACT_LEAKY: case ACT_PRELU: case ACT_SQPRELU: PU_LAUNCH(ACT_SQPRELU); break;
    default: PU_LAUNCH(-1); break;
  }
#undef PU_LAUNCH
  return true;
}

}  // namespace tbnn
