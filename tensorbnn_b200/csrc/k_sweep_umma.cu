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

#include <cstdio>
#include <cstdlib>
#include <vector>

namespace tbnn {

constexpr int US_THREADS = 384;          // warps 0-3: TMEM warps, 4: producer, 5: MMA issue, 6-11: hybrid
constexpr int US_HYB_WARP0 = 6;          // hybrid warps: converters while X streams, row workers in the tail phase
constexpr int US_NHYB = 6;
constexpr int US_HYB_THREADS = 32 * US_NHYB;
constexpr int US_ROW_WARPS = 4 + US_NHYB;      // row workers: the TMEM warps + the hybrid warps
constexpr int US_ROW_THREADS = 32 * US_ROW_WARPS;
constexpr int US_ROWS_PER_PASS = 8 * US_ROW_WARPS;   // four threads per row
constexpr int US_NSTAGE = 2;
constexpr int US_MAXNB = 4;
constexpr int US_WREG = 6;     // W-chunk float4 per hybrid thread held in registers while the X chunk flies
constexpr int US_PROF_N = 64;
constexpr int US_PROF_ROLES = 12;
constexpr int US_DZ_CG = 8 * 128 + 16;   // bytes between K column groups of the dZ operand (64 rows; +16: no bank conflicts)

// A whole warp waits on an mbarrier: ONE lane waits (parked by the hardware, see mbar_wait_parked), the others sit at
// the warp barrier.
// 384 threads spinning on try_wait saturate the shared-memory pipe that the TMA writes, the MMA operand reads and
// every LDS / STS of the working warps go through (measured: everything in the CTA ran ~8x slower).
__device__ __forceinline__ void warp_wait(uint64_t* bar, uint32_t parity) {
  if ((threadIdx.x & 31) == 0) {
    mbar_wait_parked(bar, parity);
  }
  __syncwarp();
}
__device__ __forceinline__ void tail_barrier() { asm volatile("bar.sync 1, 128;\n" ::: "memory"); }
__device__ __forceinline__ void rows_barrier() { asm volatile("bar.sync 4, %0;\n" ::"n"(US_ROW_THREADS) : "memory"); }
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

// ---- per-row helpers of the tail.  A row is shared by a QUAD of threads: each handles the elements o0, o0 + step,
// ...  Everything the helpers need from the plan is passed BY VALUE: a reference to a BlockPlan (kernel parameter or
// shared-memory copy) turns into generic-pointer loads and special-register reads inside the loops.
struct ActP { int act, out, out_p, ps; float alpha; };
__device__ __forceinline__ ActP actp(const BlockPlan& b) {
  ActP a;
  a.act = b.act; a.out = b.out; a.out_p = b.out_p; a.ps = b.ps >= 0 ? b.ps : 0; a.alpha = (float)b.alpha;
  return a;
}
// in: z in srow[o]; out: a in srow, z in zrow (when the block keeps z)
__device__ __forceinline__ void us_act_row(const ActP b, const float* Wt, float* srow, float* zrow, int o0, int step) {
  for (int o = o0; o < b.out_p; o += step) {
    float z = 0.f, a = 0.f;
    if (o < b.out) {
      z = srow[o];
      float slope = 0.f;
      if (act_keeps_z(b.act)) slope = eff_slope<float>(b.act, Wt + b.ps, o, b.alpha);
      a = act_fwd<float>(b.act, z, slope);
    }
    srow[o] = a;
    if (zrow) zrow[o] = z;
  }
}
// in: da in drow[k]; out: dz in drow, slope contribution c in zrow (when the block has slopes)
__device__ __forceinline__ void us_dact_row(const ActP pb, const float* Wt, float* drow, const float* srow, float* zrow,
                                            int k0, int step) {
  for (int k = k0; k < pb.out_p; k += step) {
    float dzp = 0.f, cp = 0.f;
    if (k < pb.out) {
      const float da = drow[k];
      if (act_keeps_z(pb.act)) {
        const float zz = zrow[k];
        const bool neg = zz < 0.f;
        const float sl = eff_slope<float>(pb.act, Wt + pb.ps, k, pb.alpha);
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
// contribution of the elements o0, o0 + step, ... to the likelihood statistic
__device__ __forceinline__ float us_lik_row(const ActP b, int lik, int OUT, const float* Wt, const float* srow,
                                            float* zrow, float* drow, const float* __restrict__ yrow, int o0, int step) {
  float stat = 0.f;
  for (int o = o0; o < b.out_p; o += step) {
    float dzo = 0.f, cc = 0.f;
    if (o < OUT) {
      const float f = srow[o], y = yrow[o];
      float df;
      if (lik == LIK_BERN) {
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
        const float sl = eff_slope<float>(b.act, Wt + b.ps, o, b.alpha);
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
// acc[j] (j < NJ) += sum_k vrow[k] * wbase[j * wstep + k] over k < n (multiple of 4): straight-line LDS.128 + FFMA,
// no guards inside (the caller picks NJ = number of vectors this thread owns)
template <int NJ>
__device__ __forceinline__ void us_dot_rows(const float* vrow, const float* wbase, int wstep, int n, float (&acc)[8]) {
  for (int k4 = 0; k4 < n; k4 += 4) {
    float av[4];
    ld4(vrow + k4, av);
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      float w[4];
      ld4(wbase + j * wstep + k4, w);
      acc[j] = fmaf(w[0], av[0], acc[j]); acc[j] = fmaf(w[1], av[1], acc[j]);
      acc[j] = fmaf(w[2], av[2], acc[j]); acc[j] = fmaf(w[3], av[3], acc[j]);
    }
  }
}
__device__ __forceinline__ void us_dot_rows_n(int nj, const float* vrow, const float* wbase, int wstep, int n,
                                              float (&acc)[8]) {
  switch (nj) {
    case 1: us_dot_rows<1>(vrow, wbase, wstep, n, acc); break;
    case 2: us_dot_rows<2>(vrow, wbase, wstep, n, acc); break;
    case 3: us_dot_rows<3>(vrow, wbase, wstep, n, acc); break;
    case 4: us_dot_rows<4>(vrow, wbase, wstep, n, acc); break;
    case 5: us_dot_rows<5>(vrow, wbase, wstep, n, acc); break;
    case 6: us_dot_rows<6>(vrow, wbase, wstep, n, acc); break;
    case 7: us_dot_rows<7>(vrow, wbase, wstep, n, acc); break;
    case 8: us_dot_rows<8>(vrow, wbase, wstep, n, acc); break;
    default: break;
  }
}

// ---- converter role (hybrid warps), plain inlined functions: lambdas capturing the register arrays by reference
// put them on the stack, and with ~212 KB of shared memory carved out the L1 is too small to hide local loads.
// Issues the loads of this thread's pieces of W1 chunk ch (no use of the values here: the L2 latency overlaps the
// X conversion of the current chunk).
__device__ __forceinline__ void us_load_w(const float* __restrict__ th, int pw, int ld_in, int FC, int ch,
                                          const int (&w_ok)[US_WREG], float4 (&wv)[US_WREG]) {
#pragma unroll
  for (int j = 0; j < US_WREG; ++j) {
    wv[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (w_ok[j] >= 0) {
      const int o = w_ok[j] & 0xff, f = ch * FC + 4 * (w_ok[j] >> 8);
      if (f + 3 < ld_in) wv[j] = *reinterpret_cast<const float4*>(th + pw + o * ld_in + f);
    }
  }
}
// lo = rn_tf32(x - trunc_tf32(x)) of the landed chunk: the raw fp32 chunk IS the hi operand (kind::tf32 ignores the
// low 13 mantissa bits, tools/umma_test.cu) and the lo x lo products are kept, so the split is exact up to the
// rounding of lo.
__device__ __forceinline__ void us_split_x(const float4* __restrict__ xh, float4* __restrict__ xl, int xq, int ht) {
#pragma unroll 4
  for (int i = ht; i < xq; i += US_HYB_THREADS) {
    const float4 v = xh[i];
    float4 l;
    l.x = umma::rn_tf32(v.x - __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u));
    l.y = umma::rn_tf32(v.y - __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u));
    l.z = umma::rn_tf32(v.z - __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u));
    l.w = umma::rn_tf32(v.w - __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u));
    xl[i] = l;
  }
}
// stacked [Whi; Wlo] chunk (rows 0..31 hi, 32..63 lo) from the registers loaded by us_load_w; the column of the
// constant feature D carries the bias (the few threads that own it fetch it here)
__device__ __forceinline__ void us_store_w(unsigned char* wh, uint32_t RGx, const float* __restrict__ th, int pb, int D,
                                           int FC, int ch, const int (&w_ok)[US_WREG], const float4 (&wv)[US_WREG]) {
  unsigned char* wl = wh + 4u * RGx;
#pragma unroll
  for (int j = 0; j < US_WREG; ++j) {
    if (w_ok[j] >= 0) {
      const int o = w_ok[j] & 0xff, kq = w_ok[j] >> 8, f = ch * FC + 4 * kq;
      float v[4] = {wv[j].x, wv[j].y, wv[j].z, wv[j].w};
      if (f + 3 >= D) {
        const float bias = (f <= D) ? th[pb + o] : 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if (f + i == D) v[i] = bias;
          else if (f + i > D) v[i] = 0.f;
        }
      }
      float4 h, l;
      umma::split_tf32(v[0], h.x, l.x); umma::split_tf32(v[1], h.y, l.y);
      umma::split_tf32(v[2], h.z, l.z); umma::split_tf32(v[3], h.w, l.w);
      const uint32_t off = (uint32_t)(o >> 3) * RGx + (uint32_t)kq * 128u + (uint32_t)(o & 7) * 16u;
      *reinterpret_cast<float4*>(wh + off) = h;
      *reinterpret_cast<float4*>(wl + off) = l;
    }
  }
}

// ------------------------------------------------------------------ the sweep
__global__ void __launch_bounds__(US_THREADS, 1)
k_sweep_umma(const __grid_constant__ ModelPlan mp, const __grid_constant__ USweepPlan up, int S,
             const float* __restrict__ theta_pad, const float* __restrict__ Xt, const float* __restrict__ Y,
             long long N, float* __restrict__ partial, double* __restrict__ stat_part, long long* prof) {
  extern __shared__ __align__(128) unsigned char smraw[];
  __shared__ uint32_t tmem_slot;
  // The plan is a __grid_constant__ kernel parameter; a reference to one of its blocks handed to a non-inlined helper
  // (or indexed at run time) becomes a GENERIC pointer into parameter space whose loads cost hundreds of cycles each.
  // The tail works on a copy in shared memory instead.
  __shared__ ModelPlan smp;
  float* sm = reinterpret_cast<float*>(smraw);
  const int c = blockIdx.y, s = blockIdx.x;
  for (int i = threadIdx.x; i < (int)(sizeof(ModelPlan) / 4); i += US_THREADS)
    reinterpret_cast<int*>(&smp)[i] = reinterpret_cast<const int*>(&mp)[i];
  // developer aid (TBNN_US_PROF=1): clock64 marks of CTA 0, one row of US_PROF_N slots per role
  if (blockIdx.x != 0 || blockIdx.y != 0) prof = nullptr;
#define US_MARK(role, idx) do { if (prof && (idx) < US_PROF_N) prof[(role) * US_PROF_N + (idx)] = clock64(); } while (0)
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const BlockPlan& b0 = mp.b[0];
  const int D = mp.D, nb = mp.nb, FC = up.FC, nch = up.nch, TRc = up.TRc, NP = up.NP;
  const int tpc = up.ntiles / S;                  // tiles per CTA
  const uint32_t RGx = (uint32_t)FC * 32u;        // bytes between 8-row groups of a forward X chunk / W chunk
  const uint32_t RGt = (uint32_t)TRc * 32u;       // bytes between 8-feature groups of a backward (transposed) X chunk
  const size_t xt_copy = (size_t)up.ntiles * nch * up.xbytes;   // bytes of one tiled copy of X
  double* red = reinterpret_cast<double*>(smraw + up.off_red);
  uint64_t* bars = reinterpret_cast<uint64_t*>(red + 16);
  uint64_t* full = bars;                          // [NSTAGE] X chunk landed (tx bytes)
  uint64_t* conv = bars + US_NSTAGE;              // [NSTAGE] converters done with the stage
  uint64_t* freeb = bars + 2 * US_NSTAGE;         // [NSTAGE] MMAs reading the stage completed
  uint64_t* zdone = bars + 3 * US_NSTAGE;         // forward MMAs of the tile completed
  uint64_t* dzready = zdone + 1;                  // the row workers published dZ1 (stacked hi / lo operand)
  uint64_t* dwdone = zdone + 2;                   // [2] backward MMAs of a chunk completed (TMEM buffer b)
  uint64_t* dwfree = zdone + 4;                   // [2] the drain warps emptied TMEM buffer b
  const float* th = theta_pad + (size_t)c * mp.Ppad;
  float* Wt = sm + (up.off_wt >> 2) - b0.pb;      // tail parameters, indexed like the padded theta (>= b0.pb)
  float* G = sm + (up.off_g >> 2) - b0.pb;        // their gradient accumulators
  float* WT = sm + (up.off_wtt >> 2);             // transposed weights of blocks >= 1: WT_l[k][o], ld = ld_out
  const uint32_t ZC = 2u * (uint32_t)TRc;         // TMEM: Z1^T pieces in columns [0, ZC), dW1 buffers behind
  const uint32_t DWC = 2u * (uint32_t)FC;         // columns of one dW1 buffer

  if (warp == 0) umma::tmem_alloc(&tmem_slot, 512);
  if (tid == 0) {
    for (int i = 0; i < US_NSTAGE; ++i) { mbar_init(&full[i], 1); mbar_init(&conv[i], US_NHYB); mbar_init(&freeb[i], 1); }
    mbar_init(zdone, 1); mbar_init(dzready, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(&dwdone[i], 1); mbar_init(&dwfree[i], 1); }
    mbar_fence_init();
  }
  // the stacked W1 / dZ1 operands have rows that are never written (outputs >= out): zero them once
  for (int i = tid; i < (up.off_dz + up.dzbytes - up.off_stage) / 16; i += US_THREADS)
    reinterpret_cast<float4*>(smraw + up.off_stage)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  // tail parameters, their transposes (backward pass of the row workers) and zeroed accumulators
  for (int i = b0.pb + tid; i < mp.Ppad; i += US_THREADS) { Wt[i] = th[i]; G[i] = 0.f; }
  for (int l = 1; l < nb; ++l) {
    const BlockPlan& b = mp.b[l];
    float* wt = WT + up.wtt_ofs[l];
    for (int e = tid; e < b.in_p * b.ld_out; e += US_THREADS) {
      const int k = e / b.ld_out, o = e - k * b.ld_out;
      wt[e] = o < b.out_p ? th[b.pw + o * b.ld_in + k] : 0.f;
    }
  }
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
            if (n > 0) mbar_wait_parked(&freeb[st], (uint32_t)((n - 1) & 1));
            unsigned char* dst = smraw + up.off_stage + st * up.stage_bytes;
            US_MARK(0, q);
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
          mbar_wait_parked(&conv[st], (uint32_t)(n & 1));
          umma::fence_after_sync();
          US_MARK(1, q);
          const int feats = min(FC, D + 1 - ch * FC);
          const int nks = (feats + 7) >> 3;
          uint64_t a = dW[st], b = dXf[st];
          for (int ks = 0; ks < nks; ++ks, a += 16, b += 16) umma::mma_tf32_ss(tbase, a, b, idF, ch > 0 || ks > 0);
          umma::commit(&freeb[st]);
          US_MARK(2, q);
        }
        umma::commit(zdone);
        mbar_wait_parked(dzready, (uint32_t)(ti & 1));
        umma::fence_after_sync();
        // ---- backward: dW1 pieces of chunk ch (TMEM buffer g & 1), K = rows of the tile
        const int nkb = (R + 7) >> 3;
        for (int ch = 0; ch < nch; ++ch, ++q, ++g) {
          const int st = q % US_NSTAGE, n = q / US_NSTAGE;
          const int bsel = g & 1, u = g >> 1;
          if (u > 0) { mbar_wait_parked(&dwfree[bsel], (uint32_t)((u - 1) & 1)); umma::fence_after_sync(); }
          mbar_wait_parked(&conv[st], (uint32_t)(n & 1));
          umma::fence_after_sync();
          US_MARK(1, q);
          const uint32_t d = tbase + ZC + (uint32_t)bsel * DWC;
          uint64_t a = dDZ, b = dXb[st];
          for (int ks = 0; ks < nkb; ++ks, a += (2 * US_DZ_CG) >> 4, b += 16) umma::mma_tf32_ss(d, a, b, idB, ks > 0);
          umma::commit(&freeb[st]);
          umma::commit(&dwdone[bsel]);
          US_MARK(2, q);
        }
      }
    }
  } else {
    // ================================================================= TMEM warps 0-3 and hybrid warps 6-11
    const bool hybrid = warp >= US_HYB_WARP0;
    const ModelPlan& P = smp;                        // shared-memory copy (see above)
    const BlockPlan& B0 = P.b[0];
    const int ht = tid - 32 * US_HYB_WARP0;          // hybrid thread index (converter role)
    const int rw = hybrid ? warp - 2 : warp;         // row-worker warp index 0..9
    const int part = lane & 3;                       // a row is shared by a quad of lanes
    const uint32_t lane_base = (uint32_t)(32 * (warp & 3));
    const int xq = up.xbytes >> 4;                   // float4 per X chunk
    const int wq = NP * (FC >> 2);                   // float4 per W chunk (rows < NP)
    // converter role: (output row | K quad << 8) of this thread's W-chunk pieces, fixed for the whole kernel
    int w_ok[US_WREG];
#pragma unroll
    for (int j = 0; j < US_WREG; ++j) {
      const int e = ht + j * US_HYB_THREADS;
      w_ok[j] = -1;
      if (hybrid && e < wq) {
        const int o = e % NP;
        if (o < b0.out) w_ok[j] = o | ((e / NP) << 8);
      }
    }

    double stat = 0.0;
    unsigned char* dzop = smraw + up.off_dz;
    float* dwx = reinterpret_cast<float*>(smraw + up.off_dwx);    // [32][FC]: warp 1's dW1 piece of the chunk
    float* out = partial + ((size_t)c * S + s) * mp.Ppad;
    int q = 0, g = 0;
    for (int ti = 0; ti < tpc; ++ti) {
      const int tile = s * tpc + ti;
      const long long r0 = N * tile / up.ntiles;
      const int R = (int)(N * (tile + 1) / up.ntiles - r0);
      // ------------------------------------------------ forward pass
      if (hybrid) {
        for (int ch = 0; ch < nch; ++ch, ++q) {
          const int st = q % US_NSTAGE, n = q / US_NSTAGE;
          unsigned char* sb = smraw + up.off_stage + st * up.stage_bytes;
          // the W1 loads are issued first and consumed after the X split: their L2 latency is hidden behind it
          float4 wv[US_WREG];
          us_load_w(th, b0.pw, b0.ld_in, FC, ch, w_ok, wv);
          if (ht == 0) US_MARK(3, q);
          warp_wait(&full[st], (uint32_t)(n & 1));
          if (ht == 0) US_MARK(4, q);
          us_split_x(reinterpret_cast<const float4*>(sb), reinterpret_cast<float4*>(sb + up.xbytes), xq, ht);
          if (ht == 0) US_MARK(8, q);
          us_store_w(sb + 2 * up.xbytes, RGx, th, b0.pb, D, FC, ch, w_ok, wv);
          if (ht == 0) US_MARK(9, q);
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) mbar_arrive(&conv[st]);
          if (ht == 0) US_MARK(5, q);
        }
      } else {
        if (tid == 0) US_MARK(6, 8 * ti);
        warp_wait(zdone, (uint32_t)(ti & 1));
        umma::fence_after_sync();
        if (tid == 0) US_MARK(6, 8 * ti + 1);
        // Z1^T (lanes = outputs) -> rows in shared memory: warp 0 (lanes 0..31 = W hi rows) stores hi*hi + hi*lo
        // into the S_0 rows, warp 1 (lanes 32..63 = W lo rows) stores lo*hi + lo*lo into the dz rows of block 0
        if (warp < 2) {
          float* dstb = sm + (warp == 0 ? b0.offS : b0.offD);
          for (int rr = 0; rr < TRc; rr += 16) {
            float v1[2][8], v2[2][8];
#pragma unroll
            for (int j = 0; j < 2; ++j)
              if (rr + 8 * j < TRc) {
                umma::tmem_ld8(umma::tmem_addr(tbase, lane_base, rr + 8 * j), v1[j]);
                umma::tmem_ld8(umma::tmem_addr(tbase, lane_base, TRc + rr + 8 * j), v2[j]);
              }
            umma::tmem_ld_wait();
            if (lane < b0.out_p) {
#pragma unroll
              for (int j = 0; j < 2; ++j)
                if (rr + 8 * j < TRc) {
#pragma unroll
                  for (int i = 0; i < 8; ++i) dstb[(rr + 8 * j + i) * b0.ld_out + lane] = v1[j][i] + v2[j][i];
                }
            }
          }
        }
        umma::fence_before_sync();
      }
      rows_barrier();
      if (tid == 0) US_MARK(6, 8 * ti + 2);
      // ------------------------------------------------ tail: four threads per row, run-time loops over the row
      // buffers in shared memory (small code: it runs once per tile), __syncwarp between the layers.  All plan
      // fields are pulled into registers per layer (see ActP above).
      for (int rbase = 0; rbase < TRc; rbase += US_ROWS_PER_PASS) {
        const int t = rbase + rw * 8 + (lane >> 2);     // row of this quad
        const bool active = t < R;
        const int tt = active ? t : 0;
        const int ld0o = B0.ld_out, op0 = B0.out_p;
        float* S0 = sm + B0.offS + tt * ld0o;
        float* D0 = sm + B0.offD + tt * ld0o;
        // ---- block 0: z1 = (lo*hi + lo*lo) + (hi*hi + hi*lo); the bias came through the constant feature
        if (active) {
          for (int o = part; o < op0; o += 4) S0[o] += D0[o];
          us_act_row(actp(B0), Wt, S0, B0.offZ >= 0 ? sm + B0.offZ + tt * ld0o : nullptr, part, 4);
        }
        __syncwarp();
        if (tid == 0 && rbase == 0) US_MARK(10, 8 * ti);
        // ---- forward through blocks 1..nb-1: thread `part` owns the outputs part, part + 4, ...
        for (int l = 1; l < nb; ++l) {
          const int in_p = P.b[l].in_p, out_p = P.b[l].out_p, ld_in = P.b[l].ld_in, ld_out = P.b[l].ld_out;
          const int pw = P.b[l].pw, pb = P.b[l].pb, offS = P.b[l].offS, offZ = P.b[l].offZ, offSp = P.b[l - 1].offS;
          const ActP ap = actp(P.b[l]);
          if (active) {
            float* Sl = sm + offS + tt * ld_out;
            const int nj = (out_p - part + 3) >> 2;
            float acc[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] = j < nj ? Wt[pb + part + 4 * j] : 0.f;
            us_dot_rows_n(nj, sm + offSp + tt * ld_in, Wt + pw + part * ld_in, 4 * ld_in, in_p, acc);
#pragma unroll
            for (int j = 0; j < 8; ++j)
              if (j < nj) Sl[part + 4 * j] = acc[j];
            us_act_row(ap, Wt, Sl, offZ >= 0 ? sm + offZ + tt * ld_out : nullptr, part, 4);
          }
          __syncwarp();
        }
        if (tid == 0 && rbase == 0) US_MARK(10, 8 * ti + 1);
        // ---- likelihood residual -> dz of the last block
        {
          const int ld_out = P.b[nb - 1].ld_out, offS = P.b[nb - 1].offS, offZ = P.b[nb - 1].offZ, offD = P.b[nb - 1].offD;
          const ActP ap = actp(P.b[nb - 1]);
          const int lik = P.lik, OUT = P.OUT;
          if (active)
            stat += (double)us_lik_row(ap, lik, OUT, Wt, sm + offS + tt * ld_out,
                                       offZ >= 0 ? sm + offZ + tt * ld_out : nullptr, sm + offD + tt * ld_out,
                                       Y + (r0 + tt) * (long long)OUT, part, 4);
        }
        __syncwarp();
        if (tid == 0 && rbase == 0) US_MARK(10, 8 * ti + 2);
        // ---- data gradient back to block 0: dz_{l-1}[k] = (sum_o W_l[o][k] dz_l[o]) * act'_{l-1};
        //      thread `part` owns k = part, part + 4, ... and reads the transposed weights WT_l[k][o]
        for (int l = nb - 1; l >= 1; --l) {
          const int in_p = P.b[l].in_p, out_p = P.b[l].out_p, ld_out = P.b[l].ld_out, offD = P.b[l].offD;
          const int ldp = P.b[l - 1].ld_out, offDp = P.b[l - 1].offD, offSp = P.b[l - 1].offS, offZp = P.b[l - 1].offZ;
          const ActP app = actp(P.b[l - 1]);
          const int wofs = up.wtt_ofs[l];
          if (active) {
            float* Dp = sm + offDp + tt * ldp;
            const int nj = (in_p - part + 3) >> 2;
            float da[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) da[j] = 0.f;
            us_dot_rows_n(nj, sm + offD + tt * ld_out, WT + wofs + part * ld_out, 4 * ld_out, out_p, da);
#pragma unroll
            for (int j = 0; j < 8; ++j)
              if (j < nj) Dp[part + 4 * j] = da[j];
            us_dact_row(app, Wt, Dp, sm + offSp + tt * ldp, offZp >= 0 ? sm + offZp + tt * ldp : nullptr, part, 4);
          }
          __syncwarp();
        }
        if (tid == 0 && rbase == 0) US_MARK(10, 8 * ti + 3);
        // ---- dZ1 as the stacked K-major A operand of the backward MMAs: rows o (hi) and 32 + o (lo), column = t
        if (t < TRc) {
          const uint32_t cb = (uint32_t)(t >> 2) * (uint32_t)US_DZ_CG + (uint32_t)(t & 3) * 4u;
          const int out0 = B0.out;
          for (int o = part; o < op0; o += 4) {
            float h, l;
            umma::split_tf32((active && o < out0) ? D0[o] : 0.f, h, l);
            const uint32_t off = cb + (uint32_t)(o >> 3) * 128u + (uint32_t)(o & 7) * 16u;
            *reinterpret_cast<float*>(dzop + off) = h;
            *reinterpret_cast<float*>(dzop + off + 512u) = l;      // row 32 + o: four row groups further
          }
        }
      }
      if (tid == 0) US_MARK(6, 8 * ti + 3);
      fence_proxy_async();
      rows_barrier();
      if (tid == 0) mbar_arrive(dzready);
      if (tid == 0) US_MARK(6, 8 * ti + 4);
      // ------------------------------------------------ backward pass
      if (hybrid) {
        for (int ch = 0; ch < nch; ++ch, ++q) {
          const int st = q % US_NSTAGE, n = q / US_NSTAGE;
          unsigned char* sb = smraw + up.off_stage + st * up.stage_bytes;
          if (ht == 0) US_MARK(3, q);
          warp_wait(&full[st], (uint32_t)(n & 1));
          if (ht == 0) US_MARK(4, q);
          us_split_x(reinterpret_cast<const float4*>(sb), reinterpret_cast<float4*>(sb + up.xbytes), xq, ht);
          if (ht == 0) US_MARK(8, q);
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) mbar_arrive(&conv[st]);
          if (ht == 0) US_MARK(5, q);
        }
      } else if (warp >= 2) {
        // ---- gradients of everything except W1 / b1 from the row buffers of this tile (64 threads)
        const int t2 = tid - 64;
        if (nb > 1) narrow_accum<float>(P, Wt, G, sm, R, t2, 64);
        if (act_has_slopes(b0.act)) {
          for (int o = t2; o < b0.out; o += 64) {
            float sc = 0.f;
            for (int r = 0; r < R; ++r) sc += sm[b0.offZ + r * b0.ld_out + o];
            const float f = b0.act == ACT_SQPRELU ? 2.f * Wt[b0.ps + o] : 1.f;
            G[b0.ps + o] += f * sc;
          }
        }
        if (tid == 64) US_MARK(6, 8 * ti + 5);
      } else {
        // ---- drain the dW1 chunks: lanes 0..31 = dZhi rows (x XThi in columns [0, FC), x XTlo in [FC, 2 FC)),
        //      lanes 32..63 = dZlo rows.  dW1[o][f] = (lo*hi + lo*lo) + (hi*lo + hi*hi), written to the partial
        const int DWS = FC + 4;                       // row stride of the staging tiles (floats): conflict-free
        float* dws = dwx + warp * 32 * DWS;           // this warp's tile [32 outputs][DWS]
        const int fq = FC >> 2, nvec = b0.out_p * fq; // float4 per chunk after the two tiles are summed
        for (int ch = 0; ch < nch; ++ch, ++g) {
          const int bsel = g & 1, u = g >> 1;
          if (tid == 0 && g == 1) US_MARK(11, 14);
          warp_wait(&dwdone[bsel], (uint32_t)(u & 1));
          umma::fence_after_sync();
          if (tid == 0 && g == 1) US_MARK(11, 15);
          const uint32_t col0 = ZC + (uint32_t)bsel * DWC;
          // phase A: TMEM -> shared memory, the two column halves (x XThi, x XTlo) summed on the way
          for (int f0 = 0; f0 < FC; f0 += 16) {
            float v1[2][8], v2[2][8];
#pragma unroll
            for (int j = 0; j < 2; ++j)
              if (f0 + 8 * j < FC) {
                umma::tmem_ld8(umma::tmem_addr(tbase, lane_base, col0 + f0 + 8 * j), v1[j]);
                umma::tmem_ld8(umma::tmem_addr(tbase, lane_base, col0 + FC + f0 + 8 * j), v2[j]);
              }
            umma::tmem_ld_wait();
            if (lane < b0.out_p) {
#pragma unroll
              for (int j = 0; j < 2; ++j)
                if (f0 + 8 * j < FC) {
                  float* dp = dws + lane * DWS + f0 + 8 * j;
                  *reinterpret_cast<float4*>(dp) = make_float4(v1[j][0] + v2[j][0], v1[j][1] + v2[j][1],
                                                               v1[j][2] + v2[j][2], v1[j][3] + v2[j][3]);
                  *reinterpret_cast<float4*>(dp + 4) = make_float4(v1[j][4] + v2[j][4], v1[j][5] + v2[j][5],
                                                                   v1[j][6] + v2[j][6], v1[j][7] + v2[j][7]);
                }
            }
          }
          if (tid == 0 && g == 1) US_MARK(11, 16);
          umma::fence_before_sync();
          drain_barrier();
          if (tid == 0) mbar_arrive(&dwfree[bsel]);       // tensor memory is free again; the rest is shared -> global
          // phase B: dW1[o][f] = (lo rows) + (hi rows), 64 threads, consecutive threads = consecutive features
          for (int e = tid; e < nvec; e += 64) {
            const int o = e / fq, qq = e - o * fq;
            const float4 h4 = *reinterpret_cast<const float4*>(dwx + o * DWS + 4 * qq);
            const float4 l4 = *reinterpret_cast<const float4*>(dwx + (32 + o) * DWS + 4 * qq);
            float r[4] = {l4.x + h4.x, l4.y + h4.y, l4.z + h4.z, l4.w + h4.w};
            const int f = ch * FC + 4 * qq;
            float* dst = out + b0.pw + o * b0.ld_in + f;
            if (f + 3 < D) {
              if (ti > 0) {
                const float4 o1 = *reinterpret_cast<const float4*>(dst);
                r[0] += o1.x; r[1] += o1.y; r[2] += o1.z; r[3] += o1.w;
              }
              *reinterpret_cast<float4*>(dst) = make_float4(r[0], r[1], r[2], r[3]);
            } else {
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                if (f + i < D) dst[i] = ti > 0 ? dst[i] + r[i] : r[i];
                else if (f + i == D) out[b0.pb + o] = ti > 0 ? out[b0.pb + o] + r[i] : r[i];
              }
            }
          }
          if (tid == 0 && g == 1) US_MARK(11, 17);
          drain_barrier();                                 // the staging tiles are rewritten by the next chunk
          if (tid == 0) US_MARK(7, g);
        }
      }
      // the row buffers are rewritten by the next tile: the TMEM warps (drains, accumulation) must be done
      if (!hybrid) tail_barrier();
      if (tid == 0) US_MARK(6, 8 * ti + 6);
    }
    // ---- likelihood statistic of the CTA, the other gradients, zero padding of the W1 rows
    const double ws = warp_sum(stat);
    if (lane == 0) red[rw] = ws;
    rows_barrier();
    if (tid == 0) {
      double tot = 0.0;
      for (int w = 0; w < US_ROW_WARPS; ++w) tot += red[w];
      stat_part[(size_t)c * S + s] = tot;
    }
    if (!hybrid) {
      for (int e = tid; e < b0.out_p * (b0.ld_in - D); e += 128) {
        const int o = e / (b0.ld_in - D), k = D + e - o * (b0.ld_in - D);
        out[b0.pw + o * b0.ld_in + k] = 0.f;
      }
      for (int i = b0.pb + b0.out_p + tid; i < mp.Ppad; i += 128) out[i] = G[i];
    }
  }
  umma::fence_before_sync();
  __syncthreads();
  if (tid == 0) US_MARK(6, 63);
  if (warp == 0) umma::tmem_dealloc(tbase, 512);
#undef US_MARK
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
  int wtt = 0;
  for (int l = 1; l < mp.nb; ++l) { up.wtt_ofs[l] = wtt; wtt += mp.b[l].in_p * mp.b[l].ld_out; }
  wtt = (wtt + 3) & ~3;
  const int fixed = buf_floats * 4 + (2 * tailp + wtt) * 4 + up.dzbytes + 16 * 8 + 16 * 8 + 512;
  // chunk width: as wide as two stages, the drain buffer and the 512 TMEM columns allow, then balanced
  const int feats = mp.D + 1;
  int best = 0;
  for (int FC = 128; FC >= 8; FC -= 8) {
    const long long stage = 2LL * TRc * FC * 4 + 64LL * FC * 4;
    if (US_NSTAGE * stage + 64LL * (FC + 4) * 4 + fixed <= (long long)smem_limit && 2 * TRc + 4 * FC <= 512) { best = FC; break; }
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
  up.off_wtt = off; off += wtt * 4;
  up.off_dwx = off; off += 64 * (up.FC + 4) * 4;
  off = (off + 127) & ~127;
  up.off_stage = off; off += US_NSTAGE * up.stage_bytes;
  up.off_dz = off; off += up.dzbytes;
  off = (off + 15) & ~15;
  // the M = 128 MMAs read 16 row groups of the stacked W1 operand (only 8 exist; rows >= 64 are never used):
  // keep those reads inside the allocation
  const int last_w = up.off_stage + (US_NSTAGE - 1) * up.stage_bytes + 2 * up.xbytes;
  off = std::max(off, last_w + 16 * up.FC * 32);
  up.off_red = off; off += 16 * 8 + 16 * 8;
  up.smem_bytes = (off + 15) & ~15;
  wp.smem_elems = up.smem_bytes / 4;
  if ((size_t)up.smem_bytes > smem_limit) return false;
  if (up.NP * (up.FC / 4) > US_WREG * US_HYB_THREADS) return false;
  return true;
}

// two tiled copies: row-major cores (forward pass) and feature-major cores (backward pass)
size_t usweep_xt_bytes(const USweepPlan& up) { return 2 * (size_t)up.ntiles * up.nch * up.xbytes; }

void launch_tile_x(const USweepPlan& up, int D, const float* X, long long N, float* Xt, cudaStream_t st) {
  dim3 g(up.ntiles, up.nch);
  k_tile_x<<<g, 256, 0, st>>>(up, D, X, N, Xt);
}

static void launch_us(const ModelPlan& wp, const USweepPlan& up, dim3 g, const float* theta_pad, const float* Xt,
                      const float* Y, long long N, float* partial, double* stat_part, cudaStream_t st) {
  cudaFuncSetAttribute(k_sweep_umma, cudaFuncAttributeMaxDynamicSharedMemorySize, up.smem_bytes);
  static const bool want_prof = getenv("TBNN_US_PROF") != nullptr;
  static long long* dprof = nullptr;
  if (want_prof && !dprof) cudaMalloc(&dprof, US_PROF_ROLES * US_PROF_N * sizeof(long long));
  if (want_prof) cudaMemsetAsync(dprof, 0, US_PROF_ROLES * US_PROF_N * sizeof(long long), st);
  k_sweep_umma<<<g, US_THREADS, up.smem_bytes, st>>>(wp, up, (int)g.x, theta_pad, Xt, Y, N, partial, stat_part,
                                                           want_prof ? dprof : nullptr);
  if (want_prof) {
    // developer aid: clock64 marks of CTA 0, relative to the first producer mark
    static int shown = 0;
    std::vector<long long> h(US_PROF_ROLES * US_PROF_N);
    cudaStreamSynchronize(st);
    cudaMemcpy(h.data(), dprof, h.size() * sizeof(long long), cudaMemcpyDeviceToHost);
    if (shown++ == 3) {
      const char* names[US_PROF_ROLES] = {"tma issue", "mma start", "mma issued", "conv W ready", "conv X landed",
                                          "conv done", "tail marks", "drain done", "conv X split", "conv W stored",
                                          "tail fine", "drain 1st ld"};
      const long long t0 = h[0];
      for (int r = 0; r < US_PROF_ROLES; ++r) {
        fprintf(stderr, "[us_prof] %-14s", names[r]);
        for (int i = 0; i < US_PROF_N; ++i) if (h[r * US_PROF_N + i]) fprintf(stderr, " %d:%lld", i, h[r * US_PROF_N + i] - t0);
        fprintf(stderr, "\n");
      }
    }
  }
}

void launch_sweep_umma(const ModelPlan& wp, const USweepPlan& up, int C, int S, const float* theta_pad,
                       const float* Xt, const float* Y, long long N, float* partial, double* stat_part,
                       cudaStream_t st) {
  dim3 g(S, C);
  launch_us(wp, up, g, theta_pad, Xt, Y, N, partial, stat_part, st);
}

}  // namespace tbnn
