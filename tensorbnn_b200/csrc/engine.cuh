// engine.cuh -- the row-tile engine: fused forward + likelihood + backward of one
// tile of training rows for one chain, entirely in shared memory (FP32 FFMA or FP64).
//
// Replaces, for one tile, the TF ops of network.predict (network.py:141-171),
// layer.predict (layer.py:266-279), the activations (activationFunctions.py),
// the likelihood residuals (likelihood.py:88-94,162-167,225-236) and TF's reverse-mode
// autodiff of all of them (invoked by TFP's leapfrog from network.py:394-408).
//
// Data layout in shared memory (T elements; all row starts 16-byte aligned):
//   X tile / activations  [TR][ld]   row = training row, K contiguous, ld/4 odd so that
//                                    float4 loads of 8 consecutive rows hit 8 distinct banks
//   W_l                   [out_p][ld_in]  (padded copy of theta; K contiguous)
//   G                     same padded layout as theta: gradient accumulators
// Thread tiles are 4x4 with interleaved rows (r = rg + i*tm) so a quarter-warp reads
// conflict-free; operands that are shared inside a quarter-warp are broadcast.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

#include "plan.h"

namespace tbnn {

// ---------------------------------------------------------------- small helpers
template <typename T> __device__ __forceinline__ void ld4(const T* p, T (&v)[4]);
template <> __device__ __forceinline__ void ld4<float>(const float* p, float (&v)[4]) {
  const float4 t = *reinterpret_cast<const float4*>(p);
  v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
template <> __device__ __forceinline__ void ld4<double>(const double* p, double (&v)[4]) {
  const double2 a = *reinterpret_cast<const double2*>(p);
  const double2 b = *reinterpret_cast<const double2*>(p + 2);
  v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
}
template <typename T> __device__ __forceinline__ void st4(T* p, const T (&v)[4]);
template <> __device__ __forceinline__ void st4<float>(float* p, const float (&v)[4]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
template <> __device__ __forceinline__ void st4<double>(double* p, const double (&v)[4]) {
  *reinterpret_cast<double2*>(p) = make_double2(v[0], v[1]);
  *reinterpret_cast<double2*>(p + 2) = make_double2(v[2], v[3]);
}

__device__ __forceinline__ float t_exp(float x) { return expf(x); }
__device__ __forceinline__ double t_exp(double x) { return exp(x); }
__device__ __forceinline__ float t_expm1(float x) { return expm1f(x); }
__device__ __forceinline__ double t_expm1(double x) { return expm1(x); }
__device__ __forceinline__ float t_tanh(float x) { return tanhf(x); }
__device__ __forceinline__ double t_tanh(double x) { return tanh(x); }
__device__ __forceinline__ float t_log(float x) { return logf(x); }
__device__ __forceinline__ double t_log(double x) { return log(x); }
__device__ __forceinline__ float t_log1p(float x) { return log1pf(x); }
__device__ __forceinline__ double t_log1p(double x) { return log1p(x); }

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Sum over the CTA; result valid in every thread.  `red` holds >= 33 doubles.
__device__ __forceinline__ double block_sum(double v, double* red) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  if (w == 0) {
    double x = lane < nw ? red[lane] : 0.0;
    x = warp_sum(x);
    if (lane == 0) red[32] = x;
  }
  __syncthreads();
  return red[32];
}

// ---------------------------------------------------------------- activations
// effective negative-side slope of a z-keeping activation
template <typename T>
__device__ __forceinline__ T eff_slope(int act, const T* slopes, int o, T alpha) {
  if (act == ACT_PRELU) return slopes[o];
  if (act == ACT_SQPRELU) { const T s = slopes[o]; return s * s; }
  return alpha;  // ACT_LEAKY
}

template <typename T> __device__ __forceinline__ T act_fwd(int act, T z, T slope) {
  switch (act) {
    case ACT_NONE: return z;
    case ACT_RELU: return z > T(0) ? z : T(0);                 // activationFunctions.py:36
    case ACT_TANH: return t_tanh(z);                           // :62
    case ACT_SIGMOID: return T(1) / (T(1) + t_exp(-z));        // :49
    case ACT_EXP: return t_exp(z);                             // :23
    case ACT_ELU: return z > T(0) ? z : t_expm1(z);            // :75
    default: return z < T(0) ? slope * z : z;                  // :105, :250-254, :412-416
  }
}

// derivative of a parameter-free activation from its OUTPUT a
template <typename T> __device__ __forceinline__ T act_deriv_from_out(int act, T a) {
  switch (act) {
    case ACT_NONE: return T(1);
    case ACT_RELU: return a > T(0) ? T(1) : T(0);
    case ACT_TANH: return T(1) - a * a;
    case ACT_SIGMOID: return a * (T(1) - a);
    case ACT_EXP: return a;
    case ACT_ELU: return a < T(0) ? a + T(1) : T(1);
    default: return T(1);
  }
}

// ---------------------------------------------------------------- tile context
template <typename T> struct TileCtx {
  T* sm;          // base of dynamic shared memory
  const T* Wp;    // padded parameters of this chain (shared or global memory)
  T* G;           // padded gradient accumulators (shared memory)
};

// Copy rows [row0,row0+nr) of X[N][D] into the X tile; zero-fill padding rows/cols.
template <typename T>
__device__ __forceinline__ void load_x_tile(const ModelPlan& mp, T* Xs, const T* __restrict__ X,
                                            long long row0, int nr) {
  const int D = mp.D, ld0 = mp.ld0, D_p = mp.D_p, TR = mp.TR;
  constexpr int VEC = 16 / sizeof(T);
  if ((D % VEC) == 0 && ((reinterpret_cast<uintptr_t>(X) & 15) == 0)) {
    const int cpr = D / VEC;                       // 16-byte chunks per row
    const int total = nr * cpr;
    for (int c = threadIdx.x; c < total; c += blockDim.x) {
      const int r = c / cpr, j = c - r * cpr;
      const T* src = X + (row0 + r) * (long long)D + j * VEC;
      const unsigned dst = (unsigned)__cvta_generic_to_shared(Xs + r * ld0 + j * VEC);
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dst), "l"(src) : "memory");
    }
    asm volatile("cp.async.commit_group;\n" ::: "memory");
    for (int c = threadIdx.x; c < nr * (D_p - D); c += blockDim.x) {
      const int r = c / (D_p - D), j = c - r * (D_p - D);
      Xs[r * ld0 + D + j] = T(0);
    }
  } else {
    for (int c = threadIdx.x; c < nr * D_p; c += blockDim.x) {
      const int r = c / D_p, j = c - r * D_p;
      Xs[r * ld0 + j] = j < D ? X[(row0 + r) * (long long)D + j] : T(0);
    }
  }
  for (int c = threadIdx.x; c < (TR - nr) * D_p; c += blockDim.x) {
    const int r = nr + c / D_p, j = c % D_p;
    Xs[r * ld0 + j] = T(0);
  }
}

__device__ __forceinline__ void wait_x_tile() {
  asm volatile("cp.async.wait_all;\n" ::: "memory");
}

// ---------------------------------------------------------------- forward of one block
template <typename T>
__device__ __forceinline__ void fwd_store(const BlockPlan& b, const T* Wp, T* S, T* Z, int r, int o,
                                          T acc) {
  T a = T(0), z = T(0);
  if (o < b.out) {
    z = acc + Wp[b.pb + o];
    T slope = T(0);
    if (act_keeps_z(b.act)) slope = eff_slope<T>(b.act, Wp + (b.ps >= 0 ? b.ps : 0), o, T(b.alpha));
    a = act_fwd<T>(b.act, z, slope);
  }
  S[r * b.ld_out + o] = a;
  if (Z) Z[r * b.ld_out + o] = z;
}

// the same with the bias and the effective slope of output o already in registers (one load per output column of a
// thread tile instead of one per accumulator)
template <typename T>
__device__ __forceinline__ void fwd_store_pre(const BlockPlan& b, T* S, T* Z, int r, int o, T acc, T bias, T slope) {
  T a = T(0), z = T(0);
  if (o < b.out) {
    z = acc + bias;
    // the two common kinds inline (uniform branches), the rest through the generic switch
    if (act_keeps_z(b.act)) a = z < T(0) ? slope * z : z;
    else if (b.act == ACT_RELU) a = z > T(0) ? z : T(0);
    else a = act_fwd<T>(b.act, z, slope);
  }
  S[r * b.ld_out + o] = a;
  if (Z) Z[r * b.ld_out + o] = z;
}
template <typename T>
__device__ __forceinline__ void load_bias_slope(const BlockPlan& b, const T* Wp, int o, T& bias, T& slope) {
  bias = T(0); slope = T(0);
  if (o < b.out) {
    bias = Wp[b.pb + o];
    if (act_keeps_z(b.act)) slope = eff_slope<T>(b.act, Wp + (b.ps >= 0 ? b.ps : 0), o, T(b.alpha));
  }
}

// S_l[r][o] = act( sum_k A[r][k] W[o][k] + b[o] )       (layer.py:276-279 + activation)
template <typename T>
__device__ void fwd_block(const ModelPlan& mp, int l, const TileCtx<T>& cx) {
  const BlockPlan& b = mp.b[l];
  const T* A = cx.sm + (l == 0 ? mp.offX : mp.b[l - 1].offS);
  const int lda = b.ld_in;
  const T* W = cx.Wp + b.pw;
  T* S = cx.sm + b.offS;
  T* Z = b.offZ >= 0 ? cx.sm + b.offZ : nullptr;
  const int tm = mp.TR >> 2, tn = b.out_p >> 2, ntile = tm * tn, kch = b.in_p >> 2;
  const int ksplit = b.ksplit;
  if (ksplit <= 1) {
    for (int t = threadIdx.x; t < ntile; t += blockDim.x) {
      const int rg = t % tm, og = t / tm;
      T acc[4][4];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = T(0);
      const T* a0 = A + rg * lda;
      const T* w0 = W + og * lda;
#pragma unroll 4
      for (int kc = 0; kc < kch; ++kc) {
        T av[4][4], wv[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i) ld4(a0 + (i * tm) * lda + 4 * kc, av[i]);
#pragma unroll
        for (int j = 0; j < 4; ++j) ld4(w0 + (j * tn) * lda + 4 * kc, wv[j]);
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int q = 0; q < 4; ++q) acc[i][j] = fma(av[i][q], wv[j][q], acc[i][j]);
      }
      T bias[4], slope[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) load_bias_slope<T>(b, cx.Wp, og + j * tn, bias[j], slope[j]);
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) fwd_store_pre<T>(b, S, Z, rg + i * tm, og + j * tn, acc[i][j], bias[j], slope[j]);
    }
    __syncthreads();
  } else {
    T* scr = cx.sm + mp.offScr;
    const int tile = threadIdx.x % ntile, ks = threadIdx.x / ntile;
    if (ks < ksplit) {
      const int rg = tile % tm, og = tile / tm;
      T acc[4][4];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = T(0);
      const T* a0 = A + rg * lda;
      const T* w0 = W + og * lda;
      for (int kc = ks; kc < kch; kc += ksplit) {
        T av[4][4], wv[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i) ld4(a0 + (i * tm) * lda + 4 * kc, av[i]);
#pragma unroll
        for (int j = 0; j < 4; ++j) ld4(w0 + (j * tn) * lda + 4 * kc, wv[j]);
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int q = 0; q < 4; ++q) acc[i][j] = fma(av[i][q], wv[j][q], acc[i][j]);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) scr[(ks * 16 + i * 4 + j) * ntile + tile] = acc[i][j];
    }
    __syncthreads();
    for (int e = threadIdx.x; e < 16 * ntile; e += blockDim.x) {
      const int ij = e / ntile, tl = e - ij * ntile;
      T s = T(0);
      for (int k2 = 0; k2 < ksplit; ++k2) s += scr[(k2 * 16 + ij) * ntile + tl];
      const int rg = tl % tm, og = tl / tm;
      fwd_store<T>(b, cx.Wp, S, Z, rg + (ij >> 2) * tm, og + (ij & 3) * tn, s);
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------- likelihood phase
// Residuals of the output layer; writes dZ_K (gradient w.r.t. the last pre-activation,
// up to the Gaussian 1/sigma^2 factor that is applied later) and returns this thread's
// contribution to the statistic (SSE for Gaussian kinds, log-likelihood for Bernoulli).
template <typename T>
__device__ T lik_phase(const ModelPlan& mp, const TileCtx<T>& cx, const T* __restrict__ Y,
                       long long row0, int nr, T* dZ) {
  const BlockPlan& b = mp.b[mp.nb - 1];
  const T* S = cx.sm + b.offS;
  T* Z = b.offZ >= 0 ? cx.sm + b.offZ : nullptr;
  const int ld = b.ld_out, outp = b.out_p, OUT = mp.OUT;
  T stat = T(0);
  const T lo = T(1e-8), hi = T(1 - 1e-7);          // likelihood.py:229-230, cast to dtype
  for (int e = threadIdx.x; e < mp.TR * outp; e += blockDim.x) {
    const int r = e / outp, o = e - r * outp;
    T dz = T(0), c = T(0);
    if (r < nr && o < OUT) {
      const T f = S[r * ld + o];
      const T y = Y[(row0 + r) * (long long)OUT + o];
      T df;
      if (mp.lik == LIK_BERN) {
        const T p = f < lo ? lo : (f > hi ? hi : f);
        stat += (T(1) - y) * t_log1p(-p) + y * t_log(p);
        df = (f < lo || f > hi) ? T(0) : (y / p - (T(1) - y) / (T(1) - p));
      } else {
        const T res = y - f;
        stat = fma(res, res, stat);
        df = res;
      }
      if (act_keeps_z(b.act)) {
        const T z = Z[r * ld + o];
        const bool neg = z < T(0);
        const T s = eff_slope<T>(b.act, cx.Wp + (b.ps >= 0 ? b.ps : 0), o, T(b.alpha));
        dz = neg ? df * s : df;
        c = neg ? z * df : T(0);
      } else {
        dz = df * act_deriv_from_out<T>(b.act, f);
      }
    }
    dZ[r * ld + o] = dz;
    if (Z && act_has_slopes(b.act)) Z[r * ld + o] = c;
  }
  __syncthreads();
  return stat;
}

// ---------------------------------------------------------------- backward of one block
// Given dZ_l: G.W_l += dZ_l^T A_{l-1};  G.b_l += colsum(dZ_l);  slope gradient of block l
// (from the contributions left in Z_l);  dZ_{l-1} = (dZ_l W_l) * act'_{l-1}.
template <typename T>
__device__ void bwd_block(const ModelPlan& mp, int l, const TileCtx<T>& cx, const T* dZ, T* dNext) {
  const BlockPlan& b = mp.b[l];
  const T* A = cx.sm + (l == 0 ? mp.offX : mp.b[l - 1].offS);
  const int lda = b.ld_in, ldz = b.ld_out, TR = mp.TR;
  const T* W = cx.Wp + b.pw;
  // (a) weight gradient: 4(o) x 4(k) tiles, reduce over the tile's rows
  {
    T* Gw = cx.G + b.pw;
    const int tn = b.out_p >> 2, tk = b.in_p >> 2, ntile = tn * tk;
    for (int t = threadIdx.x; t < ntile; t += blockDim.x) {
      const int og = t % tn, kg = t / tn;
      T acc[4][4];
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[j][q] = T(0);
      const T* dz0 = dZ + 4 * og;
      const T* a0 = A + 4 * kg;
#pragma unroll 8
      for (int r = 0; r < TR; ++r) {
        T dv[4], av[4];
        ld4(dz0 + r * ldz, dv);
        ld4(a0 + r * lda, av);
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
          for (int q = 0; q < 4; ++q) acc[j][q] = fma(dv[j], av[q], acc[j][q]);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        T g[4];
        T* gp = Gw + (4 * og + j) * lda + 4 * kg;
        ld4(gp, g);
#pragma unroll
        for (int q = 0; q < 4; ++q) g[q] += acc[j][q];
        st4(gp, g);
      }
    }
  }
  // (b) bias gradient and (d) slope gradient: column sums
  {
    for (int o = threadIdx.x; o < b.out_p; o += blockDim.x) {
      T s = T(0);
      for (int r = 0; r < TR; ++r) s += dZ[r * ldz + o];
      cx.G[b.pb + o] += s;
    }
    if (act_has_slopes(b.act)) {
      const T* Zc = cx.sm + b.offZ;
      // second half of the CTA so it overlaps with the bias sums
      for (int o = (int)blockDim.x - 1 - (int)threadIdx.x; o < b.out_p; o += blockDim.x) {
        T s = T(0);
        for (int r = 0; r < TR; ++r) s += Zc[r * ldz + o];
        const T f = b.act == ACT_SQPRELU ? T(2) * cx.Wp[b.ps + o] : T(1);
        cx.G[b.ps + o] += f * s;
      }
    }
  }
  // (c) data gradient into the previous block
  if (l > 0) {
    const BlockPlan& pb = mp.b[l - 1];
    const T* Sp = cx.sm + pb.offS;
    T* Zp = pb.offZ >= 0 ? cx.sm + pb.offZ : nullptr;
    const int tm = TR >> 2, tk = b.in_p >> 2, ntile = tm * tk, och = b.out_p >> 2;
    const bool keepz = act_keeps_z(pb.act), hass = act_has_slopes(pb.act);
    for (int t = threadIdx.x; t < ntile; t += blockDim.x) {
      const int rg = t % tm, kg = t / tm;
      T acc[4][4];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[i][q] = T(0);
      const T* dz0 = dZ + rg * ldz;
      const T* w0 = W + 4 * kg;
#pragma unroll 4
      for (int oc = 0; oc < och; ++oc) {
        T dv[4][4], wv[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i) ld4(dz0 + (i * tm) * ldz + 4 * oc, dv[i]);
#pragma unroll
        for (int j = 0; j < 4; ++j) ld4(w0 + (4 * oc + j) * lda, wv[j]);
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int q = 0; q < 4; ++q) acc[i][q] = fma(dv[i][j], wv[j][q], acc[i][q]);
      }
      T sl4[4] = {T(0), T(0), T(0), T(0)};
      if (keepz) {
#pragma unroll
        for (int q = 0; q < 4; ++q) sl4[q] = eff_slope<T>(pb.act, cx.Wp + (pb.ps >= 0 ? pb.ps : 0), 4 * kg + q, T(pb.alpha));
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int r = rg + i * tm;
        T dzv[4];
        if (keepz) {
          T zv[4], cv[4];
          ld4(Zp + r * lda + 4 * kg, zv);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const bool neg = zv[q] < T(0);
            const T s = sl4[q];
            dzv[q] = neg ? acc[i][q] * s : acc[i][q];
            cv[q] = neg ? zv[q] * acc[i][q] : T(0);
          }
          if (hass) st4(Zp + r * lda + 4 * kg, cv);
        } else {
          T av[4];
          ld4(Sp + r * lda + 4 * kg, av);
#pragma unroll
          for (int q = 0; q < 4; ++q)
            dzv[q] = pb.act == ACT_RELU ? (av[q] > T(0) ? acc[i][q] : T(0)) : acc[i][q] * act_deriv_from_out<T>(pb.act, av[q]);
        }
        st4(dNext + r * lda + 4 * kg, dzv);
      }
    }
  }
  __syncthreads();
}

// ---------------------------------------------------------------- one tile, end to end
// Forward only (predictor / display metrics): leaves f = S_K in shared memory.
template <typename T>
__device__ __forceinline__ void tile_forward(const ModelPlan& mp, const TileCtx<T>& cx) {
  for (int l = 0; l < mp.nb; ++l) fwd_block<T>(mp, l, cx);
}

// Forward + likelihood + backward.  The X tile must already be in shared memory.
template <typename T>
__device__ __forceinline__ T tile_forward_backward(const ModelPlan& mp, const TileCtx<T>& cx,
                                                   const T* __restrict__ Y, long long row0, int nr) {
  tile_forward<T>(mp, cx);
  T* dA = cx.sm + mp.offDa;
  T* dB = cx.sm + mp.offDb;
  const T stat = lik_phase<T>(mp, cx, Y, row0, nr, dA);
  T* cur = dA;
  T* nxt = dB;
  for (int l = mp.nb - 1; l >= 0; --l) {
    bwd_block<T>(mp, l, cx, cur, nxt);
    T* t = cur; cur = nxt; nxt = t;
  }
  return stat;
}

}  // namespace tbnn
