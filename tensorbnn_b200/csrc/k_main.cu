// k_main.cu -- main-chain kernels: padded layout conversion, per-CTA likelihood partial
// gradients (the row sweep), gradient assembly + priors + leapfrog update, momentum draw,
// Metropolis-Hastings select.  Reference call sites: network.py:370-392 (target),
// :394-411 (HMC transition); TFP leapfrog / MH semantics per SURVEY.md Appendix B.
#include "engine.cuh"
#include "narrow.cuh"
#include "kernels.h"
#include "philox.cuh"

namespace tbnn {

// ------------------------------------------------------------------ element decode
struct Elem { int kind; int blk; int flat; int first; };  // kind: 0 pad, 1 W, 2 b, 3 slope
__device__ __forceinline__ Elem decode_elem(const ModelPlan& mp, int i) {
  Elem e; e.kind = 0; e.blk = 0; e.flat = 0; e.first = 0;
  for (int l = 0; l < mp.nb; ++l) {
    const BlockPlan& b = mp.b[l];
    if (i < b.pb) {
      const int j = i - b.pw, o = j / b.ld_in, k = j - o * b.ld_in;
      e.blk = l;
      if (o < b.out && k < b.in) { e.kind = 1; e.flat = b.fw + o * b.in + k; e.first = (j == 0); }
      return e;
    }
    if (i < b.pb + b.out_p) {
      const int o = i - b.pb;
      e.blk = l;
      if (o < b.out) { e.kind = 2; e.flat = b.fb + o; e.first = (o == 0); }
      return e;
    }
    if (b.ps >= 0 && i < b.ps + b.out_p) {
      const int o = i - b.ps;
      e.blk = l;
      if (o < b.out) { e.kind = 3; e.flat = b.fs + o; e.first = (o == 0); }
      return e;
    }
  }
  return e;
}

// Index of padded element i inside the pair-interleaved copy of W_1 that the warp-specialised wide sweep
// reads (k_wide2.cu: [k quad][output pair][k in quad][2], w1p_quad_stride() floats per k quad); -1 if i is
// not an element of W_1.
__device__ __forceinline__ int w1p_index(const ModelPlan& mp, int i) {
  const BlockPlan& b = mp.b[0];
  if (i >= b.pb) return -1;
  const int o = i / b.ld_in, k = i - o * b.ld_in;
  if (k >= b.in_p) return -1;
  return (k >> 2) * w1p_quad_stride(b.out_p) + (o >> 1) * 8 + (k & 3) * 2 + (o & 1);
}

template <typename T>
__global__ void k_pad(const __grid_constant__ ModelPlan mp, const T* __restrict__ flat,
                      T* __restrict__ padded, T* __restrict__ w1p) {
  const int c = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= mp.Ppad) return;
  const Elem e = decode_elem(mp, i);
  const T v = e.kind ? flat[(size_t)c * mp.P + e.flat] : T(0);
  padded[(size_t)c * mp.Ppad + i] = v;
  if (w1p) {
    const int q = w1p_index(mp, i);
    if (q >= 0) w1p[(size_t)c * w1p_elems(mp) + q] = v;
  }
}
template <typename T>
__global__ void k_unpad(const __grid_constant__ ModelPlan mp, const T* __restrict__ padded,
                        T* __restrict__ flat) {
  const int c = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= mp.Ppad) return;
  const Elem e = decode_elem(mp, i);
  if (e.kind) flat[(size_t)c * mp.P + e.flat] = padded[(size_t)c * mp.Ppad + i];
}

// ------------------------------------------------------------------ row sweep
template <typename T, bool BWD>
__global__ void __launch_bounds__(NT, 1)
k_partial(const __grid_constant__ ModelPlan mp, int S, const T* __restrict__ theta_pad,
          const T* __restrict__ X, const T* __restrict__ Y, long long N, T* __restrict__ partial,
          double* __restrict__ stat_part) {
  extern __shared__ __align__(16) unsigned char smraw[];
  T* sm = reinterpret_cast<T*>(smraw);
  const int c = blockIdx.y, s = blockIdx.x;
  const T* thg = theta_pad + (size_t)c * mp.Ppad;
  TileCtx<T> cx;
  cx.sm = sm;
  // accumulators: shared memory, or (huge layers) this CTA's slice of the global partial buffer
  cx.G = mp.offG >= 0 ? sm + mp.offG : partial + ((size_t)c * S + s) * mp.Ppad;
  if (mp.offW >= 0) {
    T* Ws = sm + mp.offW;
    for (int i = 4 * threadIdx.x; i < mp.Ppad; i += 4 * blockDim.x) {
      T v[4];
      ld4(thg + i, v);
      st4(Ws + i, v);
    }
    cx.Wp = Ws;
  } else {
    cx.Wp = thg;
  }
  if (BWD)
    for (int i = threadIdx.x; i < mp.Ppad; i += blockDim.x) cx.G[i] = T(0);
  __syncthreads();
  const int TR = mp.TR;
  const long long ntile = (N + TR - 1) / TR;
  const long long t0 = ntile * s / S, t1 = ntile * (s + 1) / S;
  T stat = T(0);
  for (long long t = t0; t < t1; ++t) {
    const long long row0 = t * TR;
    const int nr = (int)((N - row0) < TR ? (N - row0) : TR);
    load_x_tile<T>(mp, sm + mp.offX, X, row0, nr);
    wait_x_tile();
    __syncthreads();
    if (BWD) {
      stat += tile_forward_backward<T>(mp, cx, Y, row0, nr);
    } else {
      tile_forward<T>(mp, cx);
      stat += lik_phase<T>(mp, cx, Y, row0, nr, sm + mp.offDa);
    }
  }
  if (BWD && mp.offG >= 0) {
    T* out = partial + ((size_t)c * S + s) * mp.Ppad;
    for (int i = 4 * threadIdx.x; i < mp.Ppad; i += 4 * blockDim.x) {
      T v[4];
      ld4(cx.G + i, v);
      st4(out + i, v);
    }
  }
  double* red = reinterpret_cast<double*>(sm + mp.offRed);
  const double tot = block_sum((double)stat, red);
  if (threadIdx.x == 0) stat_part[(size_t)c * S + s] = tot;
}

// Local reduction before an all-reduce (row-sharded sampling).
template <typename T>
__global__ void k_reduce_partials(const __grid_constant__ ModelPlan mp, int S,
                                  const T* __restrict__ partial, const double* __restrict__ stat_part,
                                  T* __restrict__ gsum) {
  const int c = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
  const int W = mp.Ppad + 4;
  if (i < mp.Ppad) {
    double g = 0.0;
    for (int s = 0; s < S; ++s) g += (double)partial[((size_t)c * S + s) * mp.Ppad + i];
    gsum[(size_t)c * W + i] = (T)g;
  } else if (i == mp.Ppad) {
    double st = 0.0;
    for (int s = 0; s < S; ++s) st += stat_part[(size_t)c * S + s];
    gsum[(size_t)c * W + i] = (T)st;
  }
}

// ------------------------------------------------------------------ priors
// value and d/dtheta of one element's prior term (layer.py:177-195, :357-375;
// activationFunctions.py:189-190, :341-346 with the passed hyper slice, Q4).
template <typename T>
__device__ __forceinline__ void prior_elem(const ModelPlan& mp, const Elem& e, const T* hy, double x,
                                           double& val, double& grad, bool need_val = true) {
  // need_val = false: gradient only (interior leapfrog steps never read the log-density; its fp64 log / log1p
  // are most of the cost of this function)
  const BlockPlan& b = mp.b[e.blk];
  const double kLog2Pi = tfc::kLog2PiCast;
  val = 0.0;
  if (e.kind == 3) {
    if (b.act == ACT_SQPRELU) {
      const double mean = (double)hy[b.ha];
      double sd = (double)hy[b.ha + 1];
      sd = fmin(fmax(sd, tfc::kClampLo), tfc::kClampHi);
      const double d = (x - mean) / sd;
      grad = -d / sd;
      if (need_val) {
        val = -0.5 * d * d;
        if (e.first) val += -0.5 * (2.0 * log(sd) + kLog2Pi);
      }
    } else {  // ACT_PRELU: exponentialLogProb(rate, slopes)
      const double r = fabs((double)hy[b.ha]);
      if (need_val) val = -r * x + log(r);
      grad = -r;
    }
    return;
  }
  const int h0 = b.hw + (e.kind == 1 ? 0 : 2);
  const double loc = (double)hy[h0];
  const double sc = (double)hy[h0 + 1] * (double)hy[h0 + 1];
  if (b.prior == PRIOR_CAUCHY) {   // +log(1+z^2) - log(pi*gamma)   (BNN_functions.py:51-56, Q1)
    const double z = (x - loc) / sc;
    if (need_val) val = log1p(z * z) - log(3.14159265358979323846 * sc);
    grad = 2.0 * z / ((1.0 + z * z) * sc);
  } else {                          // multivariateLogProb with scalar sigma (BNN_functions.py:21-32, Q2)
    const double sg = fmin(fmax(sc, tfc::kClampLo), tfc::kClampHi);
    const double d = (x - loc) / sg;
    grad = -d / sg;
    if (need_val) {
      val = -0.5 * d * d;
      if (e.first) val += -0.5 * (2.0 * log(sg) + kLog2Pi);
    }
  }
}

// ------------------------------------------------------------------ gradient assembly + update
template <typename T>
__global__ void __launch_bounds__(256)
k_finalize(const __grid_constant__ ModelPlan mp, int S, const T* __restrict__ partial,
           const double* __restrict__ stat_part, const T* __restrict__ gsum,
           const T* __restrict__ hyper, long long Ntot, T* __restrict__ theta_pad,
           T* __restrict__ mom_pad, T* __restrict__ grad_pad, const T* __restrict__ eps_dev,
           StepCoef cf, double* __restrict__ logp, double* __restrict__ stat_out,
           double* __restrict__ prior_part, unsigned* __restrict__ ticket, T* __restrict__ w1p) {
  __shared__ double red[40];
  __shared__ int is_last;
  const int c = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
  const T* hy = hyper + (size_t)c * mp.H;
  // likelihood scale 1/sigma^2 (Gaussian kinds)
  double sg = 1.0, scale = 1.0;
  if (mp.lik == LIK_GAUSS) {
    const double h = (double)hy[mp.lik_h];
    sg = fmin(fmax(h * h, tfc::kClampLo), tfc::kClampHi);
    scale = 1.0 / (sg * sg);
  } else if (mp.lik == LIK_FIXED) {
    sg = fmin(fmax(mp.fixed_sd, tfc::kClampLo), tfc::kClampHi);
    scale = 1.0 / (sg * sg);
  }
  double pv = 0.0;
  if (i < mp.Ppad) {
    const Elem e = decode_elem(mp, i);
    const size_t gi = (size_t)c * mp.Ppad + i;
    if (e.kind) {
      double g = 0.0;
      if (gsum) {
        g = (double)gsum[(size_t)c * (mp.Ppad + 4) + i];
      } else {
        for (int s = 0; s < S; ++s) g += (double)partial[((size_t)c * S + s) * mp.Ppad + i];
      }
      const T th = theta_pad[gi];
      double pg = 0.0;
      prior_elem<T>(mp, e, hy, (double)th, pv, pg, logp != nullptr);
      const T gt = (T)(g * scale + pg);
      grad_pad[gi] = gt;
      if (cf.m1 != 0.0 || cf.m2 != 0.0 || cf.m3 != 0.0) {
        const T eps = eps_dev[c];
        T p = mom_pad[gi];
        if (cf.m1 != 0.0) p = p + (T(cf.m1) * eps) * gt;
        if (cf.m2 != 0.0) p = p - (T(cf.m2) * eps) * gt;
        mom_pad[gi] = p;
        if (cf.m3 != 0.0) {
          const T tn = th + (T(cf.m3) * eps) * p;
          theta_pad[gi] = tn;
          if (w1p) {
            const int q = w1p_index(mp, i);
            if (q >= 0) w1p[(size_t)c * w1p_elems(mp) + q] = tn;
          }
        }
      }
    } else {
      grad_pad[gi] = T(0);
    }
  }
  if (logp == nullptr) return;
  const double bs = block_sum(pv, red);
  if (threadIdx.x == 0) {
    prior_part[(size_t)c * gridDim.x + blockIdx.x] = bs;
    __threadfence();
    const unsigned t = atomicAdd(&ticket[c], 1u);
    is_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  double acc = 0.0;
  for (int j = threadIdx.x; j < (int)gridDim.x; j += blockDim.x)
    acc += __ldcg(&prior_part[(size_t)c * gridDim.x + j]);
  const double prior = block_sum(acc, red);
  if (threadIdx.x == 0) {
    double st = 0.0;
    if (gsum) st = (double)gsum[(size_t)c * (mp.Ppad + 4) + mp.Ppad];
    else for (int s = 0; s < S; ++s) st += stat_part[(size_t)c * S + s];
    double ll;
    if (mp.lik == LIK_BERN) {
      ll = st;
    } else {
      const double n = (double)Ntot * (double)mp.OUT;
      ll = -0.5 * (2.0 * n * log(sg) + st * scale + n * tfc::kLog2PiCast);
    }
    logp[c] = prior + ll;
    if (stat_out) stat_out[c] = st;
    ticket[c] = 0u;
  }
}


// Same contract as k_finalize, for MANY partial vectors (single chain on all SMs): a CTA owns 32
// consecutive parameters (one 128-byte line of every partial vector) and its 8 warps split the S
// partials; fixed summation order => deterministic.
template <typename T>
__global__ void __launch_bounds__(256)
k_finalize_split(const __grid_constant__ ModelPlan mp, int S, const T* __restrict__ partial,
                 const double* __restrict__ stat_part, const T* __restrict__ hyper, long long Ntot,
                 T* __restrict__ theta_pad, T* __restrict__ mom_pad, T* __restrict__ grad_pad,
                 const T* __restrict__ eps_dev, StepCoef cf, double* __restrict__ logp,
                 double* __restrict__ stat_out, double* __restrict__ prior_part,
                 unsigned* __restrict__ ticket, T* __restrict__ w1p) {
  __shared__ double red[40];
  __shared__ double gs[8][33];
  __shared__ int is_last;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.y, i = blockIdx.x * 32 + tx;
  const T* hy = hyper + (size_t)c * mp.H;
  double g = 0.0;
  if (i < mp.Ppad) {
    const T* src = partial + (size_t)c * S * mp.Ppad + i;
    double g0 = 0.0, g1 = 0.0, g2 = 0.0, g3 = 0.0;
    int s = ty;
    for (; s + 24 < S; s += 32) {
      g0 += (double)src[(size_t)s * mp.Ppad];
      g1 += (double)src[(size_t)(s + 8) * mp.Ppad];
      g2 += (double)src[(size_t)(s + 16) * mp.Ppad];
      g3 += (double)src[(size_t)(s + 24) * mp.Ppad];
    }
    for (; s < S; s += 8) g0 += (double)src[(size_t)s * mp.Ppad];
    g = (g0 + g1) + (g2 + g3);
  }
  gs[ty][tx] = g;
  __syncthreads();
  double sg = 1.0, scale = 1.0;
  if (mp.lik == LIK_GAUSS) {
    const double h = (double)hy[mp.lik_h];
    sg = fmin(fmax(h * h, tfc::kClampLo), tfc::kClampHi);
    scale = 1.0 / (sg * sg);
  } else if (mp.lik == LIK_FIXED) {
    sg = fmin(fmax(mp.fixed_sd, tfc::kClampLo), tfc::kClampHi);
    scale = 1.0 / (sg * sg);
  }
  double pv = 0.0;
  if (ty == 0 && i < mp.Ppad) {
    const Elem e = decode_elem(mp, i);
    const size_t gi = (size_t)c * mp.Ppad + i;
    if (e.kind) {
      g = 0.0;
#pragma unroll
      for (int j = 0; j < 8; ++j) g += gs[j][tx];
      const T th = theta_pad[gi];
      double pg = 0.0;
      prior_elem<T>(mp, e, hy, (double)th, pv, pg, logp != nullptr);
      const T gt = (T)(g * scale + pg);
      grad_pad[gi] = gt;
      if (cf.m1 != 0.0 || cf.m2 != 0.0 || cf.m3 != 0.0) {
        const T eps = eps_dev[c];
        T p = mom_pad[gi];
        if (cf.m1 != 0.0) p = p + (T(cf.m1) * eps) * gt;
        if (cf.m2 != 0.0) p = p - (T(cf.m2) * eps) * gt;
        mom_pad[gi] = p;
        if (cf.m3 != 0.0) {
          const T tn = th + (T(cf.m3) * eps) * p;
          theta_pad[gi] = tn;
          if (w1p) {
            const int q = w1p_index(mp, i);
            if (q >= 0) w1p[(size_t)c * w1p_elems(mp) + q] = tn;
          }
        }
      }
    } else {
      grad_pad[gi] = T(0);
    }
  }
  if (logp == nullptr) return;
  const double bs = block_sum(pv, red);
  if (threadIdx.x == 0) {
    prior_part[(size_t)c * gridDim.x + blockIdx.x] = bs;
    __threadfence();
    const unsigned t = atomicAdd(&ticket[c], 1u);
    is_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  double acc = 0.0;
  for (int j = threadIdx.x; j < (int)gridDim.x; j += blockDim.x)
    acc += __ldcg(&prior_part[(size_t)c * gridDim.x + j]);
  const double prior = block_sum(acc, red);
  if (threadIdx.x == 0) {
    double st = 0.0;
    for (int s = 0; s < S; ++s) st += stat_part[(size_t)c * S + s];
    double ll;
    if (mp.lik == LIK_BERN) {
      ll = st;
    } else {
      const double n = (double)Ntot * (double)mp.OUT;
      ll = -0.5 * (2.0 * n * log(sg) + st * scale + n * tfc::kLog2PiCast);
    }
    logp[c] = prior + ll;
    if (stat_out) stat_out[c] = st;
    ticket[c] = 0u;
  }
}

// ------------------------------------------------------------------ momentum draw
template <typename T>
__global__ void __launch_bounds__(256)
k_momentum(const __grid_constant__ ModelPlan mp, uint64_t seed, uint64_t call,
           const T* __restrict__ injected_flat, T* __restrict__ mom_pad, double* __restrict__ ke) {
  __shared__ double red[40];
  const int c = blockIdx.x;
  double k = 0.0;
  for (int i = threadIdx.x; i < mp.Ppad; i += blockDim.x) {
    const Elem e = decode_elem(mp, i);
    T p = T(0);
    if (e.kind) {
      p = injected_flat ? injected_flat[(size_t)c * mp.P + e.flat]
                        : draw_normal<T>(seed, STREAM_MAIN, call, (uint32_t)c, (uint32_t)e.flat);
    }
    mom_pad[(size_t)c * mp.Ppad + i] = p;
    k += 0.5 * (double)p * (double)p;
  }
  k = block_sum(k, red);
  if (threadIdx.x == 0) ke[c] = k;
}

// ------------------------------------------------------------------ Metropolis-Hastings
// log_accept_ratio = safe_sum(logp' - logp + KE0 - KE1); accept iff log u < lar;
// reported accept prob = where(lar<0, exp(lar), 1)  (network.py:410-411).
template <typename T>
__global__ void __launch_bounds__(256)
k_mh(const __grid_constant__ ModelPlan mp, uint64_t seed, uint64_t call, const T* __restrict__ u_in,
     const T* __restrict__ theta0_pad, const T* __restrict__ theta1_pad,
     const T* __restrict__ mom1_pad, const double* __restrict__ logp0,
     const double* __restrict__ logp1, const double* __restrict__ ke0,
     const double* __restrict__ stat0, const double* __restrict__ stat1,
     double* __restrict__ stat_cur, T* __restrict__ theta_flat, T* __restrict__ stats) {
  __shared__ double red[40];
  __shared__ int acc_s;
  const int c = blockIdx.x;
  const size_t base = (size_t)c * mp.Ppad;
  double k1 = 0.0, sjd = 0.0;
  for (int i = threadIdx.x; i < mp.Ppad; i += blockDim.x) {
    const double p = (double)mom1_pad[base + i];
    const double d = (double)theta1_pad[base + i] - (double)theta0_pad[base + i];
    k1 += 0.5 * p * p;
    sjd += d * d;
  }
  k1 = block_sum(k1, red);
  sjd = block_sum(sjd, red);
  if (threadIdx.x == 0) {
    const double t[4] = {logp1[c], -logp0[c], ke0[c], -k1};
    bool nan = false, pinf = false, ninf = false;
    double lar = 0.0;
    for (int j = 0; j < 4; ++j) {
      nan |= isnan(t[j]);
      pinf |= (isinf(t[j]) && t[j] > 0);
      ninf |= (isinf(t[j]) && t[j] < 0);
      lar += t[j];
    }
    if (nan || (pinf && ninf)) lar = -INFINITY;
    const double u = u_in ? (double)u_in[c] : draw_uniform(seed, STREAM_MAIN, call, (uint32_t)c);
    const int acc = (log(u) < lar) ? 1 : 0;
    acc_s = acc;
    if (stats) {
      stats[c * 4 + 0] = (T)lar;
      stats[c * 4 + 1] = (T)(lar < 0.0 ? exp(lar) : 1.0);
      stats[c * 4 + 2] = (T)acc;
      stats[c * 4 + 3] = (T)(acc ? sjd : 0.0);
    }
    if (stat_cur) stat_cur[c] = acc ? stat1[c] : stat0[c];
  }
  __syncthreads();
  const T* src = acc_s ? theta1_pad : theta0_pad;
  for (int i = threadIdx.x; i < mp.Ppad; i += blockDim.x) {
    const Elem e = decode_elem(mp, i);
    if (e.kind) theta_flat[(size_t)c * mp.P + e.flat] = src[base + i];
  }
}

// ------------------------------------------------------------------ launchers
template <typename T>
void Launch<T>::pad(const ModelPlan& mp, int C, const T* flat, T* padded, cudaStream_t st, T* w1p) {
  dim3 g((mp.Ppad + 255) / 256, C);
  k_pad<T><<<g, 256, 0, st>>>(mp, flat, padded, w1p);
}
template <typename T>
void Launch<T>::unpad(const ModelPlan& mp, int C, const T* padded, T* flat, cudaStream_t st) {
  dim3 g((mp.Ppad + 255) / 256, C);
  k_unpad<T><<<g, 256, 0, st>>>(mp, padded, flat);
}
template <typename T>
void Launch<T>::partial(const ModelPlan& mp, int C, int S, bool backward, const T* theta_pad,
                        const T* X, const T* Y, long long N, T* partial, double* stat_part,
                        cudaStream_t st) {
  dim3 g(S, C);
  const size_t smem = (size_t)mp.smem_elems * sizeof(T);
  if (backward) {
    cudaFuncSetAttribute(k_partial<T, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k_partial<T, true><<<g, NT, smem, st>>>(mp, S, theta_pad, X, Y, N, partial, stat_part);
  } else {
    cudaFuncSetAttribute(k_partial<T, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k_partial<T, false><<<g, NT, smem, st>>>(mp, S, theta_pad, X, Y, N, partial, stat_part);
  }
}
template <typename T>
void Launch<T>::reduce_partials(const ModelPlan& mp, int C, int S, const T* partial,
                                const double* stat_part, T* gsum, cudaStream_t st) {
  dim3 g((mp.Ppad + 4 + 255) / 256, C);
  k_reduce_partials<T><<<g, 256, 0, st>>>(mp, S, partial, stat_part, gsum);
}
template <typename T>
void Launch<T>::finalize(const ModelPlan& mp, int C, int S, const T* partial, const double* stat_part,
                         const T* gsum, const T* hyper, long long N_total, T* theta_pad, T* mom_pad,
                         T* grad_pad, const T* eps_dev, StepCoef cf, double* logp, double* stat_out,
                         double* prior_part, unsigned* ticket, cudaStream_t st, T* w1p) {
  if (gsum == nullptr && S >= FINALIZE_SPLIT_MIN_S) {
    dim3 g2((mp.Ppad + 31) / 32, C);
    k_finalize_split<T><<<g2, 256, 0, st>>>(mp, S, partial, stat_part, hyper, N_total, theta_pad, mom_pad,
                                            grad_pad, eps_dev, cf, logp, stat_out, prior_part, ticket, w1p);
    return;
  }
  dim3 g((mp.Ppad + 255) / 256, C);
  k_finalize<T><<<g, 256, 0, st>>>(mp, S, partial, stat_part, gsum, hyper, N_total, theta_pad, mom_pad,
                                   grad_pad, eps_dev, cf, logp, stat_out, prior_part, ticket, w1p);
}
template <typename T>
void Launch<T>::momentum(const ModelPlan& mp, int C, uint64_t seed, uint64_t call,
                         const T* injected_flat, T* mom_pad, double* ke, cudaStream_t st) {
  k_momentum<T><<<C, 256, 0, st>>>(mp, seed, call, injected_flat, mom_pad, ke);
}
template <typename T>
void Launch<T>::mh(const ModelPlan& mp, int C, uint64_t seed, uint64_t call, const T* u_in,
                   const T* theta0_pad, const T* theta1_pad, const T* mom1_pad, const double* logp0,
                   const double* logp1, const double* ke0, const double* stat0, const double* stat1,
                   double* stat_cur, T* theta_flat, T* stats, cudaStream_t st) {
  k_mh<T><<<C, 256, 0, st>>>(mp, seed, call, u_in, theta0_pad, theta1_pad, mom1_pad, logp0, logp1, ke0,
                             stat0, stat1, stat_cur, theta_flat, stats);
}

// ------------------------------------------------------------------ persistent trajectory (small problems)
// One CTA per chain runs the WHOLE L-step leapfrog trajectory (L + 1 log-posterior + gradient evaluations, TFP
// order: SURVEY App. B) in one launch when the training set is a single tile: the X tile, the parameters and the
// gradient accumulators stay in shared memory, the momentum in registers, and nothing goes through HBM or the host
// between steps.  Replaces 2 (L + 1) launches of k_partial + k_finalize at the reference's own example size
// (Examples/trainRegression.py: 11 rows, 251 parameters), where launch latency was the whole cost.
// Same arithmetic as k_partial (S = 1) + k_finalize.
constexpr int TRAJ_EPT = 8;    // parameters per thread (Ppad <= TRAJ_EPT * NT)
template <typename T>
__global__ void __launch_bounds__(NT, 1)
k_traj_small(const __grid_constant__ ModelPlan mp, const T* __restrict__ X, const T* __restrict__ Y, long long N,
             const T* __restrict__ hyper, long long Ntot, T* __restrict__ theta_pad, T* __restrict__ mom_pad,
             T* __restrict__ grad_pad, const T* __restrict__ eps_dev, int L, double* __restrict__ logp_first,
             double* __restrict__ stat_first, double* __restrict__ logp_last, double* __restrict__ stat_last) {
  extern __shared__ __align__(16) unsigned char smraw[];
  T* sm = reinterpret_cast<T*>(smraw);
  const int c = blockIdx.x, tid = threadIdx.x;
  const T* hy = hyper + (size_t)c * mp.H;
  T* thg = theta_pad + (size_t)c * mp.Ppad;
  TileCtx<T> cx;
  cx.sm = sm;
  cx.G = sm + mp.offG;
  T* Ws = sm + mp.offW;
  cx.Wp = Ws;
  for (int i = tid; i < mp.Ppad; i += NT) Ws[i] = thg[i];
  T p[TRAJ_EPT], gt[TRAJ_EPT];
  Elem el[TRAJ_EPT];               // what each owned parameter is (decoded once, not once per step)
#pragma unroll
  for (int k = 0; k < TRAJ_EPT; ++k) {
    const int i = tid + k * NT;
    p[k] = i < mp.Ppad ? mom_pad[(size_t)c * mp.Ppad + i] : T(0);
    gt[k] = T(0);
    el[k] = decode_elem(mp, i < mp.Ppad ? i : 0);
  }
  load_x_tile<T>(mp, sm + mp.offX, X, 0, (int)N);
  wait_x_tile();
  double sg = 1.0, scale = 1.0;
  if (mp.lik == LIK_GAUSS) {
    const double h = (double)hy[mp.lik_h];
    sg = fmin(fmax(h * h, tfc::kClampLo), tfc::kClampHi);
    scale = 1.0 / (sg * sg);
  } else if (mp.lik == LIK_FIXED) {
    sg = fmin(fmax(mp.fixed_sd, tfc::kClampLo), tfc::kClampHi);
    scale = 1.0 / (sg * sg);
  }
  const T eps = eps_dev[c];
  double* red = reinterpret_cast<double*>(sm + mp.offRed);
  for (int j = 0; j <= L; ++j) {
    for (int i = tid; i < mp.Ppad; i += NT) cx.G[i] = T(0);
    __syncthreads();
    const T stat = tile_forward_backward<T>(mp, cx, Y, 0, (int)N);
    // the likelihood statistic (and the prior log-density below) only where a log-posterior is reported: the
    // interior steps need the barrier that completes G, not the fp64 block reductions
    const bool need_val = j == L || (j == 0 && logp_first != nullptr);
    double st = 0.0;
    if (need_val) st = block_sum((double)stat, red);
    else __syncthreads();
    // gradient assembly + leapfrog update (k_finalize): step 0 {0.5, 0, 1}, interior {1, 0, 1}, last {1, 0.5, 0}
    const T m1 = j == 0 ? T(0.5) : T(1), m2 = j == L ? T(0.5) : T(0);
    const bool move = j < L;
    double pv = 0.0;
#pragma unroll
    for (int k = 0; k < TRAJ_EPT; ++k) {
      const int i = tid + k * NT;
      if (i < mp.Ppad) {
        const Elem e = el[k];
        if (e.kind) {
          const T th = Ws[i];
          double v = 0.0, pg = 0.0;
          prior_elem<T>(mp, e, hy, (double)th, v, pg, need_val);
          pv += v;
          gt[k] = (T)((double)cx.G[i] * scale + pg);
          p[k] = p[k] + (m1 * eps) * gt[k];
          if (m2 != T(0)) p[k] = p[k] - (m2 * eps) * gt[k];
          if (move) Ws[i] = th + eps * p[k];
        } else {
          gt[k] = T(0);
        }
      }
    }
    const bool want_first = j == 0 && logp_first != nullptr;
    if (want_first || j == L) {
      const double prior = block_sum(pv, red);
      if (tid == 0) {
        double ll;
        if (mp.lik == LIK_BERN) {
          ll = st;
        } else {
          const double n = (double)Ntot * (double)mp.OUT;
          ll = -0.5 * (2.0 * n * log(sg) + st * scale + n * tfc::kLog2PiCast);
        }
        if (j == L) { logp_last[c] = prior + ll; if (stat_last) stat_last[c] = st; }
        else { logp_first[c] = prior + ll; if (stat_first) stat_first[c] = st; }
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int k = 0; k < TRAJ_EPT; ++k) {
    const int i = tid + k * NT;
    if (i < mp.Ppad) {
      const size_t gi = (size_t)c * mp.Ppad + i;
      thg[i] = Ws[i];
      mom_pad[gi] = p[k];
      grad_pad[gi] = gt[k];
    }
  }
}

template <typename T>
bool Launch<T>::traj_small_ok(const ModelPlan& mp, long long N, int S) {
  return S == 1 && N <= mp.TR && mp.offW >= 0 && mp.offG >= 0 && mp.Ppad <= TRAJ_EPT * NT;
}
template <typename T>
void Launch<T>::traj_small(const ModelPlan& mp, int C, const T* X, const T* Y, long long N, const T* hyper,
                           long long N_total, T* theta_pad, T* mom_pad, T* grad_pad, const T* eps_dev, int L,
                           double* logp_first, double* stat_first, double* logp_last, double* stat_last,
                           cudaStream_t st) {
  const size_t smem = (size_t)mp.smem_elems * sizeof(T);
  cudaFuncSetAttribute(k_traj_small<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  k_traj_small<T><<<C, NT, smem, st>>>(mp, X, Y, N, hyper, N_total, theta_pad, mom_pad, grad_pad, eps_dev, L,
                                       logp_first, stat_first, logp_last, stat_last);
}

// ------------------------------------------------------------------ persistent trajectory (narrow networks)
// Same job as k_traj_small for networks whose every width (input included) is <= 32 and whose training set is
// <= 64 rows -- the reference's own example (Examples/trainRegression.py: 11 rows, 1-10-10-10-1): one training row
// per HALF-WARP (narrow.cuh: lane owns two neurons, no CTA barrier inside a row), so a gradient evaluation is a
// handful of barriers instead of the tile engine's ~20 phases.
constexpr int NARROW_ROWS = 64;
template <typename T>
__global__ void __launch_bounds__(NT, 1)
k_traj_narrow(const __grid_constant__ ModelPlan mp, const T* __restrict__ X, const T* __restrict__ Y, int N,
              const T* __restrict__ hyper, long long Ntot, T* __restrict__ theta_pad, T* __restrict__ mom_pad,
              T* __restrict__ grad_pad, const T* __restrict__ eps_dev, int L, double* __restrict__ logp_first,
              double* __restrict__ stat_first, double* __restrict__ logp_last, double* __restrict__ stat_last) {
  extern __shared__ __align__(16) unsigned char smraw[];
  T* sm = reinterpret_cast<T*>(smraw);
  const int c = blockIdx.x, tid = threadIdx.x, hw = tid >> 4, jl = tid & 15;
  const BlockPlan& b0 = mp.b[0];
  const T* hy = hyper + (size_t)c * mp.H;
  T* thg = theta_pad + (size_t)c * mp.Ppad;
  T* Ws = sm + mp.offW;
  T* G = sm + mp.offG;
  T* Xs = sm + mp.offX;
  const int ld0 = mp.ld0, D = mp.D;
  for (int i = tid; i < mp.Ppad; i += NT) Ws[i] = thg[i];
  for (int e = tid; e < N * ld0; e += NT) {
    const int r = e / ld0, k = e - r * ld0;
    Xs[e] = k < D ? X[(long long)r * D + k] : T(0);
  }
  T p[TRAJ_EPT], gt[TRAJ_EPT];
  Elem el[TRAJ_EPT];               // what each owned parameter is (decoded once, not once per step)
#pragma unroll
  for (int k = 0; k < TRAJ_EPT; ++k) {
    const int i = tid + k * NT;
    p[k] = i < mp.Ppad ? mom_pad[(size_t)c * mp.Ppad + i] : T(0);
    gt[k] = T(0);
    el[k] = decode_elem(mp, i < mp.Ppad ? i : 0);
  }
  double sg = 1.0, scale = 1.0;
  if (mp.lik == LIK_GAUSS) {
    const double h = (double)hy[mp.lik_h];
    sg = fmin(fmax(h * h, tfc::kClampLo), tfc::kClampHi);
    scale = 1.0 / (sg * sg);
  } else if (mp.lik == LIK_FIXED) {
    sg = fmin(fmax(mp.fixed_sd, tfc::kClampLo), tfc::kClampHi);
    scale = 1.0 / (sg * sg);
  }
  const T eps = eps_dev[c];
  double* red = reinterpret_cast<double*>(sm + mp.offRed);
  __syncthreads();
  for (int j = 0; j <= L; ++j) {
    for (int i = tid; i < mp.Ppad; i += NT) G[i] = T(0);
    // ---- rows: block 0 forward (lane owns outputs jl, jl + 16), then the narrow chain of the half-warp
    T stat = T(0);
    for (int r0 = 0; r0 < N; r0 += NT / 16) {
      const int r = r0 + hw;
      const bool active = r < N;
      if (active) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const int o = jl + 16 * i;
          if (o < b0.out_p) {
            T z = T(0), a = T(0);
            if (o < b0.out) {
              z = Ws[b0.pb + o];
              const T* w = Ws + b0.pw + o * b0.ld_in;
              const T* x = Xs + r * ld0;
              for (int k = 0; k < b0.in_p; ++k) z = fma(w[k], x[k], z);
              T slope = T(0);
              if (act_keeps_z(b0.act)) slope = eff_slope<T>(b0.act, Ws + (b0.ps >= 0 ? b0.ps : 0), o, T(b0.alpha));
              a = act_fwd<T>(b0.act, z, slope);
            }
            sm[b0.offS + r * b0.ld_out + o] = a;
            if (b0.offZ >= 0) sm[b0.offZ + r * b0.ld_out + o] = z;
          }
        }
      }
      __syncwarp();
      stat += narrow_row<T>(mp, Ws, sm, active ? r : 0, jl, active, Y, r, sm + b0.offD + (active ? r : 0) * b0.ld_out);
    }
    __syncthreads();
    // ---- parameter gradients: blocks >= 1 from the row buffers (narrow_accum), block 0 against the X tile
    if (mp.nb > 1) narrow_accum<T>(mp, Ws, G, sm, N, tid, NT);
    for (int e = tid; e < b0.out_p * b0.in_p; e += NT) {
      const int o = e / b0.in_p, k = e - o * b0.in_p;
      T acc = T(0);
      for (int r = 0; r < N; ++r) acc = fma(sm[b0.offD + r * b0.ld_out + o], Xs[r * ld0 + k], acc);
      G[b0.pw + o * b0.ld_in + k] = acc;
    }
    for (int o = tid; o < b0.out_p; o += NT) {
      T sb = T(0), sc = T(0);
      for (int r = 0; r < N; ++r) {
        sb += sm[b0.offD + r * b0.ld_out + o];
        if (act_has_slopes(b0.act)) sc += sm[b0.offZ + r * b0.ld_out + o];
      }
      G[b0.pb + o] = sb;
      if (act_has_slopes(b0.act)) G[b0.ps + o] = (b0.act == ACT_SQPRELU ? T(2) * Ws[b0.ps + o] : T(1)) * sc;
    }
    // the likelihood statistic (and the prior log-density below) only where a log-posterior is reported: the
    // interior steps need the barrier that completes G, not the fp64 block reductions
    const bool need_val = j == L || (j == 0 && logp_first != nullptr);
    double st = 0.0;
    if (need_val) st = block_sum((double)stat, red);
    else __syncthreads();
    // ---- gradient assembly + leapfrog update (k_finalize): step 0 {0.5, 0, 1}, interior {1, 0, 1}, last {1, 0.5, 0}
    const T m1 = j == 0 ? T(0.5) : T(1), m2 = j == L ? T(0.5) : T(0);
    const bool move = j < L;
    double pv = 0.0;
#pragma unroll
    for (int k = 0; k < TRAJ_EPT; ++k) {
      const int i = tid + k * NT;
      if (i < mp.Ppad) {
        const Elem e = el[k];
        if (e.kind) {
          const T th = Ws[i];
          double v = 0.0, pg = 0.0;
          prior_elem<T>(mp, e, hy, (double)th, v, pg, need_val);
          pv += v;
          gt[k] = (T)((double)G[i] * scale + pg);
          p[k] = p[k] + (m1 * eps) * gt[k];
          if (m2 != T(0)) p[k] = p[k] - (m2 * eps) * gt[k];
          if (move) Ws[i] = th + eps * p[k];
        } else {
          gt[k] = T(0);
        }
      }
    }
    const bool want_first = j == 0 && logp_first != nullptr;
    if (want_first || j == L) {
      const double prior = block_sum(pv, red);
      if (tid == 0) {
        double ll;
        if (mp.lik == LIK_BERN) {
          ll = st;
        } else {
          const double n = (double)Ntot * (double)mp.OUT;
          ll = -0.5 * (2.0 * n * log(sg) + st * scale + n * tfc::kLog2PiCast);
        }
        if (j == L) { logp_last[c] = prior + ll; if (stat_last) stat_last[c] = st; }
        else { logp_first[c] = prior + ll; if (stat_first) stat_first[c] = st; }
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int k = 0; k < TRAJ_EPT; ++k) {
    const int i = tid + k * NT;
    if (i < mp.Ppad) {
      const size_t gi = (size_t)c * mp.Ppad + i;
      thg[i] = Ws[i];
      mom_pad[gi] = p[k];
      grad_pad[gi] = gt[k];
    }
  }
}

// Plan of k_traj_narrow (offsets in elements): W, G, X tile [64][ld0], row buffers S_l / Z_l / dz_l [64][ld_out].
template <typename T>
bool Launch<T>::plan_traj_narrow(const ModelPlan& mp, ModelPlan& np, size_t smem_limit) {
  if (mp.nb > 4 || mp.OUT > 32 || mp.Ppad > TRAJ_EPT * NT) return false;
  for (int l = 0; l < mp.nb; ++l)
    if (mp.b[l].in_p > 32 || mp.b[l].out_p > 32) return false;
  np = mp;
  np.TR = NARROW_ROWS;
  int cur = 0;
  np.offW = cur; cur += np.Ppad;
  np.offG = cur; cur += np.Ppad;
  np.offX = cur; cur += NARROW_ROWS * np.ld0;
  for (int l = 0; l < np.nb; ++l) {
    BlockPlan& b = np.b[l];
    b.ksplit = 1;
    b.offS = cur; cur += NARROW_ROWS * b.ld_out;
    if (act_keeps_z(b.act)) { b.offZ = cur; cur += NARROW_ROWS * b.ld_out; } else b.offZ = -1;
    b.offD = cur; cur += NARROW_ROWS * b.ld_out;
  }
  const int per8 = (int)(8 / sizeof(T));
  cur = (cur + per8 - 1) / per8 * per8;
  np.offRed = cur; cur += 64 * per8;
  np.smem_elems = cur;
  return (size_t)cur * sizeof(T) <= smem_limit;
}
template <typename T>
void Launch<T>::traj_narrow(const ModelPlan& np, int C, const T* X, const T* Y, long long N, const T* hyper,
                            long long N_total, T* theta_pad, T* mom_pad, T* grad_pad, const T* eps_dev, int L,
                            double* logp_first, double* stat_first, double* logp_last, double* stat_last,
                            cudaStream_t st) {
  const size_t smem = (size_t)np.smem_elems * sizeof(T);
  cudaFuncSetAttribute(k_traj_narrow<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  k_traj_narrow<T><<<C, NT, smem, st>>>(np, X, Y, (int)N, hyper, N_total, theta_pad, mom_pad, grad_pad, eps_dev, L,
                                        logp_first, stat_first, logp_last, stat_last);
}

template struct Launch<float>;
template struct Launch<double>;

}  // namespace tbnn
