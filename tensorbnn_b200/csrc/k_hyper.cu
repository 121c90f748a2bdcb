// k_hyper.cu -- the hyper-parameter chain on device, one CTA per chain.
// Target = closure of network.py:417-440: sum of layer.calculateHyperProbs
// (layer.py:199-242, :379-422; activationFunctions.py:194-220, :350-382) plus the
// full-data Gaussian likelihood when likelihood.mainProbsInHypers (likelihood.py:67),
// which at fixed theta only needs SSE = sum (y-f)^2.  HMC transition + the hand-rolled
// dual averaging of network.py:442-471.  Every evaluation sweeps the chain's P weights
// (needed for the Cauchy prior) with warp-shuffle reductions; no host round trip.
#include "engine.cuh"
#include "kernels.h"
#include "philox.cuh"

namespace tbnn {

constexpr int HT = 256;           // threads per CTA
constexpr int MAXH = 8 * MAXB;    // upper bound on hyper scalars
constexpr int NSLOT = 9 * MAXB;   // 3 sums x (W, b, slopes) per block

struct HyperSmem {
  double wsum[HT / 32][NSLOT];
  double sums[NSLOT];
  double g[MAXH];
  double part[MAXB + 1];
  double red[40];
};

__device__ __forceinline__ double logn(double v, double m, double s) {
  const double d = (v - m) / s;
  return -0.5 * d * d - log(s) - 0.9189385332046727;   // tfd.MultivariateNormalDiag, 1 element
}

// three sums of one tensor for the given (loc, scale^2 or sd) hypers
template <typename T>
__device__ __forceinline__ void tensor_sums(const T* __restrict__ x, int n, int mode, double loc,
                                            double sc, double (&s)[3]) {
  s[0] = s[1] = s[2] = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    double v = (double)x[i];
    if (mode == 0) {            // Cauchy
      const double z = (v - loc) / sc, q = 2.0 * z / (1.0 + z * z);
      s[0] += log1p(z * z); s[1] += q; s[2] += q * z;
    } else if (mode == 1) {     // Gaussian on v
      const double d = v - loc;
      s[0] += d * d; s[1] += d;
    } else if (mode == 2) {     // Gaussian on v^2 (SquarePrelu hyper conditional)
      const double d = v * v - loc;
      s[0] += d * d; s[1] += d;
    } else {                    // sum |v| (Prelu)
      s[0] += fabs(v);
    }
  }
}

// Evaluates the hyper target at hv[0..H) (double, shared memory); returns logp (all threads)
// and leaves the gradient in sh.g.
template <typename T>
__device__ double hyper_eval_dev(const ModelPlan& mp, const T* __restrict__ th, const double* hv,
                                 double sse, long long Ntot, HyperSmem& sh) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  // ---- sweeps
  for (int l = 0; l < mp.nb; ++l) {
    const BlockPlan& b = mp.b[l];
    for (int t = 0; t < 3; ++t) {
      double s[3] = {0.0, 0.0, 0.0};
      if (t < 2) {
        const double loc = hv[b.hw + 2 * t], h = hv[b.hw + 2 * t + 1];
        double sc = h * h;
        const int n = t == 0 ? b.out * b.in : b.out;
        const T* x = th + (t == 0 ? b.fw : b.fb);
        if (b.prior == PRIOR_CAUCHY) tensor_sums<T>(x, n, 0, loc, sc, s);
        else tensor_sums<T>(x, n, 1, loc, sc, s);
      } else if (act_has_slopes(b.act)) {
        if (b.act == ACT_SQPRELU) tensor_sums<T>(th + b.fs, b.out, 2, hv[b.ha], 1.0, s);
        else tensor_sums<T>(th + b.fs, b.out, 3, 0.0, 1.0, s);
      }
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        const double v = warp_sum(s[q]);
        if (lane == 0) sh.wsum[w][(l * 3 + t) * 3 + q] = v;
      }
    }
  }
  __syncthreads();
  for (int j = threadIdx.x; j < mp.nb * 9; j += blockDim.x) {
    double v = 0.0;
    for (int k = 0; k < HT / 32; ++k) v += sh.wsum[k][j];
    sh.sums[j] = v;
  }
  __syncthreads();
  // ---- per-block value and gradient (thread l handles block l; thread nb the likelihood)
  const double kLog2Pi = tfc::kLog2PiCast, kPi = 3.14159265358979323846;
  if (threadIdx.x < mp.nb) {
    const int l = threadIdx.x;
    const BlockPlan& b = mp.b[l];
    double val = 0.0;
    for (int t = 0; t < 2; ++t) {
      const double* s = &sh.sums[(l * 3 + t) * 3];
      const int i0 = b.hw + 2 * t, i1 = i0 + 1;
      const double loc = hv[i0], h = hv[i1], sc = h * h;
      const double n = t == 0 ? (double)b.out * b.in : (double)b.out;
      if (b.prior == PRIOR_CAUCHY) {
        val += s[0] - n * log(kPi * sc);
        val += logn(loc, 0.0, tfc::k0p2) + logn(sc, tfc::kSqrtHalf, 0.5);      // layer.py:136-153
        sh.g[i0] = -s[1] / sc - loc / (tfc::k0p2 * tfc::k0p2);
        sh.g[i1] = ((-s[2] / sc - n / sc) - (sc - tfc::kSqrtHalf) / (0.5 * 0.5)) * 2.0 * h;
      } else {
        const double sg = fmin(fmax(sc, tfc::kClampLo), tfc::kClampHi);
        const bool inside = sc >= tfc::kClampLo && sc <= tfc::kClampHi;
        val += -0.5 * (2.0 * log(sg) + s[0] / (sg * sg) + kLog2Pi);
        val += logn(loc, 0.0, tfc::k0p1) + logn(sc, 1.0, tfc::k0p1);                      // layer.py:317-334
        sh.g[i0] = s[1] / (sg * sg) - loc / (tfc::k0p1 * tfc::k0p1);
        sh.g[i1] = ((inside ? (-1.0 / sg + s[0] / (sg * sg * sg)) : 0.0) - (sc - 1.0) / (tfc::k0p1 * tfc::k0p1)) *
                   2.0 * h;
      }
    }
    if (b.act == ACT_SQPRELU) {                       // activationFunctions.py:365-380
      const double* s = &sh.sums[(l * 3 + 2) * 3];
      const double mean = hv[b.ha], sd = hv[b.ha + 1];
      const double sg = fmin(fmax(sd, tfc::kClampLo), tfc::kClampHi);
      const bool inside = sd >= tfc::kClampLo && sd <= tfc::kClampHi;
      val += -0.5 * (2.0 * log(sg) + s[0] / (sg * sg) + kLog2Pi);
      val += logn(mean, 0.0, tfc::k0p3) + logn(sd, tfc::k0p3, tfc::k0p1);
      sh.g[b.ha] = s[1] / (sg * sg) - mean / (tfc::k0p3 * tfc::k0p3);
      sh.g[b.ha + 1] = (inside ? (-1.0 / sg + s[0] / (sg * sg * sg)) : 0.0) - (sd - tfc::k0p3) / (tfc::k0p1 * tfc::k0p1);
    } else if (b.act == ACT_PRELU) {                  // activationFunctions.py:209-218
      const double* s = &sh.sums[(l * 3 + 2) * 3];
      const double r = hv[b.ha], ar = fabs(r), n = (double)b.out;
      const double sgn = r > 0 ? 1.0 : (r < 0 ? -1.0 : 0.0);
      val += -tfc::k0p3 * r + log(tfc::k0p3);
      val += -ar * s[0] + n * log(ar);
      sh.g[b.ha] = -tfc::k0p3 + sgn * (-s[0] + n / ar);
    }
    sh.part[l] = val;
  } else if (threadIdx.x == mp.nb) {
    double val = 0.0;
    if (mp.lik == LIK_GAUSS) {                        // likelihood.py:88-94 with sd = argv[-1]
      const double h = hv[mp.lik_h], sd = h * h;
      const double sg = fmin(fmax(sd, tfc::kClampLo), tfc::kClampHi);
      const bool inside = sd >= tfc::kClampLo && sd <= tfc::kClampHi;
      const double n = (double)Ntot * (double)mp.OUT;
      val = -0.5 * (2.0 * n * log(sg) + sse / (sg * sg) + n * kLog2Pi);
      sh.g[mp.lik_h] = (inside ? (-n / sg + sse / (sg * sg * sg)) : 0.0) * 2.0 * h;
    }
    sh.part[mp.nb] = val;
  }
  __syncthreads();
  double lp = 0.0;
  for (int l = 0; l <= mp.nb; ++l) lp += sh.part[l];
  return lp;
}

template <typename T>
__global__ void __launch_bounds__(HT)
k_hyper_eval(const __grid_constant__ ModelPlan mp, const T* __restrict__ theta_flat,
             const T* __restrict__ hyper, const double* __restrict__ sse, long long Ntot,
             T* __restrict__ logp, T* __restrict__ grad) {
  __shared__ HyperSmem sh;
  __shared__ double hv[MAXH];
  const int c = blockIdx.x;
  for (int j = threadIdx.x; j < mp.H; j += blockDim.x) hv[j] = (double)hyper[(size_t)c * mp.H + j];
  __syncthreads();
  const double lp = hyper_eval_dev<T>(mp, theta_flat + (size_t)c * mp.P, hv, sse ? sse[c] : 0.0, Ntot, sh);
  if (threadIdx.x == 0) logp[c] = (T)lp;
  for (int j = threadIdx.x; j < mp.H; j += blockDim.x) grad[(size_t)c * mp.H + j] = (T)sh.g[j];
}

template <typename T>
__global__ void __launch_bounds__(HT)
k_hyper_step(const __grid_constant__ ModelPlan mp, const T* __restrict__ theta_flat,
             T* __restrict__ hyper, const double* __restrict__ sse, long long Ntot, uint64_t seed,
             uint64_t call, int L, double epoch, double burnin, double hyper_step0,
             T* __restrict__ da_state, const T* __restrict__ mom_in, const T* __restrict__ u_in,
             T* __restrict__ stats) {
  __shared__ HyperSmem sh;
  __shared__ double hv[MAXH];
  __shared__ T h0[MAXH], hc[MAXH], pm[MAXH], p0[MAXH];
  __shared__ int acc_s;
  const int c = blockIdx.x, H = mp.H;
  const T* th = theta_flat + (size_t)c * mp.P;
  const double sse_c = sse ? sse[c] : 0.0;
  const T eps = da_state[c * 3 + 2];
  const T half = eps * T(0.5);
  for (int j = threadIdx.x; j < H; j += blockDim.x) {
    const T v = hyper[(size_t)c * H + j];
    h0[j] = v; hc[j] = v; hv[j] = (double)v;
    p0[j] = mom_in ? mom_in[(size_t)c * H + j]
                   : draw_normal<T>(seed, STREAM_HYPER, call, (uint32_t)c, (uint32_t)j);
  }
  __syncthreads();
  const double lp0 = hyper_eval_dev<T>(mp, th, hv, sse_c, Ntot, sh);
  for (int j = threadIdx.x; j < H; j += blockDim.x) pm[j] = p0[j] + half * (T)sh.g[j];
  __syncthreads();
  double lp1 = lp0;
  for (int s = 0; s < L; ++s) {
    for (int j = threadIdx.x; j < H; j += blockDim.x) {
      hc[j] = hc[j] + eps * pm[j];
      hv[j] = (double)hc[j];
    }
    __syncthreads();
    lp1 = hyper_eval_dev<T>(mp, th, hv, sse_c, Ntot, sh);
    for (int j = threadIdx.x; j < H; j += blockDim.x) pm[j] = pm[j] + eps * (T)sh.g[j];
    __syncthreads();
  }
  for (int j = threadIdx.x; j < H; j += blockDim.x) pm[j] = pm[j] - half * (T)sh.g[j];
  __syncthreads();
  if (threadIdx.x == 0) {
    double k0 = 0.0, k1 = 0.0;
    for (int j = 0; j < H; ++j) {
      k0 += 0.5 * (double)p0[j] * (double)p0[j];
      k1 += 0.5 * (double)pm[j] * (double)pm[j];
    }
    const double t[4] = {lp1, -lp0, k0, -k1};
    bool nan = false, pinf = false, ninf = false;
    double lar = 0.0;
    for (int j = 0; j < 4; ++j) {
      nan |= isnan(t[j]);
      pinf |= (isinf(t[j]) && t[j] > 0);
      ninf |= (isinf(t[j]) && t[j] < 0);
      lar += t[j];
    }
    if (nan || (pinf && ninf)) lar = -INFINITY;
    const double u = u_in ? (double)u_in[c] : draw_uniform(seed, STREAM_HYPER, call, (uint32_t)c);
    acc_s = (log(u) < lar) ? 1 : 0;
    const double accept = lar < 0.0 ? exp(lar) : 1.0;
    // dual averaging, network.py:457-469 (constants :241-248)
    const double target = 0.95, gamma = tfc::k0p4, t0 = 10.0, kappa = 0.75;
    const double m = epoch + 1.0, mu = (double)logf((float)(100.0 * hyper_step0));   // tf.math.log of a python float: float32
    double hda = (double)da_state[c * 3 + 0], leb = (double)da_state[c * 3 + 1];
    hda = (1.0 - 1.0 / (m + t0)) * hda + (1.0 / (m + t0)) * (target - accept);
    const double logEps = mu - hda * sqrt(m) / gamma;
    const double mk = pow(m, -kappa);
    leb = (1.0 - mk) * leb + mk * logEps;
    da_state[c * 3 + 0] = (T)hda;
    da_state[c * 3 + 1] = (T)leb;
    if (m < burnin * 0.8) da_state[c * 3 + 2] = (T)exp(leb);
    if (stats) {
      stats[c * 2 + 0] = (T)lar;
      stats[c * 2 + 1] = (T)accept;
    }
  }
  __syncthreads();
  for (int j = threadIdx.x; j < H; j += blockDim.x) hyper[(size_t)c * H + j] = acc_s ? hc[j] : h0[j];
}

template <typename T>
void Launch<T>::hyper_eval(const ModelPlan& mp, int C, const T* theta_flat, const T* hyper,
                           const double* sse, long long N_total, T* logp, T* grad, cudaStream_t st) {
  k_hyper_eval<T><<<C, HT, 0, st>>>(mp, theta_flat, hyper, sse, N_total, logp, grad);
}
template <typename T>
void Launch<T>::hyper_step(const ModelPlan& mp, int C, const T* theta_flat, T* hyper, const double* sse,
                           long long N_total, uint64_t seed, uint64_t call, int L, double epoch,
                           double burnin, double hyper_step0, T* da_state, const T* mom_in,
                           const T* u_in, T* stats, cudaStream_t st) {
  k_hyper_step<T><<<C, HT, 0, st>>>(mp, theta_flat, hyper, sse, N_total, seed, call, L, epoch, burnin,
                                    hyper_step0, da_state, mom_in, u_in, stats);
}

template struct Launch<float>;
template struct Launch<double>;

}  // namespace tbnn
