// plan.h -- model plan shared by host planner and device kernels.
//
// A network (network.py:173-191) is normalised into "blocks": one dense layer
// (layer.py:266-279) followed by at most one activation (activationFunctions.py).
// All per-chain vectors (theta, momentum, gradient) live on the device in a PADDED
// layout so every smem/global row is float4-addressable and free of bank conflicts:
//   W_l  : [out_p][ld_in]   (out_p = round4(out), ld_in = leading dim of the input buffer)
//   b_l  : [out_p]
//   s_l  : [out_p]          (prelu / squareprelu slopes; absent otherwise)
// Padding entries are exactly zero and stay zero (see engine.cuh).
#pragma once
#include <stdint.h>

namespace tbnn {

constexpr int MAXB = 8;     // dense blocks per network
constexpr int NT = 256;     // threads per CTA of the tile engine

enum Act { ACT_NONE = 0, ACT_RELU, ACT_TANH, ACT_SIGMOID, ACT_EXP, ACT_ELU, ACT_LEAKY, ACT_PRELU,
           ACT_SQPRELU };
enum Prior { PRIOR_CAUCHY = 0, PRIOR_GAUSS = 1 };
enum Lik { LIK_GAUSS = 0, LIK_FIXED = 1, LIK_BERN = 2 };

__host__ __device__ inline bool act_has_slopes(int a) { return a == ACT_PRELU || a == ACT_SQPRELU; }
__host__ __device__ inline bool act_keeps_z(int a) {
  return a == ACT_PRELU || a == ACT_SQPRELU || a == ACT_LEAKY;
}

struct BlockPlan {
  int in, out, in_p, out_p;
  int ld_in, ld_out;   // leading dims (elements) of the input / output activation buffers
  int prior, act;
  double alpha;        // leaky-relu slope
  int pw, pb, ps;      // padded offsets of W, b, slopes (ps = -1: none)
  int fw, fb, fs;      // flat offsets (network.states order)
  int hw, ha;          // hyper offsets: 4 dense hypers; activation hypers (-1: none)
  // smem (per tile) offsets, in elements
  int offS;            // output buffer  S_l = act(z)      [TRp][ld_out]
  int offZ;            // z buffer (only if act_keeps_z)    [TRp][ld_out]  (-1: none)
  int offD;            // narrow-tail plans only: dz buffer [RB][ld_out] of blocks >= 1 (-1: none)
  int ksplit;          // split-K factor of the forward GEMM
};

struct ModelPlan {
  int nb;
  BlockPlan b[MAXB];
  int D, OUT;          // input / output widths
  int D_p, ld0;        // padded input width and leading dim of the X tile
  int P, H, Ppad;
  int lik;
  double fixed_sd;
  int lik_h;           // hyper index of the Gaussian likelihood sd (-1: none)
  // tile geometry / smem layout (elements of T)
  int TR;              // rows per tile (multiple of 4)
  int offX;            // X tile  [TR][ld0]
  int offDa, offDb;    // dZ ping-pong [TR][ldmax]
  int offScr;          // split-K scratch
  int offW;            // weights (Ppad) -- -1: read from global
  int offG;            // gradient accumulators (Ppad)
  int offRed;          // block-reduction scratch (64 doubles)
  int smem_elems;      // total
  int ldmax;
};

}  // namespace tbnn

// Constants exactly as TensorFlow materialises them in the reference (quirk Q14):
// tf.cast(python_float, dtype) converts to a float32 tensor FIRST, so 2*pi and the 1e-8 clamp of
// multivariateLogProb (BNN_functions.py:23-24,30), FixedGaussianLikelihood's sd (likelihood.py:161), the
// hyper-prior locations/scales (layer.py:137-153,318-334; activationFunctions.py:144-145,301-306: float32
// distributions) and the dual-averaging constants (network.py:241-248) carry float32 rounding even in a
// float64 network.  Invisible in float32; needed for 1e-10 agreement with the reference's arithmetic in fp64.
namespace tfc {
constexpr double kLog2PiCast = 1.8378770942368803;    // log((double)(float)(2*pi)): k*log(2pi) of multivariateLogProb
constexpr double kLog2Pi = 1.8378770664093453;        // exact: inside tfd.MultivariateNormalDiag.log_prob
constexpr double kClampLo = (double)1e-8f;            // 9.99999993922529e-09
constexpr double kClampHi = 1e8;
constexpr double k0p1 = (double)0.1f, k0p2 = (double)0.2f, k0p3 = (double)0.3f, k0p4 = (double)0.4f;
constexpr double kSqrtHalf = (double)0.70710678118654757f;   // 0.5**0.5 -> float32
}  // namespace tfc

