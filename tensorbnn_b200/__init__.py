"""B200-native HMC sampler for Bayesian neural networks behind TensorBNN's Python API.

Modules mirror the reference package (network, layer, activationFunctions, likelihood, metrics,
paramAdapter, predictor); `engine` is the torch-tensor wrapper over the C ABI of libtbnn.so
(include/tbnn.h); `workloads` holds the synthetic benchmark configurations.  There is no CPU
fallback: every sampler / predictor computation runs in hand-written CUDA for sm_100a."""
__version__ = "0.1.0"
