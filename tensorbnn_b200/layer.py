"""Dense layers with the reference's interface (tensorBNN/layer.py): same class names,
constructor arguments, attributes (numTensors, numHyperTensors, name, parameters, hypers)
and methods (predict, calculateProbs, calculateHyperProbs, sample, expand).

The objects only HOLD state (torch tensors); inside network.train / network.predict /
predictor.predict the whole network runs in the fused CUDA kernels.  The per-layer methods
are kept for drop-in compatibility and are routed through the same C ABI on a one-layer
network (no torch / CPU arithmetic fallback)."""
import math

import numpy as np
import torch


def to_torch_dtype(dtype):
    """Accepts torch / numpy dtypes and anything whose name mentions float32 / float64
    (user scripts written for the reference pass tf.float32)."""
    if isinstance(dtype, torch.dtype):
        if dtype in (torch.float32, torch.float64):
            return dtype
        raise ValueError("dtype must be float32 or float64")
    s = str(getattr(dtype, "name", dtype))
    if "float64" in s or "double" in s:
        return torch.float64
    if "float32" in s or "single" in s or s == "float":
        return torch.float32
    try:
        return to_torch_dtype(np.dtype(dtype))
    except Exception:
        raise ValueError("dtype %r is not float32 / float64" % (dtype,))


def as_tensor(x, dtype):
    if isinstance(x, torch.Tensor):
        return x.detach().to(dtype)
    return torch.as_tensor(np.asarray(x), dtype=dtype)


def _log_normal_1d(v, m, s):
    return -0.5 * ((v - m) / s) ** 2 - math.log(s) - 0.5 * math.log(2.0 * math.pi)


_ENGINES = {}


def single_layer_engine(arch, lik, dtype):
    """Cached one-layer engines behind the per-layer compatibility methods."""
    from .engine import Engine
    key = (tuple(arch), tuple(lik), dtype, torch.cuda.current_device() if torch.cuda.is_available() else -1)
    if key not in _ENGINES:
        _ENGINES[key] = Engine(list(arch), lik, dtype=dtype, chains=1)
    return _ENGINES[key]


class Layer(object):
    """Base class (reference layer.py:10-98)."""

    def __init__(self, inputDims, outputDims, weights=None, biases=None, activation=None,
                 dtype=np.float32, alpha=0, seed=1):
        self.numTensors = 0
        self.numHyperTensors = 0
        self.inputDims = inputDims
        self.outputDims = outputDims
        self.dtype = dtype
        self.seed = seed
        self.name = "name"

    # -- description used by the CUDA engine
    def spec(self):
        raise NotImplementedError(
            "layer %r has no CUDA implementation: custom TensorFlow layers (customLayerDict) cannot be "
            "honoured without TensorFlow; only the built-in layer vocabulary is supported" % self.name)

    def calculateProbs(self, *args):
        return 0.0

    def calculateHyperProbs(self, hypers, tensors):
        return 0.0

    def expand(self, current):
        """Pads rank <= 1 tensors to rank 2 (reference layer.py:72-86)."""
        t = current if isinstance(current, torch.Tensor) else torch.as_tensor(np.asarray(current))
        while t.dim() < 2:
            t = t.unsqueeze(0)
        return t

    def predict(self, inputTensor, tensors):
        pass


class _DenseBase(Layer):
    _kind = "dense"
    _init_hypers = (0.0, 0.5 ** 0.5, 0.0, 0.5 ** 0.5)
    _hyperprior = (0.0, 0.2, 0.5 ** 0.5, 0.5)        # loc mean, loc sd, scale mean, scale sd

    def __init__(self, inputDims, outputDims, weights=None, biases=None, dtype=np.float32, seed=1):
        self.numTensors = 2
        self.numHyperTensors = 4
        self.inputDims = inputDims
        self.outputDims = outputDims
        self.dtype = dtype
        self.tdtype = to_torch_dtype(dtype)
        self.seed = seed
        self.name = self._kind
        # tf.cast([[...]], dtype) of python floats passes through float32 (layer.py:156-158; quirk Q14)
        self.hypers = torch.tensor([[v] for v in self._init_hypers], dtype=torch.float32).to(self.tdtype)   # [4,1]
        if weights is None:
            self.parameters = self.sample()
        else:
            self.parameters = [as_tensor(weights, self.tdtype).reshape(outputDims, inputDims),
                               as_tensor(biases, self.tdtype).reshape(outputDims, 1)]

    def spec(self):
        return (self._kind, int(self.inputDims), int(self.outputDims))

    def sample(self):
        """W ~ N(h0, sqrt(2/out)), b ~ N(h2, sqrt(2/out)) with seeds seed / seed+1
        (reference layer.py:244-264; the TF generator itself is not reproducible here)."""
        sd = (2.0 / self.outputDims) ** 0.5
        g = torch.Generator().manual_seed(int(self.seed))
        w = self.hypers[0].item() + sd * torch.randn(self.outputDims, self.inputDims, generator=g,
                                                      dtype=torch.float64)
        g = torch.Generator().manual_seed(int(self.seed) + 1)
        b = self.hypers[2].item() + sd * torch.randn(self.outputDims, 1, generator=g, dtype=torch.float64)
        return [w.to(self.tdtype), b.to(self.tdtype)]

    # -- per-layer compatibility methods, executed by the CUDA engine on a one-layer network
    def predict(self, inputTensor, tensors):
        """W @ A + b for A [in, N] (reference layer.py:266-279)."""
        w = self.expand(as_tensor(tensors[0], self.tdtype))
        b = as_tensor(tensors[1], self.tdtype).reshape(-1)
        out_dim, in_dim = w.shape
        eng = single_layer_engine([(self._kind, in_dim, out_dim)], ("fixed", 1.0), self.tdtype)
        a = as_tensor(inputTensor, self.tdtype)
        flat = torch.cat([w.reshape(-1), b]).reshape(1, -1)
        out, _ = eng.predict(flat, a.t().contiguous(), want_out=True)
        return out[0]

    def calculateHyperProbs(self, hypers, tensors):
        """Hyper-priors + prior (reference layer.py:199-242 / :379-422)."""
        w = self.expand(as_tensor(tensors[0], self.tdtype))
        b = as_tensor(tensors[1], self.tdtype).reshape(-1)
        out_dim, in_dim = w.shape
        eng = single_layer_engine([(self._kind, in_dim, out_dim)], ("bernoulli",), self.tdtype)
        if eng.N == 0:
            eng.set_data(np.zeros((1, in_dim)), np.zeros((1, out_dim)))
        hy = torch.stack([as_tensor(h, self.tdtype).reshape(()) for h in hypers]).reshape(1, 4)
        flat = torch.cat([w.reshape(-1), b]).reshape(1, -1)
        lp, _ = eng.hyper_logp_grad(flat, hy)
        return lp[0]

    def calculateProbs(self, hypers, tensors):
        """Prior of W and b given the hypers (reference layer.py:166-197 / :346-377)."""
        total = self.calculateHyperProbs(hypers, tensors)
        hv = [float(as_tensor(h, torch.float64).reshape(())) for h in hypers]
        lm, ls, sm, ss = self._hyperprior
        hp = (_log_normal_1d(hv[0], lm, ls) + _log_normal_1d(hv[1] ** 2, sm, ss)
              + _log_normal_1d(hv[2], lm, ls) + _log_normal_1d(hv[3] ** 2, sm, ss))
        return total - hp


class CauchyDenseLayer(_DenseBase):
    """Dense layer with the reference's Cauchy prior (layer.py:101-279)."""
    _kind = "dense"
    _init_hypers = (0.0, 0.5 ** 0.5, 0.0, 0.5 ** 0.5)
    _hyperprior = (0.0, 0.2, 0.5 ** 0.5, 0.5)


class GaussianDenseLayer(_DenseBase):
    """Dense layer with a Gaussian prior (layer.py:282-459)."""
    _kind = "denseGaussian"
    _init_hypers = (0.0, 1.0, 0.0, 1.0)
    _hyperprior = (0.0, 0.1, 1.0, 0.1)


DenseLayer = CauchyDenseLayer  # reference layer.py:461
