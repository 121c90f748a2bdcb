"""Activation layers with the reference's interface (tensorBNN/activationFunctions.py).
Parameter-free: Exp, Relu, Sigmoid, Tanh, Elu, Softmax; constant slope: Leaky_relu;
sampled slopes: Prelu, SquarePrelu.  Inside the sampler / predictor they are fused into
the dense layer that precedes them; the per-layer ``predict`` is routed through the CUDA
engine on an identity dense layer (exact: 1*x + 0)."""
import math

import numpy as np
import torch

from .layer import Layer, as_tensor, single_layer_engine, to_torch_dtype, _log_normal_1d


class _Activation(Layer):
    _name = "name"

    def __init__(self, inputDims=None, outputDims=None):
        self.numTensors = 0
        self.numHyperTensors = 0
        self.name = self._name
        self.inputDims = inputDims
        self.outputDims = outputDims

    def spec(self):
        return (self._name,)

    def _apply(self, inputTensor, spec, extra=None):
        a = inputTensor if isinstance(inputTensor, torch.Tensor) else torch.as_tensor(np.asarray(inputTensor))
        dt = a.dtype if a.dtype in (torch.float32, torch.float64) else torch.float32
        a = a.to(dt)
        if a.dim() == 1:
            a = a.unsqueeze(0)
        w = a.shape[0]
        eng = single_layer_engine([("denseGaussian", w, w), spec], ("fixed", 1.0), dt)
        parts = [torch.eye(w, dtype=dt).reshape(-1), torch.zeros(w, dtype=dt)]
        if extra is not None:
            parts.append(as_tensor(extra, dt).reshape(-1))
        flat = torch.cat(parts).reshape(1, -1)
        out, _ = eng.predict(flat, a.t().contiguous(), want_out=True)
        return out[0]

    def predict(self, inputTensor, _):
        return self._apply(inputTensor, self.spec())


class Exp(_Activation):
    """Exponential activation (reference activationFunctions.py:14-24)."""
    _name = "Exp"


class Relu(_Activation):
    """(reference :27-37)"""
    _name = "relu"


class Sigmoid(_Activation):
    """(reference :40-50)"""
    _name = "sigmoid"


class Tanh(_Activation):
    """(reference :53-63)"""
    _name = "tanh"


class Elu(_Activation):
    """(reference :66-76)"""
    _name = "elu"


class Softmax(_Activation):
    """(reference :79-89).  The reference normalises over the LAST axis of its [width, N]
    tensor, i.e. over rows, which couples all training rows; it is off the
    sampling path and not available in the fused kernels."""
    _name = "softmax"

    def spec(self):
        raise NotImplementedError("Softmax (normalised over rows in the reference) is not supported "
                                  "by the CUDA sampler")

    def predict(self, inputTensor, _):
        raise NotImplementedError("Softmax is not supported by the CUDA path")


class Leaky_relu(_Activation):
    """Leaky relu with a CONSTANT slope.  The reference registers the python float alpha as a
    sampled state with no gradient and no hypers, which mis-aligns every later layer
    (SURVEY App. C Q6); here it carries no state."""
    _name = "leakyrelu"

    def __init__(self, alpha=0.3, inputDims=None, outputDims=None, activation=None):
        _Activation.__init__(self, inputDims, outputDims)
        if activation is not None:
            alpha = activation
        self.alpha = float(alpha)

    def spec(self):
        return ("leakyrelu", self.alpha)

    def calculateProbs(self, *args):
        return 0.0

    def updateParameters(self, *args):
        pass


class _SlopeActivation(_Activation):
    def __init__(self, inputDims, outputDims=None, dtype=np.float32, alpha=0.2, activation=None, seed=1):
        self.numTensors = 1
        self.inputDims = inputDims
        self.outputDims = outputDims
        self.dtype = dtype
        self.tdtype = to_torch_dtype(dtype)
        self.seed = seed
        self.name = self._name
        if activation is None:
            self.parameters = [alpha * torch.ones(int(inputDims), dtype=self.tdtype)]
        else:
            self.parameters = [as_tensor(activation, self.tdtype).reshape(-1)]

    def spec(self):
        return (self._name, int(self.parameters[0].numel()))

    def predict(self, inputTensor, slopes):
        s = as_tensor(slopes[0], self.tdtype).reshape(-1)
        return self.expand(self._apply(inputTensor, (self._name, int(s.numel())), extra=s))

    def updateParameters(self, slopes):
        self.parameters = [slopes[0]]


class Prelu(_SlopeActivation):
    """Prelu with sampled slopes (reference :117-271): exponential prior with a sampled rate."""
    _name = "prelu"

    def __init__(self, inputDims, outputDims=None, dtype=np.float32, alpha=0.2, activation=None, seed=1):
        _SlopeActivation.__init__(self, inputDims, outputDims, dtype, alpha, activation, seed)
        self.numHyperTensors = 1
        self.hyperRate = 0.3
        self.hypers = [torch.tensor(0.3, dtype=torch.float32).to(self.tdtype)]      # tf.cast(0.3, dtype): float32 first (Q14)

    def exponentialLogProb(self, rate, x):
        rate = abs(float(rate))
        return -rate * x + math.log(rate)

    def calculateProbs(self, hypers, slopes=None):
        """sum_i e(rate, a_i) with the rate taken from the passed hyper slice (Q4 minimal patch)."""
        if slopes is None:
            hypers, slopes = self.hypers, hypers
        a = as_tensor(slopes[0] if isinstance(slopes, (list, tuple)) else slopes, torch.float64)
        return float(torch.sum(self.exponentialLogProb(as_tensor(hypers[0], torch.float64), a)))

    def calculateHyperProbs(self, hypers, slopes):
        a = torch.abs(as_tensor(slopes[0], torch.float64))
        r = float(as_tensor(hypers[0], torch.float64))
        return float(self.exponentialLogProb(self.hyperRate, r)
                     + torch.sum(self.exponentialLogProb(r, a)))

    def updateHypers(self, hypers):
        self.hypers = [torch.clamp(as_tensor(hypers[0], self.tdtype), min=0.01)]


class SquarePrelu(_SlopeActivation):
    """Prelu whose slope is the square of the sampled value (reference :274-433)."""
    _name = "squareprelu"

    def __init__(self, inputDims, outputDims=None, dtype=np.float32, alpha=0.2, activation=None, seed=1):
        _SlopeActivation.__init__(self, inputDims, outputDims, dtype, alpha, activation, seed)
        self.numHyperTensors = 2
        self.hypers = [torch.tensor(0.0, dtype=self.tdtype), torch.tensor(0.3, dtype=torch.float32).to(self.tdtype)]   # Q14

    @staticmethod
    def _mvlp(sd, mean, x):
        sd = min(max(float(sd), 1e-8), 1e8)
        return -0.5 * (2.0 * math.log(sd) + float(torch.sum(((x - mean) / sd) ** 2)) + math.log(2.0 * math.pi))

    def calculateProbs(self, hypers, slopes=None):
        """multivariateLogProb(sd, mean, slopes) on the UN-squared slopes, hypers from the passed
        slice (reference :341-346 with the Q4 minimal patch)."""
        if slopes is None:
            hypers, slopes = self.hypers, hypers
        a = as_tensor(slopes[0] if isinstance(slopes, (list, tuple)) else slopes, torch.float64)
        return self._mvlp(as_tensor(hypers[1], torch.float64), float(as_tensor(hypers[0], torch.float64)), a)

    def calculateHyperProbs(self, hypers, slopes):
        a2 = as_tensor(slopes[0], torch.float64) ** 2
        mean = float(as_tensor(hypers[0], torch.float64))
        sd = float(as_tensor(hypers[1], torch.float64))
        return self._mvlp(sd, mean, a2) + _log_normal_1d(mean, 0.0, 0.3) + _log_normal_1d(sd, 0.3, 0.1)

    def updateHypers(self, hypers):
        self.hypers = [hypers[0], hypers[1]]
