"""Sample loader + posterior-predictive sweep with the reference's interface
(tensorBNN/predictor.py:15-155, :314-351): predictor(directoryPath, dtype, customLayerDict,
likelihood), predict(inputMatrix, n), extractParameters, extractHyperParameters,
parameterStatistics, hyperStatistics.  ``predict`` is one fused CUDA sweep over
(stored samples x test rows) instead of a Python loop over samples and layers;
``predict_moments`` (extension) returns the posterior-predictive mean and sd without
materialising the [S, out, M] tensor.  reweight / autocorrelation are outside the
accelerated hot path."""
import math

import os

import numpy as np
import torch

from .activationFunctions import (Elu, Exp, Leaky_relu, Prelu, Relu, Sigmoid, Softmax, SquarePrelu, Tanh)
from .engine import Engine
from .layer import DenseLayer, GaussianDenseLayer, to_torch_dtype
from .likelihood import GaussianLikelihood


class predictor(object):
    def __init__(self, directoryPath, dtype, customLayerDict={}, likelihood=None, device=None):
        self.layerDict = {"Exp": Exp, "relu": Relu, "sigmoid": Sigmoid, "tanh": Tanh, "elu": Elu,
                          "softmax": Softmax, "leakyrelu": Leaky_relu, "prelu": Prelu,
                          "squareprelu": SquarePrelu, "dense": DenseLayer,
                          "denseGaussian": GaussianDenseLayer}
        if customLayerDict:
            raise NotImplementedError("custom TensorFlow layers (customLayerDict) cannot be honoured "
                                      "without TensorFlow; only the built-in layers are supported")
        self.directoryPath = directoryPath       # string-concatenated: must end with '/' (reference :48)
        self.dtype = dtype
        self.tdtype = to_torch_dtype(dtype)
        self.device = device
        self.loadNetworks()
        self.loadArchitecture()
        self.likelihood = likelihood if likelihood is not None else GaussianLikelihood(sd=0.1)
        self.weightsTrain = []
        self._engine = None

    def _load_values(self, name, count, ndmin):
        """`count` float32 values of file `name` (.txt in the reference's layout): from the binary side-car
        `name.f32` when this package's writer left a complete one, else parsed from the text."""
        side = self.directoryPath + name + ".f32"
        if os.path.exists(side) and os.path.getsize(side) >= 4 * count:
            return np.fromfile(side, dtype="<f4", count=count), True
        return np.loadtxt(self.directoryPath + name + ".txt", dtype=np.float32, ndmin=ndmin), False

    def loadNetworks(self):
        """Parses summary.txt and the per-tensor text files (reference :43-113; values are read as
        float32 regardless of dtype, Q10)."""
        summary = []
        with open(self.directoryPath + "summary.txt", "r") as fh:
            for line in fh:
                summary.append(line.split())
        numNetworks = int(summary[-2][0])
        numFiles = int(summary[-2][1])
        numMatrices = int(summary[-2][2])
        numHypers = int(summary[-1][0])
        numNetworks //= numFiles
        total = numNetworks * numFiles
        matrices, flat_parts, shapes = [], [], []
        for n in range(numMatrices):
            d1 = int(summary[n][0])
            d2 = int(summary[n][1]) if len(summary[n]) == 2 else 1
            shapes.append((d1, d2) if len(summary[n]) == 2 else (d1,))
            w0 = np.zeros((total, d1, d2), dtype=np.float32)
            for m in range(numFiles):
                w, binary = self._load_values("%d.%d" % (n, m), numNetworks * d1 * d2, 2)
                w0[m * numNetworks:(m + 1) * numNetworks] = \
                    w.reshape(numNetworks, d1, d2) if binary else w[:numNetworks * d1, :d2].reshape(numNetworks, d1, d2)
            matrices.append(torch.as_tensor(w0).to(self.tdtype))
            flat_parts.append(w0.reshape(total, -1))
        hypers = []
        if numHypers > 0:
            for m in range(numFiles):
                w, _ = self._load_values("hypers%d" % m, numNetworks * numHypers, 1)
                for k in range(numNetworks):
                    hypers.append(w[numHypers * k:numHypers * (k + 1)])
        self.numNetworks = total
        self.numMatrices = numMatrices
        self.matrices = matrices
        self.hypers = hypers
        self.vectors = [v for v in np.concatenate(flat_parts, axis=1)] if flat_parts else []
        self._flat = np.concatenate(flat_parts, axis=1) if flat_parts else np.zeros((total, 0), np.float32)
        self._shapes = shapes

    def loadArchitecture(self, architecture=None):
        """Rebuilds the layer list from architecture.txt (reference :115-130); layer sizes come from
        the tensor shapes recorded in summary.txt."""
        path = self.directoryPath + "architecture.txt" if architecture is None else architecture
        names = [line.replace("\n", "") for line in open(path, "r") if line.strip()]
        self.layers, self._arch = [], []
        ti = 0
        for name in names:
            if name not in self.layerDict:
                raise KeyError("unknown layer name %r in architecture file" % name)
            cls = self.layerDict[name]
            if name in ("dense", "denseGaussian"):
                out_d, in_d = self._shapes[ti]
                layer = cls(in_d, out_d, weights=np.zeros((out_d, in_d)), biases=np.zeros((out_d, 1)),
                            dtype=self.tdtype)
                ti += 2
            elif name in ("prelu", "squareprelu"):
                layer = cls(self._shapes[ti][0], dtype=self.tdtype)
                ti += 1
            else:
                layer = cls(inputDims=1, outputDims=1)
            self.layers.append(layer)
            self._arch.append(layer.spec())

    def _get_engine(self):
        if self._engine is None:
            self._engine = Engine(self._arch, ("fixed", 1.0), dtype=self.tdtype, chains=1, device=self.device)
        return self._engine

    def _select(self, n):
        idx = np.arange(0, self.numNetworks, n)
        return idx, np.ascontiguousarray(self._flat[idx])

    def predict(self, inputMatrix, n=1):
        """Predictions of every n-th stored network: a list of ceil(S/n) arrays [out, M]
        (reference :132-155)."""
        eng = self._get_engine()
        idx, samples = self._select(n)
        out, _ = eng.predict(samples, np.asarray(inputMatrix), want_out=True)
        res = out.cpu().numpy()
        return [res[i] for i in range(len(idx))]

    def predict_moments(self, inputMatrix, n=1):
        """Posterior-predictive mean and sd per test row, [out, M] each, fused on the device."""
        eng = self._get_engine()
        idx, samples = self._select(n)
        _, mom = eng.predict(samples, np.asarray(inputMatrix), want_out=False, want_moments=True)
        mom = mom.cpu().numpy()
        return mom[1], np.sqrt(np.maximum(mom[2], 0.0) / max(len(idx), 1))

    def extractParameters(self):
        return self.matrices

    def extractHyperParameters(self):
        return np.array(self.hypers)

    def parameterStatistics(self):
        means = [np.mean(m.numpy(), axis=0) for m in self.matrices]
        sds = [np.std(m.numpy(), axis=0) for m in self.matrices]
        return means, sds

    def hyperStatistics(self):
        hy = np.array(self.hypers)
        return np.mean(hy, axis=0), np.std(hy, axis=0)

    def reweight(self, *args, **kwargs):
        raise NotImplementedError("predictor.reweight is outside the accelerated hot path (SURVEY 8f, f3)")

    def autocorrelation(self, *args, **kwargs):
        raise NotImplementedError("autocorrelation needs emcee, which is unavailable; SURVEY 8f, f4")

    autoCorrelationLength = autocorrelation
