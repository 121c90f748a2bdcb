"""Sample loader + posterior-predictive sweep with the reference's interface
(tensorBNN/predictor.py:15-155, :314-351): predictor(directoryPath, dtype, customLayerDict,
likelihood), predict(inputMatrix, n), extractParameters, extractHyperParameters,
parameterStatistics, hyperStatistics.  ``predict`` is one fused CUDA sweep over
(stored samples x test rows) instead of a Python loop over samples and layers;
``predict_moments`` (extension) returns the posterior-predictive mean and sd without
materialising the [S, out, M] tensor.  ``trainProbs`` / ``reweight`` (reference :157-273) evaluate the
prior + hyper-prior (+ likelihood) log-density of every stored sample in batches on the device -- the stored samples
become the "chains" of the batched hyper-target kernel; ``autocorrelation`` / ``autoCorrelationLength``
(reference :275-312) run emcee's published estimator as batched FFTs in torch (emcee itself is not needed)."""
import math
import warnings

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
        # constants of stateless layers that the reference's architecture.txt cannot carry (the Leaky_relu slope):
        # side-car "layer_params.txt" written by this package's network.train, lines "<layer index> alpha <value>"
        extra = {}
        side = os.path.join(os.path.dirname(path), "layer_params.txt")
        if os.path.exists(side):
            for line in open(side, "r"):
                f = line.split()
                if len(f) == 3 and f[1] == "alpha":
                    extra[int(f[0])] = float(f[2])
        self.layers, self._arch = [], []
        ti = 0
        for li, name in enumerate(names):
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
            elif name == "leakyrelu":
                if li not in extra:
                    warnings.warn("architecture has a leakyrelu layer but no recorded slope (layer_params.txt): "
                                  "using the reference's default alpha = 0.3")
                layer = cls(alpha=extra.get(li, 0.3))
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

    # ------------------------------------------------------------------ reweighting (reference :157-273)
    def _lik_spec(self, likelihood):
        return likelihood.spec() if likelihood is not None else ("fixed", 1.0)

    def _neg_log_weights(self, arch, trainX, trainY, n, likelihood, batch=256):
        """For every n-th stored sample m: -( likelihood term + sum_layers calculateHyperProbs(hypers_m, tensors_m) ),
        the quantity predictor.trainProbs / reweight accumulate (reference :176-202, :238-268), evaluated in batches
        with the stored samples as the chains of the batched hyper-target kernel (tbnn_hyper_logp_grad).

        Decisions where the reference has no runnable behaviour (SURVEY App. C style, recorded in DESIGN.md): the
        dense layers' calculateHyperProbs is the training-time formula (the reference indexes a numpy scalar there
        and raises); FixedGaussianLikelihood: multivariateLogProb with the constructor sd (as written, :190-194);
        BernoulliLikelihood: 0 (as written, :239-243); GaussianLikelihood: sigma = stored hyper**2 as in training
        (the reference raises KeyError('sd') on this path)."""
        idx, samples = self._select(n)
        hy = np.asarray(self.hypers, dtype=np.float32)[idx] if len(self.hypers) else np.zeros((len(idx), 0), np.float32)
        lik = self._lik_spec(likelihood)
        out = np.zeros(len(idx), dtype=np.float64)
        X = Y = None
        if likelihood is not None and lik[0] != "bernoulli":
            X = np.asarray(trainX, dtype=np.float64)
            D = self._arch[0][1]
            if X.ndim == 2 and X.shape[0] == D and X.shape[1] != D:
                X = X.T                      # the reference expects trainX as [D, N] here (it transposes twice, :168,:141)
            Y = np.asarray(trainY, dtype=np.float64).reshape(X.shape[0], -1)
        for b0 in range(0, len(idx), batch):
            sl = slice(b0, min(b0 + batch, len(idx)))
            B = sl.stop - sl.start
            eng = Engine(arch, lik, dtype=self.tdtype, chains=B, device=self.device)
            H = eng.H
            hb = np.zeros((B, H), dtype=np.float32)
            k = min(H, hy.shape[1])
            hb[:, :k] = hy[sl, :k]
            if lik[0] == "gaussian" and hy.shape[1] >= 1:
                hb[:, -1] = hy[sl, -1]
            sse = None
            if X is not None:
                eng.set_data(X, Y)
                _, _, stat = eng.logp_grad(samples[sl], hb)
                sse = stat
            else:
                # no data-dependent term (Bernoulli: 0 as written; no likelihood given): the handle still wants a data set
                eng.set_data(np.zeros((1, self._arch[0][1])), np.zeros((1, eng.predict(samples[:1], np.zeros((1, self._arch[0][1])))[0].shape[1])))
            val, _ = eng.hyper_logp_grad(samples[sl], hb, sse=sse)
            val = val.double().cpu().numpy()
            if lik[0] == "fixed" and likelihood is not None:
                sd = float(np.float32(likelihood.sd))
                sd = min(max(sd, float(np.float32(1e-8))), 1e8)
                nn = float(Y.size)
                val = val + (-0.5 * (2.0 * nn * math.log(sd) + sse.double().cpu().numpy() / sd ** 2
                                     + nn * math.log(float(np.float32(2 * math.pi)))))
            out[sl] = -val
        return out

    def trainProbs(self, trainX, trainY, n, likelihood):
        """Negative log probabilities of the stored samples under the architecture they were trained with
        (reference :157-202); the training likelihood is the one given to the constructor."""
        self.weightsTrain = self._neg_log_weights(self._arch, trainX, trainY, n,
                                                  self.likelihood if likelihood is not None else None)

    def reweight(self, architecture, trainX=None, trainY=None, n=1, likelihood=None):
        """Importance weights p(theta | new priors) / p(theta | training priors) of every n-th stored sample for the
        layer kinds listed in the file ``architecture`` (same sizes, e.g. "denseGaussian" instead of "dense"),
        normalised to sum to one (reference :204-273)."""
        if len(self.weightsTrain) == 0:
            self.trainProbs(trainX, trainY, n, likelihood)
        keep_layers, keep_arch = self.layers, self._arch
        self.loadArchitecture(architecture=architecture)
        new_arch = self._arch
        self.layers, self._arch = keep_layers, keep_arch
        self.weights = self._neg_log_weights(new_arch, trainX, trainY, n, likelihood)
        d = self.weightsTrain - self.weights
        weighting = np.exp(d - d.max())              # same ratios as exp(d) / sum(exp(d)), without overflow
        return weighting / np.sum(weighting)

    # ------------------------------------------------------------------ autocorrelation (reference :275-312)
    @staticmethod
    def _acf(x):
        """emcee.autocorr.function_1d for every row of x [R, T]: FFT autocorrelation normalised by lag 0."""
        T = x.shape[1]
        nfft = 1
        while nfft < T:
            nfft <<= 1
        f = torch.fft.rfft(x - x.mean(dim=1, keepdim=True), n=2 * nfft, dim=1)
        acf = torch.fft.irfft(f * torch.conj(f), n=2 * nfft, dim=1)[:, :T]
        return acf / acf[:, :1]

    @staticmethod
    def _integrated_time(acf, c=5.0):
        """emcee.autocorr.integrated_time (one walker, one dimension per row): tau = 2 cumsum(acf) - 1 at Sokal's
        automatic window, the first M with M >= c tau(M)."""
        taus = 2.0 * torch.cumsum(acf, dim=1) - 1.0
        T = acf.shape[1]
        m = torch.arange(T, device=acf.device, dtype=acf.dtype)[None, :] < c * taus
        # np.argmin of a boolean row = first False; all True -> window T - 1
        first_false = torch.where(m.all(dim=1), torch.full((acf.shape[0],), T - 1, device=acf.device),
                                  torch.argmin(m.to(torch.int8), dim=1))
        return taus.gather(1, first_false[:, None])[:, 0]

    def _traces(self, inputData):
        """Predictive traces [M * out, S] on the device: one row per (test point, output), one column per sample."""
        eng = self._get_engine()
        _, samples = self._select(1)
        out, _ = eng.predict(samples, np.asarray(inputData), want_out=True)      # [S, out, M]
        S = out.shape[0]
        return out.reshape(S, -1).t().contiguous().double()

    def autocorrelation(self, inputData, nMax):
        """Autocorrelation function of the predictive traces averaged over the test points whose integrated time is
        finite, truncated to nMax lags (reference :275-292)."""
        x = self._traces(inputData)
        acf = self._acf(x)
        tau = self._integrated_time(acf)
        ok = ~torch.isnan(tau)
        val = acf[ok].sum(dim=0) / ok.sum()
        val = val.cpu().numpy()
        return val[:nMax] if nMax < len(val) else val

    def autoCorrelationLength(self, inputData, nMax):
        """Mean integrated autocorrelation time of the predictive traces (reference :294-312)."""
        x = self._traces(inputData)
        tau = self._integrated_time(self._acf(x))
        ok = ~torch.isnan(tau)
        val = float(tau[ok].sum() / ok.sum())
        if val > nMax:
            print("Correlation time is greater than maximum accepted value.")
        return val
