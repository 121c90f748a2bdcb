"""The ``network`` container and sampler driver with the reference's interface
(tensorBNN/network.py): __init__(dtype, inputDims, trainX, trainY, validateX, validateY),
add, setupMCMC, train, predict and the attributes users read (states, hyperStates, layers,
step_size, leapfrog).  One epoch = one main HMC transition (network.py:394-411) + one hyper
transition with dual averaging (:442-471) + one adapter update (:603-607); every
log-posterior / gradient / leapfrog / Metropolis operation runs in libtbnn.so on the GPU --
the host loop only sequences launches, reads four scalars per epoch and writes samples.

Extensions (keyword-only, reference defaults): ``chains`` (independent chains batched into
one launch), ``device``, ``seed``.
"""
import math
import os
import time

import numpy as np
import torch

from .engine import Engine
from .layer import as_tensor, to_torch_dtype
from .paramAdapter import paramAdapter


class network(object):
    def __init__(self, dtype, inputDims, trainX, trainY, validateX, validateY, *ignored,
                 chains=1, device=None, seed=50):
        # README.md:76 passes two extra positionals (mean, sd) that the reference signature
        # rejects; they are accepted and ignored.
        self.dtype = dtype
        self.tdtype = to_torch_dtype(dtype)
        self.iteration = None
        self.inputDims = int(inputDims)
        self.chains = int(chains)
        self.device = device
        self.seed = int(seed)
        self.trainX = as_tensor(trainX, self.tdtype).reshape(len(trainX), self.inputDims)
        self.trainY = as_tensor(trainY, self.tdtype)
        self.validateX = as_tensor(validateX, self.tdtype).reshape(len(validateX), self.inputDims)
        self.validateY = as_tensor(validateY, self.tdtype)
        self.states = []        # weight / bias / slope tensors (views into the flat device state once training)
        self.hyperStates = []   # hyper parameter tensors
        self.layers = []
        self.currentInnerStep = None
        self.metricList = []
        self.likelihood = None
        self._engine = None
        self._pred_engine = None
        self._theta = None      # [C, P] device
        self._hyper = None      # [C, H] device
        self._lik_appended = False

    # ------------------------------------------------------------------ construction
    def add(self, layer, parameters=None):
        """Adds a layer (reference network.py:173-191)."""
        self.layers.append(layer)
        if layer.numTensors > 0:
            src = layer.parameters if parameters is None else parameters
            for st in src:
                self.states.append(as_tensor(st, self.tdtype))
        if layer.numHyperTensors > 0:
            for st in layer.hypers:
                self.hyperStates.append(as_tensor(st, self.tdtype).reshape(-1))

    def arch_spec(self):
        return [layer.spec() for layer in self.layers]

    def setupMCMC(self, stepSizeStart=1e-3, stepSizeMin=1e-4, stepSizeMax=1e-2, stepSizeOptions=40,
                  leapfrogStart=1000, leapfogMin=100, leapFrogMax=10000, leapfrogIncrement=1,
                  hyperStepSize=1e-2, hyperLeapfrog=100, burnin=1000, cores=4, averagingSteps=10, a=4,
                  delta=0.1, strikes=5, randomSteps=10, dualAveraging=False):
        """Sets up the samplers (reference network.py:193-278; argument names as in the reference,
        including ``leapfogMin`` / ``leapFrogMax``).  ``cores``, ``strikes`` and ``dualAveraging`` are
        accepted and unused, as in the reference."""
        self.adapt = paramAdapter(stepSizeStart, leapfrogStart, stepSizeMin, stepSizeMax, stepSizeOptions,
                                  leapfogMin, leapFrogMax, leapfrogIncrement, averagingSteps,
                                  burnin / averagingSteps, a=a, delta=delta, cores=cores, strikes=strikes,
                                  randomSteps=randomSteps, device=self.device)
        self.step_size = float(stepSizeStart)
        self.leapfrog = int(leapfrogStart)
        self.cores = cores
        self.burnin = burnin
        self.target = 0.95
        self.gamma, self.t0, self.kappa = 0.4, 10.0, 0.75
        self.h = 0.0
        self.logEpsilonBar = 0.0
        self.mu = math.log(100 * hyperStepSize)
        self.dualAveraging = dualAveraging
        self.hyperStepSize0 = float(hyperStepSize)
        self.hyper_step_size = float(hyperStepSize)
        self.hyperLeapfrog = int(hyperLeapfrog)

    # ------------------------------------------------------------------ device state
    def _lik_spec(self, likelihood):
        return likelihood.spec() if likelihood is not None else ("fixed", 1.0)

    def _make_engine(self, likelihood):
        eng = Engine(self.arch_spec(), self._lik_spec(likelihood), dtype=self.tdtype, chains=self.chains,
                     device=self.device)
        eng.set_data(self.trainX, self.trainY.reshape(self.trainX.shape[0], -1))
        return eng

    def _flat_states(self):
        return torch.cat([s.reshape(-1).to(self.tdtype).cpu() for s in self.states])

    def _bind_state_views(self):
        """network.states / hyperStates become views into the flat device state of chain 0
        (or [C, ...] views when chains > 1), so user reads always see the current sample."""
        shapes = [tuple(s.shape) for s in self.states]
        off, views = 0, []
        for sh in shapes:
            n = int(np.prod(sh))
            v = self._theta[:, off:off + n]
            views.append(v[0].reshape(sh) if self.chains == 1 else v.reshape((self.chains,) + sh))
            off += n
        self.states = views
        hv = []
        for j in range(self._hyper.shape[1]):
            v = self._hyper[:, j:j + 1]
            hv.append(v[0] if self.chains == 1 else v)
        self.hyperStates = hv

    def _ensure_device_state(self, likelihood):
        if self._engine is not None:
            return
        self._engine = self._make_engine(likelihood)
        eng = self._engine
        flat = self._flat_states()
        if flat.numel() != eng.P:
            raise ValueError("states hold %d values but the network has %d parameters" % (flat.numel(), eng.P))
        hy = torch.cat([h.reshape(-1).to(self.tdtype).cpu() for h in self.hyperStates])
        if hy.numel() != eng.H:
            raise ValueError("hyperStates hold %d values but the network has %d hyper parameters"
                             % (hy.numel(), eng.H))
        self._theta = flat.reshape(1, -1).repeat(self.chains, 1).to(eng.dev).contiguous()
        if self.chains > 1:
            # independent chains start from jittered copies of the given state
            g = torch.Generator().manual_seed(self.seed)
            jit = 0.01 * torch.randn(self.chains - 1, eng.P, generator=g, dtype=torch.float64)
            self._theta[1:] += jit.to(self._theta)
        self._hyper = hy.reshape(1, -1).repeat(self.chains, 1).to(eng.dev).contiguous()
        self._bind_state_views()

    # ------------------------------------------------------------------ prediction / probabilities
    def predict(self, train, *argv):
        """Network output [out, N] on the training (train=True) or validation data
        (reference network.py:141-171), computed by the CUDA forward kernel."""
        tensors = self.states if len(argv) == 0 else argv[0]
        x = self.trainX if train else self.validateX
        if self._pred_engine is None:
            self._pred_engine = Engine(self.arch_spec(), ("fixed", 1.0), dtype=self.tdtype, chains=1,
                                       device=self.device)
        eng = self._pred_engine
        if self.chains > 1 and tensors is self.states:
            flat = self._theta
        else:
            flat = torch.cat([as_tensor(t, self.tdtype).reshape(-1).to(eng.dev) for t in tensors]).reshape(1, -1)
        key = "train" if train else "val"
        if not hasattr(self, "_xdev"):
            self._xdev = {}
        if key not in self._xdev:
            self._xdev[key] = x.to(eng.dev).contiguous()
        out, _ = eng.predict(flat, self._xdev[key], want_out=True)
        return out[0] if out.shape[0] == 1 else out

    def _log_likelihood(self, states, hyperStates, likelihood):
        eng = self._engine if self._engine is not None else self._make_engine(likelihood)
        st = self.states if states is None else states
        hs = self.hyperStates if hyperStates is None else hyperStates
        th = torch.cat([as_tensor(t, self.tdtype).reshape(-1).to(eng.dev) for t in st]).reshape(1, -1)
        hy = torch.cat([as_tensor(t, self.tdtype).reshape(-1).to(eng.dev) for t in hs]).reshape(1, -1)
        th = th.repeat(eng.chains, 1) if th.shape[0] != eng.chains else th
        hy = hy.repeat(eng.chains, 1) if hy.shape[0] != eng.chains else hy
        _, _, stat = eng.logp_grad(th, hy)
        n = float(self.trainY.numel())
        if likelihood.kind == "bernoulli":
            return stat[0]
        sd = float(hy[0, -1]) ** 2 if likelihood.kind == "gaussian" else likelihood.sd
        sd = min(max(sd, 1e-8), 1e8)
        return -0.5 * (2.0 * n * math.log(sd) + stat[0] / sd ** 2 + n * math.log(2.0 * math.pi))

    def calculateProbs(self, *argv, sd=None):
        """Log posterior of the given (or current) states under the current hypers
        (the closure of reference network.py:370-392)."""
        states = self.states if len(argv) == 0 else (argv[0] if len(argv) != len(self.states) else argv)
        self._ensure_device_state(self.likelihood)
        eng = self._engine
        th = torch.cat([as_tensor(t, self.tdtype).reshape(-1).to(eng.dev) for t in states]).reshape(1, -1)
        lp, _, _ = eng.logp_grad(th.repeat(eng.chains, 1) if self.chains > 1 else th, self._hyper)
        return lp[0]

    def metrics(self, trainPredict, trainReal, validatePredict, validateReal):
        for metric in self.metricList:
            metric.calculate(trainPredict, validatePredict, trainReal, validateReal)
            metric.display()

    # ------------------------------------------------------------------ sample store
    def _open_files(self, path, idx):
        """Text files in the reference's layout (network.py:545-559) followed by their binary side-cars
        (`<name>.f32`: the same values as raw little-endian float32 in the same order).  The reference's
        predictor ignores the side-cars; this package's predictor reads them instead of parsing ~25 bytes of text
        per value."""
        names = ["%d.%d" % (n, idx) for n in range(len(self.states))] + ["hypers%d" % idx]
        files = [open(os.path.join(path, nm + ".txt"), "wb") for nm in names]
        files += [open(os.path.join(path, nm + ".f32"), "wb") for nm in names]
        return files

    def _chain_dirs(self, folderName):
        base = os.path.join(os.getcwd(), folderName)
        if self.chains == 1:
            return [base]
        return [os.path.join(base, "chain%d" % c) for c in range(self.chains)]

    # ------------------------------------------------------------------ training
    def train(self, epochs, samplingStep, likelihood, metricList=[], adjustHypers=True, scaleExp=False,
              folderName=None, networksPerFile=1000, displaySkip=1, verbose=True):
        """Runs the sampler (reference network.py:509-670).  Samples are written in the reference's
        text layout under os.getcwd()/folderName; ``folderName=None`` means
        "do not save" (the reference crashes there, Q8).  With chains > 1 each chain gets
        folderName/chain<c>/ in the same layout."""
        startSampling = self.burnin
        self.likelihood = likelihood
        self.makeResponseLikelihood = likelihood.makeResponseLikelihood
        self.metricList = metricList
        self.adjustHypers = adjustHypers
        if not self._lik_appended:             # calling train twice must not duplicate them (Q12)
            for val in likelihood.hypers:
                self.hyperStates.append(as_tensor(val, self.tdtype).reshape(-1))
            self._lik_appended = True
        self._ensure_device_state(likelihood)
        eng = self._engine
        self.adapt.verbose = bool(verbose)
        if self.adapt.device is None:
            self.adapt.device = eng.device
        C = self.chains

        dirs, files = [], []
        if folderName is not None:
            dirs = self._chain_dirs(folderName)
            for d in dirs:
                os.makedirs(d, exist_ok=True)
                files.append(self._open_files(d, 0))
                with open(os.path.join(d, "architecture.txt"), "wb") as f:
                    for layer in self.layers:
                        f.write((layer.name + "\n").encode("utf-8"))

        da_state = torch.tensor([[self.h, self.logEpsilonBar, self.hyper_step_size]], dtype=self.tdtype)
        da_state = da_state.repeat(C, 1).to(eng.dev).contiguous()
        stats = torch.zeros(C, 4, dtype=self.tdtype, device=eng.dev)
        hstats = torch.zeros(C, 2, dtype=self.tdtype, device=eng.dev)
        host = torch.zeros(C, 9, dtype=self.tdtype).pin_memory()
        self.mainAccept = 0.0
        self.hyperAccept = 0.0
        iter_ = 0
        startTime = time.time()
        while iter_ < epochs:
            eng.hmc_step(self._theta, self._hyper, self.seed, iter_, self.step_size, self.leapfrog, stats=stats)
            if adjustHypers and eng.H > 0:
                eng.hyper_step(self._theta, self._hyper, self.seed, iter_, self.hyperLeapfrog, float(iter_),
                               float(self.burnin), self.hyperStepSize0, da_state, stats=hstats)
            host[:, 0:4].copy_(stats, non_blocking=True)
            host[:, 4:6].copy_(hstats, non_blocking=True)
            host[:, 6:9].copy_(da_state, non_blocking=True)
            torch.cuda.current_stream().synchronize()
            self.mainAccept = float(host[:, 1].mean())
            self.hyperAccept = float(host[:, 5].mean())
            self.h, self.logEpsilonBar = float(host[0, 6]), float(host[0, 7])
            self.hyper_step_size = float(host[0, 8])
            sjd = float(host[:, 3].mean())
            iter_ += 1
            self.iteration = iter_

            if verbose and iter_ % displaySkip == 0:
                print()
                print("iter:{:>2}".format(iter_))
                print("step size", self.step_size)
                print("hyper step size", self.hyper_step_size)
                print("leapfrog", self.leapfrog)
                print("Main acceptance", self.mainAccept)
                print("Hyper acceptance", self.hyperAccept)
                self.metrics(self.predict(train=True), self.trainY.to(eng.dev),
                             self.predict(train=False), self.validateY.to(eng.dev))
            step, leap = self.adapt.update(sjd=sjd)
            self.step_size = float(step)
            self.leapfrog = int(leap)

            # file rollover + summary (reference network.py:609-646, lagging summary of Q9 reproduced)
            indexShift = iter_ - startSampling - 1
            indexInterval = networksPerFile * samplingStep
            if dirs and iter_ > startSampling and indexShift % indexInterval == 0:
                fidx = int((iter_ - startSampling) // (networksPerFile * samplingStep))
                for c, d in enumerate(dirs):
                    for fh in files[c]:
                        fh.close()
                    files[c] = self._open_files(d, fidx)
                    with open(os.path.join(d, "summary.txt"), "wb") as fh:
                        for s in self.states:
                            sh = s.shape[1:] if C > 1 else s.shape
                            fh.write((" ".join(str(int(v)) for v in sh).strip() + "\n").encode("utf-8"))
                        numNetworks = indexShift // samplingStep
                        numFiles = numNetworks // networksPerFile
                        if numNetworks % networksPerFile != 0:
                            numFiles += 1
                        fh.write(("%d %d %d\n" % (numNetworks, numFiles, len(self.states))).encode("utf-8"))
                        fh.write(str(int(self._hyper.shape[1])).encode("utf-8"))
            # record a sample (reference network.py:647-663; one hyper scalar per line, Q5)
            if dirs and iter_ > startSampling and iter_ % samplingStep == 0:
                nfile = len(self.states) + 1      # text handles first, then as many side-car handles
                th = self._theta.detach().cpu().double().numpy()
                hy = self._hyper.detach().cpu().double().numpy()
                for c in range(C):
                    off = 0
                    for n, s in enumerate(self.states):
                        sh = tuple(s.shape[1:] if C > 1 else s.shape)
                        cnt = int(np.prod(sh))
                        np.savetxt(files[c][n], th[c, off:off + cnt].reshape(sh))
                        th[c, off:off + cnt].astype("<f4").tofile(files[c][nfile + n])
                        off += cnt
                    np.savetxt(files[c][nfile - 1], hy[c].reshape(-1, 1))
                    hy[c].astype("<f4").tofile(files[c][2 * nfile - 1])
            if verbose and iter_ % displaySkip == 0:
                likelihood.display(self.hyperStates)
                print("Time elapsed:", time.time() - startTime)
                startTime = time.time()
        for fl in files:
            for fh in fl:
                fh.close()
