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
        self._engine_key = None
        self._rng_calls = 0     # Philox call counter, persistent across train() calls (a second call draws fresh numbers)
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
        # tf.cast(stepSizeStart, dtype): a python float passes through float32 (network.py:237; quirk Q14)
        self.step_size = float(np.float32(stepSizeStart))
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
        shapes = self._state_shapes
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
        """Builds (or rebuilds) the engine and the flat device state.  The cache is keyed on the architecture AND the
        likelihood: calculateProbs() before train() builds an engine without the likelihood's hyper parameter, and a
        later train(GaussianLikelihood(...)) must not keep sampling with that stale engine."""
        key = (repr(self.arch_spec()), repr(self._lik_spec(likelihood)))
        if self._engine is not None and self._engine_key == key:
            return
        if self._engine is not None:
            # keep the current sample: the views die with the old flat state
            self.states = [s.detach().clone() for s in self.states]
            self.hyperStates = [h.detach().clone().reshape(-1) if self.chains == 1 else h.detach().clone()
                                for h in self.hyperStates]
        self._engine = self._make_engine(likelihood)
        self._engine_key = key
        eng = self._engine
        C = self.chains
        bound = self._theta is not None            # states already are [C, ...] views / copies of a flat state
        self._state_shapes = [tuple(s.shape[1:]) if (bound and C > 1) else tuple(s.shape) for s in self.states]

        def flat(tensors, width):
            parts = []
            for t in tensors:
                t = t.to(self.tdtype).cpu()
                per_chain = bound and C > 1 and t.dim() >= 2 and t.shape[0] == C
                parts.append(t.reshape(C, -1) if per_chain else t.reshape(1, -1))
            rows = max([q.shape[0] for q in parts] + [1])
            parts = [q if q.shape[0] == rows else q.repeat(rows, 1) for q in parts]
            out = torch.cat(parts, dim=1) if parts else torch.zeros(1, 0, dtype=self.tdtype)
            if out.shape[1] != width:
                raise ValueError("the network state holds %d values but the engine expects %d" % (out.shape[1], width))
            return out

        th = flat(self.states, eng.P)
        hy = flat(self.hyperStates, eng.H)
        if th.shape[0] == 1 and C > 1:
            th = th.repeat(C, 1)
            # independent chains start from jittered copies of the given state
            g = torch.Generator().manual_seed(self.seed)
            th[1:] += (0.01 * torch.randn(C - 1, eng.P, generator=g, dtype=torch.float64)).to(th)
        if hy.shape[0] == 1 and C > 1:
            hy = hy.repeat(C, 1)
        self._theta = th.to(eng.dev).contiguous()
        self._hyper = hy.to(eng.dev).contiguous()
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
        (the closure of reference network.py:370-392).  Needs the likelihood: call train() first or set
        ``network.likelihood``."""
        if self.likelihood is None:
            raise RuntimeError("calculateProbs needs a likelihood: set network.likelihood or call train() first")
        self._ensure_device_state(self.likelihood)
        eng = self._engine
        if len(argv) == 0:
            th = self._theta
        else:
            states = argv[0] if len(argv) != len(self.states) else argv
            th = torch.cat([as_tensor(t, self.tdtype).reshape(-1).to(eng.dev) for t in states]).reshape(1, -1)
            th = th.repeat(eng.chains, 1) if self.chains > 1 else th
        lp, _, _ = eng.logp_grad(th, self._hyper)
        return lp[0] if self.chains == 1 else lp

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
                # tf.cast(val, dtype) of python floats passes through float32 (network.py:542-543; quirk Q14)
                self.hyperStates.append(torch.tensor(val, dtype=torch.float32).to(self.tdtype).reshape(-1))
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
                # constants of stateless layers that architecture.txt cannot carry (Leaky_relu's slope is not a
                # sampled state here, Q6): side-car read back by this package's predictor
                consts = [(i, layer.alpha) for i, layer in enumerate(self.layers) if layer.name == "leakyrelu"]
                if consts:
                    with open(os.path.join(d, "layer_params.txt"), "w") as f:
                        for i, alpha in consts:
                            f.write("%d alpha %r\n" % (i, float(alpha)))

        da_state = torch.tensor([[self.h, self.logEpsilonBar, self.hyper_step_size]], dtype=self.tdtype)
        da_state = da_state.repeat(C, 1).to(eng.dev).contiguous()
        stats = torch.zeros(C, 4, dtype=self.tdtype, device=eng.dev)
        hstats = torch.zeros(C, 2, dtype=self.tdtype, device=eng.dev)
        # The host reads nine scalars per chain and epoch (accept probabilities, squared jump distance for the adapter,
        # dual-averaging state) -- but only when it needs them: the adapter can change (step size, L) only on its
        # decision epochs (every `averagingSteps`-th call), so the epochs in between are queued on the stream without
        # a synchronisation and their rows are consumed together (pinned ring, asynchronous copies).
        ring = torch.zeros(max(2, int(self.adapt.m) + 1), C, 9, dtype=self.tdtype).pin_memory()
        pending = []
        self.mainAccept = 0.0
        self.hyperAccept = 0.0

        def consume():
            if not pending:
                return
            torch.cuda.current_stream(eng.dev).synchronize()
            for slot in pending:
                row = ring[slot]
                self.mainAccept = float(row[:, 1].mean())
                self.hyperAccept = float(row[:, 5].mean())
                self.h, self.logEpsilonBar = float(row[0, 6]), float(row[0, 7])
                self.hyper_step_size = float(row[0, 8])
                step, leap = self.adapt.update(sjd=float(row[:, 3].mean()))
                self.step_size = float(step)
                self.leapfrog = int(leap)
            del pending[:]

        iter_ = 0
        startTime = time.time()
        while iter_ < epochs:
            eng.hmc_step(self._theta, self._hyper, self.seed, self._rng_calls, self.step_size, self.leapfrog, stats=stats)
            if adjustHypers and eng.H > 0:
                eng.hyper_step(self._theta, self._hyper, self.seed, self._rng_calls, self.hyperLeapfrog, float(iter_),
                               float(self.burnin), self.hyperStepSize0, da_state, stats=hstats)
            self._rng_calls += 1
            slot = len(pending)
            ring[slot, :, 0:4].copy_(stats, non_blocking=True)
            ring[slot, :, 4:6].copy_(hstats, non_blocking=True)
            ring[slot, :, 6:9].copy_(da_state, non_blocking=True)
            pending.append(slot)
            iter_ += 1
            self.iteration = iter_
            display = verbose and iter_ % displaySkip == 0
            saving = bool(folderName) and iter_ > startSampling
            if display or saving or len(pending) >= ring.shape[0] or self.adapt.calls_until_decision() < len(pending):
                consume()

            if display:
                print()
                print("iter:{:>2}".format(iter_))
                print("step size", self.step_size)
                print("hyper step size", self.hyper_step_size)
                print("leapfrog", self.leapfrog)
                print("Main acceptance", self.mainAccept)
                print("Hyper acceptance", self.hyperAccept)
                ptrain, pval = self.predict(train=True), self.predict(train=False)
                if C > 1:                              # metrics are displayed for the first chain
                    ptrain, pval = ptrain[0], pval[0]
                self.metrics(ptrain, self.trainY.to(eng.dev), pval, self.validateY.to(eng.dev))

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
        consume()
        for fl in files:
            for fh in fl:
                fh.close()
