"""Step-size / leapfrog-count adaptation with the reference's interface
(tensorBNN/paramAdapter.py): GP-UCB optimisation of the expected squared jump distance over
an (e, L) grid after Wang, Mohamed and de Freitas.  The bookkeeping (history, <=50x50 GP
solve) is host numpy in float32 like the reference (:60); the exhaustive grid search
(:158-196) runs on the device through tbnn_adapter_ucb.

Besides the reference's ``update(state)`` (which needs the previous and the new state),
``update(sjd=...)`` accepts the squared jump distance already reduced on the device by
tbnn_hmc_step, so the sampler never copies the weights to the host for adaptation.
"""
import math
import random

import numpy as np
import torch

F = np.float32


class paramAdapter(object):
    def __init__(self, e1, L1, el, eu, eNumber, Ll, Lu, lStep, m, k, a=4, delta=0.1, cores=4,
                 strikes=10, randomSteps=10, device=None, rng=None):
        self.dtype = np.float32
        self.currentE = e1
        self.currentL = L1
        self.el, self.eu = F(el), F(eu)
        self.Ll, self.Lu = F(Ll), F(Lu)
        self.eNumber = int(eNumber)
        self.eGrid = np.linspace(el, eu, num=int(eNumber)).astype(F)
        self.lGrid = np.array(range(int(Ll), int(Lu) + 1, int(lStep)), dtype=F)
        self.lNumber = len(self.lGrid)
        self.delta = F(delta)
        kappa = F(0.2)
        self.sigma = np.diag([1 / ((kappa * 2) ** 2), 1 / ((kappa * 2) ** 2)]).astype(F)
        self.k = k
        self.m = m
        self.a = F(a)
        self.cores = cores
        self.maxStrikes = 50          # the reference hard-codes 50 and ignores ``strikes`` (:92)
        self.randomSteps = randomSteps
        self.device = device
        self.rng = rng if rng is not None else random
        self.verbose = True
        self._clear()

    def _clear(self):
        self.previousGamma = []
        self.allSD = []
        self.K = np.zeros((0, 0), dtype=F)
        self.currentData = []
        self.allData = []
        self.maxR = F(1e-8)
        self.i = -2
        self.previous_state = None
        self.current_state = None
        self.strikes = 0

    def reset(self):
        """Resets the adapter (reference :143-156)."""
        if self.verbose:
            print("Reset")
        self._clear()

    # -- covariance of two (e, L) points: exp(-0.5 g1^T Sigma g2) on [-1,1]-normalised points (Q11)
    def _norm(self, gamma, el, eu):
        return np.array([-1 + 2 * (F(gamma[0]) - el) / (eu - el),
                         -1 + 2 * (F(gamma[1]) - self.Ll) / (self.Lu - self.Ll)], dtype=F)

    def calck(self, gammaI, gammaJ, el=None, eu=None, sigma=None):
        el = self.el if el is None else el
        eu = self.eu if eu is None else eu
        sigma = self.sigma if sigma is None else sigma
        return F(np.exp(F(-0.5) * F(self._norm(gammaI, el, eu) @ (sigma @ self._norm(gammaJ, el, eu)))))

    def gridSearch(self, previousGamma, inverseR, s, inverse, p, rootbeta, el, eu, sigma):
        """Exhaustive UCB arg-max on the device (first maximum in scan order wins)."""
        from .engine import adapter_ucb
        dev = self.device if self.device is not None else torch.cuda.current_device()
        e, L, _ = adapter_ucb(dev, self.eGrid, self.lGrid, np.array(previousGamma, dtype=F), inverse,
                              np.asarray(inverseR).reshape(-1), s, p, rootbeta, el, eu, self.Ll, self.Lu,
                              sigma)
        return F(e), F(L)

    @staticmethod
    def _sjd_of_states(previous_state, current_state, L):
        val = 0.0
        for old, new in zip(previous_state, current_state):
            d = torch.as_tensor(new).reshape(-1).to(torch.float32) - torch.as_tensor(old).reshape(-1).to(torch.float32)
            val += float(torch.sum(d * d)) / float(F(F(L) ** F(0.5)))
        return F(val)

    def calls_until_decision(self):
        """How many update() calls can be made before the one that may change (step size, L): that is the call made
        while ``i % m == 0 and i > 0`` (reference :231).  0 = the very next call may decide."""
        i, m = int(self.i), int(self.m)
        k = 0
        while not ((i + k) % m == 0 and (i + k) > 0):
            k += 1
        return k

    def update(self, state=None, sjd=None):
        """One adapter step (reference :199-292).  Returns (float32 step size, int32 leapfrog)."""
        if self.i < self.k - 2 and self.strikes == self.maxStrikes:
            self.el = self.el / 2
            self.eu = self.eu / 2
            self.eGrid = np.linspace(self.el, self.eu, num=self.eNumber).astype(F)
            self.k = self.k - self.i - 2
            self.reset()
            self.strikes = 0

        val = None
        if sjd is not None:
            # device-reduced |theta_new - theta_old|^2, scaled by L^-0.5 here
            had_previous = self.current_state is not None
            self.previous_state, self.current_state = self.current_state, True
            if had_previous:
                val = F(F(sjd) / F(F(self.currentL) ** F(0.5)))
        else:
            state = [torch.as_tensor(s).detach().clone() for s in state]
            self.previous_state, self.current_state = self.current_state, state
            if self.previous_state is not None:
                val = self._sjd_of_states(self.previous_state, self.current_state, self.currentL)
        if val is not None:
            if self.verbose:
                print("SJD:", float(val))
            self.currentData.append(val)
            if val < 1e-8 and self.i // self.m > self.randomSteps:
                self.strikes += 1
            else:
                self.strikes = 0

        if self.i % self.m == 0 and self.i > 0:
            u = self.rng.random()
            self.p = max(self.i / self.m - self.k + 1, 1) ** (-0.5)
            if u < self.p:
                data = np.array(self.currentData, dtype=F)
                mean, sd = F(np.mean(data)), F(np.std(data))
                self.currentData = []
                self.allData.append(mean)
                self.allSD.append(sd)
                self.maxR = F(np.max(self.allData))
                self.previousGamma.append((self.currentE, self.currentL))
                size = len(self.previousGamma)
                extra = np.array([self.calck(g, self.previousGamma[-1]) for g in self.previousGamma], dtype=F)
                newK = np.zeros((size, size), dtype=F)
                newK[:size - 1, :size - 1] = self.K
                newK[size - 1, :] = extra
                newK[:, size - 1] = extra
                self.K = newK
                self.s = self.a / self.maxR
                sigmaNu = F(np.mean(np.array(self.allSD, dtype=F)))
                A = self.K + (sigmaNu ** 2) * np.eye(size, dtype=F)
                try:
                    self.inverse = np.linalg.inv(A).astype(F)
                    if not np.all(np.isfinite(self.inverse)):
                        raise np.linalg.LinAlgError("singular")
                except np.linalg.LinAlgError:
                    self.inverse = np.linalg.inv(A + F(0.1) * np.eye(size, dtype=F)).astype(F)
                self.inverseR = self.inverse @ np.array(self.allData, dtype=F)[:, None]
                rb = (self.i / self.m + 1) ** 3 * math.pi ** 2 / (3 * float(self.delta))
                self.rootbeta = (math.log(rb) * 2) ** 0.5
                if self.i // self.m >= self.randomSteps:
                    self.currentE, self.currentL = self.gridSearch(
                        self.previousGamma, self.inverseR, self.s, self.inverse, self.p, self.rootbeta,
                        self.el, self.eu, self.sigma)
                else:
                    self.currentE = self.rng.choice(list(self.eGrid))
                    self.currentL = self.rng.choice(list(self.lGrid))
                if size == 50:
                    self.K = self.K[1:, 1:]
                    self.previousGamma = self.previousGamma[1:]
                    self.allData = self.allData[1:]
                    self.allSD = self.allSD[1:]

        self.i += 1
        return F(self.currentE), np.int32(self.currentL)
