"""paramAdapter (GP-UCB search over (step size, leapfrog count)) restated in
numpy float32.

TEST INFRASTRUCTURE (see oracle/__init__.py; parity unpinned).  Follows
paramAdapter.py:39-93 (state), :95-111 (calck), :113-141 (calcUCB),
:143-156 (reset), :158-196 (gridSearch), :199-292 (update).  Randomness
(``tf.random.uniform`` at :232 and python ``random.choice`` at :283-284) is
drawn from an injected ``random.Random`` so runs are reproducible.
"""
import math
import random

import numpy as np

F = np.float32


class OracleAdapter(object):
    def __init__(self, e1, L1, el, eu, eNumber, Ll, Lu, lStep, m, k, a=4,
                 delta=0.1, cores=4, strikes=10, randomSteps=10, rng=None):
        self.currentE = e1
        self.currentL = L1
        self.el, self.eu = F(el), F(eu)
        self.Ll, self.Lu = F(Ll), F(Lu)
        self.eNumber = int(eNumber)
        self.eGrid = np.linspace(el, eu, num=eNumber).astype(F)          # :68
        self.lGrid = np.array(range(Ll, Lu + 1, int(lStep)), dtype=F)     # :69
        self.lNumber = len(self.lGrid)
        self.delta = F(delta)
        kappa = F(0.2)
        self.sigma = np.diag([1 / ((kappa * 2) ** 2), 1 / ((kappa * 2) ** 2)]).astype(F)  # :72-74
        self.k = k
        self.m = m
        self.a = F(a)
        self.maxStrikes = 50                                             # :92 (ignores ``strikes``)
        self.randomSteps = randomSteps
        self.rng = rng if rng is not None else random.Random(0)
        self.reset()
        self.strikes = 0

    def reset(self):                                                     # :143-156
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

    def _norm(self, gamma, el, eu):
        return np.array([-1 + 2 * (F(gamma[0]) - el) / (eu - el),
                         -1 + 2 * (F(gamma[1]) - self.Ll) / (self.Lu - self.Ll)], dtype=F)

    def calck(self, gammaI, gammaJ, el, eu, sigma):                      # :95-111
        g1 = self._norm(gammaI, el, eu)
        g2 = self._norm(gammaJ, el, eu)
        return F(np.exp(F(-0.5) * F(g1 @ (sigma @ g2))))

    def calcUCB(self, testGamma, previousGamma, inverseR, s, inverse, p, rootbeta, el, eu, sigma):
        k = np.array([self.calck(g, testGamma, el, eu, sigma) for g in previousGamma], dtype=F)
        mean = F(k @ inverseR[:, 0]) * s
        variance = F(k @ (inverse @ k))
        variance = self.calck(testGamma, testGamma, el, eu, sigma) - variance
        ucb = mean + variance * F(p) * F(rootbeta)
        return ucb, mean, variance

    def ucb_surface(self, previousGamma, inverseR, s, inverse, p, rootbeta, el, eu, sigma):
        """Vectorised UCB over the whole grid, [lNumber, eNumber] (float64
        arithmetic; used to judge near-ties)."""
        el64, eu64 = float(el), float(eu)
        ge = -1 + 2 * (self.eGrid.astype(np.float64) - el64) / (eu64 - el64)
        gl = -1 + 2 * (self.lGrid.astype(np.float64) - float(self.Ll)) / (float(self.Lu) - float(self.Ll))
        sg = sigma.astype(np.float64)
        pe = np.array([-1 + 2 * (float(F(g[0])) - el64) / (eu64 - el64) for g in previousGamma])
        pl = np.array([-1 + 2 * (float(F(g[1])) - float(self.Ll)) / (float(self.Lu) - float(self.Ll))
                       for g in previousGamma])
        # k_j(e,L) = exp(-0.5 (pe_j*s00*ge + pl_j*s11*gl))   (sigma diagonal)
        kv = np.exp(-0.5 * (pe[:, None, None] * sg[0, 0] * ge[None, None, :]
                            + pl[:, None, None] * sg[1, 1] * gl[None, :, None]))
        mean = np.einsum("jle,j->le", kv, inverseR[:, 0].astype(np.float64)) * float(s)
        quad = np.einsum("ile,ij,jle->le", kv, inverse.astype(np.float64), kv)
        kself = np.exp(-0.5 * (sg[0, 0] * ge[None, :] ** 2 + sg[1, 1] * gl[:, None] ** 2))
        return mean + (kself - quad) * float(p) * float(rootbeta)

    def gridSearch(self, previousGamma, inverseR, s, inverse, p, rootbeta, el, eu, sigma):
        """:158-196 -- e fastest, L slowest, strict '>' => first maximum wins."""
        best, e, L = F(-1000000000), F(el), F(self.Ll)
        for lc in range(self.lNumber):
            for ec in range(self.eNumber):
                newE, newL = self.eGrid[ec], self.lGrid[lc]
                ucb, _, _ = self.calcUCB([newE, newL], previousGamma, inverseR, s, inverse,
                                         p, rootbeta, el, eu, sigma)
                if ucb > best:
                    best, e, L = ucb, newE, newL
        return F(e), F(L)

    def sjd(self, previous_state, current_state):                        # :219-222
        val = F(0)
        for old, new in zip(previous_state, current_state):
            d = np.asarray(new, dtype=F).reshape(-1) - np.asarray(old, dtype=F).reshape(-1)
            val += F(np.sum(np.square(d))) / F(F(self.currentL) ** F(0.5))
        return val

    def update(self, state):                                             # :199-292
        if self.i < self.k - 2 and self.strikes == self.maxStrikes:
            self.el = self.el / 2
            self.eu = self.eu / 2
            self.eGrid = np.linspace(self.el, self.eu, num=self.eNumber).astype(F)
            self.k = self.k - self.i - 2
            self.reset()
            self.strikes = 0

        self.previous_state, self.current_state = self.current_state, state

        if self.previous_state is not None:
            val = self.sjd(self.previous_state, self.current_state)
            self.currentData.append(val)
            if val < 1e-8 and self.i // self.m > self.randomSteps:
                self.strikes += 1
            else:
                self.strikes = 0

        if self.i % self.m == 0 and self.i > 0:
            u = self.rng.random()
            self.p = max(self.i / self.m - self.k + 1, 1) ** (-0.5)
            if u < self.p:
                mean = F(np.mean(np.array(self.currentData, dtype=F)))
                sd = F(np.std(np.array(self.currentData, dtype=F)))
                self.currentData = []
                self.allData.append(mean)
                self.allSD.append(sd)
                self.maxR = F(np.max(self.allData))
                self.previousGamma.append((self.currentE, self.currentL))
                size = len(self.previousGamma)
                extra = np.array([self.calck(g, self.previousGamma[-1], self.el, self.eu, self.sigma)
                                  for g in self.previousGamma], dtype=F)
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
                except np.linalg.LinAlgError:
                    self.inverse = np.linalg.inv(A + F(0.1) * np.eye(size, dtype=F)).astype(F)
                self.inverseR = self.inverse @ np.array(self.allData, dtype=F)[:, None]
                rb = (self.i / self.m + 1) ** 3 * math.pi ** 2
                rb /= (3 * float(self.delta))
                self.rootbeta = (math.log(rb) * 2) ** 0.5
                if self.i // self.m >= self.randomSteps:
                    self.currentE, self.currentL = self.gridSearch(
                        self.previousGamma, self.inverseR, self.s, self.inverse, self.p,
                        self.rootbeta, self.el, self.eu, self.sigma)
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
