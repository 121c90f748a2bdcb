"""Output likelihoods with the reference's interface (tensorBNN/likelihood.py): attributes
``hypers`` and ``mainProbsInHypers``; methods ``makeResponseLikelihood``,
``calcultateLogProb`` (sic) and ``display``.  During sampling the likelihood is evaluated
inside the fused CUDA kernel; ``makeResponseLikelihood`` is kept as a user-callable wrapper
around the same C ABI (it returns the summed log-likelihood)."""
import math

import torch


class Likelihood(object):
    kind = None

    def __init__(self, *argv, **kwargs):
        self.hypers = []
        self.mainProbsInHypers = False

    def spec(self):
        raise NotImplementedError("custom likelihoods written against TensorFlow are not supported; "
                                  "use GaussianLikelihood / FixedGaussianLikelihood / BernoulliLikelihood")

    def makeResponseLikelihood(self, *argv, **kwargs):
        """Summed log-likelihood of the training data for the given states, evaluated by the
        CUDA engine of the network bound through ``kwargs['predict']`` (network.predict)."""
        net = getattr(kwargs.get("predict"), "__self__", None)
        if net is None or not hasattr(net, "_log_likelihood"):
            raise RuntimeError("makeResponseLikelihood needs predict=network.predict of a tensorbnn_b200 network")
        return net._log_likelihood(argv[0] if len(argv) else None, kwargs.get("hyperStates"), self)

    def calcultateLogProb(self, *argv, **kwargs):
        """(sic)  The per-sample likelihood terms of predictor.trainProbs / reweight are evaluated on the device by
        predictor._neg_log_weights; this host method exists for interface compatibility only."""
        raise NotImplementedError("use predictor.trainProbs / predictor.reweight: the per-sample likelihood is "
                                  "evaluated on the device there")

    def display(self, hypers):
        pass


class GaussianLikelihood(Likelihood):
    """Gaussian output with a sampled sd: hyper = sqrt(sd), sigma = hyper**2
    (reference likelihood.py:63-133)."""
    kind = "gaussian"

    def __init__(self, *argv, **kwargs):
        self.sd0 = float(kwargs["sd"])
        self.hypers = [[self.sd0 ** 0.5]]
        self.mainProbsInHypers = True

    def spec(self):
        return ("gaussian", self.sd0)

    def display(self, hypers):
        h = hypers[-1]
        v = float(h.reshape(-1)[0]) if isinstance(h, torch.Tensor) else float(h)
        print("Loss Standard Deviation: ", v ** 2)


class FixedGaussianLikelihood(Likelihood):
    """Gaussian output with a fixed sd (reference likelihood.py:136-202)."""
    kind = "fixed"

    def __init__(self, *argv, **kwargs):
        self.hypers = []
        self.sd = float(kwargs["sd"])
        self.mainProbsInHypers = False

    def spec(self):
        return ("fixed", self.sd)


class BernoulliLikelihood(Likelihood):
    """Bernoulli output on probabilities clipped to [1e-8, 1-1e-7] (reference likelihood.py:205-243)."""
    kind = "bernoulli"

    def __init__(self, *argv, **kwargs):
        self.hypers = []
        self.mainProbsInHypers = False

    def spec(self):
        return ("bernoulli",)

    def calcultateLogProb(self, *argv, **kwargs):
        return [0.0 for _ in kwargs["hypers"]]
