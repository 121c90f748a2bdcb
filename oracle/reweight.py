"""predictor.trainProbs / reweight restated (reference predictor.py:157-273).  TEST INFRASTRUCTURE (see oracle/__init__.py).

For every stored sample m the reference accumulates  -( likelihood term + sum over layers calculateHyperProbs(hypers_m,
tensors_m) )  under the training architecture and under the new one, and returns exp(train - new) normalised to one.
As shipped, three pieces of that path raise (dense calculateHyperProbs indexes a numpy scalar, predictor.py:195-198 ->
layer.py:207-215; GaussianLikelihood.calcultateLogProb needs a keyword trainProbs never passes, likelihood.py:116-119),
so -- as for Q4 -- the behaviour is DEFINED as the minimal patch: the layer terms are the training-time hyper
conditionals (targets.layer_hyper_prob); FixedGaussianLikelihood uses multivariateLogProb with the constructor sd and
BernoulliLikelihood contributes 0, both as written (likelihood.py:190-194, :239-243); GaussianLikelihood uses
sigma = stored hyper**2 as in training.
"""
import numpy as np
import torch

from . import targets


def neg_log_weights(arch, lik, samples, hypers, X=None, Y=None):
    """samples [S, P], hypers [S, H_stored] (float64).  lik = None: no likelihood term."""
    t = lambda a: torch.tensor(np.asarray(a, dtype=np.float64))
    out = np.zeros(len(samples))
    n_layer_h = sum(targets.num_tensors(l)[1] for l in arch)
    for m in range(len(samples)):
        theta = targets.unflatten_theta(arch, t(samples[m]))
        hy = [t(hypers[m][j]) for j in range(n_layer_h)]
        val = 0.0
        ih = it = 0
        for layer in arch:
            nt, nh = targets.num_tensors(layer)
            if nh > 0:
                val = val + targets.layer_hyper_prob(layer, hy[ih:ih + nh], theta[it:it + nt])
            ih += nh
            it += nt
        if lik is not None and lik[0] == "gaussian":
            val = val + targets.log_likelihood(arch, lik, theta, t(X), t(Y), sd_hyper=t(hypers[m][-1]))
        elif lik is not None and lik[0] == "fixed":
            val = val + targets.log_likelihood(arch, lik, theta, t(X), t(Y))
        out[m] = -float(val)
    return out


def reweight(arch_train, arch_new, lik_train, lik_new, samples, hypers, X=None, Y=None):
    wt = neg_log_weights(arch_train, lik_train, samples, hypers, X, Y)
    wn = neg_log_weights(arch_new, lik_new, samples, hypers, X, Y)
    d = wt - wn
    w = np.exp(d - d.max())
    return w / w.sum()
