"""TFP HamiltonianMonteCarlo semantics as TensorBNN uses them, restated.

TEST INFRASTRUCTURE (see oracle/__init__.py; parity unpinned).  TFP is a
third-party dependency absent from /root/reference (README.md:26,32 says
TFP 0.12.2, docs/Setup.md:21 says 0.11; nothing is pinned); the algorithm
restated here is TFP's published SimpleLeapfrogIntegrator + MetropolisHastings
(SURVEY.md Appendix B), anchored on the reference call sites
network.py:394-411 (main chain), :442-471 (hyper chain + dual averaging).
All functions work on flat torch vectors in the dtype of ``theta``.
"""
import math

import numpy as np
import torch

from . import targets
from .tfconst import f32


def leapfrog(value_and_grad, theta, momentum, eps, L, logp0=None, grad0=None):
    """TFP order (SimpleLeapfrogIntegrator):
        p = p + (eps/2) g(theta)
        L times: theta += eps p ; (logp,g) = vg(theta) ; p += eps g
        p = p - (eps/2) g
    Returns theta', p', logp', g'."""
    if grad0 is None:
        logp0, grad0 = value_and_grad(theta)
    dt = theta.dtype
    eps = torch.as_tensor(eps, dtype=dt)
    half = eps * 0.5
    p = momentum + half * grad0
    th, g, lp = theta.clone(), grad0, logp0
    for _ in range(int(L)):
        th = th + eps * p
        lp, g = value_and_grad(th)
        p = p + eps * g
    p = p - half * g
    return th, p, lp, g


def log_accept_ratio(logp0, logp1, p0, p1):
    """MetropolisHastings: safe_sum([logp', -logp, 0.5 sum p^2, -0.5 sum p'^2]);
    an indeterminate / NaN sum is -inf (reject)."""
    terms = torch.stack([logp1, -logp0, 0.5 * torch.sum(p0 * p0), -0.5 * torch.sum(p1 * p1)])
    s = torch.sum(terms)
    has_pinf = bool(torch.any(terms == math.inf))
    has_ninf = bool(torch.any(terms == -math.inf))
    if bool(torch.any(torch.isnan(terms))) or (has_pinf and has_ninf):
        return torch.tensor(-math.inf, dtype=terms.dtype)
    return s


def hmc_step(value_and_grad, theta, momentum, u, eps, L):
    """bootstrap_results + one_step of sample_chain(num_results=1)
    (network.py:400-408) with injected momentum and uniform ``u``.
    Returns (new theta, log_accept_ratio, accept probability min(1,e^lar)
    as reported at network.py:410-411, accepted flag, proposal, proposal momentum)."""
    logp0, g0 = value_and_grad(theta)
    th1, p1, logp1, _ = leapfrog(value_and_grad, theta, momentum, eps, L, logp0, g0)
    lar = log_accept_ratio(logp0, logp1, momentum, p1)
    accepted = bool(math.log(u) < lar.item()) if u > 0 else bool(-math.inf < lar.item())
    new = th1 if accepted else theta
    prob = torch.where(lar < 0, torch.exp(lar), torch.ones_like(lar))
    return new, lar, prob, accepted, th1, p1


def dual_averaging(epoch, accept, h, log_eps_bar, step, hyper_step0, burnin,
                   target=0.95, gamma=0.4, t0=10.0, kappa=0.75):
    """network.py:457-469 with the constants of :241-248.  ``epoch`` is the
    0-based iteration; mu = log(100*hyperStepSize)."""
    m = epoch + 1.0
    # tf.cast(0.4, dtype) and tf.cast(tf.math.log(100*hyperStepSize), dtype) pass through float32 (Q14, oracle/tfconst.py)
    gamma = f32(gamma)
    mu = float(np.log(np.float32(100.0 * hyper_step0)))
    h = (1 - 1 / (m + t0)) * h + (1 / (m + t0)) * (target - accept)
    log_eps = mu - h * (m ** 0.5) / gamma
    log_eps_bar = (1 - m ** (-kappa)) * log_eps_bar + m ** (-kappa) * log_eps
    if m < burnin * 0.8:
        step = math.exp(log_eps_bar)
    return h, log_eps_bar, step


def make_main_vg(arch, lik, hyper_flat, X, Y):
    def vg(theta):
        return targets.main_value_and_grad(arch, lik, theta, hyper_flat, X, Y)
    return vg


def make_hyper_vg(arch, lik, theta_flat, X, Y):
    def vg(hyper):
        return targets.hyper_value_and_grad(arch, lik, theta_flat, hyper, X, Y)
    return vg
