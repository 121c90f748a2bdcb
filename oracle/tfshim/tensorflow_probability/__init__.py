"""Minimal ``tensorflow_probability`` stand-in over torch, TEST INFRASTRUCTURE ONLY.

Companion of oracle/tfshim/tensorflow (read its header first).  Holds exactly
what the reference touches:

* ``tfp.distributions.MultivariateNormalDiag(loc=, scale_diag=).log_prob`` and
  ``tfp.distributions.Bernoulli(probs=).log_prob`` (call sites layer.py:137-153,
  221-229,318-334,401-409; activationFunctions.py:309-313,375-380;
  likelihood.py:234-236), with TFP's published formulas:
  MVNDiag: -0.5*sum(((x-loc)/scale)^2) - sum(log scale) - 0.5*k*log(2*pi) over
  the event axis; Bernoulli(probs=p): multiply_no_nan(log1p(-p), 1-y) +
  multiply_no_nan(log p, y).
* ``tfp.mcmc.HamiltonianMonteCarlo`` + ``tfp.mcmc.sample_chain`` (call sites
  network.py:315-329,394-408,442-456) restated from TFP's published algorithm
  (TFP is not vendored by the reference and not installable here; the README
  names 0.12.2): bootstrap_results = one value_and_gradient; one_step = momentum
  ~ N(0,I) per state part, SimpleLeapfrogIntegrator (half step, L x {position,
  value_and_gradient, momentum}, half step back), kinetic energies through
  log-sum-exp of log-squares, nested ``safe_sum`` (an indeterminate sum is
  -inf), log u < log_accept_ratio.  Gradients of the REFERENCE'S OWN target
  closures come from torch autograd.
* ``tfp.mcmc.DualAveragingStepSizeAdaptation``: constructed by
  network.setupMCMC (network.py:275-278) and never stepped -- an inert holder.

``mcmc.TRACE`` (a list, or None) records every value_and_gradient evaluation so
fixture generation can pin per-evaluation log-posterior and gradient.
"""
import collections
import math
import sys
import types

import torch

import tensorflow as tf

_LOG_2PI = math.log(2.0 * math.pi)


def _mod(name):
    m = types.ModuleType(__name__ + "." + name)
    sys.modules[m.__name__] = m
    return m


# ---------------------------------------------------------------------------
# distributions
# ---------------------------------------------------------------------------
distributions = _mod("distributions")


class MultivariateNormalDiag(object):
    def __init__(self, loc=None, scale_diag=None, **kwargs):
        self.loc = tf.convert_to_tensor(loc)
        self.scale = tf.convert_to_tensor(scale_diag)

    def log_prob(self, value):
        x = tf.convert_to_tensor(value)
        loc = self.loc.to(x.dtype) if x.dtype.is_floating_point else self.loc
        scale = self.scale.to(loc.dtype)
        x = x.to(loc.dtype)
        z = (x - loc) / scale
        k = float(loc.shape[-1])
        return tf._wrap(-0.5 * torch.sum(z * z, dim=-1) - torch.sum(torch.log(scale), dim=-1) - 0.5 * k * _LOG_2PI)


class Bernoulli(object):
    def __init__(self, logits=None, probs=None, **kwargs):
        assert (logits is None) != (probs is None)
        self.logits = None if logits is None else tf.convert_to_tensor(logits)
        self.probs = None if probs is None else tf.convert_to_tensor(probs)

    def log_prob(self, value):
        if self.logits is None:
            p = self.probs
            lp0, lp1 = torch.log1p(-p), torch.log(p)
        else:
            s = self.logits
            lp0, lp1 = -torch.nn.functional.softplus(s), -torch.nn.functional.softplus(-s)
        event = tf.convert_to_tensor(value).to(lp0.dtype)
        return tf._wrap(tf.math.multiply_no_nan(lp0, 1 - event) + tf.math.multiply_no_nan(lp1, event))


distributions.MultivariateNormalDiag = MultivariateNormalDiag
distributions.Bernoulli = Bernoulli

# ---------------------------------------------------------------------------
# mcmc
# ---------------------------------------------------------------------------
mcmc = _mod("mcmc")
mcmc.TRACE = None
mcmc.RESULTS = None      # a list collects every one_step's MetropolisHastingsKernelResults

UncalibratedResults = collections.namedtuple(
    "UncalibratedHamiltonianMonteCarloKernelResults",
    ["log_acceptance_correction", "target_log_prob", "grads_target_log_prob", "initial_momentum",
     "final_momentum", "step_size", "num_leapfrog_steps"])
MHResults = collections.namedtuple(
    "MetropolisHastingsKernelResults",
    ["accepted_results", "is_accepted", "log_accept_ratio", "proposed_state", "proposed_results"])
StatesAndTrace = collections.namedtuple("StatesAndTrace", ["all_states", "trace"])


def _is_list_like(x):
    return isinstance(x, (list, tuple))


def _value_and_gradient(fn, parts):
    """mcmc_util.maybe_call_fn_and_grads: fn(*parts) and its gradient w.r.t.
    every part; a part the target does not depend on raises, as in TFP."""
    leaves = [p.detach().as_subclass(torch.Tensor).clone().requires_grad_(True) for p in parts]
    value = fn(*[tf._wrap(v) for v in leaves])
    grads = torch.autograd.grad(value, leaves, allow_unused=True)
    if any(g is None for g in grads):
        raise ValueError("Encountered `None` gradient.")
    value = value.detach()
    grads = [g.detach() for g in grads]
    if mcmc.TRACE is not None:
        mcmc.TRACE.append({"state": [v.detach().clone() for v in leaves], "value": value.clone(),
                           "grads": [g.clone() for g in grads]})
    return tf._wrap(value), [tf._wrap(g) for g in grads]


def _safe_sum(terms):
    x = torch.stack([torch.as_tensor(t) for t in terms], dim=-1)
    x_sum = torch.sum(x, dim=-1)
    fin = torch.isfinite(x)
    determinate = torch.all(fin | (x >= 0), dim=-1) & torch.all(fin | (x <= 0), dim=-1)
    return torch.where(determinate, x_sum, torch.full_like(x_sum, -math.inf))


def _log_sum_sq(x):
    return torch.logsumexp(2.0 * torch.log(torch.abs(x)).reshape(-1), dim=0)


class HamiltonianMonteCarlo(object):
    def __init__(self, target_log_prob_fn, step_size, num_leapfrog_steps, state_gradients_are_stopped=False,
                 name=None, **kwargs):
        self.target_log_prob_fn = target_log_prob_fn
        self.step_size = step_size
        self.num_leapfrog_steps = num_leapfrog_steps

    def _step_sizes(self, parts):
        ss = self.step_size
        ss = list(ss) if _is_list_like(ss) else [ss]
        if len(ss) == 1:
            ss = ss * len(parts)
        return [torch.as_tensor(s).detach().as_subclass(torch.Tensor).to(p.dtype) for s, p in zip(ss, parts)]

    def bootstrap_results(self, init_state):
        parts = list(init_state) if _is_list_like(init_state) else [init_state]
        value, grads = _value_and_gradient(self.target_log_prob_fn, parts)
        inner = UncalibratedResults(torch.zeros_like(value), value, grads, None, None, self.step_size,
                                    self.num_leapfrog_steps)
        return MHResults(inner, torch.ones_like(value, dtype=torch.bool), torch.zeros_like(value),
                         parts, inner)

    def one_step(self, current_state, previous_kernel_results):
        listlike = _is_list_like(current_state)
        parts = [p.detach().as_subclass(torch.Tensor) for p in (list(current_state) if listlike else [current_state])]
        acc = previous_kernel_results.accepted_results
        eps = self._step_sizes(parts)
        L = int(torch.as_tensor(self.num_leapfrog_steps).item())
        # --- UncalibratedHamiltonianMonteCarlo.one_step
        mom0 = [tf.random.normal(tuple(p.shape), dtype=p.dtype).as_subclass(torch.Tensor) for p in parts]
        target, grads = acc.target_log_prob, [g.as_subclass(torch.Tensor) for g in acc.grads_target_log_prob]
        half = [v + (0.5 * e) * g for v, e, g in zip(mom0, eps, grads)]
        state = parts
        for _ in range(L):
            state = [s + e * v for s, e, v in zip(state, eps, half)]
            target, grads = _value_and_gradient(self.target_log_prob_fn, state)
            grads = [g.as_subclass(torch.Tensor) for g in grads]
            half = [v + e * g for v, e, g in zip(half, eps, grads)]
        mom1 = [v - (0.5 * e) * g for v, e, g in zip(half, eps, grads)]
        ke0 = 0.5 * torch.exp(torch.logsumexp(torch.stack([_log_sum_sq(m) for m in mom0]), dim=0))
        ke1 = 0.5 * torch.exp(torch.logsumexp(torch.stack([_log_sum_sq(m) for m in mom1]), dim=0))
        correction = _safe_sum([ke0, -ke1])
        target = target.as_subclass(torch.Tensor)
        proposed = UncalibratedResults(correction, target, grads, mom0, mom1, self.step_size,
                                       self.num_leapfrog_steps)
        # --- MetropolisHastings.one_step
        lar = _safe_sum([target, -acc.target_log_prob.as_subclass(torch.Tensor), correction])
        u = tf.random.uniform(tuple(target.shape), dtype=target.dtype).as_subclass(torch.Tensor)
        is_accepted = torch.log(u) < lar
        nxt = [torch.where(is_accepted, new, old) for new, old in zip(state, parts)]
        accepted = UncalibratedResults(
            torch.where(is_accepted, correction, acc.log_acceptance_correction),
            torch.where(is_accepted, target, acc.target_log_prob.as_subclass(torch.Tensor)),
            [torch.where(is_accepted, gn, go.as_subclass(torch.Tensor)) for gn, go in
             zip(grads, acc.grads_target_log_prob)],
            mom0, mom1, self.step_size, self.num_leapfrog_steps)
        results = MHResults(accepted, is_accepted, lar, state, proposed)
        if mcmc.RESULTS is not None:
            mcmc.RESULTS.append(results)
        nxt = [tf._wrap(s) for s in nxt]
        return (nxt if listlike else nxt[0]), results


def sample_chain(num_results, current_state, previous_kernel_results=None, kernel=None, num_burnin_steps=0,
                 num_steps_between_results=0, trace_fn=None, return_final_kernel_results=False,
                 parallel_iterations=10, seed=None, name=None):
    assert num_burnin_steps == 0 and num_steps_between_results == 0
    listlike = _is_list_like(current_state)
    state = current_state
    if previous_kernel_results is None:
        previous_kernel_results = kernel.bootstrap_results(state)
    states, traces = [], []
    for _ in range(int(num_results)):
        state, previous_kernel_results = kernel.one_step(state, previous_kernel_results)
        states.append(list(state) if listlike else state)
        if trace_fn is not None:
            traces.append(trace_fn(state, previous_kernel_results))
    if listlike:
        all_states = [tf._wrap(torch.stack([s[i].as_subclass(torch.Tensor) for s in states]))
                      for i in range(len(states[0]))]
    else:
        all_states = tf._wrap(torch.stack([s.as_subclass(torch.Tensor) for s in states]))
    if trace_fn is None:
        return all_states
    trace = [tf._wrap(torch.stack([torch.as_tensor(t[i]).detach().as_subclass(torch.Tensor) for t in traces]))
             for i in range(len(traces[0]))]
    return StatesAndTrace(all_states, trace)


class DualAveragingStepSizeAdaptation(object):
    def __init__(self, inner_kernel, num_adaptation_steps, **kwargs):
        self.inner_kernel = inner_kernel
        self.num_adaptation_steps = num_adaptation_steps


mcmc.HamiltonianMonteCarlo = HamiltonianMonteCarlo
mcmc.sample_chain = sample_chain
mcmc.DualAveragingStepSizeAdaptation = DualAveragingStepSizeAdaptation
