"""Pins the oracle (parity unpinned against the reference, which has no
golden vectors): autograd restatement vs hand-derived analytic gradients vs
central finite differences vs closed forms."""
import math

import numpy as np
import pytest
import torch

from oracle import tfconst, analytic, targets
from tensorbnn_b200 import workloads as wl

torch.set_default_dtype(torch.float64)

ARCHS = {
    "c1a": (wl.mlp_arch([1, 10, 10, 10, 1], "denseGaussian", "tanh"), ("fixed", 0.1)),
    "c1b": (wl.mlp_arch([1, 10, 10, 10, 1], "dense", "relu"), ("gaussian", 0.1)),
    "bern": (wl.mlp_arch([7, 5, 4, 1], "dense", "relu", "sigmoid"), ("bernoulli",)),
    "sqp": (wl.mlp_arch([3, 6, 6, 2], "dense", "squareprelu"), ("gaussian", 0.2)),
    "prelu": (wl.mlp_arch([3, 6, 5, 1], "denseGaussian", "prelu"), ("fixed", 0.3)),
    "mixed": ([("dense", 4, 6), ("elu",), ("denseGaussian", 6, 5), ("Exp",), ("dense", 5, 3),
               ("leakyrelu", 0.3), ("dense", 3, 1), ("sigmoid",)], ("bernoulli",)),
}


def make_problem(key, N=13, seed=0):
    arch, lik = ARCHS[key]
    rng = np.random.default_rng(seed)
    D = arch[0][1]
    out = [l for l in arch if l[0] in ("dense", "denseGaussian")][-1][2]
    X = rng.normal(size=(N, D))
    if lik[0] == "bernoulli":
        Y = (rng.random(N) > 0.5).astype(np.float64)
    else:
        Y = rng.normal(size=(N, out)) if out > 1 else rng.normal(size=N)
    theta = wl.init_theta(arch, seed=seed + 5) * 0.7
    # perturb slopes so prelu paths are not symmetric
    theta = theta + 0.05 * rng.normal(size=theta.size)
    hyper = wl.init_hyper(arch, lik) + 0.05 * rng.normal(size=wl.init_hyper(arch, lik).size)
    return arch, lik, X, Y, theta, hyper


@pytest.mark.parametrize("key", list(ARCHS))
def test_main_autograd_vs_analytic(key):
    arch, lik, X, Y, theta, hyper = make_problem(key)
    lp, g = targets.main_value_and_grad(arch, lik, torch.tensor(theta), torch.tensor(hyper),
                                        torch.tensor(X), torch.tensor(Y))
    lp2, g2 = analytic.main_value_and_grad(arch, lik, theta, hyper, X, Y)
    assert abs(lp.item() - lp2) <= 1e-11 * max(1.0, abs(lp2))
    np.testing.assert_allclose(g.numpy(), g2, rtol=1e-9, atol=1e-10 * np.abs(g2).max())


@pytest.mark.parametrize("key", list(ARCHS))
def test_hyper_autograd_vs_analytic(key):
    arch, lik, X, Y, theta, hyper = make_problem(key)
    lp, g = targets.hyper_value_and_grad(arch, lik, torch.tensor(theta), torch.tensor(hyper),
                                         torch.tensor(X), torch.tensor(Y))
    lp2, g2 = analytic.hyper_value_and_grad(arch, lik, theta, hyper, X, Y)
    assert abs(lp.item() - lp2) <= 1e-11 * max(1.0, abs(lp2))
    np.testing.assert_allclose(g.numpy(), g2, rtol=1e-9, atol=1e-10 * max(1.0, np.abs(g2).max()))


@pytest.mark.parametrize("key", ["c1a", "c1b", "bern", "sqp", "prelu"])
def test_main_finite_differences(key):
    arch, lik, X, Y, theta, hyper = make_problem(key)
    _, g = analytic.main_value_and_grad(arch, lik, theta, hyper, X, Y)
    rng = np.random.default_rng(1)
    for idx in rng.choice(theta.size, size=12, replace=False):
        h = 1e-6
        tp, tm = theta.copy(), theta.copy()
        tp[idx] += h
        tm[idx] -= h
        fd = (analytic.main_value_and_grad(arch, lik, tp, hyper, X, Y)[0]
              - analytic.main_value_and_grad(arch, lik, tm, hyper, X, Y)[0]) / (2 * h)
        assert abs(fd - g[idx]) <= 1e-5 * max(1.0, abs(g[idx])), (idx, fd, g[idx])


@pytest.mark.parametrize("key", ["c1a", "c1b", "sqp", "prelu"])
def test_hyper_finite_differences(key):
    arch, lik, X, Y, theta, hyper = make_problem(key)
    _, g = analytic.hyper_value_and_grad(arch, lik, theta, hyper, X, Y)
    for idx in range(hyper.size):
        h = 1e-6
        hp, hm = hyper.copy(), hyper.copy()
        hp[idx] += h
        hm[idx] -= h
        fd = (analytic.hyper_value_and_grad(arch, lik, theta, hp, X, Y)[0]
              - analytic.hyper_value_and_grad(arch, lik, theta, hm, X, Y)[0]) / (2 * h)
        assert abs(fd - g[idx]) <= 2e-5 * max(1.0, abs(g[idx])), (idx, fd, g[idx])


def test_known_answer_zero_weights_c1():
    """C1 data, all-zero weights => f == 0, closed-form logp."""
    cfg = wl.c1("a")
    arch, lik = cfg["arch"], cfg["lik"]
    P = sum(math.prod(s) for s in wl.theta_shapes(arch))
    theta = np.zeros(P)
    hyper = wl.init_hyper(arch, lik)
    lp, _ = analytic.main_value_and_grad(arch, lik, theta, hyper, cfg["X"], cfg["Y"])
    y = cfg["Y"]
    # constants as TensorFlow materialises them (tf.cast of a python float goes through float32: oracle/tfconst.py, Q14)
    sd, log2pi = tfconst.f32(0.1), tfconst.LOG_2PI_MVLP
    ll = -0.5 * (2 * 11 * math.log(sd) + np.sum(y ** 2) / sd ** 2 + 11 * log2pi)
    # 4 Gaussian dense layers x (W,b), sigma=1, mu=0, all-zero tensors: -0.5*log(2pi) each (Q2)
    prior = 8 * (-0.5 * log2pi)
    assert abs(lp - (ll + prior)) < 1e-10
    lp_t, _ = targets.main_value_and_grad(arch, lik, torch.tensor(theta), torch.tensor(hyper),
                                          torch.tensor(cfg["X"]), torch.tensor(cfg["Y"]))
    assert abs(lp_t.item() - (ll + prior)) < 1e-10


def test_known_answer_single_dense_n1():
    """one dense layer, one row, each likelihood by hand."""
    X, w, b = np.array([[2.0]]), 0.5, -0.25
    f = w * 2.0 + b                                                  # 0.75
    theta = np.array([w, b])
    # Cauchy dense prior with gamma = (sqrt(.5))^2 = .5, x0 = 0  (Q1 sign)
    prior = (math.log(1 + (w / 0.5) ** 2) - math.log(math.pi * 0.5)
             + math.log(1 + (b / 0.5) ** 2) - math.log(math.pi * 0.5))
    arch = [("dense", 1, 1)]
    y = np.array([1.0])
    hy = wl.init_hyper(arch, ("fixed", 0.5))
    lp, _ = analytic.main_value_and_grad(arch, ("fixed", 0.5), theta, hy, X, y)
    ll = -0.5 * (2 * math.log(0.5) + ((1.0 - f) / 0.5) ** 2 + tfconst.LOG_2PI_MVLP)
    assert abs(lp - (prior + ll)) < 1e-12
    hy = wl.init_hyper(arch, ("gaussian", 0.09))                     # hyper = 0.3, sigma = 0.09
    lp, _ = analytic.main_value_and_grad(arch, ("gaussian", 0.09), theta, hy, X, y)
    ll = -0.5 * (2 * math.log(0.09) + ((1.0 - f) / 0.09) ** 2 + tfconst.LOG_2PI_MVLP)
    assert abs(lp - (prior + ll)) < 1e-9
    arch_b = [("dense", 1, 1), ("sigmoid",)]
    hy = wl.init_hyper(arch_b, ("bernoulli",))
    p = 1 / (1 + math.exp(-f))
    for yy in (0.0, 1.0):
        lp, _ = analytic.main_value_and_grad(arch_b, ("bernoulli",), theta, hy, X, np.array([yy]))
        ll = (1 - yy) * math.log1p(-p) + yy * math.log(p)
        assert abs(lp - (prior + ll)) < 1e-12


def test_clip_and_clamp_edges():
    """Bernoulli clip at 1e-8 / fp32(1-1e-7); sigma clamp at 1e-8."""
    arch = [("dense", 1, 1), ("sigmoid",)]
    hy = torch.tensor(wl.init_hyper(arch, ("bernoulli",)), dtype=torch.float32)
    X = torch.tensor([[1.0], [1.0]], dtype=torch.float32)
    Y = torch.tensor([0.0, 1.0], dtype=torch.float32)
    th = torch.tensor([40.0, 0.0], dtype=torch.float32)            # p -> 1, clipped
    ll = targets.log_likelihood(arch, ("bernoulli",), targets.unflatten_theta(arch, th), X, Y)
    hi = np.float32(1 - 1e-7)
    assert hi == np.float32(0.99999988)
    expect = np.log1p(-np.float64(hi)) + np.log(np.float64(hi))
    assert abs(ll.item() - expect) < 1e-4 * abs(expect)
    _, g = targets.main_value_and_grad(arch, ("bernoulli",), th, hy, X, Y)
    # likelihood gradient is zero where clipped; only the Cauchy prior term remains
    z = 40.0 / 0.5
    assert abs(g[0].item() - (2 * z / (1 + z * z)) / 0.5) < 1e-6
    th = torch.tensor([-40.0, 0.0], dtype=torch.float32)           # p -> 0, clipped at 1e-8
    ll = targets.log_likelihood(arch, ("bernoulli",), targets.unflatten_theta(arch, th), X, Y)
    assert abs(ll.item() - (math.log1p(-1e-8) + math.log(1e-8))) < 1e-4
    # sigma clamp: multivariateLogProb with sigma below 1e-8 uses 1e-8 and has zero sigma-gradient
    s = torch.tensor(1e-12, requires_grad=True)
    v = targets.multivariate_log_prob(s, 0.0, torch.tensor([1e-9]))
    lo = tfconst.CLAMP_LO
    expect = -0.5 * (2 * math.log(lo) + (1e-9 / lo) ** 2 + tfconst.LOG_2PI_MVLP)
    assert abs(v.item() - expect) < 1e-9
    v.backward()
    assert s.grad.item() == 0.0


def test_fp32_oracle_close_to_fp64():
    for key in ("c1a", "bern", "sqp"):
        arch, lik, X, Y, theta, hyper = make_problem(key, N=64)
        lp64, g64 = analytic.main_value_and_grad(arch, lik, theta, hyper, X, Y)
        f32 = torch.float32
        lp32, g32 = targets.main_value_and_grad(arch, lik, torch.tensor(theta, dtype=f32),
                                                torch.tensor(hyper, dtype=f32),
                                                torch.tensor(X, dtype=f32), torch.tensor(Y, dtype=f32))
        assert abs(lp32.item() - lp64) <= 2e-5 * abs(lp64)
        assert np.abs(g32.numpy() - g64).max() <= 2e-5 * np.abs(g64).max()
