"""Synthetic workloads C1-C5 of SURVEY.md section 8(d) (shapes of BASELINE.json's
configs).  numpy only; used by bench.py and the parity tests.

arch / lik use the vocabulary of the reference's layer ``name`` strings
(predictor.py:30-34): ("dense", in, out), ("denseGaussian", in, out),
("relu",), ("tanh",), ("sigmoid",), ("squareprelu", width), ... and
("gaussian", sd) | ("fixed", sd) | ("bernoulli",).
"""
import math

import numpy as np


def mlp_arch(dims, dense="dense", act="relu", last_act=None):
    arch = []
    for i in range(len(dims) - 1):
        arch.append((dense, dims[i], dims[i + 1]))
        a = act if i < len(dims) - 2 else last_act
        if a is not None:
            arch.append((a, dims[i + 1]) if a in ("prelu", "squareprelu") else (a,))
    return arch


def theta_shapes(arch):
    shapes = []
    for layer in arch:
        if layer[0] in ("dense", "denseGaussian"):
            shapes += [(layer[2], layer[1]), (layer[2], 1)]
        elif layer[0] in ("prelu", "squareprelu"):
            shapes.append((layer[1],))
    return shapes


def init_theta(arch, seed=0, slope=0.2, dtype=np.float64):
    """Weights/biases ~ N(0, sqrt(2/out)) (layer.py:253-262), slopes = alpha
    (activationFunctions.py:320-325); numpy default_rng(seed + 1000*layer)."""
    parts, li = [], 0
    for layer in arch:
        if layer[0] in ("dense", "denseGaussian"):
            rng = np.random.default_rng(seed + 1000 * li)
            sd = math.sqrt(2.0 / layer[2])
            parts.append(rng.normal(0.0, sd, size=layer[2] * layer[1]))
            parts.append(rng.normal(0.0, sd, size=layer[2]))
            li += 1
        elif layer[0] in ("prelu", "squareprelu"):
            parts.append(np.full(layer[1], slope))
    return np.concatenate(parts).astype(dtype)


def init_hyper(arch, lik, dtype=np.float64):
    h = []
    for layer in arch:
        if layer[0] == "dense":
            h += [0.0, 0.5 ** 0.5, 0.0, 0.5 ** 0.5]
        elif layer[0] == "denseGaussian":
            h += [0.0, 1.0, 0.0, 1.0]
        elif layer[0] == "prelu":
            h += [0.3]
        elif layer[0] == "squareprelu":
            h += [0.0, 0.3]
    if lik[0] == "gaussian":
        h.append(lik[1] ** 0.5)
    return np.asarray(h, dtype=dtype)


def _teacher(X, dims, seed, act=np.tanh):
    rng = np.random.default_rng(seed)
    a = X
    for i in range(len(dims) - 1):
        W = rng.normal(0, math.sqrt(2.0 / dims[i]), size=(dims[i], dims[i + 1]))
        a = a @ W
        if i < len(dims) - 2:
            a = act(a)
    return a


def c1(variant="a"):
    """Examples/trainRegression.py:33-36 data; C1a = as in the file (Tanh,
    GaussianDenseLayer, FixedGaussianLikelihood(0.1)); C1b = as BASELINE.json
    words it (Relu, DenseLayer, GaussianLikelihood(0.1))."""
    x = np.linspace(-2, 2, num=11)
    xv = np.linspace(-2 + 2 / 30, 2.0 - 2 / 30, num=30)
    f = lambda t: np.sin(t * math.pi * 2) * t - np.cos(t * math.pi)
    if variant == "a":
        arch, lik = mlp_arch([1, 10, 10, 10, 1], "denseGaussian", "tanh"), ("fixed", 0.1)
    else:
        arch, lik = mlp_arch([1, 10, 10, 10, 1], "dense", "relu"), ("gaussian", 0.1)
    return dict(name="C1" + variant, arch=arch, lik=lik, X=x[:, None], Y=f(x),
                Xv=xv[:, None], Yv=f(xv), eps=1e-3, L=1000, hyper_eps=1e-3, hyper_L=100, chains=1)


def c2(N=9600, D=784, seed=21, teacher_seed=3):
    """docs/ClassificationExample.md stand-in: U[0,1)^{N x 784}, labels from a
    fixed random 784-20-20-1 teacher thresholded at its median."""
    rng = np.random.default_rng(seed)
    if N * D > (1 << 27):
        # C2-L (1,048,576 x 784): float32 draws and a chunked teacher keep the host footprint at 3.3 GB
        X = rng.random((N, D), dtype=np.float32)
        t = np.concatenate([_teacher(X[i:i + 65536].astype(np.float64) - 0.5, [D, 20, 20, 1], teacher_seed)[:, 0]
                            for i in range(0, N, 65536)])
    else:
        X = rng.random((N, D))
        t = _teacher(X - 0.5, [D, 20, 20, 1], teacher_seed)[:, 0]
    Y = (t > np.median(t)).astype(np.float64)
    arch = mlp_arch([D, 20, 20, 1], "dense", "relu", "sigmoid")
    return dict(name="C2", arch=arch, lik=("bernoulli",), X=X, Y=Y, eps=1e-3, L=500,
                hyper_eps=1e-5, hyper_L=30, chains=1)


def c3(N=4096, chains=1024, width=64, seed=42):
    """Batched chains: 1-64-64-64-1 SquarePrelu, GaussianLikelihood(0.1)."""
    rng = np.random.default_rng(seed)
    x = np.linspace(-2, 2, num=N)
    y = np.sin(x * math.pi * 2) * x - np.cos(x * math.pi) + 0.1 * rng.normal(size=N)
    arch = mlp_arch([1, width, width, width, 1], "dense", "squareprelu")
    return dict(name="C3", arch=arch, lik=("gaussian", 0.1), X=x[:, None], Y=y, eps=1e-3, L=100,
                hyper_eps=1e-3, hyper_L=100, chains=chains, slope=0.1 ** 0.5)


def c4(N=4194304, D=32, width=128, seed=7, teacher_seed=11):
    """Large-N regression: N(0,1)^{N x 32}, ReLU teacher + 0.1 noise."""
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((N, D), dtype=np.float32).astype(np.float64)
    t = _teacher(X, [D, width, width, width, 1], teacher_seed, act=lambda a: np.maximum(a, 0))[:, 0]
    Y = t + 0.1 * rng.standard_normal(N)
    arch = mlp_arch([D, width, width, width, 1], "dense", "relu")
    return dict(name="C4", arch=arch, lik=("gaussian", 0.1), X=X, Y=Y, eps=1e-4, L=50,
                hyper_eps=1e-3, hyper_L=100, chains=1)


def c5(M=1048576, S=20480, width=64, seed=5):
    """Predictor sweep: S stored samples x M test rows on the C3 net."""
    arch = mlp_arch([1, width, width, width, 1], "dense", "squareprelu")
    th0 = init_theta(arch, seed=1000, slope=0.1 ** 0.5)
    rng = np.random.default_rng(seed)
    samples = th0[None, :] + 0.05 * rng.standard_normal((S, th0.size))
    X = np.linspace(-4, 4, num=M)[:, None]
    return dict(name="C5", arch=arch, samples=samples, X=X)
