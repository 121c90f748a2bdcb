"""Log-density targets of the TensorBNN sampler, restated with torch on CPU.

TEST INFRASTRUCTURE (see oracle/__init__.py; parity unpinned).

Architecture vocabulary (mirrors the ``name`` strings of the reference's layer
classes, predictor.py:30-34):

    arch = [("dense", in, out) | ("denseGaussian", in, out)
            | ("relu",) | ("tanh",) | ("sigmoid",) | ("Exp",) | ("elu",)
            | ("leakyrelu", alpha) | ("prelu", width) | ("squareprelu", width)]
    lik  = ("gaussian",) | ("fixed", sd) | ("bernoulli",)

``theta`` is the list ``network.states`` (network.py:173-187): per dense layer
W[out,in], b[out,1]; per prelu/squareprelu a slope vector [width].
``hyper`` is the list ``network.hyperStates`` (network.py:189-191,542-543):
4 scalars per dense layer, 1 per prelu, 2 per squareprelu, then the
likelihood's own (1 for the Gaussian likelihood), each a 0-d/1-element tensor.
"""
import math

import torch

from .tfconst import CLAMP_HI, CLAMP_LO, LOG_2PI_MVLP, f32

DENSE_KINDS = ("dense", "denseGaussian")
PARAM_ACTS = ("prelu", "squareprelu")
PLAIN_ACTS = ("relu", "tanh", "sigmoid", "Exp", "elu", "leakyrelu", "softmax")

LOG_2PI = math.log(2.0 * math.pi)


# ----------------------------------------------------------------------------
# bookkeeping
# ----------------------------------------------------------------------------
def num_tensors(layer):
    """numTensors / numHyperTensors of a layer (layer.py:127-128,308-309;
    activationFunctions.py:136-137,293-294; Leaky_relu is treated as
    stateless)."""
    k = layer[0]
    if k in DENSE_KINDS:
        return 2, 4
    if k == "prelu":
        return 1, 1
    if k == "squareprelu":
        return 1, 2
    return 0, 0


def theta_shapes(arch):
    shapes = []
    for layer in arch:
        k = layer[0]
        if k in DENSE_KINDS:
            shapes.append((layer[2], layer[1]))
            shapes.append((layer[2], 1))
        elif k in PARAM_ACTS:
            shapes.append((layer[1],))
    return shapes


def num_params(arch):
    return sum(math.prod(s) for s in theta_shapes(arch))


def num_hypers(arch, lik):
    h = sum(num_tensors(layer)[1] for layer in arch)
    return h + (1 if lik[0] == "gaussian" else 0)


def unflatten_theta(arch, flat):
    out, off = [], 0
    for s in theta_shapes(arch):
        n = math.prod(s)
        out.append(flat[off:off + n].reshape(s))
        off += n
    assert off == flat.numel()
    return out


def flatten_theta(theta):
    return torch.cat([t.reshape(-1) for t in theta])


def unflatten_hyper(flat):
    return [flat[i] for i in range(flat.numel())]


def initial_hypers(arch, lik, dtype=torch.float64):
    """Start values: layer.py:136-158 (Cauchy: 0, sqrt(.5), 0, sqrt(.5)),
    layer.py:317-338 (Gaussian: 0,1,0,1), activationFunctions.py:144-149
    (prelu rate .3), :301-317 (squareprelu mean 0, sd .3), likelihood.py:66
    (sqrt(sd))."""
    h = []
    for layer in arch:
        k = layer[0]
        if k == "dense":
            h += [0.0, 0.5 ** 0.5, 0.0, 0.5 ** 0.5]
        elif k == "denseGaussian":
            h += [0.0, 1.0, 0.0, 1.0]
        elif k == "prelu":
            h += [0.3]
        elif k == "squareprelu":
            h += [0.0, 0.3]
    if lik[0] == "gaussian":
        h.append(lik[1] ** 0.5 if len(lik) > 1 else 0.1 ** 0.5)
    return torch.tensor(h, dtype=dtype)


# ----------------------------------------------------------------------------
# densities
# ----------------------------------------------------------------------------
def multivariate_log_prob(sigma, mu, x):
    """BNN_functions.py:21-32.  sigma clamped to [1e-8, 1e8]; logDet and the
    k*log(2pi) term count the elements OF SIGMA (scalar sigma => once, Q2)."""
    dt = x.dtype
    sigma = torch.as_tensor(sigma, dtype=dt)
    sigma = torch.clamp(sigma, min=CLAMP_LO, max=CLAMP_HI)    # tf.cast(10**-8, dtype): float32-rounded (Q14)
    log_det = 2.0 * torch.sum(torch.log(sigma))
    k = float(sigma.numel())
    dif = (1.0 / sigma) * (x - mu)
    return -0.5 * (log_det + torch.sum(dif * dif) + k * LOG_2PI_MVLP)


def cauchy_log_prob(gamma, x0, x):
    """BNN_functions.py:51-56: +log(1+z^2) - log(pi*gamma) elementwise (the
    sign of the first term is the reference's, Q1)."""
    return torch.log(1.0 + ((x - x0) / gamma) ** 2) - torch.log(math.pi * gamma)


def log_normal_1d(v, m, s):
    """tfd.MultivariateNormalDiag(loc=[m], scale_diag=[s]).log_prob([v]); loc and scale are python lists,
    i.e. float32 constants (Q14)."""
    m, s = f32(m), f32(s)
    return -0.5 * ((v - m) / s) ** 2 - math.log(s) - 0.5 * LOG_2PI


def exponential_log_prob(rate, x):
    """activationFunctions.py:172-173."""
    rate = torch.abs(rate)
    return -rate * x + torch.log(rate)


# ----------------------------------------------------------------------------
# forward pass
# ----------------------------------------------------------------------------
def apply_activation(layer, a, tensors):
    k = layer[0]
    if k == "relu":
        return torch.relu(a)                       # activationFunctions.py:36
    if k == "tanh":
        return torch.tanh(a)                       # :62
    if k == "sigmoid":
        return torch.sigmoid(a)                    # :49
    if k == "Exp":
        return torch.exp(a)                        # :23
    if k == "elu":
        return torch.nn.functional.elu(a)          # :75
    if k == "softmax":
        return torch.softmax(a, dim=-1)            # :88 (last axis = rows)
    if k == "leakyrelu":
        return torch.where(a < 0, layer[1] * a, a)  # :105, constant slope (Q6)
    if k == "prelu":
        s = tensors[0].reshape(-1, 1)              # :250-254
        return torch.where(a < 0, s * a, a)
    if k == "squareprelu":
        s = (tensors[0] ** 2).reshape(-1, 1)       # :412-416
        return torch.where(a < 0, s * a, a)
    raise ValueError(k)


def forward(arch, theta, X):
    """network.py:160-169: A0 = X^T [D,N]; dense W@A+b (layer.py:278);
    returns [out, N]."""
    a = X.t()
    idx = 0
    for layer in arch:
        nt, _ = num_tensors(layer)
        ts = theta[idx:idx + nt]
        idx += nt
        if layer[0] in DENSE_KINDS:
            a = ts[0] @ a + ts[1].reshape(-1, 1)
        else:
            a = apply_activation(layer, a, ts)
    return a


# ----------------------------------------------------------------------------
# likelihoods (likelihood.py)
# ----------------------------------------------------------------------------
def log_likelihood(arch, lik, theta, X, Y, sd_hyper=None):
    """makeResponseLikelihood summed to a scalar (network.py:388).
    gaussian: sigma = hyperStates[-1]**2 (likelihood.py:88-94);
    fixed:    sigma = sd                  (likelihood.py:162-167);
    bernoulli: clip(f,1e-8,1-1e-7), (1-y)log1p(-p)+y log p (likelihood.py:225-236)."""
    f = forward(arch, theta, X)
    dt = f.dtype
    if lik[0] in ("gaussian", "fixed"):
        cur = f.t()
        sd = sd_hyper ** 2 if lik[0] == "gaussian" else torch.tensor(f32(lik[1]), dtype=dt)   # tf.cast(self.sd, dtype)
        sigma = torch.ones_like(cur) * sd
        real = Y.reshape(cur.shape)
        return multivariate_log_prob(sigma, cur, real)
    if lik[0] == "bernoulli":
        lo = torch.tensor(1e-8, dtype=dt)
        hi = torch.tensor(1 - 1e-7, dtype=dt)
        p = torch.clamp(f, min=lo.item(), max=hi.item())
        y = Y.t() if Y.dim() == 2 else Y
        return torch.sum((1.0 - y) * torch.log1p(-p) + y * torch.log(p))
    raise ValueError(lik)


def sse(arch, theta, X, Y):
    f = forward(arch, theta, X).t()
    return torch.sum((Y.reshape(f.shape) - f) ** 2)


# ----------------------------------------------------------------------------
# per-layer priors / hyper conditionals
# ----------------------------------------------------------------------------
def layer_prior(layer, hypers, tensors):
    """calculateProbs(hypers, tensors): layer.py:166-197 (Cauchy),
    :346-377 (Gaussian); activationFunctions.py:177-192 / :329-348 with the
    hyper slice passed in (the minimal patch of Q4)."""
    k = layer[0]
    if k == "dense":
        g_w, g_b = hypers[1] ** 2, hypers[3] ** 2
        return (torch.sum(cauchy_log_prob(g_w, hypers[0], tensors[0]))
                + torch.sum(cauchy_log_prob(g_b, hypers[2], tensors[1])))
    if k == "denseGaussian":
        return (multivariate_log_prob(hypers[1] ** 2, hypers[0], tensors[0])
                + multivariate_log_prob(hypers[3] ** 2, hypers[2], tensors[1]))
    if k == "squareprelu":
        return multivariate_log_prob(hypers[1], hypers[0], tensors[0])
    if k == "prelu":
        return torch.sum(exponential_log_prob(hypers[0], tensors[0]))
    raise ValueError(k)


def layer_hyper_prob(layer, hypers, tensors):
    """calculateHyperProbs(hypers, tensors): layer.py:199-242, :379-422;
    activationFunctions.py:194-220, :350-382.  Scale hyper-priors are
    evaluated at h**2 with no Jacobian (Q3)."""
    k = layer[0]
    if k == "dense":
        r5 = 0.5 ** 0.5
        p = (log_normal_1d(hypers[0], 0.0, 0.2) + log_normal_1d(hypers[1] ** 2, r5, 0.5)
             + log_normal_1d(hypers[2], 0.0, 0.2) + log_normal_1d(hypers[3] ** 2, r5, 0.5))
        return p + layer_prior(layer, hypers, tensors)
    if k == "denseGaussian":
        p = (log_normal_1d(hypers[0], 0.0, 0.1) + log_normal_1d(hypers[1] ** 2, 1.0, 0.1)
             + log_normal_1d(hypers[2], 0.0, 0.1) + log_normal_1d(hypers[3] ** 2, 1.0, 0.1))
        return p + layer_prior(layer, hypers, tensors)
    if k == "squareprelu":
        return (multivariate_log_prob(hypers[1], hypers[0], tensors[0] ** 2)
                + log_normal_1d(hypers[0], 0.0, 0.3) + log_normal_1d(hypers[1], 0.3, 0.1))
    if k == "prelu":
        dt = tensors[0].dtype
        return (exponential_log_prob(torch.tensor(f32(0.3), dtype=dt), hypers[0])
                + torch.sum(exponential_log_prob(hypers[0], torch.abs(tensors[0]))))
    raise ValueError(k)


# ----------------------------------------------------------------------------
# the two targets
# ----------------------------------------------------------------------------
def main_log_prob(arch, lik, theta, hyper, X, Y):
    """network.py:370-392: sum of layer priors for layers with
    numHyperTensors>0 (index advances only for those) + summed likelihood."""
    prob = 0.0
    ih = it = 0
    for layer in arch:
        nt, nh = num_tensors(layer)
        if nh > 0:
            prob = prob + layer_prior(layer, hyper[ih:ih + nh], theta[it:it + nt])
            ih += nh
            it += nt
    return prob + log_likelihood(arch, lik, theta, X, Y, sd_hyper=hyper[-1] if len(hyper) else None)


def hyper_log_prob(arch, lik, theta, hyper, X, Y):
    """network.py:417-440."""
    prob = 0.0
    ih = it = 0
    for layer in arch:
        nt, nh = num_tensors(layer)
        if nh > 0:
            prob = prob + layer_hyper_prob(layer, hyper[ih:ih + nh], theta[it:it + nt])
            ih += nh
            it += nt
    if lik[0] == "gaussian":                       # mainProbsInHypers, likelihood.py:67
        prob = prob + log_likelihood(arch, lik, theta, X, Y, sd_hyper=hyper[-1])
    return prob


# ----------------------------------------------------------------------------
# flat-vector value-and-gradient wrappers
# ----------------------------------------------------------------------------
def main_value_and_grad(arch, lik, theta_flat, hyper_flat, X, Y):
    th = theta_flat.detach().clone().requires_grad_(True)
    lp = main_log_prob(arch, lik, unflatten_theta(arch, th), unflatten_hyper(hyper_flat), X, Y)
    (g,) = torch.autograd.grad(lp, th)
    return lp.detach(), g


def hyper_value_and_grad(arch, lik, theta_flat, hyper_flat, X, Y):
    hy = hyper_flat.detach().clone().requires_grad_(True)
    lp = hyper_log_prob(arch, lik, unflatten_theta(arch, theta_flat), unflatten_hyper(hy), X, Y)
    (g,) = torch.autograd.grad(lp, hy, allow_unused=True)
    if g is None:
        g = torch.zeros_like(hy)
    return lp.detach(), g
