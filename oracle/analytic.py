"""Second, independently written CPU implementation of the main and hyper
targets with HAND-DERIVED gradients (numpy, no autograd).

TEST INFRASTRUCTURE (see oracle/__init__.py; parity unpinned).  It guards
oracle/targets.py against shared mistakes and is the blueprint the CUDA
kernels follow.  Same arch / lik / theta / hyper conventions as targets.py,
but theta and hyper are FLAT numpy vectors in ``network.states`` /
``network.hyperStates`` order (W row-major [out,in], b[out], slopes[width]).
"""
import math

import numpy as np

from .tfconst import CLAMP_HI, CLAMP_LO, LOG_2PI_MVLP, f32

LOG_2PI = math.log(2.0 * math.pi)
DENSE = ("dense", "denseGaussian")


def layout(arch):
    """Offsets of every tensor in the flat theta / hyper vectors."""
    items, p, h = [], 0, 0
    for layer in arch:
        k = layer[0]
        if k in DENSE:
            i, o = layer[1], layer[2]
            items.append(dict(kind=k, i=i, o=o, w=p, b=p + o * i, h=h))
            p += o * i + o
            h += 4
        elif k in ("prelu", "squareprelu"):
            items.append(dict(kind=k, n=layer[1], s=p, h=h))
            p += layer[1]
            h += 1 if k == "prelu" else 2
        else:
            items.append(dict(kind=k, alpha=layer[1] if k == "leakyrelu" else None))
    return items, p, h


def _act_fwd(it, z, theta):
    k = it["kind"]
    if k == "relu":
        return np.maximum(z, 0)
    if k == "tanh":
        return np.tanh(z)
    if k == "sigmoid":
        return 1.0 / (1.0 + np.exp(-z))
    if k == "Exp":
        return np.exp(z)
    if k == "elu":
        return np.where(z > 0, z, np.expm1(z))
    if k == "leakyrelu":
        return np.where(z < 0, it["alpha"] * z, z)
    if k == "prelu":
        s = theta[it["s"]:it["s"] + it["n"]][:, None]
        return np.where(z < 0, s * z, z)
    if k == "squareprelu":
        s = theta[it["s"]:it["s"] + it["n"]][:, None] ** 2
        return np.where(z < 0, s * z, z)
    raise ValueError(k)


def _act_bwd(it, z, a, da, theta, grad):
    """Returns dz; adds slope gradients into ``grad``."""
    k = it["kind"]
    if k == "relu":
        return da * (z > 0)
    if k == "tanh":
        return da * (1.0 - a * a)
    if k == "sigmoid":
        return da * a * (1.0 - a)
    if k == "Exp":
        return da * a
    if k == "elu":
        return da * np.where(z > 0, 1.0, a + 1.0)
    if k == "leakyrelu":
        return da * np.where(z < 0, it["alpha"], 1.0)
    if k == "prelu":
        sl = theta[it["s"]:it["s"] + it["n"]]
        neg = z < 0
        grad[it["s"]:it["s"] + it["n"]] += np.sum(np.where(neg, z * da, 0.0), axis=1)
        return da * np.where(neg, sl[:, None], 1.0)
    if k == "squareprelu":
        sl = theta[it["s"]:it["s"] + it["n"]]
        neg = z < 0
        grad[it["s"]:it["s"] + it["n"]] += 2.0 * sl * np.sum(np.where(neg, z * da, 0.0), axis=1)
        return da * np.where(neg, (sl ** 2)[:, None], 1.0)
    raise ValueError(k)


def loglik_and_grad(arch, lik, theta, X, Y, sd_hyper=None):
    """Summed log-likelihood, its gradient w.r.t. theta, SSE (Gaussian kinds)
    and d loglik / d sd_hyper (Gaussian likelihood only)."""
    items, P, _ = layout(arch)
    grad = np.zeros(P, dtype=theta.dtype)
    a = X.T.copy()
    tape = []
    for it in items:
        if it["kind"] in DENSE:
            W = theta[it["w"]:it["w"] + it["o"] * it["i"]].reshape(it["o"], it["i"])
            b = theta[it["b"]:it["b"] + it["o"]]
            z = W @ a + b[:, None]
            tape.append((it, a, None))
            a = z
        else:
            z = a
            a = _act_fwd(it, z, theta)
            tape.append((it, z, a))
    f = a                                           # [out, N]
    n_out, N = f.shape
    y = np.asarray(Y, dtype=theta.dtype).reshape(N, n_out).T
    sse = None
    dsd = None
    if lik[0] in ("gaussian", "fixed"):
        sd = sd_hyper ** 2 if lik[0] == "gaussian" else f32(lik[1])     # tf.cast(self.sd, dtype), Q14
        sg = min(max(sd, CLAMP_LO), CLAMP_HI)
        r = y - f
        sse = float(np.sum(r * r))
        ll = -0.5 * (2.0 * N * n_out * math.log(sg) + sse / sg ** 2 + N * n_out * LOG_2PI_MVLP)
        df = r / sg ** 2
        if lik[0] == "gaussian":
            inside = CLAMP_LO <= sd <= CLAMP_HI
            dll_dsg = -(N * n_out) / sg + sse / sg ** 3
            dsd = (dll_dsg if inside else 0.0) * 2.0 * sd_hyper
    else:
        lo = theta.dtype.type(1e-8)
        hi = theta.dtype.type(1 - 1e-7)
        p = np.clip(f, lo, hi)
        ll = float(np.sum((1.0 - y) * np.log1p(-p) + y * np.log(p)))
        df = np.where((f < lo) | (f > hi), 0.0, y / p - (1.0 - y) / (1.0 - p))
    da = df
    for it, u, v in reversed(tape):
        if it["kind"] in DENSE:
            a_in = u
            W = theta[it["w"]:it["w"] + it["o"] * it["i"]].reshape(it["o"], it["i"])
            grad[it["w"]:it["w"] + it["o"] * it["i"]] += (da @ a_in.T).reshape(-1)
            grad[it["b"]:it["b"] + it["o"]] += da.sum(axis=1)
            da = W.T @ da
        else:
            da = _act_bwd(it, u, v, da, theta, grad)
    return float(ll), grad, sse, dsd


def _prior_terms(it, theta, hyper, want_hyper_grad):
    """(value, d/dtheta as (offset, array) list, d/dhyper dict) of one layer's
    calculateProbs."""
    k = it["kind"]
    h = hyper[it["h"]:]
    dth, dh = [], {}
    if k == "dense":
        val = 0.0
        for (off, n, hx0, hg) in ((it["w"], it["o"] * it["i"], 0, 1), (it["b"], it["o"], 2, 3)):
            x = theta[off:off + n]
            g = h[hg] ** 2
            z = (x - h[hx0]) / g
            val += np.sum(np.log1p(z * z)) - n * math.log(math.pi * g)
            q = 2.0 * z / (1.0 + z * z)
            dth.append((off, q / g))
            if want_hyper_grad:
                dh[it["h"] + hx0] = np.sum(-q / g)
                dh[it["h"] + hg] = (np.sum(-q * z / g) - n / g) * 2.0 * h[hg]
        return val, dth, dh
    if k == "denseGaussian":
        val = 0.0
        for (off, n, hm, hs) in ((it["w"], it["o"] * it["i"], 0, 1), (it["b"], it["o"], 2, 3)):
            x = theta[off:off + n]
            s = h[hs] ** 2
            sg = min(max(s, CLAMP_LO), CLAMP_HI)
            d = x - h[hm]
            ss = np.sum(d * d)
            val += -0.5 * (2.0 * math.log(sg) + ss / sg ** 2 + LOG_2PI_MVLP)
            dth.append((off, -d / sg ** 2))
            if want_hyper_grad:
                dh[it["h"] + hm] = np.sum(d) / sg ** 2
                inside = CLAMP_LO <= s <= CLAMP_HI
                dh[it["h"] + hs] = ((-1.0 / sg + ss / sg ** 3) if inside else 0.0) * 2.0 * h[hs]
        return val, dth, dh
    raise ValueError(k)


def main_value_and_grad(arch, lik, theta, hyper, X, Y):
    items, P, _ = layout(arch)
    sdh = hyper[-1] if lik[0] == "gaussian" else None
    ll, grad, _, _ = loglik_and_grad(arch, lik, theta, X, Y, sdh)
    val = ll
    for it in items:
        k = it["kind"]
        if k in DENSE:
            v, dth, _ = _prior_terms(it, theta, hyper, False)
            val += v
            for off, g in dth:
                grad[off:off + g.size] += g
        elif k == "squareprelu":                   # Q4 "as written": prior on the un-squared slope
            a = theta[it["s"]:it["s"] + it["n"]]
            mean, sd = hyper[it["h"]], hyper[it["h"] + 1]
            sg = min(max(sd, CLAMP_LO), CLAMP_HI)
            val += -0.5 * (2.0 * math.log(sg) + np.sum(((a - mean) / sg) ** 2) + LOG_2PI_MVLP)
            grad[it["s"]:it["s"] + it["n"]] += -(a - mean) / sg ** 2
        elif k == "prelu":
            a = theta[it["s"]:it["s"] + it["n"]]
            r = abs(hyper[it["h"]])
            val += np.sum(-r * a) + it["n"] * math.log(r)
            grad[it["s"]:it["s"] + it["n"]] += -r
    return float(val), grad


def hyper_value_and_grad(arch, lik, theta, hyper, X, Y, sse=None):
    """Hyper target and gradient.  ``sse`` may be supplied (sufficient
    statistic of the Gaussian likelihood at fixed theta)."""
    items, _, Hl = layout(arch)
    H = hyper.size
    g = np.zeros(H, dtype=hyper.dtype)
    val = 0.0

    def logn(v, m, s):
        m, s = f32(m), f32(s)                       # float32 hyper-prior constants (Q14)
        return -0.5 * ((v - m) / s) ** 2 - math.log(s) - 0.5 * LOG_2PI

    for it in items:
        k = it["kind"]
        o = it.get("h")
        if k in DENSE:
            v, _, dh = _prior_terms(it, theta, hyper, True)
            val += v
            for idx, d in dh.items():
                g[idx] += d
            if k == "dense":
                loc, sc, lm, ls = 0.0, f32(0.2), f32(0.5 ** 0.5), 0.5
            else:
                loc, sc, lm, ls = 0.0, f32(0.1), 1.0, f32(0.1)
            for j in (0, 2):
                val += logn(hyper[o + j], loc, sc)
                g[o + j] += -(hyper[o + j] - loc) / sc ** 2
            for j in (1, 3):
                s2 = hyper[o + j] ** 2
                val += logn(s2, lm, ls)
                g[o + j] += -(s2 - lm) / ls ** 2 * 2.0 * hyper[o + j]
        elif k == "squareprelu":
            a2 = theta[it["s"]:it["s"] + it["n"]] ** 2
            mean, sd = hyper[o], hyper[o + 1]
            sg = min(max(sd, CLAMP_LO), CLAMP_HI)
            d = a2 - mean
            ss = np.sum(d * d)
            val += -0.5 * (2.0 * math.log(sg) + ss / sg ** 2 + LOG_2PI_MVLP)
            val += logn(mean, 0.0, 0.3) + logn(sd, 0.3, 0.1)
            g[o] += np.sum(d) / sg ** 2 - mean / f32(0.3) ** 2
            inside = CLAMP_LO <= sd <= CLAMP_HI
            g[o + 1] += ((-1.0 / sg + ss / sg ** 3) if inside else 0.0) - (sd - f32(0.3)) / f32(0.1) ** 2
        elif k == "prelu":
            a = np.abs(theta[it["s"]:it["s"] + it["n"]])
            r = hyper[o]
            ar = abs(r)
            sgn = 1.0 if r > 0 else (-1.0 if r < 0 else 0.0)
            val += -f32(0.3) * r + math.log(f32(0.3))
            val += np.sum(-ar * a) + it["n"] * math.log(ar)
            g[o] += -f32(0.3) + sgn * (-np.sum(a) + it["n"] / ar)
    if lik[0] == "gaussian":
        sdh = hyper[-1]
        sd = sdh ** 2
        sg = min(max(sd, CLAMP_LO), CLAMP_HI)
        if sse is None:
            _, _, sse, _ = loglik_and_grad(arch, lik, theta, X, Y, sdh)
        n = np.asarray(Y).size
        val += -0.5 * (2.0 * n * math.log(sg) + sse / sg ** 2 + n * LOG_2PI_MVLP)
        inside = CLAMP_LO <= sd <= CLAMP_HI
        g[-1] += ((-n / sg + sse / sg ** 3) if inside else 0.0) * 2.0 * sdh
    return float(val), g
