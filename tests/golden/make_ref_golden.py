"""Generates tests/golden/ref_*.json and tests/golden/ref_run/: known-answer vectors produced by the
REFERENCE'S OWN SOURCE, run unmodified.

/root/reference/tensorBNN/{BNN_functions,layer,activationFunctions,likelihood,network,paramAdapter,metrics,
predictor}.py are imported as shipped; TensorFlow / TensorFlow-Probability / emcee (not installable here) are
replaced by the small torch-backed stand-ins under oracle/tfshim/ (read their headers: only the tf.* symbols
the reference touches, TF's documented semantics; TFP's HMC restated from its published algorithm; gradients
of the reference's closures from torch autograd).  What is pinned by the reference's code itself:

* ``multivariateLogProb`` / ``cauchyLogProb`` (BNN_functions.py:7-57) incl. the clamp edges;
* every ``predict`` / ``calculateProbs`` / ``calculateHyperProbs`` of layer.py and activationFunctions.py;
* the three ``makeResponseLikelihood`` (likelihood.py:69-96,143-169,210-237) through ``network.predict``;
* ``network.train`` -> ``stepMCMC`` / ``stepMCMCNoHypers`` end to end for one epoch (network.py:280-507,509-670):
  the main target closure + its gradient at every leapfrog step, the trajectory end point, the log-accept
  ratio, the hyper target closure + gradient, the hyper trajectory and the hand-rolled dual averaging;
* ``paramAdapter`` (paramAdapter.py:39-292): a 330-update history incl. calck / calcUCB / gridSearch;
* the sample writer of ``network.train`` and the reader ``predictor.loadNetworks`` / ``predict``
  (network.py:545-663, predictor.py:43-155) on a 2-burn-in / 3-per-file / every-2nd schedule (ref_run/);
* ``metrics.SquaredError / PercentError / Accuracy``.

Momentum and Metropolis uniforms are injected through the stand-in's hooks and stored in the fixture (TF's
stateful RNG stream cannot be reproduced outside TF).  /root/reference does not exist on the
GPU box, so the vectors are committed; rerun here with

    python tests/golden/make_ref_golden.py
"""
import contextlib
import io
import json
import os
import random as pyrandom
import shutil
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
# the repo's own drop-in ``tensorBNN`` package must not shadow the reference's (a namespace package: it has no
# __init__.py), so the repo root joins sys.path only after the reference modules are imported
sys.path = [p for p in sys.path if os.path.abspath(p or ".") != ROOT]
sys.path.insert(0, os.path.join(ROOT, "oracle", "tfshim"))
sys.path.insert(0, REF)

import tensorflow as tf  # noqa: E402  (the stand-in)
import tensorflow_probability as tfp  # noqa: E402
from tensorBNN import BNN_functions as R_fn  # noqa: E402  (the reference, unmodified)
from tensorBNN import activationFunctions as R_act  # noqa: E402
from tensorBNN import layer as R_layer  # noqa: E402
from tensorBNN import likelihood as R_lik  # noqa: E402
from tensorBNN import metrics as R_met  # noqa: E402
from tensorBNN import network as R_net  # noqa: E402
from tensorBNN import paramAdapter as R_pa  # noqa: E402
from tensorBNN import predictor as R_pred  # noqa: E402

sys.path.insert(0, ROOT)
from tensorbnn_b200 import workloads as wl  # noqa: E402  (pure-numpy shape helpers)

assert tf.__file__.startswith(os.path.join(ROOT, "oracle", "tfshim"))
assert R_net.__file__.startswith(REF)

F64 = tf.float64


def r(a):
    if isinstance(a, torch.Tensor):
        a = a.detach().as_subclass(torch.Tensor).numpy()
    return [float(v) for v in np.asarray(a, dtype=np.float64).reshape(-1)]


def quiet():
    return contextlib.redirect_stdout(io.StringIO())


# ---------------------------------------------------------------------------
# reference objects from the (arch, lik) vocabulary used by the tests
# ---------------------------------------------------------------------------
ACT = {"relu": R_act.Relu, "tanh": R_act.Tanh, "sigmoid": R_act.Sigmoid, "Exp": R_act.Exp, "elu": R_act.Elu}


def build_reference_network(arch, lik, X, Y, theta, hyper, npdt=np.float64):
    tfdt = tf.float64 if npdt == np.float64 else tf.float32
    net = R_net.network(tfdt, X.shape[1], X, Y, X, Y)
    off = 0
    for layer in arch:
        if layer[0] in ("dense", "denseGaussian"):
            cls = R_layer.DenseLayer if layer[0] == "dense" else R_layer.GaussianDenseLayer
            nw = layer[1] * layer[2]
            W = tf.cast(theta[off:off + nw].reshape(layer[2], layer[1]), tfdt)
            b = tf.cast(theta[off + nw:off + nw + layer[2]].reshape(layer[2], 1), tfdt)
            off += nw + layer[2]
            net.add(cls(layer[1], layer[2], weights=W, biases=b, dtype=npdt))
        else:
            net.add(ACT[layer[0]]())
    assert off == theta.size
    nh = len(net.hyperStates)
    net.hyperStates = [tf.cast(np.array([hyper[i]]), tfdt) for i in range(nh)]      # shape [1] each, as network.add leaves them
    if lik[0] == "gaussian":
        likelihood = R_lik.GaussianLikelihood(sd=float(hyper[-1]) ** 2)      # hypers = [[sd**0.5]] (likelihood.py:66)
    elif lik[0] == "fixed":
        likelihood = R_lik.FixedGaussianLikelihood(sd=lik[1])
    else:
        likelihood = R_lik.BernoulliLikelihood()
    return net, likelihood


class Draws(object):
    """Hooks for tf.random.normal / uniform: reproducible draws, recorded."""

    def __init__(self, seed):
        self.rng = np.random.default_rng(seed)
        self.normals, self.uniforms = [], []

    def normal(self, shape, dtype):
        z = torch.tensor(self.rng.normal(size=shape), dtype=torch.float64)
        self.normals.append(z.clone())
        return z.to(dtype)

    def uniform(self, shape, dtype):
        u = torch.tensor(self.rng.random(size=shape) * 0.98 + 0.01, dtype=torch.float64)
        self.uniforms.append(u.clone())
        return u.to(dtype)

    def __enter__(self):
        tf.random.normal_hook, tf.random.uniform_hook = self.normal, self.uniform
        return self

    def __exit__(self, *a):
        tf.random.normal_hook = tf.random.uniform_hook = None


NET_CASES = {
    # name: (arch, lik, N, eps, L, hyper_eps, hyper_L)
    "ref_c1a": (wl.mlp_arch([1, 10, 10, 10, 1], "denseGaussian", "tanh"), ("fixed", 0.1), 11, 1e-3, 25, 1e-3, 12),
    "ref_c1b": (wl.mlp_arch([1, 10, 10, 10, 1], "dense", "relu"), ("gaussian", 0.1), 11, 1e-3, 25, 1e-3, 12),
    "ref_bern": (wl.mlp_arch([7, 5, 4, 1], "dense", "relu", "sigmoid"), ("bernoulli",), 23, 2e-3, 20, 2e-3, 10),
    "ref_mix": ([("dense", 4, 6), ("elu",), ("denseGaussian", 6, 5), ("Exp",), ("dense", 5, 3), ("tanh",),
                 ("dense", 3, 1), ("sigmoid",)], ("bernoulli",), 15, 1e-3, 15, 1e-3, 8),
    "ref_wide": (wl.mlp_arch([64, 8, 4, 1], "dense", "relu", "sigmoid"), ("bernoulli",), 37, 1e-3, 10, 1e-3, 8),
    "ref_2out": (wl.mlp_arch([3, 6, 6, 2], "denseGaussian", "relu"), ("gaussian", 0.2), 17, 1e-3, 20, 5e-4, 10),
    # step sizes large enough that both proposals are rejected (the MH select and the dual-averaging branch acc < 1)
    "ref_rej": (wl.mlp_arch([7, 5, 4, 1], "dense", "relu", "sigmoid"), ("bernoulli",), 23, 0.35, 12, 0.3, 10),
}


def net_case(name):
    arch, lik, N, eps, L, heps, hL = NET_CASES[name]
    rng = np.random.default_rng(sum(map(ord, name)))
    D = arch[0][1]
    out = [l for l in arch if l[0] in ("dense", "denseGaussian")][-1][2]
    if name in ("ref_c1a", "ref_c1b"):
        c = wl.c1(name[-1])
        X, Y = c["X"], np.asarray(c["Y"]).reshape(-1, 1)                  # Examples/trainRegression.py:33-36
    else:
        X = rng.random((N, D)) if D > 32 else rng.normal(size=(N, D))
        Y = (rng.random((N, out)) > 0.5).astype(np.float64) if lik[0] == "bernoulli" else rng.normal(size=(N, out))
    P = sum(int(np.prod(s)) for s in wl.theta_shapes(arch))
    theta = wl.init_theta(arch, seed=7) * (0.3 if D > 32 else 0.7) + 0.05 * rng.normal(size=P)
    hyper = wl.init_hyper(arch, lik)
    hyper = hyper + 0.05 * rng.normal(size=hyper.size)
    if lik[0] == "gaussian":          # train() appends tf.cast([[sd**0.5]], dtype): python floats pass through float32
        hyper[-1] = float(np.float32(hyper[-1]))
    burnin, iters = 1000, 1
    Yref = Y if out > 1 else Y.reshape(-1)                                # the examples pass 1-D targets
    net, likelihood = build_reference_network(arch, lik, X, Yref, theta, hyper)
    with quiet():
        net.setupMCMC(eps, eps / 2, eps * 2, 4, L, L, L + 10, 1, heps, hL, burnin, 4, 10, 4, 0.1, 5, 10)
    forward = net.predict(True, net.states)                               # network.py:141-171
    # the step sizes the sampler really uses: tf.cast(python float, dtype) rounds through float32 (network.py:237,253)
    eps, heps = float(np.asarray(r(net.step_size))[0]), float(np.asarray(r(net.hyper_step_size))[0])
    tfp.mcmc.TRACE, tfp.mcmc.RESULTS = [], []
    tmp = tempfile.mkdtemp()
    cwd = os.getcwd()
    os.chdir(tmp)
    try:
        with Draws(1234 + len(name)) as draws, quiet():
            net.train(iters, 1, likelihood, metricList=[], adjustHypers=True, folderName="run",
                      networksPerFile=10, displaySkip=10 ** 9)           # network.py:509-670, unmodified
    finally:
        os.chdir(cwd)
        shutil.rmtree(tmp)
    trace, results = tfp.mcmc.TRACE, tfp.mcmc.RESULTS
    tfp.mcmc.TRACE = tfp.mcmc.RESULTS = None
    ns, nh = len(net.states), len(net.hyperStates)
    assert len(trace) == (L + 1) + (hL + 1) and len(results) == 2
    flat = lambda parts: np.concatenate([np.asarray(r(p)) for p in parts])
    main_res, hyp_res = results
    # order of draws: main momentum parts (ns), main uniform, hyper momentum parts (nh), hyper uniform
    mom = flat(draws.normals[:ns])
    hmom = flat(draws.normals[ns:ns + nh])
    u_main, u_hyp = float(draws.uniforms[0]), float(draws.uniforms[1])
    theta_new = flat(net.states)
    hyper_after_main = flat(trace[L + 1]["state"])                        # hyper chain starts from the given hypers
    assert np.allclose(hyper_after_main, hyper if lik[0] != "gaussian" else hyper)
    d = {
        "name": name, "source": "reference source run unmodified over oracle/tfshim (make_ref_golden.py)",
        "arch": [list(l) for l in arch], "lik": list(lik), "N": int(X.shape[0]), "D": int(D), "out": int(out),
        "X": r(X), "Y": r(Y), "theta": r(theta), "hyper": r(hyper), "momentum": r(mom), "eps": eps, "L": L,
        "forward": r(forward),
        "logp": float(trace[0]["value"]), "grad": r(flat(trace[0]["grads"])),
        "hyper_logp": float(trace[L + 1]["value"]), "hyper_grad": r(flat(trace[L + 1]["grads"])),
        "traj_theta": r(flat(main_res.proposed_state)), "traj_momentum": r(flat(main_res.proposed_results.final_momentum)),
        "traj_logp": float(main_res.proposed_results.target_log_prob),
        "log_accept_ratio": float(main_res.log_accept_ratio),
        # --- the rest of the epoch (network.py:414-471)
        "u_main": u_main, "main_accepted": bool(main_res.is_accepted), "theta_after": r(theta_new),
        "main_accept_prob": float(np.asarray(r(net.mainAccept))[0]),
        "hyper_eps": heps, "hyper_L": hL, "hyper_momentum": r(hmom), "u_hyper": u_hyp, "epoch": 0, "burnin": burnin,
        "hyper_theta": r(theta_new),                                      # the hyper target sees the post-MH theta
        "hyper_traj": r(flat(hyp_res.proposed_state)), "hyper_traj_logp": float(hyp_res.proposed_results.target_log_prob),
        "hyper_log_accept_ratio": float(hyp_res.log_accept_ratio), "hyper_accepted": bool(hyp_res.is_accepted),
        "hyper_after": r(flat(net.hyperStates)), "hyper_accept_prob": float(np.asarray(r(net.hyperAccept))[0]),
        "hyper_step_after": float(np.asarray(r(net.hyper_step_size))[0]),
        "log_eps_bar_after": float(np.asarray(r(net.logEpsilonBar))[0]), "h_after": float(np.asarray(r(net.h))[0]),
        # every 5th interior evaluation of the main trajectory (per-step parity of the target along the path)
        "path_logp": [float(trace[i]["value"]) for i in range(1, L + 1)],
    }
    return d


# ---------------------------------------------------------------------------
# per-function fixtures
# ---------------------------------------------------------------------------
def grad_of(fn, tensors):
    leaves = [torch.tensor(np.asarray(t, dtype=np.float64), requires_grad=True) for t in tensors]
    val = fn(*[tf._wrap(v) for v in leaves])
    val = val if isinstance(val, torch.Tensor) else torch.as_tensor(val)
    grads = torch.autograd.grad(val.sum(), leaves, allow_unused=True)
    return val.detach(), [np.zeros(l.shape) if g is None else g.numpy() for g, l in zip(grads, leaves)]


def function_cases():
    rng = np.random.default_rng(99)
    out = {"source": "reference source run unmodified over oracle/tfshim (make_ref_golden.py)"}
    # BNN_functions.py:7-34 multivariateLogProb(sigma, mu, x): scalar sigma (k = 1, Q2) and full-shape sigma
    mv = []
    x = rng.normal(size=(5, 3))
    for sigma, mu in ((0.37, 0.2), (1e-9, 0.0), (3e8, -0.5), (np.full((5, 3), 0.81), rng.normal(size=(5, 3))),
                      (np.abs(rng.normal(size=(5, 3))) + 0.1, 0.3)):
        val, (gs, gm, gx) = grad_of(lambda s, m, xx: R_fn.multivariateLogProb(s, m, xx, dtype=F64), (sigma, mu, x))
        mv.append({"sigma": r(sigma), "sigma_shape": list(np.shape(sigma)), "mu": r(mu), "mu_shape": list(np.shape(mu)),
                   "x": r(x), "x_shape": [5, 3], "value": float(val), "d_sigma": r(gs), "d_mu": r(gm), "d_x": r(gx)})
    out["multivariateLogProb"] = mv
    # BNN_functions.py:37-57 cauchyLogProb(gamma, x0, x): elementwise
    ca = []
    for gamma, x0 in ((0.5, 0.0), (0.71, -0.2), (2.5, 1.0)):
        val, (gg, g0, gx) = grad_of(lambda g, m, xx: R_fn.cauchyLogProb(g, m, xx, dtype=F64), (gamma, x0, x))
        ca.append({"gamma": gamma, "x0": x0, "x": r(x), "value": r(val), "sum": float(val.sum()),
                   "d_gamma": r(gg), "d_x0": r(g0), "d_x": r(gx)})
    out["cauchyLogProb"] = ca
    # layers: calculateProbs / calculateHyperProbs / predict
    lay = []
    A = rng.normal(size=(4, 6))
    for kind, cls in (("dense", R_layer.DenseLayer), ("denseGaussian", R_layer.GaussianDenseLayer)):
        W, b = rng.normal(size=(3, 4)) * 0.6, rng.normal(size=(3, 1)) * 0.4
        hy = np.array([0.05, 0.8, -0.03, 0.6]) + (0.0 if kind == "dense" else 0.2)
        L = cls(4, 3, weights=tf.cast(W, F64), biases=tf.cast(b, F64), dtype=np.float64)
        sp = lambda h0, h1, h2, h3, w, bb: L.calculateProbs([h0, h1, h2, h3], [w, bb])
        sh = lambda h0, h1, h2, h3, w, bb: L.calculateHyperProbs([h0, h1, h2, h3], [w, bb])
        args = ([hy[0]], [hy[1]], [hy[2]], [hy[3]], W, b)
        pv, pg = grad_of(sp, args)
        hv, hg = grad_of(sh, args)
        pred = L.predict(tf.cast(A, F64), [tf.cast(W, F64), tf.cast(b, F64)])
        init = cls(4, 3, dtype=np.float64, seed=5)
        lay.append({"kind": kind, "W": r(W), "b": r(b), "hyper": r(hy), "A": r(A),
                    "prior": float(pv), "prior_d_hyper": r(np.concatenate([g.reshape(-1) for g in pg[:4]])),
                    "prior_d_W": r(pg[4]), "prior_d_b": r(pg[5]),
                    "hyper_prob": float(hv), "hyper_prob_d_hyper": r(np.concatenate([g.reshape(-1) for g in hg[:4]])),
                    "predict": r(pred), "initial_hypers": r(init.hypers),
                    "name": L.name, "numTensors": L.numTensors, "numHyperTensors": L.numHyperTensors})
    out["dense_layers"] = lay
    # activations with state (activationFunctions.py:117-433).  calculateProbs(slopes) uses the layer's own
    # self.hypers (Q4); updateHypers sets them, so the one-argument form pins the body for any hyper value.
    acts = []
    Z = rng.normal(size=(5, 7))
    for kind, cls in (("prelu", R_act.Prelu), ("squareprelu", R_act.SquarePrelu)):
        slopes = rng.normal(size=5) * 0.3 + 0.2
        L = cls(5, dtype=np.float64, alpha=0.25)
        default_params, default_hypers = r(L.parameters[0]), r(tf.convert_to_tensor(L.hypers))
        hy = np.array([0.45]) if kind == "prelu" else np.array([0.07, 0.33])
        L.updateHypers([tf.cast(v, F64) for v in hy])
        pv, (pgs,) = grad_of(lambda s: L.calculateProbs(s), (slopes,))
        if kind == "prelu":
            hv, hg = grad_of(lambda h0, s: L.calculateHyperProbs([h0], [s]), (hy[0], slopes))
        else:
            hv, hg = grad_of(lambda h0, h1, s: L.calculateHyperProbs([h0, h1], [s]), (hy[0], hy[1], slopes))
        pred = L.predict(tf.cast(Z, F64), [tf.cast(slopes, F64)])
        acts.append({"kind": kind, "slopes": r(slopes), "hyper": r(hy), "Z": r(Z), "prior": float(pv), "prior_d_slopes": r(pgs),
                     "hyper_prob": float(hv), "hyper_prob_d_hyper": r(np.concatenate([g.reshape(-1) for g in hg[:-1]])),
                     "predict": r(pred), "default_parameters": default_params, "default_hypers": default_hypers,
                     "name": L.name, "numTensors": L.numTensors, "numHyperTensors": L.numHyperTensors})
    out["param_activations"] = acts
    plain = []
    for name, obj in (("relu", R_act.Relu()), ("tanh", R_act.Tanh()), ("sigmoid", R_act.Sigmoid()), ("Exp", R_act.Exp()),
                      ("elu", R_act.Elu()), ("leakyrelu", R_act.Leaky_relu(alpha=0.17))):
        plain.append({"kind": name, "name": obj.name, "Z": r(Z), "predict": r(obj.predict(tf.cast(Z, F64), [])),
                      "alpha": 0.17 if name == "leakyrelu" else None})
    out["plain_activations"] = plain
    # likelihoods through a stub predict (likelihood.py:69-96,143-169,210-237), incl. the Bernoulli clip edges
    liks = []
    f = rng.normal(size=(2, 9))
    y = rng.normal(size=(9, 2))
    for kind, obj, hs in (("gaussian", R_lik.GaussianLikelihood(sd=0.3), [tf.cast([0.55], F64)]),
                          ("fixed", R_lik.FixedGaussianLikelihood(sd=0.3), [])):
        def fn(ff, hh):
            return obj.makeResponseLikelihood([None], predict=lambda train, st: ff, dtype=F64,
                                              hyperStates=[hh], realVals=tf.cast(y, F64), sd=None)
        val, (gf, gh) = grad_of(fn, (f, [0.55]))
        liks.append({"kind": kind, "sd": 0.3, "f": r(f), "f_shape": [2, 9], "y": r(y), "hyper_last": 0.55,
                     "value": float(val.sum()), "d_f": r(gf), "d_hyper": r(gh), "mainProbsInHypers": bool(obj.mainProbsInHypers),
                     "hypers": [float(v) for v in np.asarray(obj.hypers, dtype=np.float64).reshape(-1)]})
    p = np.array([[0.0, 1e-9, 1e-8, 0.3, 0.5, 0.9999, 1 - 1e-7, 1 - 1e-8, 1.0]])
    yb = np.array([1.0, 0.0, 1.0, 1.0, 0.0, 1.0, 0.0, 1.0, 0.0])
    for dt, npdt in ((F64, np.float64), (tf.float32, np.float32)):
        obj = R_lik.BernoulliLikelihood()
        leaves = torch.tensor(p.astype(npdt), requires_grad=True)
        val = obj.makeResponseLikelihood([None], predict=lambda train, st: tf._wrap(leaves), dtype=dt,
                                         realVals=tf.cast(yb.astype(npdt), dt))
        (g,) = torch.autograd.grad(val.sum(), leaves)
        liks.append({"kind": "bernoulli", "dtype": "float64" if dt == F64 else "float32", "f": r(p), "f_shape": [1, 9], "y": r(yb),
                     "values": r(val), "value": float(val.sum()), "d_f": r(g)})
    out["likelihoods"] = liks
    # metrics.py:30-135
    mets = []
    pt, pv_ = rng.random(size=(1, 12)), rng.random(size=(1, 7))
    yt, yv = (rng.random(size=12) > 0.5).astype(np.float64), (rng.random(size=7) > 0.5).astype(np.float64) + 0.5
    for name, cls in (("SquaredError", R_met.SquaredError), ("PercentError", R_met.PercentError), ("Accuracy", R_met.Accuracy)):
        for scaleExp, mean, sd in ((False, 0.0, 1.0), (True, 0.3, 1.7)):
            m = cls(scaleExp=scaleExp, mean=mean, sd=sd)
            m.calculate(tf.cast(pt, F64), tf.cast(pv_, F64), tf.cast(yt + 0.5, F64), tf.cast(yv, F64))
            vals = {k: float(np.asarray(v.numpy() if hasattr(v, "numpy") else v)) for k, v in vars(m).items()
                    if k not in ("scaleExp", "mean", "sd")}
            mets.append({"metric": name, "scaleExp": scaleExp, "mean": mean, "sd": sd, "predTrain": r(pt), "predVal": r(pv_),
                         "realTrain": r(yt + 0.5), "realVal": r(yv), "values": vals})
    out["metrics"] = mets
    return out


# ---------------------------------------------------------------------------
# paramAdapter
# ---------------------------------------------------------------------------
def adapter_case():
    """330 updates of the reference adapter on a synthetic chain whose jump size depends on (e, L):
    random proposals until i//m >= randomSteps, then grid searches; K, inverse, rootbeta recorded at each refit."""
    seed = 17
    pyrandom.seed(seed)                                   # random.choice at paramAdapter.py:283-284
    draws = Draws(5)
    args = dict(e1=1e-3, L1=30, el=1e-4, eu=1e-2, eNumber=12, Ll=10, Lu=60, lStep=5, m=4, k=25, a=4, delta=0.1,
                randomSteps=6)
    with draws, quiet():
        ad = R_pa.paramAdapter(args["e1"], args["L1"], args["el"], args["eu"], args["eNumber"], args["Ll"], args["Lu"],
                               args["lStep"], args["m"], args["k"], a=args["a"], delta=args["delta"],
                               randomSteps=args["randomSteps"])
        rng = np.random.default_rng(3)
        state = [np.zeros((3, 2), dtype=np.float32), np.zeros((3, 1), dtype=np.float32)]
        hist, sjds, refits = [], [], []
        n_gamma = 0
        for step in range(330):
            e, L = float(np.asarray(r(ad.currentE))[0]), float(np.asarray(r(ad.currentL))[0])
            # jump size peaks at e = 4e-3, L = 35
            scale = np.exp(-((e - 4e-3) / 3e-3) ** 2 - ((L - 35.0) / 20.0) ** 2)
            jump = [np.float32(scale) * rng.normal(size=s.shape).astype(np.float32) for s in state]
            prev = [s.copy() for s in state]
            state = [s + j for s, j in zip(state, jump)]
            sjds.append(float(sum(np.sum(np.square((n - o).reshape(-1).astype(np.float32))) for n, o in zip(state, prev))))
            E, Lr = ad.update([tf.cast(s, tf.float32) for s in state])
            hist.append([float(np.asarray(r(E))[0]), int(np.asarray(r(Lr))[0])])
            if len(ad.allData) != n_gamma or (len(ad.previousGamma) == 49 and hasattr(ad, "inverse")
                                              and len(refits) and refits[-1]["i"] != float(np.asarray(r(ad.i))[0]) - 1
                                              and (float(np.asarray(r(ad.i))[0]) - 1) % args["m"] == 0 and False):
                n_gamma = len(ad.allData)
            if hasattr(ad, "inverse") and (not refits or refits[-1]["step"] != step) and getattr(ad, "_seen", None) is not ad.inverse:
                ad._seen = ad.inverse
                refits.append({"step": step, "i": float(np.asarray(r(ad.i))[0]) - 1, "size": int(ad.inverse.shape[0]),
                               "inverse": r(ad.inverse), "inverseR": r(ad.inverseR), "s": float(np.asarray(r(ad.s))[0]),
                               "p": float(ad.p), "rootbeta": float(np.asarray(r(ad.rootbeta))[0]),
                               "E": hist[-1][0], "L": hist[-1][1]})
    return {"source": "reference paramAdapter run unmodified over oracle/tfshim", "args": args, "python_random_seed": seed,
            "uniforms": [float(u) for u in draws.uniforms], "sjd_unscaled": sjds, "history": hist, "refits": refits,
            "state_shapes": [[3, 2], [3, 1]], "jump_seed": 3}


# ---------------------------------------------------------------------------
# sample writer / reader
# ---------------------------------------------------------------------------
def run_case():
    """network.train writer + predictor reader on the 2 / 3 / 2 schedule: epochs = burnin + 2*3*2 + 1 = 15
    (the lagging summary then counts 6 networks in 2 files, Q9).  The directory the reference wrote is committed
    under tests/golden/ref_run/; predictions the reference's predictor makes from it are in ref_run.json."""
    arch = wl.mlp_arch([2, 4, 3, 1], "dense", "tanh")
    lik = ("gaussian", 0.1)
    rng = np.random.default_rng(8)
    X, Y = rng.normal(size=(9, 2)), rng.normal(size=9)
    P = sum(int(np.prod(s)) for s in wl.theta_shapes(arch))
    theta = wl.init_theta(arch, seed=2) * 0.5
    hyper = wl.init_hyper(arch, lik)
    net, likelihood = build_reference_network(arch, lik, X.astype(np.float32), Y.astype(np.float32), theta, hyper, npdt=np.float32)
    dst = os.path.join(HERE, "ref_run")
    shutil.rmtree(dst, ignore_errors=True)
    tmp = tempfile.mkdtemp()
    cwd = os.getcwd()
    os.chdir(tmp)
    try:
        with Draws(77), quiet():
            pyrandom.seed(4)
            net.setupMCMC(5e-3, 1e-3, 1e-2, 5, 6, 4, 10, 2, 1e-3, 5, 2, 4, 2, 4, 0.1, 5, 3)
            net.train(15, 2, likelihood, folderName="ref_run", networksPerFile=3, displaySkip=10 ** 9)
        shutil.copytree(os.path.join(tmp, "ref_run"), dst)
    finally:
        os.chdir(cwd)
        shutil.rmtree(tmp)
    files = sorted(os.listdir(dst))
    with quiet():
        pred = R_pred.predictor(dst + "/", tf.float32)
    Xt = rng.normal(size=(5, 2)).astype(np.float32)
    outs = pred.predict(Xt, n=1)
    outs2 = pred.predict(Xt, n=2)
    acf = pred.autocorrelation(Xt, 4)                   # predictor.py:275-292 over the emcee stand-in
    acl = pred.autoCorrelationLength(Xt, 4)             # :294-312
    return {"source": "reference network.train writer + predictor reader run unmodified over oracle/tfshim",
            "arch": [list(l) for l in arch], "lik": list(lik), "files": files,
            "line_counts": {f: sum(1 for _ in open(os.path.join(dst, f))) for f in files},
            "summary": open(os.path.join(dst, "summary.txt")).read(), "architecture": open(os.path.join(dst, "architecture.txt")).read(),
            "schedule": {"epochs": 15, "burnin": 2, "samplingStep": 2, "networksPerFile": 3},
            "numNetworks": int(pred.numNetworks), "matrix_shapes": [list(m.shape) for m in pred.matrices],
            "hypers": [r(h) for h in pred.hypers], "Xtest": r(Xt), "predict_n1": [r(o) for o in outs],
            "predict_n2": [r(o) for o in outs2],
            "autocorrelation_nmax4": r(acf), "autocorrelation_length": float(acl), "final_states": [r(s) for s in net.states],
            "final_hypers": r(np.concatenate([np.asarray(r(h)) for h in net.hyperStates]))}


def main():
    for name in NET_CASES:
        d = net_case(name)
        with open(os.path.join(HERE, name + ".json"), "w") as f:
            json.dump(d, f)
        print(name, "P=%d" % len(d["theta"]), "logp=%.12g" % d["logp"], "lar=%.6g" % d["log_accept_ratio"],
              "accepted", d["main_accepted"], "hyper lar=%.6g" % d["hyper_log_accept_ratio"], d["hyper_accepted"])
    for fname, fn in (("reffn_functions.json", function_cases), ("reffn_adapter.json", adapter_case), ("reffn_run.json", run_case)):
        d = fn()
        with open(os.path.join(HERE, fname), "w") as f:
            json.dump(d, f)
        print(fname, "ok")


if __name__ == "__main__":
    main()
