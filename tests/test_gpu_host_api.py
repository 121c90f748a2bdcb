"""GPU tests of the host-side mirror of the reference API beyond the sampling kernels: predictor.reweight /
trainProbs against the oracle restatement, predictor.autocorrelation (f4) against the reference's own
code run over the emcee stand-in (tests/golden/reffn_run.json), non-syncing metrics (f2), and the round-1 advisor
findings (leaky-relu slope persisted, stale engine after calculateProbs, chains > 1 display path, fresh random
numbers on a second train() call, engine on a non-current device)."""
import contextlib
import io
import json
import os

import numpy as np
import pytest
import torch

from oracle import fileformat, reweight as oreweight, targets
from tensorbnn_b200 import workloads as wl

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


def _quiet():
    return contextlib.redirect_stdout(io.StringIO())


def _ref_run():
    d = json.load(open(os.path.join(GOLD, "reffn_run.json")))
    return d, os.path.join(GOLD, "ref_run") + "/"


def test_autocorrelation_matches_reference_code():
    """predictor.autocorrelation / autoCorrelationLength on the directory the reference's train() wrote: same numbers
    as the reference's predictor (its own code over the emcee.autocorr stand-in)."""
    from tensorbnn_b200.predictor import predictor
    d, path = _ref_run()
    pred = predictor(path, torch.float32)
    Xt = np.asarray(d["Xtest"], dtype=np.float32).reshape(5, 2)
    acf = pred.autocorrelation(Xt, 4)
    assert len(acf) == 4 and rel(acf, d["autocorrelation_nmax4"]) <= 1e-5
    with _quiet():
        acl = pred.autoCorrelationLength(Xt, 4)
    assert abs(acl - d["autocorrelation_length"]) <= 1e-5 * abs(d["autocorrelation_length"])


def test_autocorrelation_of_an_ar1_trace():
    """The batched FFT estimator on traces with a known answer: AR(1) with coefficient phi has acf(k) = phi^k and
    integrated time (1 + phi) / (1 - phi)."""
    from tensorbnn_b200.predictor import predictor
    rng = np.random.default_rng(0)
    phi, T, R = 0.6, 20000, 8
    e = rng.normal(size=(R, T))
    x = np.zeros((R, T))
    for t in range(1, T):
        x[:, t] = phi * x[:, t - 1] + e[:, t]
    xt = torch.tensor(x, device="cuda")
    acf = predictor._acf(xt)
    tau = predictor._integrated_time(acf).cpu().numpy()
    assert np.abs(acf[:, 1].cpu().numpy() - phi).max() < 0.03 and np.abs(acf[:, 3].cpu().numpy() - phi ** 3).max() < 0.03
    want = (1 + phi) / (1 - phi)
    assert abs(tau.mean() - want) < 0.3 and np.abs(tau - want).max() < 0.8       # the estimator's own scatter at T = 20,000


@pytest.mark.parametrize("lik_kind", ["gaussian", "fixed", "bernoulli", None])
def test_reweight_matches_oracle(tmp_path, lik_kind):
    """Importance weights for new prior families (dense -> denseGaussian) on the reference-written run."""
    from tensorbnn_b200.likelihood import BernoulliLikelihood, FixedGaussianLikelihood, GaussianLikelihood
    from tensorbnn_b200.predictor import predictor
    d, path = _ref_run()
    rng = np.random.default_rng(5)
    X, Y = rng.normal(size=(9, 2)), rng.normal(size=9)
    lik = {"gaussian": GaussianLikelihood(sd=0.1), "fixed": FixedGaussianLikelihood(sd=0.25),
           "bernoulli": BernoulliLikelihood(), None: None}[lik_kind]
    pred = predictor(path, torch.float32, likelihood=lik if lik is not None else GaussianLikelihood(sd=0.1))
    new_arch_file = tmp_path / "arch_new.txt"
    new_arch_file.write_text("denseGaussian\ntanh\ndenseGaussian\ntanh\ndense\n")
    w = pred.reweight(str(new_arch_file), trainX=X.T, trainY=Y, n=1, likelihood=lik)      # trainX as [D, N], like the reference
    assert w.shape == (d["numNetworks"],) and abs(w.sum() - 1.0) < 1e-12 and np.all(w > 0)
    arch = [tuple(l) for l in d["arch"]]
    arch_new = [("denseGaussian", 2, 4), ("tanh",), ("denseGaussian", 4, 3), ("tanh",), ("dense", 3, 1)]
    mats, hypers = fileformat.load_networks(path)
    samples = np.concatenate([m.reshape(m.shape[0], -1) for m in mats], axis=1)
    spec = lik.spec() if lik is not None else None
    w_ref = oreweight.reweight(arch, arch_new, spec, spec, samples, np.asarray(hypers, dtype=np.float64),
                               X.astype(np.float32).astype(np.float64), Y.astype(np.float32).astype(np.float64).reshape(-1, 1))
    assert rel(w, w_ref) <= 2e-3, (w, w_ref)            # exp() of float32 log-densities of size ~1e2
    wt_ref = oreweight.neg_log_weights(arch, spec, samples, np.asarray(hypers, dtype=np.float64),
                                       X.astype(np.float32).astype(np.float64), Y.astype(np.float32).astype(np.float64).reshape(-1, 1))
    assert rel(pred.weightsTrain, wt_ref) <= 2e-5


def _small_net(chains=1, alpha=None, seed=3):
    from tensorbnn_b200.activationFunctions import Leaky_relu, Tanh
    from tensorbnn_b200.layer import DenseLayer
    from tensorbnn_b200.network import network
    rng = np.random.default_rng(1)
    X = rng.normal(size=(40, 2)).astype(np.float32)
    Y = np.sin(X[:, 0]) + 0.1 * rng.normal(size=40).astype(np.float32)
    net = network(np.float32, 2, X, Y, X[:10], Y[:10], chains=chains)
    net.add(DenseLayer(2, 5, seed=seed))
    net.add(Leaky_relu(alpha=alpha) if alpha is not None else Tanh())
    net.add(DenseLayer(5, 1, seed=seed + 2))
    with _quiet():
        net.setupMCMC(1e-3, 5e-4, 2e-3, 5, 8, 4, 12, 2, 1e-3, 5, 2, 4, 2, 4, 0.1, 5, 3)
    return net, X, Y


def test_leaky_relu_slope_survives_train_then_predict(tmp_path, monkeypatch):
    from tensorbnn_b200.likelihood import FixedGaussianLikelihood
    from tensorbnn_b200.predictor import predictor
    monkeypatch.chdir(tmp_path)
    net, X, Y = _small_net(alpha=0.11)
    net.train(9, 2, FixedGaussianLikelihood(sd=0.3), folderName="run", networksPerFile=3, verbose=False)
    assert os.path.exists(tmp_path / "run" / "layer_params.txt")
    pred = predictor(str(tmp_path / "run") + "/", torch.float32)
    assert pred._arch[1] == ("leakyrelu", 0.11)
    outs = pred.predict(X[:7])
    arch = [("dense", 2, 5), ("leakyrelu", 0.11), ("dense", 5, 1)]
    for m in range(pred.numNetworks):
        theta = [torch.tensor(np.asarray(mat[m], dtype=np.float64)) for mat in pred.matrices]
        ref = targets.forward(arch, theta, torch.tensor(X[:7].astype(np.float64))).numpy()
        assert rel(outs[m], ref) <= 2e-5
    wrong = targets.forward([("dense", 2, 5), ("leakyrelu", 0.3), ("dense", 5, 1)], theta,
                            torch.tensor(X[:7].astype(np.float64))).numpy()
    assert rel(outs[-1], wrong) > 1e-3          # the default slope would have been visibly wrong


def test_engine_is_rebuilt_when_the_likelihood_changes():
    """calculateProbs() before train() used to pin an engine without the likelihood's hyper parameter."""
    from tensorbnn_b200.likelihood import FixedGaussianLikelihood, GaussianLikelihood
    net, X, Y = _small_net()
    with pytest.raises(RuntimeError):
        net.calculateProbs()
    net.likelihood = FixedGaussianLikelihood(sd=1.0)
    lp_fixed = float(net.calculateProbs())
    assert net._engine.H == 8
    net.train(6, 1, GaussianLikelihood(sd=0.2), verbose=False)
    assert net._engine.H == 9 and len(net.hyperStates) == 9
    assert abs(float(net.hyperStates[-1]) - float(np.float32(0.2 ** 0.5))) > 1e-6       # the noise hyper was sampled
    assert np.isfinite(lp_fixed)


def test_two_chains_train_with_metrics_on_the_display_path(tmp_path, monkeypatch):
    from tensorbnn_b200.likelihood import GaussianLikelihood
    from tensorbnn_b200.metrics import SquaredError
    monkeypatch.chdir(tmp_path)
    net, X, Y = _small_net(chains=2)
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        net.train(5, 1, GaussianLikelihood(sd=0.2), metricList=[SquaredError()], folderName="run2", networksPerFile=2,
                  displaySkip=1)
    assert "training squared error" in buf.getvalue()
    assert os.path.isdir(tmp_path / "run2" / "chain0") and os.path.isdir(tmp_path / "run2" / "chain1")
    lp = net.calculateProbs()
    assert lp.shape == (2,) and torch.isfinite(lp).all()


def test_metrics_stay_on_the_device_until_displayed():
    from tensorbnn_b200.metrics import Accuracy, PercentError, SquaredError
    d = json.load(open(os.path.join(GOLD, "reffn_functions.json")))
    for c in d["metrics"]:
        cls = {"SquaredError": SquaredError, "PercentError": PercentError, "Accuracy": Accuracy}[c["metric"]]
        m = cls(scaleExp=c["scaleExp"], mean=c["mean"], sd=c["sd"])
        dev = lambda k, shape: torch.tensor(np.asarray(c[k]).reshape(shape), device="cuda", dtype=torch.float64)
        m.calculate(dev("predTrain", (1, -1)), dev("predVal", (1, -1)), dev("realTrain", (-1,)), dev("realVal", (-1,)))
        for k, want in c["values"].items():
            got = getattr(m, k)
            assert isinstance(got, torch.Tensor) and got.is_cuda          # no host read-back in calculate()
            assert abs(float(got) - want) <= 1e-12 * max(1.0, abs(want)), (c["metric"], k)


def test_second_train_call_draws_fresh_random_numbers():
    from tensorbnn_b200.likelihood import FixedGaussianLikelihood
    a, _, _ = _small_net()
    a.train(4, 1, FixedGaussianLikelihood(sd=0.3), adjustHypers=False, verbose=False)
    b, _, _ = _small_net()
    b.train(2, 1, FixedGaussianLikelihood(sd=0.3), adjustHypers=False, verbose=False)
    first = [s.clone() for s in b.states]
    b.train(2, 1, FixedGaussianLikelihood(sd=0.3), adjustHypers=False, verbose=False)
    assert b._rng_calls == 4
    for sa, sb in zip(a.states, b.states):
        assert torch.equal(sa, sb)              # 2 + 2 epochs == 4 epochs: the counter carries over
    assert any(not torch.equal(f, s) for f, s in zip(first, b.states))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_engine_on_a_non_current_device():
    from tensorbnn_b200.engine import Engine
    arch, lik = wl.mlp_arch([3, 6, 1], "dense", "tanh"), ("fixed", 0.5)
    rng = np.random.default_rng(0)
    X, Y = rng.normal(size=(30, 3)), rng.normal(size=(30, 1))
    th, hy = wl.init_theta(arch, seed=1), wl.init_hyper(arch, lik)
    torch.cuda.set_device(0)
    e0 = Engine(arch, lik, device=0)
    e1 = Engine(arch, lik, device=1)
    e0.set_data(X, Y)
    e1.set_data(X, Y)
    a = e0.logp_grad(th[None], hy[None])
    b = e1.logp_grad(th[None], hy[None])
    assert torch.cuda.current_device() == 0
    assert b[1].device.index == 1 and torch.equal(a[1].cpu(), b[1].cpu())
    z = torch.zeros(4, device="cuda")           # later torch work still lands on device 0
    assert z.device.index == 0
