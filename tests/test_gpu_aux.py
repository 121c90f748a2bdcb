"""GPU parity of the remaining hot-path rows: Philox momentum stream, adapter grid search
(paramAdapter.gridSearch), posterior-predictive sweep (predictor.predict)."""
import math
import os
import random

import numpy as np
import pytest
import torch

from oracle import adapter as oad
from oracle import targets
from tensorbnn_b200 import workloads as wl
import philox_ref

pytestmark = pytest.mark.gpu


def _engine(arch, lik, dtype, chains=1):
    from tensorbnn_b200.engine import Engine
    return Engine(arch, lik, dtype=dtype, chains=chains)


# ---------------------------------------------------------------------------- RNG
@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_momentum_stream_matches_reference_philox(dtype):
    arch, lik = wl.mlp_arch([3, 6, 5, 1], "dense", "squareprelu"), ("gaussian", 0.1)
    eng = _engine(arch, lik, dtype, chains=3)
    eng.set_data(np.zeros((2, 3)), np.zeros(2))
    seed, call = 0x1234567890ABCDEF, 77
    p, ke = eng.draw_momentum(seed, call)
    p = p.cpu().numpy()
    gen = philox_ref.normals_f32 if dtype == torch.float32 else philox_ref.normals_f64
    tol = 2e-5 if dtype == torch.float32 else 1e-12
    for c in range(3):
        ref = gen(seed, philox_ref.STREAM_MAIN, call, c, eng.P)
        assert np.abs(p[c] - ref).max() <= tol * max(1.0, np.abs(ref).max())
        assert abs(ke[c].item() - 0.5 * np.sum(p[c].astype(np.float64) ** 2)) <= 1e-5 * ke[c].item()
    # different call counter / chain => different stream
    p2, _ = eng.draw_momentum(seed, call + 1)
    assert not np.allclose(p2.cpu().numpy(), p)


def test_momentum_is_standard_normal():
    arch, lik = wl.mlp_arch([784, 20, 20, 1], "dense", "relu", "sigmoid"), ("bernoulli",)
    eng = _engine(arch, lik, torch.float32, chains=8)
    eng.set_data(np.zeros((2, 784)), np.zeros(2))
    p, _ = eng.draw_momentum(42, 0)
    x = p.cpu().numpy().ravel().astype(np.float64)
    n = x.size
    assert abs(x.mean()) < 5 / math.sqrt(n)
    assert abs(x.var() - 1) < 5 * math.sqrt(2 / n)
    assert abs(np.mean(x ** 3)) < 5 * math.sqrt(15 / n)
    assert abs(np.mean(x ** 4) - 3) < 5 * math.sqrt(96 / n)


# ---------------------------------------------------------------------------- adapter
def _adapter_state(n_hist, seed, eN=40, Ll=100, Lu=2000, lStep=7):
    rng = random.Random(seed)
    ad = oad.OracleAdapter(1e-3, 500, 1e-4, 1e-2, eN, Ll, Lu, lStep, 10, 100, randomSteps=0, rng=rng)
    nrng = np.random.default_rng(seed)
    # drive the adapter with synthetic states until it has n_hist history points
    st = [nrng.normal(size=(5, 3)), nrng.normal(size=(5, 1))]
    guard = 0
    while len(ad.previousGamma) < n_hist and guard < 5000:
        st = [s + 0.01 * nrng.normal(size=s.shape) * (1 + 0.3 * math.sin(guard / 7.0)) for s in st]
        ad.update(st)
        guard += 1
    return ad


def _conditioned_state(n_hist, seed, eN=40, Ll=100, Lu=2000, lStep=7):
    """A well-conditioned GP state built directly: distinct history points, unit noise."""
    ad = oad.OracleAdapter(1e-3, 500, 1e-4, 1e-2, eN, Ll, Lu, lStep, 10, 100, randomSteps=0)
    rng = np.random.default_rng(seed)
    pts = set()
    while len(pts) < n_hist:
        pts.add((float(ad.eGrid[rng.integers(eN)]), float(ad.lGrid[rng.integers(ad.lNumber)])))
    ad.previousGamma = sorted(pts)
    K = np.array([[ad.calck(a, b, ad.el, ad.eu, ad.sigma) for b in ad.previousGamma]
                  for a in ad.previousGamma], dtype=np.float64)
    data = rng.random(n_hist) + 0.5
    inv = np.linalg.inv(K + np.eye(n_hist) * max(1.0, 0.05 * np.abs(K).max()))
    ad.inverse = inv.astype(np.float32)
    ad.inverseR = (inv @ data[:, None]).astype(np.float32)
    ad.s = np.float32(4.0 / data.max())
    ad.p = 0.7
    ad.rootbeta = 2.5
    return ad


@pytest.mark.parametrize("n_hist,seed,driven", [(1, 0, True), (3, 1, True), (12, 2, True), (5, 4, False),
                                                (30, 3, False), (50, 5, False)])
def test_adapter_grid_search(n_hist, seed, driven):
    from tensorbnn_b200.engine import adapter_ucb
    ad = _adapter_state(n_hist, seed) if driven else _conditioned_state(n_hist, seed)
    args = (ad.previousGamma, ad.inverseR, ad.s, ad.inverse, ad.p, ad.rootbeta, ad.el, ad.eu, ad.sigma)
    e_ref, L_ref = ad.gridSearch(*args)
    e, L, ucb = adapter_ucb(0, ad.eGrid, ad.lGrid, np.array(ad.previousGamma, dtype=np.float32),
                            ad.inverse, ad.inverseR[:, 0], ad.s, ad.p, ad.rootbeta, ad.el, ad.eu,
                            ad.Ll, ad.Lu, ad.sigma)
    surf = ad.ucb_surface(*args)                     # float64 surface, [lNumber, eNumber]
    li = int(np.argmin(np.abs(ad.lGrid - L)))
    ei = int(np.argmin(np.abs(ad.eGrid - e)))
    assert ad.lGrid[li] == np.float32(L) and ad.eGrid[ei] == np.float32(e)   # a grid point
    # the device choice maximises the surface up to float32 evaluation error of the UCB values
    scale = max(1.0, float(np.abs(surf).max()))
    assert surf[li, ei] >= surf.max() - 1e-4 * scale
    assert abs(ucb - surf[li, ei]) <= 1e-4 * scale
    lr = int(np.argmin(np.abs(ad.lGrid - L_ref)))
    er = int(np.argmin(np.abs(ad.eGrid - e_ref)))
    assert abs(surf[li, ei] - surf[lr, er]) <= 1e-4 * scale       # only near-ties may differ


def test_adapter_first_maximum_tie_break():
    """A flat surface (one history point at the grid centre with zero data) must return the FIRST
    grid point in scan order (e fastest, L slowest), paramAdapter.py:184-190."""
    from tensorbnn_b200.engine import adapter_ucb
    eGrid = np.linspace(1e-4, 1e-2, 8).astype(np.float32)
    lGrid = np.arange(100, 200, 10).astype(np.float32)
    prev = np.array([[5.05e-3, 150.0]], dtype=np.float32)
    # rootbeta = 0 and KinvR = 0 make the UCB identically zero
    e, L, ucb = adapter_ucb(0, eGrid, lGrid, prev, np.ones((1, 1)), np.zeros(1), 1.0, 1.0, 0.0,
                            1e-4, 1e-2, 100.0, 190.0, np.diag([6.25, 6.25]))
    assert e == float(eGrid[0]) and L == float(lGrid[0]) and ucb == 0.0


def _engine_flags(arch, lik, dtype, flags):
    from tensorbnn_b200.engine import Engine
    return Engine(arch, lik, dtype=dtype, flags=flags)


# ---------------------------------------------------------------------------- predictor
@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
@pytest.mark.parametrize("arch", [wl.mlp_arch([1, 64, 64, 64, 1], "dense", "squareprelu"),
                                  wl.mlp_arch([5, 9, 3], "denseGaussian", "tanh"),
                                  wl.mlp_arch([784, 20, 20, 1], "dense", "relu", "sigmoid")])
def test_predict_matches_oracle(arch, dtype):
    lik = ("fixed", 0.1)
    rng = np.random.default_rng(0)
    th0 = wl.init_theta(arch, seed=3)
    S, M = 7, 333
    samples = th0[None, :] + 0.05 * rng.normal(size=(S, th0.size))
    D = arch[0][1]
    X = rng.normal(size=(M, D))
    eng = _engine(arch, lik, dtype)
    out, mom = eng.predict(samples, X, want_out=True, want_moments=True)
    out, mom = out.cpu().numpy(), mom.cpu().numpy()
    np_dt = np.float32 if dtype == torch.float32 else np.float64
    r = lambda a: torch.tensor(np.asarray(a).astype(np_dt).astype(np.float64))
    ref = np.stack([targets.forward(arch, targets.unflatten_theta(arch, r(samples[s])), r(X)).numpy()
                    for s in range(S)])
    tol = 2e-5 if dtype == torch.float32 else 1e-11
    assert np.abs(out - ref).max() <= tol * max(1.0, np.abs(ref).max())
    assert np.all(mom[0] == S)
    assert np.abs(mom[1] - ref.mean(axis=0)).max() <= 5 * tol * max(1.0, np.abs(ref).max())
    var_ref = ref.var(axis=0)
    assert np.abs(mom[2] / S - var_ref).max() <= 1e-4 * max(1e-12, var_ref.max()) + 10 * tol


def test_predict_moments_only_many_samples():
    """Fused mean/sd mode over more samples than one weight-staging chunk of rows."""
    arch, lik = wl.mlp_arch([1, 16, 16, 1], "dense", "squareprelu"), ("gaussian", 0.1)
    rng = np.random.default_rng(1)
    th0 = wl.init_theta(arch, seed=4)
    S, M = 200, 5000
    samples = th0[None, :] + 0.05 * rng.normal(size=(S, th0.size))
    X = np.linspace(-4, 4, M)[:, None]
    eng = _engine(arch, lik, torch.float32)
    _, mom = eng.predict(samples, X, want_out=False, want_moments=True)
    out, _ = eng.predict(samples, X, want_out=True, want_moments=False)
    out = out.cpu().numpy().astype(np.float64)
    mom = mom.cpu().numpy()
    assert np.abs(mom[1] - out.mean(axis=0)).max() <= 1e-5 * max(1.0, np.abs(out).max())
    assert np.abs(mom[2] / S - out.var(axis=0)).max() <= 1e-4 * out.var(axis=0).max() + 1e-9


# ---------------------------------------------------------------------------- tcgen05 predictor
UMMA_ARCHS = [wl.mlp_arch([1, 64, 64, 64, 1], "dense", "squareprelu"),
              wl.mlp_arch([3, 40, 24, 2], "denseGaussian", "tanh"),
              wl.mlp_arch([2, 32, 48, 16, 1], "dense", "relu", "sigmoid"),
              [("dense", 8, 80), ("elu",), ("dense", 80, 72), ("prelu", 72), ("dense", 72, 4)]]


@pytest.mark.parametrize("arch", UMMA_ARCHS)
@pytest.mark.parametrize("S,M", [(5, 700), (3, 128), (2, 1)])
def test_predict_tensor_core_kernel_matches_oracle(arch, S, M):
    """k_predict_umma (tcgen05, 3xTF32 from TMEM) vs the fp64 oracle and vs the FFMA predictor."""
    from tensorbnn_b200 import _lib
    lik = ("fixed", 0.1)
    rng = np.random.default_rng(2)
    th0 = wl.init_theta(arch, seed=3)
    samples = th0[None, :] + 0.05 * rng.normal(size=(S, th0.size))
    D = arch[0][1]
    X = rng.normal(size=(M, D))
    eng = _engine(arch, lik, torch.float32)
    assert eng.predict_kernel() == "k_predict_umma"
    out, mom = eng.predict(samples, X, want_out=True, want_moments=True)
    ffma = _engine_flags(arch, lik, torch.float32, _lib.FLAG_NO_UMMA)
    assert ffma.predict_kernel() == "k_predict"
    out2, mom2 = ffma.predict(samples, X, want_out=True, want_moments=True)
    r = lambda a: torch.tensor(np.asarray(a).astype(np.float32).astype(np.float64))
    ref = np.stack([targets.forward(arch, targets.unflatten_theta(arch, r(samples[s])), r(X)).numpy()
                    for s in range(S)])
    out, out2, mom, mom2 = out.cpu().numpy(), out2.cpu().numpy(), mom.cpu().numpy(), mom2.cpu().numpy()
    scale = max(1.0, np.abs(ref).max())
    assert np.abs(out - ref).max() <= 2e-5 * scale
    assert np.abs(out - out2).max() <= 2e-5 * scale
    assert np.all(mom[0] == S)
    assert np.abs(mom[1] - mom2[1]).max() <= 2e-5 * scale
    assert np.abs(mom[2] - mom2[2]).max() <= 1e-4 * max(1e-12, np.abs(mom2[2]).max()) + 1e-6


def test_predict_tensor_core_moments_accumulate_across_calls():
    """The fused (count, mean, M2) mode over many samples equals the materialised outputs' moments."""
    arch = wl.mlp_arch([1, 64, 64, 64, 1], "dense", "squareprelu")
    rng = np.random.default_rng(5)
    th0 = wl.init_theta(arch, seed=9, slope=0.1 ** 0.5)
    S, M = 64, 3000
    samples = th0[None, :] + 0.05 * rng.normal(size=(S, th0.size))
    X = np.linspace(-4, 4, M)[:, None]
    eng = _engine(arch, ("gaussian", 0.1), torch.float32)
    assert eng.predict_kernel() == "k_predict_umma"
    _, mom = eng.predict(samples, X, want_out=False, want_moments=True)
    out, _ = eng.predict(samples, X, want_out=True, want_moments=False)
    out = out.cpu().numpy().astype(np.float64)
    mom = mom.cpu().numpy()
    assert np.abs(mom[1] - out.mean(axis=0)).max() <= 1e-5 * max(1.0, np.abs(out).max())
    assert np.abs(mom[2] / S - out.var(axis=0)).max() <= 1e-4 * out.var(axis=0).max() + 1e-9


# ---------------------------------------------------------------------------- the drop-in Python surface end to end
def test_train_writes_reference_layout_and_predictor_reads_it(tmp_path, monkeypatch):
    """Examples/trainRegression.py in miniature through the reference's class API: network(...).add(...),
    setupMCMC, train -> text files in the reference's layout (+ float32 side-cars) -> predictor.predict.
    The side-cars hold exactly what parsing the text gives; predictions are finite and have the documented shape."""
    from tensorbnn_b200.activationFunctions import Tanh
    from tensorbnn_b200.layer import GaussianDenseLayer
    from tensorbnn_b200.likelihood import FixedGaussianLikelihood
    from tensorbnn_b200.metrics import SquaredError
    from tensorbnn_b200.network import network
    from tensorbnn_b200.predictor import predictor
    monkeypatch.chdir(tmp_path)
    x = np.linspace(-2, 2, 11)
    y = np.sin(x * math.pi * 2) * x - np.cos(x * math.pi)
    xv = np.linspace(-2 + 2 / 30, 2 - 2 / 30, 30)
    yv = np.sin(xv * math.pi * 2) * xv - np.cos(xv * math.pi)
    net = network(np.float32, 1, x, y, xv, yv)
    for i, (a, b) in enumerate(((1, 10), (10, 10), (10, 1))):
        net.add(GaussianDenseLayer(a, b, seed=10 + i))
        if b != 1:
            net.add(Tanh())
    net.setupMCMC(0.001, 0.0001, 0.01, 20, 50, 10, 100, 10, 0.001, 10, 20, 2, 4)      # burnin = 20
    net.train(20 + 41, 2, FixedGaussianLikelihood(sd=0.1), metricList=[SquaredError(mean=0, sd=1, scaleExp=False)],
              folderName="run", networksPerFile=10, displaySkip=1000, verbose=False)
    folder = str(tmp_path / "run") + "/"
    names = sorted(os.listdir(folder))
    assert "summary.txt" in names and "architecture.txt" in names and "0.0.txt" in names and "0.0.f32" in names
    pr = predictor(folder, np.float32)
    assert pr.numNetworks in (10, 20) and pr.numMatrices == 6     # lagging summary (Q9)
    # side-car == parsed text
    w_txt = np.loadtxt(folder + "0.0.txt", dtype=np.float32, ndmin=2)
    w_bin = np.fromfile(folder + "0.0.f32", dtype="<f4")
    assert np.array_equal(w_bin[:w_txt.size], w_txt.reshape(-1))
    out = pr.predict(xv[:, None], n=5)
    assert len(out) == pr.numNetworks // 5 and out[0].shape == (1, 30)
    assert all(np.isfinite(o).all() for o in out)
