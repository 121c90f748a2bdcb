"""GPU parity of the tcgen05 3xTF32 training sweep (csrc/k_train_umma.cu: the default for networks whose hidden blocks
are 64 or 128 wide -- the batched-chain C3 shape and the large-N C4 shape of BASELINE.json) against the fp64 oracle
and against the FP32 FFMA tile engine (TBNN_FLAG_NO_UMMA_TRAIN), plus full-size checks of the fp32 accumulation
(C4 at 262,144 rows, C2-L at 1,048,576 rows) against fp64.

Tolerance: 1e-5 relative on log-posterior and gradient (max |g - g_ref| <= 1e-5 max |g_ref|), BASELINE.json north_star.
One caveat is inherent to piecewise-linear activations in ANY fp32 evaluation (TF's included): a pre-activation
within rounding distance of its kink can land on the other side, which moves the gradient by one row's contribution
(~1/N of a weight's gradient).  The many-chain test therefore asserts the tolerance on the median and on >= 95 % of the
chains and bounds the rest by a few rows' worth.
"""
import numpy as np
import pytest
import torch

from oracle import analytic
from tensorbnn_b200 import _lib
from tensorbnn_b200 import workloads as wl

pytestmark = pytest.mark.gpu


def _engine(arch, lik, chains=1, flags=0, dtype=torch.float32):
    from tensorbnn_b200.engine import Engine
    return Engine(arch, lik, dtype=dtype, chains=chains, flags=flags)


def rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


r32 = lambda a: np.asarray(a).astype(np.float32).astype(np.float64)

SHAPES = {
    "relu64": (wl.mlp_arch([3, 64, 64, 1], "dense", "relu"), ("gaussian", 0.1), 300, 2),
    "c3s": (wl.mlp_arch([1, 64, 64, 64, 1], "dense", "squareprelu"), ("gaussian", 0.1), 200, 2),
    "prelu64": (wl.mlp_arch([2, 64, 64, 1], "denseGaussian", "prelu"), ("fixed", 0.2), 129, 2),
    "tanh64": (wl.mlp_arch([5, 64, 64, 64, 2], "denseGaussian", "tanh"), ("fixed", 0.3), 515, 3),
    "c4s": (wl.mlp_arch([32, 128, 128, 128, 1], "dense", "relu"), ("gaussian", 0.1), 700, 1),
    "bern128": (wl.mlp_arch([9, 128, 128, 1], "dense", "relu", "sigmoid"), ("bernoulli",), 1000, 2),
    "leaky128": ([("dense", 40, 128), ("leakyrelu", 0.17), ("dense", 128, 128), ("leakyrelu", 0.17),
                  ("dense", 128, 3)], ("gaussian", 0.4), 257, 1),
    "elu128deep": (wl.mlp_arch([7, 128, 128, 128, 128, 1], "dense", "elu"), ("gaussian", 0.2), 128, 1),
    "one_row": (wl.mlp_arch([3, 64, 64, 1], "dense", "relu"), ("gaussian", 0.1), 1, 1),
}


def _problem(name, seed=1):
    arch, lik, N, C = SHAPES[name]
    rng = np.random.default_rng(seed)
    D = arch[0][1]
    out = [l for l in arch if l[0] in ("dense", "denseGaussian")][-1][2]
    X = rng.normal(size=(N, D))
    Y = (rng.random(N) > 0.5).astype(np.float64) if lik[0] == "bernoulli" else rng.normal(size=(N, out))
    P = wl.init_theta(arch).size
    TH = np.stack([wl.init_theta(arch, seed=5 + 17 * c) * 0.7 + 0.05 * rng.normal(size=P) for c in range(C)])
    H = wl.init_hyper(arch, lik).size
    HY = np.stack([wl.init_hyper(arch, lik) + 0.05 * rng.normal(size=H) for c in range(C)])
    return arch, lik, X, Y, TH, HY


@pytest.mark.parametrize("name", sorted(SHAPES))
def test_train_umma_matches_oracle_and_ffma_engine(name):
    arch, lik, X, Y, TH, HY = _problem(name)
    C = TH.shape[0]
    eng = _engine(arch, lik, chains=C)
    eng.set_data(X, Y)
    assert eng.sweep_info()["kernel"] == "k_train_umma"
    lp, g, st = eng.logp_grad(TH, HY)
    lp2, g2, st2 = eng.logp_grad(TH, HY)
    assert torch.equal(lp, lp2) and torch.equal(g, g2) and torch.equal(st, st2)       # fixed summation order
    ref = _engine(arch, lik, chains=C, flags=_lib.FLAG_NO_UMMA_TRAIN)
    ref.set_data(X, Y)
    assert ref.sweep_info()["kernel"] == "k_partial"
    lpf, gf, stf = ref.logp_grad(TH, HY)
    lp, g, st, lpf, gf, stf = (t.cpu().numpy() for t in (lp, g, st, lpf, gf, stf))
    for c in range(C):
        lp_ref, g_ref = analytic.main_value_and_grad(arch, lik, r32(TH[c]), r32(HY[c]), r32(X), r32(Y))
        assert abs(lp[c] - lp_ref) <= 1e-5 * abs(lp_ref), (name, c)
        assert rel(g[c], g_ref) <= 1e-5, (name, c, rel(g[c], g_ref))
        assert abs(lp[c] - lpf[c]) <= 5e-6 * abs(lpf[c]) and abs(st[c] - stf[c]) <= 5e-6 * abs(stf[c])
        assert rel(g[c], gf[c]) <= 1e-5


@pytest.mark.parametrize("name", ["c3s", "relu64", "tanh64"])
def test_train_umma_two_threads_per_row_variant(name, monkeypatch):
    """64-wide networks run with four threads per training row by default; TBNN_TU_TPR=2 (read when the handle is
    planned) keeps two.  Both variants must agree with the oracle and with each other."""
    arch, lik, X, Y, TH, HY = _problem(name)
    C = TH.shape[0]
    eng4 = _engine(arch, lik, chains=C)
    eng4.set_data(X, Y)
    lp4, g4, _ = eng4.logp_grad(TH, HY)
    monkeypatch.setenv("TBNN_TU_TPR", "2")
    eng2 = _engine(arch, lik, chains=C)
    eng2.set_data(X, Y)
    assert eng2.sweep_info()["kernel"] == "k_train_umma"
    assert eng2.sweep_info()["smem_bytes"] != eng4.sweep_info()["smem_bytes"]          # a different plan was made
    lp2, g2, _ = eng2.logp_grad(TH, HY)
    lp4, g4, lp2, g2 = (t.cpu().numpy() for t in (lp4, g4, lp2, g2))
    for c in range(C):
        lp_ref, g_ref = analytic.main_value_and_grad(arch, lik, r32(TH[c]), r32(HY[c]), r32(X), r32(Y))
        assert abs(lp2[c] - lp_ref) <= 1e-5 * abs(lp_ref) and rel(g2[c], g_ref) <= 1e-5, (name, c, rel(g2[c], g_ref))
        assert abs(lp2[c] - lp4[c]) <= 5e-6 * abs(lp4[c]) and rel(g2[c], g4[c]) <= 1e-5


@pytest.mark.parametrize("name", ["c4s", "bern128"])      # (3-output networks plan one tile in flight either way)
def test_train_umma_128_wide_one_tile_in_flight_variant(name, monkeypatch):
    """128-wide networks run two tiles in flight by default (bias gradients as warp column sums); TBNN_TU_NT128=1
    keeps one (bias gradients from the constant-one column of the weight-gradient GEMM).  Both must agree with the
    oracle and with each other."""
    arch, lik, X, Y, TH, HY = _problem(name)
    C = TH.shape[0]
    eng2 = _engine(arch, lik, chains=C)
    eng2.set_data(X, Y)
    lp2, g2, _ = eng2.logp_grad(TH, HY)
    monkeypatch.setenv("TBNN_TU_NT128", "1")
    eng1 = _engine(arch, lik, chains=C)
    eng1.set_data(X, Y)
    assert eng1.sweep_info()["kernel"] == "k_train_umma"
    assert eng1.sweep_info()["smem_bytes"] != eng2.sweep_info()["smem_bytes"]          # a different plan was made
    lp1, g1, _ = eng1.logp_grad(TH, HY)
    lp1, g1, lp2, g2 = (t.cpu().numpy() for t in (lp1, g1, lp2, g2))
    for c in range(C):
        lp_ref, g_ref = analytic.main_value_and_grad(arch, lik, r32(TH[c]), r32(HY[c]), r32(X), r32(Y))
        assert abs(lp1[c] - lp_ref) <= 1e-5 * abs(lp_ref) and rel(g1[c], g_ref) <= 1e-5, (name, c, rel(g1[c], g_ref))
        assert abs(lp1[c] - lp2[c]) <= 5e-6 * abs(lp2[c]) and rel(g1[c], g2[c]) <= 1e-5


def test_train_umma_64_wide_three_outputs():
    """Three outputs on a 64-wide network: the last-block exchange buffer of the four-threads-per-row variant does not
    fit, the planner falls back to two threads per row (fx_stride 4)."""
    arch = wl.mlp_arch([4, 64, 64, 3], "dense", "relu")
    lik = ("gaussian", 0.3)
    rng = np.random.default_rng(7)
    N, C = 300, 2
    X, Y = rng.normal(size=(N, 4)), rng.normal(size=(N, 3))
    P, H = wl.init_theta(arch).size, wl.init_hyper(arch, lik).size
    TH = np.stack([wl.init_theta(arch, seed=3 + c) * 0.7 + 0.05 * rng.normal(size=P) for c in range(C)])
    HY = np.stack([wl.init_hyper(arch, lik) + 0.05 * rng.normal(size=H) for c in range(C)])
    eng = _engine(arch, lik, chains=C)
    eng.set_data(X, Y)
    assert eng.sweep_info()["kernel"] == "k_train_umma"
    lp, g, _ = eng.logp_grad(TH, HY)
    lp, g = lp.cpu().numpy(), g.cpu().numpy()
    for c in range(C):
        lp_ref, g_ref = analytic.main_value_and_grad(arch, lik, r32(TH[c]), r32(HY[c]), r32(X), r32(Y))
        assert abs(lp[c] - lp_ref) <= 1e-5 * abs(lp_ref) and rel(g[c], g_ref) <= 1e-5, (c, rel(g[c], g_ref))


def test_train_umma_trajectory_and_transition_match_oracle():
    """Fixed-momentum L-step trajectory end point within 1e-4 and the Metropolis decision, tcgen05 sweep inside."""
    from oracle import hmc
    arch, lik, X, Y, TH, HY = _problem("relu64")
    eng = _engine(arch, lik, chains=1)
    eng.set_data(X, Y)
    rng = np.random.default_rng(3)
    p0 = rng.normal(size=(1, TH.shape[1]))
    t = lambda a: torch.tensor(r32(a))
    vg = hmc.make_main_vg(arch, lik, t(HY[0]), t(X), t(Y))
    eps, L = 2e-5, 12
    th1, p1, lp1, _ = hmc.leapfrog(vg, t(TH[0]), t(p0[0]), eps, L)
    a, b, c, _ = eng.trajectory(TH[:1], HY[:1], p0, eps, L)
    assert rel(a.cpu().numpy()[0], th1.numpy()) <= 1e-4 and rel(b.cpu().numpy()[0], p1.numpy()) <= 1e-4
    assert abs(c.item() - lp1.item()) <= 1e-4 * abs(lp1.item())


def test_train_umma_many_chains_several_items_per_cta():
    """300 chains of the C3 network on 148 SMs: every CTA walks through two or three chains."""
    C, N = 300, 512
    cfg = wl.c3(N=N, chains=C)
    arch, lik = cfg["arch"], cfg["lik"]
    TH = np.stack([wl.init_theta(arch, seed=c, slope=cfg["slope"]) for c in range(C)])
    HY = np.tile(wl.init_hyper(arch, lik), (C, 1))
    out = {}
    for fl in (0, _lib.FLAG_NO_UMMA_TRAIN):
        eng = _engine(arch, lik, chains=C, flags=fl)
        eng.set_data(cfg["X"], cfg["Y"])
        lp, g, _ = eng.logp_grad(TH, HY)
        out[fl] = (lp.cpu().numpy().astype(np.float64), g.cpu().numpy().astype(np.float64))
    (l0, g0), (l1, g1) = out[0], out[_lib.FLAG_NO_UMMA_TRAIN]
    err = np.abs(g0 - g1).max(axis=1) / np.abs(g1).max(axis=1)
    assert np.abs(l0 - l1).max() <= 5e-6 * np.abs(l1).max()
    assert np.all(np.abs(l0 - l1) <= 1e-5 * np.abs(l1))
    assert np.median(err) <= 5e-6
    assert np.mean(err <= 1e-5) >= 0.95             # the rest: a pre-activation at its kink (see the module docstring)
    assert err.max() <= 8.0 / N
    # spot-check three chains (first, one from a CTA's second item, last) against the fp64 oracle
    for c in (0, 200, C - 1):
        if err[c] > 1e-5:
            continue
        lp_ref, g_ref = analytic.main_value_and_grad(arch, lik, r32(TH[c]), r32(HY[c]), r32(cfg["X"]), r32(cfg["Y"]))
        assert abs(l0[c] - lp_ref) <= 1e-5 * abs(lp_ref) and rel(g0[c], g_ref) <= 2e-5


def _chunked_oracle(arch, lik, theta, hyper, X, Y, chunk):
    """fp64 log-posterior and gradient over a large data set, accumulated chunk by chunk (the likelihood and its
    gradient are sums over rows; the prior is counted once)."""
    lp, g = 0.0, None
    n = 0
    for i in range(0, len(X), chunk):
        xs, ys = r32(X[i:i + chunk]), r32(Y[i:i + chunk])
        ll, gl, _, _ = analytic.loglik_and_grad(arch, lik, theta, xs, ys, hyper[-1] if lik[0] == "gaussian" else None)
        lp += ll
        g = gl if g is None else g + gl
        n += 1
    m, gm = analytic.main_value_and_grad(arch, lik, theta, hyper, r32(X[:8]), r32(Y[:8]))
    l8, g8, _, _ = analytic.loglik_and_grad(arch, lik, theta, r32(X[:8]), r32(Y[:8]), hyper[-1] if lik[0] == "gaussian" else None)
    return lp + (m - l8), g + (gm - g8)


def test_c4_shape_262144_rows_against_fp64():
    """C4 network at 262,144 rows (2,048 tiles of 128): fp32 accumulation -- per-tile tensor-memory accumulators, fp32
    vector reductions into the CTA's slice, fp64 across CTAs -- against the fp64 oracle."""
    N, D = 262144, 32
    rng = np.random.default_rng(7)
    X = rng.standard_normal((N, D), dtype=np.float32)
    Y = (np.tanh(X[:, :4].sum(axis=1)) + 0.1 * rng.standard_normal(N, dtype=np.float32)).astype(np.float32)
    arch = wl.mlp_arch([D, 128, 128, 128, 1], "dense", "relu")
    lik = ("gaussian", 0.1)
    th = (wl.init_theta(arch, seed=0) * 0.5)
    hy = wl.init_hyper(arch, lik)
    lp_ref, g_ref = _chunked_oracle(arch, lik, r32(th), r32(hy), X, Y, 16384)
    for fl, kern in ((0, "k_train_umma"), (_lib.FLAG_NO_UMMA_TRAIN, "k_partial")):
        eng = _engine(arch, lik, flags=fl)
        eng.set_data(X, Y)
        assert eng.sweep_info()["kernel"] == kern
        lp, g, _ = eng.logp_grad(th[None], hy[None])
        assert abs(lp.item() - lp_ref) <= 1e-5 * abs(lp_ref), (kern, lp.item(), lp_ref)
        assert rel(g.cpu().numpy()[0], g_ref) <= 1e-5, (kern, rel(g.cpu().numpy()[0], g_ref))


def test_c2l_full_size_1048576_rows_against_fp64():
    """C2-L (1,048,576 x 784, 784-20-20-1, Bernoulli): the wide-first-layer sweep at the size the HBM-roofline claim
    is made on, against the fp64 oracle accumulated in chunks."""
    N, D = 1048576, 784
    g = torch.Generator(device="cuda").manual_seed(21)
    Xd = torch.rand((N, D), generator=g, device="cuda", dtype=torch.float32)
    arch = wl.mlp_arch([D, 20, 20, 1], "dense", "relu", "sigmoid")
    lik = ("bernoulli",)
    th = wl.init_theta(arch, seed=0) * 0.2
    hy = wl.init_hyper(arch, lik)
    teacher = torch.tensor(wl.init_theta(arch, seed=3)[:D * 20].reshape(20, D), dtype=torch.float32, device="cuda")
    score = (Xd - 0.5) @ teacher.t()
    Yd = (score[:, 0] > score[:, 0].median()).to(torch.float32)
    eng = _engine(arch, lik)
    eng.set_data(Xd, Yd)
    assert eng.sweep_info()["kernel"] == "k_sweep_wide2"
    lp, gr, _ = eng.logp_grad(th[None], hy[None])
    X, Y = Xd.cpu().numpy(), Yd.cpu().numpy()
    lp_ref, g_ref = _chunked_oracle(arch, lik, r32(th), r32(hy), X, Y, 65536)
    assert abs(lp.item() - lp_ref) <= 1e-5 * abs(lp_ref), (lp.item(), lp_ref)
    assert rel(gr.cpu().numpy()[0], g_ref) <= 1e-5, rel(gr.cpu().numpy()[0], g_ref)


def test_posterior_predictive_moments_agree_with_oracle_sampler_within_mc_error():
    """Sampler-level parity (BASELINE.json north_star: "posterior predictive mean and sd must agree within Monte Carlo
    error"): the CUDA sampler (64 batched chains, Philox momentum, on-device leapfrog + Metropolis) and the oracle
    sampler (TFP-ordered leapfrog + MH restated on torch-CPU, its own RNG) target the same posterior of a small
    regression network at fixed hyper parameters; posterior-predictive mean and sd at five inputs must agree within
    5 standard errors, the standard errors taken from the spread of independent chains."""
    from oracle import hmc
    rng = np.random.default_rng(4)
    N = 11
    X = np.linspace(-2, 2, N)[:, None]
    Y = (np.sin(2 * X[:, 0]) + 0.1 * rng.normal(size=N))[:, None]
    arch, lik = wl.mlp_arch([1, 4, 1], "denseGaussian", "tanh"), ("fixed", 0.3)
    hy = wl.init_hyper(arch, lik)
    P = wl.init_theta(arch).size
    Xq = np.linspace(-1.5, 1.5, 5)[:, None]
    eps, L, burn, keep = 0.03, 15, 100, 300

    def summarise(draws):                 # draws [chains, keep, P] -> per-chain predictive mean / sd at Xq: [chains, 5] each
        C_, K_, _ = draws.shape
        f = np.empty((C_, K_, len(Xq)))
        for c in range(C_):
            for k in range(K_):
                th = targets_forward(draws[c, k])
                f[c, k] = th
        return f.mean(axis=1), f.std(axis=1)

    from oracle import targets as T

    def targets_forward(theta):
        return T.forward(arch, T.unflatten_theta(arch, torch.tensor(theta)), torch.tensor(Xq)).numpy()[0]

    # ---- oracle sampler: 6 independent chains
    t = lambda a: torch.tensor(np.asarray(a, dtype=np.float64))
    vg = hmc.make_main_vg(arch, lik, t(hy), t(X), t(Y))
    gen = torch.Generator().manual_seed(11)
    Co = 6
    od = np.empty((Co, keep, P))
    for c in range(Co):
        th = t(0.3 * rng.normal(size=P))
        for it in range(burn + keep):
            p = torch.randn(P, generator=gen, dtype=torch.float64)
            u = float(torch.rand(1, generator=gen, dtype=torch.float64))
            th, _, _, _, _, _ = hmc.hmc_step(vg, th, p, u, eps, L)
            if it >= burn:
                od[c, it - burn] = th.numpy()
    # ---- CUDA sampler: 64 chains in one launch per transition (fp32)
    Cg = 64
    eng = _engine(arch, lik, chains=Cg)
    eng.set_data(X, Y)
    th = eng.tensor(0.3 * rng.normal(size=(Cg, P))).clone()
    HY = np.tile(hy, (Cg, 1))
    gd = np.empty((Cg, keep, P))
    acc = []
    for it in range(burn + keep):
        s = eng.hmc_step(th, HY, 2024, it, eps, L)
        if it >= burn:
            gd[:, it - burn] = th.cpu().numpy()
            acc.append(float(s[:, 1].mean()))
    assert 0.5 < np.mean(acc) <= 1.0                       # a working sampler: the chains move
    om, os_ = summarise(od)
    gm, gs = summarise(gd)
    for a, b in ((om, gm), (os_, gs)):                      # predictive mean, then predictive sd
        se = np.sqrt(a.var(axis=0, ddof=1) / a.shape[0] + b.var(axis=0, ddof=1) / b.shape[0])
        z = np.abs(a.mean(axis=0) - b.mean(axis=0)) / se
        assert z.max() < 5.0, (z, a.mean(axis=0), b.mean(axis=0))
