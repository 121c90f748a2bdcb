"""Parity against fixtures produced by the REFERENCE'S OWN SOURCE (tests/golden/make_ref_golden.py runs
/root/reference/tensorBNN/*.py unmodified over the torch-backed TensorFlow stand-in oracle/tfshim/).

CPU part ("not gpu"): the oracle restatement (oracle/targets.py, analytic.py, hmc.py, adapter.py, fileformat.py)
must reproduce what the reference's code computed -- this is what pins the oracle.
GPU part: the CUDA path through the C ABI must reproduce the same vectors: 1e-10 relative in fp64 and 1e-5 in
fp32 for log-posterior and gradient, 1e-9 / 1e-4 for fixed-momentum trajectory end points (BASELINE.json).
"""
import glob
import json
import os
import random

import numpy as np
import pytest
import torch

from oracle import adapter as oadapter
from oracle import analytic, fileformat, hmc, targets

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")
NET_FILES = sorted(glob.glob(os.path.join(GOLD, "ref_*.json")))
NET_NAMES = [os.path.basename(f)[:-5] for f in NET_FILES]
t64 = lambda v: torch.tensor(np.asarray(v, dtype=np.float64))


def load_net(name):
    with open(os.path.join(GOLD, name + ".json")) as f:
        d = json.load(f)
    arch = [tuple(l) for l in d["arch"]]
    lik = tuple(d["lik"])
    a = lambda k, shape=None: np.asarray(d[k], dtype=np.float64).reshape(shape if shape else -1)
    return d, arch, lik, a("X", (d["N"], d["D"])), a("Y", (d["N"], d["out"])), a


def load_fn(name):
    with open(os.path.join(GOLD, "reffn_" + name + ".json")) as f:
        return json.load(f)


def rel(x, ref):
    x, ref = np.asarray(x, dtype=np.float64), np.asarray(ref, dtype=np.float64)
    return np.abs(x - ref).max() / max(np.abs(ref).max(), 1e-300)


def close(x, ref, tol):
    return abs(x - ref) <= tol * max(abs(ref), 1e-300)


def test_reference_fixtures_exist_and_say_where_they_come_from():
    assert len(NET_NAMES) >= 7
    for n in NET_NAMES:
        assert "reference source run unmodified" in load_net(n)[0]["source"]


# ---------------------------------------------------------------------------- oracle vs the reference's own code
@pytest.mark.parametrize("name", NET_NAMES)
def test_oracle_reproduces_reference_epoch(name):
    """One full epoch of network.train (network.py:509-670 -> stepMCMC :359-507) as the reference ran it."""
    d, arch, lik, X, Y, a = load_net(name)
    theta, hyper = a("theta"), a("hyper")
    f = targets.forward(arch, targets.unflatten_theta(arch, t64(theta)), t64(X))
    assert rel(f.numpy().reshape(-1), a("forward")) <= 1e-13
    # main target + gradient (closure network.py:370-392 differentiated by TFP)
    lp, g = targets.main_value_and_grad(arch, lik, t64(theta), t64(hyper), t64(X), t64(Y))
    assert close(float(lp), d["logp"], 1e-13) and rel(g.numpy(), a("grad")) <= 1e-12
    lp2, g2 = analytic.main_value_and_grad(arch, lik, theta, hyper, X, Y)
    assert close(lp2, d["logp"], 1e-12) and rel(g2, a("grad")) <= 1e-10
    # trajectory, log-accept ratio, MH select
    vg = hmc.make_main_vg(arch, lik, t64(hyper), t64(X), t64(Y))
    th1, p1, lp1, _ = hmc.leapfrog(vg, t64(theta), t64(a("momentum")), d["eps"], d["L"])
    assert rel(th1.numpy(), a("traj_theta")) <= 1e-12 and rel(p1.numpy(), a("traj_momentum")) <= 1e-11
    assert close(float(lp1), d["traj_logp"], 1e-12)
    new, lar, prob, acc, _, _ = hmc.hmc_step(vg, t64(theta), t64(a("momentum")), d["u_main"], d["eps"], d["L"])
    assert abs(float(lar) - d["log_accept_ratio"]) <= 1e-9 * max(1.0, abs(d["log_accept_ratio"]))
    assert acc == d["main_accepted"] and rel(new.numpy(), a("theta_after")) <= 1e-12
    assert close(float(prob), d["main_accept_prob"], 1e-9)
    # hyper target + gradient (closure network.py:417-440), hyper HMC, dual averaging (:457-469)
    th_h = a("hyper_theta")
    hlp, hg = targets.hyper_value_and_grad(arch, lik, t64(th_h), t64(hyper), t64(X), t64(Y))
    assert close(float(hlp), d["hyper_logp"], 1e-13) and rel(hg.numpy(), a("hyper_grad")) <= 1e-12
    hlp2, hg2 = analytic.hyper_value_and_grad(arch, lik, th_h, hyper, X, Y)
    assert close(hlp2, d["hyper_logp"], 1e-12) and rel(hg2, a("hyper_grad")) <= 1e-9
    hvg = hmc.make_hyper_vg(arch, lik, t64(th_h), t64(X), t64(Y))
    hnew, hlar, hprob, hacc, hprop, _ = hmc.hmc_step(hvg, t64(hyper), t64(a("hyper_momentum")), d["u_hyper"],
                                                     d["hyper_eps"], d["hyper_L"])
    want = d["hyper_log_accept_ratio"]
    if np.isfinite(want):
        assert abs(float(hlar) - want) <= 1e-9 * max(1.0, abs(want))
        assert rel(hprop.numpy(), a("hyper_traj")) <= 1e-11
    else:
        assert float(hlar) == want          # divergent hyper trajectory: safe_sum -> -inf -> reject
    assert hacc == d["hyper_accepted"] and rel(hnew.numpy(), a("hyper_after")) <= 1e-12
    assert close(float(hprob), d["hyper_accept_prob"], 1e-9) or float(hprob) == d["hyper_accept_prob"]
    h, leb, step = hmc.dual_averaging(d["epoch"], float(hprob), 0.0, 0.0, d["hyper_eps"], d["hyper_eps"], d["burnin"])
    assert abs(h - d["h_after"]) <= 1e-14 and abs(leb - d["log_eps_bar_after"]) <= 1e-12
    assert close(step, d["hyper_step_after"], 1e-12)


def _grads(fn, arrays):
    leaves = [torch.tensor(np.asarray(v, dtype=np.float64), requires_grad=True) for v in arrays]
    val = fn(*leaves)
    gs = torch.autograd.grad(val.sum(), leaves, allow_unused=True)
    return val.detach().numpy(), [np.zeros(l.shape) if g is None else g.numpy() for g, l in zip(gs, leaves)]


def test_oracle_densities_match_reference():
    """BNN_functions.py:7-57 as shipped, incl. the clamp edges and the scalar-sigma normalisation (Q2)."""
    d = load_fn("functions")
    for c in d["multivariateLogProb"]:
        sig = np.asarray(c["sigma"]).reshape(c["sigma_shape"])
        mu = np.asarray(c["mu"]).reshape(c["mu_shape"])
        x = np.asarray(c["x"]).reshape(c["x_shape"])
        val, (gs, gm, gx) = _grads(lambda s, m, xx: targets.multivariate_log_prob(s, m, xx), (sig, mu, x))
        assert close(float(val), c["value"], 1e-13)
        for got, key in ((gs, "d_sigma"), (gm, "d_mu"), (gx, "d_x")):
            want = np.asarray(c[key])
            assert np.abs(got.reshape(-1) - want).max() <= 1e-12 * max(np.abs(want).max(), 1e-300) + 1e-300
    for c in d["cauchyLogProb"]:
        x = np.asarray(c["x"]).reshape(5, 3)
        val, (gg, g0, gx) = _grads(lambda g, m, xx: targets.cauchy_log_prob(g, m, xx), (c["gamma"], c["x0"], x))
        assert rel(val.reshape(-1), c["value"]) <= 1e-13 and close(float(val.sum()), c["sum"], 1e-13)
        assert rel(gx.reshape(-1), c["d_x"]) <= 1e-12 and rel(gg.reshape(-1), c["d_gamma"]) <= 1e-12


def test_oracle_layers_match_reference():
    """layer.py / activationFunctions.py predict, calculateProbs, calculateHyperProbs as shipped."""
    d = load_fn("functions")
    for c in d["dense_layers"]:
        layer = (c["kind"], 4, 3)
        assert c["name"] == c["kind"] and (c["numTensors"], c["numHyperTensors"]) == targets.num_tensors(layer)
        W, b, hy = np.asarray(c["W"]).reshape(3, 4), np.asarray(c["b"]).reshape(3, 1), np.asarray(c["hyper"])
        pv, pg = _grads(lambda h, w, bb: targets.layer_prior(layer, [h[i] for i in range(4)], [w, bb]), (hy, W, b))
        assert close(float(pv), c["prior"], 1e-13)
        assert rel(pg[0], c["prior_d_hyper"]) <= 1e-12 and rel(pg[1].reshape(-1), c["prior_d_W"]) <= 1e-12
        assert rel(pg[2].reshape(-1), c["prior_d_b"]) <= 1e-12
        hv, hg = _grads(lambda h, w, bb: targets.layer_hyper_prob(layer, [h[i] for i in range(4)], [w, bb]), (hy, W, b))
        assert close(float(hv), c["hyper_prob"], 1e-13) and rel(hg[0], c["hyper_prob_d_hyper"]) <= 1e-12
        A = np.asarray(c["A"]).reshape(4, 6)
        assert rel((W @ A + b).reshape(-1), c["predict"]) <= 1e-14
        init = targets.initial_hypers([layer], ("bernoulli",))
        assert rel(np.float32(init.numpy()), c["initial_hypers"]) == 0.0     # tf.cast of python floats: float32 values
    for c in d["param_activations"]:
        layer = (c["kind"], 5)
        assert c["name"] == c["kind"] and (c["numTensors"], c["numHyperTensors"]) == targets.num_tensors(layer)
        s, hy, Z = np.asarray(c["slopes"]), np.asarray(c["hyper"]), np.asarray(c["Z"]).reshape(5, 7)
        nh = hy.size
        pv, pg = _grads(lambda h, sl: targets.layer_prior(layer, [h[i] for i in range(nh)], [sl]), (hy, s))
        assert close(float(pv), c["prior"], 1e-13) and rel(pg[1], c["prior_d_slopes"]) <= 1e-12
        hv, hg = _grads(lambda h, sl: targets.layer_hyper_prob(layer, [h[i] for i in range(nh)], [sl]), (hy, s))
        assert close(float(hv), c["hyper_prob"], 1e-13) and rel(hg[0], c["hyper_prob_d_hyper"]) <= 1e-12
        out = targets.apply_activation(layer, t64(Z), [t64(s)])
        assert rel(out.numpy().reshape(-1), c["predict"]) <= 1e-14
        assert c["default_parameters"] == [0.25] * 5
    for c in d["plain_activations"]:
        layer = (c["kind"],) if c["kind"] != "leakyrelu" else ("leakyrelu", c["alpha"])
        Z = np.asarray(c["Z"]).reshape(5, 7)
        out = targets.apply_activation(layer, t64(Z), [])
        assert rel(out.numpy().reshape(-1), c["predict"]) <= 1e-13, c["kind"]


def test_oracle_likelihoods_match_reference():
    """likelihood.py:69-96,143-169,210-237 as shipped (stub ``predict``), incl. the Bernoulli clip edges."""
    d = load_fn("functions")
    for c in d["likelihoods"]:
        f = np.asarray(c["f"]).reshape(c["f_shape"])
        if c["kind"] in ("gaussian", "fixed"):
            y = np.asarray(c["y"]).reshape(9, 2)

            def fn(ff, hh):
                cur = ff.t()
                sd = hh[0] ** 2 if c["kind"] == "gaussian" else torch.tensor(targets.f32(c["sd"]), dtype=torch.float64)
                return targets.multivariate_log_prob(torch.ones_like(cur) * sd, cur, t64(y))
            val, (gf, gh) = _grads(fn, (f, [c["hyper_last"]]))
            assert close(float(val), c["value"], 1e-13) and rel(gf.reshape(-1), c["d_f"]) <= 1e-12
            if c["kind"] == "gaussian":
                assert rel(gh, c["d_hyper"]) <= 1e-12 and c["mainProbsInHypers"]
                assert c["hypers"] == [0.3 ** 0.5]
        else:
            dt = torch.float64 if c["dtype"] == "float64" else torch.float32
            y = torch.tensor(c["y"], dtype=dt)
            p = torch.tensor(f, dtype=dt, requires_grad=True)
            lo, hi = torch.tensor(1e-8, dtype=dt), torch.tensor(1 - 1e-7, dtype=dt)
            pc = torch.clamp(p, min=lo.item(), max=hi.item())
            vals = (1.0 - y) * torch.log1p(-pc) + y * torch.log(pc)
            (g,) = torch.autograd.grad(vals.sum(), p)
            tol = 1e-13 if dt == torch.float64 else 1e-6
            assert rel(vals.detach().double().numpy().reshape(-1), c["values"]) <= tol
            assert rel(g.double().numpy().reshape(-1), c["d_f"]) <= tol


def test_oracle_adapter_replays_reference_history():
    """paramAdapter.py:39-292 as shipped: 330 updates (random proposals, then GP-UCB grid searches)."""
    d = load_fn("adapter")
    args = d["args"]

    class Replay(object):
        def __init__(self):
            self.u = list(d["uniforms"])
            self.py = random.Random(d["python_random_seed"])

        def random(self):
            return self.u.pop(0)

        def choice(self, seq):
            return self.py.choice(seq)

    ad = oadapter.OracleAdapter(args["e1"], args["L1"], args["el"], args["eu"], args["eNumber"], args["Ll"], args["Lu"],
                                args["lStep"], args["m"], args["k"], a=args["a"], delta=args["delta"],
                                randomSteps=args["randomSteps"], rng=Replay())
    rng = np.random.default_rng(d["jump_seed"])
    state = [np.zeros(s, dtype=np.float32) for s in d["state_shapes"]]
    refits = {r["step"]: r for r in d["refits"]}
    assert len(refits) >= 20 and max(r["size"] for r in d["refits"]) >= 20
    n_grid = 0
    for step, (E, L) in enumerate(d["history"]):
        e, l = float(ad.currentE), float(ad.currentL)
        scale = np.exp(-((e - 4e-3) / 3e-3) ** 2 - ((l - 35.0) / 20.0) ** 2)
        state = [s + np.float32(scale) * rng.normal(size=s.shape).astype(np.float32) for s in state]
        gotE, gotL = ad.update(state)
        assert int(gotL) == L, (step, gotL, L)
        assert abs(float(gotE) - E) <= 1e-6 * abs(E), (step, gotE, E)
        if step in refits:
            r = refits[step]
            assert ad.inverse.shape[0] == r["size"]
            assert close(float(ad.p), r["p"], 1e-6) and close(float(ad.rootbeta), r["rootbeta"], 1e-6)
            assert close(float(ad.s), r["s"], 1e-5)
            assert rel(ad.inverse.reshape(-1), r["inverse"]) <= 2e-3        # float32 inverse of a <=50x50 matrix
            n_grid += ad.i - 1 >= args["randomSteps"] * args["m"]
    assert n_grid >= 10


def test_oracle_reader_matches_reference_writer_and_predictor():
    """The directory written by the reference's network.train (tests/golden/ref_run/) parsed by the restated reader;
    forward passes on the parsed samples equal what the reference's predictor.predict returned."""
    d = load_fn("run")
    run = os.path.join(GOLD, "ref_run")
    assert sorted(os.listdir(run)) == d["files"]
    assert open(os.path.join(run, "summary.txt")).read() == d["summary"]
    mats, hypers = fileformat.load_networks(run + "/")
    assert open(os.path.join(run, "architecture.txt")).read() == d["architecture"]
    # the restated writer replays the reference's schedule: same files, same line counts, same summary (Q9 lag)
    import tempfile
    sch = d["schedule"]
    shapes = [(m[1], m[2]) for m in d["matrix_shapes"]]
    with tempfile.TemporaryDirectory() as tmp:
        fileformat.write_run(tmp, d["architecture"].split(), sch["epochs"], sch["burnin"], sch["samplingStep"],
                             sch["networksPerFile"], lambda it: [np.zeros(s) for s in shapes],
                             lambda it: np.zeros(len(d["hypers"][0])))
        assert sorted(os.listdir(tmp)) == d["files"]
        assert {f: sum(1 for _ in open(os.path.join(tmp, f))) for f in d["files"]} == d["line_counts"]
        assert open(os.path.join(tmp, "summary.txt")).read() == d["summary"]
    assert [list(m.shape) for m in mats] == d["matrix_shapes"] and mats[0].shape[0] == d["numNetworks"]
    assert rel(np.asarray(hypers).reshape(-1), np.asarray(d["hypers"]).reshape(-1)) <= 1e-7
    arch = [tuple(l) for l in d["arch"]]
    Xt = np.asarray(d["Xtest"]).reshape(5, 2)
    for n, key in ((1, "predict_n1"), (2, "predict_n2")):
        for j, m in enumerate(range(0, d["numNetworks"], n)):
            theta = [t64(mat[m]) for mat in mats]
            out = targets.forward(arch, theta, t64(Xt))
            assert rel(out.numpy().reshape(-1), d[key][j]) <= 2e-6          # the reference predicts in float32


def test_metrics_fixture_is_self_consistent():
    d = load_fn("functions")
    for c in d["metrics"]:
        pt, yt = np.asarray(c["predTrain"]), np.asarray(c["realTrain"])
        if c["metric"] == "SquaredError" and not c["scaleExp"]:
            want = np.mean((pt * c["sd"] + c["mean"] - (yt * c["sd"] + c["mean"])) ** 2)
            assert close(c["values"]["squaredErrorTrain"], want, 1e-12)


# ---------------------------------------------------------------------------- CUDA vs the reference's own code
@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
@pytest.mark.parametrize("name", NET_NAMES)
def test_cuda_reproduces_reference_epoch(name, dtype):
    from tensorbnn_b200.engine import Engine
    d, arch, lik, X, Y, a = load_net(name)
    f64 = dtype == torch.float64
    tol, ttol = (1e-10, 1e-9) if f64 else (1e-5, 1e-4)
    theta, hyper = a("theta"), a("hyper")
    eng = Engine(arch, lik, dtype=dtype, chains=1)
    eng.set_data(X, Y)
    out, _ = eng.predict(theta[None], X, want_out=True)
    assert rel(out.cpu().numpy().reshape(-1), a("forward")) <= (1e-11 if f64 else 2e-5)
    lp, g, _ = eng.logp_grad(theta[None], hyper[None])
    assert close(lp.item(), d["logp"], tol) and rel(g.cpu().numpy()[0], a("grad")) <= tol
    hlp, hg = eng.hyper_logp_grad(a("hyper_theta")[None], hyper[None])
    assert close(hlp.item(), d["hyper_logp"], tol) and rel(hg.cpu().numpy()[0], a("hyper_grad")) <= 10 * tol
    big = d["eps"] > 0.1          # ref_rej: a deliberately unstable step size, compared loosely in fp32
    th1, p1, lp1, _ = eng.trajectory(theta[None], hyper[None], a("momentum")[None], d["eps"], d["L"])
    if f64 or not big:
        assert rel(th1.cpu().numpy()[0], a("traj_theta")) <= ttol
        assert rel(p1.cpu().numpy()[0], a("traj_momentum")) <= ttol
        assert close(lp1.item(), d["traj_logp"], ttol)
    # the full main transition with the reference's momentum and uniform
    th = eng.tensor(theta[None]).clone()
    stats = eng.hmc_step(th, hyper[None], 1, 0, d["eps"], d["L"], momentum=a("momentum")[None],
                         u=np.array([d["u_main"]])).cpu().numpy()[0]
    lar = d["log_accept_ratio"]
    assert abs(stats[0] - lar) <= (1e-7 if f64 else 2e-2) * max(1.0, abs(lar))
    assert stats[2] == float(d["main_accepted"])
    if f64 or not big:
        assert rel(th.cpu().numpy()[0], a("theta_after")) <= ttol
    assert abs(stats[1] - d["main_accept_prob"]) <= (1e-7 if f64 else 2e-2)
    # the hyper transition + dual averaging
    hy = eng.tensor(hyper[None]).clone()
    da = eng.tensor(np.array([[0.0, 0.0, d["hyper_eps"]]])).clone()
    hs = eng.hyper_step(a("hyper_theta")[None], hy, 1, 0, d["hyper_L"], float(d["epoch"]), float(d["burnin"]),
                        d["hyper_eps"], da, momentum=a("hyper_momentum")[None], u=np.array([d["u_hyper"]])).cpu().numpy()[0]
    want = d["hyper_log_accept_ratio"]
    if np.isfinite(want):
        assert abs(hs[0] - want) <= (1e-7 if f64 else 5e-3) * max(1.0, abs(want))
    else:
        assert hs[0] == want
    assert rel(hy.cpu().numpy()[0], a("hyper_after")) <= (1e-9 if f64 else 1e-4)
    got = da.cpu().numpy()[0]
    datol = 1e-10 if f64 else 2e-3
    assert abs(got[0] - d["h_after"]) <= datol and abs(got[1] - d["log_eps_bar_after"]) <= datol * 10
    assert close(got[2], d["hyper_step_after"], datol * 10)


@pytest.mark.gpu
def test_cuda_adapter_replays_reference_history():
    """Product paramAdapter (host bookkeeping + tbnn_adapter_ucb grid search) against the reference's history."""
    from tensorbnn_b200.paramAdapter import paramAdapter
    d = load_fn("adapter")
    args = d["args"]

    class Replay(object):
        def __init__(self):
            self.u = list(d["uniforms"])
            self.py = random.Random(d["python_random_seed"])

        def random(self):
            return self.u.pop(0)

        def choice(self, seq):
            return self.py.choice(seq)

    ad = paramAdapter(args["e1"], args["L1"], args["el"], args["eu"], args["eNumber"], args["Ll"], args["Lu"],
                      args["lStep"], args["m"], args["k"], a=args["a"], delta=args["delta"],
                      randomSteps=args["randomSteps"], rng=Replay())
    ad.verbose = False
    rng = np.random.default_rng(d["jump_seed"])
    state = [np.zeros(s, dtype=np.float32) for s in d["state_shapes"]]
    mism = 0
    for step, (E, L) in enumerate(d["history"]):
        e, l = float(ad.currentE), float(ad.currentL)
        scale = np.exp(-((e - 4e-3) / 3e-3) ** 2 - ((l - 35.0) / 20.0) ** 2)
        state = [s + np.float32(scale) * rng.normal(size=s.shape).astype(np.float32) for s in state]
        gotE, gotL = ad.update([torch.tensor(s) for s in state])
        if int(gotL) != L or abs(float(gotE) - E) > 1e-6 * abs(E):
            mism += 1
            break
    assert mism == 0, (step, gotE, gotL, E, L)


@pytest.mark.gpu
def test_cuda_predictor_reads_reference_run():
    """Product predictor on the directory the reference's network.train wrote."""
    from tensorbnn_b200.predictor import predictor
    d = load_fn("run")
    pred = predictor(os.path.join(GOLD, "ref_run") + "/", torch.float32)
    assert pred.numNetworks == d["numNetworks"]
    Xt = np.asarray(d["Xtest"], dtype=np.float32).reshape(5, 2)
    for n, key in ((1, "predict_n1"), (2, "predict_n2")):
        outs = pred.predict(Xt, n=n)
        assert len(outs) == len(d[key])
        for o, want in zip(outs, d[key]):
            assert rel(np.asarray(o).reshape(-1), want) <= 2e-5
