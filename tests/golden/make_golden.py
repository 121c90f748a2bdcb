"""Generates tests/golden/*.json: known-answer vectors for the hot path.

The reference (TensorBNN) cannot be imported here (TensorFlow / TFP are absent and uninstallable), so these
vectors come from the fp64 CPU oracle (oracle/targets.py, torch autograd), cross-checked at generation time
against the independent analytic implementation (oracle/analytic.py).  They freeze the oracle: a later edit
of either implementation that changes a value breaks tests/test_golden.py.  Inputs are stored in full so the
fixtures are self-contained (no RNG dependency at test time).

    python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import analytic, hmc, targets  # noqa: E402
from tensorbnn_b200 import workloads as wl  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
torch.set_default_dtype(torch.float64)

CASES = {
    # name: (arch, lik, N, eps, L)
    "c1a": (wl.mlp_arch([1, 10, 10, 10, 1], "denseGaussian", "tanh"), ("fixed", 0.1), 11, 1e-3, 25),
    "c1b": (wl.mlp_arch([1, 10, 10, 10, 1], "dense", "relu"), ("gaussian", 0.1), 11, 1e-3, 25),
    "bern": (wl.mlp_arch([7, 5, 4, 1], "dense", "relu", "sigmoid"), ("bernoulli",), 23, 2e-3, 20),
    "sqp": (wl.mlp_arch([3, 6, 6, 2], "dense", "squareprelu"), ("gaussian", 0.2), 17, 1e-3, 20),
    "prelu": (wl.mlp_arch([3, 6, 5, 1], "denseGaussian", "prelu"), ("fixed", 0.3), 19, 1e-3, 20),
    "mixed": ([("dense", 4, 6), ("elu",), ("denseGaussian", 6, 5), ("Exp",), ("dense", 5, 3),
               ("leakyrelu", 0.3), ("dense", 3, 1), ("sigmoid",)], ("bernoulli",), 15, 1e-3, 15),
    "wide": (wl.mlp_arch([64, 8, 4, 1], "dense", "relu", "sigmoid"), ("bernoulli",), 37, 1e-3, 10),
}


def build(name):
    arch, lik, N, eps, L = CASES[name]
    rng = np.random.default_rng(sum(map(ord, name)))
    D = arch[0][1]
    out = [l for l in arch if l[0] in ("dense", "denseGaussian")][-1][2]
    if name in ("c1a", "c1b"):
        c = wl.c1(name[-1])
        X, Y = c["X"], np.asarray(c["Y"]).reshape(-1, 1)       # Examples/trainRegression.py:33-36
    else:
        X = rng.random((N, D)) if D > 32 else rng.normal(size=(N, D))
        Y = (rng.random((N, out)) > 0.5).astype(np.float64) if lik[0] == "bernoulli" else rng.normal(size=(N, out))
    theta = wl.init_theta(arch, seed=7) * (0.3 if D > 32 else 0.7) + 0.05 * rng.normal(size=targets.num_params(arch))
    hyper = wl.init_hyper(arch, lik)
    hyper = hyper + 0.05 * rng.normal(size=hyper.size)
    mom = rng.normal(size=theta.size)
    lp, g = targets.main_value_and_grad(arch, lik, torch.tensor(theta), torch.tensor(hyper), torch.tensor(X), torch.tensor(Y))
    lp2, g2 = analytic.main_value_and_grad(arch, lik, theta, hyper, X, Y)
    assert abs(float(lp) - lp2) <= 1e-11 * abs(lp2), (name, float(lp), lp2)
    assert np.abs(g.numpy() - g2).max() <= 1e-10 * np.abs(g2).max(), name
    hlp, hg = targets.hyper_value_and_grad(arch, lik, torch.tensor(theta), torch.tensor(hyper), torch.tensor(X), torch.tensor(Y))
    hlp2, hg2 = analytic.hyper_value_and_grad(arch, lik, theta, hyper, X, Y)
    assert abs(float(hlp) - hlp2) <= 1e-11 * abs(hlp2), name
    assert np.abs(hg.numpy() - hg2).max() <= 1e-9 * max(np.abs(hg2).max(), 1.0), name
    vg = hmc.make_main_vg(arch, lik, torch.tensor(hyper), torch.tensor(X), torch.tensor(Y))
    th1, p1, lp1, g1 = hmc.leapfrog(vg, torch.tensor(theta), torch.tensor(mom), eps, L)
    lar = hmc.log_accept_ratio(lp, lp1, torch.tensor(mom), p1)
    f = targets.forward(arch, targets.unflatten_theta(arch, torch.tensor(theta)), torch.tensor(X))
    r = lambda a: [float(v) for v in np.asarray(a, dtype=np.float64).reshape(-1)]
    return {
        "name": name, "arch": [list(l) for l in arch], "lik": list(lik), "N": int(X.shape[0]), "D": int(D), "out": int(out),
        "X": r(X), "Y": r(Y), "theta": r(theta), "hyper": r(hyper), "momentum": r(mom), "eps": eps, "L": L,
        "forward": r(f.detach().numpy()),                        # network.predict(train=True): [out, N]
        "logp": float(lp), "grad": r(g.numpy()),                  # network.py:370-392 + autodiff
        "hyper_logp": float(hlp), "hyper_grad": r(hg.numpy()),    # network.py:417-440 + autodiff
        "traj_theta": r(th1.numpy()), "traj_momentum": r(p1.numpy()), "traj_logp": float(lp1),
        "log_accept_ratio": float(lar),                           # TFP semantics, SURVEY App. B
    }


def main():
    for name in CASES:
        d = build(name)
        with open(os.path.join(HERE, name + ".json"), "w") as f:
            json.dump(d, f)
        print(name, "P=%d" % len(d["theta"]), "logp=%.12g" % d["logp"], "lar=%.6g" % d["log_accept_ratio"])


if __name__ == "__main__":
    main()
