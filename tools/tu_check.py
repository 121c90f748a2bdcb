"""Developer check of the tcgen05 training sweep (k_train_umma) against the fp64 oracle and the FFMA tile engine."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import analytic
from tensorbnn_b200 import workloads as wl
from tensorbnn_b200.engine import Engine

CASES = [
    ("relu64", wl.mlp_arch([3, 64, 64, 1], "dense", "relu"), ("gaussian", 0.1), 300, 2),
    ("c3s", wl.mlp_arch([1, 64, 64, 64, 1], "dense", "squareprelu"), ("gaussian", 0.1), 200, 2),
    ("tanh64", wl.mlp_arch([5, 64, 64, 64, 2], "denseGaussian", "tanh"), ("fixed", 0.3), 515, 3),
    ("c4s", wl.mlp_arch([32, 128, 128, 128, 1], "dense", "relu"), ("gaussian", 0.1), 700, 1),
    ("bern128", wl.mlp_arch([9, 128, 128, 1], "dense", "relu", "sigmoid"), ("bernoulli",), 1000, 2),
    ("c3full", wl.mlp_arch([1, 64, 64, 64, 1], "dense", "squareprelu"), ("gaussian", 0.1), 4096, 4),
]
only = sys.argv[1:] 
for name, arch, lik, N, C in CASES:
    if only and name not in only:
        continue
    rng = np.random.default_rng(1)
    D = arch[0][1]
    out = [l for l in arch if l[0] in ("dense", "denseGaussian")][-1][2]
    X = rng.normal(size=(N, D))
    Y = (rng.random(N) > 0.5).astype(np.float64) if lik[0] == "bernoulli" else rng.normal(size=(N, out))
    TH = np.stack([wl.init_theta(arch, seed=5 + 17 * c) * 0.7 + 0.05 * rng.normal(size=wl.init_theta(arch).size) for c in range(C)])
    HY = np.stack([wl.init_hyper(arch, lik) + 0.05 * rng.normal(size=wl.init_hyper(arch, lik).size) for c in range(C)])
    eng = Engine(arch, lik, dtype=torch.float32, chains=C)
    eng.set_data(X, Y)
    print(name, eng.sweep_info(), flush=True)
    lp, g, _ = eng.logp_grad(TH, HY)
    torch.cuda.synchronize()
    lp, g = lp.cpu().numpy(), g.cpu().numpy()
    ref = Engine(arch, lik, dtype=torch.float32, chains=C, flags=64)
    ref.set_data(X, Y)
    lp2, g2, _ = ref.logp_grad(TH, HY)
    lp2, g2 = lp2.cpu().numpy(), g2.cpu().numpy()
    r32 = lambda a: np.asarray(a).astype(np.float32).astype(np.float64)
    for c in range(C):
        lpo, go = analytic.main_value_and_grad(arch, lik, r32(TH[c]), r32(HY[c]), r32(X), r32(Y))
        e1 = np.abs(g[c] - go).max() / np.abs(go).max()
        e2 = np.abs(g2[c] - go).max() / np.abs(go).max()
        print("  chain %d: logp umma %.8g ffma %.8g oracle %.8g | grad err umma %.2e ffma %.2e" % (c, lp[c], lp2[c], lpo, e1, e2), flush=True)
        if e1 > 1e-4:
            # where is the error: per tensor
            off = 0
            for shp in wl.theta_shapes(arch):
                n = int(np.prod(shp))
                d = np.abs(g[c][off:off + n] - go[off:off + n]).max()
                print("     tensor", shp, "max abs err %.3e (ref max %.3e)" % (d, np.abs(go[off:off + n]).max()))
                off += n
    t = eng.time_sweep(TH, iters=5)
    t2 = ref.time_sweep(TH, iters=5)
    print("  sweep ms: umma %.3f  ffma %.3f" % (t[0], t2[0]), flush=True)
