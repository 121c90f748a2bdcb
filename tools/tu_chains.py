import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tensorbnn_b200 import workloads as wl
from tensorbnn_b200.engine import Engine
C = int(sys.argv[1]); N = int(sys.argv[2])
cfg = wl.c3(N=N, chains=C)
arch, lik = cfg["arch"], cfg["lik"]
theta = np.stack([wl.init_theta(arch, seed=c, slope=cfg["slope"]) for c in range(C)])
hyper = np.tile(wl.init_hyper(arch, lik), (C, 1))
out = {}
for fl in (0, 64):
    eng = Engine(arch, lik, dtype=torch.float32, chains=C, flags=fl)
    eng.set_data(cfg["X"], cfg["Y"])
    lp, g, _ = eng.logp_grad(theta, hyper)
    out[fl] = (lp.cpu().numpy().astype(np.float64), g.cpu().numpy().astype(np.float64))
(l0, g0), (l1, g1) = out[0], out[64]
err = np.abs(g0 - g1).max(axis=1) / np.abs(g1).max(axis=1)
lerr = np.abs(l0 - l1) / np.abs(l1)
order = np.argsort(-err)[:8]
print("C", C, "N", N, "worst chains", [(int(c), float("%.2e" % err[c]), float("%.2e" % lerr[c])) for c in order])
print("median grad err %.2e, chains with err > 1e-5: %d" % (np.median(err), int((err > 1e-5).sum())))
c = order[0]
off = 0
for shp in wl.theta_shapes(arch):
    n = int(np.prod(shp))
    d = np.abs(g0[c][off:off + n] - g1[c][off:off + n]).max()
    print("   chain", c, "tensor", shp, "max abs diff %.3e ref max %.3e" % (d, np.abs(g1[c][off:off + n]).max()))
    off += n
