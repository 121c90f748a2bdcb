"""Sweep time of the training row sweep at the BASELINE shapes: tcgen05 (k_train_umma) vs FFMA tile engine."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tensorbnn_b200 import workloads as wl
from tensorbnn_b200.engine import Engine
which = sys.argv[1]
flags_list = [int(f) for f in sys.argv[2:]] or [0, 64]
if which.startswith("c3:"):                      # c3:<chains>, e.g. c3:128 = one GPU's share of the 8-GPU split
    cfg = wl.c3(chains=int(which[3:]))
elif which == "c3":
    cfg = wl.c3(chains=1024)
elif which == "c3s":
    cfg = wl.c3(chains=148)
else:
    cfg = wl.c4()
arch, lik, C = cfg["arch"], cfg["lik"], cfg["chains"]
theta = np.stack([wl.init_theta(arch, seed=c, slope=cfg.get("slope", 0.2)) * cfg.get("wscale", 1.0) for c in range(C)])
hyper = np.tile(wl.init_hyper(arch, lik), (C, 1))
X = torch.tensor(cfg["X"], dtype=torch.float32).cuda()
Y = torch.tensor(np.asarray(cfg["Y"]).reshape(len(cfg["X"]), -1), dtype=torch.float32).cuda()
dims = [l for l in arch if l[0].startswith("dense")]
F = sum(l[1] * l[2] for l in dims)
flops = C * X.shape[0] * (6 * F - 2 * dims[0][1] * dims[0][2])
res = {}
for fl in flags_list:
    eng = Engine(arch, lik, dtype=torch.float32, chains=C, flags=fl)
    eng.set_data(X, Y)
    th = eng.tensor(theta)
    lp, g, _ = eng.logp_grad(th, hyper)
    torch.cuda.synchronize()
    avg, mn = eng.time_sweep(th, iters=10)
    res[fl] = (lp.cpu().numpy(), g.cpu().numpy())
    print(which, "flags", fl, eng.sweep_info(), "sweep ms avg %.3f min %.3f -> %.1f TFLOP/s fp32-equivalent, %.0f chain-steps/s"
          % (avg, mn, flops / (avg * 1e-3) / 1e12, C / (avg * 1e-3)), flush=True)
if len(res) == 2:
    (l0, g0), (l1, g1) = res[flags_list[0]], res[flags_list[1]]
    print("  max rel logp diff %.2e, grad diff %.2e (vs max |g|)" % (np.abs(l0 - l1).max() / np.abs(l1).max(),
          np.abs(g0 - g1).max() / np.abs(g1).max()))
