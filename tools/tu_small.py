import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import analytic
from tensorbnn_b200 import workloads as wl
from tensorbnn_b200.engine import Engine
G, HW, N, act = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
dims = [3] + [HW] * G + [1]
arch, lik = wl.mlp_arch(dims, "dense", act), ("gaussian", 0.1)
rng = np.random.default_rng(1)
X, Y = rng.normal(size=(N, 3)), rng.normal(size=(N, 1))
TH = (wl.init_theta(arch, seed=5) * 0.7)[None]
HY = wl.init_hyper(arch, lik)[None]
eng = Engine(arch, lik, dtype=torch.float32, chains=1)
eng.set_data(X, Y)
print(sys.argv[1:], eng.sweep_info(), flush=True)
lp, g, _ = eng.logp_grad(TH, HY)
torch.cuda.synchronize()
r32 = lambda a: np.asarray(a).astype(np.float32).astype(np.float64)
lpo, go = analytic.main_value_and_grad(arch, lik, r32(TH[0]), r32(HY[0]), r32(X), r32(Y))
print("  logp", lp.item(), lpo, "grad err", np.abs(g.cpu().numpy()[0] - go).max() / np.abs(go).max(), flush=True)
