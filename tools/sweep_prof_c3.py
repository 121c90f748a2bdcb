"""Launches the row sweep of the C3 shape (batched chains on the tile engine) a few times (developer aid for ncu)."""
import sys
sys.path.insert(0, ".")
import numpy as np
from tensorbnn_b200 import workloads as wl
from tensorbnn_b200.engine import Engine

C = int(sys.argv[1]) if len(sys.argv) > 1 else 148
cfg = wl.c3(chains=C)
arch, lik = cfg["arch"], cfg["lik"]
eng = Engine(arch, lik, chains=C)
eng.set_data(cfg["X"], cfg["Y"])
th = eng.tensor(np.stack([wl.init_theta(arch, seed=1000 + c, slope=cfg["slope"]) for c in range(C)]))
print(eng.sweep_info())
print("avg_ms %.5f min_ms %.5f" % eng.time_sweep(th, iters=5))
