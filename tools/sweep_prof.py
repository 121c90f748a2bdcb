"""Launches the row sweep of a C2-shaped problem a few times (developer aid for ncu; GPU only).
usage: python tools/sweep_prof.py [N] [flags]"""
import sys

sys.path.insert(0, ".")
from tensorbnn_b200 import workloads as wl
from tensorbnn_b200.engine import Engine

N = int(sys.argv[1]) if len(sys.argv) > 1 else 9600
flags = int(sys.argv[2]) if len(sys.argv) > 2 else 0
cfg = wl.c2(N=N)
arch, lik = cfg["arch"], cfg["lik"]
eng = Engine(arch, lik, chains=1, flags=flags)
eng.set_data(cfg["X"], cfg["Y"])
th = eng.tensor(wl.init_theta(arch, seed=0)[None] * 0.2)
print(eng.sweep_info())
print("avg_ms %.5f min_ms %.5f" % eng.time_sweep(th, iters=10))
