"""Runs a few C1 trajectories (persistent narrow kernel) -- developer aid for ncu."""
import sys
sys.path.insert(0, ".")
import numpy as np
import torch
from tensorbnn_b200 import workloads as wl
from tensorbnn_b200.engine import Engine

cfg = wl.c1("a")
arch, lik = cfg["arch"], cfg["lik"]
eng = Engine(arch, lik, chains=1)
eng.set_data(cfg["X"], cfg["Y"])
th = eng.tensor(wl.init_theta(arch, seed=1000)[None]).clone()
hy = eng.tensor(wl.init_hyper(arch, lik)[None]).clone()
for i in range(4):
    eng.hmc_step(th, hy, 1, i, 1e-3, 200)
torch.cuda.synchronize()
print("ok")
