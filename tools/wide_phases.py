"""Prints the phase clocks of one wide-sweep CTA (developer aid; GPU only)."""
import ctypes as C
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from tensorbnn_b200 import _lib, workloads as wl
from tensorbnn_b200.engine import Engine, _ptr, _stream

N = int(sys.argv[1]) if len(sys.argv) > 1 else 9600
cfg = wl.c2(N=N)
arch, lik = cfg["arch"], cfg["lik"]
eng = Engine(arch, lik, chains=1)
eng.set_data(cfg["X"], cfg["Y"])
th = eng.tensor(wl.init_theta(arch, seed=0)[None] * 0.2)
buf = (C.c_longlong * 64)()
_lib.check(eng.lib.tbnn_wide_profile(eng.h, _ptr(th), buf, _stream()))
n = int(buf[0])
t = np.array([buf[1 + i] for i in range(n)], dtype=np.int64)
print(eng.sweep_info(), "marks", n)
names = ["start", "data", "fwd", "bar1", "reduce", "narrow", "bar3", "bwd"]
print("prologue->first pass:", t[1] - t[0])
i = 1
p = 0
while i + 8 <= n - 1 and p < 7:
    seg = t[i:i + 9] if i + 9 <= n else t[i:]
    d = np.diff(seg[:9]) if len(seg) >= 9 else np.diff(seg)
    print("pass", p, dict(zip(["wait", "fwd", "bar1", "reduce", "narrow", "bar3", "bwd", "accum+next"], d.tolist())))
    i += 8
    p += 1
print("total", t[-1] - t[0], "last two marks", t[-1] - t[-2])
