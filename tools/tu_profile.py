"""clock64 timeline of CTA 0 of k_train_umma (library built with -DTBNN_TU_PROFILE, see tools/build_tu_profile.sh)."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tensorbnn_b200 import workloads as wl, _lib
from tensorbnn_b200.engine import Engine
which = sys.argv[1]
cfg = wl.c3(chains=148) if which == "c3" else wl.c4(N=148 * 128 * 8)
arch, lik, Cn = cfg["arch"], cfg["lik"], cfg["chains"]
theta = np.stack([wl.init_theta(arch, seed=c, slope=cfg.get("slope", 0.2)) for c in range(Cn)])
hyper = np.tile(wl.init_hyper(arch, lik), (Cn, 1))
eng = Engine(arch, lik, chains=Cn)
eng.set_data(cfg["X"], np.asarray(cfg["Y"]).reshape(len(cfg["X"]), -1))
lib = _lib.load()
cap = 4096
buf = (C.c_longlong * (4 * cap))()
lib.tbnn_tu_profile.argtypes = [C.POINTER(C.c_longlong), C.c_int]
eng.logp_grad(theta, hyper); torch.cuda.synchronize()
lib.tbnn_tu_profile(buf, cap)            # discard the cold launch
eng.logp_grad(theta, hyper); torch.cuda.synchronize()
n = lib.tbnn_tu_profile(buf, cap)
n0, n1 = n & 0xFFFF, n >> 16
a = np.frombuffer(buf, dtype=np.int64).reshape(2, 2, cap)
for role, cnt in ((0, n0), (1, n1)):
    tags, clk = a[role, 0, :cnt], a[role, 1, :cnt]
    base = clk[0]
    print("role", role, "entries", cnt)
    for i in range(min(cnt, 140)):
        t = int(tags[i]); k, ti, ph = (t >> 20) & 0xFF, (t >> 16) & 0xF, t & 0xFFFF
        print("  k=%3d ti=%d ph=%2d  t=%8d  dt=%6d" % (k, ti, ph, clk[i] - base, clk[i] - (clk[i - 1] if i else base)))
