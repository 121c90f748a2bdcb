"""Developer aid (GPU): compares the tcgen05 wide sweep with the generic tile engine component by component."""
import sys
sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import numpy as np
import torch
from tensorbnn_b200 import _lib
from tensorbnn_b200.engine import Engine
import test_gpu_parity as tp

for key, N, chains in (("c2s", 96, 1), ("c2s", 9600, 1), ("wide_sq", 77, 2), ("wide_single", 40, 1), ("wide_prelu", 41, 1), ("wide32", 97, 1)):
    arch, lik, X, Y, TH, HY = tp.problem(key, N, chains=chains)
    res = {}
    for name, flags in (("umma", _lib.FLAG_UMMA_SWEEP), ("generic", _lib.FLAG_NO_WIDE)):
        eng = Engine(arch, lik, dtype=torch.float32, chains=chains, flags=flags)
        eng.set_data(X, Y)
        lp, g, st = eng.logp_grad(TH, HY)
        res[name] = (lp.cpu().numpy(), g.cpu().numpy(), st.cpu().numpy(), eng.sweep_info())
    D, out = arch[0][1], arch[0][2]
    gu, gg = res["umma"][1], res["generic"][1]
    print(key, N, res["umma"][3], "logp", res["umma"][0], res["generic"][0])
    for c in range(chains):
        w_u, w_g = gu[c][:D * out].reshape(out, D), gg[c][:D * out].reshape(out, D)
        sc = np.abs(w_g).max()
        err = np.abs(w_u - w_g) / sc
        print("  chain", c, "W1 relerr max %.3e; by feature block of 8:" % err.max(),
              np.array2string(err.max(axis=0).reshape(-1, 8).max(axis=1)[:20], precision=1),
              "by output:", np.array2string(err.max(axis=1), precision=1))
        r_u, r_g = gu[c][D * out:], gg[c][D * out:]
        print("  rest relerr %.3e  b1 err %.3e" % (np.abs(r_u - r_g).max() / np.abs(r_g).max(),
                                                 np.abs(r_u[:out] - r_g[:out]).max() / np.abs(r_g[:out]).max()))
