"""Small invocations of the barrier-synchronised kernels for compute-sanitizer (memcheck / racecheck / synccheck)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tensorbnn_b200 import workloads as wl
from tensorbnn_b200.engine import Engine
case = sys.argv[1]
rng = np.random.default_rng(0)
def run(arch, lik, N, flags=0, chains=1, predict=False, traj=False):
    D = arch[0][1]
    X = rng.random((N, D)) if D > 100 else rng.normal(size=(N, D))
    Y = (rng.random(N) > 0.5).astype(np.float64) if lik[0] == "bernoulli" else rng.normal(size=(N, 1))
    th = np.stack([wl.init_theta(arch, seed=c) * 0.3 for c in range(chains)])
    hy = np.tile(wl.init_hyper(arch, lik), (chains, 1))
    eng = Engine(arch, lik, chains=chains, flags=flags)
    eng.set_data(X, Y)
    print(case, eng.sweep_info(), flush=True)
    if predict:
        out, mom = eng.predict(np.repeat(th, 3, axis=0), X, want_out=True, want_moments=True)
        print(eng.predict_kernel(), float(out.sum()))
    elif traj:
        t = eng.tensor(th).clone()
        s = eng.hmc_step(t, hy, 1, 0, 1e-4, 6)
        print(s.cpu().numpy())
    else:
        lp, g, _ = eng.logp_grad(th, hy)
        print(lp.cpu().numpy(), float(g.abs().max()))
    torch.cuda.synchronize()
if case == "wide2":
    run(wl.mlp_arch([784, 20, 20, 1], "dense", "relu", "sigmoid"), ("bernoulli",), 200)
elif case == "sweep_umma":
    run(wl.mlp_arch([784, 20, 20, 1], "dense", "relu", "sigmoid"), ("bernoulli",), 200, flags=8)
elif case == "predict_umma":
    run(wl.mlp_arch([1, 64, 64, 64, 1], "dense", "squareprelu"), ("gaussian", 0.1), 300, predict=True)
elif case == "traj_narrow":
    run(wl.mlp_arch([1, 10, 10, 10, 1], "denseGaussian", "tanh"), ("fixed", 0.1), 11, traj=True)
elif case == "train_umma64":
    run(wl.mlp_arch([1, 64, 64, 64, 1], "dense", "squareprelu"), ("gaussian", 0.1), 300, chains=2)
elif case == "train_umma128":
    run(wl.mlp_arch([32, 128, 128, 1], "dense", "relu"), ("gaussian", 0.1), 200)
