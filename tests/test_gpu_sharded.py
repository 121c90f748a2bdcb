"""Row-sharded sampling (one chain, training rows split across ranks, one ncclAllReduce of the partial
gradient + likelihood statistic per gradient evaluation inside libtbnn.so).

* world size 1: the communicator path (k_reduce_partials -> ncclAllReduce -> k_finalize from the reduced vector)
  must reproduce the plain path;
* world size 2 (skipped on a single-GPU box): two processes, each holding half of the rows, must reproduce the
  single-GPU log-posterior, gradient, trajectory end point and Metropolis decision, identically on both ranks.
"""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from tensorbnn_b200 import workloads as wl

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


CASES = {
    "c4s": (wl.mlp_arch([32, 128, 128, 128, 1], "dense", "relu"), ("gaussian", 0.1), 3000),      # tcgen05 sweep
    "c1b": (wl.mlp_arch([1, 10, 10, 10, 1], "dense", "relu"), ("gaussian", 0.1), 500),           # tile engine
    "c2s": (wl.mlp_arch([784, 20, 20, 1], "dense", "relu", "sigmoid"), ("bernoulli",), 1200),    # wide sweep
}


def _problem(name):
    arch, lik, N = CASES[name]
    rng = np.random.default_rng(11)
    D = arch[0][1]
    X = rng.random((N, D)) if D > 100 else rng.normal(size=(N, D))
    Y = (rng.random(N) > 0.5).astype(np.float64) if lik[0] == "bernoulli" else rng.normal(size=(N, 1))
    th = wl.init_theta(arch, seed=4) * (0.2 if D > 100 else 0.6)
    hy = wl.init_hyper(arch, lik)
    p0 = rng.normal(size=th.size)
    return arch, lik, X, Y, th, hy, p0


@pytest.mark.parametrize("name", sorted(CASES))
@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_world_size_one_communicator_equals_plain_path(name, dtype):
    from tensorbnn_b200.engine import Engine
    arch, lik, X, Y, th, hy, p0 = _problem(name)
    plain = Engine(arch, lik, dtype=dtype)
    plain.set_data(X, Y)
    comm = Engine(arch, lik, dtype=dtype)
    comm.comm_init(Engine.comm_unique_id(), 0, 1)
    comm.set_data(X, Y)
    tol = 2e-6 if dtype == torch.float32 else 1e-12
    a, b = plain.logp_grad(th[None], hy[None]), comm.logp_grad(th[None], hy[None])
    assert abs(a[0].item() - b[0].item()) <= tol * abs(a[0].item())
    assert rel(b[1].cpu().numpy(), a[1].cpu().numpy()) <= tol
    assert abs(a[2].item() - b[2].item()) <= tol * abs(a[2].item())
    eps, L = 1e-5, 6
    ta, tb = plain.trajectory(th[None], hy[None], p0[None], eps, L), comm.trajectory(th[None], hy[None], p0[None], eps, L)
    assert rel(tb[0].cpu().numpy(), ta[0].cpu().numpy()) <= 10 * tol and rel(tb[1].cpu().numpy(), ta[1].cpu().numpy()) <= 10 * tol
    sa = plain.hmc_step(plain.tensor(th[None]).clone(), hy[None], 3, 0, eps, L, momentum=p0[None], u=np.array([0.4]))
    sb = comm.hmc_step(comm.tensor(th[None]).clone(), hy[None], 3, 0, eps, L, momentum=p0[None], u=np.array([0.4]))
    sa, sb = sa.cpu().numpy()[0], sb.cpu().numpy()[0]
    # the log-accept ratio is a difference of log-posteriors: compare on their scale
    assert sa[2] == sb[2] and abs(sa[0] - sb[0]) <= max(1e-3, 4 * tol * abs(a[0].item()))


WORKER = r'''
import os, sys, json
sys.path.insert(0, %(root)r)
import numpy as np, torch, torch.distributed as dist
from tensorbnn_b200 import parallel
from tensorbnn_b200.engine import Engine
sys.path.insert(0, os.path.join(%(root)r, "tests"))
from test_gpu_sharded import _problem
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group(backend="nccl", device_id=torch.device("cuda", rank))
out = {}
for name in ("c4s", "c2s"):
    arch, lik, X, Y, th, hy, p0 = _problem(name)
    eng = Engine(arch, lik, dtype=torch.float32, device=rank)
    lo, hi = parallel.shard_range(len(X), rank, world)
    parallel.attach_row_sharding(eng, device=eng.dev)
    eng.set_data(X[lo:hi], Y[lo:hi])
    lp, g, st = eng.logp_grad(th[None], hy[None])
    t = eng.trajectory(th[None], hy[None], p0[None], 1e-5, 6)
    thc = eng.tensor(th[None]).clone()
    s = eng.hmc_step(thc, hy[None], 3, 0, 1e-5, 6, momentum=p0[None], u=np.array([0.4]))
    torch.cuda.synchronize()
    out[name] = {"logp": lp.item(), "grad": g.cpu().numpy()[0].tolist(), "stat": st.item(),
                 "traj": t[0].cpu().numpy()[0].tolist(), "stats": s.cpu().numpy()[0].tolist(),
                 "theta": thc.cpu().numpy()[0].tolist()}
json.dump(out, open(os.path.join(%(tmp)r, "rank%%d.json" %% rank), "w"))
dist.destroy_process_group()
'''


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_ranks_row_sharded_equal_single_gpu(tmp_path):
    import json
    from tensorbnn_b200.engine import Engine
    script = tmp_path / "worker.py"
    script.write_text(WORKER % {"root": ROOT, "tmp": str(tmp_path)})
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29613", str(script)],
                       env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    res = [json.load(open(tmp_path / ("rank%d.json" % k))) for k in range(2)]
    for name in ("c4s", "c2s"):
        arch, lik, X, Y, th, hy, p0 = _problem(name)
        one = Engine(arch, lik, dtype=torch.float32)
        one.set_data(X, Y)
        lp, g, st = one.logp_grad(th[None], hy[None])
        t = one.trajectory(th[None], hy[None], p0[None], 1e-5, 6)
        thc = one.tensor(th[None]).clone()
        s = one.hmc_step(thc, hy[None], 3, 0, 1e-5, 6, momentum=p0[None], u=np.array([0.4])).cpu().numpy()[0]
        a, b = res[0][name], res[1][name]
        # both ranks hold bit-identical results (the all-reduce returns the same bits everywhere): identical decisions
        assert a["logp"] == b["logp"] and a["grad"] == b["grad"] and a["stats"] == b["stats"] and a["theta"] == b["theta"]
        assert abs(a["logp"] - lp.item()) <= 5e-6 * abs(lp.item())
        assert rel(a["grad"], g.cpu().numpy()[0]) <= 5e-6
        assert abs(a["stat"] - st.item()) <= 5e-6 * abs(st.item())
        assert rel(a["traj"], t[0].cpu().numpy()[0]) <= 1e-4
        assert a["stats"][2] == s[2]
