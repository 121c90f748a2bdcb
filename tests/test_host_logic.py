"""CPU tests of the oracle's sampler pieces and of the product's host-side logic (no GPU):
TFP-ordered leapfrog, dual averaging, adapter bookkeeping, sample-store format, dtype handling,
Philox known answers, and world_size-2 gloo runs of the multi-GPU plumbing."""
import math
import os
import random

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import adapter as oad
from oracle import fileformat, hmc
import philox_ref


# ---------------------------------------------------------------------------- oracle HMC
def _std_normal_vg(theta):
    return -0.5 * torch.sum(theta * theta), -theta


def test_leapfrog_standard_normal_is_a_rotation():
    """on a standard normal target the leapfrog map is a rotation to O(eps^2)."""
    th0 = torch.tensor([1.0, -0.5, 0.25], dtype=torch.float64)
    p0 = torch.tensor([0.3, 0.8, -1.1], dtype=torch.float64)
    eps, L = 1e-3, 500
    th, p, lp, g = hmc.leapfrog(_std_normal_vg, th0, p0, eps, L)
    t = eps * L
    assert torch.allclose(th, th0 * math.cos(t) + p0 * math.sin(t), atol=1e-6)
    assert torch.allclose(p, p0 * math.cos(t) - th0 * math.sin(t), atol=1e-6)
    h0 = 0.5 * (th0 @ th0 + p0 @ p0)
    h1 = 0.5 * (th @ th + p @ p)
    assert abs(h1 - h0) < 1e-6                                   # O(eps^2) energy error
    # reversibility
    thb, pb, _, _ = hmc.leapfrog(_std_normal_vg, th, -p, eps, L)
    assert torch.allclose(thb, th0, atol=1e-12) and torch.allclose(-pb, p0, atol=1e-12)


def test_hmc_step_accept_reject_and_safe_sum():
    th0 = torch.tensor([0.2, -0.1], dtype=torch.float64)
    p0 = torch.tensor([1.0, 0.5], dtype=torch.float64)
    new, lar, prob, acc, prop, _ = hmc.hmc_step(_std_normal_vg, th0, p0, 1e-12, 0.1, 5)
    assert acc and torch.equal(new, prop) and 0 < prob.item() <= 1
    new, lar, prob, acc, prop, _ = hmc.hmc_step(_std_normal_vg, th0, p0, 1 - 1e-12, 0.1, 5)
    assert (math.log(1 - 1e-12) < lar.item()) == acc
    nan = torch.tensor(float("nan"), dtype=torch.float64)
    assert hmc.log_accept_ratio(torch.tensor(0.0, dtype=torch.float64), nan, p0, p0).item() == -math.inf
    inf = torch.tensor(float("inf"), dtype=torch.float64)
    assert hmc.log_accept_ratio(inf, inf, p0, p0).item() == -math.inf     # +inf and -inf present


def test_dual_averaging_constants():
    """network.py:457-469 by hand for the first epoch: m=1, t0=10, gamma=.4, kappa=.75, target=.95
    (gamma and mu are tf.cast(python float) values, i.e. float32-rounded: oracle/tfconst.py, Q14)."""
    h, leb, st = hmc.dual_averaging(0.0, 0.5, 0.0, 0.0, 1e-2, 1e-2, 1000)
    h_ref = (1 / 11) * (0.95 - 0.5)
    le = float(np.log(np.float32(100 * 1e-2))) - h_ref * 1.0 / float(np.float32(0.4))
    assert abs(h - h_ref) < 1e-15 and abs(leb - le) < 1e-15 and abs(st - math.exp(le)) < 1e-15
    # frozen after 0.8 * burnin
    _, _, st2 = hmc.dual_averaging(900.0, 0.5, 0.1, -3.0, 7e-3, 1e-2, 1000)
    assert st2 == 7e-3


# ---------------------------------------------------------------------------- adapter
def _drive(adapter_update, n, seed=0):
    nrng = np.random.default_rng(seed)
    st = [nrng.normal(size=(4, 3)), nrng.normal(size=(4, 1))]
    out = []
    for i in range(n):
        st = [s + 0.02 * nrng.normal(size=s.shape) for s in st]
        out.append(adapter_update(st))
    return out


def test_product_adapter_matches_oracle_bookkeeping():
    """Same injected randomness, random-exploration phase only (the grid search needs the GPU):
    the product's host logic must reproduce the oracle's (e, L) sequence exactly."""
    from tensorbnn_b200.paramAdapter import paramAdapter
    kw = dict(e1=1e-3, L1=200, el=1e-4, eu=1e-2, eNumber=20, Ll=100, Lu=1000, lStep=10, m=5, k=20,
              randomSteps=10 ** 6)
    a = oad.OracleAdapter(rng=random.Random(7), **kw)
    b = paramAdapter(rng=random.Random(7), **kw)
    b.verbose = False
    ra = _drive(lambda s: a.update(s), 120)
    rb = _drive(lambda s: b.update(state=[torch.tensor(x) for x in s]), 120)
    assert [(float(e), int(L)) for e, L in ra] == [(float(e), int(L)) for e, L in rb]
    assert len(a.previousGamma) == len(b.previousGamma) > 3
    np.testing.assert_allclose(a.K, b.K, rtol=1e-6)
    np.testing.assert_allclose(a.inverseR, b.inverseR, rtol=1e-4, atol=1e-6)
    assert abs(a.rootbeta - b.rootbeta) < 1e-12 and a.p == b.p


def test_product_adapter_sjd_path_equals_state_path():
    from tensorbnn_b200.paramAdapter import paramAdapter
    kw = dict(e1=1e-3, L1=200, el=1e-4, eu=1e-2, eNumber=20, Ll=100, Lu=1000, lStep=10, m=5, k=20,
              randomSteps=10 ** 6)
    a = paramAdapter(rng=random.Random(3), **kw)
    b = paramAdapter(rng=random.Random(3), **kw)
    a.verbose = b.verbose = False
    nrng = np.random.default_rng(1)
    st = [torch.tensor(nrng.normal(size=(4, 3))), torch.tensor(nrng.normal(size=(4,)))]
    prev = None
    for i in range(60):
        st = [s + 0.02 * torch.tensor(nrng.normal(size=tuple(s.shape))) for s in st]
        ra = a.update(state=st)
        sjd = 0.0 if prev is None else sum(float(torch.sum((x.float() - y.float()) ** 2)) for x, y in zip(st, prev))
        rb = b.update(sjd=sjd)
        prev = st
        assert float(ra[0]) == float(rb[0]) and int(ra[1]) == int(rb[1])


def test_adapter_reset_after_strikes():
    a = oad.OracleAdapter(1e-3, 200, 1e-4, 1e-2, 10, 100, 300, 10, m=2, k=400, randomSteps=0,
                          rng=random.Random(0))
    st = [np.zeros((2, 2))]
    el0 = float(a.el)
    for i in range(200):
        a.update(st)                                    # zero movement every epoch
    assert float(a.el) < el0                            # range was halved at least once


def test_oracle_grid_search_first_maximum():
    a = oad.OracleAdapter(1e-3, 200, 1e-4, 1e-2, 6, 100, 140, 10, m=2, k=4, randomSteps=0)
    prev = [(5.05e-3, 120.0)]
    e, L = a.gridSearch(prev, np.zeros((1, 1), np.float32), np.float32(1), np.ones((1, 1), np.float32), 1.0,
                        0.0, a.el, a.eu, a.sigma)
    assert float(e) == float(a.eGrid[0]) and float(L) == float(a.lGrid[0])


# ---------------------------------------------------------------------------- sample store
def test_file_roundtrip_and_lagging_summary(tmp_path):
    """the 6001/1000/10/50 schedule => summary '500 10 <numMatrices>', 50 samples in each
    of files 0..9; the product's loader reads what the restated writer wrote."""
    shapes = [(3, 2), (3, 1), (3,)]
    rng = np.random.default_rng(0)
    store = {}

    def state_fn(it):
        store[it] = [rng.normal(size=s) for s in shapes]
        return store[it]

    hyper_fn = lambda it: np.arange(7, dtype=np.float64) + it
    folder = str(tmp_path / "run")
    saved = fileformat.write_run(folder, ["dense", "squareprelu"], 6001, 1000, 10, 50, state_fn, hyper_fn)
    assert len(saved) == 500 and saved[0] == 1010 and saved[-1] == 6000
    last = open(os.path.join(folder, "summary.txt")).read().split("\n")
    assert last[-2] == "500 10 3" and last[-1] == "7"
    assert last[0] == "3 2" and last[1] == "3 1" and last[2] == "3"
    mats, hy = fileformat.load_networks(folder + "/")
    assert mats[0].shape == (500, 3, 2) and mats[2].shape == (500, 3, 1) and hy.shape == (500, 7)
    np.testing.assert_allclose(mats[0][0], store[1010][0].astype(np.float32), rtol=1e-6)
    np.testing.assert_allclose(mats[2][499][:, 0], store[6000][2].astype(np.float32), rtol=1e-6)
    np.testing.assert_allclose(hy[0], np.arange(7) + 1010)
    # the product's predictor loader parses the same directory identically (no GPU needed for loading)
    from tensorbnn_b200.predictor import predictor
    pr = predictor(folder + "/", np.float32)
    assert pr.numNetworks == 500 and pr.numMatrices == 3
    np.testing.assert_allclose(pr.matrices[0].numpy(), mats[0].astype(np.float32), rtol=1e-6)
    np.testing.assert_allclose(np.array(pr.hypers), hy)
    assert [l.name for l in pr.layers] == ["dense", "squareprelu"]
    assert pr._arch == [("dense", 2, 3), ("squareprelu", 3)]
    means, sds = pr.parameterStatistics()
    assert means[0].shape == (3, 2)


def test_binary_side_cars_load_identically(tmp_path):
    """`<name>.f32` side-cars (raw float32, same order) give the predictor exactly what parsing the
    reference's text files gives; an incomplete side-car is ignored."""
    from tensorbnn_b200.predictor import predictor
    shapes = [(4, 3), (4, 1), (4,)]
    rng = np.random.default_rng(1)
    folder = str(tmp_path / "run3")
    fileformat.write_run(folder, ["dense", "prelu"], 1000 + 60, 1000, 2, 10,
                         lambda it: [rng.normal(size=s) for s in shapes], lambda it: rng.normal(size=5))
    text = predictor(folder + "/", np.float32)
    for name in os.listdir(folder):
        if name.endswith(".txt") and name not in ("summary.txt", "architecture.txt"):
            np.loadtxt(os.path.join(folder, name), dtype=np.float32).astype("<f4").tofile(
                os.path.join(folder, name[:-4] + ".f32"))
    binary = predictor(folder + "/", np.float32)
    assert binary.numNetworks == text.numNetworks == 20      # lagging summary (Q9): the last file is not listed yet
    for a, b in zip(binary.matrices, text.matrices):
        assert torch.equal(a, b)
    assert np.array_equal(np.array(binary.hypers), np.array(text.hypers))
    # a truncated side-car must not be trusted
    with open(os.path.join(folder, "0.0.f32"), "wb") as fh:
        fh.write(b"\0" * 8)
    again = predictor(folder + "/", np.float32)
    assert torch.equal(again.matrices[0], text.matrices[0])


def test_partial_last_file_is_invisible(tmp_path):
    """Q9: epochs that do not end on a rollover leave the last partial file out of the summary."""
    shapes = [(2, 2), (2, 1)]
    folder = str(tmp_path / "run2")
    saved = fileformat.write_run(folder, ["dense"], 1000 + 75, 1000, 1, 50,
                                 lambda it: [np.full(s, float(it)) for s in shapes], lambda it: np.zeros(4))
    assert len(saved) == 75
    mats, _ = fileformat.load_networks(folder + "/")
    assert mats[0].shape[0] == 50


# ---------------------------------------------------------------------------- host helpers
def test_dtype_vocabulary():
    from tensorbnn_b200.layer import to_torch_dtype

    class FakeTF(object):
        name = "float32"

    assert to_torch_dtype(np.float32) == torch.float32
    assert to_torch_dtype(torch.float64) == torch.float64
    assert to_torch_dtype("float64") == torch.float64
    assert to_torch_dtype(FakeTF()) == torch.float32
    with pytest.raises(ValueError):
        to_torch_dtype(np.int32)


def test_network_bookkeeping_matches_reference_layout():
    """network.add ordering (network.py:173-191) and setupMCMC defaults / positional order (:193-198)."""
    from tensorbnn_b200.activationFunctions import Relu, SquarePrelu
    from tensorbnn_b200.layer import DenseLayer, GaussianDenseLayer
    from tensorbnn_b200.network import network
    x = np.linspace(-1, 1, 7)
    net = network(np.float32, 1, x, x ** 2, x, x ** 2, 0.0, 1.0)     # docs pass mean, sd: ignored
    net.add(DenseLayer(1, 4, seed=1))
    net.add(SquarePrelu(4, alpha=0.1 ** 0.5))
    net.add(GaussianDenseLayer(4, 1, seed=2))
    net.add(Relu())
    assert [tuple(s.shape) for s in net.states] == [(4, 1), (4, 1), (4,), (1, 4), (1, 1)]
    assert len(net.hyperStates) == 4 + 2 + 4
    assert [float(h) for h in net.hyperStates[:4]] == pytest.approx([0, 0.5 ** 0.5, 0, 0.5 ** 0.5])
    assert [float(h) for h in net.hyperStates[4:6]] == pytest.approx([0.0, 0.3])
    assert net.arch_spec() == [("dense", 1, 4), ("squareprelu", 4), ("denseGaussian", 4, 1), ("relu",)]
    net.setupMCMC(0.001, 0.0005, 0.002, 100, 500, 100, 2000, 1, 1e-5, 30, 50, 2, 2)   # docs positional call
    # tf.cast(stepSizeStart, dtype) passes the python float through float32 (network.py:237, Q14)
    assert net.step_size == float(np.float32(0.001)) and net.leapfrog == 500 and net.hyperLeapfrog == 30 and net.burnin == 50
    assert net.adapt.eNumber == 100 and net.adapt.lNumber == 1901 and net.adapt.k == 25 and net.adapt.m == 2
    assert abs(net.mu - math.log(100 * 1e-5)) < 1e-12


def test_layer_init_statistics():
    from tensorbnn_b200.layer import DenseLayer
    l = DenseLayer(50, 400, seed=3)
    w, b = l.parameters
    assert tuple(w.shape) == (400, 50) and tuple(b.shape) == (400, 1)
    assert abs(float(w.std()) - (2 / 400) ** 0.5) < 0.01 * (2 / 400) ** 0.5 * 3
    assert l.numTensors == 2 and l.numHyperTensors == 4 and l.name == "dense"


def test_philox_known_answers():
    """Random123 known-answer vectors for Philox4x32-10 pin the stream definition."""
    assert philox_ref.philox4x32_10((0, 0, 0, 0), 0, 0) == (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)
    assert philox_ref.philox4x32_10((0xffffffff,) * 4, 0xffffffff, 0xffffffff) == \
        (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)
    assert philox_ref.philox4x32_10((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), 0xa4093822, 0x299f31d0) == \
        (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)


# ---------------------------------------------------------------------------- world_size 2 (gloo)
def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from tensorbnn_b200 import parallel
    # sample-sharded predictor merge: each rank owns a block of samples of the same rows
    rng = np.random.default_rng(0)
    data = rng.normal(size=(11, 5))                                   # 11 samples x 5 rows
    lo, hi = parallel.shard_range(11, rank, world)
    mine = torch.tensor(data[lo:hi])
    cnt = torch.full((5,), float(hi - lo), dtype=torch.float64)
    mean = mine.mean(dim=0)
    m2 = ((mine - mean) ** 2).sum(dim=0)
    n, mu, s = parallel.merge_moments(cnt, mean, m2)
    ok = bool(torch.allclose(mu, torch.tensor(data.mean(axis=0))) and
              torch.allclose(s / n, torch.tensor(data.var(axis=0))) and torch.all(n == 11))
    # the one-collective variant of the same merge (float64 sums, one all-reduce)
    n2, mu2, s2 = parallel.merge_moments_reduce(cnt.float(), mean.float(), m2.float())
    ok = ok and bool(torch.allclose(mu2.double(), mu, rtol=1e-6) and torch.allclose(s2.double(), s, rtol=1e-5)
                     and torch.all(n2 == 11))
    # chain split of the default bench workload: every chain exactly once (1,024 chains over the ranks)
    owned = [(1024 * r // world, 1024 * (r + 1) // world) for r in range(world)]
    ok = ok and owned[0][0] == 0 and owned[-1][1] == 1024 and all(owned[i][1] == owned[i + 1][0] for i in range(world - 1))
    # unique-id broadcast
    payload = bytes(range(128)) if rank == 0 else bytes(128)
    got = parallel.broadcast_bytes(payload, 0)
    ok = ok and got == bytes(range(128))
    # row sharding covers every row exactly once
    blocks = [parallel.shard_range(1000003, r, world) for r in range(world)]
    ok = ok and blocks[0][0] == 0 and blocks[-1][1] == 1000003 and all(
        blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
    q.put((rank, ok))
    dist.destroy_process_group()


def test_two_process_gloo_plumbing():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + random.randint(0, 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]
