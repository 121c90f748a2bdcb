#!/usr/bin/env python
"""bench.py -- chain-leapfrog-steps/sec of the B200 HMC hot path (BASELINE.json metric).

One bench "step" = one HMC transition of every chain of the workload through the C ABI (tbnn_hmc_step: momentum draw,
bootstrap gradient, L leapfrog steps each with ONE full-data log-posterior + gradient evaluation, Metropolis select).
value = chains * L * K / device time (max over ranks).

Default workload: **C3** = BASELINE.json configs[2], the batched-chain shape the metric is quoted on ("chain-
leapfrog-steps/sec at 1/2/4/8 B200"): 1,024 independent chains of a 1-64-64-64-1 SquarePrelu network on 4,096 rows,
GaussianLikelihood, L = 100.  It fits one GPU; with --gpus N the 1,024 chains are split 1024/N per GPU with no
communication ("strong" scaling: the total work is fixed).  Other shapes: --workload c1 | c2 | c2l | c4 | c5
(c4 at N > 1: training rows sharded, one NCCL all-reduce per gradient evaluation).

  python bench.py --gpus N --steps K --warmup W            # this framework
  python bench.py --impl reference ...                     # the reference's CPU path (oracle port, all host threads)
"""
import argparse
import glob
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np
import torch

from tensorbnn_b200 import workloads as wl

METRIC = "chain-leapfrog-steps/sec"
UNIT = "leapfrog-steps/s"
DEFAULT_WORKLOAD = "c3"


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback"


def ncu_traffic(kernel, workload):
    """dram read + write bytes per launch of the dominant kernel from the newest committed `ncu --set full` capture of
    the same workload (profiles/*_<kernel>_<workload>_full_metrics.csv, raw page), or None."""
    short = {"k_sweep_wide2": "wide2", "k_train_umma": "train_umma", "k_partial": "partial", "k_predict_umma": "predict_umma"}.get(kernel, kernel)
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_%s_%s_full_metrics.csv" % (short, workload))))
    if not files and workload == "c2":
        files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_%s_full_metrics.csv" % short)))
    if not files:
        return None
    try:
        import csv
        rows = list(csv.reader(open(files[-1])))
        hdr, units, vals = rows[0], rows[1], rows[2]
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        tot = 0.0
        for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            i = hdr.index(k)
            tot += float(vals[i].replace(",", "")) * scale.get(units[i], 1.0)
        return tot
    except Exception:
        return None


def make_workload(name, rank=0, world=1):
    """(cfg, theta [C_local, P], hyper [C_local, H], first global chain index).  Chains are split across ranks."""
    if name == "c2":
        cfg = wl.c2()
    elif name == "c2l":
        cfg = wl.c2(N=1048576)
        cfg["name"] = "C2-L"
    elif name == "c1":
        cfg = wl.c1("a")
    elif name == "c3":
        cfg = wl.c3(chains=1024)
    elif name == "c4":
        cfg = wl.c4()
    else:
        raise ValueError(name)
    C_total = cfg["chains"]
    if C_total >= world and name == "c3":
        lo, hi = C_total * rank // world, C_total * (rank + 1) // world     # strong scaling over chains
    else:
        lo, hi = rank * C_total, (rank + 1) * C_total                       # one replica of the chain set per rank
    arch, lik = cfg["arch"], cfg["lik"]
    theta = np.stack([wl.init_theta(arch, seed=1000 + c, slope=cfg.get("slope", 0.2)) * cfg.get("wscale", 1.0)
                      for c in range(lo, hi)])
    if name in ("c2", "c2l"):
        theta = theta * 0.2   # moderate logits at the random start (see tests/test_gpu_parity.py)
    hyper = np.tile(wl.init_hyper(arch, lik), (hi - lo, 1))
    cfg["chains_total"] = C_total if name == "c3" else C_total * world
    cfg["chains"] = hi - lo
    return cfg, theta, hyper, lo


def config_dict(cfg, L, eps, world, rows_sharded):
    """The `config` object of the JSON line; the reference arm prints the identical object."""
    arch, lik = cfg["arch"], cfg["lik"]
    name = cfg["name"]
    if rows_sharded:
        par = "one chain, training rows sharded %d ways, one NCCL all-reduce of the partial gradient per gradient evaluation" % world
    elif name == "C3":
        par = "%d chains split %d per GPU, no communication" % (cfg["chains_total"], cfg["chains_total"] // max(world, 1))
    else:
        par = "one chain per GPU, no communication" if world > 1 else "single chain"
    return {"workload": name, "rows": int(cfg["rows_total"]), "features": int(cfg["X"].shape[1]),
            "network": "-".join(str(d) for d in [arch[0][1]] + [l[2] for l in arch if l[0].startswith("dense")]),
            "activation": [l[0] for l in arch if not l[0].startswith("dense")][0],
            "likelihood": lik[0], "chains": int(cfg["chains_total"]), "leapfrog_per_step": int(L),
            "parallelism": par,
            "l2": "flushed between timed steps (256 MB write); within a trajectory the working set stays L2-resident"}


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            pass
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for n, v in zip(names, r[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def algorithmic_bytes_per_step(cfg, P, esz, chains):
    """N*(d0+dK)*s / C_sharing + 4*P*s per chain-leapfrog-step, times the chains of one launch."""
    N, D = cfg["X"].shape
    out = 1
    return N * (D + out) * esz + chains * 4 * P * esz


def flops_per_chain_step(cfg):
    dims = [l for l in cfg["arch"] if l[0] in ("dense", "denseGaussian")]
    F = sum(l[1] * l[2] for l in dims)
    return cfg["X"].shape[0] * (6 * F - 2 * dims[0][1] * dims[0][2])


def cpu_sample(name, cfg):
    """Bounded sample of the workload for the CPU arm: (X, Y, scale, text).  One CPU chain-leapfrog-step on the sample
    costs `scale` of one on the full workload (the cost is linear in rows); chains are identical in cost."""
    X, Y = cfg["X"], np.asarray(cfg["Y"])
    rows = {"c4": 262144, "c2l": 65536}.get(name)
    if rows and rows < len(X):
        frac = rows / float(len(X))
        return X[:rows], Y[:rows], frac, "a %d-row slice (1/%d of the rows; throughput scaled by %g, the cost is linear in rows)" % (
            rows, len(X) // rows, frac)
    if name == "c3":
        return X, Y, 1.0, "one of the %d chains on all %d rows (chains cost the same; they would run one after another)" % (
            cfg["chains_total"], len(X))
    return X, Y, 1.0, "the full workload"


def run_reference(args):
    """The reference's CPU implementation of the path.  TF/TFP cannot be installed offline, so this is the oracle port
    (reference op structure restated on torch-CPU: per-layer W@A+b, separate activation op, autograd backward,
    TFP-ordered leapfrog), all host threads.  One step = one L-step trajectory (the workload's own L) on a bounded
    sample of the workload; same metric / unit / config as the GPU arm."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import hmc
    cfg, theta, hyper, _ = make_workload(args.workload, 0, 1)
    cfg["rows_total"] = cfg["X"].shape[0]
    if args.workload == "c3":
        cfg["chains_total"] = 1024
    elif args.workload != "c4":
        cfg["chains_total"] = cfg["chains_total"] * max(args.gpus, 1)
    L = args.leapfrog or cfg["L"]
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    Xs, Ys, scale, text = cpu_sample(args.workload, cfg)
    f32 = lambda a: torch.tensor(np.asarray(a), dtype=torch.float32)
    vg = hmc.make_main_vg(cfg["arch"], cfg["lik"], f32(hyper[0]), f32(Xs), f32(Ys))
    g = torch.Generator().manual_seed(0)
    th = f32(theta[0])
    times = []
    for i in range(args.warmup + args.steps):
        p = torch.randn(th.shape, generator=g)
        t0 = time.perf_counter()
        hmc.leapfrog(vg, th, p, cfg["eps"], L if i >= args.warmup else min(L, 5))
        dt = time.perf_counter() - t0
        if i >= args.warmup:
            times.append(dt)
    total = sum(times)
    value = scale * L * len(times) / total
    sample = "per step one %d-leapfrog-step trajectory (bootstrap gradient included) of %s" % (L, text)
    rows_sharded = args.workload == "c4" and args.gpus > 1
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times),
            "higher_is_better": True, "scaling": "strong" if (args.workload in ("c3", "c4")) else "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_dict(cfg, L, cfg["eps"], max(args.gpus, 1), rows_sharded),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def run_predictor(args, restore_stdout):
    """C5: posterior-predictive sweep (predictor.predict), fused per-row mean / sd mode.  One step = S_step stored
    samples x M test rows on this rank's share of the samples; samples are split across ranks with no
    communication, the per-row (count, mean, M2) triples are merged once at the end."""
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local_rank))
    from tensorbnn_b200 import parallel
    from tensorbnn_b200.engine import Engine
    S_step, M = args.pred_samples, args.pred_rows
    cfg = wl.c5(M=M, S=S_step * world)
    lo, hi = parallel.shard_range(S_step * world, rank, world)
    eng = Engine(cfg["arch"], ("gaussian", 0.1), dtype=torch.float32, device=local_rank)
    samples_h = torch.tensor(cfg["samples"][lo:hi], dtype=torch.float32).contiguous().pin_memory()
    X_h = torch.tensor(cfg["X"], dtype=torch.float32).contiguous().pin_memory()
    samples, X = samples_h.cuda(), X_h.cuda()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=eng.dev)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        eng.predict(samples, X, want_out=False, want_moments=True)
    clocks = ClockSampler(local_rank)
    barrier()
    clocks.start()
    launches0 = eng.launches
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for i in range(args.steps):
        flush.zero_()
        evs[i][0].record()
        _, mom = eng.predict(samples, X, want_out=False, want_moments=True)
        evs[i][1].record()
    barrier()
    gpu_launches = eng.launches - launches0
    clk = clocks.stop()
    dev_ms = sum(a.elapsed_time(b) for a, b in evs)
    t = torch.tensor([dev_ms], dtype=torch.float64, device=eng.dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms = float(t.item())
    value = world * S_step * M * args.steps / (dev_ms * 1e-3)
    # end to end: samples and test rows from pinned host memory, merged mean / sd back to the host
    out_h = torch.empty(2, mom.shape[1], M, dtype=torch.float32).pin_memory()
    e2e_steps = max(2, args.steps // 2)
    merge = parallel.merge_moments if args.pred_merge == "gather" else parallel.merge_moments_reduce

    def e2e_step():
        s_d, x_d = samples_h.cuda(non_blocking=True), X_h.cuda(non_blocking=True)
        _, mom = eng.predict(s_d, x_d, want_out=False, want_moments=True)
        n, mu, m2 = merge(mom[0], mom[1], mom[2])
        out_h[0].copy_(mu, non_blocking=True)
        out_h[1].copy_((m2 / torch.clamp(n - 1, min=1)).sqrt(), non_blocking=True)
        torch.cuda.current_stream().synchronize()

    e2e_step()                                          # untimed: first use of the merge collective sets up its channels
    barrier()
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        flush.zero_()
        e2e_step()
    barrier()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device=eng.dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * S_step * M * e2e_steps / float(t.item())
    if rank == 0:
        peaks, peak_src = load_peaks()
        F = sum(l[1] * l[2] for l in cfg["arch"] if l[0].startswith("dense"))
        tflops = value * 2 * F / 1e12 / world
        line = {"metric": "predictor sample-rows/sec", "value": value, "unit": "sample-rows/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "tf32x3", "data": "synthetic",
                "config": {"workload": "C5 slice", "samples_per_gpu": S_step, "test_rows": M, "network": "1-64-64-64-1",
                           "mode": "fused per-row (count, mean, M2)", "parallelism": "samples split, one all-reduce of the per-row moments at the end",
                           "l2": "flushed between timed steps (256 MB write)"},
                "e2e": {"value": e2e_value, "unit": "sample-rows/s", "h2d_bytes_per_step": int(samples_h.numel() * 4 + X_h.numel() * 4),
                        "d2h_bytes_per_step": int(out_h.numel() * 4), "steps": e2e_steps,
                        "what": "per step: samples + test rows H2D, tbnn_predict (moments), merge, mean / sd D2H"},
                "gpu_launches": int(gpu_launches), "clocks": clk,
                "roofline": {"bound": "tensor", "achieved": tflops, "peak": peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"]) / 2 / 3,
                             "unit": "TFLOP/s", "frac": tflops / (peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"]) / 2 / 3),
                             "traffic": None, "kernel": eng.predict_kernel(), "peak_source": peak_src,
                             "note": "useful fp32-equivalent flops (2F per sample-row); peak = measured bf16 dense / 2 "
                                     "(tf32) / 3 (3xTF32 issues three MMAs per product)"},
                "cpu_baseline": None}
        restore_stdout()
        print(json.dumps(line))
        sys.stdout.flush()
    if dist is not None:
        dist.destroy_process_group()


def _stdout_to_stderr():
    """Libraries (NCCL with NCCL_DEBUG=VERSION) write to fd 1; the contract is ONE JSON line on stdout.  Everything
    goes to stderr until the returned function is called."""
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)

    def restore():
        sys.stdout.flush()
        os.dup2(saved, 1)
        os.close(saved)
    return restore


def tune_step_size(eng, th, hy, eps0, L, rank):
    """Per-chain step sizes at which the chains move: halve a chain's step size until a full L-step trajectory from the
    start state is accepted with probability >= 0.5 (the reference tunes the step size as well, with its paramAdapter)."""
    C = th.shape[0]
    eps = np.full(C, float(eps0))
    stats = torch.zeros(C, 4, dtype=th.dtype, device=th.device)
    Lt = L
    for it in range(24):
        trial = th.clone()
        eng.hmc_step(trial, hy, 77 + rank, 1000 + it, eps, Lt, stats=stats)
        acc = stats[:, 1].cpu().numpy()
        bad = ~(acc >= 0.5)
        if not bad.any():
            break
        eps[bad] *= 0.5
    return eps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD)
    ap.add_argument("--leapfrog", type=int, default=0, help="override L (0 = workload default)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-tune", action="store_true", help="keep the workload's nominal step size")
    ap.add_argument("--flags", type=int, default=0, help="TBNN_FLAG_* for the engine (64 = FFMA tile engine instead of tcgen05)")
    ap.add_argument("--pred-samples", type=int, default=512, help="c5: stored samples per GPU and step")
    ap.add_argument("--pred-rows", type=int, default=1048576, help="c5: test rows")
    ap.add_argument("--pred-merge", default="reduce", help="c5 end-to-end moment merge: reduce (one float64 all-reduce) | gather")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    restore_stdout = _stdout_to_stderr()
    if args.workload == "c5":
        return run_predictor(args, restore_stdout)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local_rank))

    from tensorbnn_b200.engine import Engine
    rows_sharded = args.workload == "c4" and world > 1
    cfg, theta, hyper, chain0 = make_workload(args.workload, 0 if rows_sharded else rank, 1 if rows_sharded else world)
    if rows_sharded:
        cfg["chains_total"] = 1
    cfg["rows_total"] = cfg["X"].shape[0]
    arch, lik, C = cfg["arch"], cfg["lik"], cfg["chains"]
    L = args.leapfrog or cfg["L"]
    dt = torch.float32
    eng = Engine(arch, lik, dtype=dt, chains=C, device=local_rank, flags=args.flags)
    seed = 1 if rows_sharded else 1 + rank
    if rows_sharded:
        # one chain, training rows split across the ranks, one NCCL all-reduce of the partial gradient per
        # gradient evaluation inside libtbnn.so; every rank replays the identical chain
        from tensorbnn_b200 import parallel
        lo, hi = parallel.shard_range(len(cfg["X"]), rank, world)
        cfg["X"], cfg["Y"] = cfg["X"][lo:hi], np.asarray(cfg["Y"])[lo:hi]
        parallel.attach_row_sharding(eng, device=eng.dev)
    Xh = torch.tensor(cfg["X"], dtype=dt).contiguous().pin_memory()
    Yh = torch.tensor(np.asarray(cfg["Y"]).reshape(len(cfg["X"]), -1), dtype=dt).contiguous().pin_memory()
    eng.set_data(Xh.cuda(), Yh.cuda())
    th = eng.tensor(theta).clone()
    hy = eng.tensor(hyper).clone()
    stats = torch.zeros(C, 4, dtype=dt, device=eng.dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=eng.dev)   # > 126 MB L2

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- step sizes at which the chains move (untimed), then warm-up
    eps = np.full(C, float(cfg["eps"])) if args.no_tune else tune_step_size(eng, th, hy, cfg["eps"], L, 0 if rows_sharded else rank)
    n_warm = 0
    while True:
        eng.hmc_step(th, hy, seed, n_warm, eps, L, stats=stats)
        n_warm += 1
        acc_w = stats[:, 1].cpu().numpy()
        if not args.no_tune:
            eps = np.where(acc_w >= 0.5, eps, eps * 0.5)   # the chain moved into a stiffer region: keep adapting (untimed)
        if n_warm >= args.warmup and (args.no_tune or np.mean(acc_w >= 0.5) >= 0.9 or n_warm >= args.warmup + 16):
            break
    # ---------------- device-resident throughput ("value")
    clocks = ClockSampler(local_rank)
    barrier()
    clocks.start()
    launches0 = eng.launches
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    acc_sum = 0.0
    for i in range(args.steps):
        flush.zero_()                                   # evict L2 between timed iterations
        evs[i][0].record()
        eng.hmc_step(th, hy, seed, n_warm + i, eps, L, stats=stats)
        evs[i][1].record()
    barrier()
    gpu_launches = eng.launches - launches0
    clk = clocks.stop()
    dev_ms = sum(a.elapsed_time(b) for a, b in evs)
    t = torch.tensor([dev_ms], dtype=torch.float64, device=eng.dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms = float(t.item())
    units = 1 if rows_sharded else cfg["chains_total"]          # chains advanced by all ranks together
    value = units * L * args.steps / (dev_ms * 1e-3)
    accept = float(stats[:, 1].mean().item())
    moved = float((stats[:, 2] > 0).float().mean().item())

    # ---------------- end to end through the C ABI with HOST buffers
    th_h = th.detach().cpu().contiguous().pin_memory()
    hy_h = torch.tensor(hyper, dtype=dt).contiguous().pin_memory()
    th_o = torch.empty_like(th_h).pin_memory()
    st_o = torch.zeros(C, 4, dtype=dt).pin_memory()
    h2d = Xh.numel() * 4 + Yh.numel() * 4 + th_h.numel() * 4 + hy_h.numel() * 4
    d2h = th_o.numel() * 4 + st_o.numel() * 4
    e2e_steps = max(3, args.steps // 2)

    def e2e_step(i):
        eng.set_data_host(Xh, Yh)                       # training set from pinned host memory
        th.copy_(th_h, non_blocking=True)
        hy.copy_(hy_h, non_blocking=True)
        eng.hmc_step(th, hy, seed, 10_000 + i, eps, L, stats=stats)
        th_o.copy_(th, non_blocking=True)
        st_o.copy_(stats, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        th_h.copy_(th_o)

    e2e_step(0)
    barrier()
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        flush.zero_()
        e2e_step(1 + i)
    barrier()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device=eng.dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = units * L * e2e_steps / float(t.item())
    eng.set_data(Xh.cuda(), Yh.cuda())

    # ---------------- roofline of the dominant kernel (row sweep), CUDA events inside the library
    peaks, peak_src = load_peaks()
    avg_ms, min_ms = eng.time_sweep(th, iters=20 if avg_guess(cfg, C) else 50)
    info = eng.sweep_info()
    abytes = algorithmic_bytes_per_step(cfg, eng.P, 4, C)
    hbm_gbs = abytes / (avg_ms * 1e-3) / 1e9
    tflops = C * flops_per_chain_step(cfg) / (avg_ms * 1e-3) / 1e12
    tensor_peak = peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"]) / 2 / 3
    ffma_peak = 148 * 128 * 2 * peaks.get("sm_max_mhz", 1965.0) * 1e6 / 1e12
    roof = {"kernel": info["kernel"], "sweep": info, "launch_ms": avg_ms, "launch_ms_min": min_ms,
            "share_of_step": avg_ms * (L + 1) * args.steps / dev_ms,
            "traffic": ncu_traffic(info["kernel"], args.workload), "peak_source": peak_src,
            "algorithmic_bytes_per_launch": abytes, "algorithmic_flops_per_launch": C * flops_per_chain_step(cfg),
            "hbm_gbs": hbm_gbs, "hbm_frac": hbm_gbs / peaks["hbm_gbs"], "fp32_equivalent_tflops": tflops,
            "ffma_frac": tflops / ffma_peak}
    if rows_sharded:
        ar_avg, ar_min = eng.time_allreduce(iters=30)
        roof["allreduce"] = {"ms": ar_avg, "ms_min": ar_min, "values": int(C * (eng.P + 4)),
                             "share_of_step": ar_avg * (L + 1) * args.steps / dev_ms,
                             "what": "local reduction of the CTA partials + ncclAllReduce per gradient evaluation (CUDA events)"}
    if info["kernel"] == "k_train_umma":
        roof.update({"bound": "tensor", "achieved": tflops, "peak": tensor_peak, "unit": "TFLOP/s", "frac": tflops / tensor_peak,
                     "note": "useful fp32-equivalent flops N(6F - 2 d0 d1) per chain-step (SURVEY 8d); peak = measured sustained "
                             "bf16 dense / 2 (tf32) / 3 (3xTF32 issues three MMAs per product).  HBM and FP32-FFMA fractions "
                             "alongside (hbm_frac, ffma_frac)."})
    else:
        roof.update({"bound": "hbm", "achieved": hbm_gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": hbm_gbs / peaks["hbm_gbs"],
                     "note": "algorithmic bytes N(d0+dK)s + 4Ps per launch (SURVEY 8d); small working sets are L2-resident across "
                             "the leapfrog steps of a trajectory; FP32 FFMA also bounds the wide-first-layer shape (ffma_frac); "
                             "C1 is latency-bound (one tile), no roofline fraction is meaningful there."})

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---------------- CPU baseline beside it (rank 0, N=1 only): the oracle port on a bounded sample
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        from oracle import cpu_baseline
        Xs, Ys, scale, text = cpu_sample(args.workload, cfg)
        Lc = min(L, 20)
        v, nsteps, cores = cpu_baseline.leapfrog_steps_per_second(
            arch, lik, Xs, Ys, theta[0], hyper[0], cfg["eps"], Lc, min_seconds=10.0)
        cpu = {"value": v * scale, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": "%d leapfrog steps in %d-step trajectories of %s; torch-CPU fp32 restatement of the reference's op "
                         "sequence (TensorFlow cannot be installed offline)" % (nsteps, Lc, text)}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
            "scaling": "strong" if (rows_sharded or args.workload == "c3") else "weak", "vs_baseline": None,
            "dtype": "f32 (hidden-layer GEMMs: 3xTF32 on tcgen05)" if info["kernel"] == "k_train_umma" else "f32",
            "data": "synthetic",
            "config": config_dict(cfg, L, cfg["eps"], world, rows_sharded),
            "step_size": {"nominal": float(cfg["eps"]), "median_used": float(np.median(eps)), "min_used": float(eps.min()),
                          "how": "per-chain halving until an L-step trajectory is accepted with probability >= 0.5, from the start state and "
                                 "again after every (untimed) warm-up transition", "warmup_transitions": int(n_warm)},
            "us_per_leapfrog": 1e3 * dev_ms / (args.steps * L),
            "accept_prob_last": accept, "accepted_fraction_last": moved,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "steps": e2e_steps,
                    "what": "per step: tbnn_set_data_host(X,Y) + theta/hyper H2D + tbnn_hmc_step + theta/stats D2H"},
            "gpu_launches": int(gpu_launches),
            "clocks": clk, "roofline": roof, "cpu_baseline": cpu}
    restore_stdout()
    print(json.dumps(line))
    sys.stdout.flush()
    if dist is not None:
        dist.destroy_process_group()


def avg_guess(cfg, C):
    """True for workloads whose sweep takes milliseconds (fewer timing iterations suffice)."""
    return cfg["X"].shape[0] * C >= (1 << 20)


if __name__ == "__main__":
    main()
