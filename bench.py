#!/usr/bin/env python
"""bench.py -- chain-leapfrog-steps/sec of the B200 HMC hot path (BASELINE.json metric).

One bench "step" = one HMC transition of the main chain through the C ABI (tbnn_hmc_step:
momentum draw, bootstrap gradient, L leapfrog steps each with ONE full-data log-posterior +
gradient evaluation, Metropolis select).  value = ranks * chains * L * K / device time.

Workload at N=1: C2 = BASELINE.json configs[1] (docs ClassificationExample shape): 9,600 x 784
synthetic 2-class data, 784-20-20-1 ReLU/ReLU/Sigmoid, DenseLayer (Cauchy) priors,
BernoulliLikelihood, one chain, fixed L = 500, eps = 1e-3.  N>1: one independent chain per GPU
(chains split with no communication -> "weak" scaling).

  python bench.py --gpus N --steps K --warmup W            # this framework
  python bench.py --impl reference ...                     # the reference-equivalent CPU port
"""
import argparse
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


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback"


def ncu_traffic(kernel, workload):
    """dram read + write bytes per launch of the dominant kernel from the committed `ncu --set full` capture of the
    same workload (profiles/), or None.  Cold-cache figure: in steady state C2's 30 MB set is L2-resident."""
    if workload != "c2" or kernel != "k_sweep_wide2":
        return None
    try:
        import csv
        rows = list(csv.reader(open(os.path.join(ROOT, "profiles", "r1e_wide2_full_metrics.csv"))))
        hdr, units, vals = rows[0], rows[1], rows[2]
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        tot = 0.0
        for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            i = hdr.index(k)
            tot += float(vals[i].replace(",", "")) * scale.get(units[i], 1.0)
        return tot
    except Exception:
        return None


def make_workload(name, rank=0):
    if name == "c2":
        cfg = wl.c2()
    elif name == "c2l":
        cfg = wl.c2(N=1048576 // 4)
        cfg["name"] = "C2-L/4"
    elif name == "c1":
        cfg = wl.c1("a")
    elif name == "c3":
        cfg = wl.c3(chains=148 * 2)
    elif name == "c4":
        cfg = wl.c4()
    else:
        raise ValueError(name)
    C = cfg["chains"]
    arch, lik = cfg["arch"], cfg["lik"]
    theta = np.stack([wl.init_theta(arch, seed=1000 * rank + c, slope=cfg.get("slope", 0.2)) * cfg.get("wscale", 1.0)
                      for c in range(C)])
    if name in ("c2", "c2l"):
        theta = theta * 0.2   # moderate logits at the random start (see tests/test_gpu_parity.py)
    hyper = np.tile(wl.init_hyper(arch, lik), (C, 1))
    return cfg, theta, hyper


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


def run_reference(args):
    """The reference's CPU implementation of the path: TF/TFP cannot be installed offline, so this is
    the oracle port (reference-equivalent torch-CPU restatement), all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import cpu_baseline
    cfg, theta, hyper = make_workload(args.workload)
    cores = os.cpu_count() or 1
    L_s = args.ref_leapfrog
    times = []
    torch.set_num_threads(cores)
    from oracle import hmc
    f32 = lambda a: torch.tensor(np.asarray(a), dtype=torch.float32)
    vg = hmc.make_main_vg(cfg["arch"], cfg["lik"], f32(hyper[0]), f32(cfg["X"]), f32(cfg["Y"]))
    g = torch.Generator().manual_seed(0)
    th = f32(theta[0])
    for i in range(args.warmup + args.steps):
        p = torch.randn(th.shape, generator=g)
        t0 = time.perf_counter()
        hmc.leapfrog(vg, th, p, cfg["eps"], L_s)
        dt = time.perf_counter() - t0
        if i >= args.warmup:
            times.append(dt)
    total = sum(times)
    value = L_s * len(times) / total
    sample = "%d-leapfrog-step trajectories of the full %s workload per step (bootstrap gradient included)" % (
        L_s, cfg["name"])
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": cfg["name"], "rows": int(cfg["X"].shape[0]), "features": int(cfg["X"].shape[1]),
                       "chains": 1, "leapfrog_per_step": L_s},
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
    barrier()
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        flush.zero_()
        s_d, x_d = samples_h.cuda(non_blocking=True), X_h.cuda(non_blocking=True)
        _, mom = eng.predict(s_d, x_d, want_out=False, want_moments=True)
        n, mu, m2 = parallel.merge_moments(mom[0], mom[1], mom[2])
        out_h[0].copy_(mu, non_blocking=True)
        out_h[1].copy_((m2 / torch.clamp(n - 1, min=1)).sqrt(), non_blocking=True)
        torch.cuda.current_stream().synchronize()
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
                           "mode": "fused per-row (count, mean, M2)", "parallelism": "samples split, one merge at the end",
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--workload", default="c2")
    ap.add_argument("--leapfrog", type=int, default=0, help="override L (0 = workload default)")
    ap.add_argument("--ref-leapfrog", type=int, default=20)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--pred-samples", type=int, default=512, help="c5: stored samples per GPU and step")
    ap.add_argument("--pred-rows", type=int, default=1048576, help="c5: test rows")
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
    cfg, theta, hyper = make_workload(args.workload, rank)
    arch, lik, C = cfg["arch"], cfg["lik"], cfg["chains"]
    L = args.leapfrog or cfg["L"]
    eps = cfg["eps"]
    dt = torch.float32
    eng = Engine(arch, lik, dtype=dt, chains=C, device=local_rank)
    rows_sharded = args.workload == "c4" and world > 1
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

    # ---------------- device-resident throughput ("value")
    for i in range(args.warmup):
        eng.hmc_step(th, hy, 1 + (0 if rows_sharded else rank), i, eps, L, stats=stats)
    clocks = ClockSampler(local_rank)
    barrier()
    clocks.start()
    launches0 = eng.launches
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for i in range(args.steps):
        flush.zero_()                                   # evict L2 between timed iterations
        evs[i][0].record()
        eng.hmc_step(th, hy, 1 + (0 if rows_sharded else rank), args.warmup + i, eps, L, stats=stats)
        evs[i][1].record()
    barrier()
    gpu_launches = eng.launches - launches0
    clk = clocks.stop()
    dev_ms = sum(a.elapsed_time(b) for a, b in evs)
    t = torch.tensor([dev_ms], dtype=torch.float64, device=eng.dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms = float(t.item())
    units = 1 if rows_sharded else world          # row sharding: ONE chain advanced by all ranks together
    value = units * C * L * args.steps / (dev_ms * 1e-3)
    accept = float(stats[:, 1].mean().item())

    # ---------------- end to end through the C ABI with HOST buffers
    th_h = torch.tensor(theta, dtype=dt).contiguous().pin_memory()
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
        eng.hmc_step(th, hy, 1 + (0 if rows_sharded else rank), 10_000 + i, eps, L, stats=stats)
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
    e2e_value = units * C * L * e2e_steps / float(t.item())
    eng.set_data(Xh.cuda(), Yh.cuda())

    # ---------------- roofline of the dominant kernel (row sweep), CUDA events inside the library
    peaks, peak_src = load_peaks()
    avg_ms, min_ms = eng.time_sweep(th, iters=50)
    abytes = algorithmic_bytes_per_step(cfg, eng.P, 4, C)
    achieved = abytes / (avg_ms * 1e-3) / 1e9
    roof = {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
            "frac": achieved / peaks["hbm_gbs"], "traffic": ncu_traffic(eng.sweep_info()["kernel"], args.workload),
            "kernel": eng.sweep_info()["kernel"], "sweep": eng.sweep_info(),
            "launch_ms": avg_ms, "launch_ms_min": min_ms, "algorithmic_bytes_per_launch": abytes,
            "peak_source": peak_src,
            "fp32_tflops": C * flops_per_chain_step(cfg) / (avg_ms * 1e-3) / 1e12,
            "note": "C2's 30 MB working set is L2-resident across the leapfrog steps of a trajectory "
                    "(the real access pattern); FP32 FFMA also bounds this shape (SURVEY 8d)"}

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---------------- CPU baseline beside it (rank 0, N=1 only)
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        from oracle import cpu_baseline
        v, nsteps, cores = cpu_baseline.leapfrog_steps_per_second(
            arch, lik, cfg["X"], cfg["Y"], theta[0], hyper[0], eps, 20, min_seconds=10.0)
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": "%d leapfrog steps of the full %s workload in 20-step trajectories, torch-CPU fp32 "
                         "restatement of the reference (TensorFlow unavailable offline)" % (nsteps, cfg["name"])}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
            "scaling": "strong" if rows_sharded else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": cfg["name"], "rows": int(cfg["X"].shape[0]), "features": int(cfg["X"].shape[1]),
                       "network": "-".join(str(d) for d in [arch[0][1]] + [l[2] for l in arch if l[0].startswith("dense")]),
                       "likelihood": lik[0], "chains_per_gpu": C, "leapfrog_per_step": L, "step_size": eps,
                       "parallelism": ("rows sharded, one NCCL all-reduce per gradient evaluation" if rows_sharded else
                                       "chains split, no communication" if world > 1 else "single chain"),
                       "l2": "flushed between timed steps (256 MB write); within a trajectory the data is L2-resident"},
            "us_per_leapfrog": 1e3 * dev_ms / (args.steps * L),
            "accept_prob_last": accept,
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


if __name__ == "__main__":
    main()
