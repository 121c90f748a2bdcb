"""Reference-equivalent CPU path, timed: the restatement of oracle/targets.py + oracle/hmc.py run
in fp32 on torch-CPU with the reference's op structure (per-layer W@A+b, separate activation op,
autograd backward, TFP-ordered leapfrog), all host threads.  TensorFlow / TFP cannot be installed
offline, so this port IS the CPU baseline (kind "port").

TEST / BENCH INFRASTRUCTURE (see oracle/__init__.py): used only by bench.py's cpu_baseline leg and
`bench.py --impl reference`.
"""
import os
import time

import numpy as np
import torch

from . import hmc, targets


def leapfrog_steps_per_second(arch, lik, X, Y, theta, hyper, eps, L, min_seconds=8.0, warmup=1,
                              threads=None):
    """Runs L-step trajectories (bootstrap gradient + L leapfrog steps, as one epoch of the main chain
    does) until ``min_seconds`` of wall clock have elapsed; returns (steps/s, steps timed, cores)."""
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    f32 = lambda a: torch.tensor(np.asarray(a), dtype=torch.float32)
    Xt, Yt, th, hy = f32(X), f32(Y), f32(theta), f32(hyper)
    vg = hmc.make_main_vg(arch, lik, hy, Xt, Yt)
    g = torch.Generator().manual_seed(0)
    for _ in range(warmup):
        p = torch.randn(th.shape, generator=g)
        hmc.leapfrog(vg, th, p, eps, max(1, min(L, 3)))
    steps, t0 = 0, time.perf_counter()
    while True:
        p = torch.randn(th.shape, generator=g)
        hmc.leapfrog(vg, th, p, eps, L)
        steps += L
        dt = time.perf_counter() - t0
        if dt >= min_seconds:
            break
    return steps / dt, steps, threads


def predict_rows_per_second(arch, samples, X, min_seconds=5.0, threads=None):
    """predictor.predict on the CPU: Python loop over samples (predictor.py:143-153)."""
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    f32 = lambda a: torch.tensor(np.asarray(a), dtype=torch.float32)
    Xt = f32(X)
    S = [targets.unflatten_theta(arch, f32(s)) for s in samples]
    done, t0 = 0, time.perf_counter()
    while True:
        for th in S:
            targets.forward(arch, th, Xt)
            done += Xt.shape[0]
        dt = time.perf_counter() - t0
        if dt >= min_seconds:
            break
    return done / dt, done, threads
