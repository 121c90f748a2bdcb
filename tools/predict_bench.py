"""Times the predictor sweep on a C5 slice (developer aid; GPU only)."""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from tensorbnn_b200 import _lib, workloads as wl
from tensorbnn_b200.engine import Engine

S = int(sys.argv[1]) if len(sys.argv) > 1 else 256
M = int(sys.argv[2]) if len(sys.argv) > 2 else 1048576
cfg = wl.c5(M=M, S=S)
arch = cfg["arch"]
for flags in (0, _lib.FLAG_NO_UMMA):
    eng = Engine(arch, ("gaussian", 0.1), flags=flags)
    samples = eng.tensor(cfg["samples"])
    X = eng.tensor(cfg["X"])
    for rep in range(3):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _, mom = eng.predict(samples, X, want_out=False, want_moments=True)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
    print(eng.predict_kernel(), "S", S, "M", M, "ms", ms, "sample-rows/s %.3e" % (S * M / (ms * 1e-3)),
          "TFLOP/s %.1f" % (S * M * 2 * 8320 / (ms * 1e-3) / 1e12), "mean[0..3]", mom[1, 0, :3].tolist())
