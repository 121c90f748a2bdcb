"""Saved-sample text layout restated: the writer schedule of network.py:545-663
and the reader of predictor.py:43-130.

TEST INFRASTRUCTURE (see oracle/__init__.py; parity unpinned).
"""
import os

import numpy as np


def write_run(folder, layer_names, epochs, burnin, samplingStep, networksPerFile,
              state_fn, hyper_fn):
    """Replays the train() file logic.  ``state_fn(iter_)`` returns the list of
    state arrays after epoch ``iter_`` (1-based), ``hyper_fn(iter_)`` the flat
    hyper scalars.  Returns the list of epochs at which a sample was saved."""
    os.makedirs(folder, exist_ok=True)
    states0 = state_fn(0)
    n_states = len(states0)
    files = [open(os.path.join(folder, "%d.0.txt" % n), "wb") for n in range(n_states)]
    files.append(open(os.path.join(folder, "hypers0.txt"), "wb"))
    with open(os.path.join(folder, "architecture.txt"), "wb") as f:           # :557-559
        for name in layer_names:
            f.write((name + "\n").encode("utf-8"))
    saved = []
    iter_ = 0
    while iter_ < epochs:                                                      # :567
        iter_ += 1
        states = state_fn(iter_)
        hypers = np.asarray(hyper_fn(iter_)).reshape(-1)
        indexShift = iter_ - burnin - 1                                        # :610
        indexInterval = networksPerFile * samplingStep
        if iter_ > burnin and indexShift % indexInterval == 0:                 # :612
            for fh in files:
                fh.close()
            fidx = int((iter_ - burnin) // (networksPerFile * samplingStep))
            files = [open(os.path.join(folder, "%d.%d.txt" % (n, fidx)), "wb") for n in range(n_states)]
            files.append(open(os.path.join(folder, "hypers%d.txt" % fidx), "wb"))
            with open(os.path.join(folder, "summary.txt"), "wb") as fh:        # :629-646
                for s in states:
                    fh.write((" ".join(str(d) for d in np.shape(s)).strip() + "\n").encode("utf-8"))
                numNetworks = indexShift // samplingStep
                numFiles = numNetworks // networksPerFile
                if numNetworks % networksPerFile != 0:
                    numFiles += 1
                fh.write(("%d %d %d\n" % (numNetworks, numFiles, n_states)).encode("utf-8"))
                fh.write(str(hypers.size).encode("utf-8"))
        if iter_ > burnin and iter_ % samplingStep == 0:                       # :648
            for n in range(n_states):
                np.savetxt(files[n], np.asarray(states[n]))
            np.savetxt(files[-1], hypers.reshape(-1, 1))
            saved.append(iter_)
    for fh in files:
        fh.close()
    return saved


def load_networks(directoryPath):
    """predictor.py:43-113.  ``directoryPath`` must end with '/'.  Returns
    (matrices [list of [S,d1,d2] float32-parsed arrays], hypers [S,H])."""
    summary = []
    with open(directoryPath + "summary.txt", "r") as fh:
        for line in fh:
            summary.append(line.split())
    numNetworks = int(summary[-2][0])
    numFiles = int(summary[-2][1])
    numMatrices = int(summary[-2][2])
    numHypers = int(summary[-1][0])
    numNetworks //= numFiles
    matrices = []
    for n in range(numMatrices):
        d1 = int(summary[n][0])
        d2 = int(summary[n][1]) if len(summary[n]) == 2 else 1
        out = np.zeros((numNetworks * numFiles, d1, d2))
        for m in range(numFiles):
            w = np.loadtxt(directoryPath + "%d.%d.txt" % (n, m), dtype=np.float32, ndmin=2)
            for k in range(numNetworks):
                out[m * numNetworks + k] = w[d1 * k:d1 * (k + 1), :d2]
        matrices.append(out)
    hypers = []
    if numHypers > 0:
        for m in range(numFiles):
            w = np.loadtxt(directoryPath + "hypers%d.txt" % m, dtype=np.float32, ndmin=1)
            for k in range(numNetworks):
                hypers.append(w[numHypers * k:numHypers * (k + 1)])
    return matrices, np.array(hypers)
