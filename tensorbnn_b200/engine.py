"""Thin torch-tensor wrapper over the C ABI (include/tbnn.h).  torch is used only for
device memory, streams and torch.distributed plumbing; every computation happens in
libtbnn.so.  All tensors are [C, ...] with C = number of batched chains."""
import ctypes as C

import numpy as np
import torch

from . import _lib


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _stream(device=None):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


class Engine(object):
    def __init__(self, arch, lik, dtype=torch.float32, chains=1, device=None, flags=0):
        if not torch.cuda.is_available():
            raise RuntimeError("tensorbnn_b200 needs a CUDA device (no CPU fallback)")
        self.lib = _lib.load()
        self.arch, self.lik = list(arch), tuple(lik)
        self.dtype = dtype
        self.chains = int(chains)
        self.device = torch.cuda.current_device() if device is None else int(device)
        self.dev = torch.device("cuda", self.device)
        code = _lib.F32 if dtype == torch.float32 else _lib.F64
        if dtype not in (torch.float32, torch.float64):
            raise ValueError("dtype must be float32 or float64")
        desc, self._keep = _lib.make_desc(self.arch, self.lik, code, self.chains, self.device, flags)
        h = C.c_void_p()
        self._call(self.lib.tbnn_create, (C.byref(desc), C.byref(h)))
        self.h = h
        self.P = self.lib.tbnn_num_params(h)
        self.H = self.lib.tbnn_num_hypers(h)
        self.N = 0
        self._data = None

    def __del__(self):
        try:
            if getattr(self, "h", None):
                self.lib.tbnn_destroy(self.h)
                self.h = None
        except Exception:
            pass

    # ------------------------------------------------------------------ helpers
    def _st(self):
        """torch's current stream ON THIS ENGINE'S DEVICE (a stream handle of another device is invalid there)."""
        return _stream(self.device)

    def _call(self, fn, args):
        """Library call with this engine's device current: every tbnn_* entry point does cudaSetDevice(handle device),
        so without the guard a handle on device k would leave the thread's CUDA device changed behind torch's back."""
        with torch.cuda.device(self.device):
            _lib.check(fn(*args))

    def tensor(self, a, shape=None):
        t = torch.as_tensor(np.asarray(a), dtype=self.dtype).to(self.dev).contiguous() \
            if not isinstance(a, torch.Tensor) else a.to(self.dev, self.dtype).contiguous()
        return t.reshape(shape) if shape is not None else t

    def _eps(self, eps):
        e = np.ascontiguousarray(np.broadcast_to(np.asarray(eps, dtype=np.float64), (self.chains,)))
        return e, e.ctypes.data_as(C.POINTER(C.c_double))

    @property
    def launches(self):
        return int(self.lib.tbnn_launch_count(self.h))

    def sweep_info(self):
        """dict(kernel=..., ctas_per_chain=..., rows_per_tile=..., smem_bytes=...) of the planned row sweep."""
        k, s, r, b = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        self._call(self.lib.tbnn_sweep_info, (self.h, C.byref(k), C.byref(s), C.byref(r), C.byref(b)))
        return {"kernel": ["k_partial", "k_sweep_wide", "k_sweep_wide2", "k_sweep_umma", "k_train_umma"][k.value], "ctas_per_chain": s.value,
                "rows_per_tile": r.value, "smem_bytes": b.value}

    def predict_kernel(self):
        k = C.c_int()
        self._call(self.lib.tbnn_predict_info, (self.h, C.byref(k)))
        return ["k_predict", "k_predict_umma"][k.value]

    # ------------------------------------------------------------------ data
    def set_data(self, X, Y):
        X = self.tensor(X)
        Y = self.tensor(Y)
        self.N = X.shape[0]
        X = X.reshape(self.N, -1)
        Y = Y.reshape(self.N, -1)
        self._data = (X, Y)
        self._call(self.lib.tbnn_set_data, (self.h, _ptr(X), _ptr(Y), self.N))

    def set_data_host(self, X_host, Y_host):
        """X_host / Y_host: CPU tensors (pinned for an async copy) of the engine dtype."""
        assert X_host.dtype == self.dtype and Y_host.dtype == self.dtype and not X_host.is_cuda
        self.N = X_host.shape[0]
        self._data = (X_host, Y_host)
        self._call(self.lib.tbnn_set_data_host, (self.h, _ptr(X_host), _ptr(Y_host), self.N, self._st()))

    # ------------------------------------------------------------------ targets
    def logp_grad(self, theta, hyper):
        theta = self.tensor(theta, (self.chains, self.P))
        hyper = self.tensor(hyper, (self.chains, self.H))
        logp = torch.empty(self.chains, dtype=self.dtype, device=self.dev)
        grad = torch.empty(self.chains, self.P, dtype=self.dtype, device=self.dev)
        stat = torch.empty(self.chains, dtype=self.dtype, device=self.dev)
        self._call(self.lib.tbnn_logp_grad, (self.h, _ptr(theta), _ptr(hyper), _ptr(logp), _ptr(grad),
                                           _ptr(stat), self._st()))
        return logp, grad, stat

    def hyper_logp_grad(self, theta, hyper, sse=None):
        theta = self.tensor(theta, (self.chains, self.P))
        hyper = self.tensor(hyper, (self.chains, self.H))
        sse_t = self.tensor(sse, (self.chains,)) if sse is not None else None
        logp = torch.empty(self.chains, dtype=self.dtype, device=self.dev)
        grad = torch.empty(self.chains, self.H, dtype=self.dtype, device=self.dev)
        self._call(self.lib.tbnn_hyper_logp_grad, (self.h, _ptr(theta), _ptr(hyper), _ptr(sse_t),
                                                 _ptr(logp), _ptr(grad), self._st()))
        return logp, grad

    # ------------------------------------------------------------------ sampler
    def trajectory(self, theta, hyper, momentum, eps, L):
        theta = self.tensor(theta, (self.chains, self.P))
        hyper = self.tensor(hyper, (self.chains, self.H))
        momentum = self.tensor(momentum, (self.chains, self.P))
        th = torch.empty_like(theta)
        p = torch.empty_like(theta)
        g = torch.empty_like(theta)
        lp = torch.empty(self.chains, dtype=self.dtype, device=self.dev)
        e, ep = self._eps(eps)
        self._call(self.lib.tbnn_trajectory, (self.h, _ptr(theta), _ptr(hyper), _ptr(momentum), ep, int(L),
                                            _ptr(th), _ptr(p), _ptr(lp), _ptr(g), self._st()))
        return th, p, lp, g

    def hmc_step(self, theta, hyper, seed, counter, eps, L, momentum=None, u=None, stats=None):
        """theta [C,P] (device, engine dtype, contiguous) is updated IN PLACE."""
        assert theta.is_cuda and theta.dtype == self.dtype and theta.is_contiguous()
        hyper = self.tensor(hyper, (self.chains, self.H))
        if momentum is not None:
            momentum = self.tensor(momentum, (self.chains, self.P))
        if u is not None:
            u = self.tensor(u, (self.chains,))
        if stats is None:
            stats = torch.empty(self.chains, 4, dtype=self.dtype, device=self.dev)
        e, ep = self._eps(eps)
        self._call(self.lib.tbnn_hmc_step, (self.h, _ptr(theta), _ptr(hyper), int(seed), int(counter), ep,
                                          int(L), _ptr(momentum), _ptr(u), _ptr(stats), self._st()))
        return stats

    def draw_momentum(self, seed, counter):
        p = torch.empty(self.chains, self.P, dtype=self.dtype, device=self.dev)
        ke = torch.empty(self.chains, dtype=self.dtype, device=self.dev)
        self._call(self.lib.tbnn_draw_momentum, (self.h, int(seed), int(counter), _ptr(p), _ptr(ke), self._st()))
        return p, ke

    def time_sweep(self, theta, iters=20):
        """(avg_ms, min_ms) of the row-sweep kernel, CUDA events on the launching stream."""
        theta = self.tensor(theta, (self.chains, self.P))
        a, m = C.c_float(), C.c_float()
        self._call(self.lib.tbnn_time_sweep, (self.h, _ptr(theta), int(iters), C.byref(a), C.byref(m), self._st()))
        return float(a.value), float(m.value)

    def time_allreduce(self, iters=20):
        """(avg_ms, min_ms) of the per-gradient-evaluation exchange of the row-sharded path (collective call)."""
        a, m = C.c_float(), C.c_float()
        self._call(self.lib.tbnn_time_allreduce, (self.h, int(iters), C.byref(a), C.byref(m), self._st()))
        return float(a.value), float(m.value)

    def hyper_step(self, theta, hyper, seed, counter, hyperL, epoch, burnin, hyper_step0, da_state,
                   momentum=None, u=None, stats=None):
        """hyper [C,H] and da_state [C,3] = (h, logEpsilonBar, step) are updated IN PLACE."""
        assert hyper.is_cuda and hyper.dtype == self.dtype and hyper.is_contiguous()
        assert da_state.is_cuda and da_state.dtype == self.dtype and da_state.is_contiguous()
        theta = self.tensor(theta, (self.chains, self.P))
        if momentum is not None:
            momentum = self.tensor(momentum, (self.chains, self.H))
        if u is not None:
            u = self.tensor(u, (self.chains,))
        if stats is None:
            stats = torch.empty(self.chains, 2, dtype=self.dtype, device=self.dev)
        self._call(self.lib.tbnn_hyper_step, (self.h, _ptr(theta), _ptr(hyper), int(seed), int(counter),
                                            int(hyperL), float(epoch), float(burnin), float(hyper_step0),
                                            _ptr(da_state), _ptr(momentum), _ptr(u), _ptr(stats),
                                            self._st()))
        return stats

    # ------------------------------------------------------------------ predictor
    def predict(self, samples, X, want_out=True, want_moments=False):
        samples = self.tensor(samples)
        S = samples.shape[0]
        samples = samples.reshape(S, self.P)
        X = self.tensor(X)
        M = X.shape[0]
        X = X.reshape(M, -1)
        n_out = [l for l in self.arch if l[0] in _lib.DENSE][-1][2]
        out = torch.empty(S, n_out, M, dtype=self.dtype, device=self.dev) if want_out else None
        mom = torch.zeros(3, n_out, M, dtype=self.dtype, device=self.dev) if want_moments else None
        self._call(self.lib.tbnn_predict, (self.h, _ptr(samples), S, _ptr(X), M, _ptr(out), _ptr(mom),
                                         self._st()))
        return out, mom

    # ------------------------------------------------------------------ multi-GPU (row sharding)
    def comm_init(self, unique_id, rank, world):
        buf = (C.c_char * 128).from_buffer_copy(bytes(unique_id))
        self._call(self.lib.tbnn_comm_init, (self.h, C.cast(buf, C.c_void_p), int(rank), int(world)))

    @staticmethod
    def comm_unique_id():
        buf = (C.c_char * 128)()
        _lib.check(_lib.load().tbnn_comm_unique_id(C.cast(buf, C.c_void_p)))
        return bytes(buf)


def adapter_ucb(device, eGrid, lGrid, prev, Kinv, KinvR, s, p, rootbeta, el, eu, Ll, Lu, sigma):
    """paramAdapter.gridSearch on device; returns (e, L, ucb) as python floats."""
    lib = _lib.load()
    f = lambda a: np.ascontiguousarray(np.asarray(a, dtype=np.float32))
    eGrid, lGrid, prev, Kinv, KinvR, sigma = map(f, (eGrid, lGrid, prev, Kinv, KinvR, sigma))
    pf = C.POINTER(C.c_float)
    out = (C.c_float * 2)()
    ucb = C.c_float()
    _lib.check(lib.tbnn_adapter_ucb(int(device), eGrid.ctypes.data_as(pf), eGrid.size,
                                    lGrid.ctypes.data_as(pf), lGrid.size, prev.ctypes.data_as(pf),
                                    prev.reshape(-1, 2).shape[0], Kinv.ctypes.data_as(pf),
                                    KinvR.ctypes.data_as(pf), float(s), float(p), float(rootbeta),
                                    float(el), float(eu), float(Ll), float(Lu), sigma.ctypes.data_as(pf),
                                    out, C.byref(ucb)))
    return float(out[0]), float(out[1]), float(ucb.value)
