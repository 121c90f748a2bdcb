"""ctypes binding of libtbnn.so (include/tbnn.h).  There is no CPU fallback: if the
library is missing or a call fails, a RuntimeError is raised."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libtbnn.so")

# enums of include/tbnn.h
DENSE_CAUCHY, DENSE_GAUSSIAN = 0, 1
ACT = {"relu": 10, "tanh": 11, "sigmoid": 12, "Exp": 13, "elu": 14, "leakyrelu": 15, "prelu": 16,
       "squareprelu": 17}
DENSE = {"dense": DENSE_CAUCHY, "denseGaussian": DENSE_GAUSSIAN}
LIK = {"gaussian": 0, "fixed": 1, "bernoulli": 2}
F32, F64 = 0, 1
FLAG_NO_WIDE = 1   # tbnn_desc.flags: never use the wide-first-layer row sweep
FLAG_NO_UMMA = 2   # tbnn_desc.flags: never use the tcgen05 (tensor-core) kernels
FLAG_NO_WIDE2 = 4  # tbnn_desc.flags: phase-serial wide sweep instead of the warp-specialised one
FLAG_NO_PERSISTENT = 16  # tbnn_desc.flags: never run a trajectory as one persistent launch (k_traj_small)
FLAG_NO_NARROW = 32  # tbnn_desc.flags: persistent trajectories on the tile engine only (no k_traj_narrow)
FLAG_NO_UMMA_TRAIN = 64  # tbnn_desc.flags: hidden-layer GEMMs of the training sweep on FP32 FFMA (k_partial), not tcgen05
FLAG_UMMA_SWEEP = 8  # tbnn_desc.flags: wide-first-layer sweep on tcgen05 (k_sweep_umma) instead of FP32 FFMA2 (opt-in)

EXPORTS = ["tbnn_last_error", "tbnn_version", "tbnn_create", "tbnn_destroy", "tbnn_num_params",
           "tbnn_num_hypers", "tbnn_launch_count", "tbnn_sweep_info", "tbnn_wide_profile", "tbnn_predict_info", "tbnn_set_data", "tbnn_set_data_host",
           "tbnn_logp_grad", "tbnn_hyper_logp_grad", "tbnn_trajectory", "tbnn_hmc_step",
           "tbnn_draw_momentum", "tbnn_time_sweep", "tbnn_time_allreduce", "tbnn_hyper_step", "tbnn_adapter_ucb", "tbnn_predict", "tbnn_comm_unique_id",
           "tbnn_comm_init"]


class LayerDesc(C.Structure):
    _fields_ = [("kind", C.c_int32), ("in_dim", C.c_int32), ("out_dim", C.c_int32), ("alpha", C.c_double)]


class Desc(C.Structure):
    _fields_ = [("n_layers", C.c_int32), ("layers", C.POINTER(LayerDesc)), ("likelihood", C.c_int32),
                ("fixed_sd", C.c_double), ("dtype", C.c_int32), ("chains", C.c_int32),
                ("device", C.c_int32), ("flags", C.c_int32)]


_lib = None


def load():
    """Returns the loaded library; raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError("tensorbnn_b200: %s is missing -- build it with "
                           "`python -m tensorbnn_b200.build` (needs nvcc); there is no CPU fallback"
                           % LIB_PATH)
    lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    vp, i32, i64, u64, dbl, flt = C.c_void_p, C.c_int, C.c_int64, C.c_uint64, C.c_double, C.c_float
    pd, pf = C.POINTER(C.c_double), C.POINTER(C.c_float)
    lib.tbnn_last_error.restype = C.c_char_p
    lib.tbnn_last_error.argtypes = []
    lib.tbnn_version.restype = i32
    lib.tbnn_create.argtypes = [C.POINTER(Desc), C.POINTER(vp)]
    lib.tbnn_destroy.argtypes = [vp]
    lib.tbnn_num_params.argtypes = [vp]
    lib.tbnn_num_hypers.argtypes = [vp]
    lib.tbnn_launch_count.argtypes = [vp]
    lib.tbnn_launch_count.restype = i64
    pi = C.POINTER(C.c_int)
    lib.tbnn_sweep_info.argtypes = [vp, pi, pi, pi, pi]
    lib.tbnn_predict_info.argtypes = [vp, pi]
    lib.tbnn_wide_profile.argtypes = [vp, vp, C.POINTER(C.c_longlong), vp]
    lib.tbnn_set_data.argtypes = [vp, vp, vp, i64]
    lib.tbnn_set_data_host.argtypes = [vp, vp, vp, i64, vp]
    lib.tbnn_logp_grad.argtypes = [vp, vp, vp, vp, vp, vp, vp]
    lib.tbnn_hyper_logp_grad.argtypes = [vp, vp, vp, vp, vp, vp, vp]
    lib.tbnn_trajectory.argtypes = [vp, vp, vp, vp, pd, i32, vp, vp, vp, vp, vp]
    lib.tbnn_hmc_step.argtypes = [vp, vp, vp, u64, u64, pd, i32, vp, vp, vp, vp]
    lib.tbnn_draw_momentum.argtypes = [vp, u64, u64, vp, vp, vp]
    lib.tbnn_time_sweep.argtypes = [vp, vp, i32, pf, pf, vp]
    lib.tbnn_time_allreduce.argtypes = [vp, i32, pf, pf, vp]
    lib.tbnn_hyper_step.argtypes = [vp, vp, vp, u64, u64, i32, dbl, dbl, dbl, vp, vp, vp, vp, vp]
    lib.tbnn_adapter_ucb.argtypes = [i32, pf, i32, pf, i32, pf, i32, pf, pf, flt, flt, flt, flt, flt, flt,
                                     flt, pf, pf, pf]
    lib.tbnn_predict.argtypes = [vp, vp, i64, vp, i64, vp, vp, vp]
    lib.tbnn_comm_unique_id.argtypes = [vp]
    lib.tbnn_comm_init.argtypes = [vp, vp, i32, i32]
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        raise RuntimeError("libtbnn: " + load().tbnn_last_error().decode("utf-8", "replace"))


def make_desc(arch, lik, dtype_code, chains, device, flags=0):
    """arch / lik in the vocabulary of tensorbnn_b200.workloads."""
    n = len(arch)
    layers = (LayerDesc * n)()
    for i, layer in enumerate(arch):
        k = layer[0]
        if k in DENSE:
            layers[i] = LayerDesc(DENSE[k], int(layer[1]), int(layer[2]), 0.0)
        elif k in ("prelu", "squareprelu"):
            layers[i] = LayerDesc(ACT[k], int(layer[1]), 0, 0.0)
        elif k == "leakyrelu":
            layers[i] = LayerDesc(ACT[k], 0, 0, float(layer[1]))
        elif k in ACT:
            layers[i] = LayerDesc(ACT[k], 0, 0, 0.0)
        else:
            raise ValueError("layer kind %r is not supported by the CUDA path" % (k,))
    d = Desc(n, layers, LIK[lik[0]], float(lik[1]) if lik[0] == "fixed" else 0.0, dtype_code,
             int(chains), int(device), int(flags))
    return d, layers
