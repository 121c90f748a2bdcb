"""CPU-side checks of the drop-in boundary: libtbnn.so builds for sm_100a, loads, and exports
every symbol include/tbnn.h declares; no compute call is made without a GPU."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="session")
def lib_path():
    from tensorbnn_b200 import build
    path, _ = build.build_library()
    return path


def declared_functions():
    src = open(os.path.join(ROOT, "include", "tbnn.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(tbnn_[a-z_0-9]+)\s*\(", src)))


def test_header_declares_expected_surface():
    names = declared_functions()
    for must in ("tbnn_create", "tbnn_set_data", "tbnn_logp_grad", "tbnn_hyper_logp_grad", "tbnn_trajectory",
                 "tbnn_hmc_step", "tbnn_hyper_step", "tbnn_adapter_ucb", "tbnn_predict", "tbnn_comm_init"):
        assert must in names


def test_library_exports_every_declared_symbol(lib_path):
    lib = ctypes.CDLL(lib_path)
    for name in declared_functions():
        assert hasattr(lib, name), name


def test_binding_covers_header(lib_path):
    from tensorbnn_b200 import _lib
    assert sorted(_lib.EXPORTS) == declared_functions()
    lib = _lib.load()
    assert lib.tbnn_version() >= 100


def test_sass_is_sm100a(lib_path):
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", lib_path], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_no_gpu_fails_loudly(lib_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from tensorbnn_b200.engine import Engine
    with pytest.raises(RuntimeError):
        Engine([("dense", 1, 1)], ("fixed", 0.1))
    # and the C ABI itself refuses without a device
    from tensorbnn_b200 import _lib
    lib = _lib.load()
    desc, keep = _lib.make_desc([("dense", 1, 1)], ("fixed", 0.1), _lib.F32, 1, 0)
    h = ctypes.c_void_p()
    assert lib.tbnn_create(ctypes.byref(desc), ctypes.byref(h)) != 0
    assert len(lib.tbnn_last_error()) > 0


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "tensorbnn_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, f
