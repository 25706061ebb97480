"""CPU: the C-ABI library builds, loads and exports every symbol include/optmc.h declares; the ctypes
binding lists exactly those symbols; without a GPU the product fails loudly (no CPU fallback)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g

    g.build()
    from options_model_b200 import _lib

    return _lib


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "optmc.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(optmc_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported_and_bound(lib):
    names = _declared_symbols()
    assert len(names) >= 20
    so = ctypes.CDLL(lib.library_path())
    for n in names:
        assert hasattr(so, n), f"{n} declared in optmc.h but not exported by liboptmc.so"
    assert sorted(lib.PROTOTYPES) == names  # the binding covers the header, nothing more, nothing less
    assert lib.load_library().optmc_abi_version() == 1


def test_struct_layouts_match_header(lib):
    # sizes implied by the C declarations (x86-64 SysV): catches a drifting ctypes mirror
    assert ctypes.sizeof(lib.ModelParams) == 8 + 9 * 8
    assert ctypes.sizeof(lib.RngParams) == 16 + 16 + 8 + 8
    assert ctypes.sizeof(lib.LsmParams) == 24 + 16
    assert ctypes.sizeof(lib.LsmResult) == 16 + 8 + 8 + 32
    assert ctypes.sizeof(lib.EuropeanResult) == 24
    assert lib.load_library().optmc_lsm_gram_len(lib.BASIS_POLY2) == 8
    assert lib.load_library().optmc_lsm_gram_len(lib.BASIS_POLY3) == 11


def test_no_cpu_fallback(lib):
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    h = ctypes.c_void_p()
    rc = lib.load_library().optmc_ctx_create(0, ctypes.byref(h))
    assert rc == lib.ECUDA and not h.value
    assert b"no CPU fallback" in lib.load_library().optmc_last_error()
    from options_model_b200 import engine

    with pytest.raises(lib.OptmcError):
        engine.Engine(0)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "options-model_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, f


def test_workspace_estimate_needs_no_gpu(lib):
    import ctypes as C

    from options_model_b200 import engine as E

    b1 = E.workspace_estimate(1_000_000, 252, "f32", 1)
    assert 1_000_000 * 253 * 4 <= b1 <= 1.02 * 1_000_000 * 253 * 4 + (8 << 20)
    assert E.workspace_estimate(1_000_000, 252, "f64", 4) >= 4 * 1_000_000 * 253 * 8
    assert E.workspace_estimate(10_000, 50, "f32", 5000) == E.workspace_estimate(10_000, 50, "f32", 160)  # waves
    out = C.c_int64()
    assert lib.load_library().optmc_workspace_bytes(0, 10, 0, 1, C.byref(out)) == -1
