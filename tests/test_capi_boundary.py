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


def test_ctypes_mirrors_match_the_compiled_header(lib, tmp_path):
    """Compile include/optmc.h with gcc and compare sizeof / offsetof of every public struct with the ctypes mirror."""
    import subprocess

    pairs = {"optmc_model_params": lib.ModelParams, "optmc_rng_params": lib.RngParams, "optmc_lsm_params": lib.LsmParams,
             "optmc_lsm_result": lib.LsmResult, "optmc_european_result": lib.EuropeanResult,
             "optmc_global_result": lib.GlobalResult, "optmc_mlp_params": lib.MlpParams, "optmc_gnet_params": lib.GnetParams,
             "optmc_gnet_result": lib.GnetResult, "optmc_american_option": lib.AmericanOption,
             "optmc_price_result": lib.PriceResult, "optmc_ivnet": lib.IvNet, "optmc_batch_extras": lib.BatchExtras}
    src = open(os.path.join(ROOT, "include", "optmc.h")).read()
    declared = set(re.findall(r"typedef struct (optmc_[a-z_]+) \{", src))
    assert declared == set(pairs), declared ^ set(pairs)
    lines = ["#include <stdio.h>", "#include <stddef.h>", '#include "optmc.h"', "int main(void) {"]
    for cname, ct in pairs.items():
        lines.append(f'  printf("{cname} %zu", sizeof({cname}));')
        for fname, _ in ct._fields_:
            lines.append(f'  printf(" %zu", offsetof({cname}, {fname}));')
        lines.append('  printf("\\n");')
    lines += ["  return 0;", "}"]
    c = tmp_path / "probe.c"
    c.write_text("\n".join(lines))
    exe = tmp_path / "probe"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(c), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.strip().splitlines()
    for line in out:
        name, size, *offs = line.split()
        ct = pairs[name]
        assert ctypes.sizeof(ct) == int(size), name
        assert [getattr(ct, f).offset for f, _ in ct._fields_] == [int(o) for o in offs], name


def _build_c_example(tmp_path):
    import subprocess

    exe = tmp_path / "price_american"
    libdir = os.path.join(ROOT, "options-model_b200")
    subprocess.run(["gcc", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "price_american.c"),
                    "-o", str(exe), "-L", libdir, "-loptmc", f"-Wl,-rpath,{libdir}"], check=True)
    return exe


def test_c_client_compiles_links_and_fails_loudly_without_gpu(lib, tmp_path):
    """The boundary is a plain C ABI: a C program (no Python, no torch) compiles against include/optmc.h and links
    liboptmc.so.  Without a CUDA device it must refuse to run (exit code 2), never fall back to a CPU path."""
    import subprocess

    import torch

    exe = _build_c_example(tmp_path)
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu-marked run")
    p = subprocess.run([str(exe)], capture_output=True, text=True)
    assert p.returncode == 2 and "no CPU fallback" in p.stderr


@pytest.mark.gpu
def test_c_client_prices_configs_1_and_2(lib, tmp_path):
    import subprocess

    exe = _build_c_example(tmp_path)
    p = subprocess.run([str(exe)], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stdout + p.stderr
    out = dict(line.split(" ", 1) for line in p.stdout.strip().splitlines())
    c1 = float(out["config1"].split()[1])
    c2 = float(out["config2"].split()[1])
    assert 6.3 < c1 < 6.8 and 5.6 < c2 < 6.0  # SURVEY 8(c) pins: 6.54 (GBM, reference semantics), 5.83 (Heston)
    assert "num_simulations and num_time_steps must be positive integers." in out["error"]
