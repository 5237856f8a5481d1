"""The C-ABI library loads without a GPU and exports every symbol include/lbm.h declares; calls that
need a device fail loudly (no CPU fallback anywhere on the product path)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "lbm.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(lbm_[a-z_0-9]+)\s*\(", text)))


def test_header_and_binding_agree(lbm):
    assert declared_symbols() == sorted(lbm.cabi.EXPORTS)


def test_library_exports_every_declared_symbol(lbm):
    lib = lbm.cabi.load_library()
    for name in declared_symbols():
        assert hasattr(lib, name), name
    assert lib.lbm_abi_version() == 2
    assert lib.lbm_export_size() > 64


def test_params_struct_matches_t_param(lbm):
    """t_param (d2q9-bgk.c:81-92): four floats then four ints, 32 bytes."""
    assert C.sizeof(lbm.cabi.LbmParams) == 32
    names = [n for n, _ in lbm.cabi.LbmParams._fields_]
    assert names == ["density", "accel", "omega", "free_cells_inv", "nx", "ny", "maxIters", "reynolds_dim"]


def test_product_fails_loudly_without_gpu(lbm):
    lib = lbm.cabi.load_library()
    if lib.lbm_device_count() > 0:
        pytest.skip("a GPU is visible")
    p = lbm.decks.Params(nx=8, ny=8, maxIters=1, reynolds_dim=10, density=0.1, accel=0.005, omega=1.85,
                         free_cells_inv=1.0 / 64)
    with pytest.raises(lbm.cabi.LbmError):
        lbm.cabi.Simulation(p)
    assert lib.lbm_last_error().decode() != ""


def test_no_product_file_touches_the_oracle():
    """oracle/ is test infrastructure: nothing under the package, include/ or the Makefile's product
    rules may include or link it."""
    pkg = os.path.join(ROOT, "opencl-lattice-boltzmann_b200")
    for base, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".c", ".h", ".cu", ".cuh")):
                text = open(os.path.join(base, f), errors="ignore").read()
                assert not re.search(r'#include\s*[<"][^>"]*oracle', text), f
                assert not re.search(r"import\s+oracle_lib|from\s+oracle_lib|liboracle", text), f
