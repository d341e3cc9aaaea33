"""The C-ABI library loads on a machine without a GPU, exports every symbol that
include/ocs_b200.h declares, and fails loudly (no fallback) when there is no CUDA device."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "ocs_b200.h")).read()
    return sorted(set(re.findall(r"OCS_B200_API\s+[\w\s\*]+?\b(ocs_b200_\w+)\s*\(", src)))


def test_header_declares_the_expected_surface():
    syms = declared_symbols()
    assert len(syms) >= 34
    for must in ("ocs_b200_create", "ocs_b200_destroy", "ocs_b200_last_error", "ocs_b200_compute_states",
                 "ocs_b200_compute_adjoints", "ocs_b200_set_control_basis", "ocs_b200_set_problem_source",
                 "ocs_b200_fb_sweep", "ocs_b200_fetch_states", "ocs_b200_fetch_adjoints", "ocs_b200_evaluate",
                 "ocs_b200_single_shooting", "ocs_b200_set_fb_costate", "ocs_b200_eval_solution",
                 "ocs_b200_set_control_ufunc", "ocs_b200_compute_equilibrium"):
        assert must in syms


def test_library_exports_every_declared_symbol(ocs):
    lib = ocs._lib.load()
    for name in declared_symbols():
        assert hasattr(lib, name), f"libocs_b200.so does not export {name}"
        assert name in ocs._lib.SIGNATURES, f"{name} has no ctypes signature"
    assert lib.ocs_b200_abi_version() == 1


def test_no_cpu_fallback(ocs):
    """Without a CUDA device create() returns OCS_B200_ERR_CUDA and says why; with one, bad
    arguments are rejected with OCS_B200_ERR_ARG.  Nothing computes on the host."""
    from conftest import has_cuda
    lib = ocs._lib.load()
    h = ctypes.c_void_p()
    ts = np.array([0.0, 1.0])
    if not has_cuda():
        rc = lib.ocs_b200_create(1, ocs._lib.dp(ts), 0, None, 0, 0, ctypes.byref(h))
        assert rc == ocs._lib.ERR_CUDA and not h.value
        assert b"no CUDA device" in lib.ocs_b200_last_error(None)
        with pytest.raises(ocs.OcsError):
            ocs.CudaRK4Integrator(ts)
    rc = lib.ocs_b200_create(0, ocs._lib.dp(ts), 0, None, 0, 0, ctypes.byref(h))
    assert rc == ocs._lib.ERR_ARG and not h.value
    assert lib.ocs_b200_destroy(None) == 0


def test_product_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under the package may import or load it."""
    pkg = os.path.join(ROOT, "optimal-control-solvers_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".c", ".m")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, re.M), f
                assert "liboracle" not in text and "oracle/" not in text.replace("oracle/mlab", ""), f
