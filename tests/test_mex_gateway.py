"""The MEX gateway (optimal-control-solvers_b200/matlab/ocs_b200_mex.c) cannot be built with MATLAB here (no mex.h),
so it is compiled against tests/stubs/mex.h and driven by tests/stubs/mex_driver.c exactly the
way CudaRK4Integrator.m calls it.  CPU: it compiles, links against libocs_b200.so and turns the
library's "no CUDA device" status into a MATLAB error.  GPU: the tests/backprop_test.m scenario
through the gateway matches the golden vectors."""
import os
import struct
import subprocess

import numpy as np
import pytest

from conftest import has_cuda
from util import TOL, relerr

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "optimal-control-solvers_b200")


@pytest.fixture(scope="module")
def driver(ocs, tmp_path_factory):
    d = tmp_path_factory.mktemp("mex")
    exe = str(d / "mex_driver")
    cmd = ["gcc", "-O1", "-Wall", "-Wextra", "-I", os.path.join(ROOT, "tests", "stubs"), "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tests", "stubs", "mex_driver.c"), os.path.join(PKG, "matlab", "ocs_b200_mex.c"),
           "-L", PKG, "-locs_b200", "-Wl,-rpath," + PKG, "-o", exe]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "warning" not in r.stderr, r.stderr
    return exe, d


def _write_input(path):
    g = np.load(os.path.join(ROOT, "tests", "golden", "golden_pwlinear101.npz"))
    tspan, u = g["tspan"].reshape(-1), g["u"].reshape(-1)
    N = tspan.size - 1
    with open(path, "wb") as f:
        f.write(struct.pack("ii", N, 1))
        f.write(tspan.tobytes())
        f.write(np.array([1.5, 3.0, 0.05]).tobytes())
        f.write(np.array([1.7]).tobytes())
        f.write(u.tobytes())
    return g, N


def test_gateway_compiles_and_reports_errors_like_matlab(driver):
    exe, d = driver
    g, N = _write_input(str(d / "in.bin"))
    r = subprocess.run([exe, str(d / "in.bin"), str(d / "out.bin")], capture_output=True, text=True)
    if has_cuda():
        assert r.returncode == 0, r.stderr
    else:
        assert r.returncode == 3
        assert "ocs_b200:create" in r.stderr and "no CUDA device" in r.stderr


@pytest.mark.gpu
def test_gateway_matches_golden(driver):
    exe, d = driver
    g, N = _write_input(str(d / "in.bin"))
    r = subprocess.run([exe, str(d / "in.bin"), str(d / "out.bin")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    out = np.fromfile(str(d / "out.bin"))
    J, x, lam, dJdu = out[0], out[1:1 + 2 * (N + 1)], out[1 + 2 * (N + 1):1 + 4 * (N + 1)], out[1 + 4 * (N + 1):]
    assert relerr(J, g["J"]) < TOL
    assert relerr(x.reshape(2, N + 1, order="F"), g["x"]) < TOL
    assert relerr(lam.reshape(2, N + 1, order="F"), g["lam"]) < TOL
    assert relerr(dJdu, g["dJdu"].reshape(-1)) < TOL


@pytest.mark.gpu
def test_gateway_lifecycle_and_shape_checks(driver):
    """mexLock while an integrator lives, the mexAtExit hook frees live handles (`clear mex` with a live object),
    stale / bogus handle values are refused, and wrong input shapes never reach the C ABI (sizes come from
    ocs_b200_get_dims, not from caller scalars)."""
    exe, d = driver
    _write_input(str(d / "in.bin"))
    r = subprocess.run([exe, "--lifecycle", str(d / "in.bin")], capture_output=True, text=True)
    assert r.returncode == 0 and "lifecycle ok" in r.stdout, (r.returncode, r.stdout, r.stderr)


def test_matlab_classes_run_under_the_interpreter_up_to_the_device(ocs):
    """CPU: the MATLAB host classes parse and execute under oracle/mlab (constructor, inputParser, property
    access) until the first C-ABI call, which reports the missing device exactly like the MEX gateway would."""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import mlab_reference_stubs
    from mlab_mex_bridge import make_mex
    from oracle.mlab.interp import Interpreter
    it = Interpreter([os.path.join(PKG, "matlab"), os.path.join(ROOT, "tests", "matlab")])
    it.builtins["ocs_b200_mex"] = make_mex(ocs)
    mlab_reference_stubs.install(it, ocs)
    if has_cuda():
        fr = it.run_script("backprop_test_cuda")
        assert "J" in fr.vars
    else:
        with pytest.raises(RuntimeError, match="no CUDA device"):
            it.run_script("backprop_test_cuda")
