"""CUDA sweeps (through the C ABI) against the committed golden vectors, i.e. against the
reference's own .m files executed by oracle/mlab (tests/golden/generate_golden.py).
Tolerance 1e-10 normwise (north star); grid, basis and control-grid indices bit-exact."""
import os

import numpy as np
import pytest

from util import TEST_PARAMS, TOL, relerr

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return np.load(os.path.join(GOLD, name + ".npz"))


def _run(ocs, g, ctl_cls, K, x0, math):
    tspan = g["tspan"].reshape(-1)
    integ = ocs.CudaRK4Integrator(tspan, math=math)
    assert np.array_equal(integ.t, g["integrator__t"].reshape(-1))      # RK4Integrator.m:17-24, bit-exact
    assert np.array_equal(integ.h, g["integrator__h"].reshape(-1))
    ctl = getattr(ocs, ctl_cls)(integ.t, K, 1)
    assert np.array_equal(ctl.B, g["control__B"])                        # host basis == MATLAB-source basis
    prob = ocs.TestOCProblem(TEST_PARAMS, [[0, 1]])
    v = g["v"].reshape(-1)
    u = ctl.compute_u(v)
    x, J = integ.compute_states(prob, x0, u)
    lam, dJdu = integ.compute_adjoints(prob, u)
    integ.attach_control(ctl)
    fused = integ.evaluate(prob, x0, v)
    return integ, ctl, prob, v, u, x, J, lam, dJdu, fused


@pytest.mark.parametrize("math", ["fast", "strict"])
def test_backprop_test_script(ocs, math):
    g = load("backprop_test")
    integ, ctl, prob, v, u, x, J, lam, dJdu, fused = _run(ocs, g, "PWLinearControl", 21, 1.0, math)
    assert relerr(J, g["J"]) < TOL
    assert relerr(dJdu, g["dJdu"]) < TOL
    assert relerr(fused["J"], g["J"]) < TOL
    assert relerr(fused["dJdv"][:, 0], g["dJdv"].reshape(-1)) < TOL
    # the property the reference's script prints: adjoint gradient == finite differences
    eps = 1e-4
    fd = np.empty(21)
    for i in range(21):
        vp = v.copy()
        vp[i] += eps
        fd[i] = (integ.evaluate(prob, 1.0, vp, gradient=False)["J"][0] - fused["J"][0]) / eps
    dev = fd - fused["dJdv"][:, 0]
    assert abs(dev.max() - 4.28269e-05) < 5e-8 and abs(np.abs(dev).mean() - 3.40849e-05) < 5e-8


@pytest.mark.parametrize("math", ["fast", "strict"])
def test_pwconstant(ocs, math):
    g = load("golden_pwconstant")
    integ, ctl, prob, v, u, x, J, lam, dJdu, fused = _run(ocs, g, "PWConstantControl", 100, 1.0, math)
    k, nnz = integ.control_index()
    assert np.array_equal(k, np.argmax(g["B"], axis=0)) and np.all(nnz == 1)     # control-grid indices
    assert np.array_equal(integ.get_control_u()[0, :, 0], g["u"].reshape(-1))    # bit-exact u
    assert relerr(J, g["J"]) < TOL and relerr(x, g["x"]) < TOL
    assert relerr(lam, g["lam"]) < TOL and relerr(dJdu, g["dJdu"]) < TOL
    assert relerr(fused["dJdv"][:, 0], g["dJdv"].reshape(-1)) < TOL
    assert relerr(fused["lam0"][:, 0], g["lam"][:, 0]) < TOL


def test_pwlinear101_and_lamT(ocs):
    g = load("golden_pwlinear101")
    integ, ctl, prob, v, u, x, J, lam, dJdu, fused = _run(ocs, g, "PWLinearControl", 101, 1.7, "fast")
    assert relerr(J, g["J"]) < TOL and relerr(x, g["x"]) < TOL
    assert relerr(lam, g["lam"]) < TOL and relerr(dJdu, g["dJdu"]) < TOL
    assert relerr(fused["dJdv"][:, 0], g["dJdv"].reshape(-1)) < TOL
    integ.attach_control(None)
    integ.compute_states(prob, 1.7, u)
    lam_T, dJdu_T = integ.compute_adjoints(prob, u, g["lamT"].reshape(-1))
    assert relerr(lam_T, g["lam_T"]) < TOL and relerr(dJdu_T, g["dJdu_T"]) < TOL


def test_chebyshev(ocs):
    g = load("golden_chebyshev")
    integ, ctl, prob, v, u, x, J, lam, dJdu, fused = _run(ocs, g, "ChebyshevControl", 17, 0.8, "fast")
    assert relerr(integ.get_control_u()[0, :, 0], g["u"].reshape(-1)) < TOL
    assert relerr(J, g["J"]) < TOL and relerr(x, g["x"]) < TOL
    assert relerr(lam, g["lam"]) < TOL and relerr(dJdu, g["dJdu"]) < TOL
    assert relerr(fused["dJdv"][:, 0], g["dJdv"].reshape(-1)) < TOL
    # per-trajectory coefficients take the dense-basis kernels
    vb = np.repeat(v[:, None], 3, axis=1)
    out = integ.evaluate(prob, np.full((1, 3), 0.8), vb)
    assert relerr(out["J"], np.repeat(g["J"], 3)) < TOL
    assert relerr(out["dJdv"][:, 1], g["dJdv"].reshape(-1)) < TOL


def test_infinite_horizon(ocs):
    g = load("golden_infinite")
    uStar = float(g["uStar"].reshape(-1)[0])
    integ = ocs.CudaRK4InfiniteIntegrator(g["tspan"].reshape(-1), g["tspanExtra"].reshape(-1), uStar)
    ctl = ocs.PWLinearControl(integ.t, 101, 1)
    prob = ocs.TestOCProblem(TEST_PARAMS, [[0, 1]])
    v = g["v"].reshape(-1)
    u = ctl.compute_u(v)
    x, J = integ.compute_states(prob, 1.0, u)
    lam, dJdu = integ.compute_adjoints(prob, u)
    assert relerr(J, g["J"]) < TOL and relerr(x, g["x"]) < TOL
    assert relerr(lam, g["lam"]) < TOL and relerr(dJdu, g["dJdu"]) < TOL
    lam2, _ = integ.fetch_adjoints(segment=1, want_dJdu=False)
    assert relerr(lam2[:, :, 0], g["lam2"]) < TOL
    assert relerr(integ.fetch_states(segment=1)[:, :, 0], g["x2"]) < TOL
    integ.attach_control(ctl)
    out = integ.evaluate(prob, 1.0, v)
    assert relerr(out["J"], g["J"]) < TOL and relerr(out["dJdv"][:, 0], g["dJdv"].reshape(-1)) < TOL
