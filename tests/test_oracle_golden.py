"""Pins the CPU oracle (NumPy and C restatements) against golden vectors produced by running
the reference's UNMODIFIED .m files under oracle/mlab (tests/golden/generate_golden.py).
Structural quantities (grid, step sizes, basis matrices, control-grid indices) must be
bit-identical; sweeps agree to 1e-13 normwise (the only freedom is the summation order of
v*B / dJdu*B' products, which MATLAB leaves to BLAS)."""
import os

import numpy as np
import pytest

from util import relerr

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOL_ORACLE = 1e-13
P = (1.5, 3.0, 0.05)


def load(name):
    return np.load(os.path.join(GOLD, name + ".npz"))


def _common(oracle, g, ctl_cls, K, x0, lamT=None, check_u=True):
    tspan = np.asarray(g["tspan"]).reshape(-1)
    integ = oracle.RK4Integrator(tspan)
    assert np.array_equal(integ.t, g["integrator__t"].reshape(-1))
    assert np.array_equal(integ.h, g["integrator__h"].reshape(-1))
    ctl = getattr(oracle, ctl_cls)(integ.t, K, 1)
    assert np.array_equal(ctl.B, g["control__B"])            # bit-exact basis, hence bit-exact indices
    prob = oracle.TestOCProblem(*P)
    v = g["v"].reshape(-1)
    u = ctl.compute_u(v)
    if check_u:
        assert relerr(u, g["u"]) < TOL_ORACLE
    x, J = integ.compute_states(prob, [x0], u)
    lam, dJdu = integ.compute_adjoints(prob, u, lamT)
    return integ, ctl, prob, v, u, x, J, lam, dJdu


def test_reference_backprop_test_script(oracle):
    """tests/backprop_test.m run as is: the restatement reproduces its workspace."""
    g = load("backprop_test")
    # the script's workspace keeps the LAST finite-difference u (v(21) perturbed) and a v that went through
    # (v + eps) - eps, so u is not compared and v is the original one only to within an ulp
    integ, ctl, prob, v, u, x, J, lam, dJdu = _common(oracle, g, "PWLinearControl", 21, 1.0, check_u=False)
    assert relerr(J, g["J"]) < TOL_ORACLE
    # integrator.xK holds the stages of the last FD pass; its first component differs by the perturbation
    assert relerr(dJdu, g["dJdu"]) < TOL_ORACLE
    assert relerr(ctl.compute_dJdv(dJdu), g["dJdv"].reshape(-1)) < TOL_ORACLE
    # what the script prints: adjoint gradient vs forward differences (epsilon = 1e-4)
    dev = g["dJdv_FD"].reshape(-1) - g["dJdv"].reshape(-1)
    assert abs(dev.max() - 4.28269e-05) < 1e-9 and abs(np.abs(dev).mean() - 3.40849e-05) < 1e-9
    assert "Max Norm: 4.28269e-05" in str(g["stdout"])
    # C restatement
    out = oracle.C.evaluate(P, g["tspan"].reshape(-1), ctl.B, [[1.0]], v, want_grids=True)
    assert relerr(out["J"], g["J"]) < TOL_ORACLE
    assert relerr(out["dJdu"][:, :, 0], g["dJdu"]) < TOL_ORACLE
    assert relerr(out["dJdv"][:, 0], g["dJdv"].reshape(-1)) < TOL_ORACLE


def test_pwconstant_indices_and_sweeps(oracle):
    g = load("golden_pwconstant")
    integ, ctl, prob, v, u, x, J, lam, dJdu = _common(oracle, g, "PWConstantControl", 100, 1.0)
    assert np.array_equal(ctl.intervalStarts, g["intervalStarts"].reshape(-1))
    k = ctl.interval_index()
    assert np.array_equal(k, np.argmax(g["B"], axis=0))
    assert np.array_equal(u[0], v[k]) and np.array_equal(g["u"].reshape(-1), v[k])   # exact copies of v entries
    assert relerr(J, g["J"]) < TOL_ORACLE and relerr(x, g["x"]) < TOL_ORACLE
    assert relerr(lam, g["lam"]) < TOL_ORACLE and relerr(dJdu, g["dJdu"]) < TOL_ORACLE
    assert relerr(ctl.compute_dJdv(dJdu), g["dJdv"].reshape(-1)) < TOL_ORACLE
    Lb, Ub = ctl.compute_nlp_bounds([[0, 1]])
    assert np.array_equal(Lb, g["Lb"].reshape(-1)) and np.array_equal(Ub, g["Ub"].reshape(-1))
    assert np.array_equal(ctl.compute_initial_v(.25), g["v0"].reshape(-1))


def test_pwlinear101_and_lamT(oracle):
    g = load("golden_pwlinear101")
    integ, ctl, prob, v, u, x, J, lam, dJdu = _common(oracle, g, "PWLinearControl", 101, 1.7)
    assert np.array_equal(ctl.controlPts, g["controlPts"].reshape(-1))
    assert relerr(J, g["J"]) < TOL_ORACLE and relerr(x, g["x"]) < TOL_ORACLE
    assert relerr(lam, g["lam"]) < TOL_ORACLE and relerr(dJdu, g["dJdu"]) < TOL_ORACLE
    lam_T, dJdu_T = integ.compute_adjoints(prob, u, g["lamT"].reshape(-1))
    assert relerr(lam_T, g["lam_T"]) < TOL_ORACLE and relerr(dJdu_T, g["dJdu_T"]) < TOL_ORACLE
    assert np.array_equal(ctl.compute_initial_v(.25), g["v0"].reshape(-1))


def test_chebyshev(oracle):
    g = load("golden_chebyshev")
    integ, ctl, prob, v, u, x, J, lam, dJdu = _common(oracle, g, "ChebyshevControl", 17, 0.8)
    assert relerr(J, g["J"]) < TOL_ORACLE and relerr(x, g["x"]) < TOL_ORACLE
    assert relerr(lam, g["lam"]) < TOL_ORACLE and relerr(dJdu, g["dJdu"]) < TOL_ORACLE
    assert relerr(ctl.compute_dJdv(dJdu), g["dJdv"].reshape(-1)) < TOL_ORACLE
    assert np.array_equal(ctl.compute_initial_v(.25), g["v0"].reshape(-1))


def test_infinite_horizon(oracle):
    g = load("golden_infinite")
    tspan, tspanE = g["tspan"].reshape(-1), g["tspanExtra"].reshape(-1)
    uStar = float(g["uStar"].reshape(-1)[0])
    integ = oracle.RK4InfiniteIntegrator(tspan, tspanE, uStar)
    assert np.array_equal(integ.uStar, g["integrator__uStar"])
    ctl = oracle.PWLinearControl(integ.t, 101, 1)
    prob = oracle.TestOCProblem(*P)
    v = g["v"].reshape(-1)
    u = ctl.compute_u(v)
    x, J = integ.compute_states(prob, [1.0], u)
    lam, dJdu = integ.compute_adjoints(prob, u)
    assert relerr(J, g["J"]) < TOL_ORACLE and relerr(x, g["x"]) < TOL_ORACLE
    assert relerr(lam, g["lam"]) < TOL_ORACLE and relerr(dJdu, g["dJdu"]) < TOL_ORACLE
    assert relerr(integ.lam2, g["lam2"]) < TOL_ORACLE
    out = oracle.C.evaluate(P, tspan, ctl.B, [[1.0]], v, tspan_extra=tspanE, uStar=uStar, want_grids=True)
    assert relerr(out["J"], g["J"]) < TOL_ORACLE
    assert relerr(out["lam"][:, :, 0], g["lam"]) < TOL_ORACLE
    assert relerr(out["dJdv"][:, 0], g["dJdv"].reshape(-1)) < TOL_ORACLE
