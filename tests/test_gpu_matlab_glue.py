"""The MATLAB host side (optimal-control-solvers_b200/matlab/*.m) executed by the MATLAB-subset interpreter against the real
library on the GPU: tests/backprop_test.m of the reference with the integrator swapped for CudaRK4Integrator
(tests/matlab/backprop_test_cuda.m), the reference's own TestOCProblem and PWLinearControl classes untouched.
`ocs_b200_mex` is supplied by tests/mlab_mex_bridge.py (same C-ABI calls as the real gateway)."""
import os

import numpy as np
import pytest

from util import TOL, relerr

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"


def _interp(ocs):
    import mlab_reference_stubs
    from mlab_mex_bridge import make_mex
    from oracle.mlab.interp import Interpreter
    it = Interpreter([os.path.join(ROOT, "optimal-control-solvers_b200", "matlab"), os.path.join(ROOT, "tests", "matlab")])
    it.builtins["ocs_b200_mex"] = make_mex(ocs)
    # a GPU box has no /root/reference (and reference sources are never copied here): the reference classes the
    # script touches are stand-ins -- see tests/mlab_reference_stubs.py
    mlab_reference_stubs.install(it, ocs)
    return it


def test_reference_backprop_test_with_the_drop_in_integrator(ocs, oracle):
    it = _interp(ocs)
    fr = it.run_script("backprop_test_cuda")
    g = np.load(os.path.join(ROOT, "tests", "golden", "backprop_test.npz"))
    w = fr.vars
    assert relerr(w["J"], g["J"]) < TOL
    assert relerr(w["dJdu"], g["dJdu"]) < TOL
    assert relerr(w["dJdv"], g["dJdv"]) < TOL
    # what the reference's script prints, with the CUDA integrator inside
    maxDev, meanDev = float(np.asarray(w["maxDev"]).ravel()[0]), float(np.asarray(w["meanDev"]).ravel()[0])
    assert abs(maxDev - 4.28269e-05) < 5e-8 and abs(meanDev - 3.40849e-05) < 5e-8
    assert w["x"].shape == (2, 501) and w["lam"].shape == (2, 501) and np.all(w["lam"][-1] == 1.0)
    # batched fused evaluation through the .m class
    out = w["out"]
    assert out["J"].shape == (1, 3) and out["dJdv"].shape == (21, 3)
    assert relerr(out["J"][0, 0], g["J"]) < TOL
    # infinite horizon through CudaRK4InfiniteIntegrator.m vs the oracle
    tspan = np.asarray(w["tspan"]).ravel()
    oi = oracle.RK4InfiniteIntegrator(tspan, np.asarray(w["tspanExtra"]).ravel(), 0.723368780097983)
    op = oracle.TestOCProblem(1.5, 3.0, 0.05)
    ox, oJ = oi.compute_states(op, [1.0], np.asarray(w["uInf"]))
    olam, odu = oi.compute_adjoints(op, np.asarray(w["uInf"]))
    assert relerr(w["JInf"], oJ) < TOL and relerr(w["xInf"], ox) < TOL
    assert relerr(w["lamInf"], olam) < TOL and relerr(w["dJduInf"], odu) < TOL
    # device-side single shooting driven from MATLAB code
    ss = w["ss"]
    assert ss["v"].shape == (21, 3) and np.all(ss["converged"] > 0)
    assert np.all(ss["v"] >= 0) and np.all(ss["v"] <= 1) and np.all(ss["J"] < out["J"] + 1e-9)
    # eval_solution through the .m class: soln.u of PWLinearControl is the linear interpolant of v over controlPts
    sol = w["solnSS"]
    assert sol["x"].shape == (1, 41, 3) and sol["lam"].shape == (1, 41, 3) and sol["u"].shape == (1, 41, 3)
    tq = np.linspace(0, 10, 41)
    cp = ocs.matlab_linspace(0, 10, 21)
    for b in range(3):
        assert relerr(sol["u"][0, :, b], np.interp(tq, cp, ss["v"][:, b])) < TOL
    assert np.allclose(sol["x"][0, 0, :], [1, 1.5, 2]) and np.all(np.abs(sol["lam"][0, -1, :]) < 1e-12)
    # compute_equilibrium_cuda.m / CudaRK4Integrator.compute_equilibrium (tests/solve_test_problem.m:21-29)
    sc = lambda z: float(np.asarray(z).ravel()[0])
    assert abs(sc(w["uEq"]) - 0.723368780097983) < 1e-12 and abs(sc(w["xEq"]) - 2.73556918863414) < 1e-12
    assert sc(w["flagEq"]) == 1 and sc(w["resnormEq"]) < 1e-24
    assert np.asarray(w["uEqS"]).shape == (1, 3) and abs(np.asarray(w["uEqS"])[0, 1] - sc(w["uEq"])) < 1e-12
    for k, r in enumerate((.04, .05, .06)):
        oxs, ols, ous = oracle.solve_equilibrium(oracle.TestOCProblem(1.5, 3.0, r), [2.7], [2.2], [.7], r)
        assert abs(np.asarray(w["xEqS"])[0, k] - oxs[0]) < 1e-11 and abs(np.asarray(w["lamEqS"])[0, k] - ols[0]) < 1e-11
    # fb_sweep through the .m class, both costates
    fbD, fbC = w["fbD"], w["fbC"]
    assert np.all(np.asarray(fbD["converged"]) > 0) and np.all(np.asarray(fbC["converged"]) > 0)
    assert np.asarray(fbC["u"]).shape == (1, 1001, 3)
    assert np.max(np.abs(np.asarray(fbD["J"]) - np.asarray(fbC["J"])) / np.abs(np.asarray(fbD["J"]))) < 1e-4
