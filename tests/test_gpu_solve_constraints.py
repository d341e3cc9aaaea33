"""ControlBounds on the grid as linear constraints of the batched single_shooting solve (ocs_b200_set_solve_control_bounds):
what a finished ChebyshevControl.compute_lincon hands to fmincon (functions/single_shooting.m:106-111, A v <= b), solved on
the device by an augmented-Lagrangian iteration around the projected-gradient method.  The minimisers are compared with
SciPy's SLSQP on the CPU oracle under the explicit constraint matrix of host/controls.py::ChebyshevControl.compute_lincon."""
import importlib

import numpy as np
import pytest

from util import TEST_PARAMS, relerr

pytestmark = pytest.mark.gpu


def _setup(ocs, N=60, K=6, T=3.0):
    tspan = ocs.matlab_linspace(0.0, T, N + 1)
    integ = ocs.CudaRK4Integrator(tspan)
    ctl = ocs.ChebyshevControl(integ.t, K, 1)
    integ.attach_control(ctl)
    return tspan, integ, ctl


def test_chebyshev_solve_respects_control_bounds(ocs, oracle):
    from scipy.optimize import minimize
    tspan, integ, ctl = _setup(ocs)
    prob = ocs.TestOCProblem(TEST_PARAMS, [[0.0, 0.6]])
    cb = np.array([[0.0, 0.6]])
    Bt = 96
    x0 = np.linspace(0.6, 2.4, Bt).reshape(1, Bt)
    v0 = ctl.compute_initial_v(0.5)
    Lb, Ub = ctl.compute_nlp_bounds(cb)
    assert np.all(np.isinf(Lb)) and np.all(np.isinf(Ub))

    # without the constraints the Chebyshev solve leaves the bounds (the gap VERDICT r1 names)
    free = integ.single_shooting(prob, x0, v0, Lb, Ub, max_iter=20000, tol=1e-6)
    u_free = np.einsum("k,kj->j", free["v"][:, 0], ctl.B)
    assert free["converged"].all() and u_free.max() > 1.0

    # (a first-order method on an augmented Lagrangian: tens of thousands of cheap evaluations at this tolerance)
    integ.set_solve_control_bounds(cb, rho0=100.0, feas_tol=1e-6)
    out = integ.single_shooting(prob, x0, v0, Lb, Ub, max_iter=150000, tol=1e-5)
    viol = integ.solve_violation()
    u = np.einsum("kb,kj->bj", out["v"], ctl.B)
    assert out["converged"].all()
    assert viol.max() <= 1.0001e-6 and u.max() <= 0.6 + 2e-6 and u.min() >= -2e-6
    assert np.all(out["J"] >= free["J"] - 1e-9)                  # a constrained minimum is no better than the free one
    # the reported J is the objective of the returned coefficients, without any penalty term
    ref = oracle.C.evaluate(list(TEST_PARAMS.values()), tspan, ctl.B, x0, out["v"])
    assert relerr(out["J"], ref["J"]) < 1e-10

    A, b = ctl.compute_lincon(cb)
    for k in (0, Bt // 2, Bt - 1):
        def f(v):
            r = oracle.C.evaluate(list(TEST_PARAMS.values()), tspan, ctl.B, x0[:, k:k + 1], v.reshape(-1, 1))
            return float(r["J"][0]), r["dJdv"][:, 0].copy()
        sq = minimize(f, v0, jac=True, method="SLSQP", options=dict(ftol=1e-14, maxiter=500),
                      constraints=[dict(type="ineq", fun=lambda v: b - A @ v, jac=lambda v: -A)])
        assert sq.success
        assert abs(out["J"][k] - sq.fun) <= 1e-6 * abs(sq.fun), (k, out["J"][k], sq.fun)
        assert np.abs(u[k] - sq.x @ ctl.B).max() < 2e-3, k
    # switching the constraints off again restores the free solve
    integ.set_solve_control_bounds(None)
    again = integ.single_shooting(prob, x0, v0, Lb, Ub, max_iter=20000, tol=1e-6)
    assert np.array_equal(again["J"], free["J"])
    integ.close()


def test_driver_uses_the_constraints_for_chebyshev_and_two_controls(ocs):
    """host/drivers.single_shooting discovers compute_lincon-style constraints like the reference (:106-111); the
    generated 6-state problem with two controls stays inside its ControlBounds."""
    cg = importlib.import_module("optimal-control-solvers_b200.codegen")
    prob = cg.harvest_chain_problem(ControlBounds=((0.0, 0.5), (0.1, 0.4)))
    Bt, N = 64, 40
    rng = np.random.default_rng(3)
    x0 = rng.uniform(1.0, 2.5, (6, Bt))
    tspan = ocs.matlab_linspace(0.0, 2.0, N + 1)
    integ = ocs.CudaRK4Integrator(tspan)
    ctl = ocs.ChebyshevControl(integ.t, 5, 2)
    soln = ocs.single_shooting(prob, x0, tspan, 5, Control=ctl, Integrator=integ, u0=0.3, MaxIter=100000, TolX=1e-4)
    u = soln["u"]                                   # nc x G x Bt
    cb = np.asarray(prob.ControlBounds)
    # default feasibility tolerance of the constrained solve: 1e-4
    assert integ.solve_violation().max() <= 1e-4
    for c in range(2):
        assert u[c].min() >= cb[c, 0] - 1e-4 and u[c].max() <= cb[c, 1] + 1e-4
    assert (soln["converged"] == 1).mean() > 0.9
    # at least one bound is active somewhere: the constraints did something
    assert min(np.abs(u[0] - cb[0, 1]).min(), np.abs(u[1] - cb[1, 0]).min(), np.abs(u[1] - cb[1, 1]).min()) < 1e-4
    integ.close()
