"""The callers of the RK4 path through the host mirrors."""
import numpy as np
import pytest

from util import TEST_PARAMS, TOL, relerr

pytestmark = pytest.mark.gpu


def test_nlp_objective_with_free_initial_state(ocs, oracle):
    """single_shooting.m:143-149: the last entry of v is the free x0 and its gradient is lam(1,1)."""
    tspan = ocs.matlab_linspace(0.0, 10.0, 501)
    integ = ocs.CudaRK4Integrator(tspan)
    ctl = ocs.PWLinearControl(integ.t, 21, 1)
    prob = ocs.TestOCProblem(TEST_PARAMS, [[0, 1]])
    oprob, ointeg = oracle.TestOCProblem(**TEST_PARAMS), oracle.RK4Integrator(tspan)
    octl = oracle.PWLinearControl(ointeg.t, 21, 1)
    rng = np.random.default_rng(0)
    v = np.concatenate([.5 + .5 * rng.random(21), [1.3]])
    J, g = ocs.nlp_objective(prob, integ, ctl, [1.0], v, FreeInitStates=(0,))
    oJ, og = oracle.nlp_objective(oprob, ointeg, octl, [1.0], v, FreeInitStates=(0,))
    assert relerr(J, oJ) < TOL and relerr(g, og) < TOL and g.shape == (22,)
    # batch of 64 (v, x0) pairs
    V = np.vstack([.5 + .5 * rng.random((21, 64)), rng.uniform(.5, 2.5, (1, 64))])
    Jb, gb = ocs.nlp_objective(prob, integ, ctl, [1.0], V, FreeInitStates=(0,))
    for b in (0, 17, 63):
        oJ, og = oracle.nlp_objective(oprob, ointeg, octl, [1.0], V[:, b], FreeInitStates=(0,))
        assert relerr(Jb[b], oJ) < TOL and relerr(gb[:, b], og) < TOL
    # no free states
    J2, g2 = ocs.nlp_objective(prob, integ, ctl, [1.0], v[:-1])
    oJ2, og2 = oracle.nlp_objective(oprob, ointeg, octl, [1.0], v[:-1])
    assert relerr(J2, oJ2) < TOL and relerr(g2, og2) < TOL


def test_legacy_compute_states_and_objective(ocs, oracle):
    """functions/compute_states_and_objective.m with the k4 defect fixed, vs the oracle's restatement of
    the same legacy loop (uniform h, tHalf = t + h/2, h*(...)/6)."""
    t = ocs.matlab_linspace(0.0, 10.0, 401)
    rng = np.random.default_rng(1)
    u = rng.random((1, 401))
    prob = ocs.TestOCProblem(TEST_PARAMS, [[0, 1]])
    x, J = ocs.compute_states_and_objective(prob, 1.2, t, u)
    ox, oJ = oracle.compute_states_and_objective_legacy(oracle.TestOCProblem(**TEST_PARAMS), [1.2], t, u)
    assert x.shape == ox.shape
    assert relerr(J, oJ) < TOL and relerr(x, ox) < TOL


def test_fb_sweep_driver(ocs, oracle):
    tspan = ocs.matlab_linspace(0.0, 4.0, 201)
    integ = ocs.CudaRK4Integrator(tspan)
    prob = ocs.TestOCProblem(TEST_PARAMS, [[0, 1]])
    x0 = np.random.default_rng(2).uniform(.5, 2.5, (1, 32))
    soln = ocs.fb_sweep(prob, integ, x0, options=dict(u0=0.5, nSWEEPS=60))
    assert soln["converged"].all() and soln["u"].shape == (1, 401, 32)
    assert np.all(soln["u"] >= 0) and np.all(soln["u"] <= 1)
    oprob, ointeg = oracle.TestOCProblem(**TEST_PARAMS), oracle.RK4Integrator(tspan)
    cc = lambda t, x, lam: oracle.control_char_test_problem(oprob, t, x, lam)
    ref = oracle.fb_sweep_rk4(oprob, ointeg, x0[:, 5], np.full((1, 401), .5), cc, nSWEEPS=60)
    assert soln["sweeps"][5] == ref["sweeps"] and relerr(soln["J"][5], ref["J"]) < 1e-9
    # the optimality system holds at the fixed point: u = ControlChar(t, x, lam) on the grid
    x = integ.fetch_states(5, 1)[:, :, 0]
    lam, _ = integ.fetch_adjoints(5, 1, want_dJdu=False)
    xg, lg = oracle._nodes_to_grid(x[:-1]), oracle._nodes_to_grid(lam[:-1, :, 0])
    assert np.max(np.abs(cc(ointeg.t, xg, lg) - soln["u"][:, :, 5])) < 1e-5


def test_batched_single_shooting_matches_lbfgsb_on_the_oracle(ocs, oracle):
    """the device-side solve reaches the minimisers that a trusted CPU optimiser
    (SciPy L-BFGS-B) finds on the oracle's (J, dJdv) -- fmincon's iterates are not reproduced, its optimum is."""
    from scipy.optimize import minimize
    tspan = ocs.matlab_linspace(0.0, 10.0, 201)
    prob = ocs.TestOCProblem(TEST_PARAMS, [[0, 1]])
    x0 = np.random.default_rng(3).uniform(.5, 2.5, (1, 48))
    soln = ocs.single_shooting(prob, x0, tspan, 21, u0=0.5, TolX=1e-7, MaxIter=400)
    assert soln["converged"].all(), soln["converged"]
    assert np.all(soln["v"] >= 0) and np.all(soln["v"] <= 1)
    oprob, ointeg = oracle.TestOCProblem(**TEST_PARAMS), oracle.RK4Integrator(tspan)
    octl = oracle.PWLinearControl(ointeg.t, 21, 1)
    for b in (0, 11, 47):
        fun = lambda v: oracle.nlp_objective(oprob, ointeg, octl, x0[:, b], v)
        res = minimize(fun, np.full(21, .5), jac=True, method="L-BFGS-B", bounds=[(0, 1)] * 21,
                       options=dict(maxiter=500, ftol=1e-15, gtol=1e-9))
        assert abs(soln["J"][b] - res.fun) < 1e-7 * abs(res.fun), (soln["J"][b], res.fun)
        assert np.max(np.abs(soln["v"][:, b] - res.x)) < 2e-3
    # the device state is left at the minimisers: the reported J is the sweep's J, and the gradient is projected-zero
    integ, ctl = soln["integrator"], soln["control"]
    chk = integ.evaluate(prob, x0, soln["v"])
    assert relerr(chk["J"], soln["J"]) < 1e-14
    pg = soln["v"] - np.clip(soln["v"] - chk["dJdv"], 0, 1)
    assert np.max(np.abs(pg)) < 1e-6


def test_single_shooting_with_free_initial_state(ocs, oracle):
    from scipy.optimize import minimize
    tspan = ocs.matlab_linspace(0.0, 6.0, 121)
    prob = ocs.TestOCProblem(TEST_PARAMS, [[0, 1]])
    x0 = np.full((1, 8), 1.0)
    soln = ocs.single_shooting(prob, x0, tspan, 11, u0=0.5, FreeInitStates=(0,), FreeStateBounds=[[0.5, 2.0]],
                               TolX=1e-7, MaxIter=400)
    assert soln["v"].shape == (12, 8) and (soln["converged"] > 0).all()
    oprob, ointeg = oracle.TestOCProblem(**TEST_PARAMS), oracle.RK4Integrator(tspan)
    octl = oracle.PWLinearControl(ointeg.t, 11, 1)
    fun = lambda z: oracle.nlp_objective(oprob, ointeg, octl, [1.0], z, FreeInitStates=(0,))
    res = minimize(fun, np.concatenate([np.full(11, .5), [1.0]]), jac=True, method="L-BFGS-B",
                   bounds=[(0, 1)] * 11 + [(0.5, 2.0)], options=dict(maxiter=500, ftol=1e-15, gtol=1e-9))
    assert abs(soln["J"][0] - res.fun) < 1e-7 * abs(res.fun)
    assert abs(soln["v"][-1, 0] - res.x[-1]) < 1e-3
