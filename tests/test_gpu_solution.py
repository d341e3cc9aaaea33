"""soln.x / soln.lam / soln.u (functions/single_shooting.m:128-130, fb_sweep.m:120-124) evaluated on the device
for a batch (ocs_b200_eval_solution) against the oracle's restatements of vectorInterpolant / pchip /
compute_uFunc applied to the oracle's own grids."""
import numpy as np
import pytest

from util import TEST_PARAMS, TOL, relerr

pytestmark = pytest.mark.gpu


def _setup(ocs, oracle, N, T, K, kind):
    tspan = ocs.matlab_linspace(0.0, T, N + 1)
    integ = ocs.CudaRK4Integrator(tspan)
    prob = ocs.TestOCProblem(TEST_PARAMS, [[0.0, 1.0]])
    oprob = oracle.TestOCProblem(**TEST_PARAMS)
    ointeg = oracle.RK4Integrator(tspan)
    ctl = getattr(ocs, kind)(integ.t, K, 1)
    octl = getattr(oracle, kind)(ointeg.t, K, 1)
    return tspan, integ, prob, oprob, ointeg, ctl, octl


def _queries(tspan, nq, rng):
    """nodes, midpoints, random interior points and the two ends -- unsorted on purpose"""
    tq = np.concatenate([tspan[:: max(1, tspan.size // 7)], rng.uniform(tspan[0], tspan[-1], nq), [tspan[0], tspan[-1]]])
    return rng.permutation(tq)


@pytest.mark.parametrize("kind,K", [("PWLinearControl", 21), ("PWConstantControl", 10), ("ChebyshevControl", 9)])
def test_soln_of_a_batch(ocs, oracle, kind, K):
    tspan, integ, prob, oprob, ointeg, ctl, octl = _setup(ocs, oracle, 200, 5.0, K, kind)
    integ.attach_control(ctl)
    rng = np.random.default_rng(31)
    Bt = 333
    x0 = rng.uniform(.5, 2.5, (1, Bt))
    v = rng.random((K, Bt)) if kind != "ChebyshevControl" else np.vstack([rng.uniform(.3, .7, (1, Bt)), .05 * rng.standard_normal((K - 1, Bt))])
    integ.evaluate(prob, x0, v)
    tq = _queries(tspan, 300, rng)
    s = integ.eval_solution(tq)
    assert s["x"].shape == (1, tq.size, Bt) and s["lam"].shape == (1, tq.size, Bt) and s["u"].shape == (1, tq.size, Bt)
    for b in (0, 100, 332):
        u = octl.compute_u(v[:, b])
        ox, _ = ointeg.compute_states(oprob, x0[:, b], u)
        olam = ointeg.compute_adjoints(oprob, u, want_dJdu=False)
        assert relerr(s["x"][:, :, b], oracle.vector_interpolant(tspan, ox[:-1], "pchip")(tq)) < TOL
        assert relerr(s["lam"][:, :, b], oracle.vector_interpolant(tspan, olam[:-1], "pchip")(tq)) < TOL
        assert relerr(s["u"][:, :, b], oracle.compute_uFunc(octl, v[:, b])(tq)) < TOL
    # a sub-range and a subset of the outputs
    part = integ.eval_solution(tq[:17], b0=100, nb=5, want=("x",))
    assert set(part) == {"x"} and np.array_equal(part["x"], s["x"][:, :, 100:105][:, [list(tq).index(t) for t in tq[:17]]])
    # the host mirror's compute_uFunc agrees with the device's
    assert relerr(s["u"][:, :, 7], ctl.compute_uFunc(v[:, 7])(tq)) < TOL


def test_soln_single_trajectory_shared_v_and_short_grids(ocs, oracle):
    """the literal single-trajectory use (one x0, one v) and the 2- and 3-node grids of pchip"""
    for N in (1, 2, 5):
        tspan, integ, prob, oprob, ointeg, ctl, octl = _setup(ocs, oracle, N, 1.0, 3, "PWLinearControl")
        integ.attach_control(ctl)
        v = np.array([.2, .9, .4])
        integ.evaluate(prob, 1.3, v)
        tq = np.linspace(0, 1, 23)
        s = integ.eval_solution(tq)
        assert s["x"].shape == (1, 23)
        u = octl.compute_u(v)
        ox, _ = ointeg.compute_states(oprob, [1.3], u)
        olam = ointeg.compute_adjoints(oprob, u, want_dJdu=False)
        assert relerr(s["x"], oracle.vector_interpolant(tspan, ox[:-1], "pchip")(tq)) < TOL
        assert relerr(s["lam"], oracle.vector_interpolant(tspan, olam[:-1], "pchip")(tq)) < TOL
        assert relerr(s["u"], oracle.compute_uFunc(octl, v)(tq)) < TOL
        integ.close()


def test_soln_after_fb_sweep_and_on_the_tail(ocs, oracle):
    # fb_sweep: soln.u is the pchip interpolant of the control's grid values (fb_sweep.m:123)
    tspan, integ, prob, oprob, ointeg, ctl, octl = _setup(ocs, oracle, 100, 4.0, 5, "PWLinearControl")
    x0 = np.random.default_rng(2).uniform(.5, 2.5, (1, 40))
    integ.fb_sweep(prob, x0, np.full((1, 201), .5), nSWEEPS=60)
    tq = np.linspace(0, 4, 77)
    s = integ.eval_solution(tq)
    for b in (0, 39):
        ug = integ.get_control_u(b, 1)[:, :, 0]
        assert relerr(s["u"][:, :, b], oracle.vector_interpolant(integ.t, ug, "pchip")(tq)) < TOL
        ox, _ = ointeg.compute_states(oprob, x0[:, b], ug)
        # soln.x belongs to the sweep that met the tolerance, i.e. to the control BEFORE its last update
        # (fb_sweep.m:82,120: final_sweep(u) with the old u), so it matches x(ug) to the sweep tolerance only
        assert relerr(s["x"][:, :, b], oracle.vector_interpolant(tspan, ox[:-1], "pchip")(tq)) < 1e-6
    integ.close()
    # infinite horizon: both segments
    ts, te = ocs.matlab_linspace(0, 5, 101), ocs.matlab_linspace(5, 10, 81)
    uStar = 0.723368780097983
    inf = ocs.CudaRK4InfiniteIntegrator(ts, te, uStar)
    oinf = oracle.RK4InfiniteIntegrator(ts, te, uStar)
    u = np.full((1, 201), .4)
    inf.compute_states(prob, 1.0, u)
    inf.compute_adjoints(prob, u)
    ox, _ = oinf.compute_states(oprob, [1.0], u)
    olam = oinf.compute_adjoints(oprob, u, want_dJdu=False)
    ox2, _ = oinf.integrator2.compute_states(oprob, ox[:-1, -1], oinf.uStar)
    q0, q1 = np.linspace(0, 5, 33), np.linspace(5, 10, 33)
    s0 = inf.eval_solution(q0, want=("x", "lam"))
    s1 = inf.eval_solution(q1, segment=1, want=("x", "lam"))
    assert relerr(s0["x"], oracle.vector_interpolant(ts, ox[:-1], "pchip")(q0)) < TOL
    assert relerr(s0["lam"], oracle.vector_interpolant(ts, olam[:-1], "pchip")(q0)) < TOL
    assert relerr(s1["x"], oracle.vector_interpolant(te, ox2[:-1], "pchip")(q1)) < TOL
    assert relerr(s1["lam"], oracle.vector_interpolant(te, oinf.lam2[:-1], "pchip")(q1)) < TOL
    with pytest.raises(ocs.OcsError):
        inf.eval_solution(np.array([np.nan]))
    inf.close()
