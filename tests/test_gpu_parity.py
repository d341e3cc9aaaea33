"""Parity of the CUDA sweeps (through the C ABI) with the CPU oracle on the same seeded
inputs.  Tolerance: 1e-10 normwise relative (BASELINE.json north_star); control-grid
indices bit-exact; the strict arithmetic mode is held to 5e-14."""
import numpy as np
import pytest

from util import TEST_PARAMS, TOL, TOL_STRICT, relerr

pytestmark = pytest.mark.gpu


def _setup(ocs, oracle, N=500, T=10.0, K=21, kind="pwl", math="fast"):
    tspan = ocs.matlab_linspace(0.0, T, N + 1)
    integ = ocs.CudaRK4Integrator(tspan, math=math)
    prob = ocs.TestOCProblem(TEST_PARAMS, [[0.0, 1.0]])
    oprob = oracle.TestOCProblem(**TEST_PARAMS)
    ointeg = oracle.RK4Integrator(tspan)
    cls = dict(pwl="PWLinearControl", pwc="PWConstantControl", cheb="ChebyshevControl")[kind]
    ctl = getattr(ocs, cls)(integ.t, K, 1)
    octl = getattr(oracle, cls)(ointeg.t, K, 1)
    return tspan, integ, prob, oprob, ointeg, ctl, octl


@pytest.mark.parametrize("math", ["fast", "strict"])
def test_backprop_scenario_single_trajectory(ocs, oracle, math):
    """tests/backprop_test.m:5-27 configuration through the literal Integrator API."""
    tspan, integ, prob, oprob, ointeg, ctl, octl = _setup(ocs, oracle, math=math)
    assert np.array_equal(integ.t, ointeg.t) and np.array_equal(integ.h, ointeg.h)
    assert np.array_equal(ctl.B, octl.B)
    v = .5 + .5 * np.random.default_rng(0).random(21)
    u = ctl.compute_u(v)
    x, J = integ.compute_states(prob, 1.0, u)
    lam, dJdu = integ.compute_adjoints(prob, u)
    ox, oJ = ointeg.compute_states(oprob, [1.0], octl.compute_u(v))
    olam, odJdu = ointeg.compute_adjoints(oprob, octl.compute_u(v))
    tol = TOL if math == "fast" else TOL_STRICT
    assert x.shape == ox.shape and lam.shape == olam.shape and dJdu.shape == odJdu.shape
    assert relerr(J, oJ) < tol
    assert relerr(x, ox) < tol
    assert relerr(lam, olam) < tol
    assert relerr(dJdu, odJdu) < tol
    assert np.all(lam[-1] == 1.0)
    dJdv = ctl.compute_dJdv(dJdu)
    assert relerr(dJdv, octl.compute_dJdv(odJdu)) < tol


@pytest.mark.parametrize("kind,K", [("pwl", 21), ("pwl", 101), ("pwc", 100), ("pwc", 7), ("cheb", 17)])
@pytest.mark.parametrize("math", ["fast", "strict"])
def test_fused_evaluate_shared_v(ocs, oracle, kind, K, math):
    """nlpObjective (functions/single_shooting.m:137-150) fused, shared v, random x0."""
    tspan, integ, prob, oprob, ointeg, ctl, octl = _setup(ocs, oracle, K=K, kind=kind, math=math)
    integ.attach_control(ctl)
    rng = np.random.default_rng(1)
    v = rng.random(K) if kind != "cheb" else np.concatenate([[0.5], 0.05 * rng.standard_normal(K - 1)])
    x0 = np.random.default_rng(2).uniform(.5, 2.5, (1, 777))
    out = integ.evaluate(prob, x0, v)
    ref = oracle.C.evaluate(list(TEST_PARAMS.values()), tspan, octl.B, x0, v)
    tol = TOL if math == "fast" else TOL_STRICT
    for key in ("J", "xT", "dJdv", "lam0"):
        assert relerr(out[key], ref[key]) < tol, key
    assert not out["status"].any()


@pytest.mark.parametrize("kind,K", [("pwl", 21), ("pwc", 100), ("cheb", 17)])
def test_fused_evaluate_per_trajectory_v(ocs, oracle, kind, K):
    tspan, integ, prob, oprob, ointeg, ctl, octl = _setup(ocs, oracle, N=200, K=K, kind=kind)
    integ.attach_control(ctl)
    Bt = 333
    rng = np.random.default_rng(3)
    if kind == "cheb":
        v = np.vstack([rng.random((1, Bt)), 0.05 * rng.standard_normal((K - 1, Bt))])
    else:
        v = rng.random((K, Bt))
    x0 = rng.uniform(.5, 2.5, (1, Bt))
    out = integ.evaluate(prob, x0, v)
    ref = oracle.C.evaluate(list(TEST_PARAMS.values()), tspan, octl.B, x0, v)
    for key in ("J", "xT", "dJdv", "lam0"):
        assert relerr(out[key], ref[key]) < TOL, key


def test_pwconstant_control_grid_indices_bit_exact(ocs, oracle):
    """The interval index of every grid point and the control values themselves are exact."""
    for (T, N, K) in [(10, 500, 100), (10, 500, 20), (10, 1000, 100), (10, 1000, 50), (7, 1000, 70)]:
        tspan, integ, prob, oprob, ointeg, ctl, octl = _setup(ocs, oracle, N=N, T=float(T), K=K, kind="pwc")
        integ.attach_control(ctl)
        k, nnz = integ.control_index()
        assert np.array_equal(k, octl.interval_index())
        assert np.all(nnz == 1)
        # exact-rational index floor(j*K/(2N))
        j = np.arange(2 * N + 1)
        assert np.array_equal(k, np.minimum((j * K) // (2 * N), K - 1))
        v = np.random.default_rng(5).random(K)
        integ.compute_states(prob, 1.0, v=v)
        u = integ.get_control_u()
        assert np.array_equal(u[0, :, 0], v[k])          # bit-exact copy of a v entry
        integ.close()


def test_per_trajectory_params(ocs, oracle):
    tspan = ocs.matlab_linspace(0.0, 10.0, 301)
    integ = ocs.CudaRK4Integrator(tspan)
    Bt = 200
    rng = np.random.default_rng(7)
    P = np.stack([rng.uniform(1, 2, Bt), rng.uniform(2.5, 3.5, Bt), rng.uniform(0.01, 0.1, Bt)])
    prob = ocs.TestOCProblem(dict(c=P[0], m=P[1], r=P[2]), [[0, 1]])
    ctl = ocs.PWLinearControl(integ.t, 31, 1)
    integ.attach_control(ctl)
    v = rng.random(31)
    x0 = rng.uniform(.5, 2.5, (1, Bt))
    out = integ.evaluate(prob, x0, v)
    ref = oracle.C.evaluate(P, tspan, ctl.B, x0, v)
    for key in ("J", "xT", "dJdv", "lam0"):
        assert relerr(out[key], ref[key]) < TOL, key


def test_infinite_horizon(ocs, oracle):
    """RK4InfiniteIntegrator.m:20-30: J = J1 + J2, lamT = lam2(:,1)."""
    uStar = 0.723368780097983
    tspan = ocs.matlab_linspace(0.0, 10.0, 501)
    tspanE = ocs.matlab_linspace(10.0, 20.0, 501)
    integ = ocs.CudaRK4InfiniteIntegrator(tspan, tspanE, uStar)
    prob = ocs.TestOCProblem(TEST_PARAMS, [[0, 1]])
    ctl = ocs.PWLinearControl(integ.t, 101, 1)
    integ.attach_control(ctl)
    rng = np.random.default_rng(4)
    x0 = rng.uniform(.5, 2.5, (1, 257))
    v = np.full(101, uStar)
    out = integ.evaluate(prob, x0, v)
    ref = oracle.C.evaluate(list(TEST_PARAMS.values()), tspan, ctl.B, x0, v, tspan_extra=tspanE, uStar=uStar)
    for key in ("J", "xT", "dJdv", "lam0"):
        assert relerr(out[key], ref[key]) < TOL, key
    # single trajectory through the literal API
    oprob = oracle.TestOCProblem(**TEST_PARAMS)
    oint = oracle.RK4InfiniteIntegrator(tspan, tspanE, uStar)
    u = ctl.compute_u(v)
    x, J = integ.compute_states(prob, 1.0, u)
    lam, dJdu = integ.compute_adjoints(prob, u)
    ox, oJ = oint.compute_states(oprob, [1.0], u)
    olam, odJdu = oint.compute_adjoints(oprob, u)
    assert relerr(J, oJ) < TOL and relerr(x, ox) < TOL
    assert relerr(lam, olam) < TOL and relerr(dJdu, odJdu) < TOL
    lam2, _ = integ.fetch_adjoints(segment=1, want_dJdu=False)
    assert relerr(lam2[:, :, 0], oint.lam2) < TOL


def test_batched_u_grid_and_lamT(ocs, oracle):
    """Integrator API with one control grid per trajectory and an explicit lamT (4th argument
    of compute_adjoints, RK4Integrator.m:59)."""
    tspan, integ, prob, oprob, ointeg, ctl, octl = _setup(ocs, oracle, N=100, K=11)
    rng = np.random.default_rng(9)
    Bt = 5
    u = rng.random((1, 201, Bt))
    x0 = rng.uniform(.5, 2.5, (1, Bt))
    lamT = np.vstack([rng.standard_normal((1, Bt)), np.ones((1, Bt))])
    x, J = integ.compute_states(prob, x0, u, return_x=True)
    lam, dJdu = integ.compute_adjoints(prob, u, lamT, grids=True)
    for b in range(Bt):
        ox, oJ = ointeg.compute_states(oprob, x0[:, b], u[:, :, b])
        olam, odJdu = ointeg.compute_adjoints(oprob, u[:, :, b], lamT[:, b])
        assert relerr(J[b], oJ) < TOL and relerr(x[:, :, b], ox) < TOL
        assert relerr(lam[:, :, b], olam) < TOL and relerr(dJdu[:, :, b], odJdu) < TOL


def test_fb_sweep_rk4_grid(ocs, oracle):
    """fb_sweep.m:79-115 composed on the RK4 grid vs the oracle's composition."""
    tspan, integ, prob, oprob, ointeg, ctl, octl = _setup(ocs, oracle, N=200, T=4.0)
    G = 401
    u0 = np.full((1, G), 0.5)
    cc = lambda t, x, lam: oracle.control_char_test_problem(oprob, t, x, lam)
    rng = np.random.default_rng(11)
    x0 = rng.uniform(.5, 2.5, (1, 16))
    out = integ.fb_sweep(prob, x0, u0, nSWEEPS=60)
    for b in range(0, 16, 5):
        ref = oracle.fb_sweep_rk4(oprob, ointeg, x0[:, b], u0.copy(), cc, nSWEEPS=60)
        assert bool(out["converged"][b]) == ref["converged"]
        assert out["sweeps"][b] == ref["sweeps"]
        assert relerr(out["J"][b], ref["J"]) < 1e-9
        un = integ.get_control_u(b, 1)[:, :, 0]
        assert np.max(np.abs(un - ref["u"])) < 1e-6   # soln.u = ControlChar of the final sweep (fb_sweep.m:123)


def test_fb_sweep_continuous_costate(ocs, oracle):
    """fb_sweep with the v1 API's costate equation (adjointRHS, compute_x_lam.m:11-12) integrated
    backwards by RK4 on the grid, vs the oracle's costate_rk4_continuous composition."""
    tspan, integ, prob, oprob, ointeg, ctl, octl = _setup(ocs, oracle, N=200, T=4.0)
    G = 401
    u0 = np.full((1, G), 0.5)
    cc = lambda t, x, lam: oracle.control_char_test_problem(oprob, t, x, lam)
    x0 = np.random.default_rng(12).uniform(.5, 2.5, (1, 333))
    one = integ.fb_sweep(prob, x0, u0, nSWEEPS=1, costate="continuous")       # a single sweep: pure kernel parity
    for b in (0, 77, 332):
        ref = oracle.fb_sweep_rk4(oprob, ointeg, x0[:, b], u0.copy(), cc, nSWEEPS=1, costate="continuous")
        un = integ.get_control_u(b, 1)[:, :, 0]
        assert relerr(un, ref["u"]) < TOL
        assert relerr(one["maxChange"][b], ref["maxChange"]) < 1e-9
    out = integ.fb_sweep(prob, x0, u0, nSWEEPS=60, costate="continuous")
    dis = integ.fb_sweep(prob, x0, u0, nSWEEPS=60, costate="discrete")
    for b in (0, 77, 332):
        ref = oracle.fb_sweep_rk4(oprob, ointeg, x0[:, b], u0.copy(), cc, nSWEEPS=60, costate="continuous")
        assert bool(out["converged"][b]) == ref["converged"] and out["sweeps"][b] == ref["sweeps"]
        assert relerr(out["J"][b], ref["J"]) < 1e-9
    # both costates approximate the same continuous problem: the optimal J agree to discretisation accuracy
    assert np.max(np.abs(out["J"] - dis["J"]) / np.abs(dis["J"])) < 1e-4
    with pytest.raises(ValueError):
        integ.fb_sweep(prob, x0, u0, costate="odevr7")


def test_reduce_J(ocs, oracle):
    tspan, integ, prob, oprob, ointeg, ctl, octl = _setup(ocs, oracle, N=50, K=11)
    integ.attach_control(ctl)
    x0 = np.random.default_rng(6).uniform(.5, 2.5, (1, 5000))
    out = integ.evaluate(prob, x0, np.full(11, .5), gradient=False)
    vmin, imin = integ.reduce_J("min")
    vmax, imax = integ.reduce_J("max")
    vsum, _ = integ.reduce_J("sum")
    assert vmin == out["J"].min() and imin == int(np.argmin(out["J"]))
    assert vmax == out["J"].max() and imax == int(np.argmax(out["J"]))
    assert abs(vsum - out["J"].sum()) < 1e-9 * abs(vsum)


def test_error_conventions(ocs):
    integ = ocs.CudaRK4Integrator(ocs.matlab_linspace(0, 1, 11))
    prob = ocs.TestOCProblem(TEST_PARAMS, [[0, 1]])
    with pytest.raises(ocs.OcsError) as e:
        integ._bind(prob)
        integ.compute_adjoints(prob)           # before compute_states
    assert e.value.code == ocs._lib.ERR_STATE
    with pytest.raises(ValueError):
        integ.compute_states(prob, 1.0, np.zeros((1, 5)))


@pytest.mark.parametrize("kind,K", [("pwl", 11), ("pwc", 8), ("cheb", 6)])
def test_fb_sweep_started_from_coefficients(ocs, oracle, kind, K):
    """fb_sweep over control CANDIDATES: u0 = control.compute_u(v) evaluated on the device for every trajectory
    (ocs_b200_fb_sweep with u0 = NULL) equals handing the same grids over from the host, and the oracle."""
    tspan, integ, prob, oprob, ointeg, ctl, octl = _setup(ocs, oracle, N=120, T=3.0, K=K, kind=kind)
    integ.attach_control(ctl)
    rng = np.random.default_rng(14)
    Bt = 130
    x0 = rng.uniform(.5, 2.5, (1, Bt))
    v = rng.uniform(.1, .9, (K, Bt)) if kind != "cheb" else np.vstack([rng.uniform(.3, .7, (1, Bt)), .05 * rng.standard_normal((K - 1, Bt))])
    u0 = np.stack([ctl.compute_u(v[:, b]) for b in range(Bt)], axis=2)           # nc x G x Bt on the host
    a = integ.fb_sweep(prob, x0, v0=v, nSWEEPS=1)
    ua = integ.get_control_u(0, Bt)
    b_ = integ.fb_sweep(prob, x0, u0, nSWEEPS=1)
    ub = integ.get_control_u(0, Bt)
    assert relerr(ua, ub) < TOL and relerr(a["J"], b_["J"]) < TOL and relerr(a["maxChange"], b_["maxChange"]) < 1e-8
    cc = lambda t, x, lam: oracle.control_char_test_problem(oprob, t, x, lam)
    for b in (0, 129):
        ref = oracle.fb_sweep_rk4(oprob, ointeg, x0[:, b], octl.compute_u(v[:, b]), cc, nSWEEPS=1)
        assert relerr(ua[:, :, b], ref["u"]) < TOL and relerr(a["J"][b], ref["J"]) < TOL
    # a shared coefficient vector works too
    s = integ.fb_sweep(prob, x0, v0=v[:, 0], nSWEEPS=1)
    assert relerr(s["J"][0], a["J"][0]) < TOL
    with pytest.raises(ValueError):
        integ.fb_sweep(prob, x0)
