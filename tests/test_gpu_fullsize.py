"""BASELINE.json's full sizes on the GPU, checked through size-independent properties:
a random sample of the 2^20 trajectories against the CPU oracle (a trajectory's result does not
depend on the batch around it), status words, the device-side J reductions, and the
adjoint == finite-difference property of tests/backprop_test.m on the batch."""
import numpy as np
import pytest

from util import TEST_PARAMS, TOL, relerr

pytestmark = pytest.mark.gpu


def test_cfg2_full_size_sample_vs_oracle(ocs, oracle):
    Bt, N, K = 1 << 20, 1000, 101
    tspan = ocs.matlab_linspace(0.0, 10.0, N + 1)
    integ = ocs.CudaRK4Integrator(tspan)
    ctl = ocs.PWLinearControl(integ.t, K, 1)
    integ.attach_control(ctl)
    prob = ocs.TestOCProblem(TEST_PARAMS, [[0, 1]])
    v = np.random.default_rng(1).random(K)
    x0 = np.random.default_rng(2).uniform(.5, 2.5, (1, Bt))
    out = integ.evaluate(prob, x0, v)
    assert not out["status"].any() and np.all(np.isfinite(out["J"]))
    idx = np.random.default_rng(3).choice(Bt, 256, replace=False)
    idx[:2] = [0, Bt - 1]
    ref = oracle.C.evaluate(list(TEST_PARAMS.values()), tspan, ctl.B, x0[:, idx], v)
    assert relerr(out["J"][idx], ref["J"]) < TOL and relerr(out["xT"][:, idx], ref["xT"]) < TOL
    assert relerr(out["dJdv"][:, idx], ref["dJdv"]) < TOL and relerr(out["lam0"][:, idx], ref["lam0"]) < TOL
    # device-side batch reductions (warp-shuffle tree)
    vmin, imin = integ.reduce_J("min")
    vsum, _ = integ.reduce_J("sum")
    assert vmin == out["J"].min() and out["J"][imin] == vmin
    assert abs(vsum - out["J"].sum()) < 1e-9 * abs(vsum)
    # batch-position independence: the same x0 anywhere in the batch gives the same bits
    x0b = x0.copy()
    x0b[0, 12345] = x0[0, 777_777]
    out2 = integ.evaluate(prob, x0b, v, gradient=False)
    assert out2["J"][12345] == out["J"][777_777]
    # grids of a sub-range of the big batch (the reference's return values, on request)
    integ.evaluate(prob, x0, v)
    x = integ.fetch_states(777_777, 1)
    lam, dJdu = integ.fetch_adjoints(777_777, 1)
    oi = oracle.RK4Integrator(tspan)
    op = oracle.TestOCProblem(**TEST_PARAMS)
    u = oracle.PWLinearControl(oi.t, K, 1).compute_u(v)
    ox, oJ = oi.compute_states(op, x0[:, 777_777], u)
    olam, odJdu = oi.compute_adjoints(op, u)
    assert relerr(x[:, :, 0], ox) < TOL and relerr(lam[:, :, 0], olam) < TOL and relerr(dJdu[:, :, 0], odJdu) < TOL


def test_cfg2_full_size_adjoint_equals_finite_difference(ocs):
    """tests/backprop_test.m's property at scale: central differences in one coefficient of a
    per-trajectory v agree with the adjoint gradient for every one of 2^18 trajectories."""
    Bt, N, K = 1 << 18, 1000, 21
    integ = ocs.CudaRK4Integrator(ocs.matlab_linspace(0.0, 10.0, N + 1))
    ctl = ocs.PWLinearControl(integ.t, K, 1)
    integ.attach_control(ctl)
    prob = ocs.TestOCProblem(TEST_PARAMS, [[0, 1]])
    rng = np.random.default_rng(5)
    v = .5 + .5 * rng.random((K, Bt))
    x0 = rng.uniform(.5, 2.5, (1, Bt))
    g = integ.evaluate(prob, x0, v)["dJdv"]
    eps, k = 1e-4, 7
    vp, vm = v.copy(), v.copy()
    vp[k] += eps
    vm[k] -= eps
    Jp = integ.evaluate(prob, x0, vp, gradient=False)["J"]
    Jm = integ.evaluate(prob, x0, vm, gradient=False)["J"]
    cd = (Jp - Jm) / (2 * eps)
    assert np.max(np.abs(cd - g[k])) < 2e-8 * max(1.0, np.max(np.abs(g[k])))


def test_cfg4_full_size_infinite_horizon_sample_vs_oracle(ocs, oracle):
    """BASELINE.json configs[3]: RK4InfiniteIntegrator, 262144 trajectories x (10000 + 10000) steps, 42 GB of
    node states in HBM; a sample against the C oracle's RK4InfiniteIntegrator."""
    Bt, N, K = 1 << 18, 10_000, 101
    uStar = 0.723368780097983
    ts, te = ocs.matlab_linspace(0.0, 100.0, N + 1), ocs.matlab_linspace(100.0, 200.0, N + 1)
    integ = ocs.CudaRK4InfiniteIntegrator(ts, te, uStar)
    ctl = ocs.PWLinearControl(integ.t, K, 1)
    integ.attach_control(ctl)
    prob = ocs.TestOCProblem(TEST_PARAMS, [[0, 1]])
    v = np.full(K, uStar) + 0.05 * np.random.default_rng(1).standard_normal(K)
    x0 = np.random.default_rng(4).uniform(.5, 2.5, (1, Bt))
    out = integ.evaluate(prob, x0, v)
    assert not out["status"].any()
    idx = np.random.default_rng(3).choice(Bt, 48, replace=False)
    idx[:2] = [0, Bt - 1]
    ref = oracle.C.evaluate(list(TEST_PARAMS.values()), ts, ctl.B, x0[:, idx], v, tspan_extra=te, uStar=uStar)
    assert relerr(out["J"][idx], ref["J"]) < TOL and relerr(out["xT"][:, idx], ref["xT"]) < TOL
    assert relerr(out["dJdv"][:, idx], ref["dJdv"]) < TOL and relerr(out["lam0"][:, idx], ref["lam0"]) < TOL
    integ.close()


def test_fb_sweep_full_size_properties(ocs, oracle):
    """The fb_sweep iteration on 2^20 trajectories x 1000 steps (16.8 GB of per-trajectory controls): every
    trajectory converges, converged ones are frozen (a second call changes nothing), the fixed point satisfies
    the control characterisation, and two sampled trajectories match the oracle's composition sweep for sweep."""
    Bt, N = 1 << 20, 1000
    tspan = ocs.matlab_linspace(0.0, 10.0, N + 1)
    integ = ocs.CudaRK4Integrator(tspan)
    prob = ocs.TestOCProblem(TEST_PARAMS, [[0, 1]])
    x0 = np.random.default_rng(2).uniform(.5, 2.5, (1, Bt))
    u0 = np.full((1, 2 * N + 1), .5)
    out = integ.fb_sweep(prob, x0, u0, nSWEEPS=50)
    assert out["converged"].all() and out["maxChange"].max() <= 1.0 and out["sweeps"].max() <= 50
    oprob, ointeg = oracle.TestOCProblem(**TEST_PARAMS), oracle.RK4Integrator(tspan)
    cc = lambda t, x, lam: oracle.control_char_test_problem(oprob, t, x, lam)
    for b in (0, Bt - 1):
        ref = oracle.fb_sweep_rk4(oprob, ointeg, x0[:, b], u0.copy(), cc, nSWEEPS=50)
        assert ref["converged"] and out["sweeps"][b] == ref["sweeps"]
        assert relerr(out["J"][b], ref["J"]) < 1e-9
        un = integ.get_control_u(b, 1)[:, :, 0]
        assert np.max(np.abs(un - ref["u"])) < 1e-6
        # fixed point: u = ControlChar(t, x(u), lam(u)) up to the sweep tolerance
        x, _ = ointeg.compute_states(oprob, x0[:, b], un)
        lam = ointeg.compute_adjoints(oprob, un, want_dJdu=False)
        fix = cc(ointeg.t, oracle._nodes_to_grid(x[:-1]), oracle._nodes_to_grid(lam[:-1]))
        assert np.max(np.abs(fix - un)) < 1e-6
    integ.close()


def test_cfg3_full_size_sample_vs_oracle(ocs, oracle):
    """BASELINE.json configs[2] at full size: the generated 6-state problem, 100k Chebyshev (degree 16) control
    candidates x 1000 steps; sampled candidates against the NumPy oracle (lambdified twin of the same
    expressions)."""
    import importlib
    from gen_util import oracle_problem
    cg = importlib.import_module("optimal-control-solvers_b200.codegen")
    prob = cg.harvest_chain_problem()
    oprob = oracle_problem(oracle, prob)
    Bt, N, K = 100_000, 1000, 17
    tspan = ocs.matlab_linspace(0.0, 10.0, N + 1)
    integ = ocs.CudaRK4Integrator(tspan)
    ctl = ocs.ChebyshevControl(integ.t, K, 2)
    integ.attach_control(ctl)
    rng = np.random.default_rng(3)
    x0 = rng.uniform(1.0, 2.5, (6, Bt))
    v = np.vstack([rng.uniform(.2, .8, (2, Bt)), 0.05 * rng.standard_normal((2 * (K - 1), Bt))])
    out = integ.evaluate(prob, x0, v)
    assert not out["status"].any()
    ointeg = oracle.RK4Integrator(tspan)
    octl = oracle.ChebyshevControl(ointeg.t, K, 2)
    for b in (0, 54_321, Bt - 1):
        u = octl.compute_u(v[:, b])
        x, J = ointeg.compute_states(oprob, x0[:, b], u)
        lam, dJdu = ointeg.compute_adjoints(oprob, u)
        assert relerr(out["J"][b], J) < TOL and relerr(out["xT"][:, b], x[:-1, -1]) < TOL
        assert relerr(out["dJdv"][:, b], octl.compute_dJdv(dJdu)) < TOL and relerr(out["lam0"][:, b], lam[:, 0]) < TOL
    integ.close()


@pytest.mark.parametrize("per_traj_v", [False, True])
def test_parameter_sweep_full_size_sample_vs_oracle(ocs, oracle, per_traj_v):
    """bench.py --config cfg2p / cfg2pv at full size: 2^20 trajectories, each with its own (c, m, r)
    (tests/TestOCProblem.m:13-19 -- exactly what a parameter sweep varies) x 1000 RK4 steps; nothing is
    batch-uniform, exp(-r t) runs in the thread.  A random sample against the C oracle."""
    Bt, N, K = 1 << 20, 1000, 101
    tspan = ocs.matlab_linspace(0.0, 10.0, N + 1)
    integ = ocs.CudaRK4Integrator(tspan)
    ctl = ocs.PWLinearControl(integ.t, K, 1)
    integ.attach_control(ctl)
    rng = np.random.default_rng(31)
    c, m, r = rng.uniform(1.3, 1.7, Bt), rng.uniform(2.8, 3.2, Bt), rng.uniform(0.03, 0.07, Bt)
    prob = ocs.TestOCProblem(dict(c=c, m=m, r=r), [[0, 1]])
    v = rng.random((K, Bt)) if per_traj_v else rng.random(K)
    x0 = rng.uniform(.5, 2.5, (1, Bt))
    out = integ.evaluate(prob, x0, v)
    assert not out["status"].any()
    idx = np.random.default_rng(3).choice(Bt, 192, replace=False)
    idx[:2] = [0, Bt - 1]
    ref = oracle.C.evaluate(np.stack([c[idx], m[idx], r[idx]]), tspan, ctl.B, x0[:, idx], v[:, idx] if per_traj_v else v)
    for key in ("J", "xT", "dJdv", "lam0"):
        assert relerr(out[key][..., idx], ref[key]) < TOL, key
    # the infinite-horizon composition with per-trajectory parameters: chunked (pipelined) host path, tail included
    integ.close()
    Bs = 1 << 18
    tsE = ocs.matlab_linspace(10.0, 20.0, 501)
    inf = ocs.CudaRK4InfiniteIntegrator(ocs.matlab_linspace(0.0, 10.0, 501), tsE, 0.72)
    ctl2 = ocs.PWLinearControl(inf.t, 21, 1)
    inf.attach_control(ctl2)
    prob2 = ocs.TestOCProblem(dict(c=c[:Bs], m=m[:Bs], r=r[:Bs]), [[0, 1]])
    v2 = rng.random(21)
    out2 = inf.evaluate(prob2, x0[:, :Bs], v2)
    sel = idx[idx < Bs][:64]
    ref2 = oracle.C.evaluate(np.stack([c[sel], m[sel], r[sel]]), ocs.matlab_linspace(0.0, 10.0, 501), ctl2.B, x0[:, sel], v2,
                             tspan_extra=tsE, uStar=0.72)
    for key in ("J", "xT", "dJdv", "lam0"):
        assert relerr(out2[key][..., sel], ref2[key]) < TOL, key
    inf.close()
