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
