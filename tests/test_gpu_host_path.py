"""The host-buffer path as a MATLAB caller drives it: pageable arrays (mxArrays are never pinned) go through the
handle's pinned ring and copy threads and must deliver bit-identical results to the plain path; and the handle stays
consistent across failed or out-of-order calls (a stale batch is never swept)."""
import ctypes

import numpy as np
import pytest

from util import TEST_PARAMS, TOL, relerr

pytestmark = pytest.mark.gpu


def _setup(ocs, N=100, K=21):
    integ = ocs.CudaRK4Integrator(ocs.matlab_linspace(0.0, 4.0, N + 1))
    ctl = ocs.PWLinearControl(integ.t, K, 1)
    integ.attach_control(ctl)
    return integ, ctl, ocs.TestOCProblem(TEST_PARAMS, [[0, 1]])


def test_pageable_buffers_take_the_pinned_ring_and_agree_with_the_plain_path(ocs, monkeypatch):
    """Bt large enough for the chunked, overlapped gradient download (>= 2^18) with 44 MB of dJdv: NumPy arrays
    (pageable) through the ring, with several slot sizes, vs the driver's own pageable path (OCS_B200_STAGER=0)."""
    Bt = 1 << 18
    integ, ctl, prob = _setup(ocs)
    rng = np.random.default_rng(0)
    x0, v = rng.uniform(.5, 2.5, (1, Bt)), rng.random(21)
    monkeypatch.setenv("OCS_B200_STAGER", "0")
    ref = integ.evaluate(prob, x0, v)
    monkeypatch.setenv("OCS_B200_STAGER", "1")
    out = integ.evaluate(prob, x0, v)
    for key in ("J", "xT", "status", "dJdv", "lam0"):
        assert np.array_equal(out[key], ref[key]), key
    # the two-call interface and the grids take the same route
    _, J = integ.compute_states(prob, x0, v=v)
    lam0, dJdv = integ.compute_adjoints(prob, grids=False)
    assert np.array_equal(J, ref["J"]) and np.array_equal(dJdv, ref["dJdv"]) and np.array_equal(lam0, ref["lam0"])
    x = integ.fetch_states(1000, 3000)            # 2 x 101 x 3000 doubles = 4.8 MB
    monkeypatch.setenv("OCS_B200_STAGER", "0")
    assert np.array_equal(x, integ.fetch_states(1000, 3000))
    integ.close()


def test_pageable_upload_of_per_trajectory_coefficients(ocs, monkeypatch):
    Bt = 1 << 16
    integ, ctl, prob = _setup(ocs)
    rng = np.random.default_rng(1)
    x0, v = rng.uniform(.5, 2.5, (1, Bt)), rng.random((21, Bt))        # 11 MB of coefficients up
    monkeypatch.setenv("OCS_B200_STAGER", "0")
    ref = integ.evaluate(prob, x0, v)
    monkeypatch.setenv("OCS_B200_STAGER", "1")
    monkeypatch.setenv("OCS_B200_STAGER_SLOT_MB", "1")
    integ2, _, _ = _setup(ocs)
    out = integ2.evaluate(prob, x0, v)
    for key in ("J", "dJdv"):
        assert np.array_equal(out[key], ref[key]), key
    integ.close()
    integ2.close()


def test_pinned_buffers_bypass_the_ring(ocs):
    import torch
    Bt = 1 << 18
    integ, ctl, prob = _setup(ocs)
    L = ocs._lib
    lib = L.load()
    integ._bind(prob)
    integ._set_v(np.random.default_rng(2).random(21))
    x0 = torch.rand((1, Bt), dtype=torch.float64).mul_(2).add_(.5).pin_memory()
    J = torch.empty(Bt, dtype=torch.float64).pin_memory()
    g = torch.empty((Bt, 21), dtype=torch.float64).pin_memory()
    L.check(lib.ocs_b200_evaluate(integ._h, L.dp(x0.numpy()), Bt, None, 0, L.dp(J.numpy()), None, None, L.dp(g.numpy()), None), integ._h)
    ref = integ.evaluate(prob, x0.numpy().copy(), np.random.default_rng(2).random(21))
    assert np.array_equal(J.numpy(), ref["J"]) and np.array_equal(g.numpy().T, ref["dJdv"])
    integ.close()


def test_get_dims(ocs):
    integ = ocs.CudaRK4InfiniteIntegrator(ocs.matlab_linspace(0, 1, 11), ocs.matlab_linspace(1, 2, 6), [.7])
    lib = ocs._lib.load()
    d = np.zeros(8, dtype=np.int32)
    assert lib.ocs_b200_get_dims(integ._h, ocs._lib.ip(d)) == 0
    assert list(d[:6]) == [0, 0, 0, 0, 10, 5]
    integ.attach_control(ocs.PWConstantControl(integ.t, 5, 1))
    integ.evaluate(ocs.TestOCProblem(TEST_PARAMS, [[0, 1]]), np.ones((1, 7)), np.full(5, .5))
    assert lib.ocs_b200_get_dims(integ._h, ocs._lib.ip(d)) == 0
    assert list(d[:6]) == [1, 1, 3, 5, 10, 5] and lib.ocs_b200_get_batch(integ._h) == 7
    integ.close()


def test_a_setter_between_compute_states_and_compute_adjoints_is_caught(ocs):
    """compute_adjoints sweeps the resident batch with the batch stride Bt: a control / parameter set for another
    batch size in between must be refused, not read out of bounds (RK4Integrator.m:8-10 statefulness contract)."""
    integ, ctl, prob = _setup(ocs)
    rng = np.random.default_rng(3)
    x0 = rng.uniform(.5, 2.5, (1, 300))
    integ.compute_states(prob, x0, v=rng.random((21, 300)))
    integ._set_v(rng.random((21, 100)))
    with pytest.raises(ocs.OcsError) as e:
        integ.compute_adjoints(prob, grids=False)
    assert e.value.code == ocs._lib.ERR_ARG
    with pytest.raises(ocs.OcsError):
        integ.fetch_states(0, 10)
    # a new basis detaches the coefficients: the resident sweep is gone
    integ.compute_states(prob, x0, v=rng.random((21, 300)))
    integ.attach_control(ocs.PWLinearControl(integ.t, 5, 1))
    with pytest.raises(ocs.OcsError) as e:
        integ.compute_adjoints(prob, grids=False)
    assert e.value.code == ocs._lib.ERR_STATE
    # per-trajectory parameters of another batch size
    integ.attach_control(ctl)
    integ.compute_states(prob, x0, v=rng.random(21))
    sweep = ocs.TestOCProblem(dict(c=rng.uniform(1, 2, 50), m=3.0, r=.05), [[0, 1]])
    with pytest.raises(ocs.OcsError):
        integ.compute_adjoints(sweep, grids=False)
    integ.close()


def test_a_failed_call_does_not_leave_a_stale_batch(ocs):
    integ, ctl, prob = _setup(ocs)
    rng = np.random.default_rng(4)
    integ.compute_states(prob, rng.uniform(.5, 2.5, (1, 64)), v=rng.random(21))
    integ._set_v(rng.random((21, 64)))
    x0big = ocs._lib.f64(rng.uniform(.5, 2.5, (1, 128)))
    with pytest.raises(ocs.OcsError):          # 128 trajectories, coefficients for 64
        ocs._lib.check(integ._lib.ocs_b200_compute_states(integ._h, ocs._lib.dp(x0big), 128, None, None, None), integ._h)
    with pytest.raises(ocs.OcsError) as e:
        integ.compute_adjoints(prob, grids=False)
    assert e.value.code == ocs._lib.ERR_STATE
    assert integ._lib.ocs_b200_get_batch(integ._h) == 0
    integ.close()


def test_fetch_states_after_fb_sweep_reads_the_resident_states(ocs, oracle):
    """fb_sweep.m:82 final_sweep(u): soln.x belongs to the control the last sweep STARTED from, soln.u is
    ControlChar of it.  After one sweep the resident states are therefore x(u0), not x(uNew)."""
    N, Bt = 60, 40
    tspan = ocs.matlab_linspace(0.0, 3.0, N + 1)
    integ = ocs.CudaRK4Integrator(tspan)
    prob = ocs.TestOCProblem(TEST_PARAMS, [[0, 1]])
    rng = np.random.default_rng(6)
    x0 = rng.uniform(.5, 2.5, (1, Bt))
    u0 = rng.uniform(0, 1, (1, 2 * N + 1, Bt))
    out = integ.fb_sweep(prob, x0, u0, nSWEEPS=1)
    x = integ.fetch_states(0, Bt)
    un = integ.get_control_u(0, Bt)
    assert not np.array_equal(un, u0)
    ointeg, oprob = oracle.RK4Integrator(tspan), oracle.TestOCProblem(**TEST_PARAMS)
    for b in (0, 17, 39):
        ox, oJ = ointeg.compute_states(oprob, x0[:, b], u0[:, :, b])
        assert relerr(x[0, :, b], ox[0]) < TOL and relerr(out["J"][b], oJ) < TOL
        assert np.isnan(x[1, :-1, b]).all() and x[1, -1, b] == out["J"][b]
    # the interpolated solution reads the same store
    sol = integ.eval_solution(tspan, 0, Bt, want=("x",))
    assert relerr(sol["x"][0], x[0]) < 1e-14
    integ.close()


def test_costate_fetch_after_a_continuous_costate_solve(ocs, oracle):
    """soln.lam of a continuous-costate fb_sweep solve is that costate (compute_x_lam.m:11-14), not the discrete
    adjoint; dJdu does not exist for it."""
    N, Bt = 80, 12
    tspan = ocs.matlab_linspace(0.0, 4.0, N + 1)
    integ = ocs.CudaRK4Integrator(tspan)
    prob = ocs.TestOCProblem(TEST_PARAMS, [[0, 1]])
    x0 = np.random.default_rng(8).uniform(.5, 2.5, (1, Bt))
    out = integ.fb_sweep(prob, x0, np.full((1, 2 * N + 1), .5), nSWEEPS=80, costate="continuous")
    assert out["converged"].all()
    lam, _ = integ.fetch_adjoints(0, Bt, want_dJdu=False)
    x = integ.fetch_states(0, Bt)
    u = integ.get_control_u(0, Bt)
    ointeg, oprob = oracle.RK4Integrator(tspan), oracle.TestOCProblem(**TEST_PARAMS)
    for b in (0, 11):
        ol = oracle.costate_rk4_continuous(oprob, ointeg, x[:1, :, b], u[:, :, b])
        assert relerr(lam[:1, :, b], ol) < 1e-9          # the oracle sweeps (x(u), uNew) too; converged: |uNew - u| <= 1e-7 scale
        assert np.all(lam[1, :, b] == 1.0)
    with pytest.raises(ocs.OcsError):
        integ.fetch_adjoints(0, Bt, want_dJdu=True)
    integ.close()
