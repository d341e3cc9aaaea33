"""The weight-folded sweeps (fast build: stage weights multiplied into the hoisted tables, csrc/problem_test.cuh)
against the unfolded chain and the oracle, and the split SPG update of the batched single_shooting solve against
the one-thread-per-trajectory kernel.

The adjoint sweep folds only while lam(end) == 1.  An explicit lamT = e_m (RK4Integrator.m:63-66 builds exactly
that when nargin < 4) takes the UNFOLDED chain of the same kernel, so comparing lamT = None with lamT = e_m compares
the two chains on identical inputs."""
import os

import numpy as np
import pytest

from util import TEST_PARAMS, TOL, relerr

pytestmark = pytest.mark.gpu
P = list(TEST_PARAMS.values())


@pytest.mark.parametrize("N", [1, 2, 63, 64, 65, 127, 128, 129, 200])
@pytest.mark.parametrize("per_traj_v", [False, True])
def test_folded_and_unfolded_adjoint_chains_agree(ocs, oracle, N, per_traj_v):
    tspan = ocs.matlab_linspace(0.0, 2.0 + 0.05 * N, N + 1)
    Bt, K = 300, min(7, N + 1)
    rng = np.random.default_rng(N)
    x0 = rng.uniform(.5, 2.5, (1, Bt))
    v = rng.random((K, Bt)) if per_traj_v else rng.random(K)
    prob = ocs.TestOCProblem(TEST_PARAMS, [[0, 1]])
    with ocs.CudaRK4Integrator(tspan) as integ:
        ctl = ocs.PWLinearControl(integ.t, K, 1)
        integ.attach_control(ctl)
        folded = integ.evaluate(prob, x0, v)
        unfolded = integ.evaluate(prob, x0, v, lamT=np.array([0.0, 1.0]))
        ref = oracle.C.evaluate(P, tspan, ctl.B, x0, v)
    for key in ("J", "xT", "dJdv", "lam0"):
        assert relerr(folded[key], ref[key]) < TOL, (key, N)
        assert relerr(unfolded[key], ref[key]) < TOL, (key, N)
        assert relerr(folded[key], unfolded[key]) < 1e-12, (key, N)
    assert np.all(folded["lam0"][1] == 1.0) and np.all(unfolded["lam0"][1] == 1.0)   # lam(end,:) never changes


def test_explicit_lamT_with_other_objective_weight(ocs, oracle):
    """lamT(end) != 1 scales the objective row: only the unfolded chain is valid, and it must be the one used."""
    tspan = ocs.matlab_linspace(0.0, 5.0, 131)
    rng = np.random.default_rng(5)
    x0 = rng.uniform(.5, 2.5, (1, 64))
    v = rng.random(9)
    lamT = np.array([0.3, 2.5])
    prob = ocs.TestOCProblem(TEST_PARAMS, [[0, 1]])
    with ocs.CudaRK4Integrator(tspan) as integ:
        ctl = ocs.PWLinearControl(integ.t, 9, 1)
        integ.attach_control(ctl)
        out = integ.evaluate(prob, x0, v, lamT=lamT)
    ointeg = oracle.RK4Integrator(tspan)
    oprob = oracle.TestOCProblem(**TEST_PARAMS)
    octl = oracle.PWLinearControl(ointeg.t, 9, 1)
    u = octl.compute_u(v)
    for b in (0, 31, 63):
        ointeg.compute_states(oprob, x0[:, b], u)
        lam, dJdu = ointeg.compute_adjoints(oprob, u, lamT)
        assert relerr(out["lam0"][:, b], lam[:, 0]) < TOL
        assert relerr(out["dJdv"][:, b], octl.compute_dJdv(dJdu)) < TOL


def test_infinite_horizon_folds_behind_the_tail(ocs, oracle):
    """RK4InfiniteIntegrator: the main horizon's lamT is the tail's lam(:,1) (last row still exactly 1), so both
    segments run the folded chain; parity against the oracle's two chained passes."""
    tspan = ocs.matlab_linspace(0.0, 6.0, 151)
    tspanE = ocs.matlab_linspace(6.0, 14.0, 91)
    uStar = 0.723368780097983
    rng = np.random.default_rng(6)
    x0 = rng.uniform(.5, 2.5, (1, 40))
    v = rng.random(11)
    prob = ocs.TestOCProblem(TEST_PARAMS, [[0, 1]])
    with ocs.CudaRK4InfiniteIntegrator(tspan, tspanE, uStar) as integ:
        ctl = ocs.PWLinearControl(integ.t, 11, 1)
        integ.attach_control(ctl)
        out = integ.evaluate(prob, x0, v)
    ointeg = oracle.RK4InfiniteIntegrator(tspan, tspanE, [uStar])
    oprob = oracle.TestOCProblem(**TEST_PARAMS)
    octl = oracle.PWLinearControl(ointeg.t, 11, 1)
    u = octl.compute_u(v)
    for b in (0, 13, 39):
        _, J = ointeg.compute_states(oprob, x0[:, b], u)
        lam, dJdu = ointeg.compute_adjoints(oprob, u)
        assert relerr(out["J"][b], J) < TOL
        assert relerr(out["lam0"][:, b], lam[:, 0]) < TOL
        assert relerr(out["dJdv"][:, b], octl.compute_dJdv(dJdu)) < TOL


def _solve(ocs, K, Bt, free=False, serial=False, out=None, max_iter=400):
    if serial:
        os.environ["OCS_B200_SPG_SERIAL"] = "1"
    try:
        tspan = ocs.matlab_linspace(0.0, 6.0, 121)
        prob = ocs.TestOCProblem(TEST_PARAMS, [[0, 1]])
        with ocs.CudaRK4Integrator(tspan) as integ:
            ctl = ocs.PWLinearControl(integ.t, K, 1)
            integ.attach_control(ctl)
            x0 = np.random.default_rng(K).uniform(.5, 2.5, (1, Bt))
            Lb, Ub = ctl.compute_nlp_bounds(prob.ControlBounds)
            kw = dict(FreeInitStates=(0,), FreeStateBounds=[[0.5, 2.0]]) if free else {}
            return integ.single_shooting(prob, x0, ctl.compute_initial_v(0.5), Lb, Ub, max_iter=max_iter, tol=1e-7, out=out, **kw)
    finally:
        os.environ.pop("OCS_B200_SPG_SERIAL", None)


@pytest.mark.parametrize("K,free", [(2, False), (5, True), (16, False), (17, False), (33, True), (101, False)])
def test_split_spg_update_matches_the_serial_kernel(ocs, K, free):
    """Same iteration, two kernels: rows split over thread groups (sums reduced through shared memory) vs one
    thread per trajectory.  They differ by the rounding of the Barzilai-Borwein sums only, so the minimisers agree
    far below the solver tolerance; nz = K + nfree covers one group, a ragged last group and several groups."""
    Bt = 77
    a = _solve(ocs, K, Bt, free=free)
    b = _solve(ocs, K, Bt, free=free, serial=True)
    assert a["converged"].all() and b["converged"].all()
    assert a["v"].shape == (K + (1 if free else 0), Bt)
    assert np.max(np.abs(a["v"] - b["v"])) < 1e-5
    assert relerr(a["J"], b["J"]) < 1e-10
    assert np.all(a["v"][:K] >= 0) and np.all(a["v"][:K] <= 1)
    if free:
        assert np.all(a["v"][K] >= 0.5) and np.all(a["v"][K] <= 2.0)


@pytest.mark.parametrize("variant", [("OCS_B200_SPG_L2", "0"), ("OCS_B200_SPG_L2", "2"), ("OCS_B200_SPG_UNROLL", "1"),
                                     ("OCS_B200_SPG_UNROLL", "4")])
@pytest.mark.parametrize("K,free,Bt", [(5, True, 77), (17, False, 130), (101, False, 300), (128, False, 65), (300, False, 40)])
def test_l2_spg_update_is_bit_identical_to_the_split_kernel(ocs, variant, K, free, Bt):
    """spg_update_l2_kernel (the default) re-reads the accepted rows from L2 instead of holding them in registers; same
    row groups and summation order as the register-held split kernel (OCS_B200_SPG_L2=0), so every bit agrees -- its
    64-trajectory and unroll variants too, ragged last blocks and a batch smaller than one block included (K = 300 is
    beyond the L2 kernel's 16 row groups: both runs take the split kernel)."""
    a = _solve(ocs, K, Bt, free=free)
    os.environ[variant[0]] = variant[1]
    try:
        b = _solve(ocs, K, Bt, free=free)
    finally:
        os.environ.pop(variant[0], None)
    for key in ("v", "J", "iters", "converged"):
        assert np.array_equal(a[key], b[key]), key


@pytest.mark.parametrize("K,free,Bt,max_iter", [(5, True, 77, 400), (101, False, 1500, 400), (33, False, 300, 6), (17, True, 40, 1)])
def test_vopt_download_under_the_last_evaluation(ocs, K, free, Bt, max_iter):
    """The minimisers are final before the last forward sweep (which only leaves the device state at them), so their
    transposition and download run on the copy stream under it -- into pinned and pageable arrays alike, also when the
    iteration limit stops the solve with trajectories still open."""
    import torch
    nz = K + (1 if free else 0)
    ref = _solve(ocs, K, Bt, free=free, max_iter=max_iter)
    assert ref["converged"].all() == (max_iter == 400)
    again = _solve(ocs, K, Bt, free=free, max_iter=max_iter)
    vp = torch.full((Bt, nz), float("nan"), dtype=torch.float64).pin_memory()
    bufs = dict(v=vp.numpy().T, J=np.empty(Bt), iters=np.zeros(Bt, dtype=np.int32), converged=np.zeros(Bt, dtype=np.int32))
    res = _solve(ocs, K, Bt, free=free, out=bufs, max_iter=max_iter)
    assert res["v"] is bufs["v"]
    for key in ("v", "J", "iters", "converged"):
        assert np.array_equal(ref[key], bufs[key]) and np.array_equal(ref[key], again[key]), key
    # the coefficients on the device are the ones returned
    assert np.all(np.isfinite(bufs["v"]))


def test_single_shooting_writes_into_caller_buffers(ocs):
    K, Bt = 9, 50
    bufs = dict(v=np.full((K, Bt), np.nan, order="F"), J=np.full(Bt, np.nan), iters=np.zeros(Bt, dtype=np.int32),
                converged=np.zeros(Bt, dtype=np.int32))
    res = _solve(ocs, K, Bt, out=bufs)
    assert res["v"] is bufs["v"] and np.isfinite(bufs["v"]).all() and np.isfinite(bufs["J"]).all()
    assert bufs["converged"].all() and (bufs["iters"] > 0).all()
    ref = _solve(ocs, K, Bt)
    assert np.array_equal(ref["v"], bufs["v"]) and np.array_equal(ref["J"], bufs["J"])
    with pytest.raises(ValueError):
        _solve(ocs, K, Bt, out=dict(bufs, v=np.empty((K, Bt))))        # row-major v is refused
