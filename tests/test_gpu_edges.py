"""Edge cases of the batched sweeps: smallest and ragged sizes, tiles that do not divide N, single
knots, divergent trajectories (status word), argument validation.  All against the CPU oracle."""
import ctypes

import numpy as np
import pytest

from util import TEST_PARAMS, TOL, relerr

pytestmark = pytest.mark.gpu
P = list(TEST_PARAMS.values())


def _eval_both(ocs, oracle, N, T, kind, K, Bt, seed=0, per_traj_v=False):
    tspan = ocs.matlab_linspace(0.0, T, N + 1)
    integ = ocs.CudaRK4Integrator(tspan)
    ctl = getattr(ocs, kind)(integ.t, K, 1)
    integ.attach_control(ctl)
    rng = np.random.default_rng(seed)
    x0 = rng.uniform(.5, 2.5, (1, Bt))
    v = rng.random((K, Bt)) if per_traj_v else rng.random(K)
    if kind == "ChebyshevControl":
        v = v * 0.1
    out = integ.evaluate(ocs.TestOCProblem(TEST_PARAMS, [[0, 1]]), x0, v)
    ref = oracle.C.evaluate(P, tspan, ctl.B, x0, v)
    integ.close()
    return out, ref


@pytest.mark.parametrize("N", [1, 2, 3, 4, 5, 6, 7, 31, 32, 33, 63, 64, 65, 95, 97])
@pytest.mark.parametrize("per_traj_v", [False, True])
def test_step_counts_around_the_tile_size(ocs, oracle, N, per_traj_v):
    """N below, at and just above multiples of the 32-step shared-memory tile, and N = 1
    (one RK4 step: the reverse sweep's prefetch reads only slack)."""
    out, ref = _eval_both(ocs, oracle, N, 1.0 + 0.1 * N, "PWLinearControl", min(5, N + 1), 50, per_traj_v=per_traj_v)
    for key in ("J", "xT", "dJdv", "lam0"):
        assert relerr(out[key], ref[key]) < TOL, (key, N)


@pytest.mark.parametrize("Bt", [1, 2, 31, 127, 128, 129, 255, 256, 257, 383, 384, 385, 769])
def test_ragged_batch_sizes(ocs, oracle, Bt):
    """Batches around the block tilings (128 threads x 2 or 3 trajectories per thread)."""
    out, ref = _eval_both(ocs, oracle, 40, 2.0, "PWLinearControl", 7, Bt, seed=Bt)
    for key in ("J", "xT", "dJdv", "lam0"):
        assert out[key].shape == ref[key].shape
        assert relerr(out[key], ref[key]) < TOL, (key, Bt)


@pytest.mark.parametrize("kind,K", [("PWConstantControl", 1), ("PWConstantControl", 2), ("PWLinearControl", 2),
                                    ("PWLinearControl", 3), ("ChebyshevControl", 1), ("ChebyshevControl", 2),
                                    ("ChebyshevControl", 33), ("PWConstantControl", 150), ("PWLinearControl", 161)])
def test_smallest_and_largest_bases(ocs, oracle, kind, K):
    """One interval / two knots / a constant Chebyshev basis, and bases with more knots than RK4 steps
    (the knot window jumps by more than one per grid point)."""
    for per in (False, True):
        if kind == "ChebyshevControl" and K > 32 and per:
            continue          # run-time dense K > 32 needs a generated problem (compile-time K)
        out, ref = _eval_both(ocs, oracle, 60, 3.0, kind, K, 37, seed=K, per_traj_v=per)
        for key in ("J", "xT", "dJdv", "lam0"):
            assert relerr(out[key], ref[key]) < TOL, (key, kind, K, per)


def test_divergent_trajectories_are_flagged(ocs):
    """The reference has no NaN/Inf guard: a trajectory that blows up just propagates
    Inf/NaN into J.  Here it additionally sets its status word; the others are unaffected."""
    tspan = ocs.matlab_linspace(0.0, 10.0, 201)
    integ = ocs.CudaRK4Integrator(tspan)
    ctl = ocs.PWLinearControl(integ.t, 5, 1)
    integ.attach_control(ctl)
    x0 = np.full((1, 300), 1.0)
    x0[0, 17] = -50.0          # x' = x(m - x) - u runs away to -inf
    x0[0, 299] = -1e3
    out = integ.evaluate(ocs.TestOCProblem(TEST_PARAMS, [[0, 1]]), x0, np.full(5, .5))
    assert out["status"][17] == 1 and out["status"][299] == 1
    assert not np.isfinite(out["J"][17])
    ok = np.ones(300, bool)
    ok[[17, 299]] = False
    assert not out["status"][ok].any() and np.all(out["J"][ok] == out["J"][0])


def test_argument_validation(ocs):
    L = ocs._lib
    tspan = ocs.matlab_linspace(0.0, 1.0, 11)
    integ = ocs.CudaRK4Integrator(tspan)
    prob = ocs.TestOCProblem(TEST_PARAMS, [[0, 1]])
    ctl = ocs.PWLinearControl(integ.t, 4, 1)
    with pytest.raises(RuntimeError):
        integ.evaluate(prob, 1.0, np.zeros(4))                       # no control attached
    integ.attach_control(ctl)
    with pytest.raises(ocs.OcsError) as e:
        integ.evaluate(prob, np.ones((1, 5)), np.zeros((4, 3)))       # v batch != x0 batch
    assert e.value.code == L.ERR_ARG
    with pytest.raises(ValueError):
        integ.evaluate(prob, 1.0, np.zeros(7))                       # wrong coefficient count (host-side check)
    with pytest.raises(ValueError):
        integ.attach_control(ocs.PWLinearControl(np.linspace(0, 1, 11), 4, 1))   # basis on another grid
    bad = np.zeros((4, integ.t.size))
    bad[0, 0] = np.nan
    lib = L.load()
    assert lib.ocs_b200_set_control_basis(integ._h, 4, L.dp(L.f64(bad))) == L.ERR_ARG
    # per-trajectory parameters must match the batch
    prob_b = ocs.TestOCProblem(dict(c=np.ones(3), m=3.0, r=.05), [[0, 1]])
    integ.attach_control(ctl)
    with pytest.raises(ocs.OcsError):
        integ.evaluate(prob_b, np.ones((1, 5)), np.zeros(4))
    # fetch range checks
    integ.evaluate(prob, np.ones((1, 5)), np.zeros(4))
    with pytest.raises(ocs.OcsError):
        integ.fetch_states(3, 10)
    # the handle is still usable after errors (only CUDA errors are sticky)
    out = integ.evaluate(prob, np.ones((1, 5)), np.zeros(4))
    assert np.all(np.isfinite(out["J"]))
    # double destroy through the Python wrapper is harmless
    integ.close()
    integ.close()


@pytest.mark.parametrize("kind,K,N", [("PWLinearControl", 21, 100), ("PWLinearControl", 2, 7), ("PWLinearControl", 101, 50),
                                      ("PWConstantControl", 10, 100), ("PWConstantControl", 1, 5), ("PWConstantControl", 64, 33)])
def test_gradient_buffer_is_fully_written_without_a_zero_fill(ocs, oracle, kind, K, N):
    """The sparse-basis sweep stores every dJdv column exactly once (no zero-fill, no read-modify-write):
    a caller-owned device buffer poisoned with NaN must come back complete -- including knots that
    own very few (K > N) or all (K = 1) grid points."""
    import torch
    Bt = 333
    tspan = ocs.matlab_linspace(0.0, 3.0, N + 1)
    integ = ocs.CudaRK4Integrator(tspan)
    ctl = getattr(ocs, kind)(integ.t, K, 1)
    integ.attach_control(ctl)
    rng = np.random.default_rng(9)
    x0 = rng.uniform(.5, 2.5, (1, Bt))
    v = rng.random((K, Bt))
    dev = torch.device("cuda:0")
    x0d = torch.from_numpy(x0.copy()).to(dev)
    vd = torch.from_numpy(np.ascontiguousarray(v)).to(dev)                # SoA [(k*nc+c)][Bt]
    Jd = torch.full((Bt,), float("nan"), dtype=torch.float64, device=dev)
    gd = torch.full((K, Bt), float("nan"), dtype=torch.float64, device=dev)
    l0 = torch.full((2, Bt), float("nan"), dtype=torch.float64, device=dev)
    prob = ocs.TestOCProblem(TEST_PARAMS, [[0, 1]])
    for _ in range(2):                                                    # second pass: buffer holds old values
        integ.evaluate_dev(prob, x0d.data_ptr(), Bt, v_ptr=vd.data_ptr(), J_ptr=Jd.data_ptr(), dJdv_ptr=gd.data_ptr(),
                           lam0_ptr=l0.data_ptr())
        integ.synchronize()
        ref = oracle.C.evaluate(P, tspan, ctl.B, x0, v)
        assert torch.isfinite(gd).all()
        assert relerr(gd.cpu().numpy(), ref["dJdv"]) < TOL and relerr(Jd.cpu().numpy(), ref["J"]) < TOL
    integ.close()


def test_non_monotone_sparse_basis_falls_back_to_accumulation(ocs, oracle):
    """A user basis with at most two adjacent non-zeros per column whose knot index is NOT monotone along the
    grid (a knot is revisited): the kernel must add to the column instead of overwriting it."""
    N, K, Bt = 12, 4, 40
    tspan = ocs.matlab_linspace(0.0, 2.0, N + 1)
    integ = ocs.CudaRK4Integrator(tspan)
    G = 2 * N + 1
    B = np.zeros((K, G))
    ks = np.array([0, 1, 2, 1, 0, 2, 1, 0, 1, 2, 2, 1, 0, 0, 1, 2, 1, 0, 2, 2, 1, 0, 1, 2, 0])
    w = np.random.default_rng(3).uniform(.2, .8, G)
    B[ks, np.arange(G)] = w
    B[ks + 1, np.arange(G)] = 1 - w

    class UserControl(ocs.Control):
        nControls = 1

        def __init__(self, B):
            self.B = B
    ctl = UserControl(B)
    integ.attach_control(ctl)
    rng = np.random.default_rng(4)
    x0, v = rng.uniform(.5, 2.5, (1, Bt)), rng.random((K, Bt))
    out = integ.evaluate(ocs.TestOCProblem(TEST_PARAMS, [[0, 1]]), x0, v)
    ref = oracle.C.evaluate(P, tspan, B, x0, v)
    for key in ("J", "dJdv", "lam0"):
        assert relerr(out[key], ref[key]) < TOL, key
    integ.close()


@pytest.mark.parametrize("N", [1, 2, 3, 7, 31, 33, 65])
def test_grid_fetch_and_fb_sweep_for_odd_step_counts(ocs, oracle, N):
    """The single-trajectory kernels (fetch variants writing lam / dJdu grids, fb_sweep costate kernels) at odd
    and tiny step counts -- the unrolled gradient kernel once had a compiler problem exactly there."""
    T = 0.5 + 0.05 * N
    tspan = ocs.matlab_linspace(0.0, T, N + 1)
    integ = ocs.CudaRK4Integrator(tspan)
    ointeg = oracle.RK4Integrator(tspan)
    prob = ocs.TestOCProblem(TEST_PARAMS, [[0, 1]])
    oprob = oracle.TestOCProblem(**TEST_PARAMS)
    rng = np.random.default_rng(N)
    Bt, G = 37, 2 * N + 1
    x0 = rng.uniform(.5, 2.5, (1, Bt))
    u = rng.uniform(0, 1, (1, G, Bt))
    x, J = integ.compute_states(prob, x0, u, return_x=True)
    lam, dJdu = integ.compute_adjoints(prob, u, grids=True)
    for b in (0, 17, 36):
        ox, oJ = ointeg.compute_states(oprob, x0[:, b], u[:, :, b])
        olam, odu = ointeg.compute_adjoints(oprob, u[:, :, b])
        assert relerr(x[:, :, b], ox) < TOL and relerr(J[b], oJ) < TOL
        assert relerr(lam[:, :, b], olam) < TOL and relerr(dJdu[:, :, b], odu) < TOL
    cc = lambda t, xx, ll: oracle.control_char_test_problem(oprob, t, xx, ll)
    for mode in ("discrete", "continuous"):
        one = integ.fb_sweep(prob, x0, u, nSWEEPS=1, costate=mode)
        for b in (0, 36):
            ref = oracle.fb_sweep_rk4(oprob, ointeg, x0[:, b], u[:, :, b].copy(), cc, nSWEEPS=1, costate=mode)
            un = integ.get_control_u(b, 1)[:, :, 0]
            assert relerr(un, ref["u"]) < TOL, (mode, N)
            assert relerr(one["maxChange"][b], ref["maxChange"]) < 1e-9
    integ.close()


@pytest.mark.parametrize("uniform", [True, False])
@pytest.mark.parametrize("per_traj_v", [False, True])
@pytest.mark.parametrize("N", [1, 5, 64, 65, 200])
def test_parameter_sweep_time_exponentials_by_recurrence(ocs, oracle, uniform, per_traj_v, N):
    """Per-trajectory (c, m, r): the fast build advances exp(-r t) along the grid by one multiplication per grid
    point, re-anchored with exp() at every tile (rk4_kernels.cuh, ocs_ntexp); on a grid with unequal steps the factor
    is recomputed per step.  Tile boundaries (64 / 65 steps), a single step, shared and per-trajectory coefficients,
    against the C oracle (which calls exp() at every grid point)."""
    rng = np.random.default_rng(100 + N)
    T = 0.2 + 0.04 * N                   # steps of ~0.04-0.24: RK4 stays stable for every parameter set drawn below
    if uniform:
        tspan = ocs.matlab_linspace(0.0, T, N + 1)
    else:
        tspan = np.concatenate([[0.0], np.cumsum(rng.uniform(0.5, 1.5, N))]) * (T / N)
    Bt, K = 333, min(7, N + 1)
    integ = ocs.CudaRK4Integrator(tspan)
    ctl = ocs.PWLinearControl(integ.t, K, 1)
    integ.attach_control(ctl)
    c, m, r = rng.uniform(1.0, 2.0, Bt), rng.uniform(2.5, 3.5, Bt), rng.uniform(0.0, 0.5, Bt)
    prob = ocs.TestOCProblem(dict(c=c, m=m, r=r), [[0, 1]])
    x0 = rng.uniform(.5, 2.5, (1, Bt))
    v = rng.random((K, Bt)) if per_traj_v else rng.random(K)
    out = integ.evaluate(prob, x0, v)
    ref = oracle.C.evaluate(np.stack([c, m, r]), tspan, ctl.B, x0, v)
    assert not out["status"].any() and np.all(np.isfinite(ref["J"]))
    for key in ("J", "xT", "dJdv", "lam0"):
        assert relerr(out[key], ref[key]) < TOL, (key, N, uniform)
    # the grids (single-trajectory fetch kernels keep exp()) agree with the batched sweep's results
    x = integ.fetch_states(0, Bt)
    assert relerr(x[1, -1, :], out["J"]) < 1e-12
    integ.close()


def _unequal(rng, t0, t1, N):
    s = rng.uniform(0.5, 1.5, N)
    return t0 + np.concatenate([[0.0], np.cumsum(s)]) * ((t1 - t0) / s.sum())


@pytest.mark.parametrize("N", [1, 5, 33, 130])
def test_unequal_steps_take_the_per_step_kernels(ocs, oracle, N):
    """The fast build has two kernel sets: grids whose steps agree to rounding read ONE step size as a constant
    operand (OCS_UNIFORM_H), every other grid reads per-step sizes from the step records (kernels_builtin.cu).  A
    grid with steps between 0.5 and 1.5 of the mean must take the second set in every sweep -- grid controls,
    coefficients, fetch kernels, both fb_sweep costates -- and match the oracle like the uniform ones do."""
    rng = np.random.default_rng(300 + N)
    tspan = _unequal(rng, 0.0, 0.5 + 0.05 * N, N)
    integ = ocs.CudaRK4Integrator(tspan)
    ointeg = oracle.RK4Integrator(tspan)
    prob = ocs.TestOCProblem(TEST_PARAMS, [[0, 1]])
    oprob = oracle.TestOCProblem(**TEST_PARAMS)
    Bt, G = 37, 2 * N + 1
    x0 = rng.uniform(.5, 2.5, (1, Bt))
    u = rng.uniform(0, 1, (1, G, Bt))
    x, J = integ.compute_states(prob, x0, u, return_x=True)
    lam, dJdu = integ.compute_adjoints(prob, u, grids=True)
    for b in (0, 36):
        ox, oJ = ointeg.compute_states(oprob, x0[:, b], u[:, :, b])
        olam, odu = ointeg.compute_adjoints(oprob, u[:, :, b])
        assert relerr(x[:, :, b], ox) < TOL and relerr(J[b], oJ) < TOL
        assert relerr(lam[:, :, b], olam) < TOL and relerr(dJdu[:, :, b], odu) < TOL
    cc = lambda t, xx, ll: oracle.control_char_test_problem(oprob, t, xx, ll)
    for mode in ("discrete", "continuous"):
        integ.fb_sweep(prob, x0, u, nSWEEPS=1, costate=mode)
        ref = oracle.fb_sweep_rk4(oprob, ointeg, x0[:, 36], u[:, :, 36].copy(), cc, nSWEEPS=1, costate=mode)
        assert relerr(integ.get_control_u(36, 1)[:, :, 0], ref["u"]) < TOL, (mode, N)
    K = min(6, N + 1)
    ctl = ocs.PWLinearControl(integ.t, K, 1)
    integ.attach_control(ctl)
    for v in (rng.random(K), rng.random((K, Bt))):
        out = integ.evaluate(prob, x0, v)
        ref = oracle.C.evaluate(P, tspan, ctl.B, x0, v)
        for key in ("J", "xT", "dJdv", "lam0"):
            assert relerr(out[key], ref[key]) < TOL, (key, N)
    integ.close()


@pytest.mark.parametrize("which", ["main", "tail"])
def test_infinite_horizon_with_one_unequal_segment(ocs, oracle, which):
    """One segment of an RK4InfiniteIntegrator with unequal steps switches BOTH segments to the per-step kernels
    (the kernel set is chosen per handle)."""
    rng = np.random.default_rng(7)
    tspan = _unequal(rng, 0.0, 6.0, 150) if which == "main" else ocs.matlab_linspace(0.0, 6.0, 151)
    tspanE = _unequal(rng, 6.0, 14.0, 90) if which == "tail" else ocs.matlab_linspace(6.0, 14.0, 91)
    uStar = 0.723368780097983
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
    for b in (0, 39):
        _, J = ointeg.compute_states(oprob, x0[:, b], u)
        lam, dJdu = ointeg.compute_adjoints(oprob, u)
        assert relerr(out["J"][b], J) < TOL
        assert relerr(out["lam0"][:, b], lam[:, 0]) < TOL
        assert relerr(out["dJdv"][:, b], octl.compute_dJdv(dJdu)) < TOL
