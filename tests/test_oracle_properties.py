"""Self-checks of the CPU oracle: the properties the reference's own test pins
(tests/backprop_test.m: adjoint gradient == finite differences), RK4's order, an
mpmath 40-digit forward sweep bounding the oracle's rounding, agreement of the NumPy and
C restatements on batches, and the survey-time anchors."""
import numpy as np
import pytest

from util import relerr

P = (1.5, 3.0, 0.05)


def test_anchors_appendix_c(oracle):
    prob = oracle.TestOCProblem(*P)
    integ = oracle.RK4Integrator(oracle.matlab_linspace(0, 10, 501))
    ctl = oracle.PWLinearControl(integ.t, 21, 1)
    v = .5 + .5 * np.random.default_rng(0).random(21)
    J, dJdv = oracle.nlp_objective(prob, integ, ctl, [1.0], v)
    assert abs(J - 60.023749514116858) < 1e-12
    assert abs(integ.xK[0, -1, 0] - 2.7755466487484992) < 1e-13
    lam = integ.compute_adjoints(prob, ctl.compute_u(v), want_dJdu=False)
    assert abs(lam[0, 0] - 4.8294129510278241) < 1e-12 and np.all(lam[-1] == 1.0)


def test_adjoint_is_the_exact_discrete_gradient(oracle):
    """Central differences (epsilon = 1e-4) agree with the adjoint gradient to ~5e-10; forward
    differences show the 4e-5 truncation the reference script prints."""
    prob = oracle.TestOCProblem(*P)
    integ = oracle.RK4Integrator(oracle.matlab_linspace(0, 10, 501))
    ctl = oracle.PWLinearControl(integ.t, 21, 1)
    v = .5 + .5 * np.random.default_rng(0).random(21)
    J, dJdv = oracle.nlp_objective(prob, integ, ctl, [1.0], v)
    eps = 1e-4
    cd = np.empty(21)
    for i in range(21):
        e = np.zeros(21)
        e[i] = eps
        cd[i] = (oracle.nlp_objective(prob, integ, ctl, [1.0], v + e)[0] -
                 oracle.nlp_objective(prob, integ, ctl, [1.0], v - e)[0]) / (2 * eps)
    assert np.max(np.abs(cd - dJdv)) < 5e-9
    # free initial state: dJ/dx0 = lam(FreeInitStates, 1)   (single_shooting.m:147-148)
    Jp, g = oracle.nlp_objective(prob, integ, ctl, [1.0], np.concatenate([v, [1.0]]), FreeInitStates=(0,))
    d = (oracle.nlp_objective(prob, integ, ctl, [1.0 + eps], v)[0] - oracle.nlp_objective(prob, integ, ctl, [1.0 - eps], v)[0]) / (2 * eps)
    assert abs(g[-1] - d) < 5e-8 and Jp == J


def test_rk4_order(oracle):
    prob = oracle.TestOCProblem(*P)

    def Jn(N):
        integ = oracle.RK4Integrator(oracle.matlab_linspace(0, 10, N + 1))
        return integ.compute_states(prob, [1.0], np.full((1, 2 * N + 1), 0.5))[1]
    ref = Jn(3200)
    errs = [abs(Jn(N) - ref) for N in (50, 100, 200, 400)]
    ratios = [errs[i] / errs[i + 1] for i in range(3)]
    assert all(14.5 < r < 17.0 for r in ratios), ratios


def test_mpmath_forward_sweep_bounds_rounding(oracle):
    mp = pytest.importorskip("mpmath")
    mp.mp.dps = 40
    N = 200
    tspan = oracle.matlab_linspace(0, 10, N + 1)
    integ = oracle.RK4Integrator(tspan)
    ctl = oracle.PWLinearControl(integ.t, 11, 1)
    v = np.random.default_rng(3).random(11)
    u = ctl.compute_u(v)[0]
    prob = oracle.TestOCProblem(*P)
    _, J = integ.compute_states(prob, [1.0], u[None])
    c, m, r = (mp.mpf(x) for x in P)

    def F(t, x, uu):
        return x * (m - x) - uu, mp.e ** (-r * t) * (x * x + c * uu * uu)
    x, Jm = mp.mpf(1), mp.mpf(0)
    t, h = [mp.mpf(float(z)) for z in integ.t], [mp.mpf(float(z)) for z in integ.h]
    um = [mp.mpf(float(z)) for z in u]
    for i in range(N):
        f1 = F(t[2 * i], x, um[2 * i])
        f2 = F(t[2 * i + 1], x + h[i] / 2 * f1[0], um[2 * i + 1])
        f3 = F(t[2 * i + 1], x + h[i] / 2 * f2[0], um[2 * i + 1])
        f4 = F(t[2 * i + 2], x + h[i] * f3[0], um[2 * i + 2])
        x, Jm = (x + h[i] / 6 * (f1[0] + 2 * f2[0] + 2 * f3[0] + f4[0]),
                 Jm + h[i] / 6 * (f1[1] + 2 * f2[1] + 2 * f3[1] + f4[1]))
    assert abs(float(Jm) - J) / abs(J) < 1e-14
    assert abs(float(x) - integ.xK[0, -1, 0]) < 1e-14


def test_c_and_numpy_restatements_agree_on_batches(oracle):
    tspan = oracle.matlab_linspace(0, 10, 201)
    integ = oracle.RK4Integrator(tspan)
    prob = oracle.TestOCProblem(*P)
    rng = np.random.default_rng(5)
    x0 = rng.uniform(.5, 2.5, (1, 64))
    for cls, K in (("PWLinearControl", 21), ("PWConstantControl", 10), ("ChebyshevControl", 9)):
        ctl = getattr(oracle, cls)(integ.t, K, 1)
        v = rng.random((K, 64)) * (0.1 if cls == "ChebyshevControl" else 1.0)
        u = ctl.compute_u(v)
        x, J = integ.compute_states(prob, x0, u)
        lam, dJdu = integ.compute_adjoints(prob, u)
        out = oracle.C.evaluate(P, tspan, ctl.B, x0, v, want_grids=True)
        assert relerr(out["J"], J) < 1e-14
        assert relerr(out["lam"], lam.transpose(0, 1, 2)) < 1e-13
        assert relerr(out["dJdu"], dJdu) < 1e-13
        assert relerr(out["dJdv"], ctl.compute_dJdv(dJdu)) < 1e-13


def test_equilibrium_and_fb_sweep_composition(oracle):
    prob = oracle.TestOCProblem(*P)
    xs, ls, us = oracle.solve_equilibrium(prob, [2.7], [2.2], [.7], 0.05)
    assert abs(xs[0] - 2.73556918863414) < 1e-10 and abs(us[0] - 0.723368780097983) < 1e-10
    integ = oracle.RK4Integrator(oracle.matlab_linspace(0, 4, 201))
    cc = lambda t, x, lam: oracle.control_char_test_problem(prob, t, x, lam)
    out = oracle.fb_sweep_rk4(prob, integ, [1.0], np.full((1, 401), .5), cc, nSWEEPS=60)
    assert out["converged"] and out["maxChange"] <= 1.0
    # at the fixed point the control satisfies its characterisation
    xg, lg = oracle._nodes_to_grid(out["x"][:-1]), oracle._nodes_to_grid(out["lam"][:-1])
    assert np.max(np.abs(cc(integ.t, xg, lg) - out["u"])) < 1e-6


def test_continuous_costate_matches_an_adaptive_solve(oracle):
    """costate_rk4_continuous (the v1 adjointRHS integrated on the RK4 grid) against scipy's adaptive
    integration of the same ODE through a dense interpolant of x: the fixed-grid solve is O(h^2)
    (midpoint state = mean of the nodes), so the gap shrinks 4x per grid halving."""
    from scipy.integrate import solve_ivp
    from scipy.interpolate import CubicSpline
    prob = oracle.TestOCProblem(*P)
    errs = []
    for N in (100, 200, 400):
        integ = oracle.RK4Integrator(oracle.matlab_linspace(0, 4, N + 1))
        u = 0.3 + 0.2 * np.sin(integ.t)[None, :]
        x, _ = integ.compute_states(prob, [1.3], u)
        lam = oracle.costate_rk4_continuous(prob, integ, x[:-1], u)
        xs = CubicSpline(integ.t[0::2], x[0])
        uf = lambda t: 0.3 + 0.2 * np.sin(t)
        rhs = lambda t, l: oracle.adjoint_rhs(prob, t, np.array([xs(t)]), l, np.array([uf(t)]))
        sol = solve_ivp(rhs, [4.0, 0.0], [0.0], rtol=1e-11, atol=1e-13, dense_output=True)
        errs.append(np.max(np.abs(sol.sol(integ.t[0::2])[0] - lam[0])))
    assert errs[0] < 1e-3 and errs[1] < errs[0] / 3.5 and errs[2] < errs[1] / 3.5


def test_pchip_restatement_against_scipy(oracle):
    """oracle.pchip_eval (MATLAB pchip.m's algorithm, used for soln.x / soln.lam) against SciPy's
    PchipInterpolator -- an independent implementation of the same Fritsch-Carlson interpolant --
    on monotone, oscillating, flat and non-uniformly spaced data, including the 2- and 3-point grids."""
    from scipy.interpolate import PchipInterpolator
    rng = np.random.default_rng(5)
    for n in (2, 3, 4, 7, 50, 501):
        for kind in ("uniform", "random"):
            x = np.linspace(0, 10, n) if kind == "uniform" else np.cumsum(rng.uniform(.1, 1.0, n))
            for y in (rng.standard_normal(n), np.cumsum(rng.uniform(0, 1, n)), np.sin(x), np.where(x > x.mean(), 1.0, 0.0)):
                tq = np.concatenate([x, rng.uniform(x[0], x[-1], 200)])
                ref = PchipInterpolator(x, y)(tq) if n > 2 else np.interp(tq, x, y)
                got = oracle.pchip_eval(x, y, tq)
                assert np.max(np.abs(got - ref)) < 1e-12 * max(1.0, np.max(np.abs(y)))
                assert np.array_equal(got[:n - 1], y[:-1])     # nodes (s = 0) are reproduced exactly
    vi = oracle.vector_interpolant(np.array([0., 1., 2.]), np.array([[1., 3., 2.], [0., 0., 4.]]), "previous")
    assert np.array_equal(vi([0.0, 0.99, 1.0, 5.0]), [[1, 1, 3, 2], [0, 0, 0, 4]])
    vl = oracle.vector_interpolant(np.array([0., 1., 2.]), np.array([[1., 3., 2.]]), "linear")
    assert np.allclose(vl([0.5, 1.5, 2.0, 3.0]), [[2.0, 2.5, 2.0, 2.0]])
