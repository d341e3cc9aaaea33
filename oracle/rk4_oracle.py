"""CPU oracle for the RK4 trajectory hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.  The product path (optimal-control-solvers_b200) never
does; it fails loudly when its CUDA library is missing.

What this is: a NumPy FP64 restatement of the reference's MATLAB arithmetic, one
function per reference function, same operation order, each citing the file:line
of /root/reference it follows.  Every routine accepts a trailing batch axis
(``y`` of shape (m,) or (m, Bt)) because the reference's own F / dFdx_times_vec /
dFdu_times_vec are written column-vectorised (tests/TestOCProblem.m:23,30).

PARITY PINNING STATUS: pinned against the reference's own source.  The reference
ships no numeric golden vectors and neither MATLAB nor Octave exists in this image
(SURVEY.md section 8c), so the reference's UNMODIFIED .m files (RK4Integrator.m,
RK4InfiniteIntegrator.m, the three Control classes, TestOCProblem.m and the script
tests/backprop_test.m) are executed by the MATLAB-subset interpreter in oracle/mlab;
their outputs are committed as tests/golden/*.npz (generator:
tests/golden/generate_golden.py) and tests/test_oracle_golden.py holds this module and
the C restatement to them: grids, step sizes, basis matrices and control-grid indices
bit-identical, sweeps to 1e-13.  What that run does NOT pin are MATLAB's built-ins,
which oracle/mlab/builtins.py restates (linspace rounding, mtimes summation order, exp,
griddedInterpolant's formula): a genuine MATLAB run could differ from these vectors by a
few ulp in exp()/BLAS products, far inside the 1e-10 bar, and -- only if MathWorks'
linspace rounded differently from the documented formula -- in PWConstant indices.
Further self-checks: tests/test_oracle_properties.py (adjoint == finite differences as
tests/backprop_test.m prints, RK4 order, an mpmath 40-digit forward sweep).
"""
from __future__ import annotations

import numpy as np

__all__ = [
    "matlab_linspace", "RK4Integrator", "RK4InfiniteIntegrator",
    "PWConstantControl", "PWLinearControl", "ChebyshevControl",
    "TestOCProblem", "GenericOCProblem", "nlp_objective", "fb_sweep_rk4",
    "control_char_test_problem", "compute_states_and_objective_legacy",
    "equilibrium_residual", "solve_equilibrium",
]


# --------------------------------------------------------------------------- grids
def matlab_linspace(d1: float, d2: float, n: int) -> np.ndarray:
    """MATLAB ``linspace(d1,d2,n)``: ``d1 + (0:n1).*(d2-d1)./n1`` with both end
    points forced.  NOT numpy.linspace: that one multiplies a
    pre-divided step and flips PWConstant interval indices."""
    n = int(n)
    if n < 1:
        return np.zeros(0)
    if n == 1:
        return np.array([float(d2)])
    n1 = n - 1
    i = np.arange(0, n1 + 1, dtype=np.float64)
    y = d1 + (i * (d2 - d1)) / n1
    y[0] = d1
    y[-1] = d2
    return y


# --------------------------------------------------------------------------- problems
class TestOCProblem:
    """tests/TestOCProblem.m:13-38.  One state + running objective.
    ``F = [x(m-x)-u ; exp(-r t)(x^2 + c u^2)]``."""
    __test__ = False  # not a pytest class

    nStates = 1
    nControls = 1

    def __init__(self, c=1.5, m=3.0, r=0.05, ControlBounds=((0.0, 1.0),)):
        self.c, self.m, self.r = float(c), float(m), float(r)
        self.ControlBounds = np.asarray(ControlBounds, dtype=np.float64).reshape(-1, 2)

    def F(self, t, y, u):
        # TestOCProblem.m:22-26
        x = y[0]
        u = u[0]
        return np.stack([x * (self.m - x) - u,
                         np.exp(-self.r * t) * (x ** 2 + self.c * u ** 2)])

    def dFdx_times_vec(self, t, y, u, v):
        # TestOCProblem.m:29-33
        x = y[0]
        top = (self.m - 2 * x) * v[0] + 2 * np.exp(-self.r * t) * x * v[1]
        return np.stack([top, np.zeros_like(top)])

    def dFdu_times_vec(self, t, y, u, v):
        # TestOCProblem.m:36-38
        return (-v[0] + 2 * self.c * np.exp(-self.r * t) * u[0] * v[1])[None]


class GenericOCProblem:
    """An OCProblem (OCProblem/OCProblem.m:3-21) assembled from three NumPy
    callables -- used for the codegen problems (lambdified from the same SymPy
    expressions the CUDA backend prints)."""

    def __init__(self, nStates, nControls, F, dFdx_times_vec, dFdu_times_vec,
                 ControlBounds, control_char=None):
        self.nStates, self.nControls = nStates, nControls
        self._F, self._X, self._U = F, dFdx_times_vec, dFdu_times_vec
        self.ControlBounds = np.asarray(ControlBounds, dtype=np.float64).reshape(-1, 2)
        self._cc = control_char

    def F(self, t, y, u):
        return self._F(t, y, u)

    def dFdx_times_vec(self, t, y, u, v):
        return self._X(t, y, u, v)

    def dFdu_times_vec(self, t, y, u, v):
        return self._U(t, y, u, v)

    def control_char(self, t, x, lam):
        return self._cc(t, x, lam)


# --------------------------------------------------------------------------- integrators
def _col(u, j):
    """u(:,j) for u of shape (nc,G) or (nc,G,Bt)."""
    return u[:, j]


class RK4Integrator:
    """Integrator/RK4Integrator.m.  0-based here: grid point 2i is the left node of
    step i, 2i+1 its midpoint, 2i+2 its right node."""

    def __init__(self, tspan):
        # RK4Integrator.m:16-25
        tspan = np.asarray(tspan, dtype=np.float64)
        self.h = np.diff(tspan)
        self.nSTEPS = len(self.h)
        t = np.zeros(2 * self.nSTEPS + 1)
        t[0::2] = tspan
        t[1:-1:2] = (tspan[:-1] + tspan[1:]) / 2
        self.t = t
        self.xK = None

    def compute_states(self, prob, x0, u):
        # RK4Integrator.m:28-56
        x0 = np.asarray(x0, dtype=np.float64)
        if x0.ndim == 0:
            x0 = x0[None]
        nSTATES = x0.shape[0] + 1
        N, t, h = self.nSTEPS, self.t, self.h
        xK = np.full((nSTATES, N + 1, 4) + x0.shape[1:], np.nan)
        xK[:, 0, 0] = np.concatenate([x0, np.zeros((1,) + x0.shape[1:])])
        for i in range(N):
            F1 = prob.F(t[2 * i], xK[:, i, 0], _col(u, 2 * i))
            xK[:, i, 1] = xK[:, i, 0] + h[i] / 2 * F1
            F2 = prob.F(t[2 * i + 1], xK[:, i, 1], _col(u, 2 * i + 1))
            xK[:, i, 2] = xK[:, i, 0] + h[i] / 2 * F2
            F3 = prob.F(t[2 * i + 1], xK[:, i, 2], _col(u, 2 * i + 1))
            xK[:, i, 3] = xK[:, i, 0] + h[i] * F3
            F4 = prob.F(t[2 * i + 2], xK[:, i, 3], _col(u, 2 * i + 2))
            xK[:, i + 1, 0] = xK[:, i, 0] + h[i] / 6 * (F1 + 2 * F2 + 2 * F3 + F4)
        self.xK = xK
        x = xK[:, :, 0]
        J = x[-1, -1]
        return x, J

    def compute_adjoints(self, prob, u, lamT=None, want_dJdu=True):
        # RK4Integrator.m:59-94
        xK = self.xK
        nSTATES = xK.shape[0]
        N, t, h = self.nSTEPS, self.t, self.h
        bshape = xK.shape[3:]
        if lamT is None:
            lamT = np.zeros((nSTATES,) + bshape)
            lamT[-1] = 1
        lamT = np.asarray(lamT, dtype=np.float64)
        lam = np.full((nSTATES, N + 1) + bshape, np.nan)
        lam[:, -1] = lamT if lamT.shape == lam[:, -1].shape else \
            np.broadcast_to(lamT.reshape((nSTATES,) + (1,) * len(bshape)), lam[:, -1].shape)
        dJdk = np.full((nSTATES, N, 4) + bshape, np.nan)
        for i in range(N - 1, -1, -1):
            dJdk[:, i, 3] = h[i] / 6 * lam[:, i + 1]
            dJdx3 = prob.dFdx_times_vec(t[2 * i + 2], xK[:, i, 3], _col(u, 2 * i + 2), dJdk[:, i, 3])
            dJdk[:, i, 2] = h[i] / 3 * lam[:, i + 1] + h[i] * dJdx3
            dJdx2 = prob.dFdx_times_vec(t[2 * i + 1], xK[:, i, 2], _col(u, 2 * i + 1), dJdk[:, i, 2])
            dJdk[:, i, 1] = h[i] / 3 * lam[:, i + 1] + h[i] / 2 * dJdx2
            dJdx1 = prob.dFdx_times_vec(t[2 * i + 1], xK[:, i, 1], _col(u, 2 * i + 1), dJdk[:, i, 1])
            dJdk[:, i, 0] = h[i] / 6 * lam[:, i + 1] + h[i] / 2 * dJdx1
            lam[:, i] = lam[:, i + 1] + dJdx1 + dJdx2 + dJdx3 + \
                prob.dFdx_times_vec(t[2 * i], xK[:, i, 0], _col(u, 2 * i), dJdk[:, i, 0])
        self.dJdk = dJdk
        if not want_dJdu:
            return lam
        return lam, self.compute_dJdu(prob, u, dJdk)

    def compute_dJdu(self, prob, u, dJdk):
        # RK4Integrator.m:97-121 (grid-point by grid-point; the reference does the
        # same sums with vectorised calls over t -- identical per element)
        xK, t, N = self.xK, self.t, self.nSTEPS
        dJdu = np.zeros(u.shape[:2] + xK.shape[3:])
        U = prob.dFdu_times_vec
        # left end :101-102
        dJdu[:, 0] = U(t[0], xK[:, 0, 0], _col(u, 0), dJdk[:, 0, 0])
        for i in range(N):
            # midpoints :105-109
            j = 2 * i + 1
            dJdu[:, j] = U(t[j], xK[:, i, 1], _col(u, j), dJdk[:, i, 1]) + \
                U(t[j], xK[:, i, 2], _col(u, j), dJdk[:, i, 2])
        for i in range(1, N):
            # interior nodes :112-116
            j = 2 * i
            dJdu[:, j] = U(t[j], xK[:, i, 0], _col(u, j), dJdk[:, i, 0]) + \
                U(t[j], xK[:, i - 1, 3], _col(u, j), dJdk[:, i - 1, 3])
        # right end :119-120
        dJdu[:, -1] = U(t[-1], xK[:, N - 1, 3], _col(u, 2 * N), dJdk[:, N - 1, 3])
        return dJdu


class RK4InfiniteIntegrator:
    """Integrator/RK4InfiniteIntegrator.m:12-30."""

    def __init__(self, tspan, tspanExtra, uStar):
        self.integrator1 = RK4Integrator(tspan)
        self.integrator2 = RK4Integrator(tspanExtra)
        uStar = np.atleast_1d(np.asarray(uStar, dtype=np.float64)).reshape(-1, 1)
        self.uStar = uStar * np.ones((1, self.integrator2.t.size))  # :15
        self.t = self.integrator1.t

    def compute_states(self, prob, x0, u):
        x, J1 = self.integrator1.compute_states(prob, x0, u)          # :21
        _, J2 = self.integrator2.compute_states(prob, x[:-1, -1], self.uStar)  # :22
        self.J1, self.J2 = J1, J2
        return x, J1 + J2                                              # :23

    def compute_adjoints(self, prob, u, want_dJdu=True):
        lam2 = self.integrator2.compute_adjoints(prob, self.uStar, want_dJdu=False)   # :28
        self.lam2 = lam2
        return self.integrator1.compute_adjoints(prob, u, lam2[:, 0], want_dJdu=want_dJdu)  # :29


# --------------------------------------------------------------------------- controls
class _BasisControl:
    def compute_u(self, v):
        """``v = reshape(v,nControls,[]); u = v*B`` (PWLinearControl.m:59-62 and
        twins).  v may carry a trailing batch axis: (nc*K,) or (nc*K, Bt)."""
        v = np.asarray(v, dtype=np.float64)
        if v.ndim == 1:
            return v.reshape(self.nControls, -1, order="F") @ self.B
        Bt = v.shape[1]
        vv = v.reshape(self.nControls, -1, Bt, order="F")      # (nc, K, Bt)
        return np.einsum("ckb,kg->cgb", vv, self.B)

    def compute_dJdv(self, dJdu):
        """``dJdv = reshape(dJdu*B',[],1)`` (PWLinearControl.m:53-56 and twins)."""
        dJdu = np.asarray(dJdu, dtype=np.float64)
        if dJdu.ndim == 2:
            return (dJdu @ self.B.T).reshape(-1, order="F")
        d = np.einsum("cgb,kg->ckb", dJdu, self.B)
        return d.reshape(-1, d.shape[2], order="F")


class PWConstantControl(_BasisControl):
    """Control/PWConstantControl.m:12-55."""

    def __init__(self, t, nControlIntervals, nControls):
        self.nControlIntervals, self.nControls = int(nControlIntervals), int(nControls)
        t = np.asarray(t, dtype=np.float64)
        s = matlab_linspace(t[0], t[-1], self.nControlIntervals + 1)   # :15
        self.intervalStarts = s[:-1]                                   # :16
        self.B = self.compute_basis_matrix(t)

    def compute_basis_matrix(self, t):
        # :30-38 literal comparisons
        K, s = self.nControlIntervals, self.intervalStarts
        B = np.zeros((K, t.size))
        for i in range(K - 1):
            B[i] = ((t >= s[i]) & (t < s[i + 1])).astype(np.float64)
        B[-1] = (t >= s[-1]).astype(np.float64)
        return B

    def interval_index(self):
        """0-based interval index of every grid point (the 'control-grid index'
        that must match bit-exactly)."""
        assert np.all(self.B.sum(axis=0) == 1)
        return np.argmax(self.B, axis=0).astype(np.int32)

    def compute_nlp_bounds(self, controlBounds):
        cb = np.asarray(controlBounds, dtype=np.float64).reshape(-1, 2)
        Lb = (cb[:, :1] * np.ones((1, self.nControlIntervals))).reshape(-1, order="F")
        Ub = (cb[:, 1:] * np.ones((1, self.nControlIntervals))).reshape(-1, order="F")
        return Lb, Ub

    def compute_initial_v(self, u0):
        return np.tile(np.atleast_1d(u0).astype(np.float64), self.nControlIntervals)  # :53-55


def _tent(xp, fp, t):
    """griddedInterpolant(xp, fp, 'linear', 'nearest') evaluated at t
    (PWLinearControl.m:35-49): piecewise linear inside, nearest sample outside."""
    t = np.asarray(t, dtype=np.float64)
    out = np.empty_like(t)
    lo, hi = t < xp[0], t > xp[-1]
    out[lo], out[hi] = fp[0], fp[-1]
    mid = ~(lo | hi)
    tm = t[mid]
    k = np.clip(np.searchsorted(xp, tm, side="right") - 1, 0, len(xp) - 2)
    s = (tm - xp[k]) / (xp[k + 1] - xp[k])
    out[mid] = fp[k] + s * (fp[k + 1] - fp[k])
    return out


class PWLinearControl(_BasisControl):
    """Control/PWLinearControl.m:13-71."""

    def __init__(self, t, nControlPts, nControls):
        self.nControlPts, self.nControls = int(nControlPts), int(nControls)
        t = np.asarray(t, dtype=np.float64)
        self.controlPts = matlab_linspace(t[0], t[-1], self.nControlPts)   # :16
        self.B = self.compute_basis_matrix(t)

    def compute_basis_matrix(self, t):
        K, p = self.nControlPts, self.controlPts
        B = np.zeros((K, t.size))
        B[0] = _tent(p[0:2], np.array([1.0, 0.0]), t)                  # :35-37
        for i in range(1, K - 1):
            B[i] = _tent(p[i - 1:i + 2], np.array([0.0, 1.0, 0.0]), t)  # :40-44
        B[-1] = _tent(p[-2:], np.array([0.0, 1.0]), t)                 # :47-49
        return B

    def compute_nlp_bounds(self, controlBounds):
        cb = np.asarray(controlBounds, dtype=np.float64).reshape(-1, 2)
        Lb = (cb[:, :1] * np.ones((1, self.nControlPts))).reshape(-1, order="F")
        Ub = (cb[:, 1:] * np.ones((1, self.nControlPts))).reshape(-1, order="F")
        return Lb, Ub

    def compute_initial_v(self, u0):
        u0 = np.atleast_1d(np.asarray(u0, dtype=np.float64))
        if u0.size == 1:
            return np.tile(u0, self.nControlPts)          # :66-67
        if u0.size == self.nControlPts:
            return u0.reshape(-1, order="F")              # :68-69
        raise ValueError("compute_initial_v: unsupported u0 length")


class ChebyshevControl(_BasisControl):
    """Control/ChebyshevControl.m:13-48 (forward 3-term recurrence, not cos/acos)."""

    def __init__(self, t, nControlBasis, nControls):
        self.nControlBasis, self.nControls = int(nControlBasis), int(nControls)
        t = np.asarray(t, dtype=np.float64)
        self.controlPts = matlab_linspace(t[0], t[-1], self.nControlBasis)   # :16
        self.t0, self.t1 = t[0], t[-1]
        self.B = self.compute_basis_matrix(t)

    def compute_basis_matrix(self, t):
        K = self.nControlBasis
        B = np.zeros((K, t.size))
        tTrans = 2 * (t - t[0]) / (t[-1] - t[0]) - 1     # :23
        B[0] = 1
        if K > 1:
            B[1] = tTrans
        for i in range(2, K):
            B[i] = 2 * tTrans * B[i - 1] - B[i - 2]      # :29
        return B

    def compute_initial_v(self, u0):
        u0 = np.atleast_1d(np.asarray(u0, dtype=np.float64))
        return np.concatenate([u0, np.zeros((self.nControlBasis - 1) * self.nControls)])  # :47


# --------------------------------------------------------------------------- callers
def nlp_objective(prob, integrator, control, x0, v, FreeInitStates=()):
    """functions/single_shooting.m:137-150 -- the per-evaluation call order.
    ``FreeInitStates`` are 0-based row indices here."""
    nfree = len(FreeInitStates)
    x0 = np.array(x0, dtype=np.float64, copy=True)
    if x0.ndim == 0:
        x0 = x0[None]
    if nfree == 0:
        u = control.compute_u(v)
        _, J = integrator.compute_states(prob, x0, u)
        _, dJdu = integrator.compute_adjoints(prob, u)
        return J, control.compute_dJdv(dJdu)
    u = control.compute_u(v[:-nfree])
    x0[list(FreeInitStates)] = v[-nfree:]
    _, J = integrator.compute_states(prob, x0, u)
    lam, dJdu = integrator.compute_adjoints(prob, u)
    return J, np.concatenate([control.compute_dJdv(dJdu), lam[list(FreeInitStates), 0]])


def control_char_test_problem(prob, t, x, lam):
    """ControlChar for tests/TestOCProblem.m: the u solving dFdu_times_vec(t,y,u,[lam;1])=0
    (make_from_symbolic.m:17-23), clamped to ControlBounds (:111):
    ``u* = clamp(lam*(exp(r t)/(2c)), umin, umax)``; the batch-uniform factor
    exp(r t)/(2c) is grouped, the form the device code tabulates."""
    u = lam * (np.exp(prob.r * t) / (2 * prob.c))
    return np.minimum(prob.ControlBounds[0, 1], np.maximum(prob.ControlBounds[0, 0], u))


def adjoint_rhs(prob, t, x, lam, u):
    """prob.adjointRHS of the v1 struct API: -gradient(H, x) with H = objective + lam*stateRHS
    (functions/make_from_symbolic.m:11-14).  In terms of the v2 callbacks this is
    -dFdx_times_vec(t, [x; .], u, [lam; 1]) restricted to the state rows."""
    y = np.concatenate([x, np.zeros((1,) + x.shape[1:])])
    v = np.concatenate([lam, np.ones((1,) + lam.shape[1:])])
    return -prob.dFdx_times_vec(t, y, u, v)[:-1]


def costate_rk4_continuous(prob, integrator, x, u):
    """The backward costate solve of functions/compute_x_lam.m:11-12 / compute_x_lam_J.m:18-19
    (lam' = adjointRHS(t, x(t), lam, u(t)), lam(T) = 0) with classical RK4 on the integrator's
    fixed grid in place of odevr7.  x: (n, N+1) node states, u: (nc, G).  Midpoint state = mean
    of the neighbouring nodes (the textbook forward-backward sweep).  Returns lam (n, N+1)."""
    t, h = integrator.t, integrator.h
    n, N = x.shape[0], h.size
    lam = np.zeros((n, N + 1) + x.shape[2:])
    for i in range(N - 1, -1, -1):
        lR, xR, xL = lam[:, i + 1], x[:, i + 1], x[:, i]
        xm = (xL + xR) / 2
        hh, h2, h6 = h[i], h[i] / 2, h[i] / 6
        k1 = adjoint_rhs(prob, t[2 * i + 2], xR, lR, _col(u, 2 * i + 2))
        k2 = adjoint_rhs(prob, t[2 * i + 1], xm, lR - h2 * k1, _col(u, 2 * i + 1))
        k3 = adjoint_rhs(prob, t[2 * i + 1], xm, lR - h2 * k2, _col(u, 2 * i + 1))
        k4 = adjoint_rhs(prob, t[2 * i], xL, lR - hh * k3, _col(u, 2 * i))
        lam[:, i] = lR - h6 * (k1 + 2.0 * k2 + 2.0 * k3 + k4)
    return lam


def fb_sweep_rk4(prob, integrator, x0, u0_grid, control_char, uRelTol=1e-7, uAbsTol=1e-7,
                 nSWEEPS=50, relax=1.0, costate="discrete"):
    """The fb_sweep iteration (functions/fb_sweep.m:79-115) composed on the RK4 grid
    (SURVEY.md App. A.6 -- a NEW composition; the reference's own fb_sweep integrates
    with odevr7).  Control lives on the integrator's G grid points.

    per sweep (:94-97)   x = forward(u); lam = discrete adjoint(u)  (costate="continuous":
                         lam from costate_rk4_continuous, the v1 adjointRHS);
                         uNew_j = ControlChar(t_j, x_j, lam_j)
    midpoint x/lam are the averages of the two neighbouring nodes (documented
    deviation: the reference interpolates ODE output with pchip).
    convergence (:107-114) max_j |uNew-u| / (uRelTol |u| + uAbsTol) <= 1.
    Returns dict(u, x, lam, J, sweeps, converged, maxChange)."""
    t = integrator.t
    u = np.array(u0_grid, dtype=np.float64, copy=True)
    out = dict(converged=False, sweeps=0)
    for sweep in range(1, nSWEEPS + 1):
        x, J = integrator.compute_states(prob, x0, u)
        if costate == "continuous":     # v1 API: adjointRHS integrated backwards (compute_x_lam.m:11-12)
            lam = np.concatenate([costate_rk4_continuous(prob, integrator, x[:-1], u), np.ones((1, x.shape[1]))])
        else:
            lam = integrator.compute_adjoints(prob, u, want_dJdu=False)
        xg = _nodes_to_grid(x[:-1])
        lg = _nodes_to_grid(lam[:-1])
        uNew = control_char(t, xg, lg)
        weighted = np.abs(uNew - u) / (uRelTol * np.abs(u) + uAbsTol)
        maxChange = weighted.max()
        out.update(sweeps=sweep, maxChange=maxChange)
        if maxChange <= 1:
            out.update(converged=True, u=u, x=x, lam=lam, J=J)
            return out
        u = uNew if relax == 1.0 else (1 - relax) * u + relax * uNew
    out.update(u=u, x=x, lam=lam, J=J)
    return out


def _nodes_to_grid(a):
    """(n, N+1[,Bt]) node values -> (n, 2N+1[,Bt]) grid values, midpoints = mean of
    neighbours ((a_i + a_{i+1})/2, the same formula the grid uses for t)."""
    n, N1 = a.shape[:2]
    g = np.empty((n, 2 * N1 - 1) + a.shape[2:])
    g[:, 0::2] = a
    g[:, 1::2] = (a[:, :-1] + a[:, 1:]) / 2
    return g


def pchip_slopes(x, y):
    """Derivative values of MATLAB's shape-preserving cubic (pchip.m, local function pchipslopes;
    Fritsch & Carlson 1980): interior points where the neighbouring secants agree in sign get their
    weighted harmonic mean, all others zero; the ends use the non-centred three-point formula with
    the shape-preserving clamps.  `griddedInterpolant(x, v, 'pchip')` (functions/vectorInterpolant.m:4,9)
    documents itself as this interpolant."""
    x, y = np.asarray(x, dtype=np.float64), np.asarray(y, dtype=np.float64)
    n = x.size
    h = np.diff(x)
    delta = np.diff(y) / h
    d = np.zeros(n)
    if n == 2:
        d[:] = delta[0]
        return d
    for k in range(n - 2):                       # interior point k+1, secants delta[k], delta[k+1]
        if np.sign(delta[k]) * np.sign(delta[k + 1]) > 0:
            hs = h[k] + h[k + 1]
            w1, w2 = (h[k] + hs) / (3 * hs), (hs + h[k + 1]) / (3 * hs)
            dmax, dmin = max(abs(delta[k]), abs(delta[k + 1])), min(abs(delta[k]), abs(delta[k + 1]))
            d[k + 1] = dmin / (w1 * (delta[k] / dmax) + w2 * (delta[k + 1] / dmax))

    def end(h0, h1, d0, d1):
        v = ((2 * h0 + h1) * d0 - h0 * d1) / (h0 + h1)
        if np.sign(v) != np.sign(d0):
            return 0.0
        if np.sign(d0) != np.sign(d1) and abs(v) > abs(3 * d0):
            return 3 * d0
        return v
    d[0] = end(h[0], h[1], delta[0], delta[1])
    d[-1] = end(h[-1], h[-2], delta[-1], delta[-2])
    return d


def pchip_eval(x, y, tq):
    """griddedInterpolant(x, y, 'pchip')(tq) for one row: cubic Hermite on the interval holding each
    query (end polynomials outside the grid), local form  y_k + s (d_k + s (c + s b))."""
    x, y = np.asarray(x, dtype=np.float64), np.asarray(y, dtype=np.float64)
    tq = np.atleast_1d(np.asarray(tq, dtype=np.float64))
    d = pchip_slopes(x, y)
    h = np.diff(x)
    delta = np.diff(y) / h
    k = np.clip(np.searchsorted(x, tq, side="right") - 1, 0, x.size - 2)
    s = tq - x[k]
    c = (3 * delta[k] - 2 * d[k] - d[k + 1]) / h[k]
    b = (d[k] - 2 * delta[k] + d[k + 1]) / (h[k] * h[k])
    return y[k] + s * (d[k] + s * (c + s * b))


def vector_interpolant(x, v, interpType):
    """functions/vectorInterpolant.m:1-12: one griddedInterpolant per row of v.  interpType 'pchip',
    'linear' (end values outside the knots) or 'previous'."""
    v = np.atleast_2d(np.asarray(v, dtype=np.float64))
    x = np.asarray(x, dtype=np.float64)

    def f(t):
        t = np.atleast_1d(np.asarray(t, dtype=np.float64))
        if interpType == "pchip":
            return np.stack([pchip_eval(x, row, t) for row in v])
        if interpType == "linear":
            if x.size == 1:
                return np.repeat(v[:, :1], t.size, axis=1)
            k = np.clip(np.searchsorted(x, t, side="right") - 1, 0, x.size - 2)
            tc = np.clip(t, x[0], x[-1])
            w = (tc - x[k]) / (x[k + 1] - x[k])
            return v[:, k] + w * (v[:, k + 1] - v[:, k])
        if interpType == "previous":
            k = np.clip(np.searchsorted(x, t, side="right") - 1, 0, x.size - 1)
            return v[:, k]
        raise ValueError(interpType)
    return f


def compute_uFunc(control, v):
    """control.compute_uFunc(v): PWConstantControl.m:58-61 ('previous' over intervalStarts),
    PWLinearControl.m:74-77 ('linear' over controlPts); ChebyshevControl has none in the reference --
    the basis recurrence (:23-30) evaluated at the query time."""
    v = np.asarray(v, dtype=np.float64).reshape(control.nControls, -1, order="F")
    if isinstance(control, PWConstantControl):
        return vector_interpolant(control.intervalStarts, v, "previous")
    if isinstance(control, PWLinearControl):
        return vector_interpolant(control.controlPts, v, "linear")

    def f(t):
        t = np.atleast_1d(np.asarray(t, dtype=np.float64))
        tau = 2 * (t - control.t0) / (control.t1 - control.t0) - 1
        K = v.shape[1]
        Bq = np.zeros((K, t.size))
        Bq[0] = 1
        if K > 1:
            Bq[1] = tau
        for i in range(2, K):
            Bq[i] = 2 * tau * Bq[i - 1] - Bq[i - 2]
        return v @ Bq
    return f


def compute_states_and_objective_legacy(prob, x0, t, u):
    """functions/compute_states_and_objective.m:5-28 (v1 API, uniform step), with the
    k4 defect (u index ``uHalf(:,i+1)``, :23) replaced by the intended ``u(:,i+1)``
    (SURVEY.md App. E).  ``prob`` is the v2 problem; F = [stateRHS; objective]."""
    x0 = np.atleast_1d(np.asarray(x0, dtype=np.float64))
    nT = len(t)
    h = t[1] - t[0]
    xAug = np.zeros((x0.shape[0] + 1, nT) + x0.shape[1:])
    xAug[:, 0] = np.concatenate([x0, np.zeros((1,) + x0.shape[1:])])
    tHalf = t[:-1] + h / 2
    uHalf = .5 * (u[:, :-1] + u[:, 1:])
    for i in range(nT - 1):
        k1 = prob.F(t[i], xAug[:, i], u[:, i])
        k2 = prob.F(tHalf[i], xAug[:, i] + .5 * h * k1, uHalf[:, i])
        k3 = prob.F(tHalf[i], xAug[:, i] + .5 * h * k2, uHalf[:, i])
        k4 = prob.F(t[i + 1], xAug[:, i] + h * k3, u[:, i + 1])
        xAug[:, i + 1] = xAug[:, i] + h * (k1 + 2 * k2 + 2 * k3 + k4) / 6
    return xAug[:-1], xAug[-1, -1]


def equilibrium_residual(prob, y, r):
    """functions/compute_equilibrium.m:10-21."""
    n, nc = prob.nStates, prob.nControls
    x, lam, u = y[:n], y[n:2 * n], y[2 * n:2 * n + nc]
    val = np.zeros_like(y)
    F = prob.F(0.0, np.concatenate([x, [0.0]]), u)
    val[:n] = F[:n]
    lam1 = np.concatenate([lam, [1.0]])
    dFdx = prob.dFdx_times_vec(0.0, np.concatenate([x, [0.0]]), u, lam1)
    val[n:2 * n] = r * lam - dFdx[:n]
    val[2 * n:] = prob.dFdu_times_vec(0.0, np.concatenate([x, [0.0]]), u, lam1)
    return val


def solve_equilibrium(prob, xGuess, lamGuess, uGuess, r, iters=60):
    """Newton on equilibrium_residual with a forward-difference Jacobian -- stands in
    for lsqnonlin (compute_equilibrium.m:23-27); only used to produce uStar."""
    y = np.concatenate([np.atleast_1d(xGuess), np.atleast_1d(lamGuess), np.atleast_1d(uGuess)]).astype(float)
    for _ in range(iters):
        f = equilibrium_residual(prob, y, r)
        Jm = np.empty((y.size, y.size))
        for k in range(y.size):
            e = np.zeros_like(y)
            e[k] = 1e-7 * max(1.0, abs(y[k]))
            Jm[:, k] = (equilibrium_residual(prob, y + e, r) - equilibrium_residual(prob, y - e, r)) / (2 * e[k])
        dy = np.linalg.solve(Jm, -f)
        y = y + dy
        if np.max(np.abs(dy)) < 1e-15:
            break
    n, nc = prob.nStates, prob.nControls
    return y[:n], y[n:2 * n], y[2 * n:]
