"""Host mirror of the reference's Control classes (same names, constructor arguments
and methods): Control/PWConstantControl.m, Control/PWLinearControl.m,
Control/ChebyshevControl.m.  The objects only build the batch-invariant basis matrix B
on the integrator's grid -- with MATLAB's linspace rounding and MATLAB's literal
comparisons, so that the control-grid index of every grid point is the reference's --
and hand it to the CUDA integrator, which fuses compute_u (u = v*B) and compute_dJdv
(dJdv = dJdu*B') into the sweeps.  compute_u / compute_dJdv below are the small
single-trajectory host forms kept for API completeness (they are what single_shooting
calls, functions/single_shooting.m:139-142); batched work never goes through them.
"""
from __future__ import annotations

import numpy as np

__all__ = ["matlab_linspace", "Control", "PWConstantControl", "PWLinearControl", "ChebyshevControl"]


def matlab_linspace(d1, d2, n):
    """linspace(d1,d2,n) with MATLAB's rounding, ``d1 + (0:n-1)*(d2-d1)/(n-1)`` evaluated
    as (i*(d2-d1))/(n-1) and both ends forced.  numpy.linspace rounds differently and moves
    PWConstant interval indices."""
    n = int(n)
    if n < 1:
        return np.zeros(0)
    if n == 1:
        return np.array([float(d2)])
    n1 = n - 1
    y = d1 + (np.arange(n1 + 1, dtype=np.float64) * (d2 - d1)) / n1
    y[0], y[-1] = d1, d2
    return y


class Control:
    """Control/Control.m:4-14 (the signatures the implementations actually use)."""

    nControls = 1
    B = None

    @property
    def nBasis(self):
        return self.B.shape[0]

    def compute_u(self, v):
        """v = reshape(v, nControls, []); u = v*B."""
        v = np.asarray(v, dtype=np.float64)
        return v.reshape(self.nControls, -1, order="F") @ self.B

    def compute_dJdv(self, dJdu):
        """dJdv = reshape(dJdu*B', [], 1)."""
        return (np.asarray(dJdu, dtype=np.float64).reshape(self.nControls, -1) @ self.B.T).reshape(-1, order="F")

    def _tile_bounds(self, controlBounds):
        cb = np.asarray(controlBounds, dtype=np.float64).reshape(-1, 2)
        K = self.nBasis
        Lb = (cb[:, :1] * np.ones((1, K))).reshape(-1, order="F")
        Ub = (cb[:, 1:] * np.ones((1, K))).reshape(-1, order="F")
        return Lb, Ub


class PWConstantControl(Control):
    """Control/PWConstantControl.m:12-61."""

    def __init__(self, t, nControlIntervals, nControls):
        t = np.asarray(t, dtype=np.float64)
        self.nControlIntervals = int(nControlIntervals)
        self.nControls = int(nControls)
        s = matlab_linspace(t[0], t[-1], self.nControlIntervals + 1)          # :15
        self.intervalStarts = s[:-1].copy()                                    # :16
        self.B = self.compute_basis_matrix(t)

    def compute_basis_matrix(self, t):
        K, s = self.nControlIntervals, self.intervalStarts
        B = np.zeros((K, t.size))
        for i in range(K - 1):                                                 # :32-35
            B[i] = (t >= s[i]) & (t < s[i + 1])
        B[K - 1] = t >= s[K - 1]                                               # :37
        return B

    def compute_nlp_bounds(self, controlBounds):                               # :21-27
        return self._tile_bounds(controlBounds)

    def compute_initial_v(self, u0):                                           # :53-55
        return np.tile(np.atleast_1d(np.asarray(u0, dtype=np.float64)), self.nControlIntervals)

    def ufunc_spec(self):
        """(kind, knots) of compute_uFunc for the device (ocs_b200_set_control_ufunc)."""
        return "previous", self.intervalStarts

    def compute_uFunc(self, v):                                                # :58-61 ('previous' interpolant)
        v = np.asarray(v, dtype=np.float64).reshape(self.nControls, -1, order="F")
        s = self.intervalStarts

        def uFunc(tq):
            k = np.clip(np.searchsorted(s, np.atleast_1d(tq), side="right") - 1, 0, len(s) - 1)
            return v[:, k]
        return uFunc


def _tent(xp, fp, t):
    """griddedInterpolant(xp, fp, 'linear', 'nearest') at t: linear inside the knots,
    nearest sample outside (PWLinearControl.m:35-49)."""
    out = np.empty_like(t)
    lo, hi = t < xp[0], t > xp[-1]
    out[lo], out[hi] = fp[0], fp[-1]
    mid = ~(lo | hi)
    tm = t[mid]
    k = np.clip(np.searchsorted(xp, tm, side="right") - 1, 0, len(xp) - 2)
    out[mid] = fp[k] + (tm - xp[k]) / (xp[k + 1] - xp[k]) * (fp[k + 1] - fp[k])
    return out


class PWLinearControl(Control):
    """Control/PWLinearControl.m:13-77."""

    def __init__(self, t, nControlPts, nControls):
        t = np.asarray(t, dtype=np.float64)
        self.nControlPts = int(nControlPts)
        self.nControls = int(nControls)
        self.controlPts = matlab_linspace(t[0], t[-1], self.nControlPts)       # :16
        self.B = self.compute_basis_matrix(t)

    def compute_basis_matrix(self, t):
        K, p = self.nControlPts, self.controlPts
        B = np.zeros((K, t.size))
        B[0] = _tent(p[0:2], np.array([1.0, 0.0]), t)                          # :35-37
        for i in range(1, K - 1):                                              # :40-44
            B[i] = _tent(p[i - 1:i + 2], np.array([0.0, 1.0, 0.0]), t)
        B[K - 1] = _tent(p[K - 2:K], np.array([0.0, 1.0]), t)                  # :47-49
        return B

    def compute_nlp_bounds(self, controlBounds):                               # :21-28
        return self._tile_bounds(controlBounds)

    def compute_initial_v(self, u0):                                           # :65-71
        u0 = np.atleast_1d(np.asarray(u0, dtype=np.float64))
        if u0.size == 1:
            return np.tile(u0, self.nControlPts)
        if u0.size == self.nControlPts:
            return u0.reshape(-1, order="F")
        raise ValueError("compute_initial_v: u0 must have length 1 or nControlPts")

    def ufunc_spec(self):
        return "linear", self.controlPts

    def compute_uFunc(self, v):                                                # :74-77 ('linear' interpolant)
        v = np.asarray(v, dtype=np.float64).reshape(self.nControls, -1, order="F")
        p = self.controlPts
        return lambda tq: np.stack([np.interp(np.atleast_1d(tq), p, v[c]) for c in range(self.nControls)])


class ChebyshevControl(Control):
    """Control/ChebyshevControl.m:13-48.  The reference class is unfinished (no
    compute_uFunc, empty compute_lincon); compute_u / compute_dJdv / compute_initial_v
    keep its numerics and compute_uFunc is added so that the class is usable."""

    def __init__(self, t, nControlBasis, nControls):
        t = np.asarray(t, dtype=np.float64)
        self.nControlBasis = int(nControlBasis)
        self.nControls = int(nControls)
        self.controlPts = matlab_linspace(t[0], t[-1], self.nControlBasis)     # :16
        self._t0, self._t1 = t[0], t[-1]
        self.B = self.compute_basis_matrix(t)

    def compute_basis_matrix(self, t):
        K = self.nControlBasis
        B = np.zeros((K, t.size))
        tTrans = 2 * (t - t[0]) / (t[-1] - t[0]) - 1                           # :23
        B[0] = 1                                                               # :25
        if K > 1:
            B[1] = tTrans                                                      # :26
        for i in range(2, K):
            B[i] = 2 * tTrans * B[i - 1] - B[i - 2]                            # :29
        return B

    def compute_initial_v(self, u0):                                           # :46-48
        u0 = np.atleast_1d(np.asarray(u0, dtype=np.float64))
        return np.concatenate([u0, np.zeros((self.nControlBasis - 1) * self.nControls)])

    def compute_nlp_bounds(self, ControlBounds):
        """A box on Chebyshev coefficients does not bound u(t): the coefficient vector is free and the control
        bounds enter as linear constraints (compute_lincon)."""
        nz = self.nControlBasis * self.nControls
        return np.full(nz, -np.inf), np.full(nz, np.inf)

    def compute_lincon(self, ControlBounds):
        """[A, b] of Control/ChebyshevControl.m:51-53, which the reference left empty: ControlBounds at every grid
        point, lb <= reshape(v, nc, []) * B <= ub, as A v <= b with v = vec of the (nc x K) coefficient matrix
        (functions/single_shooting.m:106-111).  The device solve applies exactly this set through
        CudaRK4Integrator.set_solve_control_bounds without forming A."""
        cb = np.asarray(ControlBounds, dtype=np.float64).reshape(-1, 2)
        Bk = np.kron(self.B.T, np.eye(self.nControls))           # u(c, j) = sum_k v(c + nc k) B(k, j)
        A = np.vstack([Bk, -Bk])
        G = self.B.shape[1]
        b = np.concatenate([np.tile(cb[:, 1], G), -np.tile(cb[:, 0], G)])
        return A, b

    grid_bounds_as_constraints = True

    def ufunc_spec(self):
        return "chebyshev", np.array([self._t0, self._t1])

    def compute_uFunc(self, v):
        v = np.asarray(v, dtype=np.float64).reshape(self.nControls, -1, order="F")

        def uFunc(tq):
            tq = np.atleast_1d(np.asarray(tq, dtype=np.float64))
            tau = 2 * (tq - self._t0) / (self._t1 - self._t0) - 1
            Bq = np.zeros((self.nControlBasis, tq.size))
            Bq[0] = 1
            if self.nControlBasis > 1:
                Bq[1] = tau
            for i in range(2, self.nControlBasis):
                Bq[i] = 2 * tau * Bq[i - 1] - Bq[i - 2]
            return v @ Bq
        return uFunc
