"""Host mirrors of the reference functions that CALL the RK4 path.
They contain no arithmetic of their own: every sweep runs on the GPU through the integrator.

  nlp_objective                  functions/single_shooting.m:137-150 (the fmincon callback)
  compute_states_and_objective   functions/compute_states_and_objective.m (legacy v1 forward sweep)
  fb_sweep                       functions/fb_sweep.m:79-125 composed on the RK4 grid
  compute_equilibrium            functions/compute_equilibrium.m, one steady state per parameter set
"""
from __future__ import annotations

import numpy as np

from .. import _lib

__all__ = ["nlp_objective", "compute_states_and_objective", "fb_sweep", "single_shooting", "compute_equilibrium"]


def nlp_objective(prob, integrator, control, x0, v, FreeInitStates=()):
    """[J, dJdv] = nlpObjective(v) of functions/single_shooting.m:137-150 for one trajectory or a batch.
    v: (nv,) / (nv, Bt); with free initial states its last rows hold x0(FreeInitStates) (:143-145) and
    the gradient gains lam(FreeInitStates, 1) (:147-148).  FreeInitStates are 0-based row indices."""
    if integrator._control is not control:
        integrator.attach_control(control)
    free = list(FreeInitStates)
    v = np.asarray(v, dtype=np.float64)
    x0 = np.array(x0, dtype=np.float64, copy=True)
    if x0.ndim == 0:
        x0 = x0.reshape(1)
    if not free:
        out = integrator.evaluate(prob, x0, v)
        J, dJdv = out["J"], out["dJdv"]
    else:
        nf = len(free)
        vc, vx = v[:-nf], v[-nf:]
        if x0.ndim == 1 and v.ndim == 2:
            x0 = np.repeat(x0[:, None], v.shape[1], axis=1)
        x0[free] = vx                                         # x0(FreeInitStates) = v(end-nFREE_STATES+1:end)
        out = integrator.evaluate(prob, x0, vc)
        J = out["J"]
        dJdv = np.concatenate([out["dJdv"], out["lam0"][free]], axis=0)   # [compute_dJdv(dJdu); lam(FreeInitStates,1)]
    if v.ndim == 1:
        return float(J[0]), dJdv[:, 0]
    return J, dJdv


def compute_states_and_objective(prob, x0, t, u, integrator_cls=None, **kw):
    """[x, J] = compute_states_and_objective(prob, x0, t, u)  (functions/compute_states_and_objective.m:1-30):
    the v1-API RK4 forward sweep with u sampled at the nodes and the midpoint control taken as the mean of
    its neighbours (:11).  The reference indexes uHalf(:,i+1) for k4 (:23), which is out of bounds on the last
    step and the wrong sample before; the intended u(:,i+1) is used.  The same device sweep
    serves it: nodes + averaged midpoints are exactly a control on the integrator's grid.  (The legacy code's
    tHalf = t + h/2 and h*(...)/6 differ from the class path in the last bit only)"""
    from .integrators import CudaRK4Integrator
    t = np.asarray(t, dtype=np.float64).ravel()
    u = np.asarray(u, dtype=np.float64)
    if u.ndim == 1:
        u = u.reshape(1, -1)
    G = 2 * t.size - 1
    ug = np.empty((u.shape[0], G) + u.shape[2:])
    ug[:, 0::2] = u
    ug[:, 1::2] = .5 * (u[:, :-1] + u[:, 1:])                 # uHalf   :11
    integ = (integrator_cls or CudaRK4Integrator)(t, **kw)
    try:
        x, J = integ.compute_states(prob, x0, ug, return_x=True)
    finally:
        integ.close()
    return x[:-1], J                                          # x = xAug(1:end-1,:)   :27


def fb_sweep(prob, integrator, x0, u0=None, options=None):
    """soln = fb_sweep(prob, x0, tspan, options) on the RK4 grid (functions/fb_sweep.m).  options mirrors the
    reference's struct fields uRelTol, uAbsTol, nSWEEPS, u0 (fb_sweep.m:16-20,26-67) plus `relax` (1 = the
    reference's pure replacement).  Returns a dict with J, u (grid values of the converged control, = soln.u
    sampled on integrator.t), x, lam (grids on request via integrator.fetch_*), sweeps, converged; like the
    reference, a trajectory that did not converge within nSWEEPS reports converged = 0 (fb_sweep.m:77-87)."""
    opt = dict(uRelTol=1e-7, uAbsTol=1e-7, nSWEEPS=50, relax=1.0, costate="discrete")
    opt.update(options or {})
    nc = prob.nControls
    G = integrator.t.size
    if u0 is None:
        u0 = opt.get("u0", 0.0)
    u0 = np.asarray(u0, dtype=np.float64)
    if u0.ndim == 0 or u0.size == nc:                         # a constant initial control
        u0 = np.repeat(u0.reshape(nc, 1), G, axis=1)
    cb = np.asarray(prob.ControlBounds, dtype=np.float64).reshape(-1, 2)
    u0 = np.minimum(cb[:, 1].reshape((nc,) + (1,) * (u0.ndim - 1)),
                    np.maximum(cb[:, 0].reshape((nc,) + (1,) * (u0.ndim - 1)), u0))
    out = integrator.fb_sweep(prob, x0, u0, uRelTol=opt["uRelTol"], uAbsTol=opt["uAbsTol"],
                              nSWEEPS=opt["nSWEEPS"], relax=opt["relax"], costate=opt["costate"])
    out["u"] = integrator.get_control_u(0, out["J"].size)
    return out


def compute_equilibrium(prob, xGuess, lamGuess, uGuess, lb=None, ub=None, r=0.0, params=None, device=0, math="fast",
                        MaxIter=4000, tol=1e-13, integrator=None):
    """[xStar, lamStar, uStar, resnorm, residual, exitflag] = compute_equilibrium(prob, xGuess, lamGuess, uGuess,
    lb, ub, r) (functions/compute_equilibrium.m:1-33), batched: one steady state per parameter set, each solved
    by its own GPU thread (ocs_b200_compute_equilibrium).  `params` (np, nsets) overrides prob's parameters for
    a sweep; guesses may be columns (shared) or (., nsets); r scalar or (nsets,).  With one parameter set the
    outputs have the reference's shapes (column vectors, scalar resnorm/exitflag)."""
    from .integrators import CudaRK4Integrator
    own = integrator is None
    integ = integrator or CudaRK4Integrator(np.array([0.0, 1.0]), device=device, math=math)
    try:
        integ._bind(prob)
        n, nc = prob.nStates, prob.nControls
        P = np.asarray(prob.param_array() if params is None else params, dtype=np.float64)
        P = P.reshape(P.shape[0], -1)
        g = [np.asarray(a, dtype=np.float64).reshape(k, -1) for a, k in ((xGuess, n), (lamGuess, n), (uGuess, nc))]
        ng = max(a.shape[1] for a in g)
        guess = _lib.f64(np.concatenate([np.broadcast_to(a, (a.shape[0], ng)) for a in g], axis=0))
        rr = _lib.f64(np.atleast_1d(np.asarray(r, dtype=np.float64)).ravel())
        nsets = max(P.shape[1], ng, rr.size)
        Pf = _lib.f64(P)
        NY = 2 * n + nc
        lbf = None if lb is None else _lib.f64(np.asarray(lb, dtype=np.float64).ravel())
        ubf = None if ub is None else _lib.f64(np.asarray(ub, dtype=np.float64).ravel())
        for a in (lbf, ubf):
            if a is not None and a.size != NY:
                raise ValueError("lb / ub must have 2*nStates + nControls entries")
        y = np.empty((NY, nsets), order="F")
        resnorm, residual = np.empty(nsets), np.empty((NY, nsets), order="F")
        flag, iters = np.empty(nsets, dtype=np.int32), np.empty(nsets, dtype=np.int32)
        integ._ck(integ._lib.ocs_b200_compute_equilibrium(
            integ._h, _lib.dp(Pf), Pf.shape[0], Pf.shape[1], nsets, _lib.dp(guess), ng, _lib.dp(lbf), _lib.dp(ubf), _lib.dp(rr), rr.size,
            int(MaxIter), float(tol), _lib.dp(y), _lib.dp(resnorm), _lib.dp(residual), _lib.ip(flag), _lib.ip(iters)))
    finally:
        if own:
            integ.close()
    if nsets == 1:
        return y[:n, 0].copy(), y[n:2 * n, 0].copy(), y[2 * n:, 0].copy(), float(resnorm[0]), residual[:, 0].copy(), int(flag[0])
    return y[:n], y[n:2 * n], y[2 * n:], resnorm, residual, flag


def single_shooting(prob, x0, tspan, nCONTROL_PTS, Control=None, Integrator=None, u0=0.0, FreeInitStates=(),
                    FreeStateBounds=None, TolX=1e-6, MaxIter=300, **integrator_kw):
    """soln = single_shooting(prob, x0, tspan, nCONTROL_PTS, Name, Value...) (functions/single_shooting.m) for a
    batch of initial states x0 (n x Bt), solved on the device.  Same defaults as the reference: RK4 integrator on
    tspan (:41-45), PWLinearControl(integrator.t, nCONTROL_PTS, nControls) (:48-52), u0 clamped to ControlBounds
    (:56), v0 = control.compute_initial_v(u0) (:83), bounds from control.compute_nlp_bounds (:88-89), free initial
    states appended to v (:85,:91-95); MinMax = 'Max' only negates the reported J (:117-119).  fmincon's SQP is
    replaced by the device-side projected-gradient solve, so iterates differ but minimisers agree.
    Returns dict(J, v, u (grid values of the optimal control), converged, iters, integrator, control);
    soln.x / soln.lam of the reference (:128-130) are integrator.fetch_states / fetch_adjoints."""
    from .controls import PWLinearControl
    from .integrators import CudaRK4Integrator
    integ = Integrator if Integrator is not None else CudaRK4Integrator(tspan, **integrator_kw)
    nc = prob.nControls
    control = Control if Control is not None else PWLinearControl(integ.t, nCONTROL_PTS, nc)
    cb = np.asarray(prob.ControlBounds, dtype=np.float64).reshape(-1, 2)
    u0 = np.minimum(cb[:, 1], np.maximum(cb[:, 0], np.broadcast_to(np.asarray(u0, dtype=np.float64).ravel(), (nc,))))
    v0 = control.compute_initial_v(u0 if nc > 1 else u0[0])
    Lb, Ub = control.compute_nlp_bounds(cb)
    integ.attach_control(control)
    # [A, b] = control.compute_lincon(prob.ControlBounds) (:106-111): the reference's PW controls have none; a control
    # whose coefficient box does not bound u(t) (ChebyshevControl) constrains the control on the grid instead
    integ.set_solve_control_bounds(cb if getattr(control, "grid_bounds_as_constraints", False) else None)
    out = integ.single_shooting(prob, x0, v0, Lb, Ub, FreeInitStates=FreeInitStates, FreeStateBounds=FreeStateBounds,
                                max_iter=MaxIter, tol=TolX)
    J = out["J"]
    if getattr(prob, "MinMax", "Min") == "Max":
        J = -J
    # soln.u / soln.x / soln.lam (single_shooting.m:128-130) are function handles in the reference; here
    # soln["eval"](tq[, b0, nb]) evaluates all three on the device for a range of the batch
    return dict(J=J, v=out["v"], u=integ.get_control_u(0, J.size), converged=out["converged"], iters=out["iters"],
                integrator=integ, control=control,
                eval=lambda tq, b0=0, nb=None, want=("x", "lam", "u"): integ.eval_solution(tq, b0, nb, 0, want))
