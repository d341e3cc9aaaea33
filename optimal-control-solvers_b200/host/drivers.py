"""Host mirrors of the reference functions that CALL the RK4 path.
They contain no arithmetic of their own: every sweep runs on the GPU through the integrator.

  nlp_objective                  functions/single_shooting.m:137-150 (the fmincon callback)
  compute_states_and_objective   functions/compute_states_and_objective.m (legacy v1 forward sweep)
  fb_sweep                       functions/fb_sweep.m:79-125 composed on the RK4 grid
"""
from __future__ import annotations

import numpy as np

__all__ = ["nlp_objective", "compute_states_and_objective", "fb_sweep", "single_shooting"]


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
    out = integ.single_shooting(prob, x0, v0, Lb, Ub, FreeInitStates=FreeInitStates, FreeStateBounds=FreeStateBounds,
                                max_iter=MaxIter, tol=TolX)
    J = out["J"]
    if getattr(prob, "MinMax", "Min") == "Max":
        J = -J
    return dict(J=J, v=out["v"], u=integ.get_control_u(0, J.size), converged=out["converged"], iters=out["iters"],
                integrator=integ, control=control)
