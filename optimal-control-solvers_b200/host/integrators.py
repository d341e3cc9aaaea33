"""Host mirror of the Integrator plugin (Integrator/Integrator.m:6-15) backed by the
CUDA sweeps: CudaRK4Integrator replaces RK4Integrator (Integrator/RK4Integrator.m) and
CudaRK4InfiniteIntegrator replaces RK4InfiniteIntegrator
(Integrator/RK4InfiniteIntegrator.m).  Same property (`t`), same methods and argument
meaning; every array may carry a trailing batch axis (one page per trajectory).

    integrator = CudaRK4Integrator(tspan)
    x, J       = integrator.compute_states(prob, x0, u)         # u: nc x G [x Bt]
    lam, dJdu  = integrator.compute_adjoints(prob, u)           # after compute_states, like the reference

Batched fused path (compute_u and compute_dJdv folded into the sweeps):

    integrator.attach_control(control)                          # a host Control object (its B matrix)
    out = integrator.evaluate(prob, x0, v)                      # J, dJdv, lam0, xT, status for Bt trajectories

All arithmetic runs on the GPU through libocs_b200.so; there is no host fallback.
"""
from __future__ import annotations

import ctypes

import numpy as np

from .. import _lib

__all__ = ["CudaRK4Integrator", "CudaRK4InfiniteIntegrator"]

_MATH = {"fast": _lib.MATH_FAST, "strict": _lib.MATH_STRICT}


class CudaRK4Integrator:
    """RK4Integrator(tspan) on the GPU (Integrator/RK4Integrator.m:16-25)."""

    def __init__(self, tspan, device=0, math="fast", _tspanExtra=None):
        """device: a CUDA ordinal, or a sequence of distinct ordinals -- the batch is then sharded over those GPUs
        inside the one handle (ocs_b200_create_multi; MATLAB: CudaRK4Integrator(tspan, 'Device', 0:7))."""
        self._lib = _lib.load()
        tspan = _lib.f64(np.asarray(tspan, dtype=np.float64).ravel())
        self.nSTEPS = tspan.size - 1
        self._h = ctypes.c_void_p()
        te = None if _tspanExtra is None else _lib.f64(np.asarray(_tspanExtra, dtype=np.float64).ravel())
        self._N2 = 0 if te is None else te.size - 1
        self.devices = [int(d) for d in np.atleast_1d(device)]
        if np.ndim(device) == 0:
            rc = self._lib.ocs_b200_create(self.nSTEPS, _lib.dp(tspan), self._N2, _lib.dp(te), int(device),
                                           _MATH[math], ctypes.byref(self._h))
        else:
            devs = (ctypes.c_int * len(self.devices))(*self.devices)
            rc = self._lib.ocs_b200_create_multi(self.nSTEPS, _lib.dp(tspan), self._N2, _lib.dp(te), len(self.devices), devs,
                                                 _MATH[math], ctypes.byref(self._h))
        if rc != 0:
            self._h = ctypes.c_void_p()
            _lib.check(rc, None)
        self.t = np.empty(2 * self.nSTEPS + 1)
        self.h = np.empty(self.nSTEPS)
        _lib.check(self._lib.ocs_b200_get_grid(self._h, 0, _lib.dp(self.t), _lib.dp(self.h)), self._h)
        self._prob = None
        self._params = None
        self._control = None
        self._last_u = None
        self._Bt = 0
        self._single = True
        self._n = self._nc = 0

    # ------------------------------------------------------------------ plumbing
    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.ocs_b200_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):   # delete(obj)
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _ck(self, rc):
        _lib.check(rc, self._h)

    def _bind(self, prob):
        if self._prob is not prob:
            prob.bind(self._lib, self._h)
            self._prob = prob
            self._params = None
            self._n, self._nc = prob.nStates, prob.nControls
            self._after_bind(prob)
            if self._control is not None:   # generated problems are specialised on the basis
                self.attach_control(self._control)
        p = np.asarray(prob.param_array(), dtype=np.float64)
        if self._params is None or p.shape != self._params.shape or not np.array_equal(p, self._params):
            pf = _lib.f64(p.reshape(p.shape[0], -1))
            self._ck(self._lib.ocs_b200_set_params(self._h, _lib.dp(pf), pf.shape[0], pf.shape[1]))
            self._params = p.copy()

    def _after_bind(self, prob):
        pass

    def attach_control(self, control):
        """Fuse control.compute_u / compute_dJdv (u = v*B, dJdv = dJdu*B') into the sweeps."""
        self._control = control
        self._last_u = None
        if control is None:
            self._ck(self._lib.ocs_b200_set_control_basis(self._h, 0, None))
            return
        B = _lib.f64(control.B)
        if B.shape[1] != self.t.size:
            raise ValueError("control basis was not built on this integrator's grid t")
        self._ck(self._lib.ocs_b200_set_control_basis(self._h, B.shape[0], _lib.dp(B)))
        if hasattr(control, "ufunc_spec"):      # compute_uFunc on the device (eval_solution)
            kind, knots = control.ufunc_spec()
            kf = _lib.f64(np.asarray(knots, dtype=np.float64).ravel())
            self._ck(self._lib.ocs_b200_set_control_ufunc(self._h, {"previous": 1, "linear": 2, "chebyshev": 3}[kind],
                                                          kf.size, _lib.dp(kf)))

    def control_index(self):
        """(k_index, nnz) per grid point as the device stores them (0-based)."""
        G = self.t.size
        k = np.empty(G, dtype=np.int32)
        z = np.empty(G, dtype=np.int32)
        self._ck(self._lib.ocs_b200_get_control_index(self._h, _lib.ip(k), _lib.ip(z)))
        return k, z

    def _set_u(self, u):
        u = np.asarray(u, dtype=np.float64)
        G = self.t.size
        if u.ndim == 1:
            u = u.reshape(1, -1)
        if u.shape[1] != G:
            raise ValueError(f"u must be nControls x {G} (integrator.t), got {u.shape}")
        nb = 1 if u.ndim == 2 else u.shape[2]
        uf = _lib.f64(u)
        self._ck(self._lib.ocs_b200_set_control_u(self._h, _lib.dp(uf), nb))

    def _set_v(self, v):
        v = np.asarray(v, dtype=np.float64)
        nb = 1 if v.ndim == 1 else v.shape[1]
        if self._control is not None and v.shape[0] != self._nc * self._control.nBasis:
            raise ValueError(f"v must have nControls*K = {self._nc * self._control.nBasis} rows, got {v.shape[0]}")
        vf = _lib.f64(v)
        self._ck(self._lib.ocs_b200_set_control_v(self._h, _lib.dp(vf), nb))

    def _x0(self, x0):
        x0 = np.asarray(x0, dtype=np.float64)
        single = x0.ndim <= 1
        if x0.ndim == 0:
            x0 = x0.reshape(1, 1)
        elif x0.ndim == 1:
            x0 = x0.reshape(-1, 1)
        return _lib.f64(x0), single

    # ------------------------------------------------------------------ Integrator interface
    def compute_states(self, prob, x0, u=None, v=None, return_x=None):
        """[x, J] = compute_states(obj, prob, x0, u)   (RK4Integrator.m:28-56).
        u: nControls x G (shared) or nControls x G x Bt; alternatively v (coefficients of the
        attached control).  x is returned for single trajectories (return_x to override)."""
        self._bind(prob)
        if (u is None) == (v is None):
            raise ValueError("pass exactly one of u (grid values) or v (control coefficients)")
        if u is not None:
            self._set_u(u)
            self._last_u = u
        else:
            self._set_v(v)
            self._last_u = None
        x0f, single = self._x0(x0)
        Bt = x0f.shape[1]
        J = np.empty(Bt)
        self.xT = np.empty((self._n, Bt), order="F")
        self.status = np.empty(Bt, dtype=np.int32)
        self._ck(self._lib.ocs_b200_compute_states(self._h, _lib.dp(x0f), Bt, _lib.dp(J), _lib.dp(self.xT),
                                                   _lib.ip(self.status)))
        self._Bt, self._single = Bt, single
        if return_x is None:
            return_x = single
        x = self.fetch_states(0, Bt) if return_x else None
        if single:
            return (x[:, :, 0] if x is not None else None), float(J[0])
        return x, J

    def compute_adjoints(self, prob, u=None, lamT=None, nargout=2, grids=None):
        """[lam, dJdu] = compute_adjoints(obj, prob, u [, lamT])   (RK4Integrator.m:59-121).
        Uses the states of the last compute_states (RK4Integrator.m:8-10).  For a batch,
        grids=False returns (lam(:,1,:), dJdv) instead of the full grids."""
        self._bind(prob)
        if u is not None and u is not self._last_u:
            self._set_u(u)
            self._last_u = u
        Bt = self._Bt
        m = self._n + 1
        lt, lnb = None, 0
        if lamT is not None:
            lt = np.asarray(lamT, dtype=np.float64)
            lt = _lib.f64(lt.reshape(m, -1))
            lnb = lt.shape[1]
        if grids is None:
            grids = self._single
        lam0 = np.empty((m, Bt), order="F")
        dJdv = None
        if not grids and self._control is not None:
            dJdv = np.empty((self._nc * self._control.nBasis, Bt), order="F")
        self._ck(self._lib.ocs_b200_compute_adjoints(self._h, _lib.dp(lt), lnb, _lib.dp(dJdv), _lib.dp(lam0)))
        if not grids:
            return lam0, dJdv
        lam, dJdu = self.fetch_adjoints(0, Bt, want_dJdu=nargout > 1)
        if self._single:
            lam = lam[:, :, 0]
            dJdu = None if dJdu is None else dJdu[:, :, 0]
        return (lam, dJdu) if nargout > 1 else lam

    # ------------------------------------------------------------------ batched fused path
    def evaluate(self, prob, x0, v, lamT=None, gradient=True):
        """One single_shooting objective/gradient evaluation (functions/single_shooting.m:137-150)
        for every trajectory: control.compute_u -> compute_states -> compute_adjoints ->
        control.compute_dJdv, fused.  v: (nc*K,) shared or (nc*K, Bt).  Returns a dict."""
        if self._control is None:
            raise RuntimeError("attach_control(control) first")
        self._bind(prob)
        self._set_v(v)
        self._last_u = None
        x0f, single = self._x0(x0)
        Bt = x0f.shape[1]
        m = self._n + 1
        out = dict(J=np.empty(Bt), xT=np.empty((self._n, Bt), order="F"), status=np.empty(Bt, dtype=np.int32))
        lt, lnb = None, 0
        if gradient:
            out["dJdv"] = np.empty((self._nc * self._control.nBasis, Bt), order="F")
            out["lam0"] = np.empty((m, Bt), order="F")
            if lamT is not None:
                lt = _lib.f64(np.asarray(lamT, dtype=np.float64).reshape(m, -1))
                lnb = lt.shape[1]
        self._ck(self._lib.ocs_b200_evaluate(self._h, _lib.dp(x0f), Bt, _lib.dp(lt), lnb, _lib.dp(out["J"]),
                                             _lib.dp(out["xT"]), _lib.ip(out["status"]), _lib.dp(out.get("dJdv")),
                                             _lib.dp(out.get("lam0"))))
        self._Bt, self._single = Bt, single
        return out

    # ------------------------------------------------------------------ device-resident variants
    def set_stream(self, cuda_stream):
        """Launch on the given cudaStream_t (an int, e.g. torch.cuda.current_stream().cuda_stream)."""
        self._ck(self._lib.ocs_b200_set_stream(self._h, ctypes.c_void_p(int(cuda_stream) if cuda_stream else None)))

    def synchronize(self):
        self._ck(self._lib.ocs_b200_synchronize(self._h))

    def evaluate_dev(self, prob, x0_ptr, Bt, v=None, v_ptr=None, J_ptr=0, xT_ptr=0, status_ptr=0,
                     dJdv_ptr=0, lam0_ptr=0, lamT_ptr=0, lamT_nbatch=0, gradient=True, bind=True):
        """evaluate() on caller-owned DEVICE buffers in SoA layout [row][Bt] (raw pointers as ints):
        nothing crosses PCIe.  v: shared host coefficients, or v_ptr: device [(nc*K)][Bt].
        Asynchronous on the handle's stream."""
        if self._control is None:
            raise RuntimeError("attach_control(control) first")
        if bind:        # (a loop over the same problem object may skip the re-comparison of its parameter arrays)
            self._bind(prob)
        if v_ptr:
            self._ck(self._lib.ocs_b200_set_control_v_dev(self._h, ctypes.c_void_p(int(v_ptr)), int(Bt)))
        else:
            self._set_v(v)
        self._last_u = None
        P = lambda a: ctypes.c_void_p(int(a)) if a else None
        if gradient and (dJdv_ptr or lam0_ptr):
            # nlpObjective as one call: both sweeps are queued together (two lanes when that pays, ocs_b200.h)
            self._ck(self._lib.ocs_b200_evaluate_dev(self._h, P(x0_ptr), int(Bt), P(lamT_ptr), int(lamT_nbatch), P(J_ptr),
                                                     P(xT_ptr), P(status_ptr), P(dJdv_ptr), P(lam0_ptr)))
            self._Bt, self._single = int(Bt), False
            return
        self._ck(self._lib.ocs_b200_compute_states_dev(self._h, P(x0_ptr), int(Bt), P(J_ptr), P(xT_ptr), P(status_ptr)))
        self._Bt, self._single = int(Bt), False
        if gradient:
            self._ck(self._lib.ocs_b200_compute_adjoints_dev(self._h, P(lamT_ptr), int(lamT_nbatch), P(dJdv_ptr),
                                                             P(lam0_ptr)))

    def evaluate_resident(self, prob, x0=None, v=None, gradient=True):
        """nlpObjective on data that stays on the device(s) (ocs_b200_evaluate_resident): x0 (n x Bt) is uploaded and
        kept, or -- x0 None -- the resident initial states are swept again, e.g. with new coefficients v.
        Asynchronous; read the results with results() or gather_results_dev()."""
        if self._control is None:
            raise RuntimeError("attach_control(control) first")
        self._bind(prob)
        if v is not None:
            self._set_v(v)
        self._last_u = None
        if x0 is None:
            self._ck(self._lib.ocs_b200_evaluate_resident(self._h, None, 0, int(gradient)))
        else:
            x0f, single = self._x0(x0)
            self._ck(self._lib.ocs_b200_evaluate_resident(self._h, _lib.dp(x0f), x0f.shape[1], int(gradient)))
            self._Bt, self._single = x0f.shape[1], single

    def results(self, gradient=True):
        """J, xT, status (and dJdv, lam0) of the resident batch as host arrays (ocs_b200_get_results)."""
        Bt, m = self._Bt, self._n + 1
        out = dict(J=np.empty(Bt), xT=np.empty((self._n, Bt), order="F"), status=np.empty(Bt, dtype=np.int32))
        if gradient:
            out["dJdv"] = np.empty((self._nc * self._control.nBasis, Bt), order="F")
            out["lam0"] = np.empty((m, Bt), order="F")
        self._ck(self._lib.ocs_b200_get_results(self._h, _lib.dp(out["J"]), _lib.dp(out["xT"]), _lib.ip(out["status"]),
                                                _lib.dp(out.get("dJdv")), _lib.dp(out.get("lam0"))))
        return out

    def gather_results_dev(self, J_ptr=0, xT_ptr=0):
        """J [Bt] and x(T) [n][Bt] into device buffers on the root device (raw pointers as ints): the gather-to-root of
        a multi-device handle (NCCL send/recv or peer copies), a device copy on one device.  Asynchronous."""
        P = lambda a: ctypes.c_void_p(int(a)) if a else None
        self._ck(self._lib.ocs_b200_gather_results_dev(self._h, P(J_ptr), P(xT_ptr)))

    def shard_offsets(self):
        """Trajectory offsets of the resident batch per device (ndev + 1 entries)."""
        nd = ctypes.c_int()
        cap = len(self.devices) + 1
        devs = (ctypes.c_int * cap)()
        off = (ctypes.c_long * cap)()
        self._ck(self._lib.ocs_b200_get_devices(self._h, ctypes.byref(nd), devs, off, cap))
        return list(devs)[:nd.value], list(off)[:nd.value + 1]

    # ------------------------------------------------------------------ grids on request
    def fetch_states(self, b0=0, nb=None, segment=0):
        """x of RK4Integrator.m:54 for trajectories [b0, b0+nb): (m, N+1, nb)."""
        nb = self._Bt - b0 if nb is None else nb
        N = self.nSTEPS if segment == 0 else self._N2
        x = np.empty((self._n + 1, N + 1, nb), order="F")
        self._ck(self._lib.ocs_b200_fetch_states(self._h, segment, b0, nb, _lib.dp(x)))
        return x

    def fetch_adjoints(self, b0=0, nb=None, segment=0, want_dJdu=True):
        """(lam, dJdu) grids of RK4Integrator.m:68-121 for trajectories [b0, b0+nb)."""
        nb = self._Bt - b0 if nb is None else nb
        N = self.nSTEPS if segment == 0 else self._N2
        lam = np.empty((self._n + 1, N + 1, nb), order="F")
        dJdu = np.empty((self._nc, 2 * N + 1, nb), order="F") if want_dJdu else None
        self._ck(self._lib.ocs_b200_fetch_adjoints(self._h, segment, b0, nb, _lib.dp(lam), _lib.dp(dJdu)))
        return lam, dJdu

    def eval_solution(self, tq, b0=0, nb=None, segment=0, want=("x", "lam", "u")):
        """soln.x(tq), soln.lam(tq), soln.u(tq) of functions/single_shooting.m:128-130 (fb_sweep.m:120-124) for
        trajectories [b0, b0+nb), interpolated on the device: pchip through the node states / costates
        (vectorInterpolant.m), the control's uFunc (or pchip through a per-trajectory control grid).
        Returns a dict of arrays rows x nq x nb (x nq for a single-trajectory call)."""
        nb = (self._Bt - b0) if nb is None else nb
        tqf = _lib.f64(np.asarray(tq, dtype=np.float64).ravel())
        nq = tqf.size
        rows = dict(x=self._n, lam=self._n, u=self._nc)
        out = {k: np.empty((rows[k], nq, nb), order="F") for k in want}
        self._ck(self._lib.ocs_b200_eval_solution(self._h, int(segment), int(b0), int(nb), nq, _lib.dp(tqf),
                                                  _lib.dp(out.get("x")), _lib.dp(out.get("lam")), _lib.dp(out.get("u"))))
        if self._single and nb == 1:
            out = {k: a[:, :, 0] for k, a in out.items()}
        return out

    def get_control_u(self, b0=0, nb=1):
        u = np.empty((self._nc, self.t.size, nb), order="F")
        self._ck(self._lib.ocs_b200_get_control_u(self._h, b0, nb, _lib.dp(u)))
        return u

    def reduce_J(self, op="min"):
        """Batch reduction of the last J on the device: (value, index)."""
        code = {"sum": _lib.REDUCE_SUM, "min": _lib.REDUCE_MIN, "max": _lib.REDUCE_MAX}[op]
        val = ctypes.c_double()
        idx = ctypes.c_long()
        self._ck(self._lib.ocs_b200_reduce_J(self._h, code, ctypes.byref(val), ctypes.byref(idx)))
        return val.value, idx.value

    # ------------------------------------------------------------------ fb_sweep on the RK4 grid
    def fb_sweep(self, prob, x0, u0=None, uRelTol=1e-7, uAbsTol=1e-7, nSWEEPS=50, relax=1.0, costate="discrete", v0=None):
        """The sweep loop of functions/fb_sweep.m:79-115 with the RK4 forward / discrete-adjoint
        sweeps in place of odevr7, for every trajectory.  u0: nc x G shared or nc x G x Bt.
        costate="continuous" integrates the v1 API's adjointRHS (make_from_symbolic.m:11-14,
        compute_x_lam.m:11-12) backwards by RK4 instead of using the discrete adjoint."""
        self._bind(prob)
        if costate not in ("discrete", "continuous"):
            raise ValueError("costate must be 'discrete' or 'continuous'")
        self._ck(self._lib.ocs_b200_set_fb_costate(self._h, 1 if costate == "continuous" else 0))
        if u0 is None:          # start from coefficients: u0 = control.compute_u(v0), evaluated on the device
            if v0 is None or self._control is None:
                raise ValueError("fb_sweep needs u0, or v0 together with an attached control")
            self._set_v(v0)
        x0f, single = self._x0(x0)
        Bt = x0f.shape[1]
        if u0 is None:
            u0f, nbu = None, 0
        else:
            u0 = np.asarray(u0, dtype=np.float64)
            if u0.ndim == 1:
                u0 = u0.reshape(1, -1)
            nbu = 1 if u0.ndim == 2 else u0.shape[2]
            u0f = _lib.f64(u0)
        cb = _lib.f64(np.asarray(prob.ControlBounds, dtype=np.float64).reshape(-1, 2))
        out = dict(J=np.empty(Bt), sweeps=np.empty(Bt, dtype=np.int32), converged=np.empty(Bt, dtype=np.int32),
                   maxChange=np.empty(Bt))
        self._ck(self._lib.ocs_b200_fb_sweep(self._h, _lib.dp(x0f), Bt, _lib.dp(u0f), nbu, _lib.dp(cb),
                                             float(uRelTol), float(uAbsTol), int(nSWEEPS), float(relax),
                                             _lib.dp(out["J"]), _lib.ip(out["sweeps"]), _lib.ip(out["converged"]),
                                             _lib.dp(out["maxChange"])))
        self._Bt, self._single = Bt, single
        self._last_u = None
        return out

    def set_solve_control_bounds(self, ControlBounds, rho0=0.0, feas_tol=0.0, max_outer=0):
        """ControlBounds (nc x 2) at every grid point as linear constraints of single_shooting -- what a finished
        ChebyshevControl.compute_lincon would hand to fmincon (single_shooting.m:106-111); None switches them off
        (ocs_b200_set_solve_control_bounds)."""
        if ControlBounds is None:
            self._ck(self._lib.ocs_b200_set_solve_control_bounds(self._h, None, 0, 0.0, 0.0, 0))
            return
        cb = _lib.f64(np.asarray(ControlBounds, dtype=np.float64).reshape(-1, 2))
        self._ck(self._lib.ocs_b200_set_solve_control_bounds(self._h, _lib.dp(cb), cb.shape[0], float(rho0), float(feas_tol), int(max_outer)))

    def solve_violation(self):
        """Largest violation of the grid control bounds per trajectory after the last constrained solve."""
        out = np.empty(self._Bt)
        self._ck(self._lib.ocs_b200_get_solve_violation(self._h, _lib.dp(out)))
        return out

    # ------------------------------------------------------------------ batched single shooting on the device
    def single_shooting(self, prob, x0, v0, Lb, Ub, FreeInitStates=(), FreeStateBounds=None, max_iter=200, tol=1e-6,
                        out=None):
        """The fmincon solve of functions/single_shooting.m:114 for every trajectory at once, on the device
        (spectral projected gradient; see ocs_b200_single_shooting).  v0: (nc*K,) shared or (nc*K, Bt).
        Returns dict(v, J, iters, converged): v is (nc*K + nfree, Bt).  `out` may supply those four arrays
        (e.g. views of pinned host memory; v column-major)."""
        if self._control is None:
            raise RuntimeError("attach_control(control) first")
        self._bind(prob)
        x0f, single = self._x0(x0)
        Bt = x0f.shape[1]
        v0 = np.asarray(v0, dtype=np.float64)
        nb0 = 1 if v0.ndim == 1 else v0.shape[1]
        v0f = _lib.f64(v0)
        nv = self._nc * self._control.nBasis
        Lbf, Ubf = _lib.f64(np.asarray(Lb, dtype=np.float64).ravel()), _lib.f64(np.asarray(Ub, dtype=np.float64).ravel())
        if Lbf.size != nv or Ubf.size != nv:
            raise ValueError("Lb / Ub must have nControls*K entries (control.compute_nlp_bounds)")
        free = np.asarray(list(FreeInitStates), dtype=np.int32)
        nf = free.size
        flb = fub = None
        if nf:
            fb = np.asarray(FreeStateBounds, dtype=np.float64).reshape(nf, 2)
            flb, fub = _lib.f64(fb[:, 0].copy()), _lib.f64(fb[:, 1].copy())
        if out is None:
            out = dict(v=np.empty((nv + nf, Bt), order="F"), J=np.empty(Bt), iters=np.empty(Bt, dtype=np.int32),
                       converged=np.empty(Bt, dtype=np.int32))
        elif (out["v"].shape != (nv + nf, Bt) or not out["v"].flags.f_contiguous or out["v"].dtype != np.float64
              or out["J"].shape != (Bt,) or out["J"].dtype != np.float64
              or any(out[k].shape != (Bt,) or out[k].dtype != np.int32 for k in ("iters", "converged"))):
            raise ValueError("out: v (nc*K+nfree, Bt) float64 column-major, J (Bt,) float64, iters/converged (Bt,) int32")
        self._ck(self._lib.ocs_b200_single_shooting(
            self._h, _lib.dp(x0f), Bt, _lib.dp(v0f), nb0, _lib.dp(Lbf), _lib.dp(Ubf), nf,
            _lib.ip(free) if nf else None, _lib.dp(flb), _lib.dp(fub), int(max_iter), float(tol),
            _lib.dp(out["v"]), _lib.dp(out["J"]), _lib.ip(out["iters"]), _lib.ip(out["converged"])))
        self._Bt, self._single = Bt, single
        self._last_u = None
        return out

    # ------------------------------------------------------------------ measurement
    def set_profiling(self, on=True):
        self._ck(self._lib.ocs_b200_set_profiling(self._h, int(on)))

    def kernel_stats(self, reset=True):
        """Device time per kernel family since the last reset (needs set_profiling(True)): ms_forward, ms_reverse
        (= reverse sweep + projection), and the split ms_reverse_sweep / ms_project / ms_update."""
        ms = (ctypes.c_double * 5)()
        n = ctypes.c_long()
        self._ck(self._lib.ocs_b200_kernel_stats_ex(self._h, int(reset), ms, 5, ctypes.byref(n)))
        return dict(ms_forward=ms[0], ms_reverse=ms[1] + ms[2], ms_reverse_sweep=ms[1], ms_project=ms[2], ms_update=ms[3],
                    ms_fb_loop=ms[4], launches=n.value)


class CudaRK4InfiniteIntegrator(CudaRK4Integrator):
    """RK4InfiniteIntegrator(tspan, tspanExtra, uStar) on the GPU
    (Integrator/RK4InfiniteIntegrator.m:12-30): the main horizon chained with a tail under
    the constant control uStar; J = J1 + J2 and the tail's lam(:,1) is the terminal costate of
    the main horizon."""

    def __init__(self, tspan, tspanExtra, uStar, device=0, math="fast"):
        super().__init__(tspan, device=device, math=math, _tspanExtra=tspanExtra)
        self.uStar = np.atleast_1d(np.asarray(uStar, dtype=np.float64)).ravel()
        self.t2 = np.empty(2 * self._N2 + 1)
        _lib.check(self._lib.ocs_b200_get_grid(self._h, 1, _lib.dp(self.t2), None), self._h)

    def _after_bind(self, prob):
        us = _lib.f64(self.uStar)
        self._ck(self._lib.ocs_b200_set_tail_control(self._h, _lib.dp(us), us.size))

    def compute_adjoints(self, prob, u=None, nargout=2, grids=None):
        # RK4InfiniteIntegrator.m:27-30 takes no lamT: the tail sweep supplies it
        return super().compute_adjoints(prob, u, None, nargout=nargout, grids=grids)
