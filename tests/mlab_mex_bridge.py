"""`ocs_b200_mex` for the MATLAB-subset interpreter (oracle/mlab) -- TEST INFRASTRUCTURE.

The MATLAB classes in optimal-control-solvers_b200/matlab call `ocs_b200_mex(command, handle, args...)`.  MATLAB is not
in the image, so tests/test_gpu_matlab_glue.py runs those .m files under oracle/mlab; this module supplies
the built-in they call, mapping every command onto the same C-ABI entry point -- argument for argument -- as
the real gateway optimal-control-solvers_b200/matlab/ocs_b200_mex.c does (which is itself compiled and driven through
tests/stubs/mex.h by tests/test_mex_gateway.py)."""
import ctypes

import numpy as np


def make_mex(ocs):
    L = ocs._lib
    lib = L.load()

    def arr(a):
        a = np.asarray(a, dtype=np.float64)
        return L.f64(a) if a.size else None

    def hptr(h):
        v = int(np.asarray(h).reshape(-1)[0])
        if v == 0:
            raise RuntimeError("integrator handle was already deleted")
        return ctypes.c_void_p(v)

    def sc(a):
        return float(np.asarray(a).reshape(-1)[0])

    def check(rc, h):
        if rc != 0:
            msg = lib.ocs_b200_last_error(h)
            raise RuntimeError(f"ocs_b200 error {rc}: {msg.decode() if msg else ''}")

    def col(x):
        return np.asarray(x, dtype=np.float64)

    def mex(cmd, *a, nargout=1):
        if cmd == "create":
            ts, te = arr(np.asarray(a[0]).ravel()), arr(np.asarray(a[1]).ravel())
            h = ctypes.c_void_p()
            devs = [int(d) for d in np.asarray(a[2]).reshape(-1)]
            if len(devs) == 1:
                rc = lib.ocs_b200_create(ts.size - 1, L.dp(ts), (te.size - 1) if te is not None else 0, L.dp(te),
                                         devs[0], int(sc(a[3])), ctypes.byref(h))
            else:       # CudaRK4Integrator(tspan, 'Device', 0:7), as the real gateway does
                rc = lib.ocs_b200_create_multi(ts.size - 1, L.dp(ts), (te.size - 1) if te is not None else 0, L.dp(te),
                                               len(devs), (ctypes.c_int * len(devs))(*devs), int(sc(a[3])), ctypes.byref(h))
            if rc != 0:
                raise RuntimeError(f"ocs_b200 error {rc}: {lib.ocs_b200_last_error(None).decode()}")
            return np.array([[float(h.value)]])
        h = hptr(a[0])
        if cmd == "destroy":
            lib.ocs_b200_destroy(h)
            return ()
        if cmd == "get_grid":
            N = int(sc(a[2]))
            t, hh = np.empty(2 * N + 1), np.empty(N)
            check(lib.ocs_b200_get_grid(h, int(sc(a[1])), L.dp(t), L.dp(hh)), h)
            return (t.reshape(1, -1), hh.reshape(1, -1))[:max(nargout, 1)]
        if cmd == "set_problem_builtin":
            check(lib.ocs_b200_set_problem_builtin(h, int(sc(a[1]))), h)
            return ()
        if cmd == "set_problem_source":
            inc = a[5] if len(a) > 5 and isinstance(a[5], str) and a[5] else None
            check(lib.ocs_b200_set_problem_source(h, a[1].encode(), int(sc(a[2])), int(sc(a[3])), int(sc(a[4])),
                                                  inc.encode() if inc else None), h)
            return ()
        if cmd == "set_params":
            p = arr(a[1])
            check(lib.ocs_b200_set_params(h, L.dp(p), p.shape[0], p.shape[1]), h)
            return ()
        if cmd == "set_tail_control":
            u = arr(np.asarray(a[1]).ravel())
            check(lib.ocs_b200_set_tail_control(h, L.dp(u), u.size), h)
            return ()
        if cmd == "set_control_basis":
            B = arr(a[1])
            check(lib.ocs_b200_set_control_basis(h, 0 if B is None else B.shape[0], L.dp(B)), h)
            return ()
        if cmd == "set_control_v":
            v = arr(a[1])
            check(lib.ocs_b200_set_control_v(h, L.dp(v), v.shape[1]), h)
            return ()
        if cmd == "set_control_u":
            u = col(a[1])
            nb = u.shape[2] if u.ndim > 2 else 1
            uf = L.f64(u)
            check(lib.ocs_b200_set_control_u(h, L.dp(uf), nb), h)
            return ()
        if cmd == "compute_states":
            x0 = arr(a[1])
            n, Bt = x0.shape
            J, xT, st = np.empty((1, Bt)), np.empty((n, Bt), order="F"), np.empty(Bt, dtype=np.int32)
            check(lib.ocs_b200_compute_states(h, L.dp(x0), Bt, L.dp(J), L.dp(xT), L.ip(st)), h)
            return (J, xT, st.astype(np.float64).reshape(1, -1))[:max(nargout, 1)]
        if cmd == "compute_adjoints":
            lamT = arr(a[1]) if np.asarray(a[1]).size else None
            m, Bt, nv = int(sc(a[2])), int(sc(a[3])), int(sc(a[4]))
            lam0 = np.empty((m, Bt), order="F")
            dJdv = np.empty((nv, Bt), order="F") if (nargout > 1 and nv > 0) else None
            check(lib.ocs_b200_compute_adjoints(h, L.dp(lamT), lamT.shape[1] if lamT is not None else 0, L.dp(dJdv),
                                                L.dp(lam0)), h)
            return (lam0, dJdv if dJdv is not None else np.zeros((0, 0)))[:max(nargout, 1)]
        if cmd == "evaluate":
            x0 = arr(a[1])
            lamT = arr(a[2]) if np.asarray(a[2]).size else None
            nv = int(sc(a[3]))
            n, Bt = x0.shape
            J, xT, st = np.empty((1, Bt)), np.empty((n, Bt), order="F"), np.empty(Bt, dtype=np.int32)
            lam0, g = np.empty((n + 1, Bt), order="F"), np.empty((nv, Bt), order="F")
            check(lib.ocs_b200_evaluate(h, L.dp(x0), Bt, L.dp(lamT), lamT.shape[1] if lamT is not None else 0, L.dp(J),
                                        L.dp(xT), L.ip(st), L.dp(g) if nv > 0 else None, L.dp(lam0)), h)
            return (J, xT, st.astype(np.float64).reshape(1, -1), lam0, g)[:max(nargout, 1)]
        if cmd == "fetch_states":
            seg, b0, nb, m, N = (int(sc(z)) for z in a[1:6])
            x = np.empty((m, N + 1, nb), order="F")
            check(lib.ocs_b200_fetch_states(h, seg, b0, nb, L.dp(x)), h)
            return x[:, :, 0] if nb == 1 else x
        if cmd == "fetch_adjoints":
            seg, b0, nb, m, N, nc = (int(sc(z)) for z in a[1:7])
            lam = np.empty((m, N + 1, nb), order="F")
            du = np.empty((nc, 2 * N + 1, nb), order="F") if nargout > 1 else None
            check(lib.ocs_b200_fetch_adjoints(h, seg, b0, nb, L.dp(lam), L.dp(du)), h)
            sq = (lambda z: z[:, :, 0] if nb == 1 else z)
            return (sq(lam), sq(du) if du is not None else None)[:max(nargout, 1)]
        if cmd == "get_control_u":
            b0, nb, nc, G = (int(sc(z)) for z in a[1:5])
            u = np.empty((nc, G, nb), order="F")
            check(lib.ocs_b200_get_control_u(h, b0, nb, L.dp(u)), h)
            return u[:, :, 0] if nb == 1 else u
        if cmd == "fb_sweep":
            x0, u0, bounds = arr(a[1]), col(a[2]), arr(a[3])
            Bt = x0.shape[1]
            nbu = u0.shape[2] if u0.ndim > 2 else 1
            u0f = L.f64(u0)
            J, mc = np.empty((1, Bt)), np.empty((1, Bt))
            sw, cv = np.empty(Bt, dtype=np.int32), np.empty(Bt, dtype=np.int32)
            check(lib.ocs_b200_set_fb_costate(h, int(sc(a[8])) if len(a) > 8 else 0), h)
            check(lib.ocs_b200_fb_sweep(h, L.dp(x0), Bt, L.dp(u0f), nbu, L.dp(bounds), sc(a[4]), sc(a[5]), int(sc(a[6])),
                                        sc(a[7]), L.dp(J), L.ip(sw), L.ip(cv), L.dp(mc)), h)
            return (J, sw.astype(float).reshape(1, -1), cv.astype(float).reshape(1, -1), mc)[:max(nargout, 1)]
        if cmd == "set_control_ufunc":
            kn = arr(np.asarray(a[2]).ravel())
            check(lib.ocs_b200_set_control_ufunc(h, int(sc(a[1])), kn.size, L.dp(kn)), h)
            return ()
        if cmd == "eval_solution":
            seg, b0, nb = int(sc(a[1])), int(sc(a[2])), int(sc(a[3]))
            tq = arr(np.asarray(a[4]).ravel())
            n, nc = int(sc(a[5])), int(sc(a[6]))
            x = np.empty((n, tq.size, nb), order="F")
            lam = np.empty((n, tq.size, nb), order="F") if nargout > 1 else None
            u = np.empty((nc, tq.size, nb), order="F") if nargout > 2 else None
            check(lib.ocs_b200_eval_solution(h, seg, b0, nb, tq.size, L.dp(tq), L.dp(x), L.dp(lam), L.dp(u)), h)
            return (x, lam, u)[:max(nargout, 1)]
        if cmd == "compute_equilibrium":
            P, guess = arr(a[1]), arr(a[2])
            lb = arr(np.asarray(a[3]).ravel()) if np.asarray(a[3]).size else None
            ub = arr(np.asarray(a[4]).ravel()) if np.asarray(a[4]).size else None
            r = arr(np.asarray(a[5], dtype=float).ravel())
            nsets = max(P.shape[1], guess.shape[1], r.size)
            ny = guess.shape[0]
            y, rn, rs = np.empty((ny, nsets), order="F"), np.empty((1, nsets)), np.empty((ny, nsets), order="F")
            fl = np.empty(nsets, dtype=np.int32)
            check(lib.ocs_b200_compute_equilibrium(h, L.dp(P), P.shape[0], P.shape[1], nsets, L.dp(guess), guess.shape[1], L.dp(lb),
                                                   L.dp(ub), L.dp(r), r.size, int(sc(a[6])), sc(a[7]), L.dp(y), L.dp(rn),
                                                   L.dp(rs), L.ip(fl), None), h)
            return (y, rn, rs, fl.astype(float).reshape(1, -1))[:max(nargout, 1)]
        if cmd == "single_shooting":
            x0, v0, Lb, Ub = arr(a[1]), arr(a[2]), arr(np.asarray(a[3]).ravel()), arr(np.asarray(a[4]).ravel())
            free = np.asarray(a[5]).ravel().astype(np.int32)
            flb = arr(np.asarray(a[6]).ravel()) if free.size else None
            fub = arr(np.asarray(a[7]).ravel()) if free.size else None
            Bt, nv = x0.shape[1], Lb.size
            v = np.empty((nv + free.size, Bt), order="F")
            J = np.empty((1, Bt))
            it, cv = np.empty(Bt, dtype=np.int32), np.empty(Bt, dtype=np.int32)
            check(lib.ocs_b200_single_shooting(h, L.dp(x0), Bt, L.dp(v0), v0.shape[1], L.dp(Lb), L.dp(Ub), int(free.size),
                                               L.ip(free) if free.size else None, L.dp(flb), L.dp(fub), int(sc(a[8])),
                                               sc(a[9]), L.dp(v), L.dp(J), L.ip(it), L.ip(cv)), h)
            return (v, J, it.astype(float).reshape(1, -1), cv.astype(float).reshape(1, -1))[:max(nargout, 1)]
        if cmd == "set_solve_control_bounds":
            cb = np.asarray(a[1], dtype=float)
            if cb.size == 0:
                check(lib.ocs_b200_set_solve_control_bounds(h, None, 0, 0.0, 0.0, 0), h)
            else:
                cb = arr(cb.reshape(-1, 2))
                check(lib.ocs_b200_set_solve_control_bounds(h, L.dp(cb), cb.shape[0], sc(a[2]) if len(a) > 2 else 0.0,
                                                            sc(a[3]) if len(a) > 3 else 0.0, int(sc(a[4])) if len(a) > 4 else 0), h)
            return ()
        if cmd == "get_solve_violation":
            Bt = int(lib.ocs_b200_get_batch(h))
            out = np.empty((1, Bt))
            check(lib.ocs_b200_get_solve_violation(h, L.dp(out)), h)
            return (out,)
        if cmd == "reduce_J":
            val, idx = ctypes.c_double(), ctypes.c_long()
            check(lib.ocs_b200_reduce_J(h, int(sc(a[1])), ctypes.byref(val), ctypes.byref(idx)), h)
            return (np.array([[val.value]]), np.array([[float(idx.value + 1 if idx.value >= 0 else 0)]]))[:max(nargout, 1)]
        raise RuntimeError(f"unknown command {cmd!r}")

    mex._wants_nargout = True
    return mex
