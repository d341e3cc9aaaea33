"""Builds the oracle-side twin of a generated problem: the SymPy expressions the CUDA backend
printed, lambdified to NumPy (test infrastructure)."""
import numpy as np
import sympy as sp


def oracle_problem(oracle, prob):
    s = prob.sym
    n, nc = prob.nStates, prob.nControls
    pvals = [float(np.atleast_1d(prob.params[k])[0]) for k in prob.param_names]
    args = [s["t"], *s["x"], *s["u"], *s["vx"], s["vJ"], *s["lam"], *s["params"]]
    fF = sp.lambdify(args, s["F"], "numpy")
    fX = sp.lambdify(args, s["X"], "numpy")
    fU = sp.lambdify(args, s["U"], "numpy")
    fC = sp.lambdify(args, s["CC"], "numpy") if s["CC"] is not None else None

    def pack(t, y, u, v=None, lam=None):
        y = np.asarray(y, dtype=np.float64)
        u = np.asarray(u, dtype=np.float64)
        shape = np.broadcast(y[0], u[0], np.asarray(t, dtype=np.float64)).shape
        z = np.zeros(shape)
        xs = [y[i] + z for i in range(n)]
        us = [u[c] + z for c in range(nc)]
        vs = [(v[i] + z) if v is not None else z for i in range(n)]
        vJ = (v[n] + z) if v is not None else z
        ls = [(lam[i] + z) if lam is not None else z for i in range(n)]
        return [np.asarray(t, dtype=np.float64) + z, *xs, *us, *vs, vJ, *ls, *pvals], z

    def F(t, y, u):
        a, z = pack(t, y, u)
        return np.stack([np.asarray(r) + z for r in fF(*a)])

    def X(t, y, u, v):
        a, z = pack(t, y, u, v=v)
        return np.stack([np.asarray(r) + z for r in fX(*a)] + [z])

    def U(t, y, u, v):
        a, z = pack(t, y, u, v=v)
        return np.stack([np.asarray(r) + z for r in fU(*a)])

    def CC(t, x, lam):
        t = np.asarray(t, dtype=np.float64)
        uu = np.zeros((nc,) + np.broadcast(x[0], t).shape)
        a, z = pack(t, x, uu, lam=lam)
        out = np.stack([np.asarray(r) + z for r in fC(*a)])
        lo, hi = prob.ControlBounds[:, 0], prob.ControlBounds[:, 1]
        shp = (nc,) + (1,) * (out.ndim - 1)
        return np.minimum(hi.reshape(shp), np.maximum(lo.reshape(shp), out))

    return oracle.GenericOCProblem(n, nc, F, X, U, prob.ControlBounds, CC if fC is not None else None)
