"""CUDA backend of make_from_symbolic (functions/make_from_symbolic.m).

The reference derives, with the Symbolic Math Toolbox, the Hamiltonian H = f + lam*g
(:11), the adjoint right-hand side -grad_x H (:14), dH/du (:17) and the control
characterisation solving dH/du = 0 (:20-23, clamped to the bounds at :111), and turns
them into MATLAB function handles with matlabFunction (:27-31).  This backend performs
the same derivations with SymPy and prints them as CUDA `__device__` functions -- the
`struct Problem` that csrc/rk4_kernels.cuh is instantiated on -- which the library
compiles for sm_100a with NVRTC (ocs_b200_set_problem_source).  It emits the v2 plugin
triple the RK4 path consumes (OCProblem/OCProblem.m:3-21):

    F                = [g; f]                                       (objective integrand in the last row)
    dFdx_times_vec   = (dg/dx)' v_x + v_J df/dx                      (last row 0)
    dFdu_times_vec   = (dg/du)' v_x + v_J df/du
    ControlChar      = clamp(solve(dFdu_times_vec(t,y,u,[lam;1]) = 0, u), bounds)

Input contract = the reference's (:1-8): symbolic objective and state right-hand side in
the variables t, x1..xn, lam1..lamn, u1..um and the parameters named by the fields of the
params struct, in field order.

Hoisting: every maximal sub-expression that is batch-uniform -- a function of (t, params)
when the parameters are shared by the batch (HM_TIME), additionally of u when the control
is shared too (HM_TIME_CONTROL) -- is evaluated once per grid point into the record's
table instead of once per trajectory (see rk4_kernels.cuh).
"""
from __future__ import annotations

import sympy as sp
from sympy.printing.c import C99CodePrinter

from ..host.problems import OCProblem
from .. import _lib
from .fma_forms import fma_forms, inline_fused_products, order_terms

__all__ = ["make_from_symbolic", "CodegenProblem"]


class _CudaPrinter(C99CodePrinter):
    """C printer with small integer powers as products (pow(double, int) is slow and inexact)."""

    def _print_Pow(self, expr):
        b, e = expr.base, expr.exp
        if e.is_Integer and 2 <= int(e) <= 4:
            s = self.parenthesize(b, 50)
            return "(" + "*".join([s] * int(e)) + ")"
        if e.is_Integer and -4 <= int(e) <= -1:
            s = self.parenthesize(b, 50)
            return "(1.0/(" + "*".join([s] * (-int(e))) + "))"
        if e == sp.Rational(1, 2):
            return f"sqrt({self._print(b)})"
        return super()._print_Pow(expr)

    def _print_Rational(self, expr):
        return f"({int(expr.p)}.0/{int(expr.q)}.0)"

    def _print_Integer(self, expr):
        return f"{int(expr)}.0" if abs(int(expr)) < 2 ** 53 else super()._print_Integer(expr)


class _FmaPrinter(_CudaPrinter):
    """Sums with their non-product terms first: every later product contracts into an FMA (fma_forms.order_terms)."""

    def _as_ordered_terms(self, expr, order=None):
        return order_terms(super()._as_ordered_terms(expr, order=order))


def _count_ops(exprs):
    """FLOPs of the CSE'd DAG: add/sub/mul/div/function = 1 each."""
    repl, red = sp.cse(list(exprs), optimizations="basic")
    total = 0
    for e in [r for _, r in repl] + list(red):
        total += int(sp.count_ops(e, visual=False))
    return total


class CodegenProblem(OCProblem):
    """An OCProblem whose right-hand sides exist as generated CUDA source."""

    def __init__(self, nStates, nControls, params, ControlBounds, source, sym, flops, MinMax="Min"):
        self.nStates, self.nControls = nStates, nControls
        self.param_names = list(params.keys())
        self.params = dict(params)
        import numpy as np
        self.ControlBounds = np.asarray(ControlBounds, dtype=np.float64).reshape(-1, 2)
        self.source = source
        self.sym = sym            # dict of SymPy objects (for tests / documentation)
        self.flops = flops        # dict C_F, C_X, C_U
        self.MinMax = MinMax

    def param_array(self):
        import numpy as np
        vals = [np.atleast_1d(np.asarray(self.params[k], dtype=np.float64)) for k in self.param_names]
        nb = max(v.size for v in vals) if vals else 1
        if nb == 1:
            return np.array([v[0] for v in vals])
        return np.stack([np.broadcast_to(v, (nb,)) for v in vals])

    def bind(self, lib, handle):
        _lib.check(lib.ocs_b200_set_problem_source(handle, self.source.encode(), self.nStates, self.nControls,
                                                   len(self.param_names), None), handle)


def _hoist(exprs, uniform_syms, grid_syms, table, prefix="tt"):
    """Replaces, in every expression, the maximal sub-expressions whose free symbols are all
    uniform and that depend on at least one grid symbol (t or a shared control) by table symbols.
    `table` (list of (symbol, expr)) is extended in place and shared between functions."""
    known = {e: s for s, e in table}

    def visit(e):
        fs = e.free_symbols
        if not e.is_Atom and fs and fs <= uniform_syms and (fs & grid_syms):
            if e not in known:
                s = sp.Symbol(f"{prefix}{len(table)}")
                table.append((s, e))
                known[e] = s
            return known[e]
        if e.is_Atom:
            return e
        # for Add/Mul keep the uniform part of the argument list together so that it is hoisted as one
        if isinstance(e, (sp.Add, sp.Mul)):
            uni = [a for a in e.args if a.free_symbols <= uniform_syms and not a.is_Number]
            rest = [a for a in e.args if a not in uni]
            grid_uni = [a for a in uni if a.free_symbols & grid_syms]
            if len(grid_uni) >= 1 and len(uni) >= 2 and rest:
                packed = e.func(*uni)
                return e.func(visit(packed), *[visit(a) for a in rest])
        return e.func(*[visit(a) for a in e.args])
    return [visit(sp.sympify(e)) for e in exprs]


def make_from_symbolic(symObjective, symStateRHS, nStates, nControls, params, ControlBounds, MinMax="Min",
                       solve_control=True, fma=True):
    """CUDA-backend twin of functions/make_from_symbolic.m:1.

    symObjective : SymPy expression (or string) of the running cost f(t, x, u)
    symStateRHS  : sequence of nStates expressions g_i(t, x, u)
    params       : dict name -> value (scalar, or an array of per-trajectory values); the order of the
                   keys is the order of the generated parameter vector p (make_from_symbolic.m:3-4)
    ControlBounds: nControls x 2
    fma          : also emit the FMA-shaped forms of fma_forms.py for the FMA-contracted build (the strict build
                   always evaluates the expressions as derived)
    """
    n, nc = int(nStates), int(nControls)
    t = sp.Symbol("t")
    x = sp.symbols(f"x1:{n + 1}")
    lam = sp.symbols(f"lam1:{n + 1}")
    u = sp.symbols(f"u1:{nc + 1}")
    pnames = list(params.keys())
    psyms = [sp.Symbol(k) for k in pnames]
    local = {s.name: s for s in (t, *x, *lam, *u, *psyms)}
    f = sp.sympify(symObjective, locals=local)
    g = [sp.sympify(e, locals=local) for e in symStateRHS]
    if len(g) != n:
        raise ValueError("symStateRHS must have nStates rows")

    vJ = sp.Symbol("vJ")
    vx = sp.symbols(f"v1:{n + 1}")
    # F = [g; f]
    F = list(g) + [f]
    # dFdx_times_vec(v) = (dg/dx)' v_x + v_J df/dx            (make_from_symbolic.m:11-14 with lam -> v_x, 1 -> v_J)
    X = [sum(sp.diff(g[i], x[k]) * vx[i] for i in range(n)) + vJ * sp.diff(f, x[k]) for k in range(n)]
    # dFdu_times_vec(v) = (dg/du)' v_x + v_J df/du            (:17)
    U = [sum(sp.diff(g[i], u[c]) * vx[i] for i in range(n)) + vJ * sp.diff(f, u[c]) for c in range(nc)]
    # ControlChar: dH/du = 0 with H = f + lam*g               (:11,:17,:20-23)
    H = f + sum(lam[i] * g[i] for i in range(n))
    dHdu = [sp.diff(H, u[c]) for c in range(nc)]
    CC = None
    if solve_control:
        try:
            sol = sp.solve(dHdu, list(u), dict=True)
            if sol:
                CC = [sp.simplify(sol[0][u[c]]) for c in range(nc)]
        except Exception:
            CC = None

    pr = _CudaPrinter()
    prf = _FmaPrinter()
    subs_common = {}
    for i in range(n):
        subs_common[x[i]] = sp.Symbol(f"y[{i}]")
        subs_common[vx[i]] = sp.Symbol(f"v[{i}]")
        subs_common[lam[i]] = sp.Symbol(f"lam[{i}]")
    subs_common[vJ] = sp.Symbol(f"v[{n}]")
    for c in range(nc):
        subs_common[u[c]] = sp.Symbol(f"u[{c}]")
    for k, s in enumerate(psyms):
        subs_common[s] = sp.Symbol(f"p[{k}]")

    def emit_body(exprs, out, table_syms):
        sub = dict(subs_common)
        for k, s in enumerate(table_syms):
            sub[s] = sp.Symbol(f"tt[{k}]")
        repl, red = sp.cse(list(exprs), symbols=sp.numbered_symbols("z"), optimizations="basic")
        lines = []
        for s, e in repl:
            lines.append(f"            const double {s} = {pr.doprint(e.xreplace(sub))};")
        for k, e in enumerate(red):
            lines.append(f"            {out}[{k}] = {pr.doprint(e.xreplace(sub))};")
        return "\n".join(lines)

    def emit_body_fast(defs, forms, out, table_syms):
        """The FMA-shaped forms (fma_forms.py): stage constants first, then the rows."""
        sub = dict(subs_common)
        for k, s in enumerate(table_syms):
            sub[s] = sp.Symbol(f"tt[{k}]")
        lines = []
        if defs:
            repl, red = sp.cse([d for _, d in defs], symbols=sp.numbered_symbols("zc"), optimizations="basic")
            for s, e in repl:
                lines.append(f"            const double {s} = {prf.doprint(e.xreplace(sub))};")
            for (s, _), e in zip(defs, red):
                lines.append(f"            const double {s} = {prf.doprint(e.xreplace(sub))};")
        repl, red = sp.cse(list(forms), symbols=sp.numbered_symbols("z"), optimizations="basic")
        repl, red = inline_fused_products(repl, red)
        for s, e in repl:
            lines.append(f"            const double {s} = {prf.doprint(e.xreplace(sub))};")
        for k, e in enumerate(red):
            lines.append(f"            {out}[{k}] = {prf.doprint(e.xreplace(sub))};")
        return "\n".join(lines)

    uniform = {0: set(), 1: {t, *psyms}, 2: {t, *psyms, *u}}
    grid = {0: set(), 1: {t}, 2: {t, *u}}
    dynamic = dict(F=(set(x), [*x]), X=({*x, *vx, vJ}, [*vx, vJ, *x]), U=({*x, *vx, vJ}, [*vx, vJ, *x]),
                   CC=({*x, *lam}, [*lam, *x]))
    variants = {}
    fast = {}            # fast[hm][fn] = (stage constants, forms) or None: the FMA-contracted build's bodies
    tables = {}
    for hm in (0, 1, 2):
        table = []
        if hm == 0:
            variants[hm] = dict(F=F, X=X, U=U, CC=CC)
        else:
            Fh = _hoist(F, uniform[hm], grid[hm], table)
            Xh = _hoist(X, uniform[hm], grid[hm], table)
            Uh = _hoist(U, uniform[hm], grid[hm], table)
            CCh = _hoist(CC, uniform[hm], grid[hm], table) if CC is not None else None
            variants[hm] = dict(F=Fh, X=Xh, U=Uh, CC=CCh)
        fast[hm] = {}
        for fn, exprs in (("F", F), ("X", X), ("U", U), ("CC", CC)):
            fast[hm][fn] = None
            if exprs is None or not fma:
                continue
            prep = (lambda e, hm=hm: _hoist([e], uniform[hm], grid[hm], [])[0]) if hm else None
            dyn, order = dynamic[fn]
            defs, forms = fma_forms(exprs, dyn, order, prefix="c", prep=prep)
            if all(a == b for a, b in zip(forms, exprs)):
                continue                                      # nothing cheaper than the derived expressions
            if hm:
                hoisted = _hoist([d for _, d in defs] + forms, uniform[hm], grid[hm], table)
                defs = [(s_, e) for (s_, _), e in zip(defs, hoisted[:len(defs)])]
                forms = hoisted[len(defs):]
            fast[hm][fn] = (defs, forms)
        tables[hm] = table

    def per_hm(fn, out):
        parts = []
        for hm in (0, 1, 2):
            exprs = variants[hm][fn]
            kw = "if" if hm == 0 else "else if"
            tsyms = [s for s, _ in tables[hm]]
            body = emit_body(exprs, out, tsyms)
            if fast[hm][fn] is not None:
                # re-associated forms for the FMA-contracted build; the strict build keeps the derived expressions
                body = ("#if defined(OCS_STRICT) && OCS_STRICT\n" + body + "\n#else\n" +
                        emit_body_fast(*fast[hm][fn], out, tsyms) + "\n#endif")
            parts.append(f"        {kw} constexpr (HM == {hm}) {{\n{body}\n        }}")
        return "\n".join(parts)

    def table_code():
        parts = []
        for hm in (1, 2):
            kw = "if" if hm == 1 else "else if"
            exprs = [e for _, e in tables[hm]]
            body = emit_body(exprs, "tt", []) if exprs else ""
            parts.append(f"        {kw} constexpr (HM == {hm}) {{\n{body}\n        }}")
        return "\n".join(parts)

    # Time exponentials by recurrence (per-trajectory parameters; rk4_kernels.cuh, ocs_ntexp): the HM_TIME table is a
    # function of (t, p).  Every exp(a*t) in it, a free of t, becomes one per-trajectory value E_k that the sweeps
    # advance by a multiplication per grid point; table_from_texp rebuilds the table entries from E (and t, p).
    texp_src = ""
    exps = []
    for _, e in tables[1]:
        for ex in sorted(e.atoms(sp.exp), key=str):
            arg = sp.expand(ex.args[0])
            rate = sp.diff(arg, t)
            if rate != 0 and not rate.has(t) and sp.simplify(arg - rate * t) == 0 and all(sp.simplify(rate - r0) != 0 for r0, _ in exps):
                exps.append((rate, ex))
    if exps and tables[1]:
        Esyms = [sp.Symbol(f"E[{k}]") for k in range(len(exps))]
        sub_p = {s_: sp.Symbol(f"p[{k}]") for k, s_ in enumerate(psyms)}
        rate_lines = "\n".join(f"        a[{k}] = {pr.doprint(r0.xreplace(sub_p))};" for k, (r0, _) in enumerate(exps))

        def with_E(e):
            def rep(ex):
                arg = sp.expand(ex.args[0])
                rate = sp.diff(arg, t)
                for k, (r0, _) in enumerate(exps):
                    if not rate.has(t) and sp.simplify(rate - r0) == 0 and sp.simplify(arg - rate * t) == 0:
                        return Esyms[k]
                return ex
            return e.replace(lambda q: isinstance(q, sp.exp), rep)
        tbl = [with_E(e) for _, e in tables[1]]
        repl, red = sp.cse(tbl, symbols=sp.numbered_symbols("w"), optimizations="basic")
        body = [f"        const double {s_} = {pr.doprint(e.xreplace(sub_p))};" for s_, e in repl]
        body += [f"        tt[{k}] = {pr.doprint(e.xreplace(sub_p))};" for k, e in enumerate(red)]
        texp_src = f"""
    // exp(a_k t) of the HM_TIME table, advanced along the grid by recurrence when the parameters are per trajectory
    static constexpr int ntexp = {len(exps)};
    __device__ __forceinline__ static void texp_rates(const double *p, double *a)
    {{
        (void)p;
{rate_lines}
    }}
    __device__ __forceinline__ static void table_from_texp(const double *E, double t, const double *p, double *tt)
    {{
        (void)E; (void)t; (void)p;
{chr(10).join(body)}
    }}
"""

    sig = "const double *tt, double t, const double *y, const double *u, const double *p"
    unused = "(void)tt; (void)t; (void)y; (void)u; (void)p;"
    lo_hi = "\n".join(f"            o[{c}] = fmin(bounds[{2 * c + 1}], fmax(bounds[{2 * c}], o[{c}]));" for c in range(nc))
    if CC is not None:
        cc_body = per_hm("CC", "o")
        cc_src = f"""    template <int HM>
    __device__ __forceinline__ static void ControlChar(const double *tt, double t, const double *y, const double *lam,
                                                       const double *p, const double *bounds, double *o)
    {{
        (void)tt; (void)t; (void)y; (void)lam; (void)p;
        const double *u = nullptr; (void)u;
{cc_body}
        {{
{lo_hi}
        }}
    }}"""
    else:
        cc_src = """    template <int HM>
    __device__ __forceinline__ static void ControlChar(const double *, double, const double *, const double *,
                                                       const double *, const double *, double *o)
    {
        for (int c = 0; c < nc; ++c) o[c] = nan("");   // dH/du = 0 has no closed-form solution: fb_sweep unavailable
    }"""
    source = f"""// generated by optimal-control-solvers_b200/codegen/make_from_symbolic.py -- do not edit
// states: {', '.join(map(str, x))}   controls: {', '.join(map(str, u))}   params: {', '.join(pnames)}
// f = {f}
// g = {g}
namespace ocs_gen {{
struct Problem {{
    static constexpr int n = {n}, nc = {nc}, np = {len(pnames)};
    __host__ __device__ static constexpr int ntt(int hm) {{ return hm == 1 ? {len(tables[1])} : (hm == 2 ? {len(tables[2])} : 0); }}

    // batch-uniform sub-expressions, once per grid point
    template <int HM>
    __device__ __forceinline__ static void table(double t, const double *u, const double *p, double *tt)
    {{
        (void)t; (void)u; (void)p; (void)tt;
{table_code()}
    }}

{texp_src}
    // F = [g; f]   (OCProblem.m:9-11)
    template <int HM>
    __device__ __forceinline__ static void F({sig}, double *o)
    {{
        {unused}
{per_hm("F", "o")}
    }}

    // dFdx' * v, last row (zero) implicit   (OCProblem.m:13-16)
    template <int HM>
    __device__ __forceinline__ static void X({sig}, const double *v, double *o)
    {{
        {unused} (void)v;
{per_hm("X", "o")}
    }}

    // dFdu' * v   (OCProblem.m:18-19)
    template <int HM>
    __device__ __forceinline__ static void U({sig}, const double *v, double *o)
    {{
        {unused} (void)v;
{per_hm("U", "o")}
    }}

    // argmin_u H clamped to ControlBounds   (make_from_symbolic.m:20-23,111)
{cc_src}
}};
}}  // namespace ocs_gen
"""
    flops = dict(C_F=_count_ops(F), C_X=_count_ops(X), C_U=_count_ops(U))
    sym = dict(t=t, x=x, lam=lam, u=u, params=psyms, vx=vx, vJ=vJ, f=f, g=g, F=F, X=X, U=U, H=H, dHdu=dHdu, CC=CC,
               tables=tables)
    return CodegenProblem(n, nc, params, ControlBounds, source, sym, flops, MinMax=MinMax)
