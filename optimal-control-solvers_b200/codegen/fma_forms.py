"""FMA-shaped forms of the generated right-hand sides (fast arithmetic build only).

matlabFunction (functions/make_from_symbolic.m:27-31) prints the derived expressions as they come out of the symbolic
engine, and so did the first CUDA backend: `d*(-2*x1 + x2 + x6) - q*u1*x1 + x1*(m - x1)` is six FP64 instructions on a
machine whose FP64 pipe issues one fused multiply-add per lane per slot, although `x1*(c0 - x1) + d*x2 + d*x6` with the
stage constant `c0 = m - 2*d - q*u1` is four.  The sweep kernels of a several-state problem are bound by exactly that
pipe (profiles/r2_notes.md: 92 % of their issued instructions are FP64), so the instruction count of F / dFdx' v /
dFdu' v is their run time.

For every output expression this module builds candidate forms and keeps the cheapest under an FMA-aware cost:

  * the expression as derived (always a candidate; the strict build prints nothing else);
  * Horner form in the per-trajectory ("dynamic") variables -- states and adjoint weights -- of the expanded
    expression, with every coefficient (a function of t, the parameters and the stage's control: constant within one
    right-hand-side call) pulled out as a named stage constant shared by all rows;
  * the expanded terms grouped by stage constant, `sum_k c_k * (dynamic polynomial)`, which is the cheaper shape for
    rows like `-q*(v1*x1 + v2*x2 + v3*x3) + c*vJ`.

Powers of sums such as `(x1 - xs)**2` and non-polynomial functions of dynamic variables stay atomic: expanding them
would trade instructions for cancellation.  The forms differ from the derived ones by re-association only; results
move by rounding (parity tests: 1e-10 against the lambdified derived expressions, as for every fast-build kernel).
"""
from __future__ import annotations

import sympy as sp

__all__ = ["fma_cost", "dag_cost", "fma_forms", "order_terms", "inline_fused_products"]

_EXPENSIVE = 40      # exp / log / pow: tens of FP64 instructions
_DIVIDE = 10


def fma_cost(e) -> int:
    """FP64 instructions of `e` as nvcc contracts it (a*b + c is one), sums printed with their non-product terms
    first (order_terms).  Negations ride on operand modifiers and cost nothing."""
    e = sp.sympify(e)
    if e.is_Atom:
        return 0
    if isinstance(e, sp.Add):
        args = e.args
        inner = 0
        products = 0
        for a in args:
            c, r = a.as_coeff_Mul()
            if a.is_Number:
                continue
            if isinstance(r, sp.Mul) or (c not in (1, -1) and not r.is_Number):
                m = _mul_count(a)
                inner += m[0]
                products += 1 if m[1] > 0 else 0
                inner += m[1]
            else:
                inner += fma_cost(r)
        adds = len(args) - 1
        return inner + adds - min(products, adds)
    if isinstance(e, sp.Mul):
        m = _mul_count(e)
        return m[0] + m[1]
    if isinstance(e, sp.Pow):
        b, p = e.base, e.exp
        if p.is_Integer and 2 <= int(p) <= 4:
            return fma_cost(b) + int(p) - 1
        if p.is_Integer and -4 <= int(p) <= -1:
            return fma_cost(b) + (-int(p) - 1) + _DIVIDE
        if p == sp.Rational(1, 2):
            return fma_cost(b) + _DIVIDE
        return fma_cost(b) + fma_cost(p) + _EXPENSIVE
    return sum(fma_cost(a) for a in e.args) + _EXPENSIVE


def _mul_count(e):
    """(cost of the factors, multiplications) of a product; a coefficient of -1 is free."""
    c, r = e.as_coeff_Mul()
    factors = list(r.args) if isinstance(r, sp.Mul) else [r]
    inner = sum(fma_cost(f) for f in factors)
    nf = len(factors) + (0 if c in (1, -1) else 1)
    return inner, max(nf - 1, 0)


def order_terms(terms):
    """Print order of a sum: atoms and non-products first, so that every product after the first term contracts
    into an FMA (nvcc does not re-associate FP64 additions)."""
    def key(a):
        c, r = a.as_coeff_Mul()
        is_product = isinstance(r, sp.Mul) or (c not in (1, -1) and not r.is_Number)
        return (1 if is_product else 0)
    return sorted(terms, key=key)           # stable: keeps SymPy's order within each class


class _Consts:
    """Stage constants c0, c1, ... (deduplicated up to sign)."""

    def __init__(self, prefix):
        self.prefix = prefix
        self.defs = []                      # (symbol, expr)
        self.known = {}

    def get(self, coeff):
        """A symbol (possibly negated) or number standing for `coeff` in a printed row."""
        coeff = sp.sympify(coeff)
        if coeff.is_Number or coeff.is_Symbol:
            return coeff
        if (-coeff).is_Symbol:
            return coeff
        key = sp.expand(coeff)
        if key in self.known:
            return self.known[key]
        if -key in self.known:
            return -self.known[-key]
        # keep a leading minus sign outside the constant: -2*xs*E is -(c) with c = 2*xs*E
        if key.could_extract_minus_sign():
            s = sp.Symbol(f"{self.prefix}{len(self.defs)}")
            self.defs.append((s, -coeff))
            self.known[-key] = s
            return -s
        s = sp.Symbol(f"{self.prefix}{len(self.defs)}")
        self.defs.append((s, coeff))
        self.known[key] = s
        return s


def _atomise(e, dyn):
    """Replaces every sub-expression that must not be expanded -- a power of a sum, a non-polynomial function or a
    division involving dynamic variables -- by a dummy generator.  Returns (expression, {dummy: original})."""
    back = {}

    def visit(a):
        if a.is_Atom:
            return a
        if not (a.free_symbols & dyn):
            return a
        keep = False
        if isinstance(a, sp.Pow):
            b, p = a.base, a.exp
            if not (p.is_Integer and int(p) >= 1):
                keep = True
            elif isinstance(b, sp.Add):
                keep = True
        elif not isinstance(a, (sp.Add, sp.Mul)):
            keep = True
        if keep:
            for d, o in back.items():
                if o == a:
                    return d
            d = sp.Dummy(f"g{len(back)}")
            back[d] = a
            return d
        return a.func(*[visit(x) for x in a.args])
    return visit(e), back


def _candidates(e, dyn, gens_order, consts_factory):
    """[(form, new stage constants)] for one output expression."""
    out = [(e, [])]
    try:
        ea, back = _atomise(sp.sympify(e), dyn)
        gens = [g for g in gens_order if g in ea.free_symbols] + list(back.keys())
        if not gens:
            return out
        poly = sp.Poly(sp.expand(ea), *gens)
    except (sp.PolynomialError, sp.CoercionFailed, NotImplementedError):
        return out
    terms = poly.terms()
    if len(terms) > 64:
        return out
    mono = lambda m: sp.Mul(*[g ** k for g, k in zip(gens, m) if k])

    # Horner form with named coefficients
    cs = consts_factory()
    hexpr = sp.Add(*[cs.get(c) * mono(m) for m, c in terms])
    try:
        h = sp.horner(hexpr, *gens) if any(sum(m) > 0 for m, _ in terms) else hexpr
    except (sp.PolynomialError, NotImplementedError):
        h = hexpr
    out.append((h.xreplace(back), cs))

    # grouped by stage constant: sum_k c_k * poly_k(dynamic)
    cs2 = consts_factory()
    groups = {}
    order = []
    for m, c in terms:
        num, rest = sp.sympify(c).as_coeff_Mul()
        if rest.is_Number:
            num, rest = num * rest, sp.Integer(1)
        key = sp.expand(rest)
        sign = 1
        if key not in groups and -key in groups:
            key, sign = -key, -1
        if key not in groups:
            groups[key] = []
            order.append((key, rest if sign == 1 else -rest))
        groups[key].append(sign * num * mono(m))
    parts = []
    for key, rest in order:
        inner = sp.Add(*groups[key])
        try:
            inner_h = sp.horner(inner, *gens) if inner.free_symbols else inner
        except (sp.PolynomialError, NotImplementedError):
            inner_h = inner
        if rest == 1:
            parts.append(inner_h)
            continue
        if not inner.free_symbols:                 # a constant term: one stage constant
            parts.append(cs2.get(rest * inner))
            continue
        # a single monomial keeps its number inside the constant (c*vJ, not 2*c*vJ)
        if len(groups[key]) == 1:
            n1, m1 = groups[key][0].as_coeff_Mul()
            parts.append(cs2.get(n1 * rest) * m1)
        else:
            parts.append(sp.Mul(cs2.get(rest), inner_h, evaluate=False) if isinstance(inner_h, sp.Add)
                         else cs2.get(rest) * inner_h)
    out.append((sp.Add(*parts).xreplace(back), cs2))
    return out


def dag_cost(exprs, prep=None) -> int:
    """FMA-aware instruction count of a set of expressions evaluated together (common sub-expressions once).
    `prep` maps an expression to what a trajectory actually evaluates (the hoisting pass: batch-uniform
    sub-expressions become table symbols and cost nothing)."""
    exprs = [sp.sympify(e) for e in exprs]
    if prep is not None:
        exprs = [prep(e) for e in exprs]
    repl, red = sp.cse(exprs, optimizations="basic")
    return sum(fma_cost(e) for _, e in repl) + sum(fma_cost(e) for e in red)


def fma_forms(exprs, dyn, gens_order, prefix="c", prep=None):
    """Cheapest FMA-shaped form of every expression, rows chosen greedily under the joint cost of everything chosen
    so far (stage constants shared between rows count once).  Returns (stage constants [(symbol, expr)], forms)."""
    shared = _Consts(prefix)
    forms = []
    for e in exprs:
        e = sp.sympify(e)
        best = None
        for form, cs in _candidates(e, dyn, gens_order, lambda: _Consts("tmpc")):
            defs = list(shared.defs)
            trial = _Consts(prefix)
            trial.defs, trial.known = defs, dict(shared.known)
            if isinstance(cs, _Consts):
                form = form.xreplace({s_: trial.get(d) for s_, d in cs.defs})
            cost = dag_cost([d for _, d in trial.defs] + forms + [form], prep)
            if best is None or cost < best[0]:
                best = (cost, form, trial)
        forms.append(best[1])
        shared = best[2]
    return shared.defs, forms


def inline_fused_products(repl, red):
    """Undoes the part of common-sub-expression elimination that costs instructions on an FMA machine: a temporary
    `z = a*b` that is only ever added to something is one multiplication plus one addition per use, while `a*b`
    written in place contracts into the addition.  Returns (repl, red) with such temporaries substituted back."""
    repl, red = list(repl), list(red)
    k = 0
    while k < len(repl):
        s, e = repl[k]
        c, r = e.as_coeff_Mul()
        simple = isinstance(r, sp.Mul) and len(r.args) == 2 and all(a.is_Symbol for a in r.args) and c in (1, -1)
        if simple:
            other = [0]

            def scan(node, parent_is_add):
                if node == s:
                    if not parent_is_add:
                        other[0] += 1
                    return
                if node.is_Atom:
                    return
                if isinstance(node, sp.Mul) and parent_is_add:
                    cc, rr = node.as_coeff_Mul()
                    if rr == s and cc in (1, -1):
                        return
                for a in node.args:
                    scan(a, isinstance(node, sp.Add))
            for _, e2 in repl[k + 1:]:
                scan(e2, False)
            for e2 in red:
                scan(e2, False)
            if other[0] == 0:
                sub = {s: e}
                repl = repl[:k] + [(s2, e2.xreplace(sub)) for s2, e2 in repl[k + 1:]]
                red = [e2.xreplace(sub) for e2 in red]
                continue
        k += 1
    return repl, red
