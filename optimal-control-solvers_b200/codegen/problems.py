"""Symbolic problem definitions that go through the make_from_symbolic CUDA backend."""
from __future__ import annotations

from .make_from_symbolic import make_from_symbolic


def test_problem_symbolic(c=1.5, m=3.0, r=0.05, ControlBounds=((0.0, 1.0),)):
    """tests/TestOCProblem.m expressed symbolically (the same model tests/symbolic_test2.m feeds to
    make_from_symbolic, with the discount factor of TestOCProblem.m:25):
        x' = x (m - x) - u,     integrand exp(-r t) (x^2 + c u^2)."""
    return make_from_symbolic("exp(-r*t)*(x1**2 + c*u1**2)", ["x1*(m - x1) - u1"], 1, 1,
                              dict(c=c, m=m, r=r), ControlBounds)


def harvest_chain_problem(m=3.0, q=0.6, d=0.15, c=1.5, r=0.05, xs=1.8,
                          ControlBounds=((0.0, 1.0), (0.0, 1.0))):
    """The 6-state / 6-costate problem of BASELINE.json configs[2] (the reference names the
    configuration but ships no such problem; this is the build's
    definition).  Six logistic stocks on a ring coupled by diffusion -- the bundled test problem's
    dynamics x(m - x) - u (tests/TestOCProblem.m:24) per patch -- harvested by two efforts:

        x_i' = x_i (m - x_i) - q u_{c(i)} x_i + d (x_{i-1} - 2 x_i + x_{i+1}),   i = 1..6 (ring)
        c(i) = 1 for i <= 3, 2 otherwise
        integrand  exp(-r t) ( sum_i (x_i - xs)^2 + c (u_1^2 + u_2^2) )

    params (in this order): m, q, d, c, r, xs."""
    n = 6
    g = []
    for i in range(1, n + 1):
        im, ip = (i - 2) % n + 1, i % n + 1
        uc = "u1" if i <= 3 else "u2"
        g.append(f"x{i}*(m - x{i}) - q*{uc}*x{i} + d*(x{im} - 2*x{i} + x{ip})")
    f = "exp(-r*t)*(" + " + ".join(f"(x{i} - xs)**2" for i in range(1, n + 1)) + " + c*(u1**2 + u2**2))"
    return make_from_symbolic(f, g, n, 2, dict(m=m, q=q, d=d, c=c, r=r, xs=xs), ControlBounds)
