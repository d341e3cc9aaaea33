"""Host mirror of the OCProblem plugin (OCProblem/OCProblem.m:3-21).  A problem object
names the device code of its right-hand sides -- a built-in id or CUDA source printed by
the make_from_symbolic backend -- and carries parameters and ControlBounds.  The
right-hand sides themselves only exist as device functions (csrc/problem_test.cuh or
generated code); there is no host evaluation path."""
from __future__ import annotations

import numpy as np

from .. import _lib

__all__ = ["OCProblem", "TestOCProblem"]


class OCProblem:
    """Base: subclasses set nStates, nControls, ControlBounds and implement bind()."""

    nStates = 0
    nControls = 0
    ControlBounds = None
    MinMax = "Min"        # functions/single_shooting.m:11-15

    def param_array(self):
        """(np,) shared or (np, Bt) per-trajectory parameter array."""
        raise NotImplementedError

    def bind(self, lib, handle):
        raise NotImplementedError


class TestOCProblem(OCProblem):
    """tests/TestOCProblem.m:13-19: ``TestOCProblem(p, ControlBounds)`` with p.c, p.m, p.r.
    c, m, r may be scalars (shared) or arrays of length Bt (a parameter sweep)."""
    __test__ = False

    nStates = 1
    nControls = 1

    def __init__(self, p, ControlBounds):
        get = (lambda k: p[k]) if isinstance(p, dict) else (lambda k: getattr(p, k))
        self.c, self.m, self.r = get("c"), get("m"), get("r")
        self.ControlBounds = np.asarray(ControlBounds, dtype=np.float64).reshape(-1, 2)

    def param_array(self):
        vals = [np.atleast_1d(np.asarray(x, dtype=np.float64)) for x in (self.c, self.m, self.r)]
        nb = max(v.size for v in vals)
        if nb == 1:
            return np.array([v[0] for v in vals])
        return np.stack([np.broadcast_to(v, (nb,)) for v in vals])   # (3, Bt)

    def bind(self, lib, handle):
        _lib.check(lib.ocs_b200_set_problem_builtin(handle, _lib.PROBLEM_TEST), handle)
