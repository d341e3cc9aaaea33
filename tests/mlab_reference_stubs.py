"""Stand-ins for the reference classes the MATLAB glue touches, for the GPU box (where /root/reference does
not exist and reference sources must not be copied into this repository) -- TEST INFRASTRUCTURE.

  Integrator, OCProblem, Control   the abstract interfaces: empty handle classes
  TestOCProblem(p, ControlBounds)  a plain holder of c, m, r and ControlBounds (the glue only reads those)
  PWLinearControl(t, K, nc)        backed by the product's host mirror (optimal-control-solvers_b200/host/controls.py)
"""
import numpy as np

from oracle.mlab.interp import Parser, tokenize

_HOLDER = """
classdef TestOCProblem < OCProblem
   properties
      ControlBounds
      c
      m
      r
   end
   methods
      function obj = TestOCProblem(p, bounds)
         obj.c = p.c;  obj.m = p.m;  obj.r = p.r;
         obj.ControlBounds = bounds;
      end
   end
end
"""


class _ControlObject:
    """A host-mirror Control exposed to interpreted MATLAB code (fields + methods)."""
    METHODS = ("compute_u", "compute_dJdv", "compute_nlp_bounds", "compute_initial_v")

    def __init__(self, ctl):
        self.ctl = ctl

    def getfield(self, name):
        v = getattr(self.ctl, name)
        return np.atleast_2d(np.asarray(v, dtype=np.float64)) if not np.isscalar(v) else np.array([[float(v)]])

    def call_method(self, name, args, nargout):
        c = self.ctl
        if name == "compute_u":
            return [np.atleast_2d(c.compute_u(np.asarray(args[0]).reshape(-1, order="F")))]
        if name == "compute_dJdv":
            return [c.compute_dJdv(np.asarray(args[0])).reshape(-1, 1)]
        if name == "compute_nlp_bounds":
            Lb, Ub = c.compute_nlp_bounds(np.asarray(args[0]))
            return [Lb.reshape(-1, 1), Ub.reshape(-1, 1)][:max(nargout, 1)]
        if name == "compute_initial_v":
            return [np.asarray(c.compute_initial_v(np.asarray(args[0]).ravel())).reshape(-1, 1)]
        raise ValueError(name)


def _wrap_isa(it):
    """isa(control, 'PWLinearControl') for the host-mirror stand-ins."""
    base = it.builtins["isa"]

    def isa(a, cls):
        if isinstance(a, _ControlObject):
            return np.array([[float(type(a.ctl).__name__ == cls or cls == "Control")]])
        return base(a, cls)
    it.builtins["isa"] = isa


def install(it, ocs):
    _wrap_isa(it)
    for name in ("Integrator", "OCProblem", "Control"):
        it.cache[name] = ("classdef", dict(name=name, supers=["handle"], props=[], methods={}), None)
    kind, node = Parser(tokenize(_HOLDER), "<TestOCProblem holder>").parse_file()
    it.cache["TestOCProblem"] = (kind, node, None)
    it.builtins["PWLinearControl"] = lambda t, K, nc: _ControlObject(
        ocs.PWLinearControl(np.asarray(t).ravel(), int(np.asarray(K).ravel()[0]), int(np.asarray(nc).ravel()[0])))
    it.builtins["PWConstantControl"] = lambda t, K, nc: _ControlObject(
        ocs.PWConstantControl(np.asarray(t).ravel(), int(np.asarray(K).ravel()[0]), int(np.asarray(nc).ravel()[0])))
