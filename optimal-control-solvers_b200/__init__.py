"""B200-native RK4 trajectory sweeps behind the OCProblem / Control / Integrator plugin
API of DrDanRyan/Optimal-Control-Solvers.  Everything that computes goes through
libocs_b200.so (hand-written sm_100a CUDA, C ABI in include/ocs_b200.h)."""
from . import _lib
from ._lib import OcsError
from .host import (ChebyshevControl, Control, CudaRK4InfiniteIntegrator, CudaRK4Integrator, OCProblem,
                   PWConstantControl, PWLinearControl, TestOCProblem, compute_equilibrium, compute_states_and_objective, fb_sweep,
                   matlab_linspace, nlp_objective, single_shooting)

from .codegen import CodegenProblem, harvest_chain_problem, make_from_symbolic, test_problem_symbolic  # noqa: E402

__all__ = ["compute_equilibrium", "single_shooting", "nlp_objective", "compute_states_and_objective", "fb_sweep", "make_from_symbolic", "CodegenProblem", "test_problem_symbolic", "harvest_chain_problem","OcsError", "CudaRK4Integrator", "CudaRK4InfiniteIntegrator", "Control", "PWConstantControl",
           "PWLinearControl", "ChebyshevControl", "OCProblem", "TestOCProblem", "matlab_linspace"]
