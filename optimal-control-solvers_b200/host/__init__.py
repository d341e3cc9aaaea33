from .controls import ChebyshevControl, Control, PWConstantControl, PWLinearControl, matlab_linspace
from .integrators import CudaRK4InfiniteIntegrator, CudaRK4Integrator
from .problems import OCProblem, TestOCProblem
from .drivers import compute_equilibrium, compute_states_and_objective, fb_sweep, nlp_objective, single_shooting
