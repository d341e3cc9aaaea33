from .make_from_symbolic import CodegenProblem, make_from_symbolic
from .problems import harvest_chain_problem, test_problem_symbolic
