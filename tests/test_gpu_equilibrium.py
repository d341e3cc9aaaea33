"""Batched compute_equilibrium (functions/compute_equilibrium.m) on the GPU against the oracle's
Newton solve of the same residual system (oracle/rk4_oracle.py: equilibrium_residual, solve_equilibrium)."""
import importlib

import numpy as np
import pytest

from gen_util import oracle_problem
from util import TEST_PARAMS

pytestmark = pytest.mark.gpu


def test_single_set_reference_shapes(ocs, oracle):
    """tests/solve_test_problem.m:11-18: one steady state, outputs shaped like the reference's."""
    prob = ocs.TestOCProblem(TEST_PARAMS, [[0, 1]])
    oprob = oracle.TestOCProblem(**TEST_PARAMS)
    xs, ls, us, resnorm, residual, flag = ocs.compute_equilibrium(prob, [2.7], [2.2], [.7], r=0.05)
    oxs, ols, ous = oracle.solve_equilibrium(oprob, [2.7], [2.2], [.7], 0.05)
    assert flag == 1 and resnorm < 1e-24 and residual.shape == (3,)
    assert abs(xs[0] - oxs[0]) < 1e-12 and abs(ls[0] - ols[0]) < 1e-12 and abs(us[0] - ous[0]) < 1e-12
    assert abs(us[0] - 0.723368780097983) < 1e-12            # SURVEY.md App. C anchor (cfg4's uStar)
    y = np.concatenate([xs, ls, us])
    assert np.max(np.abs(oracle.equilibrium_residual(oprob, y, 0.05) - residual)) < 1e-14


def test_parameter_sweep_and_bounds(ocs, oracle):
    rng = np.random.default_rng(21)
    ns = 4097                                                 # ragged: not a multiple of the block size
    c, m, r = rng.uniform(1.3, 1.7, ns), rng.uniform(2.8, 3.2, ns), rng.uniform(0.03, 0.07, ns)
    prob = ocs.TestOCProblem(dict(c=c, m=m, r=r), [[0, 1]])
    xs, ls, us, resnorm, residual, flag = ocs.compute_equilibrium(prob, [2.7], [2.2], [.7], r=r)
    assert xs.shape == (1, ns) and flag.shape == (ns,)
    assert (flag == 1).all() and resnorm.max() < 1e-22
    for b in (0, 1, 1000, 4096):
        op = oracle.TestOCProblem(c[b], m[b], r[b])
        oxs, ols, ous = oracle.solve_equilibrium(op, [2.7], [2.2], [.7], r[b])
        assert abs(xs[0, b] - oxs[0]) < 1e-11 and abs(ls[0, b] - ols[0]) < 1e-11 and abs(us[0, b] - ous[0]) < 1e-11
    # bounds (lsqnonlin's lb/ub, compute_equilibrium.m:27): with u capped below the interior root the solution
    # sits on the bound and the residual no longer vanishes
    lb, ub = [0.0, -10.0, 0.0], [10.0, 10.0, 0.5]
    xb, lbm, ub_, rn, res, fl = ocs.compute_equilibrium(prob, [2.7], [2.2], [.4], lb=lb, ub=ub, r=r)
    assert np.all(ub_ == 0.5) and (fl == 1).all() and rn.min() > 1e-6
    with pytest.raises(ValueError):
        ocs.compute_equilibrium(prob, [2.7], [2.2], [.7], lb=[0.0], r=r)


def test_generated_six_state_problem(ocs, oracle):
    """The generated cfg3 problem (NVRTC-compiled equilibrium kernel, 14 unknowns) against the oracle."""
    cg = importlib.import_module("optimal-control-solvers_b200.codegen")
    prob = cg.harvest_chain_problem()
    oprob = oracle_problem(oracle, prob)
    oxs, ols, ous = oracle.solve_equilibrium(oprob, [2.0] * 6, [0.5] * 6, [0.3, 0.3], 0.05)
    xs, ls, us, resnorm, residual, flag = ocs.compute_equilibrium(prob, [2.0] * 6, [0.5] * 6, [0.3, 0.3], r=0.05)
    assert flag == 1 and resnorm < 1e-22
    assert np.max(np.abs(xs - oxs)) < 1e-10 and np.max(np.abs(ls - ols)) < 1e-10 and np.max(np.abs(us - ous)) < 1e-10
    # a sweep over the harvest cost through `params`
    P = np.repeat(np.asarray(prob.param_array(), dtype=float).reshape(-1, 1), 64, axis=1)
    P[3] = np.linspace(1.0, 2.0, 64)                         # params: m, q, d, c, r, xs
    out = ocs.compute_equilibrium(prob, [2.0] * 6, [0.5] * 6, [0.3, 0.3], r=0.05, params=P)
    assert (out[5] == 1).all() and out[3].max() < 1e-20
    assert np.all(np.diff(out[2][0]) < 0)                    # dearer effort -> less harvesting
