"""Two lanes (ocs_api.cu, lanes_*): a batch swept as two halves on two internal streams returns exactly what the one-pass
form returns -- the halves are independent trajectories run by the same kernels -- for the fused evaluation on host
buffers, on resident data, on device buffers, and for the fb_sweep loop with its one-sweep-late convergence check.
OCS_B200_LANES is read when the handle is created: 1 forces lanes whenever the batch can be halved, 0 forbids them."""
import importlib
import os

import numpy as np
import pytest

from util import TEST_PARAMS, relerr

pytestmark = pytest.mark.gpu


def _integ(ocs, lanes, tspan, **kw):
    old = os.environ.get("OCS_B200_LANES")
    os.environ["OCS_B200_LANES"] = "1" if lanes else "0"
    try:
        return ocs.CudaRK4Integrator(tspan, **kw)
    finally:
        if old is None:
            del os.environ["OCS_B200_LANES"]
        else:
            os.environ["OCS_B200_LANES"] = old


@pytest.mark.parametrize("Bt", [3072, 3100, 5000])
def test_generated_problem_evaluation_is_identical_with_and_without_lanes(ocs, Bt):
    cg = importlib.import_module("optimal-control-solvers_b200.codegen")
    prob = cg.harvest_chain_problem()
    tspan = ocs.matlab_linspace(0.0, 2.0, 41)
    rng = np.random.default_rng(Bt)
    x0 = rng.uniform(1.0, 2.5, (6, Bt))
    v = 0.05 * rng.standard_normal((10, Bt))
    v[:2] = rng.uniform(0.2, 0.8, (2, Bt))
    res = {}
    for lanes in (0, 1):
        integ = _integ(ocs, lanes, tspan)
        integ.attach_control(ocs.ChebyshevControl(integ.t, 5, 2))
        res[lanes] = integ.evaluate(prob, x0, v)                       # host buffers: the small-batch one-pass path
        integ.evaluate_resident(prob, x0, v)                           # resident data
        res[lanes, "r"] = integ.results()
        integ.close()
    for key in ("J", "xT", "status", "dJdv", "lam0"):
        assert np.array_equal(res[0][key], res[1][key]), key
        assert np.array_equal(res[0][key], res[1, "r"][key]), key
        assert np.array_equal(res[0][key], res[0, "r"][key]), key
    assert np.all(np.isfinite(res[1]["dJdv"])) and np.abs(res[1]["dJdv"]).max() > 0


def test_builtin_problem_on_device_buffers_with_lanes(ocs):
    """ocs_b200_evaluate_dev: the fused call on device buffers; several trajectories per thread in the built-in kernels, so
    the FMA-contracted build is compared to rounding and the strict build bit for bit."""
    import torch
    Bt, K = 6200, 11
    tspan = ocs.matlab_linspace(0.0, 5.0, 101)
    prob = ocs.TestOCProblem(TEST_PARAMS, [[0, 1]])
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev)
    g.manual_seed(3)
    x0 = (0.5 + 2.0 * torch.rand((1, Bt), dtype=torch.float64, device=dev, generator=g)).contiguous()
    vd = torch.rand((K, Bt), dtype=torch.float64, device=dev, generator=g).contiguous()
    for math, same in (("strict", np.array_equal), ("fast", lambda p, q: relerr(p, q) < 1e-13)):
        out = {}
        for lanes in (0, 1):
            integ = _integ(ocs, lanes, tspan, math=math)
            integ.attach_control(ocs.PWLinearControl(integ.t, K, 1))
            J, xT = torch.empty(Bt, dtype=torch.float64, device=dev), torch.empty((1, Bt), dtype=torch.float64, device=dev)
            st = torch.empty(Bt, dtype=torch.int32, device=dev)
            dJdv, lam0 = torch.empty((K, Bt), dtype=torch.float64, device=dev), torch.empty((2, Bt), dtype=torch.float64, device=dev)
            for _ in range(2):      # twice: the second evaluation forks from a stream that joined the first one's lanes
                integ.evaluate_dev(prob, x0.data_ptr(), Bt, v_ptr=vd.data_ptr(), J_ptr=J.data_ptr(), xT_ptr=xT.data_ptr(),
                                   status_ptr=st.data_ptr(), dJdv_ptr=dJdv.data_ptr(), lam0_ptr=lam0.data_ptr())
            integ.synchronize()
            out[lanes] = dict(J=J.cpu().numpy(), xT=xT.cpu().numpy(), dJdv=dJdv.cpu().numpy(), lam0=lam0.cpu().numpy())
            integ.close()
        for key in out[0]:
            assert same(out[0][key], out[1][key]), (math, key)


@pytest.mark.parametrize("costate", ["discrete", "continuous"])
def test_fb_sweep_lanes_and_the_late_convergence_check(ocs, costate):
    """Each half runs its own chain of sweeps; the host reads the open-trajectory counts one sweep late, so up to one
    extra sweep is queued after everything has converged -- it must change nothing (every trajectory is frozen)."""
    cg = importlib.import_module("optimal-control-solvers_b200.codegen")
    prob = cg.harvest_chain_problem()
    Bt, N = 3300, 30
    tspan = ocs.matlab_linspace(0.0, 1.5, N + 1)
    rng = np.random.default_rng(11)
    x0 = rng.uniform(1.0, 2.5, (6, Bt))
    u0 = np.full((2, 2 * N + 1), 0.4)
    for limit in (30, 3):          # everything converges well before 30 sweeps; 3 sweeps hit the limit
        res = {}
        for lanes in (0, 1):
            integ = _integ(ocs, lanes, tspan)
            res[lanes] = integ.fb_sweep(prob, x0, u0, uRelTol=1e-3, uAbsTol=1e-3, nSWEEPS=limit, relax=0.5, costate=costate)
            res[lanes]["u"] = integ.get_control_u(0, Bt)
            res[lanes]["x"] = integ.fetch_states(0, Bt)
            integ.close()
        if limit == 30:
            assert res[0]["converged"].all() and res[0]["sweeps"].max() < 30
        else:
            assert not res[0]["converged"].any() and (res[0]["sweeps"] == 3).all()
        for key in ("J", "sweeps", "converged", "maxChange", "u", "x"):
            # (x after fb_sweep is read from the node store: its running-objective row is NaN except x(end,end) = J)
            assert np.array_equal(res[0][key], res[1][key], equal_nan=(key == "x")), (limit, key)
        assert np.isfinite(res[1]["x"][:6]).all() and np.array_equal(res[1]["x"][6, -1], res[1]["J"])


def test_lanes_decision_is_automatic_for_low_occupancy_sweeps(ocs):
    """Default mode: the generated 6-state problem (8-12 resident warps per SM) at a batch of a few rounds uses lanes, the
    1-state built-in problem (4+ trajectories per thread, 16+ resident warps) does not; both agree with the forced forms."""
    cg = importlib.import_module("optimal-control-solvers_b200.codegen")
    prob = cg.harvest_chain_problem()
    tspan = ocs.matlab_linspace(0.0, 1.0, 21)
    Bt = 40_000
    rng = np.random.default_rng(2)
    x0 = rng.uniform(1.0, 2.5, (6, Bt))
    v = rng.uniform(0.2, 0.8, (6, Bt))
    out = {}
    for mode in ("auto", 0):
        integ = ocs.CudaRK4Integrator(tspan) if mode == "auto" else _integ(ocs, 0, tspan)
        integ.attach_control(ocs.ChebyshevControl(integ.t, 3, 2))
        out[mode] = integ.evaluate(prob, x0, v)
        integ.close()
    for key in ("J", "dJdv"):
        assert np.array_equal(out["auto"][key], out[0][key]), key


@pytest.mark.parametrize("which", ["builtin", "generated"])
@pytest.mark.parametrize("lanes", [0, 1])
def test_device_outputs_stay_inside_their_buffers(ocs, which, lanes):
    """compute-sanitizer is closed on this GPU pool, so out-of-bounds writes are looked for directly: every output of the
    fused device-buffer evaluation lives inside a larger NaN-filled allocation (one guard row before and after, and a
    ragged batch that ends in the middle of a block, a warp and a 16-byte pair); the guards must come back untouched and
    every element inside must have been written."""
    import torch
    dev = torch.device("cuda", 0)
    if which == "builtin":
        prob, n, K, nc = ocs.TestOCProblem(TEST_PARAMS, [[0, 1]]), 1, 11, 1
        ctl_of = lambda integ: ocs.PWLinearControl(integ.t, K, nc)
    else:
        cg = importlib.import_module("optimal-control-solvers_b200.codegen")
        prob, n, K, nc = cg.harvest_chain_problem(), 6, 5, 2
        ctl_of = lambda integ: ocs.ChebyshevControl(integ.t, K, nc)
    Bt = 3 * 1536 + 77                       # two lanes of whole blocks plus a ragged tail
    tspan = ocs.matlab_linspace(0.0, 2.0, 34)      # 33 steps: not a multiple of any tile size
    integ = _integ(ocs, lanes, tspan)
    integ.attach_control(ctl_of(integ))
    g = torch.Generator(device=dev)
    g.manual_seed(5)
    x0 = (1.0 + 1.4 * torch.rand((n, Bt), dtype=torch.float64, device=dev, generator=g)).contiguous()
    v = (0.2 + 0.5 * torch.rand((nc * K, Bt), dtype=torch.float64, device=dev, generator=g)).contiguous()
    if which == "generated":
        v[2 * nc:] *= 0.05

    def guarded(rows, dtype=torch.float64):
        fill = float("nan") if dtype == torch.float64 else -7
        return torch.full((rows + 2, Bt), fill, dtype=dtype, device=dev)
    J, xT, st, dJdv, lam0 = guarded(1), guarded(n), guarded(1, torch.int32), guarded(nc * K), guarded(n + 1)
    row = lambda t: t[1:].data_ptr()
    for _ in range(2):
        integ.evaluate_dev(prob, x0.data_ptr(), Bt, v_ptr=v.data_ptr(), J_ptr=row(J), xT_ptr=row(xT), status_ptr=row(st),
                           dJdv_ptr=row(dJdv), lam0_ptr=row(lam0))
    integ.synchronize()
    for name, t in (("J", J), ("xT", xT), ("dJdv", dJdv), ("lam0", lam0)):
        assert torch.isnan(t[0]).all() and torch.isnan(t[-1]).all(), name + ": a guard row was written"
        assert torch.isfinite(t[1:-1]).all(), name + ": an element inside was not written"
    assert (st[0] == -7).all() and (st[-1] == -7).all() and (st[1] == 0).all()
    integ.close()


@pytest.mark.parametrize("K,N,Bt", [(17, 70, 777), (5, 33, 4000), (32, 16, 130)])
def test_dense_basis_evaluation_kernel_matches_the_generic_one(ocs, K, N, Bt):
    """fb_sweep started from coefficients evaluates u0 = v*B on the device.  The dense-basis kernel (coefficients in
    registers, B tiles in shared memory, the grid cut into segments) must reproduce the generic kernel bit for bit -- same
    summation order -- for ragged batches, grids that are no multiple of the tile and K up to its register limit; with
    relax = 0 the sweep leaves the control alone, so the grid control that comes back IS u0."""
    cg = importlib.import_module("optimal-control-solvers_b200.codegen")
    prob = cg.harvest_chain_problem()
    tspan = ocs.matlab_linspace(0.0, 1.0, N + 1)
    rng = np.random.default_rng(K)
    x0 = rng.uniform(1.0, 2.5, (6, Bt))
    v0 = 0.02 * rng.standard_normal((2 * K, Bt))
    v0[:2] = rng.uniform(0.2, 0.8, (2, Bt))
    got = {}
    for generic in (False, True):
        if generic:
            os.environ["OCS_B200_BASIS_EVAL_GENERIC"] = "1"
        try:
            integ = ocs.CudaRK4Integrator(tspan)
            ctl = ocs.ChebyshevControl(integ.t, K, 2)
            integ.attach_control(ctl)
            out = integ.fb_sweep(prob, x0, None, v0=v0, nSWEEPS=1, relax=0.0)
            got[generic] = (integ.get_control_u(0, Bt), out["J"])
            integ.close()
        finally:
            os.environ.pop("OCS_B200_BASIS_EVAL_GENERIC", None)
    assert np.array_equal(got[False][0], got[True][0]) and np.array_equal(got[False][1], got[True][1])
    ref = np.einsum("ckb,kj->cjb", v0.reshape(K, 2, Bt).transpose(1, 0, 2), ctl.B)      # rows of v: k*nc + c
    assert relerr(got[False][0], ref) < 1e-14
