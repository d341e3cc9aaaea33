"""Several GPUs behind ONE handle (ocs_b200_create_multi): every host-buffer entry point must return exactly what the
single-device handle returns -- shards are independent trajectories swept by the same kernels -- for ragged shard
sizes, per-trajectory controls / parameters, ranges that straddle a shard boundary, and the device-resident path with
its gather to the root device.  Runs with however many GPUs the box shows (1: the fan-out code with one shard;
the 2-GPU call of the round exercises real shards)."""
import numpy as np
import pytest

from util import TEST_PARAMS, TOL, relerr

pytestmark = pytest.mark.gpu


def _ndev():
    import torch
    return torch.cuda.device_count()


def _pair(ocs, N=200, T=5.0, K=11, kind="PWLinearControl", devices=None, math="fast"):
    tspan = ocs.matlab_linspace(0.0, T, N + 1)
    one = ocs.CudaRK4Integrator(tspan, device=0, math=math)
    many = ocs.CudaRK4Integrator(tspan, device=devices if devices is not None else list(range(_ndev())), math=math)
    for integ in (one, many):
        integ.attach_control(getattr(ocs, kind)(integ.t, K, 1))
    return one, many


@pytest.mark.parametrize("Bt", [1, 3, 1000, 4097])
def test_evaluate_matches_single_device(ocs, Bt):
    one, many = _pair(ocs)
    prob = ocs.TestOCProblem(TEST_PARAMS, [[0, 1]])
    rng = np.random.default_rng(Bt)
    x0, v = rng.uniform(.5, 2.5, (1, Bt)), rng.random(11)
    a, b = one.evaluate(prob, x0, v), many.evaluate(prob, x0, v)
    for key in ("J", "xT", "status", "dJdv", "lam0"):
        assert np.array_equal(a[key], b[key]), key
    devs, off = many.shard_offsets()
    assert off[0] == 0 and off[-1] == Bt and len(devs) == _ndev()
    assert all(0 <= off[k + 1] - off[k] - Bt // len(devs) <= 1 for k in range(len(devs)))
    # the reductions combine the shards with the single-device tie-breaking
    for op in ("min", "max"):
        assert one.reduce_J(op) == many.reduce_J(op)
    assert relerr(many.reduce_J("sum")[0], one.reduce_J("sum")[0]) < 1e-14
    one.close()
    many.close()


@pytest.mark.parametrize("math", ["strict", "fast"])
def test_per_trajectory_inputs_are_sharded_like_the_batch(ocs, math):
    """Per-trajectory coefficients and parameters are cut like the batch.  With several trajectories per thread the
    FMA-contracted build may fuse the same source expression differently in the thread's trajectory slots, so in
    `fast` mode a trajectory's last bits can depend on its position in the (shard of the) batch; `strict`
    arithmetic (no contraction) is position-independent, hence bit-identical across any sharding."""
    Bt = 1501
    one, many = _pair(ocs, math=math)
    same = np.array_equal if math == "strict" else (lambda p, q: relerr(p, q) < 1e-13)
    rng = np.random.default_rng(7)
    prob = ocs.TestOCProblem(dict(c=rng.uniform(1.3, 1.7, Bt), m=rng.uniform(2.8, 3.2, Bt), r=rng.uniform(.03, .07, Bt)), [[0, 1]])
    x0, v = rng.uniform(.5, 2.5, (1, Bt)), rng.random((11, Bt))
    a, b = one.evaluate(prob, x0, v), many.evaluate(prob, x0, v)
    for key in ("J", "xT", "dJdv", "lam0"):
        assert same(a[key], b[key]), key
    # grids of a range that straddles every shard boundary
    lo, hi = 10, Bt - 10
    assert same(one.fetch_states(lo, hi - lo), many.fetch_states(lo, hi - lo))
    la, da = one.fetch_adjoints(lo, hi - lo)
    lb, db = many.fetch_adjoints(lo, hi - lo)
    assert same(la, lb) and same(da, db)
    assert same(one.get_control_u(lo, 40), many.get_control_u(lo, 40))
    tq = np.linspace(0.0, 5.0, 33)
    sa, sb = one.eval_solution(tq, lo, hi - lo), many.eval_solution(tq, lo, hi - lo)
    for key in ("x", "lam", "u"):
        assert same(sa[key], sb[key]), key
    one.close()
    many.close()


def test_two_call_interface_and_grid_control(ocs):
    Bt = 777
    one, many = _pair(ocs)
    for integ in (one, many):
        integ.attach_control(None)          # u on the grid: compute_u / compute_dJdv stay with the caller
    prob = ocs.TestOCProblem(TEST_PARAMS, [[0, 1]])
    rng = np.random.default_rng(3)
    x0 = rng.uniform(.5, 2.5, (1, Bt))
    u = rng.uniform(0, 1, (1, one.t.size, Bt))
    lamT = np.vstack([rng.normal(size=(1, Bt)), np.ones((1, Bt))])
    res = []
    for integ in (one, many):
        _, J = integ.compute_states(prob, x0, u)
        lam0, _ = integ.compute_adjoints(prob, u, lamT=lamT, grids=False)
        lam, dJdu = integ.fetch_adjoints(300, 200)
        res.append((J, lam0, lam, dJdu))
    for a, b in zip(*res):
        assert np.array_equal(a, b)
    one.close()
    many.close()


def test_solvers_and_equilibrium(ocs):
    Bt = 515
    one, many = _pair(ocs, N=100, K=9, math="strict")
    prob = ocs.TestOCProblem(TEST_PARAMS, [[0, 1]])
    rng = np.random.default_rng(11)
    x0 = rng.uniform(.5, 2.5, (1, Bt))
    u0 = np.full((1, one.t.size), .5)
    fa, fb = one.fb_sweep(prob, x0, u0, nSWEEPS=40), many.fb_sweep(prob, x0, u0, nSWEEPS=40)
    for key in ("J", "sweeps", "converged", "maxChange"):
        assert np.array_equal(fa[key], fb[key]), key
    assert np.array_equal(one.get_control_u(250, 20), many.get_control_u(250, 20))
    one2, many2 = _pair(ocs, N=100, K=9, math="strict")
    ctl = one2._control
    Lb, Ub = ctl.compute_nlp_bounds(prob.ControlBounds)
    v0 = rng.uniform(.2, .8, (9, Bt))
    sa = one2.single_shooting(prob, x0, v0, Lb, Ub, FreeInitStates=[0], FreeStateBounds=[[.5, 2.5]], max_iter=60)
    sb = many2.single_shooting(prob, x0, v0, Lb, Ub, FreeInitStates=[0], FreeStateBounds=[[.5, 2.5]], max_iter=60)
    for key in ("v", "J", "iters", "converged"):
        assert np.array_equal(sa[key], sb[key]), key
    ns = 301
    sweep = ocs.TestOCProblem(dict(c=rng.uniform(1.3, 1.7, ns), m=3.0, r=rng.uniform(.03, .07, ns)), [[0, 1]])
    ea = ocs.compute_equilibrium(sweep, [2.7], [2.2], [.7], r=np.asarray(sweep.r), integrator=one2)
    eb = ocs.compute_equilibrium(sweep, [2.7], [2.2], [.7], r=np.asarray(sweep.r), integrator=many2)
    for a, b in zip(ea, eb):
        assert np.array_equal(np.asarray(a), np.asarray(b))
    for integ in (one, many, one2, many2):
        integ.close()


@pytest.mark.parametrize("transport", ["nccl", "p2p"])
def test_resident_evaluation_and_gather_to_root(ocs, transport, monkeypatch):
    """ocs_b200_evaluate_resident keeps everything on the devices; ocs_b200_gather_results_dev brings J and x(T)
    to the root device (NCCL send/recv, or peer copies) and ocs_b200_get_results to the host."""
    import torch
    monkeypatch.setenv("OCS_B200_GATHER", transport)
    Bt = 2003
    one, many = _pair(ocs)
    prob = ocs.TestOCProblem(TEST_PARAMS, [[0, 1]])
    rng = np.random.default_rng(5)
    x0 = rng.uniform(.5, 2.5, (1, Bt))
    v1, v2 = rng.random(11), rng.random(11)
    ref, ref2 = one.evaluate(prob, x0, v1), one.evaluate(prob, x0, v2)
    many.evaluate_resident(prob, x0, v1, gradient=True)
    Jd = torch.full((Bt,), float("nan"), dtype=torch.float64, device="cuda:0")
    xTd = torch.full((1, Bt), float("nan"), dtype=torch.float64, device="cuda:0")
    many.gather_results_dev(Jd.data_ptr(), xTd.data_ptr())
    many.synchronize()
    torch.cuda.synchronize()
    assert np.array_equal(Jd.cpu().numpy(), ref["J"]) and np.array_equal(xTd.cpu().numpy(), ref["xT"])
    got = many.results()
    for key in ("J", "xT", "status", "dJdv", "lam0"):
        assert np.array_equal(got[key], ref[key]), key
    # a new control candidate on the resident initial states: nothing but v crosses PCIe
    many.evaluate_resident(prob, None, v2, gradient=True)
    many.gather_results_dev(Jd.data_ptr(), xTd.data_ptr())
    many.synchronize()
    torch.cuda.synchronize()
    assert np.array_equal(Jd.cpu().numpy(), ref2["J"])
    assert np.array_equal(many.results()["dJdv"], ref2["dJdv"])
    one.close()
    many.close()


def test_device_pointer_entry_points_are_refused_on_a_multi_handle(ocs):
    one, many = _pair(ocs)
    with pytest.raises(ocs.OcsError) as e:
        many.set_stream(0)
    assert e.value.code == 1 and "multi-device" in str(e.value)
    one.close()
    many.close()


def test_duplicate_devices_are_rejected(ocs):
    with pytest.raises(ocs.OcsError):
        ocs.CudaRK4Integrator(ocs.matlab_linspace(0, 1, 11), device=[0, 0])
