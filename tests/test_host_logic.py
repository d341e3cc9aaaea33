"""Host-side logic that needs no GPU: MATLAB linspace semantics, the Control mirrors' basis
matrices and helper methods, batch sharding and the gloo (CPU) gather path."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_matlab_linspace_vs_numpy(ocs):
    a = ocs.matlab_linspace(0.0, 10.0, 501)
    assert a[0] == 0.0 and a[-1] == 10.0 and a.size == 501
    # MATLAB's (i*(b-a))/n1 differs from numpy's a + i*((b-a)/n1) in the last bit at many points
    assert np.count_nonzero(a != np.linspace(0.0, 10.0, 501)) > 50
    assert np.array_equal(ocs.matlab_linspace(3.0, 3.0, 1), [3.0])


@pytest.mark.parametrize("T,N,K", [(10, 500, 100), (10, 500, 20), (10, 1000, 100), (10, 1000, 50), (7, 1000, 70),
                                   (1, 10000, 1000)])
def test_pwconstant_index_is_exact_rational(ocs, oracle, T, N, K):
    """With MATLAB's linspace rounding the interval index of grid point j is floor(j*K/(2N))
    (SURVEY.md App. B.1); the host mirror and the oracle agree bit for bit."""
    tspan = ocs.matlab_linspace(0.0, float(T), N + 1)
    t = oracle.RK4Integrator(tspan).t
    ctl = ocs.PWConstantControl(t, K, 1)
    octl = oracle.PWConstantControl(t, K, 1)
    assert np.array_equal(ctl.B, octl.B)
    assert np.all(ctl.B.sum(axis=0) == 1)
    j = np.arange(2 * N + 1)
    assert np.array_equal(np.argmax(ctl.B, axis=0), np.minimum((j * K) // (2 * N), K - 1))


@pytest.mark.parametrize("cls,K", [("PWLinearControl", 21), ("PWLinearControl", 101), ("ChebyshevControl", 17)])
def test_basis_matches_oracle(ocs, oracle, cls, K):
    t = oracle.RK4Integrator(ocs.matlab_linspace(0.0, 10.0, 501)).t
    a, b = getattr(ocs, cls)(t, K, 2), getattr(oracle, cls)(t, K, 2)
    assert np.array_equal(a.B, b.B)
    if cls == "PWLinearControl":
        assert np.allclose(a.B.sum(axis=0), 1.0, atol=1e-15) and np.all((a.B != 0).sum(axis=0) <= 2)
    v = np.random.default_rng(0).random(2 * K)
    assert np.array_equal(a.compute_u(v), b.compute_u(v))
    d = np.random.default_rng(1).random((2, t.size))
    assert np.array_equal(a.compute_dJdv(d), b.compute_dJdv(d))
    Lb, Ub = a._tile_bounds([[0, 1], [-2, 3]])
    assert np.array_equal(Lb, np.tile([0.0, -2.0], K)) and np.array_equal(Ub, np.tile([1.0, 3.0], K))


def test_control_helpers(ocs):
    t = np.linspace(0, 1, 21)
    pwl = ocs.PWLinearControl(t, 5, 1)
    assert np.array_equal(pwl.compute_initial_v(0.3), np.full(5, 0.3))
    v = np.arange(5.0)
    assert np.allclose(pwl.compute_uFunc(v)(pwl.controlPts)[0], v)
    pwc = ocs.PWConstantControl(t, 4, 1)
    assert np.array_equal(pwc.compute_uFunc(np.arange(4.0))([0.0, 0.3, 0.99])[0], [0.0, 1.0, 3.0])
    ch = ocs.ChebyshevControl(t, 4, 1)
    assert np.array_equal(ch.compute_initial_v(0.7), [0.7, 0, 0, 0])
    assert np.allclose(ch.compute_uFunc([1, 2, 3, 4])(t), ch.compute_u(np.array([1.0, 2, 3, 4])))


def test_shard_range():
    import importlib
    dist = importlib.import_module("optimal-control-solvers_b200.host.dist")
    for total, world in [(1 << 20, 8), (1000, 3), (7, 8)]:
        spans = [dist.shard_range(total, r, world) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == total
        assert all(spans[r][1] == spans[r + 1][0] for r in range(world - 1))
        sizes = [b - a for a, b in spans]
        assert max(sizes) - min(sizes) <= 1


def test_gather_over_gloo_world_size_2():
    """The N > 1 plumbing (shard, all-gather of J and x(T)) on two CPU processes over gloo."""
    script = os.path.join(ROOT, "tests", "dist_worker.py")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29611")
    procs = [subprocess.Popen([sys.executable, script, str(r), "2"], env=env, stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT) for r in range(2)]
    outs = [p.communicate(timeout=180)[0].decode() for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert all("gather ok" in o for o in outs), outs


def test_reference_arm_uses_every_host_thread_and_prints_one_line():
    """bench.py --impl reference: ONE JSON line on stdout, same metric/unit as the GPU arm, and the C oracle on every
    host thread even when the launcher exports OMP_NUM_THREADS=1 (torchrun does)."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, OMP_NUM_THREADS="1")
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["unit"] == "trajectory-steps/s" and j["higher_is_better"] is True
    assert j["metric"].startswith("FP64 trajectory-steps/sec") and j["config"]["workload"].startswith("cfg2")
    assert j["cpu_baseline"]["kind"] == "port" and j["cpu_baseline"]["cores"] == len(os.sched_getaffinity(0))
    assert j["e2e"] == {"value": j["value"], "unit": j["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert j["value"] > 0 and j["gpu_launches"] == 0
