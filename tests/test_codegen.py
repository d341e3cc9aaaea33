"""make_from_symbolic CUDA backend: derivations and emitted source (CPU) and NVRTC-compiled
sweeps against the oracle (GPU)."""
import importlib
import os
import shutil
import subprocess
import tempfile

import numpy as np
import pytest
import sympy as sp

from gen_util import oracle_problem
from util import TEST_PARAMS, TOL, relerr

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def cg(ocs):
    return importlib.import_module("optimal-control-solvers_b200.codegen")


def test_derivations_reproduce_TestOCProblem(cg):
    """SymPy derives exactly the hand-written products of tests/TestOCProblem.m:29-38 and
    the control characterisation lam exp(rt)/(2c)."""
    p = cg.test_problem_symbolic()
    s = p.sym
    x, u, v1, vJ = s["x"][0], s["u"][0], s["vx"][0], s["vJ"]
    c, m, r = s["params"]
    t = s["t"]
    assert sp.simplify(s["F"][0] - (x * (m - x) - u)) == 0
    assert sp.simplify(s["F"][1] - sp.exp(-r * t) * (x ** 2 + c * u ** 2)) == 0
    assert sp.simplify(s["X"][0] - ((m - 2 * x) * v1 + 2 * sp.exp(-r * t) * x * vJ)) == 0
    assert sp.simplify(s["U"][0] - (-v1 + 2 * c * sp.exp(-r * t) * u * vJ)) == 0
    assert sp.simplify(s["CC"][0] - s["lam"][0] * sp.exp(r * t) / (2 * c)) == 0
    assert p.flops == dict(C_F=10, C_X=8, C_U=7)
    assert "struct Problem" in p.source and "static constexpr int n = 1, nc = 1, np = 3;" in p.source
    assert p.param_names == ["c", "m", "r"] and np.array_equal(p.param_array(), [1.5, 3.0, 0.05])


def test_six_state_problem_source_compiles_for_sm100a(cg):
    """The generated source is valid CUDA for sm_100a together with the sweep kernels (nvcc here;
    NVRTC does the same at run time on the GPU box)."""
    p = cg.harvest_chain_problem()
    assert p.nStates == 6 and p.nControls == 2 and len(p.param_names) == 6
    assert p.sym["CC"] is not None
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "gen.cu")
        with open(src, "w") as f:
            f.write('#define OCS_NS ocs_gen\n#include "rk4_kernels.cuh"\n' + p.source +
                    "\nOCS_TT_KERNEL(ocs_gen::Problem, gen)\nOCS_FWD_KERNEL(ocs_gen::Problem, gen, 0, 0, 0, 0)\n"
                    "OCS_REV_KERNEL(ocs_gen::Problem, gen, 0, 1, 0, 0)\n")
        r = subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-std=c++17", "-O1",
                            "-I", os.path.join(ROOT, "optimal-control-solvers_b200", "csrc"), "-cubin", "-o",
                            os.path.join(d, "gen.cubin"), src], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr


# the options of csrc/nvrtc_problem.cu (ocs_compile_unit); JIT mode is stricter than nvcc: e.g. generic lambdas in
# the sweep bodies need -default-device, which the nvcc test above cannot notice
NVRTC_OPTS = ["--gpu-architecture=sm_100a", "--std=c++17", "-lineinfo", "-default-device"]


@pytest.mark.parametrize("strict,uniform_h", [(False, False), (False, True), (True, False)])
def test_every_kernel_unit_of_a_generated_problem_compiles_under_nvrtc(cg, strict, uniform_h):
    """What the GPU box does on demand (NVRTC, one unit per kernel variant), done here without a device: the
    table/info kernels and every sweep variant a generated problem can be asked for -- in the strict build and in the
    FMA-contracted build with per-step sizes and with the step sizes as constant operands (OCS_UNIFORM_H, what a
    linspace grid selects)."""
    nvrtc = pytest.importorskip("cuda.bindings.nvrtc")
    src_api = open(os.path.join(ROOT, "optimal-control-solvers_b200", "csrc", "nvrtc_problem.cu")).read()
    for opt in NVRTC_OPTS:
        assert '"%s"' % opt in src_api, opt          # this test and the library use the same options
    p = cg.harvest_chain_problem()
    units = ["OCS_TT_KERNEL(ocs_gen::Problem, gen)",
             "OCS_EQ_KERNEL(ocs_gen::Problem, gen)",
             "OCS_FWD_KERNEL(ocs_gen::Problem, gen, 0, 0, 0, 0)", "OCS_FWD_KERNEL(ocs_gen::Problem, gen, 1, 1, 1, 0)",
             "OCS_FWD_KERNEL(ocs_gen::Problem, gen, 1, 2, 0, 17)", "OCS_FWD_KERNEL(ocs_gen::Problem, gen, 2, 0, 0, 0)",
             "OCS_FWDG_KERNEL(ocs_gen::Problem, gen, 1, 1, 0, 0)",
             "OCS_REV_KERNEL(ocs_gen::Problem, gen, 0, 1, 0, 0)", "OCS_REV_KERNEL(ocs_gen::Problem, gen, 1, 1, 0, 0)",
             "OCS_REV_KERNEL(ocs_gen::Problem, gen, 2, 0, 1, 0)", "OCS_REVG_KERNEL(ocs_gen::Problem, gen, 1, 2, 0, 17)",
             "OCS_FB_KERNEL(ocs_gen::Problem, gen, 0)", "OCS_FBC_KERNEL(ocs_gen::Problem, gen, 1)"]
    inc = "-I" + os.path.join(ROOT, "optimal-control-solvers_b200", "csrc")
    opts = [o.encode() for o in NVRTC_OPTS + [inc, "--fmad=false" if strict else "--fmad=true"]]
    for unit in units:
        src = ('#define OCS_NS ocs_gen\n#define OCS_STRICT %d\n#define OCS_UNIFORM_H %d\n#include "rk4_kernels.cuh"\n'
               % (int(strict), int(uniform_h)) + p.source + "\n" + unit + "\n")
        err, prog = nvrtc.nvrtcCreateProgram(src.encode(), b"ocs_generated_problem.cu", 0, [], [])
        assert int(err) == 0
        (rc,) = nvrtc.nvrtcCompileProgram(prog, len(opts), opts)
        if int(rc) != 0:
            _, n = nvrtc.nvrtcGetProgramLogSize(prog)
            log = b" " * n
            nvrtc.nvrtcGetProgramLog(prog, log)
            pytest.fail("%s did not compile under NVRTC:\n%s" % (unit, log.decode(errors="replace")[:3000]))
        nvrtc.nvrtcDestroyProgram(prog)


def test_fma_forms_are_cheaper_and_equal(cg):
    """codegen/fma_forms.py: the forms printed for the FMA-contracted build are re-associations of the derived
    expressions (equal as polynomials, and to rounding at random points) with fewer FP64 instructions; the strict
    build's branch of every function still prints the derived expressions."""
    ff = importlib.import_module("optimal-control-solvers_b200.codegen.fma_forms")
    p = cg.harvest_chain_problem()
    s = p.sym
    x, vx, vJ, u, t = s["x"], s["vx"], s["vJ"], s["u"], s["t"]
    rng = np.random.default_rng(5)
    allsyms = [t, *x, *vx, vJ, *u, *s["params"]]
    point = {a: float(rng.uniform(0.3, 1.7)) for a in allsyms}
    before = after = 0
    for fn, dyn, order in (("F", set(x), [*x]), ("X", {*x, *vx, vJ}, [*vx, vJ, *x]), ("U", {*x, *vx, vJ}, [*vx, vJ, *x])):
        defs, forms = ff.fma_forms(s[fn], dyn, order)
        sub = dict(defs)
        for e, f in zip(s[fn], forms):
            full = f.xreplace(sub)
            assert sp.expand(full - e) == 0
            assert abs(float(full.xreplace(point)) - float(e.xreplace(point))) <= 1e-13 * max(1.0, abs(float(e.xreplace(point))))
        uni, grid = {t, *s["params"]}, {t}
        mfs = importlib.import_module("optimal-control-solvers_b200.codegen.make_from_symbolic")
        prep = lambda e: mfs._hoist([e], uni, grid, [])[0]
        before += ff.dag_cost(s[fn], prep)
        after += ff.dag_cost([d for _, d in defs] + forms, prep)
    assert after <= 0.85 * before, (before, after)
    assert "#if defined(OCS_STRICT) && OCS_STRICT" in p.source
    # a sum prints with its non-product terms first: c - a*b contracts into one FMA, -a*b + c need not
    a, b, c = sp.symbols("a b c")
    assert [str(q) for q in ff.order_terms([-a * b, c])] == ["c", "-a*b"]
    assert ff.fma_cost(a * b + c) == 1 and ff.fma_cost(a * b + c * a) == 2 and ff.fma_cost((a + b) * c) == 2
    assert cg.make_from_symbolic("x1**2 + u1**2", ["-x1 + u1"], 1, 1, dict(k=1.0), [[0, 1]], fma=False).source.count("#if") == 0


def test_fma_forms_leave_non_polynomial_problems_alone(cg):
    """Exponentials, divisions and powers of sums of the states are never expanded: such rows keep the derived form
    or use it as an atom, and the generated source still compiles."""
    p = cg.make_from_symbolic("exp(-r*t)*(x1 - a)**2/(1 + x2**2) + u1**2",
                              ["x1*exp(-x2) - k*u1*x1/(1 + x1)", "sqrt(1 + x1**2) - x2*u1"], 2, 1,
                              dict(r=0.1, a=1.0, k=0.5), [[0, 1]], solve_control=False)
    assert "exp(-y[1])" in p.source and "sqrt(" in p.source
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "gen.cu")
        with open(src, "w") as f:
            f.write('#define OCS_NS ocs_gen\n#define OCS_STRICT 0\n#include "rk4_kernels.cuh"\n' + p.source +
                    "\nOCS_TT_KERNEL(ocs_gen::Problem, gen)\nOCS_REV_KERNEL(ocs_gen::Problem, gen, 0, 1, 0, 0)\n")
        r = subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-std=c++17", "-O1",
                            "-I", os.path.join(ROOT, "optimal-control-solvers_b200", "csrc"), "-cubin", "-o",
                            os.path.join(d, "gen.cubin"), src], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr


def test_oracle_twin_of_generated_problem(cg, oracle):
    """The lambdified SymPy twin agrees with the oracle's hand-written TestOCProblem."""
    p = cg.test_problem_symbolic()
    g = oracle_problem(oracle, p)
    ref = oracle.TestOCProblem(**TEST_PARAMS)
    rng = np.random.default_rng(0)
    y, u, v = rng.random((2, 5)), rng.random((1, 5)), rng.random((2, 5))
    assert relerr(g.F(0.7, y, u), ref.F(0.7, y, u)) < 1e-15
    assert relerr(g.dFdx_times_vec(0.7, y, u, v), ref.dFdx_times_vec(0.7, y, u, v)) < 1e-15
    assert relerr(g.dFdu_times_vec(0.7, y, u, v), ref.dFdu_times_vec(0.7, y, u, v)) < 1e-15


# ------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("math", ["fast", "strict"])
def test_generated_test_problem_matches_builtin_and_oracle(ocs, oracle, cg, math):
    tspan = ocs.matlab_linspace(0.0, 10.0, 301)
    gen, builtin = cg.test_problem_symbolic(), ocs.TestOCProblem(TEST_PARAMS, [[0, 1]])
    rng = np.random.default_rng(0)
    x0 = rng.uniform(.5, 2.5, (1, 500))
    res = {}
    for name, prob in (("gen", gen), ("builtin", builtin)):
        integ = ocs.CudaRK4Integrator(tspan, math=math)
        ctl = ocs.PWLinearControl(integ.t, 31, 1)
        integ.attach_control(ctl)
        v = np.random.default_rng(1).random(31)
        res[name] = integ.evaluate(prob, x0, v)
        res[name + "_pv"] = integ.evaluate(prob, x0, np.random.default_rng(2).random((31, 500)))
        integ.close()
    ref = oracle.C.evaluate(list(TEST_PARAMS.values()), tspan, ctl.B, x0, v)
    for key in ("J", "xT", "dJdv", "lam0"):
        assert relerr(res["gen"][key], ref[key]) < TOL, key
        assert relerr(res["gen"][key], res["builtin"][key]) < TOL, key
        assert relerr(res["gen_pv"][key], res["builtin_pv"][key]) < TOL, key


@pytest.mark.gpu
@pytest.mark.parametrize("kind,K", [("PWLinearControl", 11), ("ChebyshevControl", 17)])
def test_six_state_problem_vs_oracle(ocs, oracle, cg, kind, K):
    """cfg3's problem (6 states, 2 controls) at a size the NumPy oracle finishes in seconds."""
    prob = cg.harvest_chain_problem()
    oprob = oracle_problem(oracle, prob)
    N, Bt = 120, 40
    tspan = ocs.matlab_linspace(0.0, 6.0, N + 1)
    integ = ocs.CudaRK4Integrator(tspan)
    ointeg = oracle.RK4Integrator(tspan)
    ctl = getattr(ocs, kind)(integ.t, K, 2)
    octl = getattr(oracle, kind)(ointeg.t, K, 2)
    assert np.array_equal(ctl.B, octl.B)
    integ.attach_control(ctl)
    rng = np.random.default_rng(3)
    x0 = rng.uniform(1.0, 2.5, (6, Bt))
    if kind == "ChebyshevControl":
        v = np.vstack([rng.uniform(.2, .8, (2, Bt)), 0.05 * rng.standard_normal((2 * (K - 1), Bt))])
    else:
        v = rng.random((2 * K, Bt))
    out = integ.evaluate(prob, x0, v)
    u = octl.compute_u(v)
    x, J = ointeg.compute_states(oprob, x0, u)
    lam, dJdu = ointeg.compute_adjoints(oprob, u)
    dJdv = octl.compute_dJdv(dJdu)
    assert relerr(out["J"], J) < TOL and relerr(out["xT"], x[:-1, -1]) < TOL
    assert relerr(out["dJdv"], dJdv) < TOL and relerr(out["lam0"], lam[:, 0]) < TOL
    # shared control + grids through the literal Integrator API
    xg, Jg = integ.compute_states(prob, x0[:, 0], v=v[:, 0])
    lg, dug = integ.compute_adjoints(prob, grids=True)
    assert relerr(xg, x[:, :, 0]) < TOL and relerr(lg, lam[:, :, 0]) < TOL and relerr(dug, dJdu[:, :, 0]) < TOL
    # fb_sweep with the generated ControlChar
    G = 2 * N + 1
    u0 = np.full((2, G), 0.4)
    fb = integ.fb_sweep(prob, x0[:, :4], u0, nSWEEPS=40, relax=0.5)
    ref = oracle.fb_sweep_rk4(oprob, ointeg, x0[:, 0], u0.copy(), oprob.control_char, nSWEEPS=40, relax=0.5)
    assert bool(fb["converged"][0]) == ref["converged"] and fb["sweeps"][0] == ref["sweeps"]
    assert relerr(fb["J"][0], ref["J"]) < 1e-8


# ------------------------------------------------------------------------------ MATLAB twin, string-emission half
def _matlab_ccode_strings(p):
    """The strings MATLAB's ccode would hand to emit_cuda_problem.m for the SymPy problem p: "  t0 = <expr>;" in the
    reference's variable names (x1.., lam1.., u1.., v1.., vJ, parameter names)."""
    s = p.sym
    one = lambda e: "  t0 = " + sp.ccode(e) + ";"
    return ([one(e) for e in s["F"]], [one(e) for e in s["X"]], [one(e) for e in s["U"]],
            [one(e) for e in s["CC"]] if s["CC"] is not None else [])


def _emit_with_matlab_twin(ocs, p):
    """Runs optimal-control-solvers_b200/matlab/emit_cuda_problem.m (unmodified) under the MATLAB-subset interpreter."""
    from oracle.mlab.interp import Frame, Interpreter, MCell
    it = Interpreter([os.path.join(ROOT, "optimal-control-solvers_b200", "matlab")])
    cF, cX, cU, cCC = _matlab_ccode_strings(p)
    out = it.call_named("emit_cuda_problem", [MCell(cF), MCell(cX), MCell(cU), MCell(cCC), float(p.nStates),
                                              float(p.nControls), MCell(list(p.param_names))], 1, Frame())
    return out[0]


@pytest.mark.parametrize("which", ["test_problem_symbolic", "harvest_chain_problem"])
def test_matlab_emission_half_runs_and_compiles(ocs, cg, which):
    """functions/make_from_symbolic.m:27-31 replaced by CUDA emission, MATLAB side: emit_cuda_problem.m executed on
    the expression strings a Symbolic Toolbox `ccode` returns; the source it prints compiles for sm_100a together
    with the sweep kernels (forward, gradient, fb_sweep and equilibrium units)."""
    p = getattr(cg, which)()
    src = _emit_with_matlab_twin(ocs, p)
    assert "static constexpr int n = %d, nc = %d, np = %d;" % (p.nStates, p.nControls, len(p.param_names)) in src
    for name in list(map(str, p.sym["x"])) + list(map(str, p.sym["u"])) + p.param_names:
        # every reference-side name was renamed: none survives as a whole word
        import re
        assert not re.search(r"\b%s\b" % re.escape(name), src.split("struct Problem", 1)[1]), name
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    with tempfile.TemporaryDirectory() as d:
        cu = os.path.join(d, "gen.cu")
        with open(cu, "w") as f:
            f.write('#define OCS_NS ocs_gen\n#include "rk4_kernels.cuh"\n' + src +
                    "\nOCS_TT_KERNEL(ocs_gen::Problem, gen)\nOCS_FWD_KERNEL(ocs_gen::Problem, gen, 1, 1, 0, 0)\n"
                    "OCS_REV_KERNEL(ocs_gen::Problem, gen, 0, 1, 0, 0)\nOCS_FB_KERNEL(ocs_gen::Problem, gen, 0)\n"
                    "OCS_EQ_KERNEL(ocs_gen::Problem, gen)\n")
        r = subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-std=c++17", "-O1",
                            "-I", os.path.join(ROOT, "optimal-control-solvers_b200", "csrc"), "-cubin", "-o",
                            os.path.join(d, "gen.cubin"), cu], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr


@pytest.mark.gpu
def test_matlab_emitted_problem_matches_the_python_backend(ocs, cg):
    """The source printed by the MATLAB twin, bound through ocs_b200_set_problem_source, gives the results of the
    Python backend's source for the same expressions (different expression order: rounding only)."""
    p = cg.harvest_chain_problem()
    twin = cg.CodegenProblem(p.nStates, p.nControls, p.params, p.ControlBounds, _emit_with_matlab_twin(ocs, p), p.sym, p.flops)
    N, Bt = 100, 64
    tspan = ocs.matlab_linspace(0.0, 5.0, N + 1)
    rng = np.random.default_rng(5)
    x0, v = rng.uniform(1.0, 2.5, (6, Bt)), rng.uniform(.2, .8, (2 * 9, Bt))
    res = []
    for prob in (p, twin):
        integ = ocs.CudaRK4Integrator(tspan)
        integ.attach_control(ocs.PWLinearControl(integ.t, 9, 2))
        out = integ.evaluate(prob, x0, v)
        fb = integ.fb_sweep(prob, x0[:, :8], np.full((2, 2 * N + 1), .4), nSWEEPS=3, relax=0.5)
        res.append((out, fb))
        integ.close()
    for key in ("J", "xT", "dJdv", "lam0"):
        assert relerr(res[1][0][key], res[0][0][key]) < 1e-12, key
    assert relerr(res[1][1]["J"], res[0][1]["J"]) < 1e-12 and relerr(res[1][1]["maxChange"], res[0][1]["maxChange"]) < 1e-9


@pytest.mark.gpu
def test_generated_problem_parameter_sweep_matches_builtin(ocs, oracle, cg):
    """Per-trajectory (c, m, r) through a generated problem: the codegen's texp_rates / table_from_texp (exp(-r t) and
    exp(r t) advanced along the grid by recurrence in the NVRTC-compiled sweeps) against the built-in problem and the
    C oracle, which call exp() at every grid point."""
    Bt, N, K = 700, 150, 9
    rng = np.random.default_rng(17)
    c, m, r = rng.uniform(1.0, 2.0, Bt), rng.uniform(2.5, 3.5, Bt), rng.uniform(0.0, 0.4, Bt)
    gen = cg.test_problem_symbolic(c=c, m=m, r=r)
    assert "static constexpr int ntexp = 2;" in gen.source
    builtin = ocs.TestOCProblem(dict(c=c, m=m, r=r), [[0, 1]])
    tspan = ocs.matlab_linspace(0.0, 6.0, N + 1)
    x0, v = rng.uniform(.5, 2.5, (1, Bt)), rng.random((K, Bt))
    res = {}
    for name, prob in (("gen", gen), ("builtin", builtin)):
        integ = ocs.CudaRK4Integrator(tspan)
        ctl = ocs.PWLinearControl(integ.t, K, 1)
        integ.attach_control(ctl)
        res[name] = integ.evaluate(prob, x0, v)
        integ.close()
    ref = oracle.C.evaluate(np.stack([c, m, r]), tspan, ctl.B, x0, v)
    for key in ("J", "xT", "dJdv", "lam0"):
        assert relerr(res["gen"][key], ref[key]) < TOL, key
        assert relerr(res["gen"][key], res["builtin"][key]) < TOL, key


@pytest.mark.gpu
@pytest.mark.parametrize("math", ["fast", "strict"])
def test_six_state_adjoint_equals_central_differences(ocs, cg, math):
    """The property of the reference's tests/backprop_test.m for the generated 6-state problem: the gradient from the
    adjoint sweep (dFdx_times_vec / dFdu_times_vec derived by SymPy, in the fast build in their FMA-shaped forms) equals
    central differences of J, which only uses F.  A wrong derivation of X or U -- which the comparison with the
    lambdified twin of the same expressions could not see -- fails here.  Coefficients of both controls and every
    initial state (lam(:,1), the free-initial-state gradient of single_shooting.m:147-148)."""
    prob = cg.harvest_chain_problem()
    Bt, N, K = 48, 200, 5
    tspan = ocs.matlab_linspace(0.0, 3.0, N + 1)
    integ = ocs.CudaRK4Integrator(tspan, math=math)
    integ.attach_control(ocs.ChebyshevControl(integ.t, K, 2))
    rng = np.random.default_rng(17)
    x0 = rng.uniform(1.0, 2.5, (6, Bt))
    v = 0.05 * rng.standard_normal((2 * K, Bt))
    v[:2] = rng.uniform(0.2, 0.8, (2, Bt))
    base = integ.evaluate(prob, x0, v)
    d = 1e-5
    for k in (0, 1, 2, 5, 2 * K - 1):                      # rows of v: (basis k, control c) at k*nc + c
        vp, vm = v.copy(), v.copy()
        vp[k] += d
        vm[k] -= d
        fd = (integ.evaluate(prob, x0, vp)["J"] - integ.evaluate(prob, x0, vm)["J"]) / (2 * d)
        assert np.abs(fd - base["dJdv"][k]).max() <= 2e-7 * max(1.0, np.abs(base["dJdv"][k]).max()), (k, math)
    for c in range(6):
        xp, xm = x0.copy(), x0.copy()
        xp[c] += d
        xm[c] -= d
        fd = (integ.evaluate(prob, xp, v)["J"] - integ.evaluate(prob, xm, v)["J"]) / (2 * d)
        assert np.abs(fd - base["lam0"][c]).max() <= 2e-7 * max(1.0, np.abs(base["lam0"][c]).max()), (c, math)
    integ.close()
