"""Generates tests/golden/*.npz by executing the reference's UNMODIFIED .m files
(/root/reference) with the MATLAB-subset interpreter oracle/mlab.

    python tests/golden/generate_golden.py

Runs only where /root/reference exists (the build container); the .npz files are
committed and are what the tests read.  Scenarios:
  backprop_test       the reference's own tests/backprop_test.m, as is (rand = default_rng(0))
  golden_*.m          driver scripts in tests/golden/scripts (MATLAB syntax) that only call the
                      reference's classes
MATLAB built-ins (linspace, mtimes, exp, griddedInterpolant, rand) are restated in
oracle/mlab/builtins.py -- see its header for what that implies.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference"

from oracle.mlab.interp import Interpreter, MObject  # noqa: E402


def run(script, seed):
    it = Interpreter([os.path.join(HERE, "scripts"), REF + "/tests", REF + "/Integrator", REF + "/Control",
                      REF + "/OCProblem", REF + "/functions"])
    it.builtins["rng"](np.array([[float(seed)]]))
    fr = it.run_script(script)
    out = {}
    for k, v in fr.vars.items():
        if isinstance(v, np.ndarray):
            out[k] = np.asarray(v, dtype=np.float64)
    # pull interesting properties out of handle objects
    for name, obj in fr.vars.items():
        if isinstance(obj, MObject):
            for pk, pv in obj.props.items():
                if isinstance(pv, np.ndarray) and pv.size:
                    out[f"{name}__{pk}"] = np.asarray(pv, dtype=np.float64)
    out["stdout"] = np.array("".join(it.output))
    return out


def main():
    if not os.path.isdir(REF):
        raise SystemExit("needs /root/reference (run in the build container)")
    for script, seed in [("backprop_test", 0), ("golden_pwconstant", 10), ("golden_pwlinear101", 11),
                         ("golden_chebyshev", 12), ("golden_infinite", 13)]:
        out = run(script, seed)
        path = os.path.join(HERE, script + ".npz")
        np.savez_compressed(path, **out)
        print(script, "->", os.path.relpath(path, ROOT), f"{os.path.getsize(path)} bytes;",
              "J =", repr(float(out["J"].reshape(-1)[0])))
        if str(out["stdout"]):
            print("   stdout:", str(out["stdout"]).strip().replace("\n", " | "))


if __name__ == "__main__":
    main()
