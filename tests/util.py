import numpy as np


def relerr(a, b):
    """Normwise relative error max|a-b| / max|b| per output array (dJdv
    components cross zero, so the 1e-10 bar of the north star is normwise)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    den = np.max(np.abs(b))
    return float(np.max(np.abs(a - b)) / (den if den > 0 else 1.0))


# the north star's FP64 bar
TOL = 1e-10
# strict arithmetic (no FMA contraction) differs from the C oracle only through exp() (<= 1 ulp each)
TOL_STRICT = 5e-14

TEST_PARAMS = dict(c=1.5, m=3.0, r=0.05)   # tests/backprop_test.m:9-11
