"""MATLAB built-ins restated for oracle/mlab/interp.py -- TEST INFRASTRUCTURE.

These are the parts of a golden-vector run that are NOT the reference's own source:

  linspace     MATLAB's formula  d1 + (0:n-1)*(d2-d1)/(n-1)  evaluated as (i*(d2-d1))/(n-1), end
               points forced.  This rounding decides PWConstant indices.
  mtimes (*)   matrix products are summed in ascending index order with separately rounded
               multiply and add (MATLAB calls BLAS; its order/FMA use are unspecified, so
               products are only comparable normwise).
  exp          numpy's (glibc) exp; MATLAB's differs by at most 1 ulp.
  griddedInterpolant(x, v, 'linear', 'nearest')
               v_k + (xq - x_k)/(x_{k+1} - x_k)*(v_{k+1} - v_k) inside, nearest sample outside;
               MathWorks does not publish its exact formula.
  rand         numpy default_rng(seed).random, column-major fill (MATLAB's generator differs;
               the reference's only use, tests/backprop_test.m:19, is unseeded anyway).

All numeric values are float64 ndarrays with ndim >= 2 (MATLAB semantics); logical results
are float 0/1 arrays carrying an `is_logical` marker via a subclass.
"""
from __future__ import annotations

import numpy as np


class Logical(np.ndarray):
    """A float array that remembers it came from a comparison (for logical indexing)."""


def as_logical(a):
    return np.asarray(a, dtype=np.float64).view(Logical)


def empty():
    return np.zeros((0, 0))


def as_array(v):
    if isinstance(v, np.ndarray):
        return v
    if isinstance(v, (int, float, bool, np.floating, np.integer)):
        return np.array([[float(v)]])
    if isinstance(v, str):
        return np.array([[float(ord(c)) for c in v]])
    raise TypeError(f"not a numeric value: {type(v).__name__}")


def scalar(v):
    a = as_array(v)
    if a.size != 1:
        raise ValueError("expected a scalar")
    return float(a.reshape(-1)[0])


def truth(v):
    a = as_array(v)
    return a.size > 0 and bool(np.all(a != 0))


def _squeeze_trailing(a):
    while a.ndim > 2 and a.shape[-1] == 1:
        a = a.reshape(a.shape[:-1])
    return a


# ------------------------------------------------------------------ operators
def mtimes(a, b):
    if a.size == 1 or b.size == 1:
        return a * b
    if a.ndim != 2 or b.ndim != 2 or a.shape[1] != b.shape[0]:
        raise ValueError(f"mtimes: inner dimensions {a.shape} x {b.shape} disagree")
    out = np.zeros((a.shape[0], b.shape[1]))
    for k in range(a.shape[1]):            # ascending k, multiply and add rounded separately
        out = out + a[:, k:k + 1] * b[k:k + 1, :]
    return out


def binop(op, x, y):
    a, b = as_array(x), as_array(y)
    if op == "+":
        return np.asarray(a + b)
    if op == "-":
        return np.asarray(a - b)
    if op == ".*":
        return np.asarray(a * b)
    if op == "./":
        return np.asarray(a / b)
    if op == "*":
        return np.asarray(mtimes(np.asarray(a), np.asarray(b)))
    if op == "/":
        if b.size != 1:
            raise ValueError("mrdivide by a matrix is not supported")
        return np.asarray(a / b)
    if op == ".^" or op == "^":
        if op == "^" and a.size != 1 and b.size != 1:
            raise ValueError("matrix power is not supported")
        if b.size == 1 and float(b.reshape(-1)[0]) == 2.0:
            return np.asarray(a * a)        # x.^2 is computed as x*x (exactly rounded square)
        return np.asarray(np.power(a, b))
    if op in ("<", "<=", ">", ">=", "==", "~="):
        f = {"<": np.less, "<=": np.less_equal, ">": np.greater, ">=": np.greater_equal,
             "==": np.equal, "~=": np.not_equal}[op]
        return as_logical(f(a, b))
    if op == "&":
        return as_logical((a != 0) & (b != 0))
    if op == "|":
        return as_logical((a != 0) | (b != 0))
    raise ValueError(f"operator {op} not supported")


def unop(op, x):
    if op == "'":
        if isinstance(x, str):
            return x
        a = as_array(x)
        if a.ndim != 2:
            raise ValueError("transpose of an N-d array")
        return np.asarray(a.T.copy())
    a = as_array(x)
    if op == "-":
        return np.asarray(-a)
    if op == "+":
        return np.asarray(a)
    if op == "~":
        return as_logical(a == 0)
    raise ValueError(op)


def colon(a, s, b):
    a, s, b = scalar(a), scalar(s), scalar(b)
    if s == 0 or (s > 0 and a > b) or (s < 0 and a < b):
        return np.zeros((1, 0))
    n = int(np.floor((b - a) / s * (1 + 1e-15) + 1e-10))
    return (a + np.arange(n + 1, dtype=np.float64) * s).reshape(1, -1)


def hcat(vals):
    vals = [v for v in vals if not (isinstance(v, np.ndarray) and v.size == 0)]
    if not vals:
        return empty()
    if all(isinstance(v, str) for v in vals):
        return "".join(vals)
    return np.concatenate([np.asarray(as_array(v)) for v in vals], axis=1)


def vcat(rows):
    rows = [r for r in rows if not (isinstance(r, np.ndarray) and r.size == 0)]
    if not rows:
        return empty()
    if len(rows) == 1:
        return rows[0]
    return np.concatenate([np.asarray(as_array(r)) for r in rows], axis=0)


# ------------------------------------------------------------------ indexing
def end_value(arr, dim, ndims_used):
    if ndims_used == 1:
        return arr.size
    if dim < ndims_used - 1:
        return arr.shape[dim] if dim < arr.ndim else 1
    # last subscript folds the remaining dimensions
    rest = arr.shape[dim:] if dim < arr.ndim else (1,)
    return int(np.prod(rest))


def _subs(ix, extent):
    """0-based integer positions of one subscript."""
    if isinstance(ix, tuple) and ix == ("all",):
        return np.arange(extent)
    if isinstance(ix, str) and ix == ":":
        return np.arange(extent)
    a = as_array(ix)
    if isinstance(a, Logical):
        return np.flatnonzero(np.asarray(a).reshape(-1, order="F") != 0)
    p = np.asarray(a).reshape(-1, order="F")
    r = np.rint(p).astype(np.int64)
    if np.any(np.abs(p - r) > 0) or np.any(r < 1):
        raise IndexError("subscript indices must be positive integers")
    return r - 1


def index(arr, idx):
    arr = np.asarray(arr)
    if len(idx) == 1:
        flat = arr.reshape(-1, order="F")
        pos = _subs(idx[0], flat.size)
        if pos.size and pos.max() >= flat.size:
            raise IndexError("index exceeds array bounds")
        out = flat[pos]
        ia = idx[0]
        if isinstance(ia, tuple) or (isinstance(ia, str)):
            return out.reshape(-1, 1)
        ia = as_array(ia)
        if arr.ndim == 2 and arr.shape[0] == 1 and not isinstance(ia, Logical):
            return out.reshape(1, -1)
        if arr.ndim == 2 and (arr.shape[1] == 1) and min(ia.shape) <= 1:
            return out.reshape(-1, 1)
        if arr.ndim == 2 and arr.shape[0] == 1:
            return out.reshape(1, -1)
        if isinstance(ia, Logical):
            return out.reshape(-1, 1)
        return out.reshape(ia.shape, order="F")
    nd = len(idx)
    shape = list(arr.shape) + [1] * max(0, nd - arr.ndim)
    if nd < len(shape):                      # fold trailing dimensions into the last subscript
        shape = shape[:nd - 1] + [int(np.prod(shape[nd - 1:]))]
    a = arr.reshape(shape, order="F")
    pos = [_subs(ix, shape[d]) for d, ix in enumerate(idx)]
    for d, p in enumerate(pos):
        if p.size and p.max() >= shape[d]:
            raise IndexError("index exceeds array bounds")
    out = a[np.ix_(*pos)]
    return _squeeze_trailing(np.asarray(out))


def index_assign(arr, idx, val):
    v = as_array(val)
    if len(idx) == 1:
        shape = arr.shape
        flat = arr.reshape(-1, order="F").copy()
        pos = _subs(idx[0], flat.size)
        if pos.size and pos.max() >= flat.size:
            if arr.ndim == 2 and min(arr.shape) <= 1:      # grow a vector
                n = pos.max() + 1
                grown = np.zeros(n)
                grown[:flat.size] = flat
                flat = grown
                shape = (1, n) if (arr.shape[0] == 1 and arr.size > 0) else (n, 1) if arr.size > 0 else (1, n)
            else:
                raise IndexError("index exceeds array bounds")
        flat[pos] = v.reshape(-1, order="F") if v.size != 1 else v.reshape(-1)[0]
        return flat.reshape(shape, order="F")
    nd = len(idx)
    shape = list(arr.shape) + [1] * max(0, nd - arr.ndim)
    a = arr.reshape(shape, order="F").copy()
    pos = [_subs(ix, shape[d]) for d, ix in enumerate(idx)]
    for d, p in enumerate(pos):
        if p.size and p.max() >= shape[d]:
            raise IndexError("index exceeds array bounds (growing N-d arrays is not supported)")
    tgt = tuple(len(p) for p in pos)
    if v.size == 1:
        a[np.ix_(*pos)] = v.reshape(-1)[0]
    else:
        if int(np.prod(tgt)) != v.size:
            raise ValueError(f"assignment dimension mismatch: target {tgt}, value {v.shape}")
        # MATLAB matches the non-singleton dimensions in order
        vt = [s for s in tgt if s != 1]
        vv = [s for s in v.shape if s != 1]
        if vt != vv:
            raise ValueError(f"assignment dimension mismatch: target {tgt}, value {v.shape}")
        a[np.ix_(*pos)] = np.asarray(v).reshape(tgt, order="F")
    return _squeeze_trailing(a) if len(arr.shape) <= 2 else a.reshape(arr.shape, order="F")


# ------------------------------------------------------------------ built-in functions
class GriddedInterpolant:
    """griddedInterpolant(x, v, method, extrap) for 1-D grids ('linear' + 'nearest'/'linear')."""

    def __init__(self, x, v, method="linear", extrap=None):
        self.x = np.asarray(as_array(x)).reshape(-1)
        self.v = np.asarray(as_array(v)).reshape(-1)
        self.method, self.extrap = method, extrap or method
        if method != "linear":
            raise ValueError(f"griddedInterpolant method {method!r} not supported")

    def __call__(self, xq):
        q = as_array(xq)
        t = np.asarray(q).reshape(-1, order="F")
        x, v = self.x, self.v
        out = np.empty_like(t)
        lo, hi = t < x[0], t > x[-1]
        mid = ~(lo | hi)
        tm = t[mid]
        k = np.clip(np.searchsorted(x, tm, side="right") - 1, 0, len(x) - 2)
        out[mid] = v[k] + (tm - x[k]) / (x[k + 1] - x[k]) * (v[k + 1] - v[k])
        if self.extrap == "nearest":
            out[lo], out[hi] = v[0], v[-1]
        else:
            out[lo] = v[0] + (t[lo] - x[0]) / (x[1] - x[0]) * (v[1] - v[0])
            out[hi] = v[-2] + (t[hi] - x[-2]) / (x[-1] - x[-2]) * (v[-1] - v[-2])
        return out.reshape(q.shape, order="F")


def matlab_linspace(d1, d2, n=100):
    d1, d2, n = scalar(d1), scalar(d2), int(scalar(n))
    if n < 1:
        return np.zeros((1, 0))
    if n == 1:
        return np.array([[d2]])
    n1 = n - 1
    y = d1 + (np.arange(n1 + 1, dtype=np.float64) * (d2 - d1)) / n1
    y[0], y[-1] = d1, d2
    return y.reshape(1, -1)


def _dims(args):
    if len(args) == 0:
        return (1, 1)
    if len(args) == 1:
        a = as_array(args[0])
        if a.size == 1:
            n = int(scalar(a))
            return (n, n)
        return tuple(int(x) for x in np.asarray(a).reshape(-1))
    return tuple(int(scalar(a)) for a in args)


class InputParser:
    """inputParser with addParameter / parse / Results (name-value pairs only)."""
    METHODS = ("addParameter", "parse")

    def __init__(self):
        from .interp import MStruct
        self.defaults = {}
        self.Results = MStruct()

    def call_method(self, name, args, nargout):
        from .interp import MStruct
        if name == "addParameter":
            self.defaults[args[0]] = args[1]
            return []
        if name == "parse":
            res = MStruct(self.defaults)
            if len(args) % 2:
                raise ValueError("inputParser.parse: name-value arguments must come in pairs")
            for k in range(0, len(args), 2):
                if args[k] not in self.defaults:
                    raise ValueError(f"inputParser: unknown parameter {args[k]!r}")
                res[args[k]] = args[k + 1]
            self.Results = res
            return []
        raise ValueError(name)

    def getfield(self, name):
        if name == "Results":
            return self.Results
        raise ValueError(f"inputParser has no field {name}")


class MatlabRaised(RuntimeError):
    """error(id, msg) raised by interpreted code."""

    def __init__(self, ident, msg):
        super().__init__(f"{ident}: {msg}")
        self.ident = ident


def named_handle(interp, name, frame):
    def call(*args, nargout=1):
        return tuple(interp.call_named(name, list(args), nargout, frame))
    call._wants_nargout = True
    return call


def make_builtins(interp):
    rng_state = {"rng": np.random.default_rng(0)}

    def size(a, dim=None, nargout=1):
        a = as_array(a) if not isinstance(a, str) else np.zeros((1, len(a)))
        if dim is not None:
            d = int(scalar(dim)) - 1
            return np.array([[float(a.shape[d] if d < a.ndim else 1)]])
        if nargout <= 1:
            return np.array([[float(s) for s in a.shape]])
        shp = list(a.shape) + [1] * max(0, nargout - a.ndim)
        if nargout < len(shp):
            shp = shp[:nargout - 1] + [int(np.prod(shp[nargout - 1:]))]
        return tuple(np.array([[float(s)]]) for s in shp[:nargout])
    size._wants_nargout = True

    def length(a):
        if isinstance(a, str):
            return np.array([[float(len(a))]])
        a = as_array(a)
        return np.array([[0.0 if a.size == 0 else float(max(a.shape))]])

    def reshape(a, *dims):
        a = as_array(a)
        if len(dims) == 1:
            shp = [int(x) for x in np.asarray(as_array(dims[0])).reshape(-1)]
        else:
            shp = []
            for d in dims:
                d = as_array(d)
                shp.append(-1 if d.size == 0 else int(scalar(d)))
        return np.asarray(a).reshape(shp, order="F")

    def repmat(a, *reps):
        a = as_array(a)
        r = _dims(reps)
        return np.tile(a, r)

    def diff(a):
        a = as_array(a)
        if a.ndim == 2 and a.shape[0] == 1:
            return a[:, 1:] - a[:, :-1]
        return a[1:] - a[:-1]

    def mmax(a, b=None, nargout=1):
        a = as_array(a)
        if b is not None:
            return np.maximum(a, as_array(b))
        v = np.asarray(a).reshape(-1, order="F") if min(a.shape) == 1 else None
        if v is not None:
            k = int(np.argmax(v))
            return (np.array([[v[k]]]), np.array([[k + 1.0]])) if nargout > 1 else np.array([[v[k]]])
        return np.max(a, axis=0, keepdims=True)
    mmax._wants_nargout = True

    def mmin(a, b=None, nargout=1):
        a = as_array(a)
        if b is not None:
            return np.minimum(a, as_array(b))
        v = np.asarray(a).reshape(-1, order="F") if min(a.shape) == 1 else None
        if v is not None:
            k = int(np.argmin(v))
            return (np.array([[v[k]]]), np.array([[k + 1.0]])) if nargout > 1 else np.array([[v[k]]])
        return np.min(a, axis=0, keepdims=True)
    mmin._wants_nargout = True

    def msum(a, dim=None):
        a = as_array(a)
        if dim is not None:
            return np.sum(a, axis=int(scalar(dim)) - 1, keepdims=True)
        if min(a.shape) == 1:
            s = 0.0
            for x in np.asarray(a).reshape(-1, order="F"):   # sequential, like a plain loop
                s = s + x
            return np.array([[s]])
        return np.sum(a, axis=0, keepdims=True)

    def rand(*dims):
        shp = _dims(dims)
        n = int(np.prod(shp))
        return rng_state["rng"].random(n).reshape(shp, order="F")

    def rng(seed=0, *a):
        rng_state["rng"] = np.random.default_rng(int(scalar(seed)))
        return empty()

    def fprintf(*args):
        if not args:
            return empty()
        fmt = args[0]
        vals = []
        for a in args[1:]:
            if isinstance(a, str):
                vals.append(a)
            else:
                vals.extend(float(x) for x in np.asarray(as_array(a)).reshape(-1, order="F"))
        py = fmt.replace("\\n", "\n").replace("\\t", "\t")
        spec = [m for m in __import__("re").finditer(r"%[-+0-9.]*[sdfge]", py)]
        out = py
        for m, v in zip(spec, vals):
            rep = (f"{v:.5e}" if isinstance(v, float) and m.group()[-1] == "s" else
                   (m.group().replace("d", "g") % v if not isinstance(v, str) else v))
            out = out.replace(m.group(), rep, 1)
        interp.output.append(out)
        return empty()

    def struct(*args):
        from .interp import MStruct
        s = MStruct()
        for k in range(0, len(args), 2):
            s[args[k]] = args[k + 1]
        return s

    def isprop(obj, name):
        from .interp import MObject
        return np.array([[float(isinstance(obj, MObject) and name in obj.props)]])

    def ismethod(obj, name):
        from .interp import MObject
        ok = isinstance(obj, MObject) and interp.find_method(obj.cls, name)[0] is not None
        return np.array([[float(ok)]])

    def isfield(s, name):
        return np.array([[float(isinstance(s, dict) and name in s)]])

    def griddedInterpolant(x, v, method="linear", extrap=None):
        gi = GriddedInterpolant(x, v, method, extrap)
        return lambda q: gi(q)

    def isa(a, cls):
        from .interp import MObject
        if isinstance(a, MObject):
            return np.array([[float(any(c["name"] == cls for c in interp.class_chain(a.cls)) or
                                    cls in [s_ for c in interp.class_chain(a.cls) for s_ in c["supers"]])]])
        if cls in ("numeric", "double", "float"):
            return np.array([[float(isinstance(a, np.ndarray))]])
        return np.array([[0.0]])

    def isequal(a, b):
        from .interp import MObject
        if isinstance(a, MObject) or isinstance(b, MObject):
            return np.array([[float(a is b)]])
        if isinstance(a, str) or isinstance(b, str):
            return np.array([[float(a == b)]])
        a, b = as_array(a), as_array(b)
        return np.array([[float(a.shape == b.shape and bool(np.all(a == b)))]])

    def merror(ident, *rest):
        msg = rest[0] if rest else ident
        raise MatlabRaised(ident, msg if isinstance(msg, str) else str(msg))

    def fieldnames(s_):
        from .interp import MCell
        return MCell(list(s_.keys()))

    def sprintf(fmt, *args):
        """MATLAB sprintf for the conversions the code generators use (%d %i %s %g %f %e, flags/width/precision),
        escapes \\n \\t \\\\ and %%; numeric array arguments are consumed element by element."""
        import re
        vals = []
        for a in args:
            if isinstance(a, str):
                vals.append(a)
            else:
                vals.extend(float(x) for x in np.asarray(as_array(a)).reshape(-1, order="F"))
        py = fmt.replace("\\n", "\n").replace("\\t", "\t").replace("\\\\", "\\")
        it = iter(vals)

        def one(m):
            if m.group() == "%%":
                return "%"
            v = next(it)
            conv = m.group()[-1]
            if conv in "di":
                return (m.group()[:-1] + "d") % int(round(v)) if not isinstance(v, str) else v
            if conv == "s":
                return (m.group() % v) if isinstance(v, str) else (m.group()[:-1] + "g") % v
            return m.group() % v
        return re.sub(r"%%|%[-+ 0#]*\d*(?:\.\d+)?[disgfe]", one, py)

    def _mregex(pat):
        # MATLAB word anchors \< \> -> Python \b
        return pat.replace("\\<", "\\b").replace("\\>", "\\b")

    def regexprep(s_, pat, rep):
        import re
        rep_py = re.sub(r"\$(\d)", r"\\\1", rep.replace("\\", "\\\\"))
        return re.sub(_mregex(pat), rep_py, s_)

    def strrep(s_, a, b):
        return s_.replace(a, b)

    bi = {
        "inputParser": lambda: InputParser(),
        "isequal": isequal, "error": merror, "fieldnames": fieldnames,
        "sprintf": sprintf, "regexprep": regexprep, "strrep": strrep,
        "num2str": lambda a: ("%g" % scalar(a)),
        "uint64": lambda a: np.asarray(as_array(a), dtype=np.float64).copy(),
        "int32": lambda a: np.round(as_array(a)),
        "logical": lambda a: as_logical(as_array(a) != 0),
        "zeros": lambda *d: np.zeros(_dims(d)),
        "ones": lambda *d: np.ones(_dims(d)),
        "nan": lambda *d: np.full(_dims(d), np.nan),
        "NaN": lambda *d: np.full(_dims(d), np.nan),
        "inf": lambda *d: np.full(_dims(d), np.inf),
        "Inf": lambda *d: np.full(_dims(d), np.inf),
        "eye": lambda *d: np.eye(*_dims(d)),
        "pi": lambda: np.array([[np.pi]]),
        "eps": lambda: np.array([[np.finfo(float).eps]]),
        "true": lambda: as_logical(np.array([[1.0]])),
        "false": lambda: as_logical(np.array([[0.0]])),
        "size": size, "length": length,
        "numel": lambda a: np.array([[float(len(a) if isinstance(a, (list, str)) else as_array(a).size)]]),
        "isempty": lambda a: np.array([[float((len(a) == 0) if isinstance(a, (str, dict, list)) else
                                             (as_array(a).size == 0 if isinstance(a, (np.ndarray, int, float)) else False))]]),
        "reshape": reshape, "repmat": repmat, "diff": diff, "linspace": matlab_linspace,
        "double": lambda a: np.asarray(as_array(a), dtype=np.float64).copy(),
        "exp": lambda a: np.exp(as_array(a)), "log": lambda a: np.log(as_array(a)),
        "sqrt": lambda a: np.sqrt(as_array(a)), "abs": lambda a: np.abs(as_array(a)),
        "sin": lambda a: np.sin(as_array(a)), "cos": lambda a: np.cos(as_array(a)),
        "floor": lambda a: np.floor(as_array(a)), "ceil": lambda a: np.ceil(as_array(a)),
        "round": lambda a: np.round(as_array(a)),
        "max": mmax, "min": mmin, "sum": msum,
        "cumsum": lambda a: np.cumsum(as_array(a), axis=1 if as_array(a).shape[0] == 1 else 0),
        "flip": lambda a, d=None: np.flip(as_array(a), axis=(int(scalar(d)) - 1) if d is not None else (1 if as_array(a).shape[0] == 1 else 0)),
        "rand": rand, "rng": rng, "fprintf": fprintf, "disp": lambda a: interp.output.append(str(a)) or empty(),
        "struct": struct, "isprop": isprop, "ismethod": ismethod, "isfield": isfield,
        "griddedInterpolant": griddedInterpolant,
        "isa": isa,
        "strcmp": lambda a, b: np.array([[float(a == b)]]),
        "any": lambda a: np.array([[float(np.any(as_array(a) != 0))]]),
        "all": lambda a: np.array([[float(np.all(as_array(a) != 0))]]),
        "isnan": lambda a: as_logical(np.isnan(as_array(a))),
        "isfinite": lambda a: as_logical(np.isfinite(as_array(a))),
        "mod": lambda a, b: np.mod(as_array(a), as_array(b)),
        "tic": lambda: empty(), "toc": lambda: np.array([[0.0]]),
        "close": lambda *a: empty(), "clear": lambda *a: empty(), "clc": lambda *a: empty(),
    }
    return bi
