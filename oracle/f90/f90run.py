"""f90run -- run UNMODIFIED Fortran 90 procedures of the reference without a Fortran compiler.

TEST INFRASTRUCTURE ONLY (like everything under oracle/): nothing in the product imports this.

Why this exists.  The reference (xiaocanli/stochastic-parker) is Fortran + MPI + HDF5 and neither
this image nor the B200 box has any Fortran front-end (profiles/r02a_fortran_probe.log), so the
C restatement in oracle/gpat_oracle.c could never be checked against the reference's own arithmetic.
This module closes that gap as far as it can be closed here: it reads the reference's source files
where they lie (/root/reference/src/modules/*.f90), translates the procedures asked for -- text
untouched, statement by statement -- into Python and executes them with Fortran's evaluation rules:

  * default-real literals (1.0, 0.1, 0.37) are single precision, `_dp` / `d0` literals double;
    real(sp) x real(dp) promotes to double, integer x real(k) to real(k), integer / integer truncates;
  * expressions are evaluated left to right inside one precedence level and parentheses are kept,
    no re-association, no FMA contraction (what gfortran does without -ffast-math / -ffp-contract=fast);
  * `x ** n` with an integer n is libgcc's __powidf2 (square-and-multiply), real exponents go to
    libm's pow / powf; sqrt, exp, log, log10, sin, cos ... are the C library's (the same glibc libm
    a gfortran-built binary would call, and the one oracle/gpat_oracle.c links);
  * assignments convert to the declared type and kind of the target; derived-type assignment copies;
  * arrays keep their declared lower bounds, sections are views, whole-array and section arithmetic
    is element-wise in the kind of the operands; array constructors, DO / DO WHILE / EXIT / CYCLE,
    block and one-line IF, CALL with scalar intent(out)/(inout) copy-back, functions with RESULT;
  * cpp conditionals (#if defined ...) are evaluated against an explicit set of defines.

What it is not: a compiler.  I/O statements are skipped (WRITE / PRINT) or rejected, there is no
SELECT CASE / WHERE / FORALL / pointers / generic interfaces (none of the procedures on the hot path
uses them), and MPI / OpenMP / HDF5 / mt_stream entry points are supplied by the caller as Python
stubs (`externals`).  Values: integers are Python ints, reals numpy float32 / float64 scalars,
logicals bools, arrays `FArray`, derived types instances of generated `FStruct` classes.
"""
import ctypes
import ctypes.util
import math
import os
import re

import numpy as np

# ------------------------------------------------------------------------------------------------
# run-time support
# ------------------------------------------------------------------------------------------------
_libm = ctypes.CDLL(ctypes.util.find_library("m") or "libm.so.6")


def _bind(name, n=1):
    d = getattr(_libm, name)
    d.restype = ctypes.c_double
    d.argtypes = [ctypes.c_double] * n
    f = getattr(_libm, name + "f")
    f.restype = ctypes.c_float
    f.argtypes = [ctypes.c_float] * n
    return d, f


_LIBM = {n: _bind(n) for n in ("sqrt", "exp", "log", "log10", "sin", "cos", "tan", "asin", "acos", "atan",
                               "sinh", "cosh", "tanh", "erf")}
_LIBM2 = {n: _bind(n, 2) for n in ("pow", "atan2", "fmod")}

f4 = np.float32
f8 = np.float64


class FortranStop(Exception):
    pass


class Undefined:
    """Value of a variable that was declared but never assigned: any use raises."""

    def __init__(self, name):
        self.name = name

    def _bad(self, *a, **k):
        raise RuntimeError(f"reference reads '{self.name}' before assigning it")

    __add__ = __radd__ = __sub__ = __rsub__ = __mul__ = __rmul__ = __truediv__ = __rtruediv__ = _bad
    __lt__ = __le__ = __gt__ = __ge__ = __neg__ = __float__ = __int__ = __bool__ = __index__ = _bad
    __array_ufunc__ = None

    def __eq__(self, o):
        self._bad()

    def __ne__(self, o):
        self._bad()

    __hash__ = None


def is_real(x):
    return isinstance(x, (np.float32, np.float64))


def _unary_math(name, x):
    if isinstance(x, FArray):
        x = x.a
    if isinstance(x, np.ndarray):
        out = np.empty_like(x)
        flat_in, flat_out = x.reshape(-1), out.reshape(-1)
        for i in range(flat_in.size):
            flat_out[i] = _unary_math(name, flat_in[i])
        return out
    d, f = _LIBM[name]
    if isinstance(x, np.float32):
        return f4(f(float(x)))
    if isinstance(x, np.float64):
        return f8(d(float(x)))
    raise TypeError(f"{name}() of a non-real value {x!r}")


def _powi(a, n):
    """libgcc __powidf2 / __powisf2: what gfortran emits for real ** integer."""
    recip = n < 0
    n = abs(n)
    one = a.dtype.type(1)
    r = one
    while True:
        if n & 1:
            r = r * a
        n >>= 1
        if n == 0:
            break
        a = a * a
    return one / r if recip else r


def f_pow(a, b):
    if isinstance(a, FArray):
        a = a.a
    if isinstance(a, np.ndarray):
        out = np.empty(a.shape, dtype=np.result_type(a.dtype, b) if is_real(b) else a.dtype)
        fi, fo = a.reshape(-1), out.reshape(-1)
        for i in range(fi.size):
            fo[i] = f_pow(fi[i] if a.dtype.kind == "f" else int(fi[i]), b)
        return out
    if isinstance(b, int):
        if isinstance(a, int):
            if b < 0:
                return 0 if abs(a) > 1 else (1 if a == 1 else (1 if b % 2 == 0 else -1) if a == -1 else 1 // 0)
            return a ** b
        return _powi(a, b)
    if isinstance(a, int):  # integer ** real -> real of b's kind
        a = b.dtype.type(a)
    if isinstance(a, np.float32) and isinstance(b, np.float32):
        return f4(_LIBM2["pow"][1](float(a), float(b)))
    return f8(_LIBM2["pow"][0](float(a), float(b)))


def f_div(a, b):
    if isinstance(a, int) and isinstance(b, int):
        q = abs(a) // abs(b)
        return q if (a >= 0) == (b >= 0) else -q
    if isinstance(a, FArray):
        a = a.a
    if isinstance(b, FArray):
        b = b.a
    return a / b


class FArray:
    """A Fortran array: numpy storage + lower bounds.  Indexing uses Fortran subscripts; a Python
    slice object inside a subscript means lo:hi:stride with an INCLUSIVE hi."""

    __array_ufunc__ = None
    __slots__ = ("a", "lb")

    def __init__(self, a, lb=None):
        object.__setattr__(self, "a", a)
        object.__setattr__(self, "lb", tuple(lb) if lb is not None else (1,) * a.ndim)

    # -- construction ---------------------------------------------------------------------------
    @staticmethod
    def alloc(dtype, bounds, fill=None):
        """bounds: list of (lb, ub)."""
        shape = [max(0, ub - lb + 1) for lb, ub in bounds]
        if dtype is object:
            a = np.empty(shape, dtype=object, order="F")
        else:
            a = np.zeros(shape, dtype=dtype, order="F")
            if dtype in (np.float32, np.float64):
                a[...] = np.nan  # unassigned reals poison whatever they touch
        if fill is not None:
            flat = a.reshape(-1, order="F") if a.size else a
            for i in range(a.size):
                flat[i] = fill()
            a = flat.reshape(shape, order="F")
        return FArray(a, [lb for lb, _ in bounds])

    def rebased(self, lbs):
        return FArray(self.a, lbs)

    # -- subscripts -----------------------------------------------------------------------------
    def _key(self, key):
        if not isinstance(key, tuple):
            key = (key,)
        if len(key) != self.a.ndim:
            raise IndexError(f"rank-{self.a.ndim} array referenced with {len(key)} subscripts")
        out, scalar = [], True
        for k, lb, n in zip(key, self.lb, self.a.shape):
            if isinstance(k, slice):
                scalar = False
                st = 1 if k.step is None else int(k.step)
                lo = lb if k.start is None else int(k.start)
                hi = lb + n - 1 if k.stop is None else int(k.stop)
                if st > 0:
                    if lo < lb or hi > lb + n - 1:
                        if hi >= lo:
                            raise IndexError(f"section {lo}:{hi} outside bounds {lb}:{lb + n - 1}")
                    cnt = max(0, (hi - lo + st) // st)
                    out.append(slice(lo - lb, lo - lb + cnt * st if cnt else lo - lb, st))
                else:
                    raise NotImplementedError("negative section stride")
            else:
                i = int(k) - lb
                if i < 0 or i >= n:
                    raise IndexError(f"subscript {k} outside bounds {lb}:{lb + n - 1}")
                out.append(i)
        return tuple(out), scalar

    def __getitem__(self, key):
        k, scalar = self._key(key)
        v = self.a[k]
        if scalar:
            kind = self.a.dtype.kind
            if kind == "i":
                return int(v)
            if kind == "b":
                return bool(v)
            return v
        return v  # a view: sections have lower bound 1

    def __setitem__(self, key, val):
        k, scalar = self._key(key)
        self._store(k, val, scalar)

    def _store(self, k, val, scalar):
        if isinstance(val, FArray):
            val = val.a
        if self.a.dtype == object:
            if scalar:
                self.a[k] = val.copy()
            elif isinstance(val, np.ndarray):
                dst = self.a[k]
                if dst.shape != val.shape:
                    raise ValueError("shape mismatch in derived-type section assignment")
                src = [v.copy() for v in val.reshape(-1)]
                df = dst.reshape(-1) if dst.flags.c_contiguous or dst.flags.f_contiguous else None
                if df is None or not np.shares_memory(df, dst):
                    for idx, s in zip(np.ndindex(dst.shape), src):
                        dst[idx] = s
                else:
                    for i, s in enumerate(src):
                        df[i] = s
            else:
                dst = self.a[k]
                for idx in np.ndindex(dst.shape):
                    dst[idx] = val.copy()
            return
        if isinstance(val, Undefined):
            val._bad()
        if self.a.dtype.kind == "i" and (is_real(val) or (isinstance(val, np.ndarray) and val.dtype.kind == "f")):
            val = np.trunc(val)
        self.a[k] = val

    def assign(self, val):
        """whole-array assignment"""
        self._store(tuple(slice(None) for _ in self.a.shape), val, False)

    def set_component(self, name, val):
        """array%component = scalar"""
        for v in self.a.reshape(-1):
            setattr(v, name, val)

    # -- arithmetic: element-wise, result is a plain ndarray (lower bounds 1) --------------------------
    def __add__(s, o): return s.a + _raw(o)
    def __radd__(s, o): return _raw(o) + s.a
    def __sub__(s, o): return s.a - _raw(o)
    def __rsub__(s, o): return _raw(o) - s.a
    def __mul__(s, o): return s.a * _raw(o)
    def __rmul__(s, o): return _raw(o) * s.a
    def __truediv__(s, o): return s.a / _raw(o)
    def __rtruediv__(s, o): return _raw(o) / s.a
    def __neg__(s): return -s.a
    def __lt__(s, o): return s.a < _raw(o)
    def __le__(s, o): return s.a <= _raw(o)
    def __gt__(s, o): return s.a > _raw(o)
    def __ge__(s, o): return s.a >= _raw(o)
    def __len__(s): return s.a.shape[0]


def _raw(x):
    return x.a if isinstance(x, FArray) else x


def realloc_assign(cur, val, dtype):
    """whole-array assignment to an ALLOCATABLE array (Fortran 2003 semantics)"""
    raw = val.a if isinstance(val, FArray) else val
    if isinstance(raw, np.ndarray) and (cur is None or cur.a.shape != raw.shape):
        lb = val.lb if isinstance(val, FArray) else (1,) * raw.ndim
        new = FArray(np.empty(raw.shape, dtype=dtype, order="F"), lb)
        new.assign(val)
        return new
    if cur is None:
        raise RuntimeError("scalar assigned to an unallocated array")
    cur.assign(val)
    return cur


class FStruct:
    """Base of generated derived-type classes.  _fields: name -> converter."""
    __slots__ = ()
    _fields = {}
    _inits = {}

    def __init__(self):
        for n in self._fields:
            object.__setattr__(self, n, self._inits[n]() if n in self._inits else Undefined(f"{type(self).__name__}%{n}"))

    def __setattr__(self, name, val):
        object.__setattr__(self, name, self._fields[name](val))

    def copy(self):
        c = object.__new__(type(self))
        for n in self._fields:
            v = getattr(self, n)
            if isinstance(v, FStruct):
                v = v.copy()
            elif isinstance(v, FArray):
                v = FArray(v.a.copy(order="F"), v.lb)
            object.__setattr__(c, n, v)
        return c

    def __repr__(self):
        return type(self).__name__ + "(" + ", ".join(f"{n}={getattr(self, n)!r}" for n in self._fields) + ")"


def cv_int(v):
    if isinstance(v, Undefined):
        v._bad()
    return int(v)  # truncation toward zero, as Fortran's real -> integer assignment


def cv_r4(v):
    if isinstance(v, Undefined):
        v._bad()
    return f4(v)


def cv_r8(v):
    if isinstance(v, Undefined):
        v._bad()
    return f8(v)


def cv_bool(v):
    if isinstance(v, Undefined):
        v._bad()
    return bool(v)


def cv_any(v):
    return v


class FStr(str):
    """A Fortran character value: substring s(i:j) is 1-based with an inclusive end, comparison ignores trailing blanks"""

    def __getitem__(self, key):
        if isinstance(key, slice):
            lo = 1 if key.start is None else int(key.start)
            hi = len(self) if key.stop is None else int(key.stop)
            return FStr(str.__getitem__(self, slice(lo - 1, hi)))
        return FStr(str.__getitem__(self, int(key) - 1))

    def __eq__(self, other):
        return str(self).rstrip() == str(other).rstrip()

    def __ne__(self, other):
        return not self.__eq__(other)

    __hash__ = str.__hash__


def cv_char(v):
    return v if isinstance(v, (FStr, Undefined)) or v is None else FStr(v)


def cv_struct(v):
    return v.copy()


class ModSpace:
    """Variables of one Fortran module, with typed assignment."""

    def __init__(self, name):
        object.__setattr__(self, "_name", name)
        object.__setattr__(self, "_conv", {})

    def declare(self, name, conv, value):
        self._conv[name] = conv
        object.__setattr__(self, name, value)

    def __setattr__(self, name, val):
        conv = self._conv.get(name)
        if conv is None:
            raise AttributeError(f"module {self._name} has no variable '{name}'")
        cur = self.__dict__.get(name)
        if isinstance(cur, FArray) and not isinstance(val, FArray):
            cur.assign(val)
            return
        object.__setattr__(self, name, conv(val) if not isinstance(val, FArray) else val)

    def __getattr__(self, name):
        raise AttributeError(f"module {self._name}: '{name}' is not declared")


class DoRange:
    __slots__ = ("start", "count", "step", "final")

    def __init__(self, a, b, c=1):
        a, b, c = int(a), int(b), int(c)
        self.start, self.step = a, c
        self.count = max(0, (b - a + c) // c)
        self.final = a + self.count * c

    def __iter__(self):
        v = self.start
        for _ in range(self.count):
            yield v
            v += self.step


# -- intrinsics -------------------------------------------------------------------------------------
def _mk_unary(name):
    return lambda x: _unary_math(name, x)


def i_abs(x):
    x = _raw(x)
    return abs(x)


def i_min(*a):
    m = a[0]
    for v in a[1:]:
        if v < m:
            m = v
    return m


def i_max(*a):
    m = a[0]
    for v in a[1:]:
        if v > m:
            m = v
    return m


def i_floor(x, kind=None):
    return int(math.floor(x))


def i_ceiling(x, kind=None):
    return int(math.ceil(x))


def i_int(x, kind=None):
    if isinstance(x, str):  # BOZ literal
        return int(x, 16)
    return int(x)


def i_nint(x, kind=None):
    x = float(x)
    return int(math.floor(x + 0.5)) if x >= 0 else -int(math.floor(-x + 0.5))


def i_real(x, kind=None):
    x = _raw(x)
    if kind is None:
        return f4(x) if not isinstance(x, np.ndarray) else x.astype(np.float32)
    t = f4 if kind == 4 else f8
    return t(x) if not isinstance(x, np.ndarray) else x.astype(t)


def i_dble(x):
    x = _raw(x)
    return f8(x) if not isinstance(x, np.ndarray) else x.astype(np.float64)


def i_mod(a, p):
    if isinstance(a, int) and isinstance(p, int):
        return a - f_div(a, p) * p
    if isinstance(a, np.float32) and isinstance(p, np.float32):
        return f4(_LIBM2["fmod"][1](float(a), float(p)))
    return f8(_LIBM2["fmod"][0](float(a), float(p)))


def i_modulo(a, p):
    if isinstance(a, int) and isinstance(p, int):
        return a - (a // p) * p
    return a - np.floor(a / p) * p


def i_sign(a, b):
    m = abs(a)
    neg = (b < 0) or (is_real(b) and b == 0 and np.signbit(b))
    return -m if neg else m


def i_epsilon(x):
    t = type(x) if is_real(x) else np.float32
    return t(np.finfo(t).eps)


def i_huge(x):
    if isinstance(x, int):
        return 2147483647
    return type(x)(np.finfo(type(x)).max)


def i_tiny(x):
    return type(x)(np.finfo(type(x)).tiny)


def i_sum(x, dim=None):
    x = _raw(x)
    flat = x.reshape(-1, order="F")
    if x.dtype.kind == "i":
        return int(flat.sum())
    s = x.dtype.type(0)
    for v in flat:  # sequential, array-element order (gfortran's inline expansion without -ffast-math)
        s = s + v
    return s


def i_maxval(x):
    x = _raw(x)
    v = x.max()
    return int(v) if x.dtype.kind == "i" else v


def i_minval(x):
    x = _raw(x)
    v = x.min()
    return int(v) if x.dtype.kind == "i" else v


def i_size(x, dim=None):
    x = _raw(x)
    return int(x.size) if dim is None else int(x.shape[dim - 1])


def i_ubound(x, dim=None):
    if isinstance(x, FArray):
        ub = [lb + n - 1 for lb, n in zip(x.lb, x.a.shape)]
    else:
        ub = list(x.shape)
    return ub[dim - 1] if dim is not None else FArray(np.array(ub, dtype=np.int64))


def i_lbound(x, dim=None):
    lb = list(x.lb) if isinstance(x, FArray) else [1] * x.ndim
    return lb[dim - 1] if dim is not None else FArray(np.array(lb, dtype=np.int64))


def i_dot_product(a, b):
    a, b = _raw(a), _raw(b)
    s = np.result_type(a.dtype, b.dtype).type(0)
    for u, v in zip(a.reshape(-1, order="F"), b.reshape(-1, order="F")):
        s = s + u * v
    return s


def i_findloc(array, value, dim=None, mask=None, kind=None, back=False):
    a = _raw(array)
    if a.ndim != 1:
        raise NotImplementedError("findloc of a rank > 1 array")
    hit = np.flatnonzero(a == value)
    if len(hit) == 0:
        return 0
    return int(hit[-1] if back else hit[0]) + 1


def i_maxloc(array, dim=None, mask=None, kind=None, back=False):
    """first location of the maximum (1-based, relative to the section: lower bound 1)"""
    a = _raw(array)
    if dim is None:
        if a.ndim != 1:
            raise NotImplementedError("maxloc without dim on a rank > 1 array")
        return np.array([int(np.argmax(a)) + 1], dtype=np.int64)
    r = np.argmax(a, axis=dim - 1) + 1   # numpy returns the first occurrence, like Fortran without back=
    return int(r) if a.ndim == 1 else r.astype(np.int64)


def i_merge(t, f, mask):
    return t if mask else f


def i_allocated(x):
    return x is not None and not isinstance(x, Undefined)


def i_present(x):
    return x is not None


def i_trim(s):
    return s.rstrip()


def i_kind(x):
    return 4 if isinstance(x, np.float32) else 8 if isinstance(x, np.float64) else 4


def i_arrcons(items):
    flat = []
    for it in items:
        it = _raw(it)
        if isinstance(it, np.ndarray):
            flat.extend(it.reshape(-1, order="F").tolist() if it.dtype.kind == "i" else list(it.reshape(-1, order="F")))
        else:
            flat.append(it)
    if all(isinstance(v, int) for v in flat):
        return np.array(flat, dtype=np.int64)
    if all(isinstance(v, bool) for v in flat):
        return np.array(flat, dtype=bool)
    t = np.float64 if any(isinstance(v, np.float64) for v in flat) else np.float32
    return np.array(flat, dtype=t)


INTRINSICS = {
    "abs": i_abs, "dabs": i_abs, "min": i_min, "max": i_max, "dmin1": i_min, "dmax1": i_max,
    "floor": i_floor, "ceiling": i_ceiling, "int": i_int, "nint": i_nint, "real": i_real, "dble": i_dble,
    "mod": i_mod, "modulo": i_modulo, "sign": i_sign, "epsilon": i_epsilon, "huge": i_huge, "tiny": i_tiny,
    "sum": i_sum, "maxval": i_maxval, "minval": i_minval, "size": i_size, "ubound": i_ubound,
    "lbound": i_lbound, "dot_product": i_dot_product, "merge": i_merge, "allocated": i_allocated,
    "present": i_present, "trim": i_trim, "kind": i_kind, "findloc": i_findloc, "maxloc": i_maxloc,
}
for _n in ("sqrt", "exp", "log", "log10", "sin", "cos", "tan", "asin", "acos", "atan", "sinh", "cosh", "tanh", "erf"):
    INTRINSICS[_n] = _mk_unary(_n)
    INTRINSICS["d" + _n] = _mk_unary(_n)
INTRINSICS["alog"] = INTRINSICS["log"]


def _atan2(y, x):
    if isinstance(y, np.float32) and isinstance(x, np.float32):
        return f4(_LIBM2["atan2"][1](float(y), float(x)))
    return f8(_LIBM2["atan2"][0](float(y), float(x)))


INTRINSICS["atan2"] = _atan2
INTRINSICS["datan2"] = _atan2

# ------------------------------------------------------------------------------------------------
# source reader: cpp conditionals, comments, continuations
# ------------------------------------------------------------------------------------------------


def _strip_comment(line):
    q = None
    for i, c in enumerate(line):
        if q:
            if c == q:
                q = None
        elif c in "'\"":
            q = c
        elif c == "!":
            return line[:i]
    return line


def read_logical_lines(path, defines=()):
    """-> list of (first_line_number, text) with comments removed and continuations joined."""
    out, cur, cur_no = [], "", 0
    stack = []  # cpp: True = taking lines
    with open(path) as f:
        for no, raw in enumerate(f, 1):
            s = raw.rstrip("\n")
            st = s.strip()
            if st.startswith("#"):
                d = st[1:].strip()
                if d.startswith("ifdef"):
                    stack.append(d.split()[1] in defines)
                elif d.startswith("ifndef"):
                    stack.append(d.split()[1] not in defines)
                elif d.startswith("if"):
                    m = re.findall(r"(!?)\s*defined\s*\(?\s*(\w+)\s*\)?", d)
                    if not m:
                        raise NotImplementedError(f"{path}:{no}: cpp expression {d!r}")
                    vals = [(name in defines) != (neg == "!") for neg, name in m]
                    stack.append(all(vals) if "&&" in d or len(vals) == 1 else any(vals))
                elif d.startswith("else"):
                    stack[-1] = not stack[-1]
                elif d.startswith("endif"):
                    stack.pop()
                elif d.startswith("define") or d.startswith("include"):
                    pass
                else:
                    raise NotImplementedError(f"{path}:{no}: cpp directive {d!r}")
                continue
            if not all(stack):
                continue
            s = _strip_comment(s).strip()
            if not s:
                continue
            if cur:
                if s.startswith("&"):
                    s = s[1:].lstrip()
                cur += " " + s
            else:
                cur, cur_no = s, no
            if cur.endswith("&"):
                cur = cur[:-1].rstrip()
                continue
            out.append((cur_no, cur))
            cur = ""
    return out


# ------------------------------------------------------------------------------------------------
# lexer
# ------------------------------------------------------------------------------------------------
_TOKEN = re.compile(r"""
    (?P<ws>\s+)
  | (?P<real>(?:\d+\.\d*|\.\d+|\d+)(?:[edED][+-]?\d+)?(?:_\w+)?)
  | (?P<dotop>\.(?:and|or|not|eqv|neqv|eq|ne|lt|le|gt|ge|true|false)\.)
  | (?P<name>[A-Za-z]\w*)
  | (?P<str>'(?:[^']|'')*'|"(?:[^"]|"")*")
  | (?P<op>\*\*|//|==|/=|<=|>=|=>|::|\(/|/\)|[-+*/<>=(),:%\[\]])
""", re.X | re.I)


def tokenize(text):
    toks, pos = [], 0
    while pos < len(text):
        m = _TOKEN.match(text, pos)
        if not m:
            raise SyntaxError(f"cannot tokenize at {text[pos:pos + 20]!r} in {text!r}")
        pos = m.end()
        k = m.lastgroup
        v = m.group()
        if k == "ws":
            continue
        if k == "real":
            # "1.and." / "1.eq." style is not used by the reference; but guard "1.e" vs "1.eq."
            isint = re.fullmatch(r"\d+(?:_\w+)?", v) is not None
            toks.append(("int" if isint else "real", v.lower()))
        elif k == "dotop":
            toks.append(("op", v.lower()))
        elif k == "name":
            toks.append(("name", v.lower()))
        elif k == "str":
            toks.append(("str", v))
        else:
            toks.append(("op", v))
    # "(/" directly followed by "=" can never be an array constructor: it is "(" "/=" -- not used here.
    toks.append(("end", ""))
    return toks


# ------------------------------------------------------------------------------------------------
# declarations
# ------------------------------------------------------------------------------------------------


class TypeSpec:
    """base: 'int' | 'r4' | 'r8' | 'bool' | 'char' | 'type:<name>'; dims: None or list of (lb_src, ub_src)
    where ub_src may be ':' (deferred / assumed shape) or '*' (assumed size)."""

    def __init__(self, base, dims=None, intent=None, parameter=False, allocatable=False, optional=False,
                 init=None, kind_src=None):
        self.base, self.dims, self.intent = base, dims, intent
        self.parameter, self.allocatable, self.optional, self.init = parameter, allocatable, optional, init
        self.kind_src = kind_src

    @property
    def is_array(self):
        return self.dims is not None

    @property
    def is_struct(self):
        return self.base.startswith("type:")


_KIND_NAMES = {"sp": 4, "dp": 8, "real32": 4, "real64": 8, "4": 4, "8": 8, "fp": 4, "c_double": 8, "c_float": 4}


def split_top(s, sep=","):
    out, depth, cur, q = [], 0, "", None
    i = 0
    while i < len(s):
        c = s[i]
        if q:
            cur += c
            if c == q:
                q = None
        elif c in "'\"":
            q = c
            cur += c
        elif c in "([":
            depth += 1
            cur += c
        elif c in ")]":
            depth -= 1
            cur += c
        elif c == sep and depth == 0:
            out.append(cur.strip())
            cur = ""
        else:
            cur += c
        i += 1
    if cur.strip() or out:
        out.append(cur.strip())
    return out


_DECL_HEAD = re.compile(r"^((real|integer|logical|character|double\s+precision)\b|type\s*\()", re.I)


def is_declaration(text):
    t = text.lower()
    if not _DECL_HEAD.match(t):
        return False
    if t.startswith("type") and not re.match(r"^type\s*\(\s*\w+\s*\)", t):
        return False
    # "real(x)" cannot start an executable statement; "type(x) ::" is a declaration
    return True


def parse_declaration(text):
    """-> list of (name, TypeSpec)"""
    t = text.strip()
    low = t.lower()
    m = re.match(r"^(double\s+precision|real|integer|logical|character|type)\s*", low)
    base_kw = m.group(1)
    rest = t[m.end():]
    kind_src = None
    if rest.startswith("("):
        depth = 0
        for i, c in enumerate(rest):
            depth += c == "("
            depth -= c == ")"
            if depth == 0:
                break
        kind_src = rest[1:i].strip()
        rest = rest[i + 1:]
    elif rest.startswith("*"):
        m2 = re.match(r"\*\s*(\d+|\(\s*\*\s*\))", rest)
        kind_src = m2.group(1)
        rest = rest[m2.end():]
    if base_kw.startswith("double"):
        base = "r8"
    elif base_kw == "real":
        k = (kind_src or "sp").lower().replace("kind=", "").strip()
        if k not in _KIND_NAMES:
            raise NotImplementedError(f"real kind {kind_src!r}")
        base = "r4" if _KIND_NAMES[k] == 4 else "r8"
    elif base_kw == "integer":
        base = "int"
    elif base_kw == "logical":
        base = "bool"
    elif base_kw == "character":
        base = "char"
    else:
        base = "type:" + kind_src.lower()
    if "::" in rest:
        attrs_s, ents_s = rest.split("::", 1)
    else:
        attrs_s, ents_s = "", rest
    dims, intent, parameter, allocatable, optional = None, None, False, False, False
    for a in split_top(attrs_s):
        al = a.lower().strip()
        if not al:
            continue
        if al.startswith("dimension"):
            dims = parse_dims(a[a.index("(") + 1:a.rindex(")")])
        elif al.startswith("intent"):
            intent = re.sub(r"\s", "", al[al.index("(") + 1:al.rindex(")")])
        elif al == "parameter":
            parameter = True
        elif al in ("allocatable", "pointer"):
            allocatable = True
        elif al == "optional":
            optional = True
        elif al in ("save", "target", "private", "public", "value", "volatile", "contiguous") or al.startswith("bind"):
            pass
        else:
            raise NotImplementedError(f"attribute {a!r} in {text!r}")
    out = []
    for e in split_top(ents_s):
        init = None
        # entity [ (dims) ] [ *len ] [ = init ]
        m3 = re.match(r"^(\w+)\s*(\(.*?\))?\s*(?:\*\s*\w+\s*)?(?:=\s*(.*))?$", e, re.S)
        if not m3:
            raise SyntaxError(f"cannot parse entity {e!r} in {text!r}")
        # the lazy (\(.*?\)) may cut nested parentheses: redo with a depth scan
        name = m3.group(1).lower()
        tail = e[len(m3.group(1)):].lstrip()
        edims = dims
        if tail.startswith("("):
            depth = 0
            for i, c in enumerate(tail):
                depth += c == "("
                depth -= c == ")"
                if depth == 0:
                    break
            edims = parse_dims(tail[1:i])
            tail = tail[i + 1:].lstrip()
        if tail.startswith("*"):
            tail = re.sub(r"^\*\s*\w+\s*", "", tail)
        if tail.startswith("="):
            init = tail[1:].strip()
        out.append((name, TypeSpec(base, edims, intent, parameter, allocatable, optional, init, kind_src)))
    return out


def parse_dims(s):
    dims = []
    for d in split_top(s):
        if d == ":":
            dims.append((None, ":"))
        elif d == "*":
            dims.append(("1", "*"))
        else:
            parts = split_top(d, ":")
            if len(parts) == 1:
                dims.append(("1", parts[0]))
            else:
                dims.append((parts[0] or "1", parts[1] if parts[1] else ":"))
    return dims


# ------------------------------------------------------------------------------------------------
# program model
# ------------------------------------------------------------------------------------------------


class Proc:
    def __init__(self, module, kind, name, args, result, lines, path):
        self.module, self.kind, self.name, self.args, self.result = module, kind, name, args, result
        self.lines, self.path = lines, path
        self.decls = {}       # name -> TypeSpec
        self.uses = []        # (module, only-list or None)
        self.body = []        # executable logical lines
        self.out_scalars = []  # indices of scalar dummies copied back to the caller
        self._scan()

    def _scan(self):
        in_body = False
        for no, text in self.lines:
            low = text.lower()
            if not in_body:
                if low.startswith("use ") or low.startswith("use,"):
                    self.uses.append(parse_use(text))
                    continue
                if low.startswith("implicit") or low.startswith("save") or low.startswith("external"):
                    continue
                if is_declaration(text) and ("::" in text or not re.match(r"^\w+\s*\(.*\)\s*=", text)):
                    for n, ts in parse_declaration(text):
                        self.decls[n] = ts
                    continue
                in_body = True
            self.body.append((no, text))
        for i, a in enumerate(self.args):
            ts = self.decls.get(a)
            if ts is None:
                raise SyntaxError(f"{self.name}: dummy argument {a} is not declared")
            if not ts.is_array and not ts.is_struct and ts.intent != "in" and ts.base != "char":
                self.out_scalars.append(i)


def parse_use(text):
    m = re.match(r"^use\s*(?:,\s*\w+\s*)?(?:::)?\s*(\w+)\s*(?:,\s*only\s*:\s*(.*))?$", text, re.I | re.S)
    if not m:
        raise SyntaxError(f"cannot parse {text!r}")
    only = None
    if m.group(2) is not None:
        only = []
        for it in split_top(m.group(2)):
            if "=>" in it:
                loc, rem = [x.strip().lower() for x in it.split("=>")]
                only.append((loc, rem))
            elif it:
                only.append((it.lower(), it.lower()))
    return m.group(1).lower(), only


class Module:
    def __init__(self, name, path):
        self.name, self.path = name, path
        self.decls = {}    # name -> TypeSpec (variables and parameters)
        self.types = {}    # derived type name -> list of (field, TypeSpec)
        self.uses = []
        self.procs = {}
        self.space = ModSpace(name)


class Program:
    def __init__(self, defines=()):
        self.defines = set(defines)
        self.modules = {}
        self.externals = {}     # name -> python callable (subroutines return () or a tuple)
        self.ext_values = {}    # name -> value for names of modules that are not loaded (mpi, hdf5 ...)
        self.ext_out = {}       # external name -> indices of actual arguments that receive the returned tuple
        self.compiled = {}      # (module, proc) -> python function
        self.struct_classes = {}
        self.sources = {}       # (module, proc) -> generated python (for debugging)
        self._consts = {}
        self.glob = {"f4": f4, "f8": f8, "f_div": f_div, "f_pow": f_pow, "DoRange": DoRange, "FArray": FArray,
                     "Undefined": Undefined, "cv_int": cv_int, "cv_r4": cv_r4, "cv_r8": cv_r8, "cv_bool": cv_bool,
                     "cv_any": cv_any, "cv_char": cv_char, "cv_struct": cv_struct, "I": INTRINSICS, "i_arrcons": i_arrcons,
                     "FortranStop": FortranStop, "np": np, "_prog": self, "realloc_assign": realloc_assign}

    def const(self, value):
        """pooled literal: built once, referenced by name from the generated code"""
        key = (type(value).__name__, repr(value))
        name = self._consts.get(key)
        if name is None:
            name = f"K{len(self._consts)}"
            self._consts[key] = name
            self.glob[name] = value
        return name

    # -- loading ----------------------------------------------------------------------------------
    def load_file(self, path):
        lines = read_logical_lines(path, self.defines)
        i, n = 0, len(lines)
        while i < n:
            no, text = lines[i]
            m = re.match(r"^module\s+(\w+)\s*$", text, re.I)
            if m and not text.lower().startswith("module procedure"):
                mod = Module(m.group(1).lower(), path)
                self.modules[mod.name] = mod
                i = self._load_module(mod, lines, i + 1)
            else:
                i += 1  # program units other than modules are ignored

    def _load_module(self, mod, lines, i):
        n = len(lines)
        # specification part
        while i < n:
            no, text = lines[i]
            low = text.lower()
            if low == "contains":
                i += 1
                break
            if re.match(r"^end\s*module", low):
                return i + 1
            if low.startswith("use ") or low.startswith("use,"):
                mod.uses.append(parse_use(text))
            elif re.match(r"^type\s*(,.*)?(::)?\s*\w+\s*$", low) and not re.match(r"^type\s*\(", low):
                tname = re.match(r"^type\s*(?:,.*?)?(?:::)?\s*(\w+)\s*$", low).group(1)
                fields = []
                i += 1
                while not re.match(r"^end\s*type", lines[i][1].lower()):
                    ft = lines[i][1]
                    if is_declaration(ft):
                        fields.extend(parse_declaration(ft))
                    i += 1
                mod.types[tname] = fields
            elif low.startswith("interface"):
                while not re.match(r"^end\s*interface", lines[i][1].lower()):
                    i += 1
            elif is_declaration(text):
                for name, ts in parse_declaration(text):
                    mod.decls[name] = ts
            i += 1
        # procedures
        while i < n:
            no, text = lines[i]
            low = text.lower()
            if re.match(r"^end\s*module", low):
                return i + 1
            m = re.match(r"^(?:(?:recursive|pure|elemental)\s+)*(?:(?:real|integer|logical|double\s+precision|type)\s*(?:\([^)]*\))?\s+)?"
                         r"(subroutine|function)\s+(\w+)\s*(?:\((.*?)\))?\s*(?:result\s*\(\s*(\w+)\s*\))?\s*$", low)
            if not m:
                raise SyntaxError(f"{mod.path}:{no}: expected a procedure, got {text!r}")
            kind, name = m.group(1), m.group(2)
            args = [a.strip() for a in (m.group(3) or "").split(",") if a.strip()]
            result = m.group(4) or (name if kind == "function" else None)
            j = i + 1
            body = []
            while not re.match(rf"^end\s*{kind}\b", lines[j][1].lower()) and lines[j][1].lower() != "end":
                if lines[j][1].lower() == "contains":
                    raise NotImplementedError(f"{mod.path}:{lines[j][0]}: internal procedures")
                body.append(lines[j])
                j += 1
            try:
                mod.procs[name] = Proc(mod, kind, name, args, result, body, mod.path)
            except (SyntaxError, NotImplementedError) as e:
                mod.procs[name] = e  # only an error if somebody asks for this procedure
            i = j + 1
        return i

    # -- name resolution --------------------------------------------------------------------------
    def find_symbol(self, modname, name, seen=None):
        """-> ('var', module, name) | ('proc', module, name) | ('type', module, name) | None"""
        seen = seen if seen is not None else set()
        if modname in seen:
            return None
        seen.add(modname)
        mod = self.modules.get(modname)
        if mod is None:
            return None
        if name in mod.decls:
            return ("var", modname, name)
        if name in mod.procs:
            return ("proc", modname, name)
        if name in mod.types:
            return ("type", modname, name)
        for um, only in mod.uses:
            r = self._through_use(um, only, name, seen)
            if r:
                return r
        return None

    def _through_use(self, um, only, name, seen):
        if only is None:
            return self.find_symbol(um, name, seen)
        for loc, rem in only:
            if loc == name:
                return self.find_symbol(um, rem, seen)
        return None

    def resolve(self, proc, name):
        for um, only in proc.uses:
            r = self._through_use(um, only, name, set())
            if r:
                return r
        return self.find_symbol(proc.module.name, name)

    # -- module data ------------------------------------------------------------------------------
    def struct_class(self, modname, tname):
        key = (modname, tname)
        if key in self.struct_classes:
            return self.struct_classes[key]
        mod = self.modules[modname]
        fields, inits = {}, {}
        cls = type(tname, (FStruct,), {"__slots__": tuple(n for n, _ in mod.types[tname])})
        self.struct_classes[key] = cls
        for n, ts in mod.types[tname]:
            fields[n] = self.converter(ts, modname)
            if ts.is_array:
                inits[n] = (lambda ts=ts: self.make_array(ts, modname))
                fields[n] = cv_any
            elif ts.is_struct:
                sub = self.type_of(ts, modname)
                inits[n] = (lambda sub=sub: sub())
            elif ts.init is not None:
                val = fields[n](self.eval_const(ts.init, modname))
                inits[n] = (lambda val=val: val)
        cls._fields, cls._inits = fields, inits
        return cls

    def type_of(self, ts, modname):
        tname = ts.base[5:]
        r = self.find_symbol(modname, tname)
        if r is None or r[0] != "type":
            raise NameError(f"derived type {tname} not found from module {modname}")
        return self.struct_class(r[1], r[2])

    def converter(self, ts, modname=None):
        return {"int": cv_int, "r4": cv_r4, "r8": cv_r8, "bool": cv_bool, "char": cv_char}.get(ts.base, cv_struct)

    def np_dtype(self, ts):
        return {"int": np.int64, "r4": np.float32, "r8": np.float64, "bool": bool}.get(ts.base, object)

    def eval_const(self, src, modname):
        """evaluate a specification / initialisation expression in the scope of a module"""
        fake = _ScopeForModule(self, modname)
        code = ExprCompiler(self, fake).compile(src)
        return eval(code, self.glob_for(fake))

    def make_array(self, ts, modname, scope_eval=None):
        ev = scope_eval or (lambda s: self.eval_const(s, modname))
        bounds = []
        for lb, ub in ts.dims:
            if ub in (":", "*"):
                return None  # deferred: the harness allocates
            bounds.append((int(ev(lb)), int(ev(ub))))
        if ts.is_struct:
            cls = self.type_of(ts, modname)
            return FArray.alloc(object, bounds, fill=cls)
        return FArray.alloc(self.np_dtype(ts), bounds)

    def init_module_data(self, order=None):
        """create every module variable (parameters evaluated; scalars Undefined; explicit-shape arrays allocated)"""
        done = set()

        def visit(mn):
            if mn in done or mn not in self.modules:
                return
            done.add(mn)
            mod = self.modules[mn]
            for um, _ in mod.uses:
                visit(um)
            for name, ts in mod.decls.items():
                conv = self.converter(ts, mn)
                try:
                    self._declare_one(mod, mn, name, ts, conv)
                except (NameError, KeyError, SyntaxError, NotImplementedError):
                    mod.space.declare(name, cv_any, Undefined(f"{mn}::{name}"))
        for mn in (order or list(self.modules)):
            visit(mn)

    def _declare_one(self, mod, mn, name, ts, conv):
        if ts.is_array:
            val = None if ts.allocatable else self.make_array(ts, mn)
            if val is not None and ts.init is not None:
                val.assign(self.eval_const(ts.init, mn))
            mod.space.declare(name, cv_any, val)
        elif ts.is_struct:
            mod.space.declare(name, conv, self.type_of(ts, mn)())
        elif ts.init is not None:
            mod.space.declare(name, conv, conv(self.eval_const(ts.init, mn)))
        else:
            mod.space.declare(name, conv, Undefined(f"{mn}::{name}"))

    def allocate(self, modname, name, bounds):
        """what ALLOCATE(name(lb:ub, ...)) does, for the harness"""
        mod = self.modules[modname]
        ts = mod.decls[name]
        if ts.is_struct:
            arr = FArray.alloc(object, bounds, fill=self.type_of(ts, modname))
        else:
            arr = FArray.alloc(self.np_dtype(ts), bounds)
        object.__setattr__(mod.space, name, arr)
        return arr

    # -- compilation ------------------------------------------------------------------------------
    def glob_for(self, scope):
        g = dict(self.glob)
        for mn, mod in self.modules.items():
            g["M_" + mn] = mod.space
        g["X"] = self.ext_values
        g["E"] = self.externals
        g["C"] = self.compiled
        return g

    def get(self, modname, name):
        key = (modname, name)
        if key not in self.compiled:
            p = self.modules[modname].procs[name]
            if isinstance(p, Exception):
                raise p
            self.compiled[key] = None  # recursion guard: calls go through C[...] at run time
            src = ProcCompiler(self, p).compile()
            self.sources[key] = src
            g = self.glob_for(None)
            exec(compile(src, f"<f90:{modname}:{name}>", "exec"), g)
            self.compiled[key] = g["P_" + name]
        return self.compiled[key]


class _ScopeForModule:
    """name resolution for expressions that live in a module's specification part"""

    def __init__(self, prog, modname):
        self.prog, self.modname = prog, modname
        self.locals = {}

    def lookup(self, name):
        return self.prog.find_symbol(self.modname, name)


# ------------------------------------------------------------------------------------------------
# expression compiler (Fortran tokens -> Python source)
# ------------------------------------------------------------------------------------------------
_REL = {"==": "==", "/=": "!=", "<": "<", "<=": "<=", ">": ">", ">=": ">=",
        ".eq.": "==", ".ne.": "!=", ".lt.": "<", ".le.": "<=", ".gt.": ">", ".ge.": ">="}


class ExprCompiler:
    def __init__(self, prog, scope):
        self.prog, self.scope = prog, scope
        self.toks, self.i = None, 0

    # scope protocol: scope.locals: name -> TypeSpec ; scope.lookup(name) -> symbol tuple or None
    def compile(self, src):
        self.toks, self.i = tokenize(src), 0
        code = self.expr()
        if self.peek()[0] != "end":
            raise SyntaxError(f"trailing tokens {self.toks[self.i:]} in {src!r}")
        return code

    def compile_tokens(self, toks):
        self.toks, self.i = toks + [("end", "")], 0
        code = self.expr()
        if self.peek()[0] != "end":
            raise SyntaxError(f"trailing tokens {self.toks[self.i:]}")
        return code

    def peek(self, k=0):
        return self.toks[min(self.i + k, len(self.toks) - 1)]

    def next(self):
        t = self.toks[self.i]
        self.i += 1
        return t

    def accept(self, val):
        if self.peek()[1] == val and self.peek()[0] in ("op", "name"):
            self.i += 1
            return True
        return False

    def expect(self, val):
        if not self.accept(val):
            raise SyntaxError(f"expected {val!r}, got {self.peek()} in {self.toks}")

    def expr(self):
        left = self.or_()
        while self.peek()[1] in (".eqv.", ".neqv."):
            op = self.next()[1]
            right = self.or_()
            left = f"({left} {'==' if op == '.eqv.' else '!='} {right})"
        return left

    def or_(self):
        left = self.and_()
        while self.peek()[1] == ".or.":
            self.next()
            left = f"({left} or {self.and_()})"
        return left

    def and_(self):
        left = self.not_()
        while self.peek()[1] == ".and.":
            self.next()
            left = f"({left} and {self.not_()})"
        return left

    def not_(self):
        if self.peek()[1] == ".not.":
            self.next()
            return f"(not {self.not_()})"
        return self.rel()

    def rel(self):
        left = self.concat()
        if self.peek()[0] == "op" and self.peek()[1] in _REL:
            op = _REL[self.next()[1]]
            right = self.concat()
            return f"({left} {op} {right})"
        return left

    def concat(self):
        left = self.add()
        while self.peek() == ("op", "//"):
            self.next()
            left = f"({left} + {self.add()})"
        return left

    def add(self):
        if self.peek() == ("op", "-"):
            self.next()
            left = f"(-{self.mul()})"
        elif self.peek() == ("op", "+"):
            self.next()
            left = self.mul()
        else:
            left = self.mul()
        while self.peek() in (("op", "+"), ("op", "-")):
            op = self.next()[1]
            left = f"({left} {op} {self.mul()})"
        return left

    def mul(self):
        left = self.pow_()
        while self.peek() in (("op", "*"), ("op", "/")):
            op = self.next()[1]
            right = self.pow_()
            left = f"({left} * {right})" if op == "*" else f"f_div({left}, {right})"
        return left

    def pow_(self):
        base = self.primary()
        if self.peek() == ("op", "**"):
            self.next()
            if self.peek() == ("op", "-"):
                self.next()
                ex = f"(-{self.pow_()})"
            elif self.peek() == ("op", "+"):
                self.next()
                ex = self.pow_()
            else:
                ex = self.pow_()
            return f"f_pow({base}, {ex})"
        return base

    # -- literals ---------------------------------------------------------------------------------
    def real_literal(self, v):
        kind = 4
        m = re.fullmatch(r"([\d.]+(?:[ed][+-]?\d+)?)(?:_(\w+))?", v)
        num, suffix = m.group(1), m.group(2)
        if "d" in num:
            kind = 8
            num = num.replace("d", "e")
        if suffix is not None:
            if suffix not in _KIND_NAMES:
                raise NotImplementedError(f"kind suffix _{suffix}")
            kind = _KIND_NAMES[suffix]
        return self.prog.const(f4(num) if kind == 4 else f8(float(num)))

    def primary(self):
        k, v = self.next()
        if k == "int":
            return str(int(v.split("_")[0]))
        if k == "real":
            return self.real_literal(v)
        if k == "str":
            body = v[1:-1].replace(v[0] * 2, v[0])
            return repr(body)
        if (k, v) == ("op", ".true."):
            return "True"
        if (k, v) == ("op", ".false."):
            return "False"
        if (k, v) == ("op", "("):
            e = self.expr()
            self.expect(")")
            return f"({e})"
        if (k, v) in (("op", "(/"), ("op", "[")):
            close = "/)" if v == "(/" else "]"
            items = []
            if not self.accept(close):
                while True:
                    items.append(self.expr())
                    if self.accept(close):
                        break
                    self.expect(",")
            return f"i_arrcons([{', '.join(items)}])"
        if k == "name":
            if v == "z" and self.peek()[0] == "str":  # BOZ: Z'123'
                return repr(self.next()[1][1:-1])
            return self.designator(v)
        raise SyntaxError(f"unexpected token {(k, v)} in {self.toks}")

    # -- designators and calls --------------------------------------------------------------------------
    def arglist(self):
        """after '(' -> (positional codes, keyword codes, has_section)"""
        pos, kw, section = [], {}, False
        if self.accept(")"):
            return pos, kw, section
        while True:
            if self.peek()[0] == "name" and self.peek(1) == ("op", "="):
                name = self.next()[1]
                self.next()
                kw[name] = self.expr()
            else:
                code, is_sec = self.subscript()
                section |= is_sec
                pos.append(code)
            if self.accept(")"):
                break
            self.expect(",")
        return pos, kw, section

    def subscript(self):
        """expr | [lo]:[hi][:st]"""
        lo = hi = st = None
        if self.peek() != ("op", ":") and self.peek() != ("op", "::"):
            lo = self.expr()
            if self.peek() not in (("op", ":"), ("op", "::")):
                return lo, False
        if self.accept("::"):
            st = self.expr()
        else:
            self.expect(":")
            if self.peek()[1] not in (",", ")", ":") or self.peek()[0] != "op":
                hi = self.expr()
            if self.accept(":"):
                st = self.expr()
        return f"slice({lo}, {hi}, {st})", True

    def base_ref(self, name):
        """python expression for a bare name, and what it is"""
        loc = self.scope.locals.get(name)
        if loc is not None:
            return f"v_{name}", "local", loc
        sym = self.scope.lookup(name)
        if sym is not None:
            kind, mn, nm = sym
            if kind == "var":
                return f"M_{mn}.{nm}", "modvar", self.prog.modules[mn].decls[nm]
            if kind == "proc":
                return f"C[({mn!r}, {nm!r})]", "proc", self.prog.modules[mn].procs[nm]
        if name in INTRINSICS:
            return f"I[{name!r}]", "intrinsic", None
        if name in self.prog.externals:
            return f"E[{name!r}]", "external", None
        return f"X[{name!r}]", "extvalue", None

    def designator(self, name):
        code, what, info = self.base_ref(name)
        if what == "proc":
            self.prog.get(info.module.name, info.name) if not isinstance(info, Exception) else None
        first = True
        while True:
            if self.peek() == ("op", "("):
                self.next()
                pos, kw, section = self.arglist()
                if first and what in ("proc", "intrinsic", "external"):
                    if what == "proc" and not isinstance(info, Exception):
                        args = self.map_keywords(info, pos, kw)
                        code = f"{code}({', '.join(args)})"
                        if info.kind == "function" and info.out_scalars:
                            # a function that also sets scalar dummies returns (result, out1, ...): copy the
                            # outs back into the actuals (local scalars only) inside the expression
                            self.prog._tmp = getattr(self.prog, "_tmp", 0) + 1
                            t = f"_t{self.prog._tmp}"
                            parts = [f"({t} := {code})"]
                            for k, ai in enumerate(info.out_scalars):
                                src = args[ai] if ai < len(args) else "None"
                                m = re.fullmatch(r"v_(\w+)", src)
                                if m and m.group(1) in self.scope.locals:
                                    conv = _CONV.get(self.scope.locals[m.group(1)].base, "cv_any")
                                    parts.append(f"({src} := {conv}({t}[{k + 1}]))")
                                elif src != "None" and info.decls[info.args[ai]].intent != "in":
                                    if not re.fullmatch(r"[\w.()\[\]' ,+-]*", src) or "v_" in src or "M_" in src:
                                        raise NotImplementedError(
                                            f"function {info.name}: actual {src} of a non-intent(in) scalar dummy")
                            parts.append(f"{t}[0]")
                            code = "(" + ", ".join(parts) + ")[-1]"
                    else:
                        args = pos + [f"{k}={c}" for k, c in kw.items()]
                        code = f"{code}({', '.join(args)})"
                elif first and what == "extvalue":
                    args = pos + [f"{k}={c}" for k, c in kw.items()]
                    code = f"E[{name!r}]({', '.join(args)})"  # unknown name used as a function: must be a stub
                else:
                    if kw:
                        raise SyntaxError("keyword in array subscript")
                    code = f"{code}[{', '.join(pos)}]"
            elif self.peek() == ("op", "%"):
                self.next()
                comp = self.next()[1]
                code = f"{code}.{comp}"
            else:
                break
            first = False
        return code

    def map_keywords(self, proc, pos, kw):
        args = list(pos)
        if kw or len(args) < len(proc.args):
            for a in proc.args[len(args):]:
                args.append(kw.pop(a, "None"))
            if kw:
                raise SyntaxError(f"unknown keyword arguments {list(kw)} to {proc.name}")
        return args


# ------------------------------------------------------------------------------------------------
# procedure compiler
# ------------------------------------------------------------------------------------------------
_CONV = {"int": "cv_int", "r4": "cv_r4", "r8": "cv_r8", "bool": "cv_bool", "char": "cv_char"}


class ProcCompiler:
    def __init__(self, prog, proc):
        self.prog, self.proc = prog, proc
        self.locals = dict(proc.decls)
        self.ec = ExprCompiler(prog, self)
        self.out = []
        self.ind = 1
        self.tmp = 0

    def lookup(self, name):
        return self.prog.resolve(self.proc, name)

    def emit(self, s):
        self.out.append("    " * self.ind + s)

    def ret_stmt(self):
        p = self.proc
        if p.kind == "function":
            if p.out_scalars:
                return f"return (v_{p.result}, " + "".join(f"v_{p.args[i]}, " for i in p.out_scalars) + ")"
            return f"return v_{p.result}"
        return "return (" + "".join(f"v_{p.args[i]}, " for i in p.out_scalars) + ")"

    def compile(self):
        p = self.proc
        self.out.append(f"def P_{p.name}({', '.join('v_' + a + '=None' for a in p.args)}):")
        # dummy arrays take the bounds their declaration gives them
        for a in p.args:
            ts = p.decls[a]
            if ts.is_array:
                lbs = ", ".join(self.ec.compile(lb) if lb is not None else "1" for lb, _ in ts.dims)
                self.emit(f"v_{a} = _prog.as_dummy(v_{a}, ({lbs},), {len(ts.dims)})")
            elif ts.base == "char":
                self.emit(f"v_{a} = cv_char(v_{a})")
        for name, ts in p.decls.items():
            if name in p.args:
                continue
            if ts.parameter:
                conv = _CONV.get(ts.base, "cv_any")
                self.emit(f"v_{name} = {conv}({self.ec.compile(ts.init)})")
            elif ts.is_array:
                if ts.allocatable or any(ub in (":", "*") for _, ub in ts.dims):
                    self.emit(f"v_{name} = None")
                else:
                    b = ", ".join(f"({self.ec.compile(lb)}, {self.ec.compile(ub)})" for lb, ub in ts.dims)
                    if ts.is_struct:
                        cls = self.struct_ref(ts)
                        self.emit(f"v_{name} = FArray.alloc(object, [{b}], fill={cls})")
                    else:
                        dt = {"int": "np.int64", "r4": "np.float32", "r8": "np.float64", "bool": "bool"}[ts.base]
                        self.emit(f"v_{name} = FArray.alloc({dt}, [{b}])")
                    if ts.init is not None:
                        self.emit(f"v_{name}.assign({self.ec.compile(ts.init)})")
            elif ts.is_struct:
                self.emit(f"v_{name} = {self.struct_ref(ts)}()")
            elif ts.init is not None:
                self.emit(f"v_{name} = {_CONV.get(ts.base, 'cv_any')}({self.ec.compile(ts.init)})")
            else:
                self.emit(f"v_{name} = Undefined({(p.name + ':' + name)!r})")
        self.block(p.body, 0, len(p.body))
        self.emit(self.ret_stmt())
        return "\n".join(self.out) + "\n"

    def struct_ref(self, ts):
        tname = ts.base[5:]
        r = self.lookup(tname)
        if r is None or r[0] != "type":
            raise NameError(f"{self.proc.name}: derived type {tname} not found")
        self.prog.struct_class(r[1], r[2])
        return f"_prog.struct_classes[({r[1]!r}, {r[2]!r})]"

    # -- statements -------------------------------------------------------------------------------
    def block(self, lines, i, end):
        """compile lines[i:end]; emits at least 'pass'"""
        start_len = len(self.out)
        while i < end:
            i = self.statement(lines, i, end)
        if len(self.out) == start_len:
            self.emit("pass")

    def find_end(self, lines, i, end, open_re, close_re):
        depth = 0
        j = i
        while j < end:
            low = lines[j][1].lower()
            low = re.sub(r"^\w+\s*:\s*(?=(do|if)\b)", "", low)
            if open_re(low):
                depth += 1
            elif close_re(low):
                depth -= 1
                if depth == 0:
                    return j
            j += 1
        raise SyntaxError(f"{self.proc.name}: unterminated block starting at line {lines[i][0]}")

    @staticmethod
    def _is_if_then(low):
        return re.match(r"^if\s*\(.*\)\s*then$", low) is not None

    @staticmethod
    def _is_do(low):
        return re.match(r"^do(\s|$)", low) is not None

    def statement(self, lines, i, end):
        no, text = lines[i]
        low = text.lower()
        try:
            return self._statement(lines, i, end, no, text, low)
        except (SyntaxError, NotImplementedError, NameError) as e:
            raise type(e)(f"{self.proc.path}:{no}: {text!r}: {e}") from None

    def _statement(self, lines, i, end, no, text, low):
        # block IF
        if self._is_if_then(low):
            close = self.find_end(lines, i, end, self._is_if_then, lambda s: re.match(r"^end\s*if$", s) is not None)
            # split the branches at depth 1
            marks, depth = [i], 0
            for j in range(i, close + 1):
                lj = lines[j][1].lower()
                if self._is_if_then(lj):
                    depth += 1
                elif re.match(r"^end\s*if$", lj):
                    depth -= 1
                elif depth == 1 and (re.match(r"^else\s*if\s*\(.*\)\s*then$", lj) or lj == "else"):
                    marks.append(j)
            marks.append(close)
            for k in range(len(marks) - 1):
                head = lines[marks[k]][1]
                hl = head.lower()
                if k == 0:
                    cond = head[head.index("("):head.lower().rindex("then")].strip()
                    self.emit(f"if {self.ec.compile(cond)}:")
                elif hl == "else":
                    self.emit("else:")
                else:
                    cond = head[head.index("("):head.lower().rindex("then")].strip()
                    self.emit(f"elif {self.ec.compile(cond)}:")
                self.ind += 1
                self.block(lines, marks[k] + 1, marks[k + 1])
                self.ind -= 1
            return close + 1
        # DO loops
        if self._is_do(low):
            close = self.find_end(lines, i, end, self._is_do, lambda s: re.match(r"^end\s*do$", s) is not None)
            head = text[2:].strip()
            hl = head.lower()
            if not head:
                self.emit("while True:")
                self.ind += 1
                self.block(lines, i + 1, close)
                self.ind -= 1
            elif hl.startswith("while"):
                cond = head[5:].strip()
                self.emit(f"while {self.ec.compile(cond)}:")
                self.ind += 1
                self.block(lines, i + 1, close)
                self.ind -= 1
            else:
                var, rng = head.split("=", 1)
                var = var.strip().lower()
                parts = split_top(rng)
                self.tmp += 1
                it = f"_do{self.tmp}"
                self.emit(f"{it} = DoRange({', '.join(self.ec.compile(p) for p in parts)})")
                self.emit(f"for {self.target_name(var)} in {it}:")
                self.ind += 1
                self.block(lines, i + 1, close)
                self.ind -= 1
                self.emit("else:")
                self.emit(f"    {self.target_name(var)} = {it}.final")
            return close + 1
        # one-line IF
        m = re.match(r"^if\s*\(", low)
        if m:
            depth, j = 0, text.index("(")
            for j in range(text.index("("), len(text)):
                depth += text[j] == "("
                depth -= text[j] == ")"
                if depth == 0:
                    break
            cond, rest = text[text.index("("):j + 1], text[j + 1:].strip()
            self.emit(f"if {self.ec.compile(cond)}:")
            self.ind += 1
            self.statement([(no, rest)], 0, 1)
            self.ind -= 1
            return i + 1
        if low in ("exit",):
            self.emit("break")
            return i + 1
        if low in ("cycle",):
            self.emit("continue")
            return i + 1
        if low == "return":
            self.emit(self.ret_stmt())
            return i + 1
        if low == "continue":
            self.emit("pass")
            return i + 1
        if re.match(r"^(write|print)\b", low):
            self.io_write(text)
            return i + 1
        if re.match(r"^stop\b", low):
            self.emit(f"raise FortranStop({text!r})")
            return i + 1
        if re.match(r"^(open|close|flush)\b\s*\(", low):
            self.emit("pass  # " + low[:40].replace("\n", " "))
            return i + 1
        if re.match(r"^(read|rewind|backspace|inquire)\b\s*\(", low):
            self.emit(f"raise NotImplementedError({('I/O statement: ' + text[:60])!r})")
            return i + 1
        if low.startswith("call ") or low.startswith("call\t"):
            self.call_stmt(text[4:].strip())
            return i + 1
        m = re.match(r"^allocate\s*\((.*)\)$", text, re.I | re.S)
        if m:
            for item in split_top(m.group(1)):
                if re.match(r"^(stat|source|mold)\s*=", item, re.I):
                    raise NotImplementedError("allocate with stat/source")
                name = re.match(r"^(\w+)", item).group(1).lower()
                dims = parse_dims(item[item.index("(") + 1:item.rindex(")")])
                b = ", ".join(f"({self.ec.compile(lb)}, {self.ec.compile(ub)})" for lb, ub in dims)
                self.emit(f"{self.alloc_code(name, b)}")
            return i + 1
        m = re.match(r"^deallocate\s*\((.*)\)$", text, re.I | re.S)
        if m:
            for item in split_top(m.group(1)):
                name = item.strip().lower()
                if name in self.locals:
                    self.emit(f"v_{name} = None")
                else:
                    sym = self.lookup(name)
                    self.emit(f"object.__setattr__(M_{sym[1]}, {sym[2]!r}, None)")
            return i + 1
        # assignment
        self.assignment(text)
        return i + 1

    def io_write(self, text):
        """WRITE / PRINT: nothing is written; the values of the output list go to the hook E['__write__']
        (unit, [values]) when every item is an ordinary expression, so a harness can read what the
        reference would have put in its files."""
        m = re.match(r"^write\s*\(", text, re.I)
        try:
            if m:
                depth = 0
                for j in range(m.end() - 1, len(text)):
                    depth += text[j] == "("
                    depth -= text[j] == ")"
                    if depth == 0:
                        break
                ctl, items = split_top(text[m.end():j]), text[j + 1:].strip()
                unit = ctl[0] if ctl else "*"
                unit = re.sub(r"^unit\s*=\s*", "", unit, flags=re.I)
                unit_code = "'*'" if unit == "*" else self.ec.compile(unit)
            else:
                rest = text[5:].strip()
                parts = split_top(rest)
                unit_code, items = "'*'", ", ".join(parts[1:])
            codes = [self.ec.compile(it) for it in split_top(items)] if items else []
            self.emit(f"E['__write__']({unit_code}, [{', '.join(codes)}])")
        except (SyntaxError, NameError, NotImplementedError):
            self.emit("pass  # " + text[:40].lower().replace("\n", " "))

    def alloc_code(self, name, bounds):
        if name in self.locals:
            ts, tgt = self.locals[name], f"v_{name} = "
        else:
            sym = self.lookup(name)
            if sym is None or sym[0] != "var":
                raise NameError(f"allocate of unknown array {name}")
            ts = self.prog.modules[sym[1]].decls[sym[2]]
            tgt = None
        if ts.is_struct:
            rhs = f"FArray.alloc(object, [{bounds}], fill={self.struct_ref(ts)})"
        else:
            dt = {"int": "np.int64", "r4": "np.float32", "r8": "np.float64", "bool": "bool"}[ts.base]
            rhs = f"FArray.alloc({dt}, [{bounds}])"
        if tgt:
            return tgt + rhs
        return f"object.__setattr__(M_{sym[1]}, {sym[2]!r}, {rhs})"

    def target_name(self, var):
        if var in self.locals:
            return f"v_{var}"
        sym = self.lookup(var)
        if sym and sym[0] == "var":
            return f"M_{sym[1]}.{sym[2]}"
        raise NameError(f"assignment to unknown variable {var}")

    def split_assignment(self, text):
        depth, q = 0, None
        for j, c in enumerate(text):
            if q:
                if c == q:
                    q = None
            elif c in "'\"":
                q = c
            elif c in "([":
                depth += 1
            elif c in ")]":
                depth -= 1
            elif c == "=" and depth == 0:
                if text[j + 1:j + 2] == "=" or text[j - 1] in "<>/=":
                    continue
                if text[j + 1:j + 2] == ">":
                    raise NotImplementedError("pointer assignment")
                return text[:j].strip(), text[j + 1:].strip()
        raise SyntaxError("not an assignment")

    def lhs_parts(self, lhs):
        """-> (python code of the designator, last-part kind: 'name'|'index'|'comp', info)"""
        toks = tokenize(lhs)
        self.ec.toks, self.ec.i = toks, 0
        k, name = self.ec.next()
        if k != "name":
            raise SyntaxError(f"bad assignment target {lhs!r}")
        parts = []  # ('idx', [codes], section) | ('comp', name)
        while True:
            if self.ec.peek() == ("op", "("):
                self.ec.next()
                pos, kw, section = self.ec.arglist()
                parts.append(("idx", pos, section))
            elif self.ec.peek() == ("op", "%"):
                self.ec.next()
                parts.append(("comp", self.ec.next()[1]))
            else:
                break
        if self.ec.peek()[0] != "end":
            raise SyntaxError(f"bad assignment target {lhs!r}")
        return name, parts

    def store(self, lhs, rhs_code):
        """emit `lhs = rhs_code` with Fortran semantics"""
        name, parts = self.lhs_parts(lhs)
        if name in self.locals:
            ts, base, is_local = self.locals[name], f"v_{name}", True
        else:
            sym = self.lookup(name)
            if sym is None or sym[0] != "var":
                raise NameError(f"assignment to unknown variable {name}")
            ts, base, is_local = self.prog.modules[sym[1]].decls[sym[2]], f"M_{sym[1]}.{sym[2]}", False
        if not parts:
            if ts.is_array and ts.allocatable:
                # Fortran 2003 (re)allocation on assignment, gfortran's default: a shape mismatch gives the
                # left-hand side the shape of the right-hand side, lower bounds 1 for an expression
                dt = {"int": "np.int64", "r4": "np.float32", "r8": "np.float64", "bool": "bool"}.get(ts.base, "object")
                if is_local:
                    self.emit(f"{base} = realloc_assign({base}, {rhs_code}, {dt})")
                else:
                    self.emit(f"object.__setattr__(M_{sym[1]}, {sym[2]!r}, realloc_assign({base}, {rhs_code}, {dt}))")
            elif ts.is_array:
                self.emit(f"{base}.assign({rhs_code})")
            elif is_local:
                conv = _CONV.get(ts.base, "cv_struct")
                self.emit(f"{base} = {conv}({rhs_code})")
            else:
                self.emit(f"{base} = {rhs_code}")  # ModSpace converts
            return
        code = base
        for k, part in enumerate(parts[:-1]):
            code = f"{code}[{', '.join(part[1])}]" if part[0] == "idx" else f"{code}.{part[1]}"
        last = parts[-1]
        if last[0] == "idx":
            self.emit(f"{code}[{', '.join(last[1])}] = {rhs_code}")
        else:
            # array%comp = value (all elements) when the parent is a whole derived-type array
            whole_array = (len(parts) == 1 and ts.is_array)
            if whole_array:
                self.emit(f"{code}.set_component({last[1]!r}, {rhs_code})")
            elif len(parts) >= 2 and parts[-2][0] == "idx" and parts[-2][2]:
                self.emit(f"for _e in ({code}).reshape(-1): _e.{last[1]} = {rhs_code}")
            else:
                self.emit(f"{code}.{last[1]} = {rhs_code}")

    def assignment(self, text):
        lhs, rhs = self.split_assignment(text)
        self.store(lhs, self.ec.compile(rhs))

    def call_stmt(self, text):
        m = re.match(r"^(\w+)\s*(\((.*)\))?$", text, re.S)
        if not m:
            raise SyntaxError(f"cannot parse call {text!r}")
        name = m.group(1).lower()
        arg_srcs = split_top(m.group(3)) if m.group(3) and m.group(3).strip() else []
        sym = self.lookup(name)
        if sym and sym[0] == "proc":
            callee = self.prog.modules[sym[1]].procs[sym[2]]
            if isinstance(callee, Exception):
                raise callee
            self.prog.get(sym[1], sym[2])
            pos, kw = [], {}
            srcs = {}
            for a in arg_srcs:
                mk = re.match(r"^(\w+)\s*=(?!=)\s*(.*)$", a, re.S)
                if mk:
                    kw[mk.group(1).lower()] = mk.group(2)
                else:
                    pos.append(a)
            ordered = list(pos)
            for a in callee.args[len(pos):]:
                ordered.append(kw.pop(a, None))
            codes = [self.ec.compile(s) if s is not None else "None" for s in ordered]
            self.tmp += 1
            r = f"_r{self.tmp}"
            self.emit(f"{r} = C[({sym[1]!r}, {sym[2]!r})]({', '.join(codes)})")
            for k, ai in enumerate(callee.out_scalars):
                src = ordered[ai] if ai < len(ordered) else None
                if src is not None and self.is_variable(src):
                    self.store(src, f"{r}[{k}]")
            return
        codes = []
        for a in arg_srcs:
            mk = re.match(r"^(\w+)\s*=(?!=)\s*(.*)$", a, re.S)
            codes.append(f"{mk.group(1).lower()}={self.ec.compile(mk.group(2))}" if mk else self.ec.compile(a))
        outs = self.prog.ext_out.get(name)
        if outs:
            self.tmp += 1
            r = f"_r{self.tmp}"
            self.emit(f"{r} = E[{name!r}]({', '.join(codes)})")
            for k, ai in enumerate(outs):
                if ai < len(arg_srcs) and self.is_variable(arg_srcs[ai]):
                    name_, parts_ = self.lhs_parts(arg_srcs[ai])
                    ts_ = self.locals.get(name_)
                    if ts_ is not None and ts_.is_array and not parts_:
                        continue  # arrays are filled in place by the stub
                    self.store(arg_srcs[ai], f"{r}[{k}]")
        else:
            self.emit(f"E[{name!r}]({', '.join(codes)})")

    def is_variable(self, src):
        try:
            name, _ = self.lhs_parts(src)
        except SyntaxError:
            return False
        if name in self.locals:
            return not self.locals[name].parameter
        sym = self.lookup(name)
        return bool(sym and sym[0] == "var" and not self.prog.modules[sym[1]].decls[sym[2]].parameter)


def _as_dummy(self, val, lbs, rank):
    """give an actual argument the bounds of the dummy array it is associated with"""
    if val is None:
        return None
    a = val.a if isinstance(val, FArray) else val
    if not isinstance(a, np.ndarray):
        raise TypeError("scalar passed where the reference declares an array dummy")
    if a.ndim != rank:
        if rank == 1:  # sequence association: element order
            flat = a.reshape(-1, order="F")
            if a.size and not np.shares_memory(flat, a):
                raise NotImplementedError("non-contiguous actual for an explicit-shape dummy")
            a = flat
        else:
            raise NotImplementedError("rank-changing argument association")
    return FArray(a, lbs)


Program.as_dummy = _as_dummy
