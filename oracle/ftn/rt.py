"""Run-time support of the Fortran interpreter: arrays with arbitrary lower bounds in column-major order, the intrinsics the
reference uses with gfortran's evaluation order (sequential SUM / DOT_PRODUCT / MATMUL, integer powers by repeated
multiplication, libm sin/cos), and list-directed / formatted / stream I/O.  TEST INFRASTRUCTURE, see oracle/ftn/README.md."""
from __future__ import annotations

import copy
import math
import os
import re
import struct
import sys
import time as _time

import numpy as np

R8 = np.float64
R4 = np.float32
np.seterr(all="ignore")

DT = {"i4": np.int32, "i2": np.int16, "i8": np.int64, "r8": np.float64, "r4": np.float32, "l": np.bool_, "o": object}


class FortranStop(Exception):
    pass


class FS:
    """A Fortran subscript triplet lo:hi:st (None = omitted)."""
    __slots__ = ("lo", "hi", "st")

    def __init__(self, lo=None, hi=None, st=None):
        self.lo, self.hi, self.st = lo, hi, st


def _unwrap(v):
    return v.d if isinstance(v, FArray) else v


def _wrap(d):
    if isinstance(d, np.ndarray):
        if d.ndim == 0:
            return d[()]
        return FArray(d)
    return d


class FArray:
    """numpy array in Fortran order + lower bounds.  Scalars come out as numpy scalars (reals) or Python ints."""
    __slots__ = ("d", "lb")

    def __init__(self, d, lb=None):
        self.d = d
        self.lb = tuple(lb) if lb is not None else (1,) * d.ndim

    @staticmethod
    def new(code, bounds):
        shape = tuple(max(0, int(hi) - int(lo) + 1) for lo, hi in bounds)
        if code == "o":
            d = np.empty(shape, dtype=object, order="F")
        elif code.startswith("c"):
            d = np.empty(shape, dtype=object, order="F")
            d[...] = " " * int(code[1:] or 1)
        else:
            d = np.zeros(shape, dtype=DT[code], order="F")
        return FArray(d, tuple(int(lo) for lo, hi in bounds))

    # -- shape
    @property
    def shape(self):
        return self.d.shape

    @property
    def size(self):
        return self.d.size

    def ub(self, k):
        return self.lb[k] + self.d.shape[k] - 1

    # -- subscripts
    def _index(self, idx):
        if not isinstance(idx, tuple):
            idx = (idx,)
        if len(idx) != self.d.ndim:
            raise IndexError(f"rank mismatch: {len(idx)} subscripts for rank {self.d.ndim}")
        out, scalar, nvec = [], True, 0
        for k, s in enumerate(idx):
            lb = self.lb[k]
            if isinstance(s, FS):
                scalar = False
                lo = (s.lo if s.lo is not None else lb) - lb
                hi = (s.hi if s.hi is not None else lb + self.d.shape[k] - 1) - lb
                st = 1 if s.st is None else int(s.st)
                if st > 0:
                    out.append(slice(int(lo), int(hi) + 1, st) if hi >= lo else slice(0, 0))
                else:
                    stop = int(hi) - 1
                    out.append(slice(int(lo), stop if stop >= 0 else None, st) if lo >= hi else slice(0, 0))
            elif isinstance(s, FArray):
                scalar = False; nvec += 1
                out.append(s.d.astype(np.int64) - lb)
            elif isinstance(s, np.ndarray):
                scalar = False; nvec += 1
                out.append(s.astype(np.int64) - lb)
            else:
                i = int(s) - lb
                if i < 0 or i >= self.d.shape[k]:
                    raise IndexError(f"subscript {int(s)} of dimension {k + 1} outside [{lb},{self.ub(k)}]")
                out.append(i)
        if nvec > 1:
            # several vector subscripts: outer product indexing
            vec_pos = [k for k, o in enumerate(out) if isinstance(o, np.ndarray)]
            shapes = np.ix_(*[out[k] for k in vec_pos])
            for k, sh in zip(vec_pos, shapes):
                out[k] = sh
        return tuple(out), scalar

    def __getitem__(self, idx):
        t, scalar = self._index(idx)
        v = self.d[t]
        if scalar:
            if self.d.dtype.kind == "i":
                return int(v)
            if self.d.dtype.kind == "b":
                return bool(v)
            return v
        return FArray(v)

    def __setitem__(self, idx, val):
        t, scalar = self._index(idx)
        v = _unwrap(val)
        if self.d.dtype.kind in "iu" and isinstance(v, (float, np.floating)):
            v = int(v)
        elif self.d.dtype.kind in "iu" and isinstance(v, np.ndarray) and v.dtype.kind == "f":
            v = np.trunc(v)
        if self.d.dtype == object and scalar and isinstance(self.d[t], str) and isinstance(v, str):
            n = len(self.d[t])
            v = v[:n].ljust(n)
        elif self.d.dtype == object and not isinstance(v, (str, np.ndarray)) and v is not None and not scalar:
            # broadcasting one derived-type value over a section: each element gets its own copy
            sub = self.d[t]
            for ix in np.ndindex(sub.shape):
                sub[ix] = copy.deepcopy(v)
            return
        elif self.d.dtype == object and scalar and not isinstance(v, str):
            v = copy.deepcopy(v)
        self.d[t] = v

    def assign(self, val):
        """Whole-array assignment (shape-conforming array or scalar broadcast)."""
        self[tuple(FS() for _ in range(self.d.ndim))] = val

    def copy(self):
        return FArray(self.d.copy(order="F"), self.lb)

    def __deepcopy__(self, memo):
        if self.d.dtype == object:
            d = np.empty(self.d.shape, dtype=object, order="F")
            for ix in np.ndindex(self.d.shape):
                d[ix] = copy.deepcopy(self.d[ix], memo)
            return FArray(d, self.lb)
        return FArray(self.d.copy(order="F"), self.lb)

    def rebound(self, bounds):
        """Explicit-shape dummy argument: same storage, new bounds (sequence association when the shapes differ).
        bounds: list of (lo, hi) with hi None for an assumed-size last dimension."""
        shape = []
        for lo, hi in bounds:
            shape.append(None if hi is None else max(0, int(hi) - int(lo) + 1))
        if None in shape:
            known = 1
            for sv in shape:
                if sv is not None:
                    known *= sv
            shape[shape.index(None)] = self.d.size // max(known, 1)
        shape = tuple(shape)
        lb = tuple(int(lo) for lo, hi in bounds)
        if shape == self.d.shape:
            return FArray(self.d, lb)
        flat = self.d.reshape(-1, order="F") if self.d.flags.f_contiguous else None
        if flat is None or not np.shares_memory(flat, self.d):
            raise ValueError("sequence association needs contiguous storage")
        n = 1
        for sv in shape:
            n *= sv
        return FArray(flat[:n].reshape(shape, order="F"), lb)

    def relb1(self):
        """Assumed-shape dummy: lower bounds 1."""
        return FArray(self.d)

    # -- arithmetic (element-wise, Fortran semantics for integer division)
    def __add__(self, o): return FArray(self.d + _unwrap(o))
    def __radd__(self, o): return FArray(_unwrap(o) + self.d)
    def __sub__(self, o): return FArray(self.d - _unwrap(o))
    def __rsub__(self, o): return FArray(_unwrap(o) - self.d)
    def __mul__(self, o): return FArray(self.d * _unwrap(o))
    def __rmul__(self, o): return FArray(_unwrap(o) * self.d)
    def __neg__(self): return FArray(-self.d)
    def __pos__(self): return self


def is_int(v):
    return isinstance(v, (int, np.integer)) and not isinstance(v, (bool, np.bool_))


def _is_int_like(v):
    if isinstance(v, np.ndarray):
        return v.dtype.kind in "iu"
    return is_int(v)


def fdiv(a, b):
    """a / b: integer operands divide towards zero."""
    if type(a) is int and type(b) is int:
        q = abs(a) // abs(b)
        return q if (a >= 0) == (b >= 0) else -q
    ua, ub = _unwrap(a), _unwrap(b)
    if _is_int_like(ua) and _is_int_like(ub):
        q = np.abs(ua) // np.abs(ub)
        r = q * np.sign(ua) * np.sign(ub)
        return _wrap(r) if isinstance(r, np.ndarray) else int(r)
    return _wrap(ua / ub)


def _powi(x, n, cache=None):
    """x**n for a real x and an integer n > 0 as libgcc's __powidf2 computes it (square and multiply from the low bit): gfortran
    expands real**integer inline only for exponents -1..2 unless -funsafe-math-optimizations is given (trans-expr.c,
    gfc_conv_cst_int_power) and otherwise emits __builtin_powi, which the middle end leaves as the library call without fast-math.
    For n <= 4 this coincides with the multiplication chains of GCC's power tree; x**5 is x*((x*x)*(x*x))."""
    y = x if (n & 1) else None
    n >>= 1
    while n:
        x = x * x
        if n & 1:
            y = x if y is None else y * x
        n >>= 1
    return y


def fpow(a, b):
    ua, ub = _unwrap(a), _unwrap(b)
    if is_int(ub):
        n = int(ub)
        if _is_int_like(ua):
            if n >= 0:
                return _wrap(ua ** n)
            return _wrap(np.where(np.abs(ua) == 1, ua ** (-n), 0)) if isinstance(ua, np.ndarray) else (ua ** (-n) if abs(ua) == 1 else 0)
        if n == 0:
            return _wrap(ua * 0 + 1)
        r = _powi(ua, abs(n))
        return _wrap(r if n > 0 else 1.0 / r)
    return _wrap(np.power(ua, ub))


def cmp(op, a, b):
    ua, ub = _unwrap(a), _unwrap(b)
    if isinstance(ua, str) or isinstance(ub, str):
        n = max(len(ua), len(ub))
        ua, ub = ua.ljust(n), ub.ljust(n)
    if op == "eq": r = ua == ub
    elif op == "ne": r = ua != ub
    elif op == "lt": r = ua < ub
    elif op == "le": r = ua <= ub
    elif op == "gt": r = ua > ub
    else: r = ua >= ub
    return _wrap(r) if isinstance(r, np.ndarray) else bool(r)


def land(a, b):
    ua, ub = _unwrap(a), _unwrap(b)
    if isinstance(ua, np.ndarray) or isinstance(ub, np.ndarray):
        return _wrap(np.logical_and(ua, ub))
    return bool(ua) and bool(ub)


def lor(a, b):
    ua, ub = _unwrap(a), _unwrap(b)
    if isinstance(ua, np.ndarray) or isinstance(ub, np.ndarray):
        return _wrap(np.logical_or(ua, ub))
    return bool(ua) or bool(ub)


def lnot(a):
    ua = _unwrap(a)
    return _wrap(np.logical_not(ua)) if isinstance(ua, np.ndarray) else (not bool(ua))


# ---- conversions on assignment -----------------------------------------------------------------------------
def to_int(v):
    v = _unwrap(v)
    if isinstance(v, np.ndarray):
        return FArray(np.trunc(v).astype(np.int32) if v.dtype.kind == "f" else v)
    return int(v)


def to_r8(v):
    v = _unwrap(v)
    if isinstance(v, np.ndarray):
        return FArray(v.astype(np.float64))
    return np.float64(v)


def to_r4(v):
    v = _unwrap(v)
    if isinstance(v, np.ndarray):
        return FArray(v.astype(np.float32))
    return np.float32(v)


def to_char(v, n):
    if not isinstance(v, str):
        v = str(v)
    return v[:n].ljust(n)


def set_substr(s, lo, hi, v):
    """s(lo:hi) = v"""
    lo = 1 if lo is None else int(lo)
    hi = len(s) if hi is None else int(hi)
    n = hi - lo + 1
    if n <= 0:
        return s
    return s[:lo - 1] + v[:n].ljust(n) + s[hi:]


def substr(s, lo, hi):
    lo = 1 if lo is None else int(lo)
    hi = len(s) if hi is None else int(hi)
    return s[lo - 1:hi] if hi >= lo else ""


def array_cons(items):
    """Array constructor [a, b, c(:), ...] -> rank-1 array."""
    flat = []
    for it in items:
        u = _unwrap(it)
        if isinstance(u, np.ndarray):
            flat.extend(u.reshape(-1, order="F").tolist() if u.dtype == object else list(u.reshape(-1, order="F")))
        else:
            flat.append(u)
    if flat and all(isinstance(x, str) for x in flat):
        d = np.empty(len(flat), dtype=object)
        d[:] = flat
        return FArray(d)
    if flat and all(is_int(x) for x in flat):
        return FArray(np.array([int(x) for x in flat], dtype=np.int32))
    if flat and all(isinstance(x, (np.float32,)) or is_int(x) for x in flat):
        return FArray(np.array(flat, dtype=np.float32))
    return FArray(np.array(flat, dtype=np.float64))


# ---- intrinsics ----------------------------------------------------------------------------------------------
def _seqsum(d):
    """Sum in array element order, one addition after the other (what gfortran's inlined SUM does without -ffast-math)."""
    if d.size == 0:
        return d.dtype.type(0)
    if d.dtype.kind == "f":
        return np.add.accumulate(d.reshape(-1, order="F"))[-1]
    return int(d.sum())


def f_sum(a, dim=None, mask=None):
    d = _unwrap(a)
    if not isinstance(d, np.ndarray):
        return d
    if mask is not None:
        d = np.where(_unwrap(mask), d, 0)
    if dim is None:
        return _seqsum(d)
    ax = int(dim) - 1
    if d.dtype.kind == "f":
        return _wrap(np.take(np.add.accumulate(d, axis=ax), -1, axis=ax))
    return _wrap(d.sum(axis=ax))


def f_dot_product(a, b):
    return _seqsum(_unwrap(a) * _unwrap(b))


def f_matmul(a, b):
    A, B = _unwrap(a), _unwrap(b)
    dt = np.result_type(A.dtype, B.dtype)
    if A.ndim == 2 and B.ndim == 1:
        c = np.zeros(A.shape[0], dtype=dt)
        for j in range(A.shape[1]):
            c = c + A[:, j] * B[j]
        return FArray(c)
    if A.ndim == 1 and B.ndim == 2:
        c = np.zeros(B.shape[1], dtype=dt)
        for j in range(A.shape[0]):
            c = c + A[j] * B[j, :]
        return FArray(c)
    c = np.zeros((A.shape[0], B.shape[1]), dtype=dt, order="F")
    for j in range(A.shape[1]):
        c = c + np.multiply.outer(A[:, j], B[j, :])
    return FArray(np.asfortranarray(c))


def f_transpose(a):
    return FArray(np.asfortranarray(_unwrap(a).T))


def f_reshape(src, shape, *rest):
    s = _unwrap(src)
    shp = tuple(int(x) for x in _unwrap(shape).reshape(-1))
    return FArray(np.asfortranarray(s.reshape(-1, order="F")[: int(np.prod(shp))].reshape(shp, order="F")))


def f_abs(a):
    u = _unwrap(a)
    if isinstance(u, np.ndarray):
        return FArray(np.abs(u))
    return abs(u)


def f_sqrt(a):
    u = _unwrap(a)
    if isinstance(u, np.ndarray):
        return FArray(np.sqrt(u))
    if isinstance(u, np.float32):
        return np.sqrt(u)
    return np.float64(math.sqrt(u)) if u >= 0 else np.float64("nan")


def _libm(fn):
    def g(a):
        u = _unwrap(a)
        if isinstance(u, np.ndarray):
            return FArray(np.array([fn(float(x)) for x in u.reshape(-1, order="F")], dtype=u.dtype).reshape(u.shape, order="F"))
        r = fn(float(u))
        return np.float32(r) if isinstance(u, np.float32) else np.float64(r)
    return g


def _safe(fn):
    def h(x):
        try:
            return fn(x)
        except (ValueError, OverflowError):
            return float("nan")
    return h


f_sin, f_cos, f_tan = _libm(math.sin), _libm(math.cos), _libm(math.tan)
f_acos, f_asin, f_atan = _libm(_safe(math.acos)), _libm(_safe(math.asin)), _libm(math.atan)
f_exp, f_log = _libm(_safe(math.exp)), _libm(_safe(math.log))
f_log10 = _libm(_safe(math.log10))


def f_atan2(a, b):
    return np.float64(math.atan2(float(a), float(b)))


def f_floor(a):
    u = _unwrap(a)
    if isinstance(u, np.ndarray):
        return FArray(np.floor(u).astype(np.int32))
    return int(math.floor(u))


def f_ceiling(a):
    return int(math.ceil(_unwrap(a)))


def f_nint(a):
    u = _unwrap(a)
    if isinstance(u, np.ndarray):
        return FArray(np.where(u >= 0, np.floor(u + 0.5), -np.floor(-u + 0.5)).astype(np.int32))
    u = float(u)
    return int(math.floor(u + 0.5)) if u >= 0 else -int(math.floor(-u + 0.5))


def f_int(a, kind=None):
    u = _unwrap(a)
    if isinstance(u, np.ndarray):
        return FArray(np.trunc(u).astype(np.int32))
    return int(u)


def f_real(a, kind=None):
    u = _unwrap(a)
    t = np.float64 if (kind is not None and int(kind) == 8) else np.float32
    if isinstance(u, np.ndarray):
        return FArray(u.astype(t))
    return t(u)


def f_dble(a):
    u = _unwrap(a)
    if isinstance(u, np.ndarray):
        return FArray(u.astype(np.float64))
    return np.float64(u)


def f_mod(a, b):
    if is_int(a) and is_int(b):
        return int(math.fmod(int(a), int(b)))
    return np.float64(math.fmod(float(a), float(b)))


def f_modulo(a, b):
    if is_int(a) and is_int(b):
        return int(a) % int(b)
    return np.float64(float(a) - math.floor(float(a) / float(b)) * float(b))


def f_sign(a, b):
    r = abs(a)
    return r if b >= 0 else -r


def _minmax(fn, npfn, args):
    if any(isinstance(x, FArray) for x in args):
        r = _unwrap(args[0])
        for x in args[1:]:
            r = npfn(r, _unwrap(x))
        return _wrap(r)
    r = args[0]
    for x in args[1:]:
        if fn(x, r):
            r = x
    if any(isinstance(x, np.float64) for x in args):
        return np.float64(r)
    return r


def f_max(*args):
    return _minmax(lambda x, r: x > r, np.maximum, args)


def f_min(*args):
    return _minmax(lambda x, r: x < r, np.minimum, args)


def f_maxval(a, dim=None):
    d = _unwrap(a)
    if dim is not None:
        return _wrap(d.max(axis=int(dim) - 1))
    if d.size == 0:
        return -2147483647 - 1 if d.dtype.kind in "iu" else -np.finfo(d.dtype).max
    r = d.max()
    return int(r) if d.dtype.kind in "iu" else r


def f_minval(a, dim=None):
    d = _unwrap(a)
    if dim is not None:
        return _wrap(d.min(axis=int(dim) - 1))
    if d.size == 0:
        return 2147483647 if d.dtype.kind in "iu" else np.finfo(d.dtype).max
    r = d.min()
    return int(r) if d.dtype.kind in "iu" else r


def f_maxloc(a, dim=None):
    d = _unwrap(a)
    if d.ndim == 1:
        i = int(np.argmax(d)) + 1
        return i if dim is not None else FArray(np.array([i], dtype=np.int32))
    ix = np.unravel_index(int(np.argmax(d.reshape(-1, order="F"))), d.shape, order="F")
    return FArray(np.array([i + 1 for i in ix], dtype=np.int32))


def f_minloc(a, dim=None):
    d = _unwrap(a)
    if d.ndim == 1:
        i = int(np.argmin(d)) + 1
        return i if dim is not None else FArray(np.array([i], dtype=np.int32))
    ix = np.unravel_index(int(np.argmin(d.reshape(-1, order="F"))), d.shape, order="F")
    return FArray(np.array([i + 1 for i in ix], dtype=np.int32))


def f_size(a, dim=None):
    if dim is None:
        return int(a.d.size)
    return int(a.d.shape[int(dim) - 1])


def f_lbound(a, dim=None):
    return a.lb[int(dim) - 1] if dim is not None else FArray(np.array(a.lb, dtype=np.int32))


def f_ubound(a, dim=None):
    return a.ub(int(dim) - 1) if dim is not None else FArray(np.array([a.ub(k) for k in range(a.d.ndim)], dtype=np.int32))


def f_count(m):
    return int(np.count_nonzero(_unwrap(m)))


def f_any(m):
    return bool(np.any(_unwrap(m)))


def f_all(m):
    return bool(np.all(_unwrap(m)))


def f_isnan(a):
    u = _unwrap(a)
    if isinstance(u, np.ndarray):
        return FArray(np.isnan(u))
    return bool(u != u)


def f_trim(s):
    return s.rstrip(" ")


def f_adjustl(s):
    n = len(s)
    return s.lstrip(" ").ljust(n)


def f_adjustr(s):
    n = len(s)
    return s.rstrip(" ").rjust(n)


def f_len_trim(s):
    return len(s.rstrip(" "))


def f_len(s):
    return len(s)


def f_index(s, sub, back=None):
    return (s.rfind(sub) if back else s.find(sub)) + 1


def f_achar(i):
    return chr(int(i))


def f_iachar(c):
    return ord(c[0]) if c else 32


def f_cpu_time():
    return np.float64(_time.process_time())


def f_date_and_time_values():
    t = _time.localtime()
    ms = int((_time.time() % 1) * 1000)
    return FArray(np.array([t.tm_year, t.tm_mon, t.tm_mday, 0, t.tm_hour, t.tm_min, t.tm_sec, ms], dtype=np.int32))


def f_epsilon(a):
    return np.finfo(type(a)).eps if isinstance(a, (np.float32, np.float64)) else np.float64(np.finfo(np.float64).eps)


def f_huge(a):
    if is_int(a):
        return 2147483647
    return np.finfo(type(a)).max


def f_tiny(a):
    return np.finfo(type(a)).tiny


def deep(v):
    """Value copy for intrinsic assignment of derived types."""
    return copy.deepcopy(v)


# ---- I/O --------------------------------------------------------------------------------------------------------
class Unit:
    def __init__(self, path, form, access, fh, lines=None):
        self.path, self.form, self.access, self.fh = path, form, access, fh
        self.lines, self.pos = lines, 0      # formatted sequential reads work on the list of records
        self.pending = ""                    # formatted write with advance='no'


class IO:
    def __init__(self, cwd=".", echo=False):
        self.units = {}
        self.cwd = cwd
        self.echo = echo          # print what the program writes to unit * / 6
        self.stdout_lines = []

    def _path(self, name):
        name = name.strip()
        return name if os.path.isabs(name) else os.path.join(self.cwd, name)

    def open(self, unit, file=None, status=None, action=None, form=None, access=None, position=None, iostat=None):
        unit = int(unit)
        form = (form or "formatted").strip().lower()
        access = (access or "sequential").strip().lower()
        status = (status or "unknown").strip().lower()
        position = (position or "asis").strip().lower()
        path = self._path(file)
        if unit in self.units:
            self.close(unit)
        if status == "old" and not os.path.exists(path):
            if iostat is not None:
                return 2
            raise FortranStop(f"open: file not found: {path}")
        if form == "unformatted":
            mode = "r+b" if os.path.exists(path) and status != "replace" else "w+b"
            fh = open(path, mode)
            if position == "append":
                fh.seek(0, 2)
            self.units[unit] = Unit(path, form, access, fh)
        else:
            lines = []
            if os.path.exists(path) and status != "replace" and (action or "").strip().lower() != "write":
                with open(path, "r") as f:
                    lines = f.read().split("\n")
                if lines and lines[-1] == "":
                    lines.pop()
            u = Unit(path, form, access, None, lines)
            if position == "append":
                u.pos = len(lines)
            elif (action or "").strip().lower() != "read" and status in ("unknown", "new", "replace") and position != "append":
                # a file opened for writing from the start: earlier contents are overwritten record by record; keep it simple
                u.pos = 0
            u.dirty = False
            self.units[unit] = u
        return 0

    def close(self, unit, status=None):
        unit = int(unit)
        u = self.units.pop(unit, None)
        if u is None:
            return
        if u.form == "unformatted":
            u.fh.close()
        else:
            if getattr(u, "dirty", False):
                if u.pending:
                    u.lines[u.pos:] = [u.pending]; u.pos += 1
                with open(u.path, "w") as f:
                    f.write("\n".join(u.lines[: max(u.pos, 0)] if u.truncate_on_close else u.lines))
                    f.write("\n")

    def rewind(self, unit):
        u = self.units.get(int(unit))
        if u is None:
            return
        if u.form == "unformatted":
            u.fh.seek(0)
        else:
            u.pos = 0

    def flush(self, unit):
        pass

    def close_all(self):
        for k in list(self.units):
            self.close(k)

    # -- formatted input
    def read_line(self, unit):
        """Next record of a formatted unit; None at end of file."""
        u = self.units.get(int(unit))
        if u is None:
            raise FortranStop(f"read from unit {unit} which is not open")
        if u.pos >= len(u.lines):
            return None
        ln = u.lines[u.pos]
        u.pos += 1
        return ln

    # -- formatted output
    def write_line(self, unit, text, advance=True):
        if unit is None or int(unit) == 6:
            if advance:
                self.stdout_lines.append(text)
                if self.echo:
                    print(text)
            else:
                self.stdout_lines.append(text)
            return
        u = self.units.get(int(unit))
        if u is None:
            # gfortran would create fort.N; the reference never relies on that
            self.open(unit, file=f"fort.{int(unit)}")
            u = self.units[int(unit)]
        u.dirty = True
        u.truncate_on_close = True
        text = u.pending + text
        if not advance:
            u.pending = text
            return
        u.pending = ""
        for piece in text.split("\n"):
            if u.pos < len(u.lines):
                u.lines[u.pos] = piece
                del u.lines[u.pos + 1:]
            else:
                u.lines.append(piece)
            u.pos += 1

    # -- stream I/O
    def write_stream(self, unit, values):
        u = self.units[int(unit)]
        for v in values:
            u.fh.write(to_bytes(v))

    def read_stream(self, unit, code, count=None):
        """One item: code as in DT (or 'cN'); count None = scalar, else number of elements (returned as flat numpy array)."""
        u = self.units[int(unit)]
        if code.startswith("c"):
            n = int(code[1:] or 1)
            b = u.fh.read(n)
            if len(b) < n:
                raise EOFError
            return b.decode("latin1")
        dt = np.dtype(DT[code])
        n = 1 if count is None else int(count)
        b = u.fh.read(dt.itemsize * n)
        if len(b) < dt.itemsize * n:
            raise EOFError
        a = np.frombuffer(b, dtype=dt).copy()
        if count is None:
            v = a[0]
            return int(v) if dt.kind in "iu" else v
        return a


def to_bytes(v):
    if isinstance(v, FArray):
        return np.asfortranarray(v.d).tobytes(order="F")
    if isinstance(v, np.ndarray):
        return v.tobytes(order="F")
    if isinstance(v, (bool, np.bool_)):
        return struct.pack("<i", 1 if v else 0)
    if is_int(v):
        return struct.pack("<i", int(v))
    if isinstance(v, np.float32):
        return struct.pack("<f", float(v))
    if isinstance(v, (float, np.floating)):
        return struct.pack("<d", float(v))
    if isinstance(v, str):
        return v.encode("latin1")
    raise TypeError(f"cannot write {type(v)} to a stream")


# -- list-directed input
_tok_re = re.compile(r"""\s*(?:'((?:[^']|'')*)'|"((?:[^"]|"")*)"|([^\s,]+))\s*,?""")


def list_tokens(line):
    out, pos = [], 0
    s = line
    while pos < len(s):
        if s[pos:].strip() == "":
            break
        m = _tok_re.match(s, pos)
        if not m or m.end() == pos:
            break
        if m.group(1) is not None:
            out.append(("s", m.group(1).replace("''", "'")))
        elif m.group(2) is not None:
            out.append(("s", m.group(2).replace('""', '"')))
        else:
            tok = m.group(3)
            if tok.startswith("!"):   # gfortran does not treat '!' as a comment in list-directed input, but inputs may rely on '/' only
                pass
            rm = re.match(r"^(\d+)\*(.+)$", tok)
            if rm:
                out.extend([("t", rm.group(2))] * int(rm.group(1)))
            else:
                out.append(("t", tok))
        pos = m.end()
    return out


class ListReader:
    """List-directed READ from an internal file (a string) or from a formatted unit."""

    def __init__(self, io, src):
        self.io, self.src = io, src
        self.toks, self.k = [], 0
        self.internal = isinstance(src, str)
        self.first = True
        self.eof = False

    def _more(self):
        if self.internal:
            if not self.first:
                self.eof = True
                return False
            self.first = False
            self.toks, self.k = list_tokens(self.src), 0
            return True
        ln = self.io.read_line(self.src)
        if ln is None:
            self.eof = True
            return False
        self.toks, self.k = list_tokens(ln), 0
        return True

    def next(self, code):
        while self.k >= len(self.toks):
            if not self._more():
                raise EOFError
        kind, t = self.toks[self.k]
        self.k += 1
        if code.startswith("c"):
            return t
        if kind == "s":
            raise ValueError(f"list-directed read: string {t!r} where a number is expected")
        if code in ("i4", "i2", "i8"):
            try:
                return int(t)
            except ValueError:
                raise ValueError(f"Bad integer for item in list input: {t!r}")
        if code in ("r8", "r4"):
            tt = t.lower().replace("d", "e")
            if re.match(r"^[+-]?(\d+\.?\d*|\.\d+)[+-]\d+$", tt):   # 1.0-3
                tt = re.sub(r"([0-9.])([+-]\d+)$", r"\1e\2", tt)
            v = float(tt)
            return np.float32(v) if code == "r4" else np.float64(v)
        if code == "l":
            return t.lower().lstrip(".").startswith("t")
        raise ValueError(f"list-directed read of type {code}")


# -- formatted output
class Fmt:
    """A subset of Fortran format specifications: A[w] I[w[.m]] F/E/ES/D/G w.d, nX, '/', literals, repeat counts and groups,
    format reversion."""

    def __init__(self, spec):
        spec = spec.strip()
        if spec.startswith("(") and spec.endswith(")"):
            spec = spec[1:-1]
        self.items = self._parse(spec)

    def _parse(self, s):
        items, i, n = [], 0, len(s)
        while i < n:
            c = s[i]
            if c in " ,":
                i += 1; continue
            if c in "'\"":
                j = i + 1; buf = []
                while j < n:
                    if s[j] == c:
                        if j + 1 < n and s[j + 1] == c:
                            buf.append(c); j += 2; continue
                        break
                    buf.append(s[j]); j += 1
                items.append(("lit", "".join(buf))); i = j + 1; continue
            if c == "/":
                items.append(("nl",)); i += 1; continue
            if c == ":":
                items.append(("colon",)); i += 1; continue
            m = re.match(r"(\d+)?\(", s[i:])
            if m:
                rep = int(m.group(1) or 1)
                j = i + m.end(); depth = 1; k = j
                while k < n and depth:
                    if s[k] == "(": depth += 1
                    elif s[k] == ")": depth -= 1
                    k += 1
                items.append(("grp", rep, self._parse(s[j:k - 1]))); i = k; continue
            m = re.match(r"(\d+)[xX]", s[i:])
            if m:
                items.append(("x", int(m.group(1)))); i += m.end(); continue
            m = re.match(r"(\d+)?(ES|EN|[AIFEDGLaifedgl]|es|en)(\d+)?(?:\.(\d+))?(?:[eE](\d+))?", s[i:])
            if m:
                rep = int(m.group(1) or 1)
                items.append(("ed", rep, m.group(2).upper(), int(m.group(3)) if m.group(3) else None, int(m.group(4)) if m.group(4) else None))
                i += m.end(); continue
            m = re.match(r"[tT][lLrR]?(\d+)", s[i:])
            if m:
                items.append(("tab", int(m.group(1)))); i += m.end(); continue
            raise ValueError(f"format: cannot parse {s[i:]!r}")
        return items

    @staticmethod
    def _ed(code, w, d, v):
        if code == "A":
            sv = v if isinstance(v, str) else str(v)
            if w is None:
                return sv
            return sv[:w].rjust(w) if len(sv) < w else sv[:w]
        if code == "L":
            return ("T" if v else "F").rjust(w or 1)
        if code == "I":
            iv = int(v)
            sv = str(abs(iv))
            if d is not None:
                sv = sv.rjust(d, "0")
            if iv < 0:
                sv = "-" + sv
            if w is None or w == 0:
                return sv
            return sv.rjust(w) if len(sv) <= w else "*" * w
        fv = float(v)
        if code == "F":
            sv = f"{fv:.{d or 0}f}"
            if w == 0 or w is None:
                return sv
            if len(sv) > w and sv.startswith("0."):
                sv = sv[1:]
            elif len(sv) > w and sv.startswith("-0."):
                sv = "-" + sv[2:]
            return sv.rjust(w) if len(sv) <= w else "*" * w
        if code in ("E", "D", "G"):
            dd = d or 0
            if fv == 0 or dd == 0:
                digits, ex = "0" * dd, 0
            else:
                # dd significant digits, correctly rounded from the exact binary value (as the Fortran run-time library does)
                m = f"{abs(fv):.{dd - 1}e}"
                mant, e10 = m.split("e")
                digits, ex = mant.replace(".", ""), int(e10) + 1
            sign = "-" if (fv < 0 or (fv == 0 and math.copysign(1.0, fv) < 0)) else ""
            sv = f"{sign}0.{digits}{'D' if code == 'D' else 'E'}{'+' if ex >= 0 else '-'}{abs(ex):02d}"
            if w and len(sv) > w and sv.startswith("0."):
                sv = sv[1:]
            elif w and len(sv) > w and sv.startswith("-0."):
                sv = "-" + sv[2:]
            return sv.rjust(w or len(sv)) if (not w or len(sv) <= w) else "*" * w
        if code in ("ES", "EN"):
            sv = f"{fv:.{d or 0}E}"
            m = re.match(r"(.*)E([+-])(\d+)$", sv)
            sv = f"{m.group(1)}E{m.group(2)}{int(m.group(3)):02d}"
            return sv.rjust(w or len(sv)) if (not w or len(sv) <= w) else "*" * w
        return str(v)

    def render(self, values):
        vals = []
        for v in values:
            u = _unwrap(v)
            if isinstance(u, np.ndarray):
                vals.extend(u.reshape(-1, order="F").tolist() if u.dtype == object else list(u.reshape(-1, order="F")))
            else:
                vals.append(u)
        out = []
        state = {"k": 0, "stop": False}

        def run(items, top):
            for it in items:
                if state["stop"]:
                    return
                if it[0] == "lit":
                    out.append(it[1])
                elif it[0] == "nl":
                    out.append("\n")
                elif it[0] == "x":
                    out.append(" " * it[1])
                elif it[0] == "tab":
                    pass
                elif it[0] == "colon":
                    if state["k"] >= len(vals):
                        state["stop"] = True; return
                elif it[0] == "grp":
                    for _ in range(it[1]):
                        run(it[2], False)
                        if state["stop"]:
                            return
                elif it[0] == "ed":
                    for _ in range(it[1]):
                        if state["k"] >= len(vals):
                            state["stop"] = True; return
                        out.append(self._ed(it[2], it[3], it[4], vals[state["k"]]))
                        state["k"] += 1

        has_ed = any(self._has_ed(it) for it in self.items)
        first = True
        while first or (state["k"] < len(vals) and has_ed):
            if not first:
                out.append("\n")
                state["stop"] = False
            first = False
            run(self.items, True)
        return "".join(out)

    def _has_ed(self, it):
        if it[0] == "ed":
            return True
        if it[0] == "grp":
            return any(self._has_ed(x) for x in it[2])
        return False


_fmt_cache = {}


def format_values(fmt, values):
    f = _fmt_cache.get(fmt)
    if f is None:
        f = _fmt_cache[fmt] = Fmt(fmt)
    return f.render(values)


def list_directed(values):
    """Approximation of gfortran's list-directed output (diagnostics only; nothing the pins compare goes through it)."""
    out = []
    for v in values:
        u = _unwrap(v)
        if isinstance(u, np.ndarray):
            seq = u.reshape(-1, order="F").tolist()
        else:
            seq = [u]
        for x in seq:
            if isinstance(x, str):
                out.append(x)
            elif isinstance(x, (bool, np.bool_)):
                out.append(" T" if x else " F")
            elif is_int(x):
                out.append(f"{int(x):12d}")
            elif isinstance(x, np.float32):
                out.append(f"  {float(x):.8G}    ")
            else:
                out.append(f"  {float(x):.17G}     ")
    return " " + "".join(out)
