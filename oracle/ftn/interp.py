"""Evaluator of the small Fortran interpreter: executes the parse trees of oracle/ftn/parse.py with the run-time support of
oracle/ftn/rt.py.  TEST INFRASTRUCTURE, see oracle/ftn/README.md.

Every statement and expression is turned once into a Python closure taking the frame (a dict: variable name -> value) of the
running procedure.  Values: Python int (every integer kind), numpy float64 / float32 scalars, bool, str (fixed length, blank
padded), rt.FArray (arrays with bounds, column-major), Struct (derived types), None (unallocated allocatable).

Evaluation rules that matter for the pins:
  * operators are applied exactly as written, left to right within a precedence level, parentheses kept; no re-association,
    no fused multiply-add, no extended precision: every operation rounds to the type Fortran gives it (real(4) literals stay
    single until promoted by the other operand);
  * SUM / DOT_PRODUCT / MATMUL accumulate sequentially in array element order; x**n with integer n multiplies as libgcc's __powidf2
    does (what gfortran emits for real**integer beyond x**2 without fast-math);
  * !$OMP lines are comments, i.e. the program runs as its serial build (the order the C oracle restates);
  * scalar arguments are passed by copy-in / copy-out, arrays and derived types by reference (equivalent for a conforming
    program).
"""
from __future__ import annotations

import copy
import glob
import os

import numpy as np

from . import rt
from .parse import Node, parse_source
from .rt import FArray, FS, FortranStop

FArray.__array_ufunc__ = None      # numpy scalars defer to FArray.__radd__ & co.

F8 = np.float64
F4 = np.float32
EXIT, CYCLE, RETURN = 1, 2, 3


class _Absent:
    def __repr__(self):
        return "<absent optional argument>"


ABSENT = _Absent()


class InterpError(Exception):
    pass


class Struct:
    """A derived-type value: t = TypeRT, f = {component: value}."""
    __slots__ = ("t", "f")

    def __init__(self, t):
        self.t = t
        self.f = {}

    def __deepcopy__(self, memo):
        s = Struct(self.t)
        for k, v in self.f.items():
            s.f[k] = copy.deepcopy(v, memo) if isinstance(v, (FArray, Struct)) else v
        return s

    def __repr__(self):
        return f"<{self.t.name} {list(self.f)}>"


class VarInfo:
    __slots__ = ("name", "code", "clen", "tname", "dims", "alloc", "param", "optional", "intent", "save", "init", "dummy", "line", "value")

    def __init__(self, name, code, clen=None, tname=None, dims=None):
        self.name, self.code, self.clen, self.tname, self.dims = name, code, clen, tname, dims
        self.alloc = self.param = self.optional = self.save = self.dummy = self.value = False
        self.intent = None
        self.init = None
        self.line = 0


class TypeRT:
    def __init__(self, node, module, scope):
        self.name, self.node, self.module, self.scope = node.name, node, module, scope
        self.bindings = dict(node.bindings)
        self.info = None          # component name -> VarInfo, filled on first use


class ModuleRT:
    def __init__(self, node):
        self.name, self.node = node.name, node
        self.vars, self.info, self.types, self.procs = {}, {}, {}, {}
        self.uses = list(node.uses)
        self.ready = False
        self.scope = None


class ProcRT:
    def __init__(self, node, module, host):
        self.node, self.module, self.host = node, module, host
        self.name, self.kind = node.name, node.kind
        self.internal = {}
        self.compiled = False
        self.argnames = list(node.args)
        self.resname = None
        self.scope = None
        self.entry = self.body = None
        self.ncalls = 0
        self.cname = getattr(node, "cname", None)      # bind(C) interface: the procedure lives in the bound C library


# ---------------------------------------------------------------------------------------------------------------------
def conv_like(old, new):
    """Value `new` converted for storage in a scalar variable currently holding `old` (intrinsic assignment)."""
    t = type(old)
    if t is F8:
        return new if type(new) is F8 else F8(new)
    if t is int:
        if type(new) is int:
            return new
        if isinstance(new, (bool, np.bool_)):
            raise InterpError("logical assigned to integer")
        return int(new)
    if t is F4:
        return new if type(new) is F4 else F4(new)
    if t is bool:
        return bool(new)
    if t is str:
        n = len(old)
        if not isinstance(new, str):
            raise InterpError(f"non-character value {new!r} assigned to character variable")
        return new if len(new) == n else new[:n].ljust(n)
    if t is Struct:
        if not isinstance(new, Struct):
            raise InterpError(f"{type(new)} assigned to a derived-type variable")
        return copy.deepcopy(new)
    raise InterpError(f"conv_like: unsupported target {t} <- {type(new)}")


def conv_code(code, clen, v):
    """Scalar value converted to the declared type `code`."""
    if code == "r8":
        return v if type(v) is F8 else F8(v)
    if code in ("i4", "i2", "i8"):
        return v if type(v) is int else int(v)
    if code == "r4":
        return F4(v)
    if code == "l":
        return bool(v)
    if code == "c":
        return v if clen is None or len(v) == clen else v[:clen].ljust(clen)
    return v


def scalar_like(v):
    return not isinstance(v, (FArray, Struct)) and v is not None and v is not ABSENT


class Scope:
    """Compile-time name resolution of one procedure (or of a module's specification part when proc is None)."""

    def __init__(self, interp, module, proc=None, host=None):
        self.interp, self.module, self.proc, self.host = interp, module, proc, host
        self.info = {}
        self.static = {}
        self.uses = []
        self.types = {}
        self._mods = None

    def modules(self):
        if self._mods is None:
            seen, out = set(), []

            def add(name):
                m = self.interp.modules.get(name)
                if m is None or m.name in seen:
                    return
                seen.add(m.name)
                self.interp.ready(m)
                out.append(m)
                for u in m.uses:
                    add(u)
            for u in self.uses:
                add(u)
            if self.module is not None:
                if self.module.name not in seen:
                    seen.add(self.module.name)
                    out.insert(0, self.module)
                for u in self.module.uses:
                    add(u)
            if self.host is not None:
                for m in self.host.modules():
                    if m.name not in seen:
                        seen.add(m.name); out.append(m)
            self._mods = out
        return self._mods

    def lookup_var(self, name):
        """-> ('local', info) | ('static', dict, info) | ('host', depth, info) | ('module', dict, info) | None"""
        i = self.info.get(name)
        if i is not None:
            if self.proc is None:
                return ("module", self.module.vars, i)
            if i.save or i.param:
                return ("static", self.static, i)
            return ("local", i)
        if self.host is not None:
            r = self.host.lookup_var(name)
            if r is not None:
                if r[0] == "local":
                    return ("host", 1, r[1])
                if r[0] == "host":
                    return ("host", r[1] + 1, r[2])
                return r
        for m in self.modules():
            i = m.info.get(name)
            if i is not None:
                return ("module", m.vars, i)
        return None

    def lookup_proc(self, name):
        p = self.proc
        while p is not None:
            if name in p.internal:
                return p.internal[name]
            p = p.host
        for m in self.modules():
            if name in m.procs:
                return m.procs[name]
        return self.interp.externals.get(name)

    def lookup_type(self, name):
        if name in self.types:
            return self.types[name]
        if self.host is not None:
            t = self.host.types.get(name)
            if t is not None:
                return t
        for m in self.modules():
            if name in m.types:
                return m.types[name]
        raise InterpError(f"unknown derived type {name}")


# ---------------------------------------------------------------------------------------------------------------------
class Interp:
    def __init__(self, cwd=".", echo=False):
        self.modules = {}
        self.externals = {}
        self.programs = {}
        self.io = rt.IO(cwd, echo)
        self.forked = False
        self.stats = {}
        self.clib = None           # ctypes library that bind(C) interfaces resolve into (bind_c_library)
        self.builtin_subs = {
            "omp_set_num_threads": lambda *a, **k: None,
            "mywait": lambda *a, **k: None,
            "flush": lambda *a, **k: None,
        }

    # -- loading
    ISO_C_BINDING = """
module iso_c_binding
    implicit none
    integer, parameter :: c_int = 4, c_short = 2, c_long = 8, c_long_long = 8, c_size_t = 8, c_double = 8, c_float = 4, c_char = 1, c_bool = 1
    character(len=1), parameter :: c_null_char = achar(0)
    type c_ptr
        integer(8) :: addr = 0
    end type
    type(c_ptr) :: c_null_ptr
end module iso_c_binding
"""

    def bind_c_library(self, lib):
        """bind(C) interfaces call into this library (a path or a ctypes.CDLL); iso_c_binding becomes available."""
        import ctypes
        self.clib = ctypes.CDLL(lib) if isinstance(lib, str) else lib
        if "iso_c_binding" not in self.modules:
            self.load_text(self.ISO_C_BINDING, "<iso_c_binding>")
        return self

    def load(self, paths):
        for p in paths:
            with open(p) as f:
                self.load_text(f.read(), p)
        return self

    def load_text(self, text, fname):
        if True:
            units = parse_source(text, fname)
            for u in units:
                if u.t == "module":
                    m = ModuleRT(u)
                    self.modules[m.name] = m
                    for pn in u.procs:
                        m.procs[pn.name] = self._mkproc(pn, m, None)
                    for pn in getattr(u, "interfaces", []):
                        m.procs[pn.name] = self._mkproc(pn, m, None)
                elif u.kind == "program":
                    self.programs[u.name] = self._mkproc(u, None, None)
                else:
                    self.externals[u.name] = self._mkproc(u, None, None)
        return self

    def _mkproc(self, node, module, host):
        p = ProcRT(node, module, host)
        for q in node.procs:
            p.internal[q.name] = self._mkproc(q, module, p)
        return p

    def ready(self, m):
        """Specification part of a module: types, parameters, variables (in source order)."""
        if m.ready:
            return
        m.ready = True
        sc = m.scope = Scope(self, m)
        for u in m.uses:
            if u in self.modules:
                self.ready(self.modules[u])
        for td in m.node.types:
            m.types[td.name] = TypeRT(td, m, sc)
        cp = Compiler(self, sc)
        for d in m.node.decls:
            for ent in d.ents:
                info = cp.varinfo(d, ent)
                m.info[info.name] = info
                sc.info[info.name] = info
                m.vars[info.name] = cp.initial_value(info, {})

    # -- running
    def run_program(self, name=None):
        p = self.programs[name] if name else next(iter(self.programs.values()))
        try:
            self.invoke(p, [], None, None)
        except FortranStop as e:
            self.io.close_all()
            return str(e)
        self.io.close_all()
        return None

    def call(self, name, *args, module=None):
        p = self.modules[module].procs[name] if module else self.externals[name]
        fr = self.invoke(p, list(args), None, None)
        return fr

    def compile_proc(self, p):
        p.compiled = True
        node = p.node
        host_scope = None
        if p.host is not None:
            if not p.host.compiled:
                self.compile_proc(p.host)
            host_scope = p.host.scope
        sc = p.scope = Scope(self, p.module, p, host_scope)
        sc.uses = list(node.uses)
        for td in node.types:
            sc.types[td.name] = TypeRT(td, p.module, sc)
        cp = Compiler(self, sc)
        for d in node.decls:
            for ent in d.ents:
                info = cp.varinfo(d, ent)
                sc.info[info.name] = info
        for a in p.argnames:
            if a not in sc.info:
                raise InterpError(f"{p.name}: dummy argument {a} is not declared")
            sc.info[a].dummy = True
        if p.kind == "function":
            p.resname = node.result or node.name
            if p.resname not in sc.info:
                if node.rtype is None:
                    raise InterpError(f"function {p.name}: result type unknown")
                sc.info[p.resname] = cp.varinfo(Node("decl", spec=node.rtype, attrs={}, ents=[], line=node.line),
                                                 Node("entity", name=p.resname, dims=None, clen=None, init=None))
        p.entry = cp.entry_code(p)
        p.body = cp.block(node.body)

    def invoke(self, p, args, kwargs, host_fr):
        if not p.compiled:
            self.compile_proc(p)
        p.ncalls += 1
        fr = {}
        if p.host is not None:
            fr["$host"] = host_fr
        names = p.argnames
        n = len(args)
        if n > len(names):
            raise InterpError(f"{p.name}: {n} arguments for {len(names)} dummies")
        for i, nm in enumerate(names):
            fr[nm] = args[i] if i < n else ABSENT
        if kwargs:
            for k, v in kwargs.items():
                if k not in fr:
                    raise InterpError(f"{p.name}: no dummy argument {k}")
                fr[k] = v
        if p.cname is not None:
            from . import cbind
            return cbind.call(self, p, fr)
        p.entry(fr)
        p.body(fr)
        return fr


# ---------------------------------------------------------------------------------------------------------------------
class Compiler:
    def __init__(self, interp, scope):
        self.I, self.sc = interp, scope

    # ---- declarations
    def const(self, e, fr=None):
        return self.expr(e)(fr if fr is not None else {})

    def varinfo(self, d, ent):
        spec = d.spec
        base = spec.base
        kind = None
        if spec.kind is not None and spec.kind.t != "star":
            kind = int(self.const(spec.kind))
        if base == "integer":
            code = {None: "i4", 4: "i4", 2: "i2", 8: "i8", 1: "i2"}[kind]
        elif base == "real":
            code = "r8" if kind == 8 else "r4"
        elif base == "logical":
            code = "l"
        elif base == "character":
            code = "c"
        elif base == "type":
            code = "t"
        else:
            raise InterpError(f"unsupported type {base}")
        clen = None
        if code == "c":
            clen = ent.clen if ent.clen is not None else spec.len
            if clen is None:
                clen = Node("num", k="i", v="1")
        dims = ent.dims if ent.dims is not None else d.attrs.get("dimension")
        info = VarInfo(ent.name, code, clen, spec.tname, dims)
        info.alloc = bool(d.attrs.get("allocatable")) or bool(d.attrs.get("pointer"))
        info.param = bool(d.attrs.get("parameter"))
        info.optional = bool(d.attrs.get("optional"))
        info.value = bool(d.attrs.get("value"))
        info.intent = d.attrs.get("intent")
        info.save = bool(d.attrs.get("save")) or (ent.init is not None)
        info.init = ent.init
        info.line = d.line
        return info

    def clen_of(self, info, fr):
        if info.clen is None:
            return None
        if info.clen.t == "star":
            return None
        return int(self.expr(info.clen)(fr))

    def bounds_of(self, info, fr):
        out = []
        for lo, hi in info.dims:
            l = 1 if lo is None else int(self.expr(lo)(fr))
            h = None if (hi is None or hi.t == "star") else int(self.expr(hi)(fr))
            out.append((l, h))
        return out

    def scalar_default(self, info, fr):
        c = info.code
        if c == "r8":
            return F8(0.0)
        if c in ("i4", "i2", "i8"):
            return 0
        if c == "r4":
            return F4(0.0)
        if c == "l":
            return False
        if c == "c":
            n = self.clen_of(info, fr)
            return " " * (n if n is not None else 1)
        if c == "t":
            return self.instantiate(self.sc.lookup_type(info.tname))
        raise InterpError(c)

    def new_array(self, info, bounds, fr):
        c = info.code
        if c == "c":
            n = self.clen_of(info, fr)
            return FArray.new("c%d" % (n or 1), bounds)
        if c == "t":
            a = FArray.new("o", bounds)
            td = self.sc.lookup_type(info.tname)
            flat = a.d.reshape(-1, order="F") if a.d.size else a.d
            for k in range(a.d.size):
                flat[k] = self.instantiate(td)
            if a.d.size and not np.shares_memory(flat, a.d):
                a.d[...] = flat.reshape(a.d.shape, order="F")
            return a
        return FArray.new(c, bounds)

    def initial_value(self, info, fr):
        """Value of a variable at the start of its life (module variable, saved local, component)."""
        if info.dims is not None:
            deferred = any(hi is None for lo, hi in info.dims)
            if info.alloc or deferred:
                if info.init is None:
                    return None
                v = self.const(info.init, fr)
                if not isinstance(v, FArray):
                    raise InterpError(f"{info.name}: array initialiser expected")
                return v
            a = self.new_array(info, self.bounds_of(info, fr), fr)
            if info.init is not None:
                a.assign(self.const(info.init, fr))
            return a
        v = self.scalar_default(info, fr)
        if info.init is not None:
            v = conv_like(v, self.const(info.init, fr))
        return v

    def instantiate(self, td):
        if td.info is None:
            td.info = {}
            cp = Compiler(self.I, td.scope)
            td.cp = cp
            for d in td.node.comps:
                for ent in d.ents:
                    td.info[ent.name] = cp.varinfo(d, ent)
        s = Struct(td)
        cp = td.cp
        for name, info in td.info.items():
            s.f[name] = cp.initial_value(info, {})
        return s

    # ---- procedure entry: dummies re-bounded, locals created
    def entry_code(self, p):
        sc = self.sc
        steps = []
        consts = {}
        for name, info in sc.info.items():
            if info.dummy:
                if info.dims is not None:
                    steps.append(self.bind_array_dummy(name, info))
                elif info.code == "c" and info.clen is not None and info.clen.t != "star":
                    steps.append(self.bind_char_dummy(name, info))
                continue
            if info.param or info.save:
                sc.static[name] = self.initial_value(info, sc.static)
                continue
            if info.dims is not None:
                deferred = any(hi is None for lo, hi in info.dims)
                if info.alloc or deferred:
                    consts[name] = None
                else:
                    steps.append(self.local_array(name, info))
            elif info.code == "t":
                steps.append(self.local_struct(name, info))
            elif info.code == "c":
                if info.clen.t == "num":
                    consts[name] = " " * int(info.clen.v)
                else:
                    steps.append(self.local_char(name, info))
            else:
                consts[name] = self.scalar_default(info, {})

        def entry(fr):
            fr.update(consts)
            for s in steps:
                s(fr)
        return entry

    def bind_array_dummy(self, name, info):
        dims = info.dims
        assumed_shape = all(hi is None for lo, hi in dims)
        los = [None if lo is None else self.expr(lo) for lo, hi in dims]
        his = [None if (hi is None or hi.t == "star") else self.expr(hi) for lo, hi in dims]
        alloc = info.alloc
        pname = self.sc.proc.name

        def bind(fr):
            v = fr[name]
            if v is ABSENT or v is None:
                return
            if not isinstance(v, FArray):
                raise InterpError(f"{pname}: scalar actual argument for array dummy {name} (sequence association from an element is not supported)")
            if alloc:
                return
            if assumed_shape:
                lb = tuple(1 if l is None else int(l(fr)) for l in los)
                if len(lb) != v.d.ndim:
                    raise InterpError(f"{pname}: rank mismatch for assumed-shape dummy {name}")
                if v.lb != lb:
                    fr[name] = FArray(v.d, lb)
                return
            b = [(1 if l is None else int(l(fr)), None if h is None else int(h(fr))) for l, h in zip(los, his)]
            fr[name] = v.rebound(b)
        return bind

    def bind_char_dummy(self, name, info):
        le = self.expr(info.clen)

        def bind(fr):
            v = fr[name]
            if isinstance(v, str):
                n = int(le(fr))
                if len(v) != n:
                    fr[name] = v[:n].ljust(n)
        return bind

    def local_array(self, name, info):
        los = [None if lo is None else self.expr(lo) for lo, hi in info.dims]
        his = [self.expr(hi) for lo, hi in info.dims]
        const_bounds = all(self.is_const(e) for e in [lo for lo, hi in info.dims if lo is not None] + [hi for lo, hi in info.dims])
        if info.code in ("r8", "r4", "i4", "i2", "i8", "l") and const_bounds:
            b = [(1 if l is None else int(l({})), int(h({}))) for l, h in zip(los, his)]
            proto = FArray.new(info.code, b)

            def mk(fr):
                fr[name] = FArray(np.zeros(proto.d.shape, dtype=proto.d.dtype, order="F"), proto.lb)
            return mk

        def mk2(fr):
            b = [(1 if l is None else int(l(fr)), int(h(fr))) for l, h in zip(los, his)]
            fr[name] = self.new_array(info, b, fr)
        return mk2

    def is_const(self, e):
        """Literal, or named constant known at compile time (conservative)."""
        if e.t == "num":
            return True
        if e.t == "un":
            return self.is_const(e.e)
        if e.t == "paren":
            return self.is_const(e.e)
        if e.t == "bin":
            return self.is_const(e.l) and self.is_const(e.r)
        if e.t == "desig" and len(e.parts) == 1 and e.parts[0].args is None:
            r = self.sc.lookup_var(e.parts[0].name)
            return r is not None and r[-1].param
        return False

    def local_struct(self, name, info):
        td = self.sc.lookup_type(info.tname)

        def mk(fr):
            fr[name] = self.instantiate(td)
        return mk

    def local_char(self, name, info):
        le = self.expr(info.clen)

        def mk(fr):
            fr[name] = " " * int(le(fr))
        return mk

    # ---- variables
    def var_access(self, name):
        """-> (getter, store) for a named variable, or None when the name is not a variable.  store(fr, v) replaces the binding."""
        r = self.sc.lookup_var(name)
        if r is None:
            return None
        k = r[0]
        if k == "local":
            def get(fr):
                return fr[name]

            def put(fr, v):
                fr[name] = v
            return get, put, r[1]
        if k in ("static", "module"):
            d = r[1]

            def get(fr):
                return d[name]

            def put(fr, v):
                d[name] = v
            return get, put, r[2]
        depth = r[1]
        if depth == 1:
            def get(fr):
                return fr["$host"][name]

            def put(fr, v):
                fr["$host"][name] = v
        else:
            def get(fr):
                for _ in range(depth):
                    fr = fr["$host"]
                return fr[name]

            def put(fr, v):
                for _ in range(depth):
                    fr = fr["$host"]
                fr[name] = v
        return get, put, r[2]

    # ---- expressions
    def expr(self, e):
        t = e.t
        if t == "num":
            if e.k == "i":
                v = int(e.v)
            elif e.k == "r8":
                v = F8(float(e.v))
            else:
                v = F4(e.v)
            return lambda fr: v
        if t == "str":
            v = e.v
            return lambda fr: v
        if t == "log":
            v = bool(e.v)
            return lambda fr: v
        if t == "paren":
            return self.expr(e.e)
        if t == "un":
            x = self.expr(e.e)
            if e.op == "neg":
                return lambda fr: -x(fr)
            return lambda fr: rt.lnot(x(fr))
        if t == "bin":
            return self.binop(e)
        if t == "arr":
            return self.array_constructor(e)
        if t == "desig":
            return self.desig(e)[0]
        if t == "star":
            return lambda fr: None
        raise InterpError(f"expression node {t}")

    def binop(self, e):
        op = e.op
        l, r = self.expr(e.l), self.expr(e.r)
        if op == "+":
            return lambda fr: l(fr) + r(fr)
        if op == "-":
            return lambda fr: l(fr) - r(fr)
        if op == "*":
            return lambda fr: l(fr) * r(fr)
        if op == "/":
            def div(fr):
                a, b = l(fr), r(fr)
                if type(a) is F8 and type(b) is F8:
                    return a / b
                return rt.fdiv(a, b)
            return div
        if op == "**":
            if e.r.t == "num" and e.r.k == "i" and int(e.r.v) == 2:
                def sq(fr):
                    a = l(fr)
                    return a * a
                return sq
            return lambda fr: rt.fpow(l(fr), r(fr))
        if op == "//":
            return lambda fr: l(fr) + r(fr)
        if op in ("eq", "ne", "lt", "le", "gt", "ge"):
            import operator
            pyop = {"eq": operator.eq, "ne": operator.ne, "lt": operator.lt, "le": operator.le, "gt": operator.gt, "ge": operator.ge}[op]

            def cmpf(fr):
                a, b = l(fr), r(fr)
                ta, tb = type(a), type(b)
                if (ta is int or ta is F8) and (tb is int or tb is F8):
                    return bool(pyop(a, b))
                return rt.cmp(op, a, b)
            return cmpf
        if op == "and":
            def andf(fr):
                a = l(fr)
                if a is False:
                    return False
                return rt.land(a, r(fr))
            return andf
        if op == "or":
            def orf(fr):
                a = l(fr)
                if a is True:
                    return True
                return rt.lor(a, r(fr))
            return orf
        if op == "eqv":
            return lambda fr: rt.cmp("eq", l(fr), r(fr))
        if op == "neqv":
            return lambda fr: rt.cmp("ne", l(fr), r(fr))
        raise InterpError(f"operator {op}")

    def ac_values(self, items):
        """Closures producing the flattened value list of an array constructor / I/O list (implied-do expanded)."""
        parts = []
        for it in items:
            if it.t == "ido":
                parts.append(("ido", self.implied_do_values(it)))
            else:
                parts.append(("v", self.expr(it)))

        def run(fr):
            out = []
            for k, f in parts:
                if k == "v":
                    out.append(f(fr))
                else:
                    out.extend(f(fr))
            return out
        return run

    def implied_do_values(self, it):
        acc = self.var_access(it.var)
        if acc is None:
            raise InterpError(f"implied-do variable {it.var} is not declared")
        put = acc[1]
        lo, hi = self.expr(it.lo), self.expr(it.hi)
        st = self.expr(it.st) if it.st is not None else None
        inner = self.ac_values(it.items)

        def run(fr):
            out = []
            s = 1 if st is None else int(st(fr))
            i, h = int(lo(fr)), int(hi(fr))
            while (i <= h) if s > 0 else (i >= h):
                put(fr, i)
                out.extend(inner(fr))
                i += s
            put(fr, i)
            return out
        return run

    def array_constructor(self, e):
        vals = self.ac_values(e.items)
        return lambda fr: rt.array_cons(vals(fr))

    # ---- designators
    def subscripts(self, args):
        """-> closure giving the tuple of subscripts (ints, FS triplets, arrays), flag: every subscript is a plain expression."""
        fs = []
        plain = True
        for a in args:
            if a.t == "slice":
                plain = False
                lo = self.expr(a.lo) if a.lo is not None else None
                hi = self.expr(a.hi) if a.hi is not None else None
                st = self.expr(a.st) if a.st is not None else None
                fs.append(("s", lo, hi, st))
            elif a.t == "kw":
                raise InterpError("keyword in subscript list")
            else:
                fs.append(("e", self.expr(a)))
        if plain:
            es = [f[1] for f in fs]
            if len(es) == 1:
                e0 = es[0]
                return (lambda fr: (e0(fr),)), True
            if len(es) == 2:
                e0, e1 = es
                return (lambda fr: (e0(fr), e1(fr))), True
            if len(es) == 3:
                e0, e1, e2 = es
                return (lambda fr: (e0(fr), e1(fr), e2(fr))), True
            if len(es) == 4:
                e0, e1, e2, e3 = es
                return (lambda fr: (e0(fr), e1(fr), e2(fr), e3(fr))), True
            return (lambda fr: tuple(x(fr) for x in es)), True

        def run(fr):
            out = []
            for f in fs:
                if f[0] == "e":
                    out.append(f[1](fr))
                else:
                    out.append(FS(None if f[1] is None else f[1](fr), None if f[2] is None else f[2](fr), None if f[3] is None else f[3](fr)))
            return tuple(out)
        return run, False

    @staticmethod
    def index_get(v, idx):
        if type(v) is FArray:
            return v[idx]
        if isinstance(v, str):
            s = idx[0]
            if isinstance(s, FS):
                return rt.substr(v, s.lo, s.hi)
            raise InterpError("character variable subscripted like an array")
        if v is None:
            raise InterpError("reference to an unallocated array")
        raise InterpError(f"subscript applied to {type(v)}")

    def fast_elem(self, base_get, sub):
        """Array element with plain subscripts: bounds-checked direct access for all-integer subscripts."""
        index_get = self.index_get

        def get(fr):
            v = base_get(fr)
            idx = sub(fr)
            if type(v) is FArray:
                d, lb = v.d, v.lb
                try:
                    if len(idx) == 1:
                        i = idx[0]
                        if type(i) is int:
                            i -= lb[0]
                            if i < 0:
                                raise IndexError
                            x = d[i]
                        else:
                            return v[idx]
                    else:
                        for i in idx:
                            if type(i) is not int:
                                return v[idx]
                        t = tuple([i - l for i, l in zip(idx, lb)])
                        if min(t) < 0 or len(t) != d.ndim:
                            raise IndexError
                        x = d[t]
                except IndexError:
                    raise InterpError(f"subscript {idx} out of bounds (lower {lb}, shape {d.shape})")
                k = d.dtype.kind
                if k == "f":
                    return x
                if k == "i":
                    return int(x)
                if k == "b":
                    return bool(x)
                return x
            return index_get(v, idx)
        return get

    def desig(self, e):
        """-> (getter, setter or None).  setter(fr, value) performs an intrinsic assignment to the designated object."""
        parts = e.parts
        p0 = parts[0]
        acc = self.var_access(p0.name)
        if acc is None:
            # function reference (user or intrinsic); further parts apply to its result
            get = self.funcref(p0)
            setter = None
            rest = parts[1:]
            if not rest:
                return get, None
            cur_get = get
            info = None
        else:
            vget, vput, info = acc
            cur_get = vget
            rest = parts[1:]
            if p0.args is None and not rest:
                return vget, self.var_setter(vget, vput, info, p0.name)
            if p0.args is not None:
                sub, plain = self.subscripts(p0.args)
                base = vget
                if not rest and p0.sub is None:
                    g = self.fast_elem(base, sub) if plain else (lambda fr: self.index_get(base(fr), sub(fr)))
                    return g, self.elem_setter(base, sub, vput)
                cur_get = self.fast_elem(base, sub) if plain else (lambda fr, base=base, sub=sub: self.index_get(base(fr), sub(fr)))
                if p0.sub is not None:
                    sub2, _ = self.subscripts(p0.sub)
                    inner = cur_get
                    if not rest:
                        return (lambda fr: self.index_get(inner(fr), sub2(fr))), self.substr_of_elem_setter(base, sub, sub2)
                    raise InterpError("substring followed by a component")
        # component chain
        for k, p in enumerate(rest):
            last = k == len(rest) - 1
            name = p.name
            container = cur_get
            comp_get = self.comp_getter(container, name)
            if p.args is None:
                if last:
                    return comp_get, self.comp_setter(container, name)
                cur_get = comp_get
                continue
            sub, plain = self.subscripts(p.args)
            if last and p.sub is None:
                # could be a type-bound function reference: decided at run time on the first evaluation
                return self.comp_elem_or_call(container, name, p, sub, plain), self.elem_setter(comp_get, sub, None)
            cur_get = self.fast_elem(comp_get, sub) if plain else (lambda fr, cg=comp_get, sub=sub: self.index_get(cg(fr), sub(fr)))
            if p.sub is not None:
                sub2, _ = self.subscripts(p.sub)
                inner = cur_get
                if last:
                    return (lambda fr: self.index_get(inner(fr), sub2(fr))), self.substr_of_elem_setter(comp_get, sub, sub2)
                raise InterpError("substring followed by a component")
        raise InterpError("designator")

    def comp_getter(self, container, name):
        def get(fr):
            o = container(fr)
            if type(o) is Struct:
                return o.f[name]
            if type(o) is FArray and o.d.dtype == object:
                flat = [x.f[name] for x in o.d.reshape(-1, order="F")]
                if flat and isinstance(flat[0], Struct):
                    a = np.empty(len(flat), dtype=object); a[:] = flat
                    return FArray(a.reshape(o.d.shape, order="F"))
                return rt.array_cons(flat) if o.d.ndim == 1 else FArray(np.array(flat).reshape(o.d.shape, order="F"))
            raise InterpError(f"component {name} of {type(o)}")
        return get

    def comp_setter(self, container, name):
        def put(fr, v):
            o = container(fr)
            if type(o) is Struct:
                self.store_into(o.f, name, v, o.t.info.get(name))
            elif type(o) is FArray and o.d.dtype == object:
                for x in o.d.reshape(-1, order="F"):
                    self.store_into(x.f, name, v, x.t.info.get(name))
            else:
                raise InterpError(f"component {name} of {type(o)}")
        return put

    def comp_elem_or_call(self, container, name, p, sub, plain):
        comp_get = self.comp_getter(container, name)
        elem = self.fast_elem(comp_get, sub) if plain else (lambda fr: self.index_get(comp_get(fr), sub(fr)))
        args = self.actual_args(p.args) if all(a.t != "slice" for a in p.args) else None
        I = self.I

        def get(fr):
            o = container(fr)
            if type(o) is Struct and name not in o.f and name in o.t.bindings:
                proc = o.t.module.procs[o.t.bindings[name]]
                return self.call_proc(proc, fr, args, obj=o)
            return elem(fr)
        return get

    @staticmethod
    def store_into(d, name, v, info):
        """Intrinsic assignment to the variable d[name] (scalar, whole array, derived type, allocatable)."""
        old = d[name]
        if type(old) is FArray:
            if isinstance(v, FArray) and v.d.shape != old.d.shape and info is not None and info.alloc:
                d[name] = Compiler.array_copy_for(info, v)
            else:
                old.assign(v)
        elif old is None:
            if isinstance(v, FArray):
                d[name] = Compiler.array_copy_for(info, v)
            else:
                raise InterpError(f"assignment of a scalar to the unallocated array {name}")
        else:
            d[name] = conv_like(old, v)

    @staticmethod
    def array_copy_for(info, v):
        """Copy of array value v with the element type of the declaration `info` (allocation on assignment)."""
        if info is None or info.code in ("c", "t"):
            return copy.deepcopy(FArray(v.d, (1,) * v.d.ndim))
        return FArray(np.array(v.d, dtype=rt.DT[info.code], order="F"), (1,) * v.d.ndim)

    def var_setter(self, vget, vput, info, name):
        if info.param:
            def bad(fr, v):
                raise InterpError(f"assignment to the named constant {name}")
            return bad
        if info.dims is None and info.code != "t":
            code = info.code
            if code == "r8":
                def put(fr, v):
                    vput(fr, v if type(v) is F8 else F8(v))
                return put
            if code in ("i4", "i2", "i8"):
                def put(fr, v):
                    if type(v) is not int:
                        if isinstance(v, (FArray, bool, np.bool_)):
                            raise InterpError(f"bad value {type(v)} for integer {name}")
                        v = int(v)
                    vput(fr, v)
                return put

            def put(fr, v):
                vput(fr, conv_like(vget(fr), v))
            return put
        array_copy_for = self.array_copy_for

        def put(fr, v):
            old = vget(fr)
            if type(old) is FArray:
                if isinstance(v, FArray) and v.d.shape != old.d.shape and info.alloc:
                    vput(fr, array_copy_for(info, v))
                else:
                    old.assign(v)
            elif old is None:
                if not isinstance(v, FArray):
                    raise InterpError(f"assignment of a scalar to the unallocated array {name}")
                vput(fr, array_copy_for(info, v))
            elif old is ABSENT:
                raise InterpError(f"assignment to the absent optional argument {name}")
            else:
                vput(fr, conv_like(old, v))
        return put

    def elem_setter(self, base, sub, vput):
        def put(fr, v):
            a = base(fr)
            idx = sub(fr)
            if type(a) is FArray:
                try:
                    a[idx] = v
                except IndexError as ex:
                    raise InterpError(str(ex))
            elif isinstance(a, str):
                s = idx[0]
                if not isinstance(s, FS) or vput is None:
                    raise InterpError("bad substring assignment")
                vput(fr, rt.set_substr(a, s.lo, s.hi, v))
            else:
                raise InterpError(f"subscripted assignment to {type(a)}")
        return put

    def substr_of_elem_setter(self, base, sub, sub2):
        def put(fr, v):
            a = base(fr)
            idx = sub(fr)
            s = sub2(fr)[0]
            a[idx] = rt.set_substr(a[idx], s.lo, s.hi, v)
        return put

    # ---- function references and calls
    def actual_args(self, args):
        """-> list of (keyword or None, value closure, setter or None)"""
        out = []
        for a in args:
            kw = None
            if a.t == "kw":
                kw, a = a.name, a.e
            if a.t == "desig":
                g, s = self.desig(a)
            else:
                g, s = self.expr(a), None
            out.append((kw, g, s))
        return out

    def call_proc(self, proc, fr, args, obj=None):
        """Calls a user procedure from frame fr; returns the function result (None for subroutines)."""
        I = self.I
        vals = [obj] if obj is not None else []
        kwv = None
        for kw, g, s in args:
            if kw is None:
                vals.append(g(fr))
            else:
                if kwv is None:
                    kwv = {}
                kwv[kw] = g(fr)
        host_fr = None
        if proc.host is not None:
            me = self.sc.proc
            if proc.host is me:
                host_fr = fr
            elif me is not None and proc.host is me.host:
                host_fr = fr["$host"]
            else:
                raise InterpError(f"call of internal procedure {proc.name} from outside its host")
        cfr = I.invoke(proc, vals, kwv, host_fr)
        # copy-out of scalars (and of arrays allocated by the callee)
        names = proc.argnames
        off = 1 if obj is not None else 0
        pinfo = proc.scope.info
        k = off
        for kw, g, s in args:
            if kw is None:
                nm = names[k]; passed = vals[k]; k += 1
            else:
                nm = kw; passed = kwv[kw]
            if s is None:
                continue
            new = cfr[nm]
            if new is passed:
                continue
            inf = pinfo[nm]
            if inf.intent == "in":
                continue
            if scalar_like(new) or inf.alloc or passed is None:
                if inf.dims is not None and not inf.alloc:
                    continue        # re-bounded view of the caller's array
                if type(new) is str and type(passed) is str and len(new) < len(passed):
                    new = new + passed[len(new):]      # character dummy shorter than the actual: the tail is untouched
                s(fr, new)
        if proc.resname is not None:
            return cfr[proc.resname]
        return None

    def funcref(self, p):
        name = p.name
        if name in ("c_loc", "c_associated") and p.args is not None and self.I.clib is not None:
            from . import cbind
            a0 = self.expr(p.args[0])
            cptr_t = self.sc.lookup_type("c_ptr")
            if name == "c_loc":
                return lambda fr: cbind.c_loc(self, cptr_t, a0(fr))
            return lambda fr: a0(fr).f["addr"] != 0
        proc = self.sc.lookup_proc(name)
        if proc is not None and p.args is not None:
            args = self.actual_args(p.args)
            return lambda fr: self.call_proc(proc, fr, args)
        fn = INTRINSICS.get(name)
        if fn is None or p.args is None:
            raise InterpError(f"{name}: not a variable, procedure or known intrinsic (in {self.sc.proc.name if self.sc.proc else self.sc.module.name})")
        pos, kws = [], []
        for a in p.args:
            if a.t == "kw":
                kws.append((a.name, self.expr(a.e)))
            elif a.t == "slice":
                raise InterpError(f"{name}: array section syntax on an intrinsic")
            else:
                pos.append(self.expr(a))
        if not kws:
            if len(pos) == 1:
                a0 = pos[0]
                return lambda fr: fn(a0(fr))
            if len(pos) == 2:
                a0, a1 = pos
                return lambda fr: fn(a0(fr), a1(fr))
            return lambda fr: fn(*[a(fr) for a in pos])
        return lambda fr: fn(*[a(fr) for a in pos], **{k: v(fr) for k, v in kws})

    # ---- statements
    def block(self, stmts):
        cs = [self.stmt(s) for s in stmts]
        if len(cs) == 1:
            return cs[0]

        def run(fr):
            for c in cs:
                r = c(fr)
                if r is not None:
                    return r
        return run

    def stmt(self, s):
        try:
            f = getattr(self, "s_" + s.t)(s)
        except InterpError as ex:
            if "line " in str(ex):
                raise
            raise InterpError(f"{ex} [{self.where()} line {s.line}]")
        line = s.line
        where = self.where()

        def guarded(fr):
            try:
                return f(fr)
            except (InterpError, FortranStop):
                raise
            except Exception as ex:
                raise InterpError(f"{type(ex).__name__}: {ex} [{where} line {line}]") from ex
        if s.t in ("assign",):
            # hot statements: the try block costs little, keep the location for diagnostics
            return guarded
        return guarded

    def where(self):
        if self.sc.proc is not None:
            return f"{os.path.basename(getattr(self.sc.proc.node, 'file', '') or '')}:{self.sc.proc.name}"
        return self.sc.module.name if self.sc.module else "?"

    def s_assign(self, s):
        get, put = self.desig(s.lhs)
        if put is None:
            raise InterpError("assignment to something that is not a variable")
        rhs = self.expr(s.rhs)

        def run(fr):
            put(fr, rhs(fr))
        return run

    def s_if(self, s):
        br = [(self.expr(c), self.block(b)) for c, b in s.branches]
        orelse = self.block(s.orelse) if s.orelse else None
        if len(br) == 1 and orelse is None:
            c0, b0 = br[0]

            def run1(fr):
                if c0(fr):
                    return b0(fr)
            return run1

        def run(fr):
            for c, b in br:
                if c(fr):
                    return b(fr)
            if orelse is not None:
                return orelse(fr)
        return run

    def s_do(self, s):
        acc = self.var_access(s.var)
        if acc is None:
            raise InterpError(f"do variable {s.var} is not declared")
        put = acc[1]
        lo, hi = self.expr(s.lo), self.expr(s.hi)
        st = self.expr(s.st) if s.st is not None else None
        body = self.block(s.body) if s.body else (lambda fr: None)

        def run(fr):
            i, h = int(lo(fr)), int(hi(fr))
            step = 1 if st is None else int(st(fr))
            n = (h - i + step) // step
            for _ in range(n if n > 0 else 0):
                put(fr, i)
                r = body(fr)
                if r is not None:
                    if r == EXIT:
                        return None
                    if r == RETURN:
                        return r
                i += step
            put(fr, i)
        return run

    def s_dowhile(self, s):
        cond = self.expr(s.cond)
        body = self.block(s.body) if s.body else (lambda fr: None)

        def run(fr):
            while cond(fr):
                r = body(fr)
                if r is not None:
                    if r == EXIT:
                        return None
                    if r == RETURN:
                        return r
        return run

    def s_select(self, s):
        sel = self.expr(s.sel)
        cases = []
        for vals, body in s.cases:
            cv = []
            for v in vals:
                if v[0] == "v":
                    cv.append(("v", self.expr(v[1])))
                else:
                    cv.append(("r", self.expr(v[1]) if v[1] is not None else None, self.expr(v[2]) if v[2] is not None else None))
            cases.append((cv, self.block(body) if body else (lambda fr: None)))
        default = self.block(s.default) if s.default else None

        def run(fr):
            x = sel(fr)
            if isinstance(x, str):
                x = x.rstrip(" ")
            for cv, body in cases:
                for v in cv:
                    if v[0] == "v":
                        y = v[1](fr)
                        if isinstance(y, str):
                            y = y.rstrip(" ")
                        if x == y:
                            return body(fr)
                    else:
                        if (v[1] is None or x >= v[1](fr)) and (v[2] is None or x <= v[2](fr)):
                            return body(fr)
            if default is not None:
                return default(fr)
        return run

    def s_associate(self, s):
        binds = []
        for name, e in s.pairs:
            g = self.expr(e)
            info = VarInfo(name, "t")
            self.sc.info[name] = info
            binds.append((name, g))
        body = self.block(s.body)

        def run(fr):
            for name, g in binds:
                fr[name] = g(fr)
            return body(fr)
        return run

    def s_cycle(self, s):
        return lambda fr: CYCLE

    def s_exit(self, s):
        return lambda fr: EXIT

    def s_return(self, s):
        return lambda fr: RETURN

    def s_continue(self, s):
        return lambda fr: None

    def s_stop(self, s):
        msg = self.expr(s.msg) if s.msg is not None else None

        def run(fr):
            raise FortranStop("STOP " + (str(msg(fr)) if msg else ""))
        return run

    def s_call(self, s):
        parts = s.target.parts
        last = parts[-1]
        argnodes = last.args or []
        if len(parts) == 1:
            name = last.name
            proc = self.sc.lookup_proc(name)
            if proc is not None:
                args = self.actual_args(argnodes)

                def run(fr):
                    self.call_proc(proc, fr, args)
                return run
            return self.builtin_call(name, argnodes)
        # type-bound procedure: object designator = all parts but the last
        objget = self.desig(Node("desig", parts=parts[:-1]))[0]
        mname = last.name
        args = self.actual_args(argnodes)

        def runm(fr):
            o = objget(fr)
            if type(o) is not Struct:
                raise InterpError(f"type-bound call {mname} on {type(o)}")
            tgt = o.t.bindings.get(mname)
            if tgt is None:
                raise InterpError(f"type {o.t.name} has no binding {mname}")
            self.call_proc(o.t.module.procs[tgt], fr, args, obj=o)
        return runm

    def builtin_call(self, name, argnodes):
        I = self.I
        args = self.actual_args(argnodes)
        if name == "cpu_time":
            s = args[0][2]
            return lambda fr: s(fr, rt.f_cpu_time())
        if name == "date_and_time":
            tgt = None
            for k, (kw, g, s) in enumerate(args):
                if kw == "values" or (kw is None and k == 3):
                    tgt = s
            if tgt is None:
                return lambda fr: None
            return lambda fr: tgt(fr, rt.f_date_and_time_values())
        if name == "myfork":
            s = args[0][2]

            def fork(fr):
                I.forked = True
                s(fr, 0)
            return fork
        if name == "myexit":
            def myexit(fr):
                if I.forked:
                    I.forked = False
                    return RETURN
                raise FortranStop("exit(%s)" % args[0][1](fr))
            return myexit
        if name == "c_f_pointer":
            from . import cbind
            src = args[0][1]
            acc = self.var_access(argnodes[1].parts[0].name)
            shp = args[2][1] if len(args) > 2 else None
            return lambda fr: acc[1](fr, cbind.c_f_pointer(acc[2], src(fr), None if shp is None else shp(fr)))
        if name in I.builtin_subs:
            fn = I.builtin_subs[name]
            return lambda fr: fn(*[g(fr) for kw, g, s in args])
        raise InterpError(f"call of unknown procedure {name}")

    def s_allocate(self, s):
        todo = []
        for it in s.items:
            parts = it.parts
            last = parts[-1]
            if last.args is None:
                raise InterpError("allocate without a shape")
            bl = []
            for a in last.args:
                if a.t == "slice":
                    bl.append((self.expr(a.lo) if a.lo is not None else None, self.expr(a.hi)))
                else:
                    bl.append((None, self.expr(a)))
            if len(parts) == 1:
                acc = self.var_access(last.name)
                if acc is None:
                    raise InterpError(f"allocate: {last.name} is not a variable")
                vget, vput, info = acc
                todo.append(("v", vput, info, bl, self))
            else:
                cont = self.desig(Node("desig", parts=parts[:-1]))[0]
                todo.append(("c", cont, last.name, bl, None))

        def run(fr):
            for kind, a, b, bl, cp in todo:
                bounds = [(1 if lo is None else int(lo(fr)), int(hi(fr))) for lo, hi in bl]
                if kind == "v":
                    a(fr, cp.new_array(b, bounds, fr))
                else:
                    o = a(fr)
                    if type(o) is not Struct:
                        raise InterpError(f"allocate of a component of {type(o)}")
                    info = o.t.info[b]
                    o.f[b] = o.t.cp.new_array(info, bounds, {})
        return run

    def s_deallocate(self, s):
        todo = []
        for it in s.items:
            parts = it.parts
            if len(parts) == 1:
                acc = self.var_access(parts[0].name)
                todo.append(("v", acc[1], None))
            else:
                cont = self.desig(Node("desig", parts=parts[:-1]))[0]
                todo.append(("c", cont, parts[-1].name))

        def run(fr):
            for kind, a, b in todo:
                if kind == "v":
                    a(fr, None)
                else:
                    a(fr).f[b] = None
        return run

    # ---- input / output
    def io_ctl(self, s):
        """unit closure (None = *), unit setter (internal files), format: None (unformatted) | '*' | closure."""
        ctl, kws = s.ctl, s.kws
        unit_node = ctl[0] if ctl else kws.get("unit")
        fmt_node = ctl[1] if len(ctl) > 1 else kws.get("fmt")
        unit_get = unit_set = None
        if unit_node is not None and unit_node.t != "star":
            if unit_node.t == "desig":
                unit_get, unit_set = self.desig(unit_node)
            else:
                unit_get = self.expr(unit_node)
        if fmt_node is None:
            fmt = None
        elif fmt_node.t == "star":
            fmt = "*"
        else:
            fmt = self.expr(fmt_node)
        return unit_get, unit_set, fmt

    def s_write(self, s):
        I = self.I
        unit_get, unit_set, fmt = self.io_ctl(s)
        vals = self.ac_values(s.items)
        adv = self.expr(s.kws["advance"]) if "advance" in s.kws else None
        iostat = self.desig(s.kws["iostat"])[1] if "iostat" in s.kws else None

        def run(fr):
            values = vals(fr)
            u = unit_get(fr) if unit_get is not None else None
            if fmt is None:
                I.io.write_stream(u, values)
            else:
                text = rt.list_directed(values) if fmt == "*" else rt.format_values(fmt(fr), values)
                if isinstance(u, str):
                    unit_set(fr, text)
                else:
                    advance = True if adv is None else adv(fr).strip().lower() != "no"
                    I.io.write_line(u, text, advance)
            if iostat is not None:
                iostat(fr, 0)
        return run

    def io_targets(self, items):
        out = []
        for it in items:
            if it.t == "ido":
                acc = self.var_access(it.var)
                out.append(("ido", acc[1], self.expr(it.lo), self.expr(it.hi), self.expr(it.st) if it.st is not None else None, self.io_targets(it.items)))
            else:
                g, st = self.desig(it)
                if st is None:
                    raise InterpError("read into something that is not a variable")
                out.append(("v", g, st))
        return out

    @staticmethod
    def code_of(v):
        if type(v) is int:
            return "i4"
        if type(v) is F8:
            return "r8"
        if type(v) is F4:
            return "r4"
        if type(v) is bool:
            return "l"
        if isinstance(v, str):
            return "c%d" % len(v)
        raise InterpError(f"I/O of {type(v)}")

    def s_read(self, s):
        I = self.I
        unit_get, unit_set, fmt = self.io_ctl(s)
        targets = self.io_targets(s.items)
        iostat = self.desig(s.kws["iostat"])[1] if "iostat" in s.kws else None
        code_of = self.code_of
        loc = f"{self.where()} line {s.line}"

        def each(fr, tg, fn):
            for t in tg:
                if t[0] == "v":
                    fn(t[1], t[2])
                else:
                    _, put, lo, hi, st, inner = t
                    step = 1 if st is None else int(st(fr))
                    i, h = int(lo(fr)), int(hi(fr))
                    while (i <= h) if step > 0 else (i >= h):
                        put(fr, i)
                        each(fr, inner, fn)
                        i += step
                    put(fr, i)

        def run(fr):
            u = unit_get(fr) if unit_get is not None else None
            try:
                if fmt is None:
                    def rd(g, st):
                        old = g(fr)
                        if type(old) is FArray:
                            code = {"f8": "r8", "f4": "r4", "i4": "i4", "i2": "i2", "i8": "i8", "b1": "l"}[old.d.dtype.kind + str(old.d.dtype.itemsize)]
                            a = I.io.read_stream(u, code, old.d.size)
                            st(fr, FArray(a.reshape(old.d.shape, order="F")))
                        else:
                            st(fr, I.io.read_stream(u, code_of(old)))
                    each(fr, targets, rd)
                elif fmt == "*":
                    if not targets and not isinstance(u, str):
                        if I.io.read_line(u) is None:
                            raise EOFError
                    else:
                        lr = rt.ListReader(I.io, u)

                        def rd(g, st):
                            old = g(fr)
                            if type(old) is FArray:
                                k = old.d.dtype.kind
                                code = "c" if k == "O" else {"f8": "r8", "f4": "r4", "i4": "i4", "i2": "i2", "i8": "i8", "b1": "l"}[k + str(old.d.dtype.itemsize)]
                                vals = [lr.next(code) for _ in range(old.d.size)]
                                if k == "O":
                                    a = np.empty(len(vals), dtype=object); a[:] = vals
                                else:
                                    a = np.array(vals, dtype=old.d.dtype)
                                st(fr, FArray(a.reshape(old.d.shape, order="F")))
                            else:
                                st(fr, lr.next(code_of(old)))
                        each(fr, targets, rd)
                else:
                    f = fmt(fr).strip().lower().replace(" ", "")
                    if f not in ("(a)",) and not f.startswith("(a"):
                        raise InterpError(f"formatted read with {f!r} is not supported")
                    line = u if isinstance(u, str) else I.io.read_line(u)
                    if line is None:
                        raise EOFError

                    def rd(g, st):
                        st(fr, line)
                    each(fr, targets, rd)
            except EOFError:
                if iostat is None:
                    raise FortranStop(f"READ: end of file [{loc}]")
                iostat(fr, -1)
                return
            except ValueError as ex:
                if iostat is None:
                    raise FortranStop(f"READ: {ex} [{loc}]")
                iostat(fr, 5010)
                return
            if iostat is not None:
                iostat(fr, 0)
        return run

    def kwvals(self, s, names):
        out = {}
        for k in names:
            if k in s.kws:
                out[k] = self.expr(s.kws[k])
        return out

    def s_open(self, s):
        I = self.I
        unit = self.expr(s.ctl[0]) if s.ctl else self.expr(s.kws["unit"])
        kw = self.kwvals(s, ("file", "status", "action", "form", "access", "position"))
        iostat = self.desig(s.kws["iostat"])[1] if "iostat" in s.kws else None

        def run(fr):
            rc = I.io.open(unit(fr), iostat=iostat, **{k: v(fr) for k, v in kw.items()})
            if iostat is not None:
                iostat(fr, rc or 0)
        return run

    def s_close(self, s):
        I = self.I
        unit = self.expr(s.ctl[0]) if s.ctl else self.expr(s.kws["unit"])
        return lambda fr: I.io.close(unit(fr))

    def s_rewind(self, s):
        I = self.I
        unit = self.expr(s.ctl[0])
        return lambda fr: I.io.rewind(unit(fr))

    def s_flush(self, s):
        return lambda fr: None

    def s_backspace(self, s):
        I = self.I
        unit = self.expr(s.ctl[0])

        def run(fr):
            u = I.io.units.get(int(unit(fr)))
            if u is not None and u.lines is not None and u.pos > 0:
                u.pos -= 1
        return run

    def s_inquire(self, s):
        I = self.I
        f = self.expr(s.kws["file"])
        ex = self.desig(s.kws["exist"])[1]
        return lambda fr: ex(fr, os.path.exists(I.io._path(f(fr))))


def _present(x):
    return x is not ABSENT


def _allocated(x):
    return x is not None


def _ieee_is_finite(x):
    u = rt._unwrap(x)
    if isinstance(u, np.ndarray):
        return FArray(np.isfinite(u))
    return bool(np.isfinite(u))


def _merge(a, b, m):
    ua, ub, um = rt._unwrap(a), rt._unwrap(b), rt._unwrap(m)
    if isinstance(um, np.ndarray):
        return FArray(np.where(um, ua, ub))
    return a if um else b


INTRINSICS = {
    "abs": rt.f_abs, "dabs": rt.f_abs, "iabs": rt.f_abs,
    "sqrt": rt.f_sqrt, "dsqrt": rt.f_sqrt,
    "sin": rt.f_sin, "dsin": rt.f_sin, "cos": rt.f_cos, "dcos": rt.f_cos, "tan": rt.f_tan, "dtan": rt.f_tan,
    "acos": rt.f_acos, "dacos": rt.f_acos, "asin": rt.f_asin, "dasin": rt.f_asin, "atan": rt.f_atan, "datan": rt.f_atan,
    "atan2": rt.f_atan2, "datan2": rt.f_atan2,
    "exp": rt.f_exp, "dexp": rt.f_exp, "log": rt.f_log, "dlog": rt.f_log, "log10": rt.f_log10,
    "floor": rt.f_floor, "ceiling": rt.f_ceiling, "nint": rt.f_nint, "int": rt.f_int, "real": rt.f_real, "dble": rt.f_dble,
    "float": rt.f_real, "mod": rt.f_mod, "modulo": rt.f_modulo, "sign": rt.f_sign, "dsign": rt.f_sign,
    "max": rt.f_max, "min": rt.f_min, "dmax1": rt.f_max, "dmin1": rt.f_min, "max0": rt.f_max, "min0": rt.f_min,
    "maxval": rt.f_maxval, "minval": rt.f_minval, "maxloc": rt.f_maxloc, "minloc": rt.f_minloc,
    "sum": rt.f_sum, "dot_product": rt.f_dot_product, "matmul": rt.f_matmul, "transpose": rt.f_transpose, "reshape": rt.f_reshape,
    "size": rt.f_size, "lbound": rt.f_lbound, "ubound": rt.f_ubound, "count": rt.f_count, "any": rt.f_any, "all": rt.f_all,
    "isnan": rt.f_isnan, "ieee_is_nan": rt.f_isnan, "ieee_is_finite": _ieee_is_finite, "merge": _merge,
    "trim": rt.f_trim, "adjustl": rt.f_adjustl, "adjustr": rt.f_adjustr, "len_trim": rt.f_len_trim, "len": rt.f_len,
    "index": rt.f_index, "achar": rt.f_achar, "char": rt.f_achar, "iachar": rt.f_iachar, "ichar": rt.f_iachar,
    "epsilon": rt.f_epsilon, "huge": rt.f_huge, "tiny": rt.f_tiny,
    "present": _present, "allocated": _allocated,
}

REF_ORDER = ["ConstParams", "FlowCondition", "SolidSolver", "Solidbody", "FluidDomain", "LBMBlockComm", "Util", "main"]


def load_reference(ref="/root/reference", cwd=".", echo=False):
    """Interpreter with the reference's own source files loaded (the list and order of its Makefile:33)."""
    paths = [os.path.join(ref, "src", n + ".f90") for n in REF_ORDER]
    missing = [p for p in paths if not os.path.exists(p)]
    if missing:
        raise FileNotFoundError(f"reference sources not found: {missing}")
    return Interp(cwd=cwd, echo=echo).load(paths)
