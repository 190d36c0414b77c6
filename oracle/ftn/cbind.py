"""bind(C) interoperability of the Fortran interpreter: calls of `bind(C)` interface procedures go to a ctypes library with the
argument association the Fortran standard prescribes (VALUE dummies by value, everything else by reference; arrays as the address of
their first element; bind(C) derived types as C structs; type(c_ptr) as void*).  TEST INFRASTRUCTURE: it lets the ISO_C_BINDING shim
fortran/fsilbm_gpu.f90 -- which cannot be compiled in this image -- be EXECUTED against libfsilbm_b200.so by tests/."""
from __future__ import annotations

import ctypes as C

import numpy as np

from .rt import FArray

SCALAR = {"i4": C.c_int, "i2": C.c_short, "i8": C.c_longlong, "r8": C.c_double, "r4": C.c_float, "l": C.c_bool}
NPDT = {"i4": np.int32, "i2": np.int16, "i8": np.int64, "r8": np.float64, "r4": np.float32}


def _is_cptr(v):
    return hasattr(v, "t") and getattr(v.t, "name", "") == "c_ptr"


def c_loc(cp, cptr_type, x):
    """c_loc(x): the address of an array (which must be contiguous: a TARGET variable, not a section copy)."""
    if not isinstance(x, FArray) or x.d.dtype == object or not x.d.flags.f_contiguous:
        raise TypeError("c_loc needs a contiguous numeric array")
    p = cp.instantiate(cptr_type)
    p.f["addr"] = int(x.d.ctypes.data)
    return p


def c_f_pointer(info, cptr, shape):
    """c_f_pointer(cptr, fptr, shape): an array over C memory.  character(kind=c_char) arrays are copied (read-only use)."""
    addr = int(cptr.f["addr"])
    n = int(np.prod(shape.d)) if shape is not None else 1
    shp = tuple(int(v) for v in shape.d) if shape is not None else (1,)
    if info.code == "c":
        raw = C.string_at(addr, n) if addr else b"\0" * n
        a = np.empty(n, dtype=object)
        a[:] = [chr(b) for b in raw]
        return FArray(a.reshape(shp, order="F"))
    ct = SCALAR[info.code]
    arr = np.ctypeslib.as_array(C.cast(addr, C.POINTER(ct)), shape=(n,))
    return FArray(arr.reshape(shp, order="F"))


def _struct_class(td, cp):
    """ctypes.Structure of a bind(C) derived type, components in declaration order."""
    if getattr(td, "cstruct", None) is None:
        cp.instantiate(td)            # fills td.info
        fields = []
        for name, info in td.info.items():
            ct = SCALAR[info.code]
            if info.dims is not None:
                n = 1
                for lo, hi in td.cp.bounds_of(info, {}):
                    n *= hi - lo + 1
                ct = ct * n
            fields.append((name, ct))
        td.cstruct = type("c_" + td.name, (C.Structure,), {"_fields_": fields})
    return td.cstruct


def call(I, proc, fr):
    """Invokes the C function behind a bind(C) interface; fr holds the actual arguments under the dummy names."""
    if I.clib is None:
        raise RuntimeError(f"{proc.name}: bind(C) procedure called but no C library is bound (Interp.bind_c_library)")
    fn = getattr(I.clib, proc.cname)
    sc = proc.scope
    from .interp import Compiler, Struct
    cp = Compiler(I, sc)
    cargs, post = [], []
    for nm in proc.argnames:
        info, v = sc.info[nm], fr[nm]
        if info.dims is not None:                                   # array: address of the first element
            if not isinstance(v, FArray):
                raise TypeError(f"{proc.name}: array expected for {nm}")
            if info.code == "c":
                flat = v.d.reshape(-1, order="F")
                buf = C.create_string_buffer(bytes(ord(ch[0]) if ch else 0 for ch in flat), len(flat))
                cargs.append(buf)
                if info.intent != "in":
                    post.append(lambda v=v, buf=buf: v.assign(FArray(np.array([chr(b) for b in buf.raw], dtype=object).reshape(v.d.shape, order="F"))))
            elif info.code == "t":                                   # array of type(c_ptr)
                flat = v.d.reshape(-1, order="F")
                cargs.append((C.c_void_p * len(flat))(*[int(p.f["addr"]) for p in flat]))
            else:
                d = v.d
                if d.dtype != NPDT[info.code]:
                    raise TypeError(f"{proc.name}: {nm} is {d.dtype}, the interface says {info.code}")
                if d.flags.f_contiguous:
                    cargs.append(C.c_void_p(d.ctypes.data))
                else:                                                # a section: copy in / copy out
                    tmp = np.asfortranarray(d)
                    cargs.append(C.c_void_p(tmp.ctypes.data))
                    if info.intent != "in":
                        post.append(lambda d=d, tmp=tmp: d.__setitem__(Ellipsis, tmp))
                    post.append(lambda tmp=tmp: None)                # keeps tmp alive over the call
        elif info.code == "t":
            if _is_cptr(v):
                if info.value:
                    cargs.append(C.c_void_p(int(v.f["addr"])))
                else:                                                # type(c_ptr), intent(out): void**
                    box = C.c_void_p(int(v.f["addr"]))
                    cargs.append(C.byref(box))
                    post.append(lambda v=v, box=box: v.f.__setitem__("addr", int(box.value or 0)))
            else:                                                    # bind(C) struct by reference
                cls = _struct_class(v.t, cp)
                st = cls()
                for name, _ct in cls._fields_:
                    val = v.f[name]
                    if isinstance(val, FArray):
                        getattr(st, name)[:] = [x.item() if hasattr(x, "item") else x for x in val.d.reshape(-1, order="F")]
                    else:
                        setattr(st, name, val.item() if hasattr(val, "item") else val)
                cargs.append(C.byref(st))
                if info.intent != "in":
                    def back(v=v, st=st, cls=cls):
                        for name, _ct in cls._fields_:
                            val = v.f[name]
                            if isinstance(val, FArray):
                                val.assign(FArray(np.array(list(getattr(st, name)), dtype=val.d.dtype)))
                            else:
                                v.f[name] = type(val)(getattr(st, name))
                    post.append(back)
        elif info.code == "c":
            raise TypeError(f"{proc.name}: scalar character argument {nm} is not supported")
        else:
            ct = SCALAR[info.code]
            pv = v.item() if hasattr(v, "item") else v
            if info.value:
                cargs.append(ct(pv))
            else:                                                    # scalar by reference
                box = ct(pv)
                cargs.append(C.byref(box))
                if info.intent != "in":
                    def back(nm=nm, box=box, code=info.code):
                        fr[nm] = int(box.value) if code in ("i4", "i2", "i8") else (np.float64(box.value) if code == "r8" else np.float32(box.value))
                    post.append(back)
    rinfo = sc.info[proc.resname] if proc.resname else None
    if rinfo is None:
        fn.restype = None
    elif rinfo.code == "t":
        fn.restype = C.c_void_p
    else:
        fn.restype = SCALAR[rinfo.code]
    fn.argtypes = None
    res = fn(*cargs)
    for f in post:
        f()
    if rinfo is not None:
        if rinfo.code == "t":
            p = cp.instantiate(sc.lookup_type("c_ptr"))
            p.f["addr"] = int(res or 0)
            fr[proc.resname] = p
        elif rinfo.code in ("i4", "i2", "i8"):
            fr[proc.resname] = int(res)
        else:
            fr[proc.resname] = np.float64(res) if rinfo.code == "r8" else np.float32(res)
    return fr
