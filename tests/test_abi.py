"""CPU checks of the drop-in boundary: the C-ABI library builds, loads and exports every symbol
include/fsilbm.h declares; without a GPU every compute entry fails loudly (no fallback)."""
import ctypes
import os

import pytest

import fsilbm3d_b200 as F
from fsilbm3d_b200.build import build


def test_library_builds_for_sm100a():
    path = build()
    assert os.path.exists(path)


def test_every_declared_symbol_is_exported():
    declared = F.declared_symbols()
    assert len(declared) >= 30
    exported = F.exported_symbols()
    assert set(declared) == set(exported), set(declared) - set(exported)


def test_sass_is_sm100a():
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", F.library_path()], capture_output=True, text=True).stdout
    assert "sm_100a" in out, out


def test_no_oracle_linkage():
    """The product never links or imports the oracle."""
    import subprocess
    out = subprocess.run(["ldd", F.library_path()], capture_output=True, text=True).stdout
    assert "oracle" not in out
    pkg = os.path.dirname(F.__file__)
    for root, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(root, fn)).read()
                assert "import oracle" not in text and "from oracle" not in text and "fsilbm_oracle" not in text, fn


def test_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(F.FsilbmError):
        F.LBMBlock(8, 8, 8)


def test_slab_range_partitions_x():
    for X in (7, 64, 1024):
        for n in (1, 2, 3, 8):
            if n > X:
                continue
            spans = [F.slab_range(X, r, n) for r in range(n)]
            assert spans[0][0] == 0 and sum(c for _, c in spans) == X
            for (o0, c0), (o1, _) in zip(spans, spans[1:]):
                assert o0 + c0 == o1


def _header_prototypes():
    import re
    text = open(os.path.join(os.path.dirname(F.__file__), "..", "include", "fsilbm.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    protos = {}
    for _, name, args in re.findall(r"^\s*((?:const\s+)?(?:int|long long|const char \*))\s*(fsilbm_\w+)\s*\(([^;]*?)\)\s*;", text, flags=re.M):
        a = " ".join(args.split())
        protos[name] = [] if a in ("void", "") else [re.search(r"(\w+)(\[\d*\])?$", x.strip()).group(1) for x in a.split(",")]
    return protos


def _fortran_interfaces():
    """{C symbol: (dummy names, declared names)} of the bind(C) interfaces in fortran/fsilbm_gpu.f90 (statement level: the file
    cannot be compiled here, so the check is textual, on logical lines joined by the interpreter's free-form lexer)."""
    import re
    from oracle.ftn.lex import logical_lines
    src = open(os.path.join(os.path.dirname(F.__file__), "..", "fortran", "fsilbm_gpu.f90")).read()
    lines = [s for _, s in logical_lines(src)]
    out, cur = {}, None
    for s in lines:
        m = re.match(r"(?i)^(integer\(c_int\)|integer\(c_long_long\)|type\(c_ptr\))\s+function\s+(\w+)\s*\((.*?)\)\s*bind\(C,\s*name='(\w+)'\)", s)
        if m:
            dummies = [x.strip() for x in m.group(3).split(",") if x.strip()]
            assert m.group(2) == m.group(4)
            cur = out[m.group(4)] = (dummies, [])
            continue
        if cur is not None:
            if re.match(r"(?i)^end function", s):
                cur = None
            elif "::" in s and not s.lower().startswith("import"):
                for ent in s.split("::", 1)[1].split(","):
                    ent = ent.strip()
                    if ent and not ent[0].isdigit() and ent != "*)":
                        cur[1].append(re.match(r"\w+", ent).group(0))
    return out, lines


def test_fortran_shim_binds_every_symbol():
    """fortran/fsilbm_gpu.f90 holds one bind(C) interface per symbol of include/fsilbm.h, with the same argument count and
    order (names may differ where a C name is a Fortran keyword), every dummy declared exactly once."""
    protos = _header_prototypes()
    ifaces, lines = _fortran_interfaces()
    assert set(protos) == set(F.declared_symbols())
    assert set(ifaces) == set(protos), (set(protos) - set(ifaces), set(ifaces) - set(protos))
    rename = {"value": "val"}
    for name, cargs in protos.items():
        dummies, declared = ifaces[name]
        assert [rename.get(a, a).lower() for a in cargs] == [d.lower() for d in dummies], name
        assert sorted(d.lower() for d in dummies) == sorted(d.lower() for d in declared), name
    # every wrapper named in the public list exists, and every call point the integration guide routes through the shim has one
    text = "\n".join(lines).lower()
    import re
    public = re.search(r"public :: (fsilbm_check.*?)\n", text).group(1)
    names = [n.strip() for n in public.split(",")]
    assert len(names) >= 45
    for n in names:
        assert re.search(rf"(subroutine|function) {n}\b", text), n
    for needed in ("gpu_extract_interpolate_layer", "gpu_interpolation_father_to_son", "gpu_deliver_son_to_father", "gpu_pair_create",
                   "gpu_comm_init", "gpu_set_option", "gpu_write_flow_window", "gpu_fluid_flux", "gpu_probe_velocity", "gpu_turbulent_statistic"):
        assert needed in names
    # block structure: every opener has its end
    for kw in ("subroutine", "function", "module", "interface", "type"):
        opens = len(re.findall(rf"(?m)^(?:[\w()]+\s+)?{kw}\b(?!\()", text)) - len(re.findall(rf"(?m)^end {kw}", text)) if kw != "type" else 0
        ends = len(re.findall(rf"(?m)^end {kw}", text))
        if kw != "type":
            assert opens == ends, (kw, opens, ends)


def test_every_option_is_documented():
    """Every key fsilbm_set_option accepts is described in the header's option table (and nothing else is listed there)."""
    import re
    ROOT = os.path.join(os.path.dirname(F.__file__), "..")
    src = open(os.path.join(ROOT, "fsilbm3d_b200", "csrc", "fsilbm_api.cu")).read()
    body = src[src.index("int fsilbm_set_option("):]
    body = body[:body.index("\n}\n")]
    keys = set(re.findall(r'strcmp\(key, "([a-z_0-9]+)"\)', body))
    assert len(keys) >= 10
    header = open(os.path.join(ROOT, "include", "fsilbm.h")).read()
    table = header[header.index("Tuning/testing switches"):header.index("int fsilbm_set_option(")]
    listed = set(re.findall(r'^ \*   "([a-z_0-9]+)"', table, flags=re.M))
    assert keys == listed, (keys - listed, listed - keys)
