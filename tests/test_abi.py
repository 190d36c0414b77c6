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
