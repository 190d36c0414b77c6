"""Format helpers of fsilbm3d_b200.flow_io that need no GPU."""
import numpy as np

from fsilbm3d_b200 import flow_io


def test_name10_is_fortran_nint_zero_padded():
    assert flow_io._name10(0.12) == "0000012000"
    assert flow_io._name10(0.000005) == "0000000001"     # nint(0.5) = 1 (away from zero), unlike Python's round
    assert flow_io._name10(123.456789) == "0012345679"


def test_e20_10_matches_fortran_edit_descriptor():
    assert flow_io._e20_10(0.12) == "    0.1200000000E+00"
    assert flow_io._e20_10(-1234.5) == "   -0.1234500000E+04"
    assert flow_io._e20_10(0.0) == "    0.0000000000E+00"
    assert flow_io._e20_10(9.99999999999e-5) == "    0.1000000000E-03"
    assert all(len(flow_io._e20_10(v)) == 20 for v in (1e-30, 3.14, -2.5e17))


def test_continue_file_reader_round_trip(tmp_path):
    rng = np.random.default_rng(1)
    f = rng.uniform(size=(19, 3, 4, 5))
    p = tmp_path / "c"
    with open(p, "wb") as fh:
        fh.write(np.array([1, 42], np.int32).tobytes()); fh.write(np.array([0.75]).tobytes())
        fh.write(np.array([0.5, 1.5, 2.5, 0.25]).tobytes()); fh.write(np.array([3, 4, 5], np.int32).tobytes()); fh.write(f.tobytes())
    n, step, t, blocks = flow_io.read_continue_file(str(p))
    assert (n, step, t) == (1, 42, 0.75) and np.array_equal(blocks[0]["fIn"], f) and blocks[0]["dh"] == 0.25
