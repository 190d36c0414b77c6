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


class _Slab:
    def __init__(self, xDim, xOffset, xLocal):
        self.xDim, self.xOffset, self.xLocal, self.yDim, self.zDim = xDim, xOffset, xLocal, 10, 12
        self.xmin = self.ymin = self.zmin = 0.0
        self.dh = 1.0


def test_flow_window_is_the_slab_part_of_the_output_window():
    """the staging buffer of write_flow must have the size fsilbm_block_write_flow_window fills on an x-slab."""
    assert flow_io.flow_window(_Slab(32, 0, 32), 2) == (2, 28, 6, 8)          # whole block
    assert flow_io.flow_window(_Slab(32, 0, 8), 2) == (2, 6, 6, 8)            # first slab loses the offset planes
    assert flow_io.flow_window(_Slab(32, 8, 8), 2) == (8, 8, 6, 8)            # interior slab: all its planes
    assert flow_io.flow_window(_Slab(32, 24, 8), 2) == (24, 6, 6, 8)          # last slab
    assert flow_io.flow_window(_Slab(32, 0, 2), 2)[1] == 0                    # window misses the slab


def test_restart_helpers_refuse_slabs():
    import pytest
    with pytest.raises(ValueError, match="x-slab"):
        flow_io.write_continue_blocks([_Slab(32, 8, 8)], 1, 0.5, root="/nonexistent")
    with pytest.raises(ValueError, match="x-slab"):
        flow_io.regrid_from_continue(_Slab(32, 8, 8), [])


def test_async_flow_window_refuses_a_staging_array_of_the_wrong_size():
    """The library fills nfields x window floats through a raw pointer: the mirror checks the array before the call."""
    import pytest
    from fsilbm3d_b200.fluid_domain import LBMBlock
    slab = _Slab(32, 8, 8)
    slab._h = -1
    with pytest.raises(ValueError, match="staging array"):
        LBMBlock.write_flow_window_async(slab, np.empty(4 * 8 * 6 * 8 - 1, dtype=np.float32), 2, 1)
    with pytest.raises(ValueError, match="staging array"):
        LBMBlock.write_flow_window_async(slab, np.empty(4 * 8 * 6 * 8, dtype=np.float32), 2, 2)     # 13 fields wanted
    with pytest.raises(Exception) as ei:                                                             # right size: reaches the library
        LBMBlock.write_flow_window_async(slab, np.empty(4 * 8 * 6 * 8, dtype=np.float32), 2, 1)
    assert not isinstance(ei.value, ValueError)
