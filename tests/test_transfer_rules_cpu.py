"""The two rules the merged father<->son transfer launches rest on (csrc/refine_kernels.cu), restated in numpy and held against the
pass-by-pass restatement of interpolate_fIn (oracle/numpy_restatement.interpolate_plane, LBMBlockComm.f90:808-905):

1. linear scheme, node outside the periodic closures: the value is (row + row) * 0.5 of rows that are themselves F or (F + F) * 0.5,
   with offsets that depend on the parity of the node's face indices only -- the fast path of pair_f2s_kernel;
2. a son node on an edge or corner keeps the value of the LAST coupled face that holds it -- the rule by which all faces of a pair
   run in one launch.
"""
import numpy as np
import pytest

from oracle.numpy_restatement import interpolate_plane


def fast_path_node(F, bF, B, A):
    """pair_f2s_kernel's fast path for the 1-based son node (B, A); F is one population's coarse plane, flat [aF * bF]."""
    be, ae = B % 2 == 0, A % 2 == 0
    c = ((B - 1) if be else B) // 2 + 1
    r0 = ((A - 1) if ae else A) // 2 + 1
    r1 = (A + 1) // 2 + 1
    o0, o1 = (r0 - 1) * bF + (c - 1), (r1 - 1) * bF + (c - 1)
    lo = (F[o0] + F[o0 + 1]) * 0.5 if be else F[o0]
    if not ae:
        return lo
    hi = (F[o1] + F[o1 + 1]) * 0.5 if be else F[o1]
    return (lo + hi) * 0.5


@pytest.mark.parametrize("bS,aS", [(9, 7), (8, 7), (9, 6), (8, 6), (33, 17), (3, 3)])
def test_fast_path_equals_the_interpolation_passes(bS, aS):
    rng = np.random.default_rng(20261017 + bS * 100 + aS)
    bT, aT = (bS - 1 if bS % 2 == 0 else bS), (aS - 1 if aS % 2 == 0 else aS)
    bF, aF = (bT + 1) // 2 + 1, (aT + 1) // 2 + 2          # the layer buffers may be wider than what the passes read
    F = rng.uniform(-1.0, 1.0, size=(aF, bF))
    want = interpolate_plane(F, aS, bS, 1)
    flat = F.reshape(-1)
    for A in range(1, aT + 1):
        for B in range(1, bT + 1):
            got = fast_path_node(flat, bF, B, A)
            assert got == want[A - 1, B - 1], (B, A)           # bit for bit: the same additions in the same order


def test_last_face_wins_on_edges_and_corners():
    """Six faces written one after the other (the reference, LBMBlockComm.f90:669) against every node taken by the last face of the
    list that holds it (the merged launch): identical, for every subset of coupled faces."""
    X, Y, Z = 5, 4, 6
    planes = [(0, 0), (0, X - 1), (1, 0), (1, Y - 1), (2, 0), (2, Z - 1)]      # (axis, boundary plane) of faces j = 0..5
    idx = np.indices((X, Y, Z))
    for mask in range(1, 64):
        faces = [j for j in range(6) if mask >> j & 1]
        seq = np.full((X, Y, Z), -1)
        for j in faces:                                       # sequential: later faces overwrite
            ax, pl = planes[j]
            seq[idx[ax] == pl] = j
        par = np.full((X, Y, Z), -1)
        for k, j in enumerate(faces):                         # merged: a node is left to any later face whose plane holds it
            ax, pl = planes[j]
            mine = idx[ax] == pl
            for j2 in faces[k + 1:]:
                ax2, pl2 = planes[j2]
                mine &= idx[ax2] != pl2
            assert np.all(par[mine] == -1)                    # no node is written twice
            par[mine] = j
        assert np.array_equal(seq, par), faces
