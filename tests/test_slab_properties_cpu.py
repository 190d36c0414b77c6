"""Property tests (hypothesis) of the slab-decomposition host logic: slab_range partitions the x-planes, halo_plan pairs up
neighbours consistently, ibm_box_participants covers every box plane exactly once with the owner of that plane."""
from hypothesis import given, settings, strategies as st

import pytest

import fsilbm3d_b200 as F
from fsilbm3d_b200.block_comm import halo_plan


@settings(max_examples=200, deadline=None)
@given(st.integers(1, 16), st.integers(1, 400))
def test_slab_range_partitions(nranks, extra):
    X = nranks + extra - 1
    spans = [F.slab_range(X, r, nranks) for r in range(nranks)]
    assert spans[0][0] == 0 and sum(c for _, c in spans) == X
    assert all(c >= 1 for _, c in spans)
    assert max(c for _, c in spans) - min(c for _, c in spans) <= 1          # balanced, remainders to the last ranks
    assert [c for _, c in spans] == sorted(c for _, c in spans)
    for (o0, c0), (o1, _) in zip(spans, spans[1:]):
        assert o0 + c0 == o1


@settings(max_examples=100, deadline=None)
@given(st.integers(1, 16), st.booleans())
def test_halo_plan_is_symmetric(nranks, periodic):
    for r in range(nranks):
        left, right, up, dn = halo_plan(r, nranks, periodic)
        if right >= 0:
            assert halo_plan(right, nranks, periodic)[0] == r      # my right neighbour's left neighbour is me
        if left >= 0:
            assert halo_plan(left, nranks, periodic)[1] == r
        if not periodic:
            assert (left == -1) == (r == 0) and (right == -1) == (r == nranks - 1)
        assert set(up).isdisjoint(dn) and len(up) == len(dn) == 5


@settings(max_examples=300, deadline=None)
@given(st.integers(1, 8), st.integers(0, 60), st.data())
def test_box_participants_cover_the_box(nranks, extra, data):
    X = nranks * 3 + extra
    slabs = [F.slab_range(X, r, nranks) for r in range(nranks)]
    x0 = data.draw(st.integers(0, X - 1))
    length = data.draw(st.integers(1, X))
    runs, leader = F.ibm_box_participants(x0, length, slabs, X)
    owner = {}
    for r, (off, cnt) in enumerate(slabs):
        for x in range(off, off + cnt):
            owner[x] = r
    assert runs[0][1] == 0 and runs[-1][2] == length and leader == owner[x0]
    for (r, d0, d1), nxt in zip(runs, runs[1:] + [None]):
        assert d0 < d1 and all(owner[(x0 + d) % X] == r for d in range(d0, d1))
        if nxt is not None:
            assert nxt[1] == d1 and nxt[0] != r                           # contiguous, maximal
    # a rank takes part iff it owns a plane of the box
    assert {r for r, _, _ in runs} == {owner[(x0 + d) % X] for d in range(length)}


def test_son_slab_plan_owner_and_remote_registrations():
    """Which rank creates a son and which neighbours register it (fsilbm_pair_create_remote): the cases of multi_rank_case.refinement_case
    and the limits."""
    import fsilbm3d_b200 as F
    slabs = [F.slab_range(40, r, 2) for r in range(2)]                     # planes 0..19 | 20..39
    assert F.son_slab_plan(0.0, 1.0, slabs, 5.0, 17, 0.5) == (0, [])       # planes 5..13: inside rank 0
    assert F.son_slab_plan(0.0, 1.0, slabs, 15.0, 17, 0.5) == (0, [1])     # planes 15..23: five on rank 0, four on rank 1
    assert F.son_slab_plan(0.0, 1.0, slabs, 16.0, 17, 0.5) == (1, [0])     # planes 16..24: four | five
    assert F.son_slab_plan(0.0, 1.0, slabs, 26.0, 17, 0.5) == (1, [])
    slabs8 = [F.slab_range(64, r, 8) for r in range(8)]                    # 8 planes each
    assert F.son_slab_plan(0.0, 1.0, slabs8, 20.0, 25, 0.5) == (3, [2, 4])   # planes 20..32: 4 | 8 | 1 -> owner 3, both neighbours
    with pytest.raises(ValueError, match="two neighbours"):
        F.son_slab_plan(0.0, 1.0, slabs8, 10.0, 61, 0.5)                   # planes 10..40: five ranks
    with pytest.raises(ValueError, match="not inside"):
        F.son_slab_plan(0.0, 1.0, slabs, 35.0, 17, 0.5)
    # every father plane of the footprint belongs to the owner or to a registered neighbour, for any placement that is accepted
    for xmin2 in range(0, 2 * 56):
        try:
            owner, remote = F.son_slab_plan(0.0, 1.0, slabs8, 0.5 * xmin2 if xmin2 % 2 == 0 else float(xmin2 // 2), 13, 0.5)
        except ValueError:
            continue
        assert all(abs(r - owner) == 1 for r in remote) and len(remote) <= 2
