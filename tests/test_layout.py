"""Implicit-layout arithmetic: the (level, position) addressing used by the CUDA kernels
(oibvh_b200/csrc/common.cuh: level_count / level_offset / tree_size) against the reference's
implicit<->real mapping (include/cuda/oibvh.cuh:56-182) as restated in the oracle and, when built,
against the reference's own host functions."""
import numpy as np
import pytest

SIZES = [2, 3, 4, 5, 6, 7, 8, 9, 12, 13, 31, 32, 33, 100, 255, 256, 257, 1000, 1023, 1024, 1025, 4097, 69451]


def ceil_log2(t):
    l = 0
    while (1 << l) < t:
        l += 1
    return l


def level_count(T, L, l):
    s = L - l
    return (T + (1 << s) - 1) >> s


def level_offset(T, L, l):
    """closed form used on the device"""
    if l == 0:
        return 0
    vl = (1 << L) - T
    v = vl >> (L - l + 1)
    return (1 << l) - 1 - 2 * v + bin(v).count("1")


def test_closed_form_offset_is_prefix_sum_of_counts():
    for T in SIZES + [2 ** 20, 10 ** 6, 4 * 2 ** 20 + 7, 2 ** 24 + 12345]:
        L = ceil_log2(T)
        run = 0
        for l in range(L + 1):
            assert level_offset(T, L, l) == run, (T, l)
            run += level_count(T, L, l)
        assert level_count(T, L, L) == T and level_count(T, L, 0) == 1


@pytest.mark.parametrize("T", SIZES)
def test_level_addressing_matches_reference_mapping(port, T):
    L = ceil_log2(T)
    vl = (1 << L) - T
    N = port.get_size(T)
    assert level_offset(T, L, L) + T == N  # tree_size
    step = 1 if N < 5000 else 97
    for l in range(L + 1):
        cnt = level_count(T, L, l)
        assert cnt == port.level_real_count(l, L, vl)
        off = level_offset(T, L, l)
        for p in list(range(0, cnt, step)) + [cnt - 1]:
            implicit = (1 << l) - 1 + p
            assert port.implicit_to_real(implicit, L, vl) == off + p
            assert port.real_to_implicit(off + p, L, vl) == implicit
            if l < L:  # right child kept <=> 2p+1 < cnt(l+1)
                assert port.have_rchild(implicit, L, vl) == (2 * p + 1 < level_count(T, L, l + 1))
        assert port.most_right_valid(l, L, vl) == (1 << l) - 1 + cnt - 1


@pytest.mark.parametrize("T", [2, 3, 5, 13, 257, 1000, 4097])
def test_port_layout_equals_unmodified_reference(port, ref, T):
    L = ceil_log2(T)
    vl = (1 << L) - T
    assert port.get_size(T) == ref.lib.ref_oibvh_get_size(T)
    N = port.get_size(T)
    for r in range(N):
        i = port.real_to_implicit(r, L, vl)
        assert i == ref.lib.ref_oibvh_real_to_implicit(r, L, vl)
        if i > 0:  # the reference's root case relies on a shift wrap
            assert ref.lib.ref_oibvh_implicit_to_real(i, L, vl) == r
        lev = (i + 1).bit_length() - 1
        if lev < L:
            assert port.have_rchild(i, L, vl) == bool(ref.lib.ref_oibvh_have_rchild(i, L, vl))


def test_layout_properties_random_sizes(port):
    """hypothesis: for random T up to 2^26 (the node-position field of a BVTT record) and random (level, position),
    the device's closed-form addressing equals the restated reference mapping, both directions"""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=300, deadline=None)
    @given(st.integers(min_value=2, max_value=1 << 26), st.data())
    def check(T, data):
        L = ceil_log2(T)
        vl = (1 << L) - T
        N = port.get_size(T)
        assert N == 2 * T - 1 + bin(vl).count("1") == level_offset(T, L, L) + T
        l = data.draw(st.integers(min_value=0, max_value=L))
        cnt = level_count(T, L, l)
        assert cnt == port.level_real_count(l, L, vl)
        p = data.draw(st.integers(min_value=0, max_value=cnt - 1))
        implicit = (1 << l) - 1 + p
        real = level_offset(T, L, l) + p
        assert port.implicit_to_real(implicit, L, vl) == real
        assert port.real_to_implicit(real, L, vl) == implicit
        if l < L:
            assert port.have_rchild(implicit, L, vl) == (2 * p + 1 < level_count(T, L, l + 1))
            # the children of (l, p) are (l+1, 2p) and, when kept, (l+1, 2p+1): parents are never virtual
            assert 2 * p < level_count(T, L, l + 1)

    check()
