"""Executable model (numpy) of the sort planned in DESIGN.md §6.1: ONE stable global partition by equal-count ranges
of a fine key histogram, then range-local stable sorts that fit shared memory. Not product code: it pins down the
range rule, the stability argument and the fallback condition before the CUDA version is written, and
tests/test_sort_model.py checks it against numpy's stable argsort on the bench meshes' Morton keys.

  fine bin  b = key >> (30 - FINE_BITS)                       (histogram computed together with the keys)
  range     consecutive fine bins; a bin opens a new range when it STARTS in another output window (start // WINDOW,
            start = exclusive prefix of the histogram) than the bin before it, or when it or the bin before it is heavy
            (> CAPACITY - WINDOW keys: a heavy bin is a range of its own). Ranges are contiguous in the output and
            never exceed CAPACITY; range ids are an inclusive scan of the "opens a range" flags
  fallback  a single fine bin above CAPACITY  ->  the present 4-pass global sort (returned as None here)
  partition chunk c of the input (one CTA) counts its keys per range; offsets = range start + sum of the counts of the
            chunks before c (the counts-matrix row scan of today's kernel); keys keep their input order inside a
            (chunk, range) cell, so the partition is stable
  local     every range is sorted stably by the full key inside one CTA (shared memory), in place
"""
import numpy as np

KEY_BITS = 30


def plan_ranges(keys, fine_bits=16, window=4096, capacity=8192):
    """-> (fine bin of every key, histogram, bin starts, range id of every bin) or None when a single fine bin exceeds
    the capacity (fallback). A bin starts a new range when it is heavy (> capacity - window keys), when the bin before
    it was heavy, or when it starts in another output window than the bin before it; range ids are the inclusive scan
    of these flags -- all data-parallel over the 2^fine_bits bins. Light ranges hold < window + (capacity - window)
    keys, heavy bins are ranges of their own."""
    fine = (keys >> (KEY_BITS - fine_bits)).astype(np.int64)
    hist = np.bincount(fine, minlength=1 << fine_bits)
    start = np.cumsum(hist) - hist
    if hist.max() > capacity:
        return None
    # rules are stated on the NON-EMPTY bins (on the device: a scan that carries "last non-empty bin" along)
    ne = np.flatnonzero(hist > 0)
    heavy = hist[ne] > capacity - window
    win = start[ne] // window
    opens = np.ones(len(ne), bool)
    opens[1:] = heavy[1:] | heavy[:-1] | (win[1:] != win[:-1])
    range_of_bin = np.zeros(len(hist), np.int64)
    range_of_bin[ne] = np.cumsum(opens) - 1
    return fine, hist, start, range_of_bin


def msd_equal_count_sort(keys, n_chunks=296, fine_bits=16, window=4096, capacity=8192):
    """-> permutation (sorted position -> input index) equal to a stable sort by key, or None (fallback needed)"""
    keys = np.asarray(keys, np.uint32)
    T = len(keys)
    plan = plan_ranges(keys, fine_bits, window, capacity)
    if plan is None:
        return None
    fine, hist, start, range_of_bin = plan
    rng = range_of_bin[fine]                       # range id of every key
    n_ranges = int(rng.max()) + 1 if T else 0
    # range extents in the output
    range_count = np.bincount(rng, minlength=n_ranges)
    range_start = np.cumsum(range_count) - range_count
    assert range_count.max() <= capacity, "range rule violated"
    # ---- stable global partition, chunk by chunk (what the CTAs do with the counts matrix) ----
    chunk = -(-T // n_chunks)
    counts = np.zeros((n_ranges, n_chunks), np.int64)
    for c in range(n_chunks):
        seg = rng[c * chunk:(c + 1) * chunk]
        if len(seg):
            counts[:, c] = np.bincount(seg, minlength=n_ranges)
    offs = range_start[:, None] + np.cumsum(counts, axis=1) - counts      # [range, chunk] -> first output slot
    part = np.empty(T, np.int64)                                           # output slot -> input index
    for c in range(n_chunks):
        lo, hi = c * chunk, min((c + 1) * chunk, T)
        if lo >= hi:
            continue
        seg = rng[lo:hi]
        order = np.argsort(seg, kind="stable")                             # in-chunk stable ranking by range
        sorted_rng = seg[order]
        first = np.searchsorted(sorted_rng, sorted_rng, side="left")       # rank inside the (chunk, range) cell
        slot = offs[sorted_rng, c] + (np.arange(hi - lo) - first)
        part[slot] = lo + order
    # ---- range-local stable sorts (one CTA each, shared memory) ----
    perm = np.empty(T, np.int64)
    for r in range(n_ranges):
        a, b = range_start[r], range_start[r] + range_count[r]
        if a == b:
            continue
        ids = part[a:b]
        perm[a:b] = ids[np.argsort(keys[ids], kind="stable")]
    return perm


if __name__ == "__main__":
    import os
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    import oracle
    P = oracle.Port()
    pos, faces = bench.make_meshes()
    keys = P.morton_keys(pos, faces, P.mesh_aabb(pos))
    perm = msd_equal_count_sort(keys)
    plan = plan_ranges(keys, 16, 4096, 8192)
    if plan is not None:
        rng = plan[3][plan[0]]
        cnt = np.bincount(rng)
        print("ranges %d, keys per range: mean %.0f max %d; heaviest fine bin %d" % ((cnt > 0).sum(), cnt[cnt > 0].mean(), cnt.max(), plan[1].max()))
    print("fallback" if perm is None else ("equal to stable argsort: %s" % np.array_equal(perm, np.argsort(keys, kind="stable"))))
