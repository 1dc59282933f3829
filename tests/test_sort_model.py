"""The planned MSD equal-count sort (tools/sort_model.py) as an executable model: same permutation as
a stable sort by key -- i.e. as thrust::stable_sort_by_key (src/cuda/oibvhTree.cu:295-296) -- or an explicit fallback."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import sort_model  # noqa: E402
from oibvh_b200 import meshgen  # noqa: E402


def test_model_equals_stable_sort_on_mesh_keys(port):
    for pos, faces in (meshgen.blob(256, 192, seed=1234), meshgen.icosphere(5)):
        faces = meshgen.shuffle_faces(faces)
        keys = port.morton_keys(pos, faces, port.mesh_aabb(pos))
        perm = sort_model.msd_equal_count_sort(keys, n_chunks=37, window=1024, capacity=4096)
        assert perm is not None
        assert np.array_equal(perm, np.argsort(keys, kind="stable"))
        assert np.array_equal(perm, port.stable_sort_perm(keys))


def test_model_ties_small_inputs_and_fallback():
    rng = np.random.default_rng(3)
    # heavy ties: few distinct keys, must stay in input order inside a key
    keys = (rng.integers(0, 50, 20000).astype(np.uint32) << 14) | rng.integers(0, 3, 20000).astype(np.uint32)
    perm = sort_model.msd_equal_count_sort(keys, n_chunks=8, fine_bits=16, window=512, capacity=2048)
    assert perm is not None and np.array_equal(perm, np.argsort(keys, kind="stable"))
    for n in (1, 2, 3, 31, 1000):
        k = rng.integers(0, 1 << 30, n).astype(np.uint32)
        assert np.array_equal(sort_model.msd_equal_count_sort(k, n_chunks=5), np.argsort(k, kind="stable"))
    # a heavy fine bin (> capacity - window) is a range of its own; above the capacity the model asks for the fallback
    heavy = np.concatenate([rng.integers(0, 1 << 30, 3000).astype(np.uint32), np.full(3500, 123 << 14, np.uint32),
                            rng.integers(0, 1 << 30, 3000).astype(np.uint32)])
    rng.shuffle(heavy)
    perm = sort_model.msd_equal_count_sort(heavy, n_chunks=7, window=1024, capacity=4096)
    assert perm is not None and np.array_equal(perm, np.argsort(heavy, kind="stable"))
    plan = sort_model.plan_ranges(heavy, 16, 1024, 4096)
    sizes = np.bincount(plan[3][plan[0]])
    assert sizes.max() <= 4096 and 3500 in sizes
    assert sort_model.msd_equal_count_sort(np.full(5000, 123 << 14, np.uint32), window=1024, capacity=4096) is None


def test_transcribed_kernels_equal_stable_sort():
    """thread-level transcription of sort_msd.cu (tools/msd_kernel_emulation.py): plan == model, result == stable sort"""
    import msd_kernel_emulation as emu
    rng = np.random.default_rng(9)
    cases = [rng.integers(0, 1 << 30, 6000).astype(np.uint32),
             ((rng.integers(0, 30, 5000).astype(np.uint32) << 22) | rng.integers(0, 3, 5000).astype(np.uint32)),
             rng.integers(0, 1 << 30, 300).astype(np.uint32)]
    for keys in cases:
        hist = np.bincount(keys >> 14, minlength=emu.BINS)
        n, fallback, range_of_bin, range_start = emu.plan_kernel(hist, len(keys))
        fine, h, start, want_rob = sort_model.plan_ranges(keys, 16, emu.WIN, emu.CAP)
        assert not fallback and np.array_equal(range_of_bin[h > 0], want_rob[h > 0])
        k, v = emu.sort(keys, G=3)
        want = np.argsort(keys, kind="stable")
        assert np.array_equal(v, want) and np.array_equal(k, keys[want])
    assert emu.sort(np.full(9000, 77 << 14, np.uint32)) is None  # one bin above the range capacity: fallback


def test_transcribed_fallback_equals_stable_sort():
    """the 4-pass LSD path built into the experimental kernel (msd_lsd4_fallback), transcribed at thread level"""
    import msd_kernel_emulation as emu
    rng = np.random.default_rng(11)
    for keys, G in ((rng.integers(0, 1 << 30, 4000).astype(np.uint32), 3),
                    (np.full(2500, 77 << 14, np.uint32) | rng.integers(0, 4, 2500).astype(np.uint32), 2)):
        k, v = emu.lsd4_fallback(keys, G)
        want = np.argsort(keys, kind="stable")
        assert np.array_equal(v, want) and np.array_equal(k, keys[want])


def test_model_property_random_distributions():
    """hypothesis: for arbitrary key multisets the model either returns the stable-sort permutation or asks for the
    fallback, and every planned range respects the capacity"""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=60, deadline=None)
    @given(st.integers(min_value=1, max_value=4000), st.integers(min_value=1, max_value=30),
           st.integers(min_value=0, max_value=2 ** 31 - 1), st.sampled_from([(256, 1024), (512, 2048), (1024, 2048)]))
    def check(n, spread_bits, seed, wc):
        window, capacity = wc
        rng = np.random.default_rng(seed)
        keys = (rng.integers(0, 1 << spread_bits, n).astype(np.uint32) << np.uint32(30 - spread_bits))
        keys |= rng.integers(0, 1 << max(1, 30 - spread_bits), n).astype(np.uint32) & np.uint32(rng.integers(0, 1 << 14))
        keys &= np.uint32((1 << 30) - 1)
        perm = sort_model.msd_equal_count_sort(keys, n_chunks=int(rng.integers(1, 9)), window=window, capacity=capacity)
        hist = np.bincount(keys >> 14, minlength=1 << 16)
        if perm is None:
            assert hist.max() > capacity
            return
        assert np.array_equal(perm, np.argsort(keys, kind="stable"))
        plan = sort_model.plan_ranges(keys, 16, window, capacity)
        assert np.bincount(plan[3][plan[0]]).max() <= capacity

    check()
