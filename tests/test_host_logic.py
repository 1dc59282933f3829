"""Host-side logic of the package that needs no GPU: the Mesh mirror (matrices, m_aabb, m_center, transform)
against golden values from the reference's glm/Mesh code, mesh generators, and the multi-rank pair-list
gather (world_size 2 over gloo)."""
import os
import socket

import numpy as np
import pytest

import oibvh_b200 as ob
from conftest import assert_bit_equal
from oibvh_b200 import meshgen


def test_mesh_matches_reference_mesh(golden):
    g = golden["transforms"]
    m = ob.Mesh(g["pos0"], g["faces"])
    assert_bit_equal(m.m_aabb, g["aabb0"], "m_aabb")
    assert_bit_equal(m.m_center, g["center0"], "m_center")
    steps = [("rot", (0, 0, 1), 1.0), ("rot", (1, 0, 0), 1.0), ("tr", (1.0, 0.0, 0.0)), ("rot", (0.3, -0.5, 0.8), 37.5),
             ("tr", (-0.25, 0.125, 3.0))]
    for i, st in enumerate(steps):
        if st[0] == "rot":
            M = m.transform_matrix_rotate(st[1], st[2])
        else:
            M = m.transform_matrix_translate(st[1])
        # sin/cos come from numpy here and from libm in glm: allow the matrix 1 ulp, then apply the GOLDEN matrix
        np.testing.assert_allclose(M, g[f"M{i}"], rtol=3e-7, atol=3e-7)
        m.transform(g[f"M{i}"])
        assert_bit_equal(m.m_positions, g[f"pos{i + 1}"], f"positions after step {i}")
        assert_bit_equal(m.m_center, g[f"center{i + 1}"], f"center after step {i}")
    assert_bit_equal(m.m_aabb, g["aabb0"], "m_aabb is never updated by transforms (mesh.cpp:91-98)")


def test_translate_matrix_is_exact(golden):
    g = golden["transforms"]
    assert_bit_equal(ob.mat_translate(ob.mat_identity(), (1.0, 0.0, 0.0)), g["M2"], "glm::translate")


def test_generators_are_deterministic_and_sized():
    p, f = meshgen.blob(1024, 512)
    assert f.shape == (2 ** 20, 3) and p.shape == (513 * 1024, 3) and p.dtype == np.float32 and f.dtype == np.uint32
    p2, f2 = meshgen.blob(1024, 512)
    assert np.array_equal(p, p2) and np.array_equal(f, f2)
    assert int(f.max()) < len(p)
    p, f = meshgen.icosphere(3)
    assert f.shape == (20 * 4 ** 3, 3) and len(p) == 10 * 4 ** 3 + 2
    p, f = meshgen.terrain(64, 48)
    assert f.shape == (2 * 64 * 48, 3) and int(f.max()) == len(p) - 1
    p, f = meshgen.cube()
    assert f.shape == (12, 3)
    p, f = meshgen.quad()
    assert f.shape == (2, 3) and (p[:, 2] == 0).all()


def test_shard_prefix():
    off, tot = ob.shard_prefix([3, 0, 5, 2])
    assert off == [0, 3, 3, 8] and tot == 10


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _gather_worker(rank, world, port_no, q):
    import torch
    import torch.distributed as dist
    from oibvh_b200 import distributed as obd
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port_no)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # rank r owns r+2 pair records whose first field encodes the owner
        n = rank + 2
        local = torch.arange(n * 4, dtype=torch.int32).reshape(n, 4) + 1000 * rank
        full = obd.gather_pairs(local, n)
        q.put((rank, full.numpy().copy()))
    finally:
        dist.destroy_process_group()


def test_gather_pairs_world2_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port_no = _free_port()
    procs = [ctx.Process(target=_gather_worker, args=(r, 2, port_no, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    want = np.concatenate([np.arange(8).reshape(2, 4), np.arange(12).reshape(3, 4) + 1000]).astype(np.int32)
    for r in range(2):
        assert np.array_equal(out[r], want)


def _block_worker(rank, world, port_no, q):
    import torch
    import torch.distributed as dist
    from oibvh_b200 import distributed as obd
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port_no)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        cap = 8
        # what a scene's device range looks like: 32-record counter block (row 0 = cand, pairs, flags, -) + pair list
        n = [3, 11][rank]  # rank 1 overflows the fixed exchange
        block = torch.full((obd.HEAD_RECORDS + cap, 4), -7, dtype=torch.int32)
        block[0] = torch.tensor([5 * n, n, 0, 0], dtype=torch.int32)
        m = min(n, cap)
        block[obd.HEAD_RECORDS:obd.HEAD_RECORDS + m] = torch.arange(m * 4, dtype=torch.int32).reshape(m, 4) + 1000 * rank
        out = torch.empty((world * (obd.HEAD_RECORDS + cap), 4), dtype=torch.int32)
        obd.gather_blocks(out, block)
        counts, parts, truncated = obd.unpack_blocks(out, world, cap)
        q.put((rank, counts, [p.numpy().copy() for p in parts], truncated))
    finally:
        dist.destroy_process_group()


def test_single_collective_block_exchange_world2_gloo():
    """[counter block | pair list] of every rank in ONE all-gather (the multi-GPU frame's only exchange)"""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port_no = _free_port()
    procs = [ctx.Process(target=_block_worker, args=(r, 2, port_no, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    for rank, counts, parts, truncated in res:
        assert counts == [3, 11] and truncated  # rank 1 holds 11 > cap = 8 records: reported, never silently dropped
        assert np.array_equal(parts[0], np.arange(12).reshape(3, 4))
        assert np.array_equal(parts[1], np.arange(32).reshape(8, 4) + 1000)


def _cap_worker(rank, world, port_no, q):
    import torch
    import torch.distributed as dist
    from oibvh_b200 import distributed as obd
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port_no)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n_local = [700, 1100, 90][rank]          # 4x -> 2800 / 4400 / 360: a rank-local rule would give 4096 / 8192 / 4096
        ceiling = [1 << 19, 1 << 19, 1 << 20][rank]
        cap = obd.agree_capacity(n_local, floor=4096, ceiling=ceiling)
        cap_small = obd.agree_capacity(n_local, floor=4096, ceiling=[6000, 5000, 7000][rank])
        # the exchange itself with the agreed size, world = 3
        block = torch.full((obd.HEAD_RECORDS + cap, 4), rank, dtype=torch.int32)
        block[0] = torch.tensor([0, n_local, 0, 0], dtype=torch.int32)
        out = torch.empty((world * (obd.HEAD_RECORDS + cap), 4), dtype=torch.int32)
        obd.gather_blocks(out, block)
        counts, parts, truncated = obd.unpack_blocks(out, world, cap)
        q.put((rank, cap, cap_small, counts, [int(p[0, 0]) for p in parts], truncated))
    finally:
        dist.destroy_process_group()


def test_exchange_capacity_is_agreed_across_ranks_world3_gloo():
    """regression: the fixed exchange size must not depend on the rank's own pair count (it hung at 4 GPUs)"""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port_no = _free_port()
    procs = [ctx.Process(target=_cap_worker, args=(r, 3, port_no, q)) for r in range(3)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(3)]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    for rank, cap, cap_small, counts, firsts, truncated in res:
        assert cap == 8192 and cap_small == 5000          # MAX of the counts, MIN of the ceilings: same on every rank
        assert counts == [700, 1100, 90] and firsts == [0, 1, 2] and not truncated
