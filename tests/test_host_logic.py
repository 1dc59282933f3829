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


class _FakeScene:
    """records what distributed.attach / detach do to a scene (the real calls need a GPU)"""

    class _Ctx:
        def synchronize(self):
            pass

    def __init__(self, rank):
        self.rank_id, self.calls, self.ctx = rank, [], self._Ctx()

    def set_shard(self, rank, world):
        self.calls.append(("set_shard", rank, world))

    def mgpu_export(self):
        self.calls.append(("export",))
        return bytes(range(160))

    def mgpu_attach(self, handle):
        self.calls.append(("attach", bytes(handle)))

    def mgpu_detach(self):
        self.calls.append(("detach",))


def _attach_worker(rank, world, port_no, q):
    import torch.distributed as dist
    from oibvh_b200 import distributed as obd
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port_no)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sc = _FakeScene(rank)
        obd.attach(sc, rank, world)
        obd.detach(sc)
        q.put((rank, sc.calls))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_mgpu_handle_transport_gloo(world):
    """set-up of the peer-mapped pair list: rank 0 exports, the 160-byte handle travels by broadcast, every other rank
    attaches exactly those bytes; nothing else is exchanged (the frames themselves need no collective)"""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port_no = _free_port()
    procs = [ctx.Process(target=_attach_worker, args=(r, world, port_no, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=180) for _ in range(world))
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert res[0] == [("set_shard", 0, world), ("export",), ("detach",)]
    for r in range(1, world):
        assert res[r] == [("set_shard", r, world), ("attach", bytes(range(160))), ("detach",)]
