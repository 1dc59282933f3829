"""Multi-GPU detection through the C ABI (oibvh_mgpu_*): every rank appends its hits to rank 0's pair list through a
peer mapping; the union must equal the single-GPU pair set. The protocol runs on ONE device too (several contexts
play the ranks one after the other), which is what the single-GPU test box exercises; the tests that need two
devices / two processes skip themselves there (run them with `gpurun --gpus 2`)."""
import multiprocessing as mp
import os

import numpy as np
import pytest

import oibvh_b200 as ob
from oibvh_b200 import meshgen
from test_gpu_collide import oracle_pairs

pytestmark = pytest.mark.gpu


def _meshes(port, seed=10, nu=80, nv=64):
    pos, faces = meshgen.blob(nu, nv, seed=seed)
    faces = meshgen.shuffle_faces(faces, seed=seed + 1)
    posB = port.transform_positions(pos, ob.mat_translate(ob.mat_identity(), (0.7, 0.2, 0.1)))
    return [(pos, faces), (posB, faces)]


def _rank_scene(device, meshes, rank, world):
    """a rank = its own context, its own replica of every tree, its own scene"""
    ctx = ob.Context(device)
    sc = ob.Scene(ctx)
    trees = []
    for pos, faces in meshes:
        t = ob.OibvhTree(ob.Mesh(pos, faces), ctx=ctx)
        t.build()
        sc.addOibvhTree(t)
        trees.append(t)
    sc.set_shard(rank, world)
    return ctx, sc, trees


@pytest.mark.parametrize("world", [2, 3, 8])
def test_ranks_on_one_device_gather_into_rank0(port, world):
    """every rank's narrow phase writes into rank 0's list; rank 0's detection returns the union"""
    meshes = _meshes(port)
    want, _ = oracle_pairs(port, meshes)
    ranks = [_rank_scene(0, meshes, r, world) for r in range(world)]
    root_ctx, root, _ = ranks[0]
    root.reserve(pair_records=1 << 17)
    handle = root.mgpu_export()
    assert len(handle) == 160
    for _, sc, _ in ranks[1:]:
        sc.mgpu_attach(handle)
    local_total = 0
    for frame in range(3):  # the protocol words are frame-numbered: several frames in a row
        root.mgpu_open_frame()
        root_ctx.synchronize()
        local = []
        for ctx, sc, _ in ranks[1:]:
            sc.detect_async(4, 3)
            local.append(sc.counts()[0])  # a remote rank reports its own hit count
            with pytest.raises(ob.OibvhError):
                sc.m_intTriPairs  # ... but the list lives on rank 0
        root.detect_async(4, 3)
        n, _ = root.counts()
        assert n == len(want), f"frame {frame}: gathered {n} pairs, want {len(want)}"
        got = root.canonical_pairs()
        assert np.array_equal(got, want), f"frame {frame}"
        local_total = sum(local)
        assert 0 < local_total < n  # the remote ranks contributed, and so did rank 0
    for _, sc, _ in ranks:
        sc.mgpu_detach()
    # back in single-GPU mode the same scenes work on their own again
    root.set_shard(0, 1)
    root.detect_async(4, 3)
    assert np.array_equal(root.canonical_pairs(), want)


def test_a_missing_rank_is_reported_not_hung(port):
    """rank 1 never launches its frame: rank 0's bounded wait times out and get_counts fails loudly"""
    meshes = _meshes(port, nu=32, nv=24)
    ranks = [_rank_scene(0, meshes, r, 2) for r in range(2)]
    root = ranks[0][1]
    handle = root.mgpu_export()
    ranks[1][1].mgpu_attach(handle)
    root.detect_async(4, 3)
    with pytest.raises(ob.OibvhError):
        root.counts()


def test_multi_gpu_queues_are_fixed(port):
    """a multi-GPU scene cannot regrow its pair list (other ranks map it): overflow is an error, reserve first"""
    pos, faces = meshgen.blob(320, 200, seed=9)
    meshes = [(pos, faces), (pos.copy(), faces.copy())]  # coincident meshes: far more than 2^19 pairs
    ranks = [_rank_scene(0, meshes, r, 2) for r in range(2)]
    root_ctx, root, _ = ranks[0]
    handle = root.mgpu_export()
    ranks[1][1].mgpu_attach(handle)
    with pytest.raises(ob.OibvhError):
        root.reserve(pair_records=1 << 22)
    root.mgpu_open_frame()
    root_ctx.synchronize()
    ranks[1][1].detect_async(4, 3)
    ranks[1][0].synchronize()
    root.detect_async(4, 3)
    with pytest.raises(ob.OibvhError):
        root.counts()


def test_two_devices_one_process(port):
    """ranks on two GPUs of one process (peer access instead of IPC), frames enqueued without host sync in between"""
    if ob.device_count() < 2:
        pytest.skip("needs two CUDA devices")
    meshes = _meshes(port, nu=160, nv=128)
    want, _ = oracle_pairs(port, meshes)
    ranks = [_rank_scene(r, meshes, r, 2) for r in range(2)]
    root = ranks[0][1]
    root.reserve(pair_records=1 << 18)
    ranks[1][1].mgpu_attach(root.mgpu_export())
    for frame in range(4):
        ranks[1][1].detect_async(4, 3)  # waits on the device for rank 0 to open the frame
        root.detect_async(4, 3)         # opens the frame, does its share, waits on the device for rank 1's DONE
        assert np.array_equal(root.canonical_pairs(), want), f"frame {frame}"
    for _, sc, _ in ranks:
        sc.mgpu_detach()


def _ipc_worker(rank, world, handle_q, go_q, done_q, frames):
    import oracle
    port = oracle.Port()
    meshes = _meshes(port, nu=160, nv=128)
    ctx, sc, _ = _rank_scene(rank, meshes, rank, world)
    sc.mgpu_attach(handle_q.get(timeout=120))
    done_q.put(("attached", rank))
    for _ in range(frames):
        go_q.get(timeout=120)
        sc.detect_async(4, 3)
        done_q.put(("frame", rank, sc.counts()[0]))
    go_q.get(timeout=120)
    sc.mgpu_detach()
    done_q.put(("detached", rank))


def test_two_processes_cuda_ipc(port):
    """one process per GPU: the handle crosses the process boundary as plain bytes (CUDA IPC underneath)"""
    if ob.device_count() < 2:
        pytest.skip("needs two CUDA devices")
    meshes = _meshes(port, nu=160, nv=128)
    want, _ = oracle_pairs(port, meshes)
    ctx0, root, _ = _rank_scene(0, meshes, 0, 2)
    root.reserve(pair_records=1 << 18)
    handle = root.mgpu_export()
    mpc = mp.get_context("spawn")
    handle_q, go_q, done_q = mpc.Queue(), mpc.Queue(), mpc.Queue()
    frames = 3
    p = mpc.Process(target=_ipc_worker, args=(1, 2, handle_q, go_q, done_q, frames))
    p.start()
    try:
        handle_q.put(handle)
        assert done_q.get(timeout=300)[0] == "attached"
        for frame in range(frames):
            go_q.put(1)
            root.detect_async(4, 3)
            n, _ = root.counts()
            msg = done_q.get(timeout=120)
            assert msg[0] == "frame" and 0 < msg[2] < n
            assert np.array_equal(root.canonical_pairs(), want), f"frame {frame}"
        go_q.put(1)
        assert done_q.get(timeout=120)[0] == "detached"
        root.mgpu_detach()
    finally:
        p.join(60)
        if p.is_alive():
            p.kill()
    assert p.exitcode == 0


@pytest.mark.parametrize("world", [2, 3])
def test_cpp_host_threads_drive_the_mgpu_abi(world):
    """tests/cpp/mgpu_threads: one C++ host thread per rank through the C ABI (replicated trees, peer-mapped pair list,
    frames without host synchronisation between the ranks). With more ranks than devices (a single-GPU box) the ranks
    share a device and the test orders the frames on the host with oibvh_mgpu_open_frame."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "tests", "cpp", "mgpu_threads")
    if not os.path.exists(exe):
        subprocess.check_call(["make", "-s", "-C", os.path.join(root, "tests", "cpp")])
    res = subprocess.run([exe, str(world), "6"], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout + res.stderr
    first = res.stdout.splitlines()[0].split()
    assert first[:5] == ["ok", "ranks", str(world), "frames", "6"] and int(first[6]) > 0, res.stdout
