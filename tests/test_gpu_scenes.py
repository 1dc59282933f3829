"""Scene-level parity at the shapes of BASELINE.json configs 3-5 (reduced so the CPU oracle finishes in seconds):
deforming mesh vs static obstacle (refit only), many small bodies, large flat terrain vs a mesh."""
import numpy as np
import pytest

import oibvh_b200 as ob
import oracle
from conftest import assert_bit_equal
from oibvh_b200 import meshgen

pytestmark = pytest.mark.gpu


def oracle_scene(port, meshes):
    built = [port.build(p, f) for p, f in meshes]
    pairs, ncand = port.detect([(b["nodes"], b["faces"], p) for b, (p, f) in zip(built, meshes)])
    return oracle.canonical_pairs(pairs, [b["perm"] for b in built]), ncand, built


def test_deforming_mesh_vs_static_obstacle(ctx, port):
    """configs[2] shape: per-frame refit (no rebuild) of a deforming sphere against a static icosphere"""
    pos, faces = meshgen.uv_sphere(384, 256)                     # 196 608 triangles, deforms
    faces = meshgen.shuffle_faces(faces)
    opos, ofaces = meshgen.icosphere(5, radius=0.6, center=(0.9, 0.2, 0.1))  # 20 480 triangles, static
    cloth = ob.OibvhTree(ob.Mesh(pos, faces), ctx=ctx)
    cloth.build()
    rock = ob.OibvhTree(ob.Mesh(opos, ofaces), ctx=ctx)
    rock.build()
    sc = ob.Scene(ctx)
    sc.addOibvhTree(cloth)
    sc.addOibvhTree(rock)
    oc = port.build(pos, faces)
    orock = port.build(opos, ofaces)
    seen = set()
    for frame in range(3):
        p2 = meshgen.cloth_positions(pos, frame, amp=0.08)
        cloth.set_positions(p2)
        cloth.refit(upload=False)
        sc.detectCollision(ob.DeviceType.GPU0, 4, 0)
        nodes = port.refit(p2, oc["faces"])
        assert_bit_equal(cloth.m_aabbTree, nodes, f"refit frame {frame}")
        pp, nc = port.detect([(nodes, oc["faces"], p2), (orock["nodes"], orock["faces"], opos)])
        want = oracle.canonical_pairs(pp, [oc["perm"], orock["perm"]])
        assert sc.getCandidateCount() == nc
        assert np.array_equal(sc.canonical_pairs(), want), f"frame {frame}"
        seen.add(len(want))
    assert min(seen) > 100 and len(seen) > 1


def test_many_bodies(ctx, port):
    """configs[3] shape: 96 small bodies (more than the 64-entry shared-memory object cache, well past nothing like
    the reference's 256-object limit matters here) scattered in a box so that a few percent touch"""
    rng = np.random.default_rng(11)
    meshes = []
    for k in range(96):
        c = rng.uniform(-2.0, 2.0, size=3)
        if k % 3 == 0:
            p, f = meshgen.cube(0.35, c)
        elif k % 3 == 1:
            p, f = meshgen.icosphere(2, radius=0.4, center=c)
        else:
            p, f = meshgen.blob(12, 9, seed=k, radius=0.4, center=c)
            f = f[:len(f) - (k % 7) - 1]
        meshes.append((p, np.ascontiguousarray(f)))
    want, ncand, _ = oracle_scene(port, meshes)
    trees = []
    sc = ob.Scene(ctx)
    for p, f in meshes:
        t = ob.OibvhTree(ob.Mesh(p, f), ctx=ctx)
        t.build()
        sc.addOibvhTree(t)
        trees.append(t)
    for entry, expand in [(4, 3), (0, 0), (2, 1)]:
        sc.detectCollision(ob.DeviceType.GPU0, entry, expand)
        assert sc.getCandidateCount() == ncand
        assert np.array_equal(sc.canonical_pairs(), want)
    touching = {tuple(r) for r in want[:, :2].tolist()}
    assert len(touching) >= 10
    assert max(max(a, b) for a, b in touching) >= 64  # pairs beyond the cached object range are exercised


def test_terrain_vs_mesh(ctx, port):
    """configs[4] shape: a large height field (very different depth from the body pressed into it)"""
    tpos, tfaces = meshgen.terrain(512, 384, height=0.25)            # 393 216 triangles, L = 19
    bpos, bfaces = meshgen.blob(96, 64, seed=5, radius=0.8, center=(0.3, 0.1, -0.2))  # 12 288 triangles, L = 14
    meshes = [(tpos, meshgen.shuffle_faces(tfaces)), (bpos, bfaces)]
    want, ncand, _ = oracle_scene(port, meshes)
    sc = ob.Scene(ctx)
    for p, f in meshes:
        t = ob.OibvhTree(ob.Mesh(p, f), ctx=ctx)
        t.build()
        sc.addOibvhTree(t)
    sc.detectCollision(ob.DeviceType.GPU0, 4, 0)
    assert sc.getCandidateCount() == ncand
    assert np.array_equal(sc.canonical_pairs(), want)
    assert len(want) > 500


def test_streaming_sort_path_beyond_single_wave(ctx, port):
    """T above the cooperative sort's capacity takes the onesweep (look-back) kernels: bit-exact as well"""
    pos, faces = meshgen.blob(1200, 640, seed=8)  # 1 536 000 triangles
    faces = meshgen.shuffle_faces(faces)
    t = ob.OibvhTree(ob.Mesh(pos, faces), ctx=ctx)
    t.build()
    o = port.build(pos, faces)
    d = t.download()
    assert np.array_equal(t.sorted_keys(), o["keys"])
    assert np.array_equal(d["perm"], o["perm"])
    assert_bit_equal(d["nodes"], o["nodes"], "1.5M build")


def test_six_hundred_cubes(ctx, port):
    """well past the reference's 256-object limit (collide.cu:91-93): 600 cubes, 179 700 object pairs seeded on the
    device, most pruned by the root test in round 0"""
    rng = np.random.default_rng(3)
    meshes = []
    for k in range(600):
        c = rng.uniform(-4.0, 4.0, size=3)
        p, f = meshgen.cube(0.3, c)
        meshes.append((p, f))
    want, ncand, _ = oracle_scene(port, meshes)
    sc = ob.Scene(ctx)
    keep = []
    for p, f in meshes:
        t = ob.OibvhTree(ob.Mesh(p, f), ctx=ctx)
        t.build()
        sc.addOibvhTree(t)
        keep.append(t)
    sc.detectCollision(ob.DeviceType.GPU0, 4, 0)
    assert sc.getCandidateCount() == ncand
    assert np.array_equal(sc.canonical_pairs(), want)
    assert len({tuple(r) for r in want[:, :2].tolist()}) > 50
    # shards of a seeded (many-body) scene partition the pair set too
    parts = []
    for rank in range(3):
        sc.set_shard(rank, 3)
        sc.detectCollision(ob.DeviceType.GPU0, 4, 0)
        parts.append(sc.canonical_pairs())
    allp = np.concatenate(parts)
    assert len(allp) == len(want)
    allp = allp[np.lexsort((allp[:, 3], allp[:, 2], allp[:, 1], allp[:, 0]))]
    assert np.array_equal(allp, want)
