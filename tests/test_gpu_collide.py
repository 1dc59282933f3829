"""GPU parity, broad + narrow phase: colliding-pair SETS bit-exact after canonical sorting."""
import numpy as np
import pytest

import oibvh_b200 as ob
import oracle
from oibvh_b200 import meshgen
from conftest import assert_bit_equal

pytestmark = pytest.mark.gpu


def make_scene(ctx, meshes):
    """meshes: list of (pos, faces) already in world space"""
    trees = []
    sc = ob.Scene(ctx)
    for pos, faces in meshes:
        t = ob.OibvhTree(ob.Mesh(pos, faces), ctx=ctx)
        t.build()
        sc.addOibvhTree(t)
        trees.append(t)
    return sc, trees


def oracle_pairs(port, meshes):
    built = [port.build(p, f) for p, f in meshes]
    pairs, ncand = port.detect([(b["nodes"], b["faces"], p) for b, (p, f) in zip(built, meshes)])
    return oracle.canonical_pairs(pairs, [b["perm"] for b in built]), ncand


def test_golden_two_spheres(ctx, golden):
    """known answers frozen from the reference's SimpleCollide (and SURVEY.md Appendix A: 456 pairs at n=64)"""
    g = golden["collide"]
    for n in (16, 64):
        pos, faces, posB, want = (g[f"sphere{n}_{k}"] for k in ("pos", "faces", "posB", "pairs"))
        sc, _ = make_scene(ctx, [(pos, faces), (posB, faces)])
        sc.detectCollision(ob.DeviceType.GPU0, 4, 3)
        assert sc.getIntTriPairCount() == len(want)
        assert np.array_equal(sc.canonical_pairs(), want)
    assert sc.getIntTriPairCount() == 456 and sc.getCandidateCount() == 2093


def test_golden_three_bodies(ctx, golden):
    g = golden["collide"]
    meshes = [(g[f"body{k}_pos"], g[f"body{k}_faces"]) for k in range(3)]
    sc, _ = make_scene(ctx, meshes)
    sc.detectCollision(ob.DeviceType.GPU0, 4, 3)
    assert np.array_equal(sc.canonical_pairs(), g["bodies_pairs"])
    raw = sc.m_intTriPairs
    assert (raw[:, 0] < raw[:, 1]).all()  # bvhIndex[0] < bvhIndex[1]  (scene.cu:195-196)
    assert len(np.unique(raw, axis=0)) == len(raw)  # no duplicates


@pytest.mark.parametrize("entry,expand", [(4, 3), (0, 1), (0, 0), (1, 1), (2, 2), (7, 1), (12, 4), (30, 8)])
def test_pair_set_independent_of_entry_and_expand(ctx, port, entry, expand):
    pos, faces = meshgen.blob(72, 50, seed=3)
    faces = meshgen.shuffle_faces(faces)[:7001]
    posB = port.transform_positions(pos, ob.mat_translate(ob.mat_identity(), (0.8, 0.15, -0.1)))
    meshes = [(pos, faces), (posB, faces[:4099])]  # different depths: L = 13 and 13 / T not power of two
    want, ncand = oracle_pairs(port, meshes)
    sc, _ = make_scene(ctx, meshes)
    sc.detectCollision(ob.DeviceType.GPU0, entry, expand)
    assert sc.getCandidateCount() == ncand
    assert np.array_equal(sc.canonical_pairs(), want)
    assert len(want) > 50


def test_different_depths_and_tiny_trees(ctx, port):
    big = meshgen.blob(128, 96, seed=4)
    tiny = meshgen.cube(0.3, (0.9, 0.1, 0.0))
    two = (tiny[0], tiny[1][:2])  # T = 2, the minimum
    three = (meshgen.cube(0.25, (-0.8, 0.3, 0.2))[0], meshgen.cube()[1][:3])
    meshes = [big, tiny, two, three]
    want, ncand = oracle_pairs(port, meshes)
    sc, _ = make_scene(ctx, meshes)
    for entry, expand in [(4, 3), (0, 1)]:
        sc.detectCollision(ob.DeviceType.GPU0, entry, expand)
        assert sc.getCandidateCount() == ncand
        assert np.array_equal(sc.canonical_pairs(), want)
    assert len(want) > 0


def test_no_contact_and_single_object(ctx):
    a = meshgen.icosphere(2)
    b = meshgen.icosphere(2, center=(5.0, 0, 0))
    sc, _ = make_scene(ctx, [a, b])
    sc.detectCollision(ob.DeviceType.GPU0, 4, 3)
    assert sc.getIntTriPairCount() == 0 and sc.m_intTriPairs.shape == (0, 4)
    sc1, _ = make_scene(ctx, [a])
    sc1.detectCollision()
    assert sc1.getIntTriPairCount() == 0


def test_touching_and_degenerate_triangles(ctx, port):
    """identical meshes in place: every shared vertex / edge / coincident triangle is a touching case where an
    FMA-contracted SAT can flip; UV-sphere poles add zero-area triangles"""
    pos, faces = meshgen.uv_sphere(24)
    meshes = [(pos, faces), (pos.copy(), faces.copy())]
    want, ncand = oracle_pairs(port, meshes)
    sc, _ = make_scene(ctx, meshes)
    sc.detectCollision(ob.DeviceType.GPU0, 3, 2)
    assert sc.getCandidateCount() == ncand
    assert np.array_equal(sc.canonical_pairs(), want)
    assert len(want) > len(faces)


def test_refit_between_detections_is_honoured(ctx, port):
    """scene.cu:264-268 re-reads every tree on each detectCollision; here trees are referenced, so a refit shows"""
    pos, faces = meshgen.blob(60, 40, seed=7)
    mA, mB = ob.Mesh(pos, faces), ob.Mesh(pos, faces)
    tA = ob.OibvhTree(mA, ctx=ctx)
    tA.build()
    tB = ob.OibvhTree(tA, mB)
    sc = ob.Scene(ctx)
    sc.addOibvhTree(tA)
    sc.addOibvhTree(tB)
    oa = port.build(pos, faces, mA.m_aabb)
    counts = []
    for frame in range(4):
        mB.rotateX(1.0)  # main.cpp:248-252
        mB.translate((0.3, 0.0, 0.0))
        tB.refit()
        sc.detectCollision(ob.DeviceType.GPU0, 4, 3)
        nodesB = port.refit(mB.m_positions, oa["faces"])
        pp, _ = port.detect([(oa["nodes"], oa["faces"], pos), (nodesB, oa["faces"], mB.m_positions)])
        want = oracle.canonical_pairs(pp, [oa["perm"], oa["perm"]])
        assert np.array_equal(sc.canonical_pairs(), want), f"frame {frame}"
        counts.append(len(want))
    assert len(set(counts)) > 1


def test_queue_growth_on_overflow(ctx, port):
    """a dense contact (coincident meshes) overflows the initial 1M-entry queues: they must regrow, not drop"""
    pos, faces = meshgen.blob(200, 128, seed=9)
    meshes = [(pos, faces), (pos.copy(), faces.copy())]
    sc, _ = make_scene(ctx, meshes)
    sc.detectCollision(ob.DeviceType.GPU0, 4, 3)
    want, ncand = oracle_pairs(port, meshes)
    assert sc.getCandidateCount() == ncand
    assert np.array_equal(sc.canonical_pairs(), want)


def test_every_queue_regrows_from_the_smallest_reservation(ctx, port):
    """work queue, candidate list and pair list all start at the minimum (2^16 records) on a contact that needs several
    times that in each: the detection is repeated with larger buffers until nothing overflows, and loses nothing"""
    pos, faces = meshgen.blob(120, 80, seed=4)
    meshes = [(pos, faces), (pos.copy(), faces.copy())]  # coincident: every triangle touches its neighbours' copies
    want, ncand = oracle_pairs(port, meshes)
    assert ncand > 3 * (1 << 16) and len(want) > (1 << 16)
    sc, _ = make_scene(ctx, meshes)
    sc.reserve(1 << 16, 1 << 16, 1 << 16)
    sc.detectCollision(ob.DeviceType.GPU0, 4, 0)
    assert sc.getCandidateCount() == ncand
    assert np.array_equal(sc.canonical_pairs(), want)
    # the grown buffers are kept: the next detection needs no repeat and gives the same set
    sc.detect_async(4, 0)
    assert sc.counts() == (len(want), ncand)
    assert np.array_equal(sc.canonical_pairs(), want)


def test_shards_partition_the_pair_set(ctx, port):
    pos, faces = meshgen.blob(80, 64, seed=10)
    posB = port.transform_positions(pos, ob.mat_translate(ob.mat_identity(), (0.7, 0.2, 0.1)))
    meshes = [(pos, faces), (posB, faces)]
    want, _ = oracle_pairs(port, meshes)
    for world in (2, 3, 8):
        parts = []
        for rank in range(world):
            sc, _ = make_scene(ctx, meshes)
            sc.set_shard(rank, world)
            sc.detectCollision(ob.DeviceType.GPU0, 4, 3)
            parts.append(sc.canonical_pairs())
        allp = np.concatenate(parts)
        assert len(allp) == len(want), f"world={world}: shards overlap or miss pairs"
        allp = allp[np.lexsort((allp[:, 3], allp[:, 2], allp[:, 1], allp[:, 0]))]
        assert np.array_equal(allp, want)
        assert sum(len(p) > 0 for p in parts) >= 2


def test_graph_replay_of_a_whole_frame(ctx, port):
    pos, faces = meshgen.blob(64, 64, seed=12)
    mA, mB = ob.Mesh(pos, faces), ob.Mesh(pos, faces)
    tA = ob.OibvhTree(mA, ctx=ctx)
    tA.build()
    tB = ob.OibvhTree(tA, mB)
    M = mB.transform_matrix_rotate((0, 0, 1), 2.0)
    T2 = mB.transform_matrix_translate((0.25, 0.0, 0.0))
    sc = ob.Scene(ctx)
    sc.addOibvhTree(tA)
    sc.addOibvhTree(tB)

    def frame():
        tA.build()
        tB.transform(M)
        tB.transform(T2)
        tB.refit(upload=False)
        sc.detect_async(4, 3)

    frame()
    sc.counts()  # warm-up: buffers exist
    before = ctx.launch_count()
    ctx.capture_begin()
    frame()
    graph = ctx.capture_end()
    assert ctx.launch_count() == before  # capture does not execute
    oa = port.build(pos, faces, mA.m_aabb)
    posB = port.transform_positions(port.transform_positions(pos, M), T2)
    for it in range(3):
        graph.launch()
        n, c = sc.counts()
        posB = port.transform_positions(port.transform_positions(posB, M), T2)
        nodesB = port.refit(posB, oa["faces"])
        pp, nc = port.detect([(oa["nodes"], oa["faces"], pos), (nodesB, oa["faces"], posB)])
        assert (n, c) == (len(pp), nc), f"replay {it}"
        assert np.array_equal(sc.canonical_pairs(), oracle.canonical_pairs(pp, [oa["perm"], oa["perm"]]))
    assert ctx.launch_count() > before


@pytest.mark.gpu
def test_pair_vertex_stream_and_wireframes(ctx, port, golden):
    """f3: Scene::convertToVertexArray / OibvhTree::convertToVertexArray as device-side gathers, bit-exact against the
    oracle and (through canonical ordering) against the stream frozen from the unmodified reference"""
    import torch
    import oibvh_b200 as ob
    g = golden["vertex_streams"]
    pos, faces, posB = g["pos"], g["faces"], g["posB"]
    ta = ob.OibvhTree(ob.Mesh(pos, faces), ctx=ctx)
    ta.build()
    tb = ob.OibvhTree(ob.Mesh(posB, faces), ctx=ctx)
    tb.build()
    sc = ob.Scene(ctx)
    sc.addOibvhTree(ta)
    sc.addOibvhTree(tb)
    sc.detectCollision(ob.DeviceType.GPU0, 4, 3)
    pairs = sc.m_intTriPairs
    da, db = ta.download(), tb.download()
    got = sc.convertToVertexArray()
    want = oracle.pair_vertices(pairs, [(da["faces"], pos), (db["faces"], posB)])
    assert_bit_equal(got, want, "pair vertex stream")
    # same multiset of 72-byte records as the reference's stream (its pair order differs)
    ref_rec = np.ascontiguousarray(g["pair_vertices"]).view(np.uint32).reshape(-1, 18)
    got_rec = np.ascontiguousarray(got).view(np.uint32).reshape(-1, 18)
    key = lambda r: r[np.lexsort(r.T[::-1])]
    assert np.array_equal(key(got_rec), key(ref_rec))
    # device-buffer variant: enqueue-only, capacity-clamped
    n = len(pairs)
    buf = torch.full((n + 5, 18), -1.0, dtype=torch.float32, device="cuda")
    stream = torch.cuda.ExternalStream(ctx.stream, device=0)
    with torch.cuda.stream(stream):
        sc.pair_vertices_device(buf.data_ptr(), n + 5)
    ctx.synchronize()
    assert_bit_equal(buf[:n].cpu().numpy().reshape(-1, 3), want, "device stream")
    assert (buf[n:] == -1).all()
    small = torch.full((4, 18), -1.0, dtype=torch.float32, device="cuda")
    sc.pair_vertices_device(small.data_ptr(), 3)
    ctx.synchronize()
    assert_bit_equal(small[:3].cpu().numpy().reshape(-1, 3), want[:18], "clamped stream")
    assert (small[3:] == -1).all()
    # wireframes of the first 256 nodes
    v, i = ta.convertToVertexArray()
    wv, wi = oracle.box_wireframe(da["nodes"], 256, n_prims=len(faces))
    assert_bit_equal(v, wv, "box corners")
    assert np.array_equal(i, wi)
    # tree A has the reference's node boxes only if its leaf order matches; the golden boxes are over UNSORTED faces,
    # so compare through the oracle on the golden nodes instead
    gv, gi = oracle.box_wireframe(g["nodesA"], 256, n_prims=len(faces))
    assert_bit_equal(gv, g["box_vertices"], "golden box corners")
    for t in (ta, tb):
        t.close()


# ---- opt-in extensions: temporal coherence and self-collision ------------------------------------
def test_temporal_coherence_equals_detection_from_the_roots(ctx, port):
    """a coherent scene records a BVTT cut once and starts the following detections from it: over a 100-frame rotation
    (refit only) the pair set of every frame equals that of a detection from the roots, and the oracle's on samples"""
    pos, faces = meshgen.blob(96, 72, seed=21)
    faces = meshgen.shuffle_faces(faces)
    mA, mB = ob.Mesh(pos, faces), ob.Mesh(pos, faces)
    tA = ob.OibvhTree(mA, ctx=ctx)
    tA.build()
    tB = ob.OibvhTree(tA, mB)
    M0 = mB.transform_matrix_translate((0.8, 0.1, 0.05))
    mB.transform(M0)
    tB.transform(M0)
    tB.refit(upload=False)
    plain, coh = ob.Scene(ctx), ob.Scene(ctx)
    for sc in (plain, coh):
        sc.addOibvhTree(tA)
        sc.addOibvhTree(tB)
    coh.set_coherence(True)
    oa = port.build(pos, faces, mA.m_aabb)
    sizes = set()
    for frame in range(100):
        R = mB.transform_matrix_rotate((0.3, 1.0, 0.2), 1.0)
        mB.transform(R)
        tB.transform(R)
        tB.refit(upload=False)
        plain.detect_async(4, 0)
        coh.detect_async(4, 0)
        want = plain.canonical_pairs()
        assert coh.counts() == plain.counts(), f"frame {frame}"
        assert np.array_equal(coh.canonical_pairs(), want), f"frame {frame}"
        sizes.add(len(want))
        if frame % 33 == 0:
            nodesB = port.refit(mB.m_positions, oa["faces"])
            pp, _ = port.detect([(oa["nodes"], oa["faces"], pos), (nodesB, oa["faces"], mB.m_positions)])
            assert np.array_equal(want, oracle.canonical_pairs(pp, [oa["perm"], oa["perm"]])), f"frame {frame}"
    assert len(sizes) > 10  # the contact really changed over the rotation
    # a large jump (the bodies far apart, then deeply interpenetrating) is still exact: the cut is complete
    for shift in ((5.0, 0.0, 0.0), (-5.6, -0.1, 0.0)):
        M = mB.transform_matrix_translate(shift)
        mB.transform(M)
        tB.transform(M)
        tB.refit(upload=False)
        plain.detect_async(4, 0)
        coh.detect_async(4, 0)
        assert np.array_equal(coh.canonical_pairs(), plain.canonical_pairs())
    assert len(plain.canonical_pairs()) > 0
    # a rebuild changes the face order: the next coherent detection records a new cut, and stays exact
    tB.build()
    plain.detect_async(4, 0)
    coh.detect_async(4, 0)
    coh.detect_async(4, 0)
    assert np.array_equal(coh.canonical_pairs(), plain.canonical_pairs())
    # replayed from a graph
    ctx.capture_begin()
    tB.transform(mB.transform_matrix_rotate((0, 0, 1), 0.5))
    tB.refit(upload=False)
    coh.detect_async(4, 0)
    g = ctx.capture_end()
    for _ in range(3):
        g.launch()
    n_graph = coh.counts()
    plain.detect_async(4, 0)
    assert plain.counts() == n_graph and np.array_equal(coh.canonical_pairs(), plain.canonical_pairs())
    g.close()


@pytest.mark.parametrize("cut_depth", [3, 6, 9, 30])
def test_temporal_coherence_cut_depths_and_three_bodies(ctx, port, golden, cut_depth):
    """any cut depth (also deeper than the trees: the cut is then the root pairs) and more than two objects"""
    g = golden["collide"]
    meshes = [(g[f"body{k}_pos"], g[f"body{k}_faces"]) for k in range(3)]
    sc, trees = make_scene(ctx, meshes)
    sc.set_coherence(True, cut_depth)
    for _ in range(3):
        sc.detect_async(4, 3)
        assert np.array_equal(sc.canonical_pairs(), g["bodies_pairs"])
    M = ob.mat_translate(ob.mat_identity(), (0.05, -0.02, 0.03))
    trees[1].transform(M)
    trees[1].refit(upload=False)
    sc.detect_async(4, 3)
    got = sc.canonical_pairs()
    sc2 = ob.Scene(ctx)
    for t in trees:
        sc2.addOibvhTree(t)
    sc2.detectCollision(ob.DeviceType.GPU0, 4, 3)
    assert np.array_equal(got, sc2.canonical_pairs())


def _self_collision_oracle(port, pos, faces):
    """brute force: all pairs a < b of one mesh with overlapping boxes and no common vertex, through the oracle's SAT"""
    boxes = port.leaf_aabbs(pos, faces)
    out = []
    lo, hi = boxes[:, :3], boxes[:, 3:]
    for a in range(len(faces) - 1):
        ov = np.all((lo[a] <= hi[a + 1:]) & (hi[a] >= lo[a + 1:]), axis=1)
        for b in np.nonzero(ov)[0] + a + 1:
            if len(set(faces[a]) & set(faces[b])):
                continue
            if port.tri_tri(pos[faces[a]], pos[faces[b]]):
                out.append((0, 0, a, b))
    return np.array(out, np.uint32).reshape(-1, 4)


def test_self_collision_matches_brute_force(ctx, port):
    """opt-in: a mesh against itself -- a folded sheet whose two halves pass through each other"""
    pos, faces = meshgen.terrain(24, 20, height=0.0, size=(2.0, 2.0))
    p = pos.astype(np.float64)
    fold = p[:, 0] > 0.1
    p[fold, 1] = 0.35 - 0.9 * (p[fold, 0] - 0.1)   # the right part is bent up and back down through the left part
    p[fold, 0] = 0.1 - 0.8 * (p[fold, 0] - 0.1)
    pos = p.astype(np.float32)
    faces = meshgen.shuffle_faces(faces, seed=3)
    want = _self_collision_oracle(port, pos, faces)
    assert len(want) > 10
    t = ob.OibvhTree(ob.Mesh(pos, faces), ctx=ctx)
    t.build()
    sc = ob.Scene(ctx)
    sc.addOibvhTree(t)
    sc.detectCollision(ob.DeviceType.GPU0, 4, 3)
    assert sc.getIntTriPairCount() == 0  # without the option a single object has no pairs (scene.cu:195-196)
    sc.set_self_collision(True)
    for entry, expand in ((4, 3), (0, 0), (2, 2)):
        sc.detectCollision(ob.DeviceType.GPU0, entry, expand)
        raw = sc.m_intTriPairs
        assert (raw[:, 0] == 0).all() and (raw[:, 1] == 0).all() and (raw[:, 2] < raw[:, 3]).all()
        perm = t.download()["perm"]
        got = np.stack([raw[:, 0], raw[:, 1], perm[raw[:, 2]], perm[raw[:, 3]]], 1)
        got[:, 2:] = np.sort(got[:, 2:], axis=1)  # input face ids of a pair, smaller first
        got = got[np.lexsort((got[:, 3], got[:, 2]))]
        assert np.array_equal(got, want[np.lexsort((want[:, 3], want[:, 2]))]), (entry, expand)
    # together with a second object and with temporal coherence
    pos2, faces2 = meshgen.icosphere(3, radius=0.3, center=(0.0, 0.1, 0.0))
    t2 = ob.OibvhTree(ob.Mesh(pos2, faces2), ctx=ctx)
    t2.build()
    sc.addOibvhTree(t2)
    sc.detectCollision(ob.DeviceType.GPU0, 4, 3)
    both = sc.m_intTriPairs
    sc.set_coherence(True)
    for _ in range(2):
        sc.detect_async(4, 3)
        again = sc.m_intTriPairs
        assert np.array_equal(again[np.lexsort(again.T[::-1])], both[np.lexsort(both.T[::-1])])
    assert (both[:, 0] == both[:, 1]).sum() == len(want) and (both[:, 0] != both[:, 1]).sum() > 0
