"""Pins the CPU restatement (oracle/oibvh_oracle.c) to the UNMODIFIED reference CPU classes (oracle/_ref),
run here in the build container. Skipped where oracle/_ref is absent."""
import numpy as np
import pytest

import oracle
from conftest import assert_bit_equal
from oibvh_b200 import meshgen


@pytest.mark.parametrize("T", [2, 3, 5, 6, 7, 12, 13, 100, 255, 256, 257, 1000, 4097, 20000])
def test_tree_equals_simplebvh_bfs(port, ref, T):
    pos, faces = meshgen.blob(128, 100, seed=4)
    faces = meshgen.shuffle_faces(faces, seed=T)[:T]
    m = ref.mesh_create(pos, faces)
    b = ref.bvh_create(m)
    ref.bvh_build(b)
    aabbs, tri = ref.bvh_dump_bfs(b)
    nodes = port.tree_from_faces(pos, faces)
    assert aabbs.shape[0] == port.get_size(T)
    assert_bit_equal(nodes, aabbs, "node AABBs")
    # leaves are the last T entries, leaf k <-> face k
    assert np.array_equal(tri[-T:], np.arange(T))
    assert (tri[:-T] == -1).all()
    # refit after a deformation
    pos2 = meshgen.cloth_positions(pos, 3)
    ref.mesh_set_positions(m, pos2)
    ref.bvh_refit(b)
    aabbs2, _ = ref.bvh_dump_bfs(b)
    assert_bit_equal(port.refit(pos2, faces), aabbs2, "refit AABBs")
    assert_bit_equal(port.mesh_aabb(pos), ref.mesh_aabb(m), "mesh aabb")
    ref.bvh_destroy(b)
    ref.mesh_destroy(m)


def test_tri_tri_equals_reference(port, ref):
    rng = np.random.default_rng(5)
    n = 3000
    p = rng.normal(size=(n, 9)).astype(np.float32)
    q = (rng.normal(size=(n, 9)) * 0.8).astype(np.float32)
    q[:300, :3] = p[:300, :3]
    q[300:500] = p[300:500]
    p[500:600, 6:] = p[500:600, :3]
    big = rng.normal(size=(200, 9)).astype(np.float32) * np.float32(1e19)  # overflow -> inf/NaN paths
    p[600:800] = big
    got = [port.tri_tri(p[i], q[i]) for i in range(n)]
    want = [ref.tri_tri(p[i], q[i]) for i in range(n)]
    assert got == want
    assert 0.05 < np.mean(want) < 0.95


def test_aabb_overlap_inclusive(port, ref):
    a = np.array([0, 0, 0, 1, 1, 1], np.float32)
    for b, want in [([1, 1, 1, 2, 2, 2], True), ([1.0000001, 0, 0, 2, 1, 1], False), ([-1, -1, -1, 0, 0, 0], True),
                    ([0.2, 0.2, 0.2, 0.3, 0.3, 0.3], True), ([0, 0, 1.5, 1, 1, 2], False)]:
        b = np.array(b, np.float32)
        assert port.aabb_overlap(a, b) == ref.aabb_overlap(a, b) == want


@pytest.mark.parametrize("n,shift", [(16, (1.0, 0.1, 0.05)), (48, (0.7, -0.2, 0.3)), (64, (1.0, 0.1, 0.05))])
def test_detect_equals_simplecollide(port, ref, n, shift):
    pos, faces = port.gen_uv_sphere(n)
    M = ref.glm_translate(shift)
    posB = port.transform_positions(pos, M)
    want = oracle.canonical_pairs(ref.detect_meshes([(pos, faces), (posB, faces)]))
    # same (input-order) trees
    nodesA, nodesB = port.tree_from_faces(pos, faces), port.tree_from_faces(posB, faces)
    got, ncand = port.detect([(nodesA, faces, pos), (nodesB, faces, posB)])
    assert np.array_equal(oracle.canonical_pairs(got), want)
    # Morton-ordered trees give the same SET once face ids are mapped back
    bA, bB = port.build(pos, faces), port.build(posB, faces)
    got2, ncand2 = port.detect([(bA["nodes"], bA["faces"], pos), (bB["nodes"], bB["faces"], posB)])
    assert ncand2 == ncand
    assert np.array_equal(oracle.canonical_pairs(got2, [bA["perm"], bB["perm"]]), want)
    if n == 64 and shift == (1.0, 0.1, 0.05):  # SURVEY.md Appendix A known answers
        assert len(want) == 456 and ncand == 2093
        assert want[:3, 2:].tolist() == [[1024, 1601], [1025, 1600], [1025, 1601]]


def test_three_bodies_mixed_sizes(port, ref):
    meshes = []
    for k, (gen, t) in enumerate([(meshgen.blob(20, 17, seed=1), (0, 0, 0)), (meshgen.icosphere(2), (0.7, 0.1, 0)),
                                  (meshgen.blob(33, 9, seed=3), (-0.6, 0.3, 0.2))]):
        pos, faces = gen
        faces = meshgen.shuffle_faces(faces, seed=k)[: len(faces) - 5 * k - 1]
        pos = port.transform_positions(pos, ref.glm_translate(t))
        meshes.append((pos, faces))
    want = oracle.canonical_pairs(ref.detect_meshes(meshes))
    built = [port.build(p, f) for p, f in meshes]
    got, _ = port.detect([(b["nodes"], b["faces"], p) for b, (p, f) in zip(built, meshes)])
    assert np.array_equal(oracle.canonical_pairs(got, [b["perm"] for b in built]), want)
    assert len(want) > 0 and len(set(map(tuple, want[:, :2].tolist()))) >= 2


def test_transform_equals_mesh_transform(port, ref):
    pos, faces = meshgen.blob(16, 12, seed=9)
    m = ref.mesh_create(pos, faces)
    cur = pos
    for axis, ang in [((0, 0, 1), 1.0), ((1, 0, 0), 1.0), ((0.3, -0.5, 0.8), 37.5)]:
        M = ref.glm_rotate_about(ref.mesh_center(m), axis, ang)
        ref.mesh_rotate(m, axis, ang)
        cur = port.transform_positions(cur, M)
        assert_bit_equal(cur, ref.mesh_positions(m, len(pos)), "rotated positions")
    ref.mesh_destroy(m)


def test_vertex_streams_equal_reference(port, ref):
    """oracle.pair_vertices / box_wireframe == unmodified SimpleCollide::convertToVertexArray / makeCube"""
    pos, faces = meshgen.blob(24, 16, seed=5)
    posB = (pos + np.float32([0.9, 0.1, 0.05])).astype(np.float32)
    ms = [ref.mesh_create(pos, faces), ref.mesh_create(posB, faces)]
    bs = [ref.bvh_create(m) for m in ms]
    col = ref.collide_create()
    for b in bs:
        ref.bvh_build(b)
        ref.collide_add(col, b)
    pairs = ref.collide_detect(col)
    assert len(pairs) > 50
    assert_bit_equal(oracle.pair_vertices(pairs, [(faces, pos), (faces, posB)]), ref.collide_vertex_array(col), "verts")
    aabbs, _ = ref.bvh_dump_bfs(bs[0])
    v1, i1 = ref.box_wireframe(aabbs[:256])
    v2, i2 = oracle.box_wireframe(aabbs, 256, n_prims=len(faces))
    assert_bit_equal(v2, v1, "box corners")
    assert np.array_equal(i1, i2)
    ref.collide_destroy(col)
    for b in bs:
        ref.bvh_destroy(b)
    for m in ms:
        ref.mesh_destroy(m)
