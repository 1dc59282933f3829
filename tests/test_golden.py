"""The oracle restatement against the committed golden vectors (tests/golden/*.npz, frozen from the unmodified
reference CPU code by tools/make_golden.py). These run everywhere, including the GPU box where
/root/reference does not exist."""
import numpy as np

import oracle
from conftest import assert_bit_equal


def test_trees(port, golden):
    g = golden["trees"]
    pos, faces_all = g["pos"], g["faces"]
    sizes = sorted(int(k[1:-6]) for k in g.files if k.endswith("_aabbs"))
    assert sizes[:4] == [2, 3, 5, 6] and 2400 in sizes
    for T in sizes:
        want = g[f"T{T}_aabbs"]
        assert want.shape[0] == port.get_size(T)
        assert_bit_equal(port.tree_from_faces(pos, faces_all[:T]), want, f"T={T}")


def test_tritri(port, golden):
    g = golden["tritri"]
    got = np.array([port.tri_tri(p, q) for p, q in zip(g["p"], g["q"])], np.uint8)
    assert np.array_equal(got, g["hit"])
    assert 100 < g["hit"].sum() < len(g["hit"]) - 100


def test_two_spheres(port, golden):
    g = golden["collide"]
    for n in (16, 64):
        pos, faces, posB, want = (g[f"sphere{n}_{k}"] for k in ("pos", "faces", "posB", "pairs"))
        p2, f2 = port.gen_uv_sphere(n)  # the generator itself is libm-dependent: check it reproduces here
        if np.array_equal(p2.view(np.uint32), pos.view(np.uint32)):
            assert np.array_equal(f2, faces)
        bA, bB = port.build(pos, faces), port.build(posB, faces)
        got, ncand = port.detect([(bA["nodes"], bA["faces"], pos), (bB["nodes"], bB["faces"], posB)])
        assert np.array_equal(oracle.canonical_pairs(got, [bA["perm"], bB["perm"]]), want)
    assert len(g["sphere64_pairs"]) == 456


def test_three_bodies(port, golden):
    g = golden["collide"]
    meshes = [(g[f"body{k}_pos"], g[f"body{k}_faces"]) for k in range(3)]
    built = [port.build(p, f) for p, f in meshes]
    got, _ = port.detect([(b["nodes"], b["faces"], p) for b, (p, f) in zip(built, meshes)])
    assert np.array_equal(oracle.canonical_pairs(got, [b["perm"] for b in built]), g["bodies_pairs"])
    assert len(g["bodies_pairs"]) > 0


def test_transforms(port, golden):
    g = golden["transforms"]
    cur = g["pos0"]
    assert_bit_equal(port.mesh_aabb(cur), g["aabb0"], "mesh aabb")
    i = 0
    while f"M{i}" in g.files:
        cur = port.transform_positions(cur, g[f"M{i}"])
        assert_bit_equal(cur, g[f"pos{i + 1}"], f"step {i}")
        i += 1
    assert i == 5


def test_vertex_streams(golden):
    """pair vertex stream (SimpleCollide::convertToVertexArray) and node-box wireframes (makeCube) restatements"""
    g = golden["vertex_streams"]
    trees = [(g["faces"], g["pos"]), (g["faces"], g["posB"])]
    assert_bit_equal(oracle.pair_vertices(g["pairs"], trees), g["pair_vertices"], "pair vertices")
    assert len(g["pair_vertices"]) == 6 * len(g["pairs"]) > 0
    v, i = oracle.box_wireframe(g["nodesA"], 256, n_prims=len(g["faces"]))
    assert_bit_equal(v, g["box_vertices"], "box corners")
    assert np.array_equal(i, g["box_indices"])
    assert len(i) == 24 * 256
