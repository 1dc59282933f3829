"""GPU parity, build/refit/transform: the CUDA path through the C ABI against the CPU oracle, bit-exact."""
import os
import numpy as np
import pytest

import oibvh_b200 as ob
from conftest import assert_bit_equal
from oibvh_b200 import meshgen

pytestmark = pytest.mark.gpu

# edge sizes: T=2 (minimum), odd, 2^k, 2^k+-1, one chunk (1024) +-1, several chunks, vl = P-1 (T = 2^k + 1)
SIZES = [2, 3, 4, 5, 6, 7, 12, 13, 31, 33, 100, 255, 256, 257, 1000, 1023, 1024, 1025, 2049, 4095, 4096, 4097, 5000, 20480, 34816]


def check_build(ctx, port, pos, faces, aabb=None):
    mesh = ob.Mesh(pos, faces)
    if aabb is not None:
        mesh.m_aabb = np.asarray(aabb, np.float32)
    tree = ob.OibvhTree(mesh, ctx=ctx)
    tree.build()
    T, V, N, depth = tree.info()
    assert (T, V, N) == (len(faces), len(pos), port.get_size(len(faces)))
    assert depth == int(np.floor(np.log2(N)))  # getDepth() = ilog2(N)
    want = port.build(pos, faces, mesh.m_aabb)
    got = tree.download()
    assert np.array_equal(tree.sorted_keys(), want["keys"]), "sorted Morton keys"
    assert np.array_equal(got["perm"], want["perm"]), "stable sort permutation"
    assert np.array_equal(got["faces"], want["faces"]), "sorted faces"
    assert_bit_equal(got["nodes"], want["nodes"], "node AABBs")
    return tree, mesh, want


@pytest.mark.parametrize("T", SIZES)
def test_build_sizes(ctx, port, T):
    pos, faces = meshgen.blob(160, 110, seed=T)
    faces = meshgen.shuffle_faces(faces, seed=T + 1)[:T]
    check_build(ctx, port, pos, faces)


def test_build_is_stable_on_ties(ctx, port):
    """many faces share a Morton cell (coarse mesh AABB): ties must keep input order (stable sort)"""
    pos, faces = meshgen.blob(64, 64, seed=2)
    aabb = np.array([-1000, -1000, -1000, 1000, 1000, 1000], np.float32)  # all keys collapse to a few cells
    tree, mesh, want = check_build(ctx, port, pos, meshgen.shuffle_faces(faces), aabb)
    keys = tree.sorted_keys()
    assert len(np.unique(keys)) < 64
    perm = tree.download()["perm"]
    for k in np.unique(keys)[:8]:
        seg = perm[keys == k]
        assert (np.diff(seg.astype(np.int64)) > 0).all()


def test_planar_mesh_nan_axis(ctx, port):
    """the reference's own objects/cube.obj is a planar quad: 0/0 on the flat axis must quantise to 0"""
    pos, faces = meshgen.quad()
    check_build(ctx, port, pos, faces)
    # a larger planar grid
    pos, faces = meshgen.terrain(40, 30, height=0.0)
    check_build(ctx, port, pos, meshgen.shuffle_faces(faces))


def test_signed_zero_and_duplicates(ctx, port):
    pos, faces = meshgen.blob(32, 32, seed=8)
    pos = pos.copy()
    pos[::7, 0] = -0.0
    pos[3::7, 0] = 0.0
    pos[::5, 1] = 0.0
    faces = np.concatenate([faces, faces[:100]])  # duplicate faces
    check_build(ctx, port, pos, meshgen.shuffle_faces(faces))


def test_refit_after_deformation(ctx, port):
    pos, faces = meshgen.blob(96, 80, seed=5)
    faces = meshgen.shuffle_faces(faces)[:15001]
    tree, mesh, want = check_build(ctx, port, pos, faces)
    for frame in range(3):
        pos2 = meshgen.cloth_positions(pos, frame)
        mesh.m_positions = pos2
        tree.refit()
        assert_bit_equal(tree.m_aabbTree, port.refit(pos2, want["faces"]), f"refit frame {frame}")
        assert np.array_equal(tree.m_faces, want["faces"])  # refit keeps the face order


def test_clone_translate_refit_like_main_cpp(ctx, port):
    """main.cpp:127-135: tree2 = OibvhTree(tree1, mesh2); mesh2.translate; tree2.refit()"""
    pos, faces = meshgen.blob(64, 50, seed=6)
    t1, m1, want = check_build(ctx, port, pos, meshgen.shuffle_faces(faces))
    m2 = m1.copy()
    t2 = ob.OibvhTree(t1, m2)
    assert t2.m_buildDone and t2.info() == t1.info()
    m2.translate((1.0, 0.0, 0.0))
    t2.refit()
    M = m1.transform_matrix_translate((1.0, 0.0, 0.0))
    pos2 = port.transform_positions(pos, M)
    assert_bit_equal(m2.m_positions, pos2, "host Mesh::transform mirror")
    assert_bit_equal(t2.m_aabbTree, port.refit(pos2, want["faces"]), "refit of the clone")
    assert_bit_equal(t1.m_aabbTree, want["nodes"], "source tree untouched")


def test_device_transform_matches_oracle(ctx, port, golden):
    g = golden["transforms"]
    mesh = ob.Mesh(g["pos0"], g["faces"])
    tree = ob.OibvhTree(mesh, ctx=ctx)
    tree.build()
    i = 0
    while f"M{i}" in g.files:
        tree.transform(g[f"M{i}"])
        assert_bit_equal(tree.m_positions, g[f"pos{i + 1}"], f"device transform step {i}")
        i += 1
    tree.refit(upload=False)
    want = port.build(g["pos0"], g["faces"], mesh.m_aabb)
    assert_bit_equal(tree.m_aabbTree, port.refit(g[f"pos{i}"], want["faces"]), "refit on device-resident positions")


def test_transform_refit_many_equals_separate_calls(ctx, port):
    """oibvh_tree_transform_refit_many (transform of one tree under the refit of another, two streams) == transform then
    refit, tree by tree, bit for bit; also captured into a graph with host matrices"""
    specs = [meshgen.blob(120, 90, seed=31), meshgen.blob(100, 80, seed=32), meshgen.icosphere(5)]
    trees, meshes = [], []
    for pos, faces in specs:
        m = ob.Mesh(pos, meshgen.shuffle_faces(faces))
        t = ob.OibvhTree(m, ctx=ctx)
        t.build()
        trees.append(t)
        meshes.append(m)
    mats = np.stack([m.transform_matrix_rotate((0.2, 1.0, -0.4), 3.0 + i) for i, m in enumerate(meshes)])
    apply = [True, False, True]
    want_pos = [port.transform_positions(m.m_positions, M) if a else m.m_positions for m, M, a in zip(meshes, mats, apply)]
    ob.transform_refit_many(trees, mats, apply)
    for t, p in zip(trees, want_pos):
        d = t.download()
        assert_bit_equal(t.m_positions, p, "positions")
        assert_bit_equal(d["nodes"], port.refit(p, d["faces"]), "nodes")
    ctx.capture_begin()
    ob.transform_refit_many(trees, mats, apply)
    g = ctx.capture_end()
    g.launch()
    ctx.synchronize()
    g.close()
    for t, p, M, a in zip(trees, want_pos, mats, apply):
        p2 = port.transform_positions(p, M) if a else p  # capturing records, the one replay executes
        assert_bit_equal(t.m_positions, p2, "positions after the replay")
        assert_bit_equal(t.download()["nodes"], port.refit(p2, t.download()["faces"]), "nodes after the replay")


def test_golden_trees_on_gpu(ctx, golden):
    """node arrays frozen from the reference's SimpleBVH: refit over the SAME face order must reproduce them"""
    g = golden["trees"]
    pos, faces_all = g["pos"], g["faces"]
    for T in (2, 3, 5, 13, 257, 1000, 2400):
        faces = np.ascontiguousarray(faces_all[:T])
        tree = ob.OibvhTree(ob.Mesh(pos, faces), ctx=ctx)
        tree.build()
        d = tree.download()
        # golden is over input order; compare leaf boxes through the permutation and the root directly
        want = g[f"T{T}_aabbs"]
        N = want.shape[0]
        assert_bit_equal(d["nodes"][N - T:], want[N - T:][d["perm"]], f"T={T} leaves")
        assert_bit_equal(d["nodes"][0], want[0], f"T={T} root")


def test_errors(ctx):
    pos, faces = meshgen.cube()
    with pytest.raises(ob.OibvhError):
        ob.OibvhTree(ob.Mesh(pos, faces[:1]), ctx=ctx)  # T = 1
    bad = faces.copy()
    bad[0, 0] = 99
    with pytest.raises(ob.OibvhError):
        ob.OibvhTree(ob.Mesh(pos, bad), ctx=ctx)  # index out of range
    t = ob.OibvhTree(ob.Mesh(pos, faces), ctx=ctx)
    with pytest.raises(ob.OibvhError):
        t.refit()  # refit before build
    sc = ob.Scene(ctx)
    with pytest.raises(ob.OibvhError):
        sc.addOibvhTree(t)  # Scene::addOibvhTree asserts m_buildDone (scene.cu:97)
    with pytest.raises(ob.OibvhError):
        sc.detectCollision(ob.DeviceType.CPU)


def test_build_many_equals_separate_builds(ctx, port):
    """oibvh_tree_build_many sorts the keys of several trees in one cooperative launch: same results"""
    specs = [(meshgen.blob(300, 200, seed=21), None), (meshgen.blob(160, 110, seed=22), 30001),
             (meshgen.icosphere(4), None)]
    trees, wants = [], []
    for (pos, faces), cut in specs:
        faces = meshgen.shuffle_faces(faces)
        if cut:
            faces = np.ascontiguousarray(faces[:cut])
        mesh = ob.Mesh(pos, faces)
        trees.append(ob.OibvhTree(mesh, ctx=ctx))
        wants.append(port.build(pos, faces, mesh.m_aabb))
    ob.build_many(trees)
    for t, w in zip(trees, wants):
        d = t.download()
        assert np.array_equal(t.sorted_keys(), w["keys"])
        assert np.array_equal(d["perm"], w["perm"])
        assert_bit_equal(d["nodes"], w["nodes"], "build_many nodes")
    ob.build_many(trees[:1])  # n = 1 falls back to a plain build
    assert_bit_equal(trees[0].download()["nodes"], wants[0]["nodes"], "build_many(1)")


def test_many_small_trees_in_one_launch(ctx, port):
    """many-body path: trees of <= 4096 triangles build / refit one CTA each, all in one launch; mixed with large
    trees in the same call. Results must equal the oracle tree by tree."""
    rng = np.random.default_rng(11)
    specs = []
    for i in range(40):
        kind = i % 5
        if kind == 0:
            pos, faces = meshgen.cube()
        elif kind == 1:
            pos, faces = meshgen.icosphere(2)
        elif kind == 2:
            pos, faces = meshgen.icosphere(3)
        elif kind == 3:
            pos, faces = meshgen.blob(64, 40, seed=i)
            faces = faces[: int(rng.integers(2, 4097))]
        else:
            pos, faces = meshgen.blob(64, 40, seed=i)
            faces = faces[:4096]
        specs.append((pos, meshgen.shuffle_faces(np.ascontiguousarray(faces), seed=i)))
    specs.append((meshgen.blob(120, 90, seed=77)[0], meshgen.shuffle_faces(meshgen.blob(120, 90, seed=77)[1])))  # large
    specs.append((meshgen.icosphere(4)[0], meshgen.shuffle_faces(meshgen.icosphere(4)[1])))  # large
    meshes = [ob.Mesh(p, f) for p, f in specs]
    trees = [ob.OibvhTree(m, ctx=ctx) for m in meshes]
    wants = [port.build(p, f, m.m_aabb) for (p, f), m in zip(specs, meshes)]
    before = ctx.launch_count()
    ob.build_many(trees)
    launches = ctx.launch_count() - before
    assert launches <= 1 + 3 + 3, f"40 small trees must share one launch (got {launches} launches)"
    for t, w in zip(trees, wants):
        d = t.download()
        assert np.array_equal(t.sorted_keys(), w["keys"])
        assert np.array_equal(d["perm"], w["perm"])
        assert np.array_equal(d["faces"], w["faces"])
        assert_bit_equal(d["nodes"], w["nodes"], "build_many (small) nodes")
    # one rigid transform per tree in one launch, then the refit of all of them
    mats = np.stack([m.transform_matrix_rotate((0.3, 1.0, 0.2), 10.0 + i) if i % 2 else
                     m.transform_matrix_translate((0.1 * i, -0.2, 0.05 * i)) for i, m in enumerate(meshes)])
    before = ctx.launch_count()
    ob.transform_many(trees, mats)
    ob.refit_many(trees)
    assert ctx.launch_count() - before == 1 + 1 + 2
    for i, (t, w) in enumerate(zip(trees, wants)):
        pos2 = port.transform_positions(specs[i][0], mats[i])
        assert_bit_equal(t.m_positions, pos2, f"transform_many tree {i}")
        assert_bit_equal(t.download()["nodes"], port.refit(pos2, w["faces"]), f"refit_many tree {i}")
    # same list again: cached device tables, same results; then a different list
    ob.refit_many(trees)
    ob.refit_many(trees[5:20])
    ob.build_many(trees[:7])
    for t, m, (p, f), M in list(zip(trees, meshes, specs, mats))[:7]:
        w2 = port.build(port.transform_positions(p, M), f, m.m_aabb)  # keys from the moved vertices, original mesh box
        assert np.array_equal(t.download()["perm"], w2["perm"])
        assert_bit_equal(t.download()["nodes"], w2["nodes"], "rebuild after the transform")


@pytest.mark.parametrize("T", [5000, 65536, 200000, 1048576])
def test_single_wave_sort_builds_the_same_tree(ctx, port, T):
    """the cooperative 3 x 10-bit sort (sort_lsd.cu): one tree over the whole grid, two trees side by side in one
    launch (oibvh_tree_build_many), eagerly and replayed from a graph (the kernel re-arms its control block)"""
    nu = max(16, int((T / 2) ** 0.5) + 2)
    pos, faces = meshgen.blob(nu, nu, seed=7)
    faces = meshgen.shuffle_faces(faces)[:T]
    want = port.build(pos, faces, port.mesh_aabb(pos))
    t = ob.OibvhTree(ob.Mesh(pos, faces), ctx=ctx)
    for _ in range(2):  # twice: the second launch starts from the control block the first one left behind
        t.build()
        got = t.download()
        assert np.array_equal(t.sorted_keys(), want["keys"])
        assert np.array_equal(got["perm"], want["perm"])
        assert_bit_equal(got["nodes"], want["nodes"], "nodes")
    faces2 = np.ascontiguousarray(faces[::-1][: max(2, (2 * T) // 3)])
    t2 = ob.OibvhTree(ob.Mesh(pos, faces2), ctx=ctx)
    want2 = port.build(pos, faces2, port.mesh_aabb(pos))
    ob.build_many([t, t2])
    ctx.capture_begin()
    ob.build_many([t, t2])
    g = ctx.capture_end()
    for _ in range(3):
        g.launch()
    ctx.synchronize()
    a, b = t.download(), t2.download()
    g.close()
    assert np.array_equal(a["perm"], want["perm"]) and np.array_equal(b["perm"], want2["perm"])
    assert_bit_equal(a["nodes"], want["nodes"], "build_many nodes A")
    assert_bit_equal(b["nodes"], want2["nodes"], "build_many nodes B")
    t.close()
    t2.close()


def test_sort_heavy_ties_pole_mesh_and_all_equal_keys(ctx, port):
    """key distributions that break bucket-based sorts: the poles of a UV sphere put thousands of triangles into one
    Morton cell, and a mesh box far larger than the mesh collapses EVERY key to one value (the order is then the
    input order: stability is all that is left)"""
    pos, faces = port.gen_uv_sphere(384)  # 294 912 triangles, 768 zero-area triangles at the poles
    faces = meshgen.shuffle_faces(faces, seed=5)
    check_build(ctx, port, pos, faces)
    coarse = np.array([-32, -32, -32, 32, 32, 32], np.float32)  # a few thousand occupied cells, thousands of ties each
    tree, _, _ = check_build(ctx, port, pos, faces, coarse)
    assert np.bincount(np.unique(tree.sorted_keys(), return_inverse=True)[1]).max() > 500
    aabb = np.array([1e6, 1e6, 1e6, 3e6, 3e6, 3e6], np.float32)  # every centroid quantises to cell 0
    tree, mesh, want = check_build(ctx, port, pos, faces, aabb)
    assert len(np.unique(tree.sorted_keys())) == 1
    assert np.array_equal(tree.download()["perm"], np.arange(len(faces), dtype=np.uint32))
