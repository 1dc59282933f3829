"""BASELINE.json's full sizes: bit-exact against the oracle where it finishes in seconds (1M), and through
size-independent properties beyond that (4M): sortedness, permutation, parent = union of children, leaf boxes,
refit idempotence, pair-set invariance under traversal parameters."""
import numpy as np
import pytest

import oibvh_b200 as ob
import oracle
from conftest import assert_bit_equal
from oibvh_b200 import meshgen

pytestmark = pytest.mark.gpu


def level_tables(T):
    L = int(np.ceil(np.log2(T)))
    cnt = [(T + (1 << (L - l)) - 1) >> (L - l) for l in range(L + 1)]
    off = np.concatenate([[0], np.cumsum(cnt)]).astype(np.int64)
    return L, cnt, off


def check_tree_properties(nodes, faces, perm, keys, pos, faces_in):
    T = len(faces)
    L, cnt, off = level_tables(T)
    assert nodes.shape[0] == off[-1]
    # sortedness + permutation + stability
    assert (np.diff(keys.astype(np.int64)) >= 0).all()
    assert np.array_equal(np.sort(perm), np.arange(T, dtype=np.uint32))
    same = keys[1:] == keys[:-1]
    assert (perm[1:][same] > perm[:-1][same]).all(), "ties must keep input order"
    assert np.array_equal(faces, faces_in[perm])
    # leaves
    tri = pos[faces]  # [T,3,3]
    leaves = nodes[off[L]:off[L] + T]
    assert_bit_equal(leaves[:, :3], tri.min(axis=1), "leaf min")
    assert_bit_equal(leaves[:, 3:], tri.max(axis=1), "leaf max")
    # every parent is the union of its kept children
    for l in range(L - 1, -1, -1):
        par = nodes[off[l]:off[l] + cnt[l]]
        ch = nodes[off[l + 1]:off[l + 1] + cnt[l + 1]]
        left = ch[0::2][:cnt[l]]
        right = np.concatenate([ch[1::2], left[len(ch[1::2]):]])[:cnt[l]]  # missing right child -> left itself
        want = np.concatenate([np.minimum(left[:, :3], right[:, :3]), np.maximum(left[:, 3:], right[:, 3:])], 1)
        assert np.array_equal(par, want), f"level {l}"


def test_one_million_bit_exact(ctx, port):
    """configs[1]: 2^20 triangles per mesh; the oracle restatement finishes in about a second per build"""
    pos, faces = meshgen.blob(1024, 512, seed=1234)
    faces = meshgen.shuffle_faces(faces)
    mA = ob.Mesh(pos, faces)
    tA = ob.OibvhTree(mA, ctx=ctx)
    tA.build()
    oa = port.build(pos, faces, mA.m_aabb)
    d = tA.download()
    assert np.array_equal(tA.sorted_keys(), oa["keys"])
    assert np.array_equal(d["perm"], oa["perm"])
    assert_bit_equal(d["nodes"], oa["nodes"], "1M build")
    mB = mA.copy()
    tB = ob.OibvhTree(tA, mB)
    M = mB.transform_matrix_translate((1.55, 0.1, 0.05))
    mB.transform(M)
    tB.transform(M)
    R = mB.transform_matrix_rotate((0, 0, 1), 1.0)
    mB.transform(R)
    tB.transform(R)
    tB.refit(upload=False)
    posB = port.transform_positions(port.transform_positions(pos, M), R)
    assert_bit_equal(tB.m_positions, posB, "device transform")
    nodesB = port.refit(posB, oa["faces"])
    assert_bit_equal(tB.m_aabbTree, nodesB, "1M refit")
    sc = ob.Scene(ctx)
    sc.addOibvhTree(tA)
    sc.addOibvhTree(tB)
    pp, nc = port.detect([(oa["nodes"], oa["faces"], pos), (nodesB, oa["faces"], posB)])
    want = oracle.canonical_pairs(pp, [oa["perm"], oa["perm"]])
    for entry, expand in [(4, 3), (0, 1)]:
        sc.detectCollision(ob.DeviceType.GPU0, entry, expand)
        assert sc.getCandidateCount() == nc
        assert np.array_equal(sc.canonical_pairs(), want)
    assert len(want) > 1000


def test_non_power_of_two_million(ctx):
    pos, faces = meshgen.blob(1000, 500, seed=5)  # T = 1 000 000
    faces = meshgen.shuffle_faces(faces)
    t = ob.OibvhTree(ob.Mesh(pos, faces), ctx=ctx)
    t.build()
    d = t.download()
    check_tree_properties(d["nodes"], d["faces"], d["perm"], t.sorted_keys(), pos, faces)
    before = d["nodes"].copy()
    t.refit(upload=False)  # idempotence: same positions -> same tree
    assert np.array_equal(t.m_aabbTree.view(np.uint32), before.view(np.uint32))


def test_four_million_properties(ctx):
    """configs[2] scale: 4 194 304 + 7 triangles (L = 23, virtual leaves on the right)"""
    pos, faces = meshgen.blob(2048, 1025, seed=6)
    faces = np.ascontiguousarray(meshgen.shuffle_faces(faces)[:4 * 2 ** 20 + 7])
    t = ob.OibvhTree(ob.Mesh(pos, faces), ctx=ctx)
    t.build()
    d = t.download()
    check_tree_properties(d["nodes"], d["faces"], d["perm"], t.sorted_keys(), pos, faces)
    pos2 = meshgen.cloth_positions(pos, 2)
    t.set_positions(pos2)
    t.refit(upload=False)
    n2 = t.m_aabbTree
    check_tree_properties(n2, d["faces"], d["perm"], t.sorted_keys(), pos2, faces)


def test_sixteen_million_terrain_vs_million_body(ctx, port):
    """configs[4] at full size: 16.8 M-triangle terrain (L = 25 levels, streaming onesweep sort) against a
    2^20-triangle body pressed into it; the CPU restatement needs ~2 s for this"""
    tpos, tfaces = meshgen.terrain(2897, 2897, height=0.3, size=(8.0, 8.0))
    bpos, bfaces = meshgen.blob(1024, 512, seed=5, radius=1.5, center=(0.2, 0.9, -0.3))
    terrain = ob.OibvhTree(ob.Mesh(tpos, tfaces), ctx=ctx)
    body = ob.OibvhTree(ob.Mesh(bpos, bfaces), ctx=ctx)
    terrain.build()
    body.build()
    ot, obody = port.build(tpos, tfaces), port.build(bpos, bfaces)
    d = terrain.download()
    assert np.array_equal(d["perm"], ot["perm"])
    assert_bit_equal(d["nodes"], ot["nodes"], "16.8M build")
    sc = ob.Scene(ctx)
    sc.addOibvhTree(terrain)
    sc.addOibvhTree(body)
    sc.detectCollision(ob.DeviceType.GPU0, 4, 0)
    pp, nc = port.detect([(ot["nodes"], ot["faces"], tpos), (obody["nodes"], obody["faces"], bpos)])
    assert sc.getCandidateCount() == nc
    assert np.array_equal(sc.canonical_pairs(), oracle.canonical_pairs(pp, [ot["perm"], obody["perm"]]))
    assert len(pp) > 5000
    terrain.close()
