"""Generate tests/golden/*.npz from the UNMODIFIED reference CPU code (oracle/_ref/liboibvh_ref.so).

Run in the build container (needs /root/reference): `python tools/make_golden.py`.
The fixtures freeze, for seeded inputs:
  * SimpleBVH node AABBs in BFS order (== oibvh array order) for several primitive counts,
  * SimpleCollide::detect pair sets (bvhA, bvhB, faceA, faceB) for two- and three-body scenes,
  * triangleIntersect verdicts for random / touching / degenerate triangle pairs,
  * glm translate / rotate matrices and Mesh::transform results, Mesh::m_aabb and m_center,
  * SimpleCollide::convertToVertexArray vertex streams and makeCube node-box wireframes.
The reference's GPU path cannot run here (no GPU in the container) and ships no fixtures of its own
(SURVEY.md §4), so these are the known answers every parity test is anchored to.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from oibvh_b200 import meshgen  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def tri_cases(rng, n):
    """random nearby triangles + exactly touching + degenerate ones"""
    p = rng.normal(size=(n, 3, 3)).astype(np.float32)
    q = (rng.normal(size=(n, 3, 3)) * 0.7 + rng.normal(size=(n, 1, 3)) * 0.8).astype(np.float32)
    k = n // 8
    q[:k, 0] = p[:k, 0]                      # shared vertex
    q[k:2 * k, :2] = p[k:2 * k, :2]          # shared edge
    q[2 * k:3 * k] = p[2 * k:3 * k]          # identical
    q[3 * k:4 * k, 2] = q[3 * k:4 * k, 1]    # degenerate (zero area) second triangle
    p[4 * k:5 * k, 2] = p[4 * k:5 * k, 0]    # degenerate first triangle
    q[5 * k:6 * k] = p[5 * k:6 * k] + np.float32(1e-3)  # nearly coincident
    # coplanar pairs
    p[6 * k:7 * k, :, 2] = 0
    q[6 * k:7 * k, :, 2] = 0
    return p.reshape(n, 9), q.reshape(n, 9)


def vertex_streams(R):
    """SimpleCollide::convertToVertexArray output (in the reference's own pair order) and node-box wireframes
    (unmodified makeCube) for a small two-body scene -> tests/golden/vertex_streams.npz"""
    P = oracle.Port()
    spos, sfaces = P.gen_uv_sphere(16)
    mA, mB = R.mesh_create(spos, sfaces), R.mesh_create(spos, sfaces)
    R.mesh_translate(mB, (1.0, 0.1, 0.05))
    posB = R.mesh_positions(mB, len(spos))
    bA, bB = R.bvh_create(mA), R.bvh_create(mB)
    R.bvh_build(bA)
    R.bvh_build(bB)
    c = R.collide_create()
    R.collide_add(c, bA)
    R.collide_add(c, bB)
    pairs = R.collide_detect(c)
    verts = R.collide_vertex_array(c)
    aabbs, _ = R.bvh_dump_bfs(bA)
    box_v, box_i = R.box_wireframe(aabbs[:256])
    np.savez_compressed(os.path.join(OUT, "vertex_streams.npz"), pos=spos, faces=sfaces, posB=posB, pairs=pairs,
                        pair_vertices=verts, nodesA=aabbs, box_vertices=box_v, box_indices=box_i)
    R.collide_destroy(c)
    for h in (bA, bB):
        R.bvh_destroy(h)
    for h in (mA, mB):
        R.mesh_destroy(h)


def main():
    assert oracle.ref_available(), "build oracle/_ref first: make -C oracle ref"
    R = oracle.Ref()
    os.makedirs(OUT, exist_ok=True)
    if "--vertex-streams-only" in sys.argv:
        vertex_streams(R)
        return
    rng = np.random.default_rng(20240917)

    # ---- trees ----
    trees = {}
    pos, faces_all = meshgen.blob(40, 30, seed=11)
    faces_all = meshgen.shuffle_faces(faces_all, seed=5)
    for T in (2, 3, 5, 6, 7, 12, 13, 100, 255, 256, 257, 1000, 2400):
        faces = np.ascontiguousarray(faces_all[:T])
        m = R.mesh_create(pos, faces)
        b = R.bvh_create(m)
        R.bvh_build(b)
        aabbs, tri = R.bvh_dump_bfs(b)
        trees[f"T{T}_aabbs"] = aabbs
        trees[f"T{T}_tri"] = tri
        R.bvh_destroy(b)
        R.mesh_destroy(m)
    np.savez_compressed(os.path.join(OUT, "trees.npz"), pos=pos, faces=faces_all, **trees)

    # ---- collisions ----
    col = {}
    # (a) the survey's known-answer scene at a small size: two UV spheres, B translated by (1, 0.1, 0.05)
    P = oracle.Port()
    for n in (16, 64):
        spos, sfaces = P.gen_uv_sphere(n)
        mA, mB = R.mesh_create(spos, sfaces), R.mesh_create(spos, sfaces)
        R.mesh_translate(mB, (1.0, 0.1, 0.05))
        posB = R.mesh_positions(mB, len(spos))
        bA, bB = R.bvh_create(mA), R.bvh_create(mB)
        R.bvh_build(bA)
        R.bvh_build(bB)
        c = R.collide_create()
        R.collide_add(c, bA)
        R.collide_add(c, bB)
        pairs = oracle.canonical_pairs(R.collide_detect(c))
        col[f"sphere{n}_pos"] = spos
        col[f"sphere{n}_faces"] = sfaces
        col[f"sphere{n}_posB"] = posB
        col[f"sphere{n}_pairs"] = pairs
        R.collide_destroy(c)
        for h in (bA, bB):
            R.bvh_destroy(h)
        for h in (mA, mB):
            R.mesh_destroy(h)
    # (b) three bodies with different, non-power-of-two primitive counts, rotated + translated
    bodies = []
    specs = [(meshgen.blob(24, 21, seed=1), None), (meshgen.icosphere(3), ((0.8, 0.2, 0.1), (0, 0, 1), 17.0)),
             (meshgen.blob(30, 13, seed=2), ((-0.5, 0.6, 0.3), (1, 0, 0), -33.0))]
    c = R.collide_create()
    keep = []
    for k, ((bpos, bfaces), xf) in enumerate(specs):
        bfaces = meshgen.shuffle_faces(bfaces, seed=k)[: len(bfaces) - (k + 1) * 3]
        m = R.mesh_create(bpos, bfaces)
        if xf:
            R.mesh_rotate(m, xf[1], xf[2])
            R.mesh_translate(m, xf[0])
        p_now = R.mesh_positions(m, len(bpos))
        b = R.bvh_create(m)
        R.bvh_build(b)
        R.collide_add(c, b)
        keep += [m, b]
        col[f"body{k}_pos0"] = bpos
        col[f"body{k}_pos"] = p_now
        col[f"body{k}_faces"] = bfaces
        bodies.append(k)
    col["bodies_pairs"] = oracle.canonical_pairs(R.collide_detect(c))
    np.savez_compressed(os.path.join(OUT, "collide.npz"), **col)

    # ---- triangle-triangle ----
    p, q = tri_cases(rng, 4096)
    hit = np.array([R.tri_tri(p[i], q[i]) for i in range(len(p))], np.uint8)
    np.savez_compressed(os.path.join(OUT, "tritri.npz"), p=p, q=q, hit=hit)

    # ---- transforms / mesh bookkeeping ----
    xf = {}
    bpos, bfaces = meshgen.blob(16, 12, seed=9)
    m = R.mesh_create(bpos, bfaces)
    xf["pos0"] = bpos
    xf["faces"] = bfaces
    xf["aabb0"] = R.mesh_aabb(m)
    xf["center0"] = R.mesh_center(m)
    steps = [("rot", (0, 0, 1), 1.0), ("rot", (1, 0, 0), 1.0), ("tr", (1.0, 0.0, 0.0)), ("rot", (0.3, -0.5, 0.8), 37.5),
             ("tr", (-0.25, 0.125, 3.0))]
    for i, st in enumerate(steps):
        if st[0] == "rot":
            xf[f"M{i}"] = R.glm_rotate_about(R.mesh_center(m), st[1], st[2])
            R.mesh_rotate(m, st[1], st[2])
        else:
            xf[f"M{i}"] = R.glm_translate(st[1])
            R.mesh_translate(m, st[1])
        xf[f"pos{i + 1}"] = R.mesh_positions(m, len(bpos))
        xf[f"center{i + 1}"] = R.mesh_center(m)
    xf["steps"] = np.array([repr(s) for s in steps])
    np.savez_compressed(os.path.join(OUT, "transforms.npz"), **xf)
    vertex_streams(R)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
