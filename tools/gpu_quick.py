"""quick GPU sanity run used during development: build/refit/detect vs the oracle on a few sizes"""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import oibvh_b200 as ob
from oibvh_b200 import meshgen
import oracle

P = oracle.Port()
print("devices", ob.device_count())
ctx = ob.default_context()

def check_tree(pos, faces, tag):
    mesh = ob.Mesh(pos, faces)
    t = ob.OibvhTree(mesh); t.build()
    d = t.download()
    o = P.build(pos, faces, mesh.m_aabb)
    ok_k = np.array_equal(t.sorted_keys(), o['keys'])
    ok_p = np.array_equal(d['perm'], o['perm'])
    ok_f = np.array_equal(d['faces'], o['faces'])
    ok_n = np.array_equal(d['nodes'].view(np.uint32), o['nodes'].view(np.uint32))
    print(f"{tag}: T={len(faces)} keys={ok_k} perm={ok_p} faces={ok_f} nodes={ok_n}")
    if not ok_n:
        bad = np.where((d['nodes'].view(np.uint32) != o['nodes'].view(np.uint32)).any(1))[0]
        print("   first bad nodes", bad[:10], len(bad))
    return t, mesh, o

for T in [2, 3, 5, 13, 100, 257, 1000, 1024, 1025, 4097, 8192]:
    pos, faces = meshgen.uv_sphere(64)
    faces = meshgen.shuffle_faces(faces)[:T]
    check_tree(pos, faces, "sphere")

pos, faces = meshgen.blob(136, 128)
check_tree(pos, meshgen.shuffle_faces(faces), "blob35k")
pos, faces = meshgen.blob(512, 256)
t0 = time.time(); check_tree(pos, meshgen.shuffle_faces(faces), "blob262k"); print(time.time() - t0)

# collision KAT: two UV spheres n=64
pos, faces = P.gen_uv_sphere(64)
mA = ob.Mesh(pos, faces); mB = mA.copy()
tA = ob.OibvhTree(mA); tA.build()
tB = ob.OibvhTree(tA, mB)
mB.translate((1.0, 0.1, 0.05)); tB.refit()
sc = ob.Scene(); sc.addOibvhTree(tA); sc.addOibvhTree(tB)
for (e, k) in [(4, 3), (0, 1), (2, 2), (0, 0)]:
    sc.detectCollision(ob.DeviceType.GPU0, e, k)
    print("detect", e, k, sc.getIntTriPairCount(), sc.getCandidateCount(), sc.round_stats())
cp = sc.canonical_pairs()
oa = P.build(pos, faces, mA.m_aabb)
nodesB = P.refit(mB.m_positions, oa['faces'])
pp, nc = P.detect([(oa['nodes'], oa['faces'], pos), (nodesB, oa['faces'], mB.m_positions)])
ref = oracle.canonical_pairs(pp, [oa['perm'], oa['perm']])
print("oracle", len(pp), nc, "pair set equal:", np.array_equal(cp, ref))
print("launches", ctx.launch_count())
