"""Tiny driver for ncu captures of the detection kernel on the bench scene."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, oibvh_b200 as ob
pos, faces = bench.make_meshes()
mA = ob.Mesh(pos, faces); mB = mA.copy()
tA = ob.OibvhTree(mA); tA.build()
tB = ob.OibvhTree(tA, mB)
M0 = mB.transform_matrix_translate(bench.OFFSET_B); mB.transform(M0); tB.transform(M0); tB.build()
sc = ob.Scene(); sc.addOibvhTree(tA); sc.addOibvhTree(tB)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 4):
    sc.detect_async(4, 0); print(sc.counts())
