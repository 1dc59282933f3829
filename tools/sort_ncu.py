"""Tiny driver for ncu captures of the build kernels: two 2^20-triangle trees, build_many x N."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, oibvh_b200 as ob
pos, faces = bench.make_meshes()
mA = ob.Mesh(pos, faces)
tA = ob.OibvhTree(mA); tA.build()
tB = ob.OibvhTree(tA, mA.copy()); tB.build()
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 4):
    ob.build_many([tA, tB])
tA.ctx.synchronize()
print("done")
