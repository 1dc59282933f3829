"""A few eager frames of the bench workload for ncu (no timing claims are ever taken from a run under ncu)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import argparse
import bench
import oibvh_b200 as ob

ap = argparse.ArgumentParser()
ap.add_argument("--frames", type=int, default=3)
ap.add_argument("--nu", type=int, default=bench.NU)
ap.add_argument("--nv", type=int, default=bench.NV)
a = ap.parse_args()
pos, faces = bench.make_meshes(a.nu, a.nv)
mA = ob.Mesh(pos, faces); mB = mA.copy()
ctx = ob.Context(0)
tA = ob.OibvhTree(mA, ctx=ctx); tA.build()
tB = ob.OibvhTree(tA, mB)
M0 = mB.transform_matrix_translate(bench.OFFSET_B); mB.transform(M0); tB.transform(M0)
R = mB.transform_matrix_rotate((0, 0, 1), 1.0)
tB.build()
sc = ob.Scene(ctx); sc.addOibvhTree(tA); sc.addOibvhTree(tB)
for i in range(a.frames):
    tA.build(); tB.build(); tB.transform(R); tA.refit(upload=False); tB.refit(upload=False)
    sc.detect_async(bench.ENTRY_LEVEL, bench.EXPAND_LEVELS)
    print(i, sc.counts(), sc.round_stats())
