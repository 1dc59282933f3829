"""A few eager frames of the bench workload for ncu (no timing claims are ever taken from a run under ncu).
Set-up launches: 3 (build A) + 1 (transform B) + 3 (build B) = 7; every frame after that is the bench frame's 9 kernels
(2 key kernels, 1 sort, 2 build emits, 1 transform, 2 refit emits, 1 detection)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import argparse
import numpy as np
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
mats = np.stack([ob.mat_identity(), R])
for i in range(a.frames):
    ob.build_many([tA, tB])
    ob.transform_refit_many([tA, tB], mats, [False, True])
    sc.detect_async(bench.ENTRY_LEVEL, bench.EXPAND_LEVELS)
    print(i, sc.counts())
