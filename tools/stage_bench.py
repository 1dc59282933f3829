"""Development micro-benchmark: per-stage CUDA-event times of the bench frame (eager) + whole-frame graph time."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import argparse
import numpy as np
import torch
import bench
import oibvh_b200 as ob

ap = argparse.ArgumentParser()
ap.add_argument("--frames", type=int, default=30)
ap.add_argument("--nu", type=int, default=bench.NU)
ap.add_argument("--nv", type=int, default=bench.NV)
ap.add_argument("--entry", type=int, default=bench.ENTRY_LEVEL)
ap.add_argument("--expand", type=int, default=bench.EXPAND_LEVELS)
ap.add_argument("--check", action="store_true")
ap.add_argument("--grid-order", action="store_true", help="keep the generator's face order instead of shuffling it")
a = ap.parse_args()
pos, faces = bench.make_meshes(a.nu, a.nv, shuffle=not a.grid_order)
mA = ob.Mesh(pos, faces); mB = mA.copy()
ctx = ob.Context(0)
stream = torch.cuda.ExternalStream(ctx.stream, device=0)
tA = ob.OibvhTree(mA, ctx=ctx); tA.build()
tB = ob.OibvhTree(tA, mB)
M0 = mB.transform_matrix_translate(bench.OFFSET_B); mB.transform(M0); tB.transform(M0)
R = mB.transform_matrix_rotate((0, 0, 1), 1.0)
tB.build()
sc = ob.Scene(ctx); sc.addOibvhTree(tA); sc.addOibvhTree(tB)
def frame():
    ob.build_many([tA, tB]); ob.transform_refit_many([tA, tB], np.stack([ob.mat_identity(), R]), [False, True])
    sc.detect_async(a.entry, a.expand)
for _ in range(3):
    frame(); sc.counts()
ctx.enable_timing(True)
acc = {k: [] for k in ob.STAGES}
for _ in range(a.frames):
    frame(); ms = ctx.stage_ms(); sc.counts()
    for k in acc: acc[k].append(ms[k])
ctx.enable_timing(False)
T = len(faces)
print(f"T={T} x2  per-frame stage ms (median): " + "  ".join(f"{k}={np.median(v)*1e3:.1f}us" for k, v in acc.items()))
print("  build/tree=%.1fus refit/tree=%.1fus" % (np.median(acc['build'])*500, np.median(acc['refit'])*500), "counts", sc.counts(), "rounds", sc.round_stats())
print("  phase cycles:", sc.phase_cycles())
ctx.capture_begin(); frame(); g = ctx.capture_end()
for _ in range(3): g.launch()
ctx.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(stream)
for _ in range(a.frames): g.launch()
e1.record(stream); torch.cuda.synchronize()
print(f"  graph frame: {e0.elapsed_time(e1)/a.frames*1e3:.1f} us")
if a.check:
    import oracle
    P = oracle.Port()
    oa = P.build(pos, faces, mA.m_aabb)
    d = tA.download()
    print("  build parity:", np.array_equal(d['nodes'].view(np.uint32), oa['nodes'].view(np.uint32)), np.array_equal(d['perm'], oa['perm']))
    posB = tB.m_positions
    dB = tB.download()
    nodesB = P.refit(posB, dB['faces'])
    print("  refit parity:", np.array_equal(dB['nodes'].view(np.uint32), nodesB.view(np.uint32)))
    nodesB = P.refit(posB, oa['faces'])
    pp, nc = P.detect([(oa['nodes'], oa['faces'], pos), (nodesB, oa['faces'], posB)])
    print("  pairs parity:", np.array_equal(sc.canonical_pairs(), oracle.canonical_pairs(pp, [oa['perm'], oa['perm']])), len(pp), nc)
