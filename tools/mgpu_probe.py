"""Where a multi-GPU detection spends its time: two ranks on two devices of ONE process (peer access), the bench scene,
phase cycles of both ranks' kernels and CUDA-event times of the detection on each device."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench, oibvh_b200 as ob
W = int(sys.argv[1]) if len(sys.argv) > 1 else 2
pos, faces = bench.make_meshes()
ranks = []
for r in range(W):
    dev = r % ob.device_count()
    ctx = ob.Context(dev)
    mA = ob.Mesh(pos, faces); mB = mA.copy()
    tA = ob.OibvhTree(mA, ctx=ctx); tA.build()
    tB = ob.OibvhTree(tA, mB)
    M0 = mB.transform_matrix_translate(bench.OFFSET_B); mB.transform(M0); tB.transform(M0); tB.build()
    sc = ob.Scene(ctx); sc.addOibvhTree(tA); sc.addOibvhTree(tB)
    ranks.append((ctx, sc, torch.cuda.ExternalStream(ctx.stream, device=dev), dev))
plain = ranks[0][1]
plain.detect_async(4, 0); print("single GPU:", plain.counts(), "phase cycles", plain.phase_cycles())
def timed_single(n=50):
    ctx, sc, st, dev = ranks[0]
    torch.cuda.set_device(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(n): sc.detect_async(4, 0)
    e1.record(st); ctx.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
print(f"single GPU detect: {timed_single():.1f} us")
for r, (ctx, sc, st, dev) in enumerate(ranks):
    sc.set_shard(r, W)
ranks[0][1].reserve(pair_records=1 << 18)
h = ranks[0][1].mgpu_export()
for ctx, sc, st, dev in ranks[1:]:
    sc.mgpu_attach(h)
def frame():
    for ctx, sc, st, dev in ranks[1:]:
        sc.detect_async(4, 0)
    ranks[0][1].detect_async(4, 0)
for _ in range(5):
    frame()
    for ctx, sc, st, dev in ranks: sc.counts()
ev = []
for ctx, sc, st, dev in ranks:
    torch.cuda.set_device(dev)
    ev.append((torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)))
n = 50
for (ctx, sc, st, dev), (e0, e1) in zip(ranks, ev):
    torch.cuda.set_device(dev); e0.record(st)
for _ in range(n): frame()
for (ctx, sc, st, dev), (e0, e1) in zip(ranks, ev):
    torch.cuda.set_device(dev); e1.record(st)
for ctx, sc, st, dev in ranks: ctx.synchronize()
for r, ((ctx, sc, st, dev), (e0, e1)) in enumerate(zip(ranks, ev)):
    print(f"rank {r} (device {dev}): {e0.elapsed_time(e1)/n*1e3:.1f} us per detection, counts {sc.counts()}, phase cycles {sc.phase_cycles()}")
