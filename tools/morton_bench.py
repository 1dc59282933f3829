import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, bench, oibvh_b200 as ob
ctx = ob.Context(0); stream = torch.cuda.ExternalStream(ctx.stream, device=0)
pos, faces = bench.make_meshes()
t = ob.OibvhTree(ob.Mesh(pos, faces), ctx=ctx)
ctx.enable_timing(True)
for name, fn in (("build", t.build),):
    for _ in range(5): fn()
    ctx.stage_ms()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(50): fn()
    e1.record(stream); torch.cuda.synchronize()
    print(f"{name}: {e0.elapsed_time(e1)/50*1e3:.1f} us")
