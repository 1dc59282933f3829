import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, bench, oibvh_b200 as ob
from oibvh_b200 import meshgen
ctx = ob.Context(0)
stream = torch.cuda.ExternalStream(ctx.stream, device=0)
for (nu, nv) in [(740, 512), (1024, 512), (1024, 1024), (2048, 1024)]:
    pos, faces = meshgen.blob(nu, nv, seed=1); faces = meshgen.shuffle_faces(faces)
    t = ob.OibvhTree(ob.Mesh(pos, faces), ctx=ctx); t.build(); ctx.synchronize()
    for name, fn in (("refit", lambda: t.refit(upload=False)), ("build", t.build)):
        for _ in range(5): fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 50
        e0.record(stream)
        for _ in range(n): fn()
        e1.record(stream); torch.cuda.synchronize()
        T = len(faces); V = len(pos); N = t.info()[2]
        us = e0.elapsed_time(e1) / n * 1e3
        b = (12 * T + 12 * V + 24 * N) if name == "refit" else (112 * T + 24 * V + 24 * N)
        print(f"T={T:8d} {name}: {us:7.1f} us  {b/us/1e3:7.0f} GB/s  ({b/us/1e3/6547.8*100:4.1f}% of measured HBM peak)")
    t.close()
