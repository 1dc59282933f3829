"""configs[4]-scale check: 16.8 M-triangle terrain vs a 1 M-triangle body pressed into it (streaming sort path,
deep tree L = 24 vs L = 20, large BVTT front). Compares against the CPU oracle and prints device timings."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import oibvh_b200 as ob, oracle
from oibvh_b200 import meshgen

P = oracle.Port()
ctx = ob.Context(0); stream = torch.cuda.ExternalStream(ctx.stream, device=0)
t0 = time.time()
tpos, tfaces = meshgen.terrain(2897, 2897, height=0.3, size=(8.0, 8.0))
bpos, bfaces = meshgen.blob(1024, 512, seed=5, radius=1.5, center=(0.2, 0.9, -0.3))
print("meshes", len(tfaces), len(bfaces), f"{time.time()-t0:.1f}s")
terrain = ob.OibvhTree(ob.Mesh(tpos, tfaces), ctx=ctx)
body = ob.OibvhTree(ob.Mesh(bpos, bfaces), ctx=ctx)
def timed(fn, n=5):
    fn(); ctx.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(n): fn()
    e1.record(stream); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
T, V, N, _ = terrain.info()
ms = timed(terrain.build); print(f"terrain build {ms*1e3:.0f} us  ({(112*T+24*V+24*N)/ms/1e6/6547.8*100:.1f}% of HBM peak)")
ms = timed(lambda: terrain.refit(upload=False)); print(f"terrain refit {ms*1e3:.0f} us  ({(12*T+12*V+24*N)/ms/1e6/6547.8*100:.1f}% of HBM peak)")
body.build()
sc = ob.Scene(ctx); sc.addOibvhTree(terrain); sc.addOibvhTree(body)
sc.detectCollision(ob.DeviceType.GPU0, 4, 0)
ms = timed(lambda: sc.detect_async(4, 0)); n, c = sc.counts()
print(f"detect {ms*1e3:.0f} us  candidates {c} pairs {n} rounds {sc.round_stats()}")
if "--check" in sys.argv:
    t0 = time.time()
    ot = P.build(tpos, tfaces); obd = P.build(bpos, bfaces)
    print(f"oracle builds {time.time()-t0:.1f}s")
    d = terrain.download()
    print("terrain nodes bit-exact:", np.array_equal(d["nodes"].view(np.uint32), ot["nodes"].view(np.uint32)), "perm:", np.array_equal(d["perm"], ot["perm"]))
    t0 = time.time()
    pp, nc = P.detect([(ot["nodes"], ot["faces"], tpos), (obd["nodes"], obd["faces"], bpos)])
    print(f"oracle detect {time.time()-t0:.1f}s", len(pp), nc)
    print("pair set equal:", np.array_equal(sc.canonical_pairs(), oracle.canonical_pairs(pp, [ot["perm"], obd["perm"]])))
