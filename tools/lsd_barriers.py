"""Per-CTA wall-clock arrival / departure at every grid barrier of the cooperative sort (PROFILE variant)."""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, bench, oibvh_b200 as ob
pos, faces = bench.make_meshes()
two = len(sys.argv) > 1 and sys.argv[1] == "two"
mA = ob.Mesh(pos, faces)
tA = ob.OibvhTree(mA); tA.build()
trees = [tA]
if two:
    tB = ob.OibvhTree(tA, mA.copy()); tB.build(); trees.append(tB)
for _ in range(3):
    ob.build_many(trees) if two else tA.build()
tA.ctx.synchronize()
buf = np.zeros((320, 4, 4, 2), np.uint64)
ob._lib.oibvh_debug_lsd_barriers(buf.ctypes.data_as(ctypes.c_void_p))
b = buf[:296].astype(np.int64)
t0 = b[:, 0, 3, 0].min()
def q(x): return "min %6d p50 %6d p90 %6d max %6d" % (x.min(), np.percentile(x, 50), np.percentile(x, 90), x.max())
print("kernel entry spread (ns):", q(b[:, 0, 3, 0] - t0))
for p in range(3):
    print(f"pass {p} start: {q(b[:, p, 3, 0] - t0)}")
    for k, nm in enumerate("ABC"):
        if p == 2 and k == 2: continue
        arr, lv = b[:, p, k, 0] - t0, b[:, p, k, 1] - t0
        print(f"  bar{nm} arrive: {q(arr)}   leave: {q(lv)}   last-arrive -> first-leave {lv.min() - arr.max()} ns")
        late = np.argsort(arr)[-5:]
        print(f"        latest arrivals: ctas {late.tolist()} at {arr[late].tolist()}")
