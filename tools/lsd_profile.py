"""Phase stamps of the cooperative LSD sort (needs the -DOIBVH_PROFILE variant: make -C oibvh_b200/csrc profile)."""
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
buf = np.zeros((4, 2, 12), np.uint64)
rc = ob._lib.oibvh_debug_lsd_profile(buf.ctypes.data_as(ctypes.c_void_p))
names = ["load+zero", "rank", "digit+reorder", "barA", "rowscan", "barB", "bases", "store", "barC"]
for p in range(4):
    if buf[p, 0, 0] == 0: continue
    for w, nm in enumerate(("cta0", "last")):
        d = np.diff(buf[p, w, :10].astype(np.int64))
        print(f"pass {p} {nm}: " + " ".join(f"{n}={x}" for n, x in zip(names, d)), "total", int(buf[p, w, 9] - buf[p, w, 0]))
last = max(p for p in range(4) if buf[p, 0, 0])
print("whole kernel cta0:", int(buf[last, 0, 9] - buf[0, 0, 0]), "cycles")
