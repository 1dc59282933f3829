import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, bench, oibvh_b200 as ob
pos, faces = bench.make_meshes()
t = ob.OibvhTree(ob.Mesh(pos, faces)); t.build(); t.build(); t.ctx.synchronize()
buf = np.zeros((4, 2, 12), np.uint64)
rc = ob._lib.oibvh_debug_coop_profile(buf.ctypes.data_as(ctypes.c_void_p))
names = ["load+zero", "rank", "digit", "barA", "rowscan", "barB", "base", "scatter+store", "barC"]
for p in range(4):
    for w, nm in enumerate(("cta0", "last")):
        d = np.diff(buf[p, w, :10].astype(np.int64))
        print(f"pass {p} {nm}: " + " ".join(f"{n}={x}" for n, x in zip(names, d)), "total", int(buf[p, w, 9] - buf[p, w, 0]))
print("whole kernel cta0:", int(buf[3, 0, 9] - buf[0, 0, 0]), "cycles")
