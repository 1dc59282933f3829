import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, bench, oibvh_b200 as ob
pos, faces = bench.make_meshes()
t = ob.OibvhTree(ob.Mesh(pos, faces)); t.build(); t.build(); t.ctx.synchronize()
buf = np.zeros((4, 2, 8), np.uint64)
rc = ob._lib.oibvh_debug_sort_profile(buf.ctypes.data_as(ctypes.c_void_p))
names = ["load", "rank", "digit", "lookback", "scatter-smem", "store"]
for p in range(4):
    for w, nm in enumerate(("tile0", "last")):
        d = np.diff(buf[p, w, :7].astype(np.int64))
        print(f"pass {p} {nm}: " + " ".join(f"{n}={x}" for n, x in zip(names, d)), "total", int(buf[p, w, 6] - buf[p, w, 0]))
