"""Where the detection kernel's warps spend their time (PROFILE variant: make -C oibvh_b200/csrc profile)."""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, bench, oibvh_b200 as ob
pos, faces = bench.make_meshes()
mA = ob.Mesh(pos, faces); mB = mA.copy()
tA = ob.OibvhTree(mA); tA.build()
tB = ob.OibvhTree(tA, mB)
M0 = mB.transform_matrix_translate(bench.OFFSET_B); mB.transform(M0); tB.transform(M0); tB.build()
sc = ob.Scene(); sc.addOibvhTree(tA); sc.addOibvhTree(tB)
for _ in range(3):
    sc.detect_async(4, 0); sc.counts()
buf = np.zeros(16, np.uint64)
hops = np.zeros((33, 4), np.uint64)
ob._lib.oibvh_debug_collide_profile(buf.ctypes.data_as(ctypes.c_void_p), 1)
ob._lib.oibvh_debug_collide_hops(hops.ctypes.data_as(ctypes.c_void_p), 1)
sc.detect_async(4, 0); print("counts", sc.counts(), "phase cycles", sc.phase_cycles())
ob._lib.oibvh_debug_collide_profile(buf.ctypes.data_as(ctypes.c_void_p), 0)
W = 148 * 21  # traversal warps of the grid: 24 warps per CTA, 3 of them auxiliary
names = ["window", "empty polls", "setup+tests", "push", "narrow", "total", "full polls", "setup alone"]
for i, n in enumerate(names):
    print(f"  {n:12s} {int(buf[i]) / W:10.0f} cycles per warp")
print("  empty polls %d  batches %d  items %d  candidate flushes %d" % tuple(int(x) for x in buf[9:13]))
items = max(1, int(buf[11]))
print("  phase 2 per item: parameter shuffles %.0f cycles, box loads + tests %.0f cycles; set-up+tests+staging in all %.0f per item, %.0f per batch"
      % (int(buf[14]) / items, int(buf[13]) / items, int(buf[2]) / items, int(buf[2]) / max(1, int(buf[10]))))
ob._lib.oibvh_debug_collide_hops(hops.ctypes.data_as(ctypes.c_void_p), 0)
t0 = int(hops[32, 0])
for l in range(32):
    if hops[l, 1] and t0:  # hop timeline: only in a -DOIBVH_PROFILE_HOPS build (it perturbs the cycle counts above)
        print(f"  level {l:2d}: first taken {int(hops[l,0])-t0:7d} ns  last taken {int(hops[l,1])-t0:7d} ns  last finished {int(hops[l,2])-t0:7d} ns")
