"""Development probe (-DOIBVH_PROFILE build): wall-clock life of every CTA of one refit launch."""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, bench, oibvh_b200 as ob
pos, faces = bench.make_meshes()
t = ob.OibvhTree(ob.Mesh(pos, faces)); t.build()
for _ in range(5): t.refit(upload=False)
t.ctx.synchronize()
buf = np.zeros((4096, 4), np.uint64)
ob._lib.oibvh_debug_emit_profile(buf.ctypes.data_as(ctypes.c_void_p))
fin = buf[0].astype(np.int64)   # block 0 = finisher CTA (stamps 0 and 3)
b = buf[1:1025].astype(np.int64)
t0 = b[:, 0].min()
start, w1, w0a, w0b = (b[:, k] - t0 for k in range(4))
print("kernel span (first CTA start -> last chunk done / finisher done): %.1f / %.1f us" % ((b[:, 1].max() - t0) / 1e3, (fin[3] - t0) / 1e3))
order = np.argsort(start)
print("CTA starts (us) percentiles 0/25/50/58/60/75/100:", np.percentile(start, [0, 25, 50, 58, 60, 75, 100]) / 1e3)
life = w1 - start
print("CTA life to chunk done (us): first-wave median %.2f, second-wave median %.2f, min %.2f max %.2f" % (
    np.median(life[start < 2000]) / 1e3, np.median(life[start >= 2000]) / 1e3, life.min() / 1e3, life.max() / 1e3))
print("n first-wave (start < 2us):", int((start < 2000).sum()))
h, e = np.histogram(start / 1e3, bins=12)
print("start histogram:", list(zip(np.round(e[:-1], 1), h)))
h, e = np.histogram(w1 / 1e3, bins=12)
print("chunk-done histogram:", list(zip(np.round(e[:-1], 1), h)))
ph = np.zeros((4096, 8), np.uint64)
ob._lib.oibvh_debug_emit_phases(ph.ctypes.data_as(ctypes.c_void_p))
ph = ph[1:1025, :6].astype(np.int64)
names = ["faces arrive", "vertices arrive", "stage+issue bulk", "heights 3-7", "bulk read done"]
d = np.diff(ph, axis=1) / 1e3
w1m = start < 2000
for nm, col in zip(names, d.T):
    print(f"  {nm:18s} wave1 median {np.median(col[w1m]):6.2f} us (p90 {np.percentile(col[w1m], 90):5.2f})   wave2 median {np.median(col[~w1m]):6.2f} us (p90 {np.percentile(col[~w1m], 90):5.2f})")
print("  warp total          wave1 median %.2f  wave2 median %.2f" % (np.median((ph[:, 5] - ph[:, 0])[w1m]) / 1e3, np.median((ph[:, 5] - ph[:, 0])[~w1m]) / 1e3))
print("  CTA start -> warp entry: median %.2f us" % (np.median(ph[:, 0] - b[:, 0]) / 1e3))
