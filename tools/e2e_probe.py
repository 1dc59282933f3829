"""Development probe: where does the end-to-end (host buffers in, pair list out) frame time go?"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench, oibvh_b200 as ob
from oibvh_b200 import distributed as obd

pos, faces = bench.make_meshes()
mA = ob.Mesh(pos, faces); mB = mA.copy()
ctx = ob.Context(0)
stream = torch.cuda.ExternalStream(ctx.stream, device=0)
tA = ob.OibvhTree(mA, ctx=ctx); tA.build()
tB = ob.OibvhTree(tA, mB)
M0 = mB.transform_matrix_translate(bench.OFFSET_B); mB.transform(M0); tB.transform(M0)
R = mB.transform_matrix_rotate((0, 0, 1), 1.0)
tB.build()
sc = ob.Scene(ctx); sc.addOibvhTree(tA); sc.addOibvhTree(tB)
hostA = torch.from_numpy(mA.m_positions.copy()).pin_memory()
frames = []
mb = mB.copy()
for _ in range(4):
    mb.transform(R); frames.append(torch.from_numpy(mb.m_positions.copy()).pin_memory())
pair_host = torch.empty((1 << 16, 4), dtype=torch.int32).pin_memory()
dev = torch.device("cuda", 0)

def upload(i):
    tA.set_positions_from_host_ptr(hostA.data_ptr()); tB.set_positions_from_host_ptr(frames[i % 4].data_ptr())
def compute():
    tA.build(); tA.refit(upload=False); tB.build(); tB.refit(upload=False); sc.detect_async(bench.ENTRY_LEVEL, bench.EXPAND_LEVELS)
def readback_torch():
    n, _ = sc.counts(); ptr, n = sc.device_pairs()
    local = obd.pairs_tensor_from_device_ptr(ptr, n, dev)
    with torch.cuda.stream(stream):
        pair_host[:n].copy_(local[:n], non_blocking=True)
    stream.synchronize(); return n
def readback_abi():
    n, _ = sc.counts()
    ob._check(ob._lib.oibvh_scene_get_pairs(sc._h, ob._vp(pair_host.data_ptr()))); return n

def timeit(name, fn, n=50):
    for i in range(3): fn(i)
    ctx.synchronize(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(n): fn(i)
    ctx.synchronize(); torch.cuda.synchronize()
    print(f"{name:50s} {(time.perf_counter()-t0)*1e6/n:8.1f} us/frame", flush=True)

timeit("full (bench path: torch readback)", lambda i: (upload(i), compute(), readback_torch()))
timeit("full (C-ABI get_pairs readback)", lambda i: (upload(i), compute(), readback_abi()))
timeit("no upload: compute + torch readback", lambda i: (compute(), readback_torch()))
timeit("no upload: compute + abi readback", lambda i: (compute(), readback_abi()))
timeit("no upload: compute + counts only", lambda i: (compute(), sc.counts()))
timeit("no upload: compute, no sync per frame", lambda i: compute())
timeit("upload + refit both + sync", lambda i: (upload(i), tA.refit(upload=False), tB.refit(upload=False), ctx.synchronize()))
d = torch.empty_like(hostA, device=dev)
def h2d(i):
    d.copy_(hostA, non_blocking=True); d.copy_(frames[i % 4], non_blocking=True); torch.cuda.synchronize()
timeit("torch H2D of the same two buffers + sync", h2d)
t0 = time.perf_counter()
for i in range(2000): tA.info()
print("ctypes call overhead (info): %.2f us" % ((time.perf_counter() - t0) * 1e6 / 2000))
ctx.capture_begin(); compute(); g = ctx.capture_end()
timeit("no upload: graph(compute) + counts", lambda i: (g.launch(), sc.counts()))
timeit("upload + graph(compute) + abi readback", lambda i: (upload(i), g.launch(), readback_abi()))
