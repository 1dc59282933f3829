"""configs[3]-scale check: a many-body scene (default 4096 instanced bodies in a box, cubes / icospheres / small
blobs mixed). Per frame: one rigid transform per body, refit of every tree, inter-object broad + narrow phase.
Prints device timings of the batched (*_many) path next to the one-launch-per-object path, and with --check compares
the pair set of a smaller scene against the CPU oracle."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import oibvh_b200 as ob
from oibvh_b200 import meshgen


def make_scene(ctx, n, seed=0):
    rng = np.random.default_rng(seed)
    side = int(np.ceil(n ** (1.0 / 3.0)))
    protos = [meshgen.cube(), meshgen.icosphere(1), meshgen.icosphere(2), meshgen.icosphere(3),
              meshgen.blob(48, 32, seed=3)]
    built = {}
    meshes, trees = [], []
    for i in range(n):
        k = int(rng.integers(len(protos)))
        pos, faces = protos[k]
        cell = np.array([i % side, (i // side) % side, i // (side * side)], np.float32)
        # bodies of diameter ~2 x scale on a unit grid: neighbours touch
        c = (cell + rng.uniform(-0.15, 0.15, 3)).astype(np.float32)
        scale = np.float32(rng.uniform(0.35, 0.6))
        p = (pos * scale + c).astype(np.float32)
        m = ob.Mesh(p, faces)
        meshes.append(m)
        trees.append(ob.OibvhTree(m, ctx=ctx))
    return meshes, trees


def timed(ctx, stream, fn, n=10):
    fn(); ctx.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(n): fn()
    e1.record(stream); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


def main():
    n = int(sys.argv[sys.argv.index("--bodies") + 1]) if "--bodies" in sys.argv else 4096
    ctx = ob.Context(0); stream = torch.cuda.ExternalStream(ctx.stream, device=0)
    t0 = time.time()
    meshes, trees = make_scene(ctx, n)
    tris = sum(t.info()[0] for t in trees)
    print(f"{n} bodies, {tris} triangles, created in {time.time()-t0:.1f}s")
    per_tree = list(trees)
    trees = ob.TreeBatch(trees)
    us = timed(ctx, stream, lambda: ob.build_many(trees), 5)
    print(f"build_many: {us:8.1f} us  ({tris/us:.1f} Mtris/s)")
    if n <= 512:
        us1 = timed(ctx, stream, lambda: [t.build() for t in per_tree], 3)
        print(f"per-tree builds: {us1:8.1f} us")
    sc = ob.Scene(ctx)
    for t in trees: sc.addOibvhTree(t)
    rng = np.random.default_rng(1)
    mats = np.stack([m.transform_matrix_rotate(rng.normal(size=3).astype(np.float32), 0.5) for m in meshes])
    us_x = timed(ctx, stream, lambda: ob.transform_many(trees, mats))
    us_r = timed(ctx, stream, lambda: ob.refit_many(trees))
    sc.detectCollision(ob.DeviceType.GPU0, 4, 0)
    us_d = timed(ctx, stream, lambda: sc.detect_async(4, 0))
    npairs, ncand = sc.counts()
    print(f"transform_many {us_x:.1f} us, refit_many {us_r:.1f} us, detect {us_d:.1f} us "
          f"(candidates {ncand}, pairs {npairs}, rounds {sc.round_stats()})")
    print("  detect phase cycles (seeds, traversal, narrow rest, clean-up):", sc.phase_cycles())
    def frame():
        ob.transform_many(trees, mats); ob.refit_many(trees); sc.detect_async(4, 0)
    print(f"frame (transform + refit + detect): {timed(ctx, stream, frame):.1f} us")
    dmats = torch.from_numpy(mats.reshape(-1, 16).copy()).cuda()
    ctx.synchronize(); torch.cuda.synchronize()
    ctx.capture_begin()
    ob.transform_many(trees, None, device_ptr=dmats.data_ptr()); ob.refit_many(trees); sc.detect_async(4, 0)
    graph = ctx.capture_end()
    print(f"frame as one CUDA graph:            {timed(ctx, stream, graph.launch):.1f} us")
    if n <= 512:
        def frame1():
            for t, M in zip(trees, mats): t.transform(M)
            for t in trees: t.refit(upload=False)
            sc.detect_async(4, 0)
        print(f"frame, one launch per object:      {timed(ctx, stream, frame1, 3):.1f} us")
    if "--check" in sys.argv:
        import oracle
        P = oracle.Port()
        t0 = time.time()
        objs, perms = [], []
        for t, m in zip(trees, meshes):
            pos = t.m_positions
            d = t.download()
            w = P.refit(pos, d["faces"])
            assert np.array_equal(w.view(np.uint32), d["nodes"].view(np.uint32))
            objs.append((w, d["faces"], pos)); perms.append(d["perm"])
        pp, nc = P.detect(objs)
        print(f"oracle {time.time()-t0:.1f}s: pairs {len(pp)} candidates {nc}")
        sc.detectCollision(ob.DeviceType.GPU0, 4, 0)
        print("pair set equal:", np.array_equal(sc.canonical_pairs(), oracle.canonical_pairs(pp, perms)))


main()
