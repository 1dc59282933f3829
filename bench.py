#!/usr/bin/env python
"""bench.py -- oibvh collision pipeline on B200: ms/frame (build + refit + broad + narrow) at 1M tris.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's own CPU path (oracle/_ref)
    torchrun --nproc-per-node N bench.py --gpus N ...        # one rank per GPU, NCCL

Workload (BASELINE.json configs[1]): two 2^20-triangle "bunny stand-in" blobs (the reference's bunny.obj is
missing from its checkout), body B one radius away so the surfaces intersect along a curve, rotated 1 degree
per frame about its centre (rigid transform, main.cpp:248-252). One step = one frame =
    build(A) + build(B) + transform(B) + refit(A) + refit(B) + broad phase + narrow phase   (SURVEY.md §8d).
Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "ms/frame (build+refit+broad+narrow) at 1M tris"
UNIT = "ms/frame"
NU, NV = 1024, 512            # blob(1024, 512) -> 2^20 faces, 525 312 vertices
OFFSET_B = (1.55, 0.1, 0.05)  # surfaces intersect along a closed curve
# The reference calls detectCollision(GPU0, 4, 3) (main.cpp:284). Both values are traversal hints -- the pair set
# does not depend on them -- so the bench keeps entry level 4 and lets the kernel pick the
# levels per round from the front size (expand_levels = 0).
ENTRY_LEVEL, EXPAND_LEVELS = 4, 0


def workload_config(args, T, V):
    return {
        "workload": "two 2^20-triangle synthetic blobs (bunny stand-in), rigid 1 deg/frame rotation of body B, "
                    "full pipeline per frame: build x2 + refit x2 + broad + narrow (BASELINE.json configs[1])",
        "tris_per_mesh": int(T), "verts_per_mesh": int(V), "meshes": 2,
        "entry_level": ENTRY_LEVEL, "expand_levels": EXPAND_LEVELS,
        "l2": "no explicit flush: one frame streams ~190 MB of distinct buffers (> 126 MB L2)",
        "parallelism": f"replicated BVH, seed front sharded over {args.gpus} GPU(s), NCCL all-gather of pair lists"
        if args.gpus > 1 else "single GPU",
    }


def make_meshes(nu=NU, nv=NV, shuffle=True):
    from oibvh_b200 import meshgen
    pos, faces = meshgen.blob(nu, nv, seed=1234)
    if shuffle:
        faces = meshgen.shuffle_faces(faces, seed=7)  # asset order is arbitrary: make the sort do real work
    return pos, faces


# --------------------------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi) -- runs during the timed region
# --------------------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region. NVML in a thread every ~2 ms (the timed region
    of the default run is ~14 ms, shorter than one nvidia-smi period); falls back to `nvidia-smi -lms 100`."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []       # nvidia-smi fallback: (time, csv line)
        self.samples = []     # NVML: (time, sm MHz, reasons bitmask)
        self.sm_max = None
        self.nvml = None
        self.stop_flag = False
        self.thread = None

    def _nvml_index(self):
        # CUDA_VISIBLE_DEVICES remaps CUDA ordinals; NVML enumerates physical devices
        vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
        ids = [v for v in vis.split(",") if v.strip() != ""]
        if ids and all(v.strip().isdigit() for v in ids) and self.gpu < len(ids):
            return int(ids[self.gpu])
        return self.gpu

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self._nvml_index())
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)  # fails here rather than in the thread
            self.nvml = (pynvml, h)
            self.thread = threading.Thread(target=self._pump_nvml, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump_nvml(self):
        pynvml, h = self.nvml
        get_reasons = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
            getattr(pynvml, "nvmlDeviceGetCurrentClocksThrottleReasons")
        while not self.stop_flag:
            try:
                clk = float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
                self.samples.append((time.time(), clk, int(get_reasons(h))))
            except Exception:
                pass
            time.sleep(0.002)

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.nvml is not None:
            time.sleep(0.01)
            self.stop_flag = True
            pynvml = self.nvml[0]
            names = {"hw_slowdown": getattr(pynvml, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                     "hw_thermal_slowdown": getattr(pynvml, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                     "sw_thermal_slowdown": getattr(pynvml, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                     "sw_power_cap": getattr(pynvml, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
            inside = [(c, r) for ts, c, r in self.samples if t0 <= ts <= t1] or [(c, r) for _, c, r in self.samples]
            reasons = sorted(n for n, bit in names.items() if any(r & bit for _, r in inside))
            return {"sm_mhz": float(np.median([c for c, _ in inside])) if inside else None, "sm_max_mhz": self.sm_max,
                    "reasons": reasons, "samples": len(inside), "source": "NVML, ~2 ms period, inside the timed region"}
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], [], set()
        for ts, line in self.lines:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                clk, mx = float(f[1]), float(f[2])
            except ValueError:
                continue
            if t0 - 0.05 <= ts <= t1 + 0.15:
                sm.append(clk)
                smax.append(mx)
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        if not sm:  # region shorter than the sampling period: fall back to every sample we have
            for ts, line in self.lines:
                f = [x.strip() for x in line.split(",")]
                try:
                    sm.append(float(f[1]))
                    smax.append(float(f[2]))
                except (ValueError, IndexError):
                    pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": float(max(smax)) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm), "source": "nvidia-smi -lms 100"}


# --------------------------------------------------------------------------------------------------------------
# CPU arms (oracle): reported baselines, never the product path
# --------------------------------------------------------------------------------------------------------------
def cpu_frame_port(port, pos, faces, posB, aabb):
    """one frame with the C restatement (Morton build like the GPU path)"""
    t0 = time.perf_counter()
    a = port.build(pos, faces, aabb)
    b = port.build(posB, faces, aabb)
    na = port.refit(pos, a["faces"])
    nb = port.refit(posB, b["faces"])
    pairs, ncand = port.detect([(na, a["faces"], pos), (nb, b["faces"], posB)])
    return (time.perf_counter() - t0) * 1e3, len(pairs), ncand


def host_info():
    """nproc and CPU model of the box the CPU baseline ran on"""
    model = None
    try:
        for l in open("/proc/cpuinfo"):
            if l.lower().startswith("model name"):
                model = l.split(":", 1)[1].strip()
                break
    except OSError:
        pass
    return {"nproc": os.cpu_count(), "cpu_model": model, "threads_used": 1}


def cpu_baseline_port(pos, faces, posB, aabb, frames=3):
    import oracle
    port = oracle.Port()
    ms = [cpu_frame_port(port, pos, faces, posB, aabb) for _ in range(frames)]
    return {"value": float(np.median([m[0] for m in ms])), "unit": UNIT, "cores": 1, "kind": "port", "host": host_info(),
            "sample": f"full workload (2 x {len(faces)} tris), median of {frames} frames, single thread "
                      "(the reference CPU path is single-threaded)", "pairs": ms[0][1], "candidates": ms[0][2]}


REF_GPU_EXE = os.path.join(ROOT, "oracle", "_ref", "ref_gpu_bench")
REF_GPU_NU, REF_GPU_NV = 256, 192


def ours_frame_ms(ob, ctx, stream, pos, faces, frames=30):
    """this library's frame (same stages as the main workload, graph replay) on arbitrary meshes: used to put a number
    beside the reference GPU path on the bounded sample it can run"""
    import torch
    mesh_a = ob.Mesh(pos, faces)
    mesh_b = mesh_a.copy()
    ta = ob.OibvhTree(mesh_a, ctx=ctx)
    ta.build()
    tb = ob.OibvhTree(ta, mesh_b)
    M0 = mesh_b.transform_matrix_translate(OFFSET_B)
    mesh_b.transform(M0)
    tb.transform(M0)
    M_rot = mesh_b.transform_matrix_rotate((0.0, 0.0, 1.0), 1.0)
    tb.build()
    sc = ob.Scene(ctx)
    sc.addOibvhTree(ta)
    sc.addOibvhTree(tb)

    def frame():
        ob.build_many([ta, tb])
        tb.transform(M_rot)
        ob.refit_many([ta, tb])
        sc.detect_async(ENTRY_LEVEL, EXPAND_LEVELS)
    for _ in range(3):
        frame()
        sc.counts()
    ctx.capture_begin()
    frame()
    g = ctx.capture_end()
    g.launch()
    sc.counts()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(frames):
        g.launch()
    e1.record(stream)
    ctx.synchronize()
    torch.cuda.synchronize()
    n_pairs, _ = sc.counts()
    g.close()
    sc.close()
    for t in (ta, tb):
        t.close()
    return {"value": e0.elapsed_time(e1) / frames, "unit": UNIT, "pairs_last_frame": int(n_pairs),
            "timing": "CUDA events around graph replays, inputs resident (the reference's figure includes its own "
                      "host<->device copies, which are part of its API)"}




def reference_gpu_baseline(pos, faces, frames=12, keep=None):
    """The UNMODIFIED reference GPU path (its five .cu files recompiled for sm_100a by `make -C oracle refgpu`,
    driven by oracle/ref_gpu_main.cu) on the same meshes and the same frame, on this GPU: the recompiled kernels
    this library replaces. A reported baseline like cpu_baseline -- never on the product path. Returns None when the
    binary was not prebuilt (it needs /root/reference at build time)."""
    import subprocess
    import tempfile
    if not os.path.exists(REF_GPU_EXE):
        return None
    with tempfile.TemporaryDirectory() as d:
        mesh, out = os.path.join(d, "mesh.bin"), os.path.join(d, "pairs.bin")
        with open(mesh, "wb") as f:
            f.write(np.array([len(pos), len(faces)], np.uint32).tobytes())
            f.write(np.ascontiguousarray(pos, np.float32).tobytes())
            f.write(np.ascontiguousarray(faces, np.uint32).tobytes())
            f.write(np.array(list(OFFSET_B) + [0.0, 0.0, 1.0, 1.0], np.float32).tobytes())
        try:
            res = subprocess.run([REF_GPU_EXE, mesh, str(frames), out], capture_output=True, text=True, timeout=60)
        except subprocess.TimeoutExpired:
            return {"unavailable": "reference GPU path timed out"}
        if res.returncode != 0:
            return {"unavailable": f"reference GPU path failed (rc {res.returncode}): {res.stderr.strip()[:200]}"}
        rows = np.array([[float(x) for x in l.split()[2:]] for l in res.stdout.splitlines() if l.startswith("frame")])
        if keep is not None:
            keep["pairs"] = np.fromfile(out, np.uint32).reshape(-1, 8)
            keep["posB"] = np.fromfile(out + ".posB", np.float32).reshape(-1, 3)
    rows = rows[min(2, len(rows) - 1):]  # the first frames pay allocations inside thrust
    med = np.median(rows, axis=0)
    return {"value": float(med[3]), "unit": UNIT, "kind": "reference GPU path (unmodified src/cuda/*.cu, nvcc sm_100a)",
            "stage_ms": {"build": float(med[0]), "refit": float(med[1]), "detect": float(med[2])},
            "pairs": int(rows[-1][4]), "frames": int(len(rows)),
            "timing": "host clock around the reference's own synchronous calls (it copies trees, faces and vertices "
                      "host<->device inside every build/refit/detect): build A + build B, rotate B + refit A + refit B, "
                      "detectCollision(GPU0, 4, 3)"}


def run_reference_arm(args):
    """The reference's own CPU implementation: unmodified SimpleBVH::build/refit + SimpleCollide::detect
    (oracle/_ref, compiled from /root/reference by oracle/Makefile). Single-threaded like the reference."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle
    T_full = 2 * NU * NV
    line = {"impl": "reference", "metric": METRIC, "unit": UNIT, "higher_is_better": False, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "gpu_launches": 0}
    if oracle.ref_available():
        # largest size at which the unmodified reference is well-defined: SimpleCollide::detect keeps a depth
        # histogram `int a[19]` (src/cpu/simpleCollide.cpp:63-68), overrun for trees deeper than 18 = 2^18 faces
        nu, nv = 512, 256
        # generator (file-like, spatially coherent) face order: SimpleBVH splits faces in INPUT order without any
        # spatial sort (src/cpu/simpleBVH.cpp:122-159), so on the shuffled order the GPU arm gets, its boxes span
        # the whole mesh and detect() degenerates to O(n^2) (90 s at 65 K faces). The pair SET is order-independent.
        pos, faces = make_meshes(nu, nv, shuffle=False)
        T = len(faces)
        R = oracle.Ref()
        mA, mB = R.mesh_create(pos, faces), R.mesh_create(pos, faces)
        R.mesh_translate(mB, OFFSET_B)
        times = []
        n_pairs = 0
        steps = max(1, min(args.steps, 20))
        for it in range(args.warmup + steps):
            R.mesh_rotate(mB, (0, 0, 1), 1.0)
            t0 = time.perf_counter()
            bA, bB = R.bvh_create(mA), R.bvh_create(mB)
            R.bvh_build(bA)
            R.bvh_build(bB)
            R.bvh_refit(bA)
            R.bvh_refit(bB)
            c = R.collide_create()
            R.collide_add(c, bA)
            R.collide_add(c, bB)
            n_pairs = len(R.collide_detect(c))
            dt = (time.perf_counter() - t0) * 1e3
            R.collide_destroy(c)
            R.bvh_destroy(bA)
            R.bvh_destroy(bB)
            if it >= args.warmup:
                times.append(dt)
        scale = T_full / T
        raw = float(np.mean(times))
        value = raw * scale
        kind, sample = "reference", (f"unmodified reference CPU classes on 2 x {T} tris (largest size where "
                                     f"SimpleCollide's int a[19] depth histogram is in bounds), faces in generator order, {steps} frames; "
                                     f"ms/frame scaled x{scale:.1f} by triangle count to 2 x {T_full}")
        extra = {"raw_ms_per_sample_frame": raw, "pairs_in_sample": n_pairs, "steps_run": steps}
    else:
        pos, faces = make_meshes()
        port = oracle.Port()
        aabb = port.mesh_aabb(pos)
        import oibvh_b200 as ob
        posB = port.transform_positions(pos, ob.mat_translate(ob.mat_identity(), OFFSET_B))
        steps = max(1, min(args.steps, 5))
        ms = [cpu_frame_port(port, pos, faces, posB, aabb)[0] for _ in range(steps)]
        value = float(np.mean(ms))
        kind, sample = "port", f"oracle port (oracle/_ref not built), full 2 x {T_full} tris, {steps} frames"
        extra = {"steps_run": steps}
    line.update({"value": value, "ms_per_step": value,
                 "config": {"workload": "two 2^20-triangle synthetic blobs, build x2 + refit x2 + detect per frame, "
                                        "reference CPU path (SimpleBVH + SimpleCollide)"},
                 "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": kind, "sample": sample,
                                  "host": host_info()},
                 "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
    line.update(extra)
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------------------------
def run_gpu_arm(args):
    import torch
    import oibvh_b200 as ob

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if ob.device_count() < 1:
        raise SystemExit("bench.py needs a CUDA device: oibvh_b200 has no CPU fallback")
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        import datetime
        # a collective that does not complete within three minutes aborts the job instead of hanging it
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank),
                                timeout=datetime.timedelta(seconds=180))
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world} (launch with torchrun for N>1)"
    dev = local_rank if world > 1 else 0
    torch.cuda.set_device(dev)

    pos, faces = make_meshes()
    T, V = len(faces), len(pos)
    mesh_a = ob.Mesh(pos, faces)
    mesh_b = mesh_a.copy()
    ctx = ob.Context(dev)
    stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
    tree_a = ob.OibvhTree(mesh_a, ctx=ctx)
    tree_a.build()
    tree_b = ob.OibvhTree(tree_a, mesh_b)
    M0 = mesh_b.transform_matrix_translate(OFFSET_B)
    mesh_b.transform(M0)
    tree_b.transform(M0)
    M_rot = mesh_b.transform_matrix_rotate((0.0, 0.0, 1.0), 1.0)  # about B's centre, which the rotation fixes
    tree_b.build()
    scene = ob.Scene(ctx)
    scene.addOibvhTree(tree_a)
    scene.addOibvhTree(tree_b)
    scene.set_shard(rank, world)

    def frame():
        ob.build_many([tree_a, tree_b])
        tree_b.transform(M_rot)
        ob.refit_many([tree_a, tree_b])  # two independent refit launches, enqueued on two streams (fork / join)
        scene.detect_async(ENTRY_LEVEL, EXPAND_LEVELS)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up (eager), then capture the frame as one CUDA graph ----
    for _ in range(max(args.warmup, 3)):
        frame()
        scene.counts()
    ctx.capture_begin()
    frame()
    graph = ctx.capture_end()
    for _ in range(2):
        graph.launch()
        scene.counts()

    # multi-GPU: the pair list (and its count) is all-gathered on the same stream right behind the frame graph,
    # fixed-size so that no host round trip sits between frames
    from oibvh_b200 import distributed as obd
    gather = None
    if dist is not None:
        # fixed-size exchange (no host round trip between frames): 4x this rank's warm-up pair count, rounded up to a
        # power of two, at least 4096 records -- the all-gather moves world x cap x 16 bytes per frame
        # (agreed across ranks by all-reduce: per-rank counts differ, and a collective whose size differs per rank hangs)
        n_warm, _ = scene.counts()
        cap = obd.agree_capacity(n_warm, floor=4096, ceiling=scene.pair_capacity(), device=torch.device("cuda", dev))
        pairs_ptr, _ = scene.device_pairs()
        ctr_ptr = scene.device_counters()
        tdev = torch.device("cuda", dev)
        HEAD = obd.HEAD_RECORDS  # the 512-byte counter block is the head of the pair-list allocation
        assert pairs_ptr == ctr_ptr + HEAD * 16, "counter block does not precede the pair list"
        # ONE all-gather per frame: [counter block | first `cap` pair records] of every rank
        block_view = obd.block_view(ctr_ptr, cap, tdev)
        block_all = torch.empty((world * (HEAD + cap), 4), dtype=torch.int32, device=tdev)
        ctr_all = block_all.view(world, HEAD + cap, 4)[:, 0, :]        # row 0 of a block: cand, pairs, overflow, -
        pairs_all = block_all.view(world, HEAD + cap, 4)[:, HEAD:, :]  # rank r's pairs: pairs_all[r, :count_r]

        def gather():
            with torch.cuda.stream(stream):
                obd.gather_blocks(block_all, block_view)
        gather()
        torch.cuda.synchronize()

    # ---- timed region 1: K frames, inputs resident in HBM, one graph launch per frame ----
    sampler = ClockSampler(dev)
    sampler.start()
    time.sleep(0.3)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    l0 = ctx.launch_count()
    w0 = time.time()
    ev0.record(stream)
    for _ in range(args.steps):
        graph.launch()
        if gather is not None:
            gather()
    ev1.record(stream)
    barrier()
    w1 = time.time()
    l1 = ctx.launch_count()
    dev_ms = ev0.elapsed_time(ev1)
    n_pairs, n_cand = scene.counts()
    clocks = sampler.stop(w0, w1)
    ms_per_step = dev_ms / args.steps
    if dist is not None:
        t = torch.tensor([ms_per_step], dtype=torch.float64, device=f"cuda:{dev}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_per_step = float(t.item())

    # ---- timed region 2 (same K frames, eager launches + CUDA events around each stage): roofline inputs ----
    ctx.enable_timing(True)
    stage = {k: 0.0 for k in ob.STAGES}
    for _ in range(args.steps):
        frame()
        ms = ctx.stage_ms()
        scene.counts()
        for k in stage:
            stage[k] += ms[k]
    ctx.enable_timing(False)
    stage = {k: v / args.steps for k, v in stage.items()}
    # the roofline kernel on its own: K back-to-back refit launches of one mesh
    ra0, ra1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tree_a.refit(upload=False)
    ctx.synchronize()
    ra0.record(stream)
    for _ in range(args.steps):
        tree_a.refit(upload=False)
    ra1.record(stream)
    ctx.synchronize()
    torch.cuda.synchronize()
    refit_alone_ms = ra0.elapsed_time(ra1) / args.steps
    rounds = [r for r in scene.round_stats() if r]
    # broad and narrow phase share one persistent kernel: split its time by the SM-cycle stamps of its phases
    cyc = scene.phase_cycles()
    if cyc and sum(cyc) > 0:
        share = cyc[-1] / float(sum(cyc))
        both = stage["broad"] + stage["narrow"]
        stage["narrow"], stage["broad"] = both * share, both * (1.0 - share)

    # ---- timed region 3: end to end through the C ABI with HOST buffers (pinned), H2D + D2H inside ----
    n_e2e = args.steps
    host_a = torch.from_numpy(mesh_a.m_positions.copy()).pin_memory()
    rot_frames = []
    mb = mesh_b.copy()
    for _ in range(4):  # a short cycle of distinct host position buffers for body B
        mb.transform(M_rot)
        rot_frames.append(torch.from_numpy(mb.m_positions.copy()).pin_memory())
    pair_host = torch.empty((max(4 * n_pairs, 1 << 16), 4), dtype=torch.int32).pin_memory()
    if dist is not None:
        gathered_host = torch.empty((world * (HEAD + cap), 4), dtype=torch.int32).pin_memory()

    def e2e_upload(i):
        tree_a.set_positions_from_host_ptr(host_a.data_ptr())
        tree_b.set_positions_from_host_ptr(rot_frames[i % len(rot_frames)].data_ptr())

    def e2e_frame(i, upload=True, prefetch=None):
        # both uploads are enqueued first (they run on the library's copy stream); body A's build + refit overlap
        # body B's upload, which is why the builds are issued per tree here
        if upload:
            e2e_upload(i)
        tree_a.build()
        tree_a.refit(upload=False)
        tree_b.build()
        tree_b.refit(upload=False)
        scene.detect_async(ENTRY_LEVEL, EXPAND_LEVELS)
        if prefetch is not None:
            e2e_upload(prefetch)  # double buffering: the next step's H2D runs under this step's kernels
        if dist is None:
            # one C-ABI call: waits for the frame, reads the counters, copies the pair list into the pinned buffer
            return scene.get_pairs_into(pair_host.data_ptr())
        # N > 1: the fixed-size exchange of the timed region (counter blocks + padded pair lists, two all-gathers on
        # the frame's stream, no host round trip in between), then ONE device->host read of everything
        gather()
        with torch.cuda.stream(stream):
            gathered_host.copy_(block_all, non_blocking=True)
        stream.synchronize()
        counts_now, _, truncated = obd.unpack_blocks(gathered_host, world, cap)
        if truncated:  # a shard outgrew the fixed exchange: exact (slower) variable-size gather
            ptr, n = scene.device_pairs()
            local = obd.pairs_tensor_from_device_ptr(ptr, n, torch.device("cuda", dev))
            with torch.cuda.stream(stream):
                full = obd.gather_pairs(local, n)
                pair_host[:full.shape[0]].copy_(full, non_blocking=True)
            stream.synchronize()
            return int(full.shape[0])
        return int(sum(counts_now))  # rank r's pairs: gathered_host.view(world, HEAD + cap, 4)[r, HEAD:HEAD + counts_now[r]]

    for i in range(2):
        e2e_frame(i)
    barrier()
    t0 = time.perf_counter()
    tot_pairs = 0
    for i in range(n_e2e):
        tot_pairs += e2e_frame(i)
    barrier()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / n_e2e
    if dist is not None:
        t = torch.tensor([e2e_ms], dtype=torch.float64, device=f"cuda:{dev}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())
    h2d = 2 * 12 * V
    d2h = CTR_BYTES + 16 * (tot_pairs // max(n_e2e, 1)) if dist is None else 16 * world * (HEAD + cap)
    # the same loop with the NEXT step's host->device copies enqueued before this step's result is awaited (every
    # step still uploads its own inputs and reads its own result; reported beside e2e, not instead of it)
    e2e_upload(0)
    e2e_frame(0, upload=False, prefetch=1)
    barrier()
    t0 = time.perf_counter()
    for i in range(1, n_e2e + 1):
        e2e_frame(i, upload=False, prefetch=i + 1)
    barrier()
    e2e_pipe_ms = (time.perf_counter() - t0) * 1e3 / n_e2e
    ctx.synchronize()
    if dist is not None:
        t = torch.tensor([e2e_pipe_ms], dtype=torch.float64, device=f"cuda:{dev}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_pipe_ms = float(t.item())

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6.65 TB/s"
        N = 2 * T - 1 + bin((1 << (T - 1).bit_length()) - T).count("1")
        refit_bytes = 12 * T + 12 * V + 24 * N                 # SURVEY.md §8d, per mesh, per launch
        build_bytes = 112 * T + 24 * V + 24 * N
        refit_ms = stage["refit"] / 2.0                        # two refit launches per frame, running concurrently
        build_ms = stage["build"] / 2.0
        traffic = None
        try:  # DRAM bytes per launch of the roofline kernel from the committed ncu --set full capture
            tr = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json")))
            if tr.get("T") == T:
                traffic = int(tr["dram_bytes_read_per_launch"]) + int(tr["dram_bytes_write_per_launch"])
        except Exception:
            pass
        refit_gbs = refit_bytes / (refit_ms * 1e-3) / 1e9 if refit_ms > 0 else 0.0
        build_gbs = build_bytes / (build_ms * 1e-3) / 1e9 if build_ms > 0 else 0.0
        line = {
            "metric": METRIC, "value": ms_per_step, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": False,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, T, V),
            "frame_mtris_per_s": 2 * T / (ms_per_step * 1e-3) / 1e6,
            "build_mtris_per_s": T / (build_ms * 1e-3) / 1e6 if build_ms > 0 else None,
            "stage_ms": stage, "collide_phase_cycles": cyc, "pairs": n_pairs if dist is None else int(ctr_all[:, 1].sum().item()),
            "pairs_this_rank": n_pairs, "candidates": n_cand, "bvtt_rounds": rounds,
            "gather": None if dist is None else {"records_per_rank": int(cap), "collectives_per_frame": 1,
                                                 "bytes_per_frame": int(world * (HEAD + cap) * 16),
                                                 "truncated": bool(int(ctr_all[:, 1].max().item()) > cap)},
            "gpu_launches": int(l1 - l0),
            "wall_ms_per_step": (w1 - w0) * 1e3 / args.steps,
            "clocks": clocks,
            "e2e": {"value": e2e_ms, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)},
            "e2e_double_buffered": {"value": e2e_pipe_ms, "unit": UNIT,
                                    "note": "next step's H2D enqueued under this step's kernels; same bytes per step"},
            "roofline": {"bound": "hbm", "kernel": "tree_emit_kernel<false> (refit: leaf AABBs + whole bottom-up "
                         "reduction, one launch per mesh)", "achieved": refit_gbs, "peak": peak, "unit": "GB/s",
                         "frac": refit_gbs / peak, "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": refit_bytes, "ms_per_launch": refit_ms,
                         "timing": "CUDA events around the frame's refit stage in the eager second pass over the same K "
                                   "frames; the stage is the two refit launches (one per mesh) enqueued on two streams, "
                                   "ms_per_launch = stage time / 2, achieved = bytes of both / stage time",
                         "one_launch_alone": {"ms_per_launch": refit_alone_ms,
                                              "achieved": refit_bytes / (refit_alone_ms * 1e-3) / 1e9,
                                              "frac": refit_bytes / (refit_alone_ms * 1e-3) / 1e9 / peak,
                                              "timing": "K back-to-back launches on one mesh, CUDA events"}},
            "stage_share": {k: (v / sum(stage.values()) if sum(stage.values()) > 0 else 0.0) for k, v in stage.items()},
            "roofline_build": {"bound": "hbm", "stage": "build (per tree) = morton keys + cooperative 4-pass radix sort (both trees in one launch) + emit",
                               "achieved": build_gbs, "peak": peak, "unit": "GB/s", "frac": build_gbs / peak,
                               "algorithmic_bytes_per_build": build_bytes, "ms_per_build": build_ms},
        }
        if world == 1 and not args.no_cpu_baseline:
            # the CPU baseline runs on the positions body B has NOW (after every rotation of the timed regions); the
            # GPU count for exactly these positions is reported beside its pair count
            posB = tree_b.m_positions
            ob.build_many([tree_a, tree_b])
            scene.detect_async(ENTRY_LEVEL, EXPAND_LEVELS)
            gpu_now = scene.counts()
            line["cpu_baseline"] = cpu_baseline_port(pos, faces, posB, mesh_a.m_aabb)
            line["cpu_baseline"]["gpu_pairs_same_positions"] = int(gpu_now[0])
            line["cpu_baseline"]["gpu_candidates_same_positions"] = int(gpu_now[1])
            # The unmodified reference GPU path only completes this scene up to ~10^5 triangles per body: it emits
            # BVTT children untested, so at 2 x 196 608 triangles its front outgrows its fixed 10 M-node buffers and
            # its unchecked level loop never terminates (measured on B200). Bounded sample: 2 x 98 304 triangles, with
            # this library timed on the very same meshes beside it.
            s_pos, s_faces = make_meshes(REF_GPU_NU, REF_GPU_NV)
            ref_gpu = reference_gpu_baseline(s_pos, s_faces)
            if ref_gpu is not None:
                ref_gpu["sample"] = (f"2 x {len(s_faces)} triangles (largest size of this scene the reference GPU path "
                                     "completes; it overflows its fixed 10M-node BVTT buffers at 2 x 196608)")
                if "unavailable" not in ref_gpu:
                    ref_gpu["ours_same_sample"] = ours_frame_ms(ob, ctx, stream, s_pos, s_faces)
                line["reference_gpu"] = ref_gpu
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


CTR_BYTES = 64 * 4


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    # a wedged collective or GPU must not hold the driver for ever: fail the process loudly after ten minutes
    def _watchdog():
        sys.stderr.write("bench.py: no result after 600 s -- aborting (rank %s)\n" % os.environ.get("RANK", "0"))
        sys.stderr.flush()
        os._exit(3)
    wd = threading.Timer(600.0, _watchdog)
    wd.daemon = True
    wd.start()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
