#!/usr/bin/env python
"""bench.py -- oibvh collision pipeline on B200: ms/frame (build + refit + broad + narrow) at 1M tris.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's own CPU path (oracle/_ref)
    torchrun --nproc-per-node N bench.py --gpus N ...        # one rank per GPU

Headline workload (BASELINE.json configs[1]): two 2^20-triangle "bunny stand-in" blobs (the reference's bunny.obj is
missing from its checkout), body B one radius away so the surfaces intersect along a curve, rotated 1 degree per frame
about its centre (rigid transform, main.cpp:248-252). One step = one frame =
    build(A) + build(B) + transform(B) + refit(A) + refit(B) + broad phase + narrow phase   (SURVEY.md §8d).
The same line carries one sub-record per BASELINE.json config (`configs`): the ~70 K bunny-sized pair, the 4 M deforming
mesh, the 4096-body scene and the 16.8 M terrain, each with stage times, roofline fractions, the CPU path beside it and a
parity flag. At N > 1 every rank keeps a replica of every tree, the BVTT front is dealt to the ranks and every rank's
narrow phase appends its hits to rank 0's pair list through a peer mapping (no collective per frame).
Prints ONE JSON line (rank 0).
"""
import argparse
import hashlib
import importlib.util
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
NU0, NV0 = 136, 128           # configs[0]: 34 816 faces per mesh (~70 K in total, the bunny pair of main.cpp:127-151)
OFFSET_B = (1.55, 0.1, 0.05)  # surfaces intersect along a closed curve
# The reference calls detectCollision(GPU0, 4, 3) (main.cpp:284). Both values are traversal hints -- the pair set
# does not depend on them -- so the bench keeps entry level 4 and lets the kernel pick the
# levels per round from the front size (expand_levels = 0).
ENTRY_LEVEL, EXPAND_LEVELS = 4, 0
HBM_FALLBACK_GBS = 6650.0


def load_meshgen():
    """the mesh generators WITHOUT importing the product package (whose __init__ loads the CUDA library): the
    reference arm must not map oibvh_b200/liboibvh_b200.so"""
    spec = importlib.util.spec_from_file_location("oibvh_meshgen", os.path.join(ROOT, "oibvh_b200", "meshgen.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def make_meshes(nu=NU, nv=NV, shuffle=True):
    mg = load_meshgen()
    pos, faces = mg.blob(nu, nv, seed=1234)
    if shuffle:
        faces = mg.shuffle_faces(faces, seed=7)  # asset order is arbitrary: make the sort do real work
    return pos, faces


def tree_nodes(T):
    return 2 * T - 1 + bin((1 << (T - 1).bit_length()) - T).count("1")


# algorithmic bytes per mesh and launch (the numerators of every roofline fraction)
def bytes_keys(T, V):
    return 12 * T + 12 * V + 4 * T


def bytes_sort(T):
    return 4 * T + 4 * (8 * T + 8 * T)  # histogram read + four 8-bit passes of (key, id) in and out -- fixed numerator


def bytes_emit(T, V):
    return 4 * T + 12 * T + 12 * V + 12 * T + 24 * tree_nodes(T)


def bytes_build(T, V):
    return 112 * T + 24 * V + 24 * tree_nodes(T)  # = keys + sort + emit


def bytes_refit(T, V):
    return 12 * T + 12 * V + 24 * tree_nodes(T)


def bytes_transform(V):
    return 24 * V


def bytes_detect(rounds, n_cand, n_pairs):
    """broad: sum over rounds (56 B per tested node pair + 8 B per emitted child) + 8 B per candidate;
    narrow: 104 B per candidate + 16 B per hit"""
    fronts = list(rounds) + [0]
    broad = sum(56 * fronts[i] + 8 * fronts[i + 1] for i in range(len(rounds))) + 8 * n_cand
    return broad + 104 * n_cand + 16 * n_pairs


def measured_peak():
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        if "hbm_gbs" in peaks:
            return float(peaks["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        pass
    return HBM_FALLBACK_GBS, "fallback 6.65 TB/s (B200_PROFILING.md)"


def committed_traffic():
    """DRAM bytes per launch from the committed `ncu --set full` capture (profiles/r02_traffic.json), by kernel"""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json")))
    except Exception:
        return {}


def pair_hash(canon):
    return hashlib.sha1(np.ascontiguousarray(canon, dtype=np.uint32).tobytes()).hexdigest()[:16]


# --------------------------------------------------------------------------------------------------------------
# clocks sampler -- runs during the timed region
# --------------------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region. NVML in a thread every ~2 ms (the timed region
    of the default run is ~50 ms, shorter than one nvidia-smi period); falls back to `nvidia-smi -lms 100`."""
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


def cpu_baseline_port(pos, faces, posB, aabb, frames=3):
    import oracle
    port = oracle.Port()
    ms = [cpu_frame_port(port, pos, faces, posB, aabb) for _ in range(frames)]
    return {"value": float(np.median([m[0] for m in ms])), "unit": UNIT, "cores": 1, "kind": "port", "host": host_info(),
            "sample": f"full workload (2 x {len(faces)} tris), median of {frames} frames, single thread "
                      "(the reference CPU path is single-threaded)", "pairs": ms[0][1], "candidates": ms[0][2]}


def reference_cpu_frames(R, pos, faces, steps, warmup):
    """the reference's CPU classes on the bench scene: per frame rotate B (untimed: the reference rotates through a
    CUDA kernel, src/utils/mesh.cpp:187-213), then SimpleBVH x2 build + refit, SimpleCollide::detect (timed)"""
    mA, mB = R.mesh_create(pos, faces), R.mesh_create(pos, faces)
    R.mesh_translate(mB, OFFSET_B)
    times, n_pairs = [], 0
    for it in range(warmup + steps):
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
        if it >= warmup:
            times.append(dt)
    return times, n_pairs


def run_reference_arm(args):
    """The reference's own CPU implementation of the path, on the box's host cores: SimpleBVH::build/refit +
    SimpleCollide::detect compiled from /root/reference by oracle/Makefile. Single-threaded like the reference (it has
    no threading: SURVEY.md quick facts). `value` is MEASURED at the metric's full size (2 x 2^20 faces) with
    oracle/_ref/liboibvh_ref_deep.so -- the reference classes with one token changed, `int a[19]` -> `int a[64]`, the
    diagnostic depth histogram that trees deeper than 18 levels overrun (src/cpu/simpleCollide.cpp:63-68) -- on the
    SAME meshes and poses as the GPU arm, faces in generator order: SimpleBVH splits faces in INPUT order without any
    spatial sort (src/cpu/simpleBVH.cpp:122-159), so on the shuffled order the GPU arm is given its boxes span the whole
    mesh and detect() degenerates to O(n^2) (90 s at 65 K faces); the pair SET is order-independent. The unmodified
    library is timed beside it on configs[0] (~35 K x 2) and on its largest valid size (2 x 262 144)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle
    T_full = 2 * NU * NV
    line = {"impl": "reference", "metric": METRIC, "unit": UNIT, "higher_is_better": False, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "gpu_launches": 0}
    extra = {}
    steps = max(1, min(args.steps, 12))   # ~0.5 s per frame: the run stays within a few minutes
    warm = min(args.warmup, 2)
    if oracle.ref_deep_available():
        pos, faces = make_meshes(shuffle=False)
        times, n_pairs = reference_cpu_frames(oracle.Ref(oracle.REF_DEEP_PATH), pos, faces, steps, warm)
        value = float(np.mean(times))
        kind = "reference"
        sample = (f"reference CPU classes (SimpleBVH + SimpleCollide; depth histogram int a[19] -> int a[64], "
                  f"oracle/Makefile refdeep) on the full 2 x {T_full} tris, same meshes and poses as the GPU arm, faces in "
                  f"generator order, {steps} frames measured (not extrapolated), 1 thread")
        extra.update({"same_config": True, "pairs_last_frame": n_pairs, "steps_run": steps})
    else:
        port = oracle.Port()
        pos, faces = make_meshes()
        aabb = port.mesh_aabb(pos)
        M = np.eye(4, dtype=np.float32)
        M[3, :3] = OFFSET_B  # column-major translation
        posB = port.transform_positions(pos, M.reshape(16))
        steps = max(1, min(args.steps, 5))
        ms = [cpu_frame_port(port, pos, faces, posB, aabb)[0] for _ in range(steps)]
        value = float(np.mean(ms))
        kind, sample = "port", f"oracle port (oracle/_ref not built), full 2 x {T_full} tris, {steps} frames, 1 thread"
        extra.update({"same_config": True, "steps_run": steps})
    if oracle.ref_available():
        # the UNMODIFIED library where it is valid: configs[0] and its largest size (tree depth 18)
        R = oracle.Ref()
        for tag, (nu, nv) in (("configs0_unmodified", (NU0, NV0)), ("largest_unmodified", (512, 256))):
            p, f = make_meshes(nu, nv, shuffle=False)
            t, n = reference_cpu_frames(R, p, f, 3, 1)
            extra[tag] = {"value": float(np.mean(t)), "unit": UNIT, "tris_per_mesh": int(len(f)), "pairs": n,
                          "kind": "reference (unmodified)", "frames": 3}
    line.update({"value": value, "ms_per_step": value,
                 "config": {"workload": "two 2^20-triangle synthetic blobs (bunny stand-in), rigid 1 deg/frame rotation "
                                        "of body B, full pipeline per frame: build x2 + refit x2 + broad + narrow "
                                        "(BASELINE.json configs[1]); reference CPU path (SimpleBVH + SimpleCollide)",
                            "tris_per_mesh": T_full, "meshes": 2},
                 "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": kind, "sample": sample,
                                  "host": host_info()},
                 "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
    line.update(extra)
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------------------
# the reference GPU path (unmodified .cu files recompiled for sm_100a), a reported baseline on a bounded sample
# --------------------------------------------------------------------------------------------------------------
REF_GPU_EXE = os.path.join(ROOT, "oracle", "_ref", "ref_gpu_bench")
REF_GPU_NU, REF_GPU_NV = 256, 192


def reference_gpu_baseline(pos, faces, frames=12, keep=None):
    """The UNMODIFIED reference GPU path (its five .cu files recompiled for sm_100a by `make -C oracle refgpu`,
    driven by oracle/ref_gpu_main.cu) on the same meshes and the same frame, on this GPU: the recompiled kernels
    this library replaces. A reported baseline like cpu_baseline -- never on the product path. Returns None when the
    binary was not prebuilt (it needs /root/reference at build time)."""
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


# --------------------------------------------------------------------------------------------------------------
# GPU arm
def bind_to_gpu_cpus(dev):
    """N > 1: run this rank on the CPUs next to its GPU (NVML's CPU affinity of the device), BEFORE any pinned buffer is
    allocated, so that the end-to-end uploads are served by the memory of the GPU's own socket. Without it every rank's
    pinned buffers land wherever torchrun happened to start the process, and at N = 8 half of the uploads cross the
    socket interconnect. Best effort: returns a short description, or None when NVML / affinity are not available."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(dev)
        ncpu = os.cpu_count() or 64
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return f"{len(cpus)} CPUs ({min(cpus)}-{max(cpus)})"
    except Exception:
        return None


# --------------------------------------------------------------------------------------------------------------
class Gpu:
    """device, context, stream and the rank plumbing shared by every scenario"""

    def __init__(self, args):
        import torch
        import oibvh_b200 as ob
        self.torch, self.ob = torch, ob
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if ob.device_count() < 1:
            raise SystemExit("bench.py needs a CUDA device: oibvh_b200 has no CPU fallback")
        self.dist = None
        if self.world > 1:
            import datetime
            import torch.distributed as dist
            torch.cuda.set_device(local_rank)
            # a collective that does not complete within three minutes aborts the job instead of hanging it
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank),
                                    timeout=datetime.timedelta(seconds=180))
            self.dist = dist
        assert self.world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={self.world} (launch with torchrun for N>1)"
        self.dev = local_rank if self.world > 1 else 0
        torch.cuda.set_device(self.dev)
        self.numa = bind_to_gpu_cpus(self.dev) if self.world > 1 else None
        self.ctx = ob.Context(self.dev)
        self.stream = torch.cuda.ExternalStream(self.ctx.stream, device=self.dev)

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, x):
        if self.dist is None:
            return float(x)
        t = self.torch.tensor([float(x)], dtype=self.torch.float64, device=f"cuda:{self.dev}")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def events(self):
        return self.torch.cuda.Event(enable_timing=True), self.torch.cuda.Event(enable_timing=True)

    def timed(self, fn, n):
        """n calls of fn on the context stream between two CUDA events, barrier + synchronize on both sides; ms per call,
        max over ranks"""
        e0, e1 = self.events()
        self.barrier()
        e0.record(self.stream)
        for _ in range(n):
            fn()
        e1.record(self.stream)
        self.barrier()
        return self.max_over_ranks(e0.elapsed_time(e1) / n)


class Scenario:
    """one BASELINE.json config: trees + scene + the per-frame work, measured the same way for every config"""

    def __init__(self, g, name, workload):
        self.g, self.name, self.workload = g, name, workload
        self.trees, self.scene = [], None
        self.frames_per_graph = 1
        self.meta = {}

    # -- to be provided by the builder functions below --
    def frame(self):
        raise NotImplementedError

    def attach(self):
        """size the queues from an unsharded detection of the initial state, remember its pair set, then (N > 1) enter
        multi-GPU mode: the gathered set of the same state must equal it"""
        g, sc = self.g, self.scene
        sc.set_shard(0, 1)
        sc.detect_async(ENTRY_LEVEL, EXPAND_LEVELS)
        n_pairs, n_cand = sc.counts()
        rounds = [r for r in sc.round_stats() if r]
        self.single = {"pairs": n_pairs, "candidates": n_cand, "hash": pair_hash(sc.canonical_pairs())}
        if g.world > 1:
            from oibvh_b200 import distributed as obd
            sc.reserve(4 * max(rounds + [1]) + (1 << 16), 4 * n_cand + (1 << 16), 4 * n_pairs + (1 << 16))
            obd.attach(sc, g.rank, g.world)
            sc.detect_async(ENTRY_LEVEL, EXPAND_LEVELS)
            n_gathered, _ = sc.counts()
            if g.rank == 0:
                h = pair_hash(sc.canonical_pairs())
                self.meta["matches_single_gpu"] = bool(h == self.single["hash"] and n_gathered == n_pairs)
                self.meta["pair_hash_sharded"] = h
            g.barrier()

    def measure(self, steps, warmup):
        """graph replay (value) + eager pass with CUDA events around every stage and kernel"""
        g, sc, ctx = self.g, self.scene, self.g.ctx
        for _ in range(max(warmup, 3)):
            self.frame()
            sc.counts()
        ctx.capture_begin()
        for _ in range(self.frames_per_graph):
            self.frame()
        graph = ctx.capture_end()
        for _ in range(2):
            graph.launch()
            sc.counts()
        launches = max(1, steps // self.frames_per_graph)
        l0 = ctx.launch_count()
        ms = g.timed(graph.launch, launches) / self.frames_per_graph
        l1 = ctx.launch_count()
        n_pairs, n_cand = sc.counts()
        g.barrier()
        graph.close()
        # eager pass: stage clocks (max over ranks per stage)
        ctx.enable_timing(True)
        acc = {}
        reps = max(3, min(steps, 30))
        for _ in range(reps):
            self.frame()
            st = ctx.stage_ms()
            sc.counts()
            for k, v in st.items():
                acc[k] = acc.get(k, 0.0) + v
        ctx.enable_timing(False)
        stage = {k: g.max_over_ranks(v / reps) for k, v in acc.items()}
        rounds = [r for r in sc.round_stats() if r]
        cyc = sc.phase_cycles()
        # broad and narrow phase are ONE kernel and the narrow phase runs inside the traversal's warps (a leaf pair is
        # tested by the warp that finds it): stage["broad"] is the whole detection, stage["narrow"] stays 0
        g.barrier()
        return {"frame_ms": ms, "stage_ms": stage, "pairs": int(n_pairs), "candidates": int(n_cand), "bvtt_rounds": rounds,
                "collide_phase_cycles": cyc, "gpu_launches": int(l1 - l0), "frames_timed": launches * self.frames_per_graph}

    def close(self):
        if self.g.world > 1 and self.scene is not None:
            from oibvh_b200 import distributed as obd
            obd.detach(self.scene)
        if self.scene is not None:
            self.scene.close()
        for t in self.trees:
            t.close()
        self.trees, self.scene = [], None


def two_body(g, nu, nv, name, workload):
    ob = g.ob
    s = Scenario(g, name, workload)
    pos, faces = make_meshes(nu, nv)
    s.pos, s.faces = pos, faces
    s.mesh_a = ob.Mesh(pos, faces)
    s.mesh_b = s.mesh_a.copy()
    ta = ob.OibvhTree(s.mesh_a, ctx=g.ctx)
    ta.build()
    tb = ob.OibvhTree(ta, s.mesh_b)
    M0 = s.mesh_b.transform_matrix_translate(OFFSET_B)
    s.mesh_b.transform(M0)
    tb.transform(M0)
    s.M_rot = s.mesh_b.transform_matrix_rotate((0.0, 0.0, 1.0), 1.0)  # about B's centre, which the rotation fixes
    tb.build()
    s.trees = [ta, tb]
    s.scene = ob.Scene(g.ctx)
    s.scene.addOibvhTree(ta)
    s.scene.addOibvhTree(tb)

    mats = np.stack([ob.mat_identity(), s.M_rot]).astype(np.float32)

    def frame():
        ob.build_many([ta, tb])
        # rigid transform of B + refit of both: B's transform runs under A's refit (two streams, fork / join)
        ob.transform_refit_many([ta, tb], mats, apply=[False, True])
        s.scene.detect_async(ENTRY_LEVEL, EXPAND_LEVELS)
    s.frame = frame
    s.meta.update({"tris_per_mesh": int(len(faces)), "verts_per_mesh": int(len(pos)), "meshes": 2,
                   "per_frame": "build x2 + transform + refit x2 + broad + narrow"})
    return s


def deforming(g):
    """configs[2]: a 4 M-triangle deforming mesh (per-frame sinusoidal vertex displacement: refit only, topology fixed)
    against a static 327 680-triangle obstacle"""
    ob, torch = g.ob, g.torch
    mg = load_meshgen()
    s = Scenario(g, "configs[2]", "deforming 4 194 304-triangle blob vs static 327 680-triangle icosphere: per frame "
                                  "new vertex positions (device-resident) + refit of the deforming mesh + broad + narrow")
    pos, faces = mg.blob(2048, 1024, seed=11)
    faces = mg.shuffle_faces(faces, seed=3)
    opos, ofaces = mg.icosphere(7, radius=0.7, center=(1.35, 0.1, 0.0))
    s.pos, s.faces, s.opos, s.ofaces = pos, faces, opos, ofaces
    td = ob.OibvhTree(ob.Mesh(pos, faces), ctx=g.ctx)
    to = ob.OibvhTree(ob.Mesh(opos, ofaces), ctx=g.ctx)
    td.build()
    to.build()
    s.trees = [td, to]
    s.scene = ob.Scene(g.ctx)
    s.scene.addOibvhTree(td)
    s.scene.addOibvhTree(to)
    s.frames_per_graph = 4
    s.deform_host = [mg.cloth_positions(pos, f, amp=0.03) for f in range(s.frames_per_graph)]
    s.deform_dev = [torch.from_numpy(p).to(f"cuda:{g.dev}") for p in s.deform_host]
    s.k = 0

    def frame():
        td.set_positions_from_device(s.deform_dev[s.k % s.frames_per_graph].data_ptr())
        s.k += 1
        td.refit(upload=False)
        s.scene.detect_async(ENTRY_LEVEL, EXPAND_LEVELS)
    s.frame = frame
    s.meta.update({"tris": [int(len(faces)), int(len(ofaces))], "per_frame": "positions + refit(4 M) + broad + narrow"})
    return s


def many_body(g, n=4096):
    """configs[3]: n instanced small bodies (cubes / icospheres / small blobs) on a jittered grid so that neighbours touch;
    per frame one rigid transform per body, refit of every tree, inter-object broad + narrow phase"""
    ob, torch = g.ob, g.torch
    mg = load_meshgen()
    s = Scenario(g, "configs[3]", f"{n} instanced bodies (cubes, icospheres, small blobs) in a box, per-object trees: per "
                                  "frame transform_many + refit_many + inter-object broad + narrow")
    rng = np.random.default_rng(0)
    side = int(np.ceil(n ** (1.0 / 3.0)))
    protos = [mg.cube(), mg.icosphere(1), mg.icosphere(2), mg.icosphere(3), mg.blob(48, 32, seed=3)]
    s.meshes = []
    for i in range(n):
        p0, f0 = protos[int(rng.integers(len(protos)))]
        cell = np.array([i % side, (i // side) % side, i // (side * side)], np.float32)
        c = (cell + rng.uniform(-0.15, 0.15, 3)).astype(np.float32)
        scale = np.float32(rng.uniform(0.35, 0.6))
        m = ob.Mesh((p0 * scale + c).astype(np.float32), f0)
        s.meshes.append(m)
        s.trees.append(ob.OibvhTree(m, ctx=g.ctx))
    batch = ob.TreeBatch(s.trees)
    ob.build_many(batch)
    s.scene = ob.Scene(g.ctx)
    for t in s.trees:
        s.scene.addOibvhTree(t)
    rng = np.random.default_rng(1)
    s.mats = np.stack([m.transform_matrix_rotate(rng.normal(size=3).astype(np.float32), 0.5) for m in s.meshes])
    s.dmats = torch.from_numpy(s.mats.reshape(-1, 16).copy()).to(f"cuda:{g.dev}")
    ob.transform_many(batch, None, device_ptr=s.dmats.data_ptr())  # one eager call builds the cached tables
    ob.refit_many(batch)

    def frame():
        ob.transform_many(batch, None, device_ptr=s.dmats.data_ptr())
        ob.refit_many(batch)
        s.scene.detect_async(ENTRY_LEVEL, EXPAND_LEVELS)
    s.frame = frame
    s.meta.update({"bodies": n, "tris": int(sum(t.info()[0] for t in s.trees)),
                   "per_frame": "transform_many + refit_many + broad + narrow"})
    return s


def terrain_scene(g):
    """configs[4]: 16.8 M-triangle static terrain vs a 1 M-triangle body pressed into it (deep trees, large BVTT front);
    per frame the body moves (rigid transform + refit), the terrain tree is built once"""
    ob = g.ob
    mg = load_meshgen()
    s = Scenario(g, "configs[4]", "16 785 218-triangle static terrain vs 1 048 576-triangle body pressed into it: per frame "
                                  "transform + refit of the body + broad + narrow (terrain tree built once)")
    tpos, tfaces = mg.terrain(2897, 2897, height=0.3, size=(8.0, 8.0))
    bpos, bfaces = mg.blob(1024, 512, seed=5, radius=1.5, center=(0.2, 0.9, -0.3))
    s.tpos, s.tfaces, s.bpos, s.bfaces = tpos, tfaces, bpos, bfaces
    tt = ob.OibvhTree(ob.Mesh(tpos, tfaces), ctx=g.ctx)
    s.mesh_b = ob.Mesh(bpos, bfaces)
    tb = ob.OibvhTree(s.mesh_b, ctx=g.ctx)
    e0, e1 = g.events()
    tt.build()
    g.ctx.synchronize()
    e0.record(g.stream)
    tt.build()
    e1.record(g.stream)
    g.ctx.synchronize()
    g.torch.cuda.synchronize()
    s.terrain_build_ms = e0.elapsed_time(e1)
    tb.build()
    s.trees = [tt, tb]
    s.scene = ob.Scene(g.ctx)
    s.scene.addOibvhTree(tt)
    s.scene.addOibvhTree(tb)
    s.M_rot = s.mesh_b.transform_matrix_rotate((0.0, 1.0, 0.0), 0.5)

    def frame():
        tb.transform(s.M_rot)
        tb.refit(upload=False)
        s.scene.detect_async(ENTRY_LEVEL, EXPAND_LEVELS)
    s.frame = frame
    s.meta.update({"tris": [int(len(tfaces)), int(len(bfaces))], "per_frame": "transform + refit(1 M body) + broad + narrow"})
    return s


def coherence_record(g, s, n=20):
    """f4 beside the from-the-roots traversal, on a refit-only config: the same trees in a second, coherent scene (it
    records a BVTT cut once and starts every detection from it); detection-only times, CUDA events, same pair set"""
    ob = g.ob
    coh = ob.Scene(g.ctx)
    for t in s.trees:
        coh.addOibvhTree(t)
    coh.reserve(candidate_records=1 << 22)
    coh.set_coherence(True)
    plain = ob.Scene(g.ctx)
    for t in s.trees:
        plain.addOibvhTree(t)
    out = {}
    for name, sc in (("from_roots", plain), ("from_recorded_cut", coh)):
        for _ in range(3):
            sc.detect_async(ENTRY_LEVEL, EXPAND_LEVELS)
            sc.counts()
        out[name + "_ms"] = g.timed(lambda: sc.detect_async(ENTRY_LEVEL, EXPAND_LEVELS), n)
        out[name + "_pairs"] = int(sc.counts()[0])
    out["pair_set_equal"] = bool(np.array_equal(plain.canonical_pairs(), coh.canonical_pairs()))
    plain.close()
    coh.close()
    return out


def cpu_refit_detect(s, moving, label):
    """CPU path beside a refit-only config, rank 0 at N = 1: the oracle port refits the moving trees from the positions
    the device holds NOW and detects on the same trees; also the parity check of this config at full size (node arrays
    bit-exact, pair sets equal). Static trees are taken as the device built them (their builds are covered by
    tests/test_gpu_large.py); one frame, single thread."""
    import oracle
    port = oracle.Port()
    sc = s.scene
    sc.detect_async(ENTRY_LEVEL, EXPAND_LEVELS)
    n_gpu, c_gpu = sc.counts()
    gpu_pairs = sc.canonical_pairs()
    objs, perms, nodes_ok = [], [], True
    t_cpu = 0.0
    for i, t in enumerate(s.trees):
        d = t.download()
        pos = t.m_positions
        if i in moving:
            t0 = time.perf_counter()
            nodes = port.refit(pos, d["faces"])
            t_cpu += time.perf_counter() - t0
            nodes_ok = nodes_ok and np.array_equal(nodes.view(np.uint32), d["nodes"].view(np.uint32))
        else:
            nodes = d["nodes"]
        objs.append((nodes, d["faces"], pos))
        perms.append(d["perm"])
    t0 = time.perf_counter()
    pp, nc = port.detect(objs)
    t_cpu += time.perf_counter() - t0
    same = bool(np.array_equal(gpu_pairs, oracle.canonical_pairs(pp, perms)) and nc == c_gpu)
    return ({"value": t_cpu * 1e3, "unit": UNIT, "cores": 1, "kind": "port", "sample": label + ", 1 frame, 1 thread",
             "pairs": int(len(pp)), "candidates": int(nc)},
            {"nodes_bit_exact": bool(nodes_ok), "pair_set_equal": same, "checked_against": "oracle port, full size"})


def run_gpu_arm(args):
    g = Gpu(args)
    ob, torch, ctx, dist = g.ob, g.torch, g.ctx, g.dist
    rank, world, dev = g.rank, g.world, g.dev
    want = set(int(x) for x in args.configs.split(",")) if args.configs else {0, 1, 2, 3, 4}
    want.add(1)
    peak, peak_src = measured_peak()
    traffic = committed_traffic()

    def frac(nbytes, ms):
        gbs = nbytes / (ms * 1e-3) / 1e9 if ms > 0 else 0.0
        return {"bytes": int(nbytes), "ms": ms, "achieved": gbs, "frac": gbs / peak}

    # ==========================================================================================================
    # configs[1]: the headline
    # ==========================================================================================================
    s1 = two_body(g, NU, NV, "configs[1]", "two 2^20-triangle synthetic blobs (bunny stand-in), rigid 1 deg/frame rotation "
                  "of body B, full pipeline per frame: build x2 + refit x2 + broad + narrow (BASELINE.json configs[1])")
    T, V = len(s1.faces), len(s1.pos)
    tree_a, tree_b, scene = s1.trees[0], s1.trees[1], s1.scene
    s1.attach()
    sampler = ClockSampler(dev)
    sampler.start()
    time.sleep(0.3)
    w0 = time.time()
    m1 = s1.measure(args.steps, args.warmup)
    w1 = time.time()
    clocks = sampler.stop(w0, w1)
    ms_per_step = m1["frame_ms"]
    stage = m1["stage_ms"]

    # the refit kernel on its own: K back-to-back launches on one mesh
    tree_a.refit(upload=False)
    ctx.synchronize()
    refit_alone_ms = g.timed(lambda: tree_a.refit(upload=False), max(args.steps, 10))

    # ---- end to end through the C ABI with HOST buffers (pinned): H2D of both position arrays + D2H of the result ----
    n_e2e = args.steps
    host_a = torch.from_numpy(s1.mesh_a.m_positions.copy()).pin_memory()
    rot_frames = []
    mb = s1.mesh_b.copy()
    for _ in range(4):  # a short cycle of distinct host position buffers for body B
        mb.transform(s1.M_rot)
        rot_frames.append(torch.from_numpy(mb.m_positions.copy()).pin_memory())
    pair_host = torch.empty((max(8 * m1["pairs"], 1 << 16), 4), dtype=torch.int32).pin_memory()

    # N > 1: every rank uploads both position arrays itself (the BVH is replicated by recomputation). Eight ranks pulling
    # 12.6 MB each from host memory contend for it (e2e 0.53 ms at N = 1, 0.77 ms at N = 8); crossing PCIe once on rank 0
    # and broadcasting over NVLink with NCCL was measured SLOWER (0.82 ms: the collective couples the ranks every frame).
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
        if rank == 0:
            # one C-ABI call: waits for the frame (at N > 1: for every rank's hits), reads the counters, copies the
            # gathered pair list into the pinned buffer
            return scene.get_pairs_into(pair_host.data_ptr())
        return scene.counts()[0]  # the other ranks only pace themselves on their own frame

    for i in range(2):
        e2e_frame(i)
    g.barrier()
    t0 = time.perf_counter()
    tot_pairs = 0
    for i in range(n_e2e):
        tot_pairs += e2e_frame(i)
    g.barrier()
    e2e_ms = g.max_over_ranks((time.perf_counter() - t0) * 1e3 / n_e2e)
    h2d = 2 * 12 * V
    d2h = 4 * 128 + 16 * (tot_pairs // max(n_e2e, 1))
    # the same loop with the NEXT step's host->device copies enqueued before this step's result is awaited (every
    # step still uploads its own inputs and reads its own result; reported beside e2e, not instead of it)
    e2e_upload(0)
    e2e_frame(0, upload=False, prefetch=1)
    g.barrier()
    t0 = time.perf_counter()
    for i in range(1, n_e2e + 1):
        e2e_frame(i, upload=False, prefetch=i + 1)
    g.barrier()
    e2e_pipe_ms = g.max_over_ranks((time.perf_counter() - t0) * 1e3 / n_e2e)
    ctx.synchronize()

    line = None
    if rank == 0:
        # per-kernel table: bytes per frame / CUDA-event time of the eager pass (both trees' launches of a
        # kind run concurrently on two streams, so the pair is timed together)
        det_ms = stage["broad"] + stage["narrow"]
        kern = {
            "morton_hist_kernel x2 (keys)": frac(2 * bytes_keys(T, V), stage["keys"]),
            "lsd_sort_kernel x1 (both trees, 3 x 10-bit passes)": frac(2 * bytes_sort(T), stage["sort"]),
            "tree_emit_kernel<build> x2": frac(2 * bytes_emit(T, V), stage["emit"]),
            "tree_emit_kernel<refit> x2 + transform_kernel x1 (B's transform runs under A's refit)":
                frac(2 * bytes_refit(T, V) + bytes_transform(V), stage["refit"]),
            "collide_kernel x1 (broad + narrow)": frac(bytes_detect(m1["bvtt_rounds"], m1["candidates"], m1["pairs"]), det_ms),
        }
        for k, v in kern.items():
            v["traffic"] = traffic.get(k.split(" ")[0])
            v["share_of_frame"] = v["ms"] / sum(x["ms"] for x in kern.values())
        dominant = max(kern, key=lambda k: kern[k]["ms"])
        dk = kern[dominant]
        refit_ms = stage["refit"] / 2.0
        build_ms = stage["build"] / 2.0
        line = {
            "metric": METRIC, "value": ms_per_step, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": False,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": s1.workload, "tris_per_mesh": int(T), "verts_per_mesh": int(V), "meshes": 2,
                       "entry_level": ENTRY_LEVEL, "expand_levels": EXPAND_LEVELS,
                       "l2": "no explicit flush: one frame streams ~190 MB of distinct buffers (> 126 MB L2)",
                       "parallelism": (f"replicated BVH, BVTT front of round 0 dealt to {world} GPUs, every rank's narrow phase "
                                       "appends to rank 0's pair list over NVLink (peer mapping, no collective per frame)")
                       if world > 1 else "single GPU"},
            "frame_mtris_per_s": 2 * T / (ms_per_step * 1e-3) / 1e6,
            "build_mtris_per_s": T / (build_ms * 1e-3) / 1e6 if build_ms > 0 else None,
            "stage_ms": stage, "collide_phase_cycles": m1["collide_phase_cycles"], "pairs": m1["pairs"],
            "candidates": m1["candidates"], "bvtt_rounds": m1["bvtt_rounds"],
            "sharded_phase_ms": det_ms,
            "matches_single_gpu": s1.meta.get("matches_single_gpu"), "pair_hash": s1.single["hash"],
            "gpu_launches": m1["gpu_launches"],
            "clocks": clocks,
            "e2e": {"value": e2e_ms, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)},
            "e2e_double_buffered": {"value": e2e_pipe_ms, "unit": UNIT,
                                    "note": "next step's H2D enqueued under this step's kernels; same bytes per step"},
            "cpu_binding": g.numa,
            "roofline": {"bound": "hbm", "kernel": dominant + " -- the kernel with the largest share of the frame",
                         "achieved": dk["achieved"], "peak": peak, "unit": "GB/s", "frac": dk["frac"],
                         "traffic": dk["traffic"], "traffic_source": "committed ncu --set full capture "
                         "(profiles/r02_traffic.json), not measured in this run" if dk["traffic"] else None,
                         "peak_source": peak_src, "algorithmic_bytes_per_launch": dk["bytes"], "ms_per_launch": dk["ms"],
                         "timing": "CUDA events on the launching stream around the kernel in an eager pass over the same "
                                   "frames (the graph-replayed frame is the headline value)"},
            "kernels": kern,
            "roofline_refit": dict(frac(bytes_refit(T, V), refit_ms),
                                   stage="refit (per tree; two launches share the machine, ms = stage / 2)",
                                   one_launch_alone=frac(bytes_refit(T, V), refit_alone_ms)),
            "roofline_build": dict(frac(bytes_build(T, V), build_ms),
                                   stage="build (per tree) = keys + cooperative 3-pass radix sort (both trees in one "
                                         "launch) + emit; ms = stage / 2"),
            "stage_share": {k: (stage[k] / sum(stage[x] for x in ob.STAGES)) for k in ob.STAGES},
        }
        if world == 1 and not args.no_cpu_baseline:
            # the CPU baseline runs on the positions body B has NOW (after every rotation of the timed regions); the
            # GPU count for exactly these positions is reported beside its pair count
            posB = tree_b.m_positions
            ob.build_many([tree_a, tree_b])
            scene.detect_async(ENTRY_LEVEL, EXPAND_LEVELS)
            gpu_now = scene.counts()
            line["cpu_baseline"] = cpu_baseline_port(s1.pos, s1.faces, posB, s1.mesh_a.m_aabb)
            line["cpu_baseline"]["gpu_pairs_same_positions"] = int(gpu_now[0])
            line["cpu_baseline"]["gpu_candidates_same_positions"] = int(gpu_now[1])
            line["parity"] = {"pair_count_equal": bool(gpu_now[0] == line["cpu_baseline"]["pairs"] and
                                                       gpu_now[1] == line["cpu_baseline"]["candidates"]),
                              "checked_against": "oracle port, full size (bit-exact trees and pair sets: tests/)"}
            # The unmodified reference GPU path only completes this scene up to ~10^5 triangles per body: it emits
            # BVTT children untested, so at 2 x 196 608 triangles its front outgrows its fixed 10 M-node buffers and
            # its unchecked level loop never terminates (measured on B200). Bounded sample: 2 x 98 304 triangles.
            s_pos, s_faces = make_meshes(REF_GPU_NU, REF_GPU_NV)
            ref_gpu = reference_gpu_baseline(s_pos, s_faces)
            if ref_gpu is not None:
                ref_gpu["sample"] = (f"2 x {len(s_faces)} triangles (largest size of this scene the reference GPU path "
                                     "completes; it overflows its fixed 10M-node BVTT buffers at 2 x 196608)")
                if "unavailable" not in ref_gpu:
                    sref = two_body(g, REF_GPU_NU, REF_GPU_NV, "reference-gpu sample", "")
                    sref.attach()
                    mm = sref.measure(20, 3)
                    ref_gpu["ours_same_sample"] = {
                        "value": mm["frame_ms"], "unit": UNIT, "pairs_last_frame": mm["pairs"],
                        "timing": "CUDA events around graph replays, inputs resident (the reference's figure includes "
                                  "its own host<->device copies, which are part of its API)"}
                    sref.close()
                line["reference_gpu"] = ref_gpu
    s1.close()

    # ==========================================================================================================
    # the other BASELINE.json configs: one sub-record each
    # ==========================================================================================================
    records = []
    sub_steps = max(8, min(args.steps, 40))

    def finish(s, m, extra):
        rec = {"config": s.name, "workload": s.workload}
        rec.update(s.meta)
        rec.update({"ms_per_frame": m["frame_ms"], "stage_ms": m["stage_ms"],
                    "sharded_phase_ms": m["stage_ms"]["broad"] + m["stage_ms"]["narrow"], "pairs": m["pairs"],
                    "candidates": m["candidates"], "bvtt_rounds": m["bvtt_rounds"], "gpu_launches": m["gpu_launches"],
                    "frames_timed": m["frames_timed"], "pair_hash_initial_state": s.single["hash"]})
        rec.update(extra)
        records.append(rec)

    if 0 in want:
        s = two_body(g, NU0, NV0, "configs[0]", "two 34 816-triangle blobs (the ~70 K-triangle bunny pair of main.cpp:127-151), "
                     "full pipeline per frame like configs[1]")
        s.attach()
        m = s.measure(sub_steps, 3)
        extra = {}
        if rank == 0:
            T0, V0 = len(s.faces), len(s.pos)
            extra["roofline"] = {"build": frac(bytes_build(T0, V0), m["stage_ms"]["build"] / 2),
                                 "refit": frac(bytes_refit(T0, V0), m["stage_ms"]["refit"] / 2)}
            if world == 1 and not args.no_cpu_baseline:
                import oracle
                posB = s.trees[1].m_positions
                ob.build_many(s.trees)
                s.scene.detect_async(ENTRY_LEVEL, EXPAND_LEVELS)
                have = s.scene.canonical_pairs()
                port = oracle.Port()
                ms = [cpu_frame_port(port, s.pos, s.faces, posB, s.mesh_a.m_aabb) for _ in range(3)]
                extra["cpu_baseline"] = {"value": float(np.median([x[0] for x in ms])), "unit": UNIT, "cores": 1,
                                         "kind": "port", "sample": "full config, 3 frames, 1 thread"}
                if oracle.ref_available():
                    # The UNMODIFIED reference CPU classes on the same meshes and poses, faces in GENERATOR order (on
                    # the shuffled order SimpleBVH's input-order boxes degenerate: 22 s for this config); its face ids
                    # are mapped through the shuffle so that the two pair sets compare record for record.
                    R = oracle.Ref()
                    gen_pos, gen_faces = make_meshes(NU0, NV0, shuffle=False)
                    shuffle = np.random.default_rng(7).permutation(len(gen_faces))  # meshgen.shuffle_faces(seed=7)
                    assert np.array_equal(gen_faces[shuffle], s.faces)
                    t0 = time.perf_counter()
                    ref_pairs = R.detect_meshes([(gen_pos, gen_faces), (posB, gen_faces)])
                    ref_ms = (time.perf_counter() - t0) * 1e3
                    mapped = have.copy()
                    mapped[:, 2], mapped[:, 3] = shuffle[have[:, 2]], shuffle[have[:, 3]]
                    extra["cpu_reference_unmodified"] = {
                        "value": ref_ms, "unit": "ms (SimpleBVH build x2 + SimpleCollide::detect)", "kind": "reference",
                        "cores": 1, "note": "unmodified reference CPU classes, same meshes and poses as the GPU frame"}
                    extra["parity"] = {"pair_set_equal": bool(np.array_equal(oracle.canonical_pairs(mapped),
                                                                             oracle.canonical_pairs(ref_pairs))),
                                       "checked_against": "unmodified reference CPU classes (oracle/_ref), full size"}
        finish(s, m, extra)
        s.close()

    if 2 in want:
        s = deforming(g)
        s.attach()
        m = s.measure(sub_steps, 3)
        extra = {}
        if rank == 0:
            Td, Vd = len(s.faces), len(s.pos)
            extra["roofline"] = {"refit": frac(bytes_refit(Td, Vd), m["stage_ms"]["refit"])}
            if world == 1:
                extra["temporal_coherence"] = coherence_record(g, s)
            if world == 1 and not args.no_cpu_baseline:
                extra["cpu_baseline"], extra["parity"] = cpu_refit_detect(s, {0}, "refit(4 M) + detect")
        finish(s, m, extra)
        s.close()
        del s

    if 3 in want:
        s = many_body(g, args.bodies)
        s.attach()
        m = s.measure(sub_steps, 3)
        extra = {}
        if rank == 0:
            tot_T = sum(t.info()[0] for t in s.trees)
            tot_V = sum(t.info()[1] for t in s.trees)
            tot_N = sum(t.info()[2] for t in s.trees)
            extra["roofline"] = {"refit_many": frac(12 * tot_T + 12 * tot_V + 24 * tot_N, m["stage_ms"]["refit"])}
            extra["tris"] = int(tot_T)
            if world == 1 and not args.no_cpu_baseline:
                extra["cpu_baseline"], extra["parity"] = cpu_refit_detect(s, set(range(len(s.trees))),
                                                                          f"refit x{len(s.trees)} + detect")
        finish(s, m, extra)
        s.close()
        del s

    if 4 in want:
        s = terrain_scene(g)
        s.attach()
        m = s.measure(sub_steps, 3)
        extra = {}
        if rank == 0:
            Tt, Vt = len(s.tfaces), len(s.tpos)
            Tb, Vb = len(s.bfaces), len(s.bpos)
            extra["roofline"] = {"terrain_build_once": frac(bytes_build(Tt, Vt), s.terrain_build_ms),
                                 "refit_body": frac(bytes_refit(Tb, Vb), m["stage_ms"]["refit"])}
            if world == 1:
                extra["temporal_coherence"] = coherence_record(g, s)
            if world == 1 and not args.no_cpu_baseline:
                extra["cpu_baseline"], extra["parity"] = cpu_refit_detect(s, {1}, "refit(1 M body) + detect vs 16.8 M terrain")
        finish(s, m, extra)
        s.close()
        del s

    if rank == 0:
        line["configs"] = records
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--configs", default="", help="comma-separated BASELINE.json config indices to measure (default: all; "
                                                  "configs[1], the headline, always runs)")
    ap.add_argument("--bodies", type=int, default=4096, help="bodies of the many-body scene (configs[3])")
    args = ap.parse_args()
    # a wedged collective or GPU must not hold the driver for ever: fail the process loudly after ten minutes
    def _watchdog():
        sys.stderr.write("bench.py: no result after 900 s -- aborting (rank %s)\n" % os.environ.get("RANK", "0"))
        sys.stderr.flush()
        os._exit(3)
    wd = threading.Timer(900.0, _watchdog)
    wd.daemon = True
    wd.start()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
