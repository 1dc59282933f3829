"""The headless scenario driver (apps/oibvh_headless) on the five BASELINE.json scenario
shapes at reduced size: the pair set of the last frame, in canonical form, must equal the CPU oracle's on the scene
the driver dumped (positions as the detection saw them)."""
import os
import struct
import subprocess

import numpy as np
import pytest

import oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "apps", "oibvh_headless")


def read_scene(path):
    raw = open(path, "rb").read()
    (n,), off = struct.unpack_from("<I", raw, 0), 4
    bodies = []
    for _ in range(n):
        V, T = struct.unpack_from("<II", raw, off)
        off += 8
        pos = np.frombuffer(raw, np.float32, 3 * V, off).reshape(V, 3).copy()
        off += 12 * V
        faces = np.frombuffer(raw, np.uint32, 3 * T, off).reshape(T, 3).copy()
        off += 12 * T
        aabb = np.frombuffer(raw, np.float32, 6, off).copy()
        off += 24
        bodies.append((pos, faces, aabb))
    assert off == len(raw)
    return bodies


def run_driver(tmp_path, *args):
    if not os.path.exists(EXE):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "apps")])
    dump = tmp_path / "pairs.bin"
    res = subprocess.run([EXE, *args, "--dump", str(dump)], capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout + res.stderr
    frames = [l.split() for l in res.stdout.splitlines() if l.startswith("frame")]
    pairs = np.fromfile(dump, dtype=np.uint32).reshape(-1, 4)
    return frames, pairs, read_scene(str(dump) + ".scene"), res.stdout


def oracle_pairs(port, bodies):
    built = [port.build(p, f, a) for p, f, a in bodies]
    got, ncand = port.detect([(b["nodes"], b["faces"], p) for b, (p, _, _) in zip(built, bodies)])
    return oracle.canonical_pairs(got, [b["perm"] for b in built]), ncand


@pytest.mark.gpu
@pytest.mark.parametrize("args", [
    ("--config", "1", "--frames", "2"),                                  # ~35 K triangles x 2, refit + detect per frame
    ("--config", "2", "--scale", "0.02", "--frames", "2"),               # full pipeline per frame (rebuild + refit)
    ("--config", "3", "--scale", "0.01", "--frames", "2"),               # deforming mesh vs static obstacle
    ("--config", "4", "--bodies", "40", "--frames", "2"),                # many bodies, one launch per step
    ("--config", "5", "--scale", "0.002", "--frames", "2"),              # terrain vs body
    ("--config", "1", "--scale", "0.05", "--subdivide", "1", "--frames", "1", "--entry", "0", "--expand", "1"),
], ids=["bunny_pair", "rebuild_per_frame", "deforming", "many_body", "terrain", "loop_subdivided"])
def test_scenarios_match_oracle(tmp_path, port, ctx, args):
    frames, pairs, bodies, out = run_driver(tmp_path, *args)
    want, ncand = oracle_pairs(port, bodies)
    assert len(frames) >= 2, out
    assert int(frames[-1][3]) == len(pairs) == len(want), (frames[-1], len(want))
    assert int(frames[-1][5]) == ncand
    assert np.array_equal(pairs, want)
    assert len(want) > 0, "scenario does not collide: the check would be vacuous"


@pytest.mark.gpu
def test_obj_round_trip_through_driver(tmp_path, ctx):
    """--dump-mesh writes the body as .obj; loading that file back gives the same frame-by-frame counts"""
    obj = tmp_path / "body.obj"
    a = run_driver(tmp_path, "--config", "1", "--scale", "0.05", "--frames", "2", "--dump-mesh", str(obj))
    b = run_driver(tmp_path, "--config", "1", "--frames", "2", "--obj", str(obj))
    assert [f[2:6] for f in a[0]] == [f[2:6] for f in b[0]]
    assert np.array_equal(a[1], b[1])
