"""bench.py keeps the driver's contract: one JSON line with the agreed keys, measured through the CUDA path."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_bench_line_has_the_contract_keys(ctx):
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "12", "--warmup", "3",
                          "--no-cpu-baseline", "--configs", "0,3", "--bodies", "96"], capture_output=True, text=True,
                         timeout=600, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [l for l in res.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, res.stdout[-2000:]
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline"):
        assert k in d, k
    assert d["n_gpus"] == 1 and d["steps"] == 12 and d["warmup"] >= 3 and d["higher_is_better"] is False
    assert d["unit"] == "ms/frame" and d["data"] == "synthetic" and d["dtype"] == "f32" and d["vs_baseline"] is None
    assert "workload" in d["config"] and "l2" in d["config"] and "model" not in d["config"]
    assert 0.05 < d["value"] < 5.0 and abs(d["value"] - d["ms_per_step"]) < 1e-9
    # 9 kernels per frame: two key kernels, one sort, two emits, one transform, two refits, one detection
    assert d["gpu_launches"] == 9 * d["steps"]
    e = d["e2e"]
    assert e["unit"] == "ms/frame" and e["value"] > d["value"] and e["h2d_bytes_per_step"] == 2 * 12 * d["config"]["verts_per_mesh"]
    assert e["d2h_bytes_per_step"] > 0
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert 0.0 < r["frac"] < 1.0 and (r["traffic"] is None or r["traffic"] > 0)
    c = d["clocks"]
    assert set(("sm_mhz", "sm_max_mhz", "reasons")) <= set(c)
    assert d["pairs"] > 0 and d["candidates"] >= d["pairs"]
    # per-kernel roofline table: every frame kernel with its SURVEY.md bytes, its time and its share; `roofline` names
    # the one with the largest share
    k = d["kernels"]
    assert len(k) == 5 and abs(sum(v["share_of_frame"] for v in k.values()) - 1.0) < 1e-6
    assert r["kernel"].startswith(max(k, key=lambda n: k[n]["ms"]))
    # one sub-record per requested BASELINE.json config (configs[1] is the line itself)
    recs = {c["config"]: c for c in d["configs"]}
    assert set(recs) == {"configs[0]", "configs[3]"}
    for c in recs.values():
        assert c["ms_per_frame"] > 0 and c["pairs"] > 0 and c["sharded_phase_ms"] > 0 and "roofline" in c


@pytest.mark.gpu
def test_reference_arm_line(ctx):
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-2000:]
    d = json.loads([l for l in res.stdout.splitlines() if l.startswith("{")][-1])
    assert d["impl"] == "reference" and d["unit"] == "ms/frame" and d["higher_is_better"] is False
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] == 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["value"] > 1.0
