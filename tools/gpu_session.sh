#!/bin/bash
# One GPU session: sanitizer on a small detection, the whole GPU test suite, then the stage / frame timings.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader
echo "== racecheck + memcheck, small detections"
timeout 600 compute-sanitizer --tool racecheck python -m pytest "tests/test_gpu_collide.py::test_golden_two_spheres" "tests/test_gpu_collide.py::test_golden_three_bodies" -m gpu -x -q 2>&1 | tail -4
timeout 600 compute-sanitizer --tool memcheck python -m pytest "tests/test_gpu_collide.py::test_golden_two_spheres" "tests/test_gpu_scenes.py" -m gpu -x -q 2>&1 | tail -4
echo "== pytest -m gpu"
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
echo "== stage bench"
timeout 300 python tools/stage_bench.py --frames 30 --check 2>&1 | tail -7
timeout 300 python tools/many_body.py 2>&1 | tail -6
timeout 300 python tools/big_scene.py 2>&1 | tail -4
