#!/bin/bash
# One GPU session: the whole GPU test suite, then the stage / frame timings of the bench scene.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader
echo "== pytest -m gpu"
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
echo "== stage bench"
timeout 300 python tools/stage_bench.py --frames 30 --check 2>&1 | tail -7
timeout 300 python tools/refit_bench.py 2>&1 | grep -E "build|refit"
