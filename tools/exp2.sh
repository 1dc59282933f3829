#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
tail -4 gpurun_out/pytest_gpu.log
timeout 300 python tools/refit_bench.py 2>&1 | tail -8
timeout 300 python tools/stage_bench.py --frames 30 2>&1 | tail -5
