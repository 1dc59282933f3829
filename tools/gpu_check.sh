#!/bin/bash
# GPU box regression pass used during development: parity tests, refit / frame / big-scene timings
( timeout 900 python -m pytest tests -m gpu -x -q ) 2>&1 | tail -3
timeout 300 python tools/refit_bench.py 2>&1 | grep refit
timeout 300 python tools/stage_bench.py --frames 40 2>&1 | tail -4
timeout 200 python tools/big_scene.py 2>&1 | tail -6
