#!/bin/bash
for lv in 0 8 9; do
  echo "== OIBVH_DENSE_SEED_LEVEL=$lv"
  OIBVH_DENSE_SEED_LEVEL=$lv timeout 200 python tools/stage_bench.py --frames 40 --check 2>&1 | tail -7
done
timeout 600 python -m pytest tests/test_gpu_collide.py tests/test_gpu_scenes.py tests/test_gpu_large.py -m gpu -x -q 2>&1 | tail -4
