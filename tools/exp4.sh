#!/bin/bash
( timeout 900 python -m pytest tests/test_gpu_tree.py tests/test_gpu_large.py tests/test_gpu_scenes.py -m gpu -x -q ) 2>&1 | tail -4
timeout 300 python tools/refit_bench.py 2>&1 | tail -8
OIBVH_B200_LIB=$PWD/oibvh_b200/variants/lib_PROFILE.so timeout 100 python tools/emit_profile.py 2>&1 | tail -9
timeout 300 python tools/stage_bench.py --frames 40 --check 2>&1 | tail -7
