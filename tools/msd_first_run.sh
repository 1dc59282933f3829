#!/bin/bash
# First GPU session of the experimental MSD sort (oibvh_b200/csrc/sort_msd.cu). Everything runs under
# `timeout` so that a wedged cooperative kernel cannot hold the box: memcheck on a small tree first, then the gated
# parity test, then build timings with and without OIBVH_SORT_MSD=1.
mkdir -p gpurun_out
export OIBVH_TEST_EXPERIMENTAL=1
timeout 300 compute-sanitizer --tool memcheck python -m pytest \
  "tests/test_gpu_tree.py::test_experimental_msd_sort_builds_the_same_tree[5000]" -m gpu -x -q > gpurun_out/msd_memcheck.log 2>&1
tail -5 gpurun_out/msd_memcheck.log
timeout 300 python -m pytest tests/test_gpu_tree.py -k experimental_msd -m gpu -x -q 2>&1 | tail -8
for v in 0 1; do
  echo "== OIBVH_SORT_MSD=$v"
  OIBVH_SORT_MSD=$v timeout 200 python tools/refit_bench.py 2>&1 | grep build
  OIBVH_SORT_MSD=$v timeout 200 python tools/stage_bench.py --frames 30 --check 2>&1 | tail -7   # two-tree launch inside the frame graph
done
