#!/bin/bash
# Round-end evidence run on the GPU box: parity tests, clean bench line, ncu launch list of the bench command, one
# ncu --set full capture of the frame's kernels, sanitizers over the kernels written this round.
R=${1:-r02}
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/${R}_pytest_gpu.log 2>&1
tail -3 gpurun_out/${R}_pytest_gpu.log
timeout 900 python bench.py > gpurun_out/${R}_bench_line.json 2> gpurun_out/${R}_bench.err
timeout 600 python bench.py --impl reference --steps 10 --warmup 1 > gpurun_out/${R}_bench_line_reference_arm.json 2>> gpurun_out/${R}_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv \
  --log-file gpurun_out/${R}_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --configs 1 > gpurun_out/ncu_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"tree_emit|morton_hist|lsd_sort|collide_kernel|transform_kernel" \
  --launch-skip 16 -c 9 -f -o gpurun_out/${R}_frame_full python tools/profile_frame.py --frames 3 > gpurun_out/ncu_full.log 2>&1
ncu -i gpurun_out/${R}_frame_full.ncu-rep --page raw --csv > gpurun_out/${R}_frame_full_raw.csv 2>/dev/null
timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_collide.py tests/test_gpu_mgpu.py tests/test_gpu_tree.py -m gpu -q -x \
  > gpurun_out/${R}_sanitizer_memcheck.log 2>&1
tail -4 gpurun_out/${R}_sanitizer_memcheck.log
timeout 600 compute-sanitizer --tool racecheck python -m pytest "tests/test_gpu_collide.py::test_golden_three_bodies" "tests/test_gpu_collide.py::test_self_collision_matches_brute_force" "tests/test_gpu_tree.py::test_build_sizes[34816]" -m gpu -q -x \
  > gpurun_out/${R}_sanitizer_racecheck.log 2>&1
tail -3 gpurun_out/${R}_sanitizer_racecheck.log
ls -la gpurun_out | tail -12
