#!/bin/bash
# Round-end evidence run on the GPU box: parity tests, clean bench line, ncu launch list of the bench command, one
# ncu --set full capture of the frame's kernels, sanitizer over the kernels added this round.
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q ) > gpurun_out/pytest_gpu.log 2>&1
tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 50 --warmup 5 > gpurun_out/bench_final.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv \
  --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"tree_emit|morton_hist|coop_sort|collide_kernel|transform_kernel" \
  --launch-skip 12 -c 14 -o gpurun_out/frame_full python tools/profile_frame.py --frames 3 > gpurun_out/ncu_full.log 2>&1
ncu -i gpurun_out/frame_full.ncu-rep --page raw --csv > gpurun_out/frame_full_raw.csv 2>/dev/null
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_collide.py tests/test_gpu_headless.py -m gpu -q -x \
  > gpurun_out/sanitizer_memcheck_new.log 2>&1
tail -4 gpurun_out/sanitizer_memcheck_new.log
timeout 300 compute-sanitizer --tool racecheck python __graft_entry__.py smoke > gpurun_out/sanitizer_racecheck_smoke.log 2>&1
tail -3 gpurun_out/sanitizer_racecheck_smoke.log
ls -la gpurun_out | head -30
