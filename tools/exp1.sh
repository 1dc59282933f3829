#!/bin/bash
# GPU box: validate HEAD (tests + bench), then gather diagnostics for the refit kernel
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
timeout 600 python bench.py > gpurun_out/bench.log 2>&1
echo "== baseline" > gpurun_out/refit_variants.log
timeout 300 python tools/refit_bench.py >> gpurun_out/refit_variants.log 2>&1
for v in GATHER_LINEAR GATHER_CONSEC; do
  echo "== $v" >> gpurun_out/refit_variants.log
  OIBVH_B200_LIB=$PWD/oibvh_b200/variants/lib_$v.so timeout 300 python tools/refit_bench.py >> gpurun_out/refit_variants.log 2>&1
done
M=l1tex__data_pipe_lsu_wavefronts_mem_lg.sum,l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld_lookup_hit.sum,l1tex__throughput.avg.pct_of_peak_sustained_elapsed,l1tex__data_pipe_lsu_wavefronts.sum,l1tex__lsu_writeback_active.sum,sm__cycles_elapsed.max,gpu__time_duration.sum,lts__t_sectors_op_read.sum,lts__throughput.avg.pct_of_peak_sustained_elapsed,smsp__warps_issue_stalled_long_scoreboard_per_warp_active.pct,smsp__warps_issue_stalled_lg_throttle_per_warp_active.pct
timeout 600 ncu --metrics $M --clock-control none -k regex:"tree_emit|morton_hist" -c 12 --csv --log-file gpurun_out/l1tex_emit.csv python tools/stage_bench.py --frames 2 > gpurun_out/ncu_l1.log 2>&1
tail -3 gpurun_out/pytest_gpu.log; cat gpurun_out/refit_variants.log; tail -c 600 gpurun_out/bench.log
