#!/bin/bash
# N-GPU session: the two-device / two-process tests of the peer-mapped pair list, then bench.py under torchrun.
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader
echo "== mgpu tests"
timeout 600 python -m pytest tests/test_gpu_mgpu.py -m gpu -x -q 2>&1 | tail -6
echo "== bench N=$N"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
   bench.py --gpus $N --steps 100 --warmup 5 > gpurun_out/bench_n$N.log 2> gpurun_out/bench_n$N.err
echo "rc=$?"; tail -c 1500 gpurun_out/bench_n$N.err
