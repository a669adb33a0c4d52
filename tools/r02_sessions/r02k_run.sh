#!/bin/bash
# round 2, step k (N GPUs): bench.py under torchrun exactly as the driver launches it (reduced steps), with extras and cabi_multi
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L | wc -l
( time timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 3 --warmup 2 > gpurun_out/r02k_bench_n$N.json 2> gpurun_out/r02k_bench_n$N.err ) 2> gpurun_out/r02k_bench_n$N.time; echo "bench n$N rc=$?"; tail -3 gpurun_out/r02k_bench_n$N.time
timeout 600 python -m pytest tests -m gpu -q -k "multi_gpu or peer_reduce" > gpurun_out/r02k_tests_multi_n$N.log 2>&1; echo "multi tests rc=$?"; tail -2 gpurun_out/r02k_tests_multi_n$N.log
