#!/bin/bash
# round 2, step w (N GPUs): final build, driver-style bench with all extras (+ cabi_multi at N > 1)
N=${1:-1}
mkdir -p gpurun_out
if [ "$N" = "1" ]; then
  ( time timeout 1500 python bench.py --gpus 1 --steps 3 --warmup 2 > gpurun_out/r02w_bench_n1.json 2> gpurun_out/r02w_bench_n1.err ) 2> gpurun_out/r02w_bench_n1.time; echo "bench rc=$?"; tail -3 gpurun_out/r02w_bench_n1.time
else
  ( time timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus $N --steps 3 --warmup 2 > gpurun_out/r02w_bench_n$N.json 2> gpurun_out/r02w_bench_n$N.err ) 2> gpurun_out/r02w_bench_n$N.time; echo "bench n$N rc=$?"; tail -3 gpurun_out/r02w_bench_n$N.time
fi
