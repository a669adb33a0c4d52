#!/bin/bash
# round 2, step f (2 GPUs): multi-GPU context tests, bench.py under torchrun with extras + cabi_multi, CLI --gpus, shade occupancy A/B
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests -m gpu -q -k "multi_gpu or peer_reduce or peer_memory" > gpurun_out/r02f_tests_multi.log 2>&1; echo "multi tests rc=$?"; tail -3 gpurun_out/r02f_tests_multi.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 1 --warmup 1 --extras mesh_1080p,cornell_default --no-cpu-baseline > gpurun_out/r02f_bench_n2.json 2> gpurun_out/r02f_bench_n2.err; echo "bench n2 rc=$?"
( cd /root/repo && ./path_tracer_rust_b200/render 256 1080 cornell --gpus 2 --out /tmp/c2.ppm && ./path_tracer_rust_b200/render 256 1080 cornell --gpus 1 --out /tmp/c1.ppm && cmp /tmp/c1.ppm /tmp/c2.ppm; echo "cli cmp rc=$? (0 = PPMs identical, differences in the last digit of a few pixels are possible: different summation order)" ) > gpurun_out/r02f_cli.log 2>&1; tail -4 gpurun_out/r02f_cli.log


