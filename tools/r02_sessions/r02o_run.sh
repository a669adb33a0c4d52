#!/bin/bash
# round 2, step o (2 GPUs): smoke() with the multi-GPU part, bench.py --gpus 2 WITHOUT torchrun (one process, ptb_create_multi), time-budget skip path
mkdir -p gpurun_out
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02o_smoke.log 2>&1; echo "smoke rc=$?"; tail -4 gpurun_out/r02o_smoke.log
timeout 900 python bench.py --gpus 2 --steps 2 --warmup 1 --extras mesh_1080p,cornell_default --no-cpu-baseline > gpurun_out/r02o_bench_single_process_n2.json 2> gpurun_out/r02o_bench_single_process_n2.err; echo "single-process bench rc=$?"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29547 bench.py --gpus 2 --steps 1 --warmup 1 --spp 256 --extras mesh_1080p,synthetic4k --time-budget 25 --no-cpu-baseline > gpurun_out/r02o_bench_budget.json 2> gpurun_out/r02o_bench_budget.err; echo "budget bench rc=$?"
