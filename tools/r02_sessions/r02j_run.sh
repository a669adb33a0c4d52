#!/bin/bash
# round 2, step j: the driver's own commands at reduced step counts (bench with all extra workloads, reference arm), then the ncu evidence
# of the bench command itself: launch list + one full capture of the headline kernel (a 258-spp launch as bench.py issues them)
mkdir -p gpurun_out
( time timeout 1500 python bench.py --gpus 1 --steps 5 --warmup 2 > gpurun_out/r02j_bench_n1.json 2> gpurun_out/r02j_bench_n1.err ) 2> gpurun_out/r02j_bench_n1.time; echo "bench rc=$?"; cat gpurun_out/r02j_bench_n1.time | tail -3
( time timeout 600 python bench.py --impl reference --gpus 1 --steps 2 --warmup 1 > gpurun_out/r02j_bench_ref.json 2> gpurun_out/r02j_bench_ref.err ) 2> gpurun_out/r02j_bench_ref.time; echo "ref rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02j_launches_bench.csv \
   python bench.py --spp 258 --steps 1 --warmup 1 --extras mesh_1080p,cornell_default --no-cpu-baseline > gpurun_out/r02j_ncu_list.log 2>&1; echo "list rc=$?"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'k_render' --launch-skip 1 -c 1 -f \
   -o gpurun_out/prof_bench_k_render_r02j python bench.py --spp 258 --steps 1 --warmup 1 --extras none --no-cpu-baseline > gpurun_out/r02j_ncu_full.log 2>&1; echo "full rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_render' -c 1 -f \
   -o gpurun_out/prof_sp_k_render_r02j python tools/profile_render.py cornell 450 300 100 1 > gpurun_out/r02j_ncu_sp.log 2>&1; echo "sp rc=$?"
