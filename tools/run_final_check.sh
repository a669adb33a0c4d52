#!/bin/bash
# parity tests, smoke, headline bench (plain), then ONE ncu capture of a short k_render launch of the same scene / resolution
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/final_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/final_tests.log
tail -3 gpurun_out/final_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1; echo "smoke rc=$?"
timeout 600 python bench.py > gpurun_out/final_bench_default.json 2> gpurun_out/final_bench_default.err; rc=$?; echo "bench rc=$rc"
tail -1 gpurun_out/final_bench_default.json | cut -c1-200
timeout 300 python bench.py --workload cornell_default --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/final_c1.json
python -c "import json; d=json.load(open('gpurun_out/final_c1.json')); print('C1', d['value'], d['e2e']['value'])"
if [ $rc -eq 0 ]; then
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_render --launch-skip 1 -c 1 -f \
      -o gpurun_out/prof_bench_k_render_r01k python bench.py --spp 32 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/final_ncu.log 2>&1
  echo "ncu rc=$?"
fi
