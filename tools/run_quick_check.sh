#!/bin/bash
# parity tests, smoke, headline bench line, C1 (no profiler)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/quick_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/quick_tests.log
tail -3 gpurun_out/quick_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/quick_smoke.log 2>&1; echo "smoke rc=$?"
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/quick_bench_default.json 2> gpurun_out/quick_bench_default.err; echo "bench rc=$?"
tail -1 gpurun_out/quick_bench_default.json | cut -c1-200
timeout 300 python bench.py --workload cornell_default --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/quick_c1.json
python -c "import json; d=json.load(open('gpurun_out/quick_c1.json')); print('C1', d['value'], d['e2e']['value'])"
timeout 200 python bench.py --workload three_spheres_1080p --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/quick_c2.json
python -c "import json; d=json.load(open('gpurun_out/quick_c2.json')); print('C2 three-spheres', d['value'], d['e2e']['value'])"
