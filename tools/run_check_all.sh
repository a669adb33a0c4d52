#!/bin/bash
# One GPU call: parity tests, smoke, the headline bench line and the other BASELINE configs without their CPU legs.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/check_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/check_tests.log
tail -3 gpurun_out/check_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/check_smoke.log 2>&1; echo "smoke rc=$?"
timeout 600 python bench.py > gpurun_out/check_bench_default.json 2> gpurun_out/check_bench_default.err; echo "bench rc=$?"
tail -1 gpurun_out/check_bench_default.json | cut -c1-400
out=gpurun_out/check_matrix.jsonl
: > $out
timeout 300 python bench.py --workload cornell_default --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 >> $out
timeout 300 python bench.py --workload single_sphere_1080p --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 >> $out
timeout 300 python bench.py --workload three_spheres_1080p --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 >> $out
timeout 300 python bench.py --workload mesh_1080p --steps 2 --warmup 1 --no-cpu-baseline 2>/dev/null | tail -1 >> $out
timeout 400 python bench.py --workload synthetic4k --steps 1 --warmup 1 --no-cpu-baseline 2>/dev/null | tail -1 >> $out
python - <<'PY'
import json
for l in open('gpurun_out/check_matrix.jsonl'):
    d = json.loads(l)
    print(f"{d['config']['workload'][:60]:60s} {d['value']:10.1f} Mpaths/s  {d['mray_segments_per_s']:10.1f} Mseg/s  e2e {d['e2e']['value']:10.1f}  ms/step {d['ms_per_step']:.1f}")
PY
