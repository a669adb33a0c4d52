#!/bin/bash
# ray reordering between bounces (wf_sort): parity tests first, then the two BVH workloads with the sort off / octant-major / cell-major
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/sort_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/sort_tests.log
tail -3 gpurun_out/sort_tests.log
out=gpurun_out/sort_experiment.jsonl
: > $out
for m in 0 1 2; do
  timeout 300 python bench.py --workload synthetic4k --spp 64 --steps 1 --warmup 1 --no-cpu-baseline --opt wf_sort=$m 2>/dev/null | tail -1 >> $out
  timeout 300 python bench.py --workload mesh_1080p --spp 128 --steps 2 --warmup 1 --no-cpu-baseline --opt wf_sort=$m 2>/dev/null | tail -1 >> $out
done
python - <<'PY'
import json
for l in open('gpurun_out/sort_experiment.jsonl'):
    d = json.loads(l)
    print(f"{d['config']['workload'][:44]:44s} {str(d['config']['backend_options']):16s} {d['value']:9.1f} Mpaths/s {d['mray_segments_per_s']:9.1f} Mseg/s  ms/step {d['ms_per_step']:.1f} launches {d['gpu_launches']}")
PY
