#!/bin/bash
# Runs every BASELINE.json configuration on one GPU and collects the bench lines (gpurun_out/matrix_n1.jsonl).
set -u
out=gpurun_out/matrix_n1.jsonl
: > $out
python bench.py --workload cornell_default --steps 5 --warmup 3 2>/dev/null | tail -1 >> $out
python bench.py --workload single_sphere_1080p --steps 5 --warmup 3 2>/dev/null | tail -1 >> $out
python bench.py --workload three_spheres_1080p --steps 5 --warmup 3 2>/dev/null | tail -1 >> $out
python bench.py --workload mesh_1080p --steps 2 --warmup 1 2>/dev/null | tail -1 >> $out
python bench.py --workload synthetic4k --steps 1 --warmup 1 2>/dev/null | tail -1 >> $out
python - <<'PY'
import json
for l in open('gpurun_out/matrix_n1.jsonl'):
    d = json.loads(l)
    cb = d.get('cpu_baseline') or {}
    print(f"{d['config']['workload'][:60]:60s} {d['value']:10.1f} Mpaths/s  {d['mray_segments_per_s']:10.1f} Mseg/s  e2e {d['e2e']['value']:10.1f}  "
          f"frac {d['roofline']['frac']}  cpu {cb.get('value')} ({cb.get('cores')} cores)  ms/step {d['ms_per_step']:.1f}")
PY
