#!/bin/bash
# BASELINE configs C1-C3 on one GPU (C4 = default bench.py, C5 = --workload synthetic4k take minutes each)
out=gpurun_out/matrix_small_n1.jsonl
: > $out
python bench.py --workload cornell_default --steps 5 --warmup 3 2>/dev/null | tail -1 >> $out
python bench.py --workload single_sphere_1080p --steps 5 --warmup 3 2>/dev/null | tail -1 >> $out
python bench.py --workload three_spheres_1080p --steps 5 --warmup 3 2>/dev/null | tail -1 >> $out
python bench.py --workload mesh_1080p --steps 2 --warmup 1 2>/dev/null | tail -1 >> $out
python - <<'PY'
import json
for l in open('gpurun_out/matrix_small_n1.jsonl'):
    d = json.loads(l)
    cb = d.get('cpu_baseline') or {}
    print(f"{d['config']['workload'][:50]:50s} {d['value']:10.1f} Mpaths/s  {d['mray_segments_per_s']:10.1f} Mseg/s  e2e {d['e2e']['value']:10.1f}  "
          f"frac {d['roofline']['frac']:.3f}  cpu {cb.get('value'):.3f} ({cb.get('cores')} cores)")
PY
