#!/bin/bash
# round-2 baseline: BVH workloads as shipped at the end of round 1 (numbers + launch lists + full captures of the wavefront kernels)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv > gpurun_out/r02a_smi.log 2>&1
timeout 300 python bench.py --workload mesh_1080p --spp 128 --steps 2 --warmup 1 --no-cpu-baseline 2>gpurun_out/r02a_mesh.err | tail -1 > gpurun_out/r02a_mesh.json
timeout 600 python bench.py --workload synthetic4k --spp 8 --steps 2 --warmup 1 --no-cpu-baseline 2>gpurun_out/r02a_syn.err | tail -1 > gpurun_out/r02a_syn.json
python - <<'PY'
import json
for n in ("mesh","syn"):
    try:
        d=json.load(open(f"gpurun_out/r02a_{n}.json")); print(n, d["value"], d["e2e"]["value"], d["roofline"]["tests_per_segment"])
    except Exception as e: print(n, "failed", e)
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r02a_launches_mesh.csv \
   python tools/profile_render.py mesh 1920 1080 8 2 > gpurun_out/r02a_ncu_mesh.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_wf_(trace|shade)' --launch-skip 4 -c 2 -f \
   -o gpurun_out/prof_wf_r02a_mesh python tools/profile_render.py mesh 1920 1080 4 1 > gpurun_out/r02a_ncu_mesh_full.log 2>&1
echo "ncu rc=$?"
