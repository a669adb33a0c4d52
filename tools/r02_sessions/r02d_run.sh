#!/bin/bash
# round 2, step d: adaptive trace chunk + explicit stack addressing; paths-in-flight sweep on the BVH workloads; tile the small configs
mkdir -p gpurun_out
S=synthetic4k:8; M=mesh_1080p:128
tools/r02_exp.sh r02d \
  "$S:wavefront_paths=8388608" "$S:wavefront_paths=16777216" "$S:wavefront_paths=33554432" "$S:wavefront_paths=67108864" \
  "$S:wavefront_paths=33554432,bvh_top_levels=0" "$S:wavefront_paths=33554432,bvh_top_levels=5,wf_trace_threads=512" "$S:wavefront_paths=33554432,wf_descend_min=8" "$S:wavefront_paths=33554432,wf_descend_min=16" \
  "$M:wavefront_paths=8388608" "$M:wavefront_paths=16777216" "$M:wavefront_paths=33554432" "$M:wavefront_paths=67108864" "$M:wavefront_paths=33554432,bvh_top_levels=0" \
  "cornell_default:100:" "cornell_default:100:wavefront_paths=33554432"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r02d_launches_mesh.csv \
   python tools/profile_render.py mesh 1920 1080 8 2 > gpurun_out/r02d_ncu_mesh.log 2>&1
timeout 900 python bench.py --steps 1 --warmup 1 --extras cornell_default,three_spheres_1080p,mesh_1080p > gpurun_out/r02d_bench_extras.json 2> gpurun_out/r02d_bench_extras.err; echo "bench rc=$?"
