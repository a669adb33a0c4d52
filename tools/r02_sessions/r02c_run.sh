#!/bin/bash
# round 2, step c: GPU suite (release + bounds-checking build) on the two-ended wavefront queue, then A/B on the BVH workloads
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/r02c_tests.log 2>&1; echo "tests rc=$?" | tee -a gpurun_out/r02c_tests.log
tail -4 gpurun_out/r02c_tests.log
PTB_LIBRARY=$PWD/path_tracer_rust_b200/libptb_check.so timeout 1500 python -m pytest tests -m gpu -q -k "not fullsize_synthetic and not statistics" > gpurun_out/r02c_tests_check.log 2>&1; echo "check-build tests rc=$?" | tee -a gpurun_out/r02c_tests_check.log
tail -3 gpurun_out/r02c_tests_check.log
S=synthetic4k:8; M=mesh_1080p:128
tools/r02_exp.sh r02c \
  "$S:" "$S:wf_trace_variant=1,wf_leaf_min=8" "$S:bvh_top_levels=4" "$S:wavefront_paths=16777216" "$S:wavefront_paths=33554432" "$S:wavefront_paths=67108864" \
  "$S:wavefront_paths=33554432,bvh_top_levels=4" "$S:wavefront_paths=33554432,wf_refill=4" "$S:wavefront_paths=33554432,wf_refill=12" "$S:wavefront_paths=33554432,bvh_leaf_max=4" \
  "$M:" "$M:wf_trace_variant=1" "$M:wavefront_paths=33554432" "$M:wavefront_paths=2097152" \
  "cornell_default:100:" "cornell_default:100:integrator=1" "single_sphere_1080p:256:" "three_spheres_1080p:256:"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r02c_launches_mesh.csv \
   python tools/profile_render.py mesh 1920 1080 8 2 > gpurun_out/r02c_ncu_mesh.log 2>&1
