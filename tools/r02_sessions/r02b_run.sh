#!/bin/bash
# round 2, step b: GPU suite on the rebuilt wavefront integrator, then A/B of the trace-kernel variants on the two BVH workloads
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/r02b_tests.log 2>&1; echo "tests rc=$?" | tee -a gpurun_out/r02b_tests.log
tail -3 gpurun_out/r02b_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02b_smoke.log 2>&1; echo "smoke rc=$?"
S=synthetic4k:8; M=mesh_1080p:128
tools/r02_exp.sh r02b \
  "$S:" "$S:wf_trace_variant=1" "$S:wf_trace_variant=1,wf_leaf_min=8" "$S:wf_trace_variant=1,wf_leaf_min=24" \
  "$S:wf_trace_variant=1,wf_descend_min=20" "$S:wf_trace_variant=1,wf_descend_min=6" \
  "$S:bvh_top_levels=4" "$S:wf_trace_variant=1,bvh_top_levels=4" "$S:wf_trace_variant=1,bvh_top_levels=5,wf_trace_threads=1024" \
  "$S:wf_trace_variant=1,wavefront_paths=33554432" "$S:wavefront_paths=33554432" \
  "$M:" "$M:wf_trace_variant=1" "$M:wf_trace_variant=1,bvh_top_levels=4" "$M:integrator=1" \
  "cornell_default:100:" "three_spheres_1080p:256:"
