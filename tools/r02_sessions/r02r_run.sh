#!/bin/bash
# round 2, step r: compressed eight-wide BVH in the trace kernel (bvh_wide=1): parity, then A/B on the BVH workloads
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fuzz.py -m gpu -q -x > gpurun_out/r02r_tests_fuzz.log 2>&1; echo "fuzz rc=$?"; tail -3 gpurun_out/r02r_tests_fuzz.log
S=synthetic4k:8; M=mesh_1080p:128
tools/r02_exp.sh r02r "$M:bvh_wide=1" "$M:" "$S:bvh_wide=1" "$S:" "$S:bvh_wide=1,wf_descend_min=24" "$S:bvh_wide=1,wf_trace_threads=256"
