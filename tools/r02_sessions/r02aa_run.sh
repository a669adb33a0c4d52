#!/bin/bash
# round 2, step aa: eight-wide collapse by the SAH cost recurrence vs greedy
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "fuzz or fullsize_synthetic_lockstep" > gpurun_out/r02aa_tests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/r02aa_tests.log
S=synthetic4k:8
tools/r02_exp.sh r02aa "$S:bvh_wide_sah=1" "$S:bvh_wide_sah=0" "mesh_1080p:128:bvh_wide=1,bvh_wide_sah=1"
