#!/bin/bash
# round 2, step v: eight-wide kernel with a parked primitive group: parity, then thresholds
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "fuzz or fullsize_synthetic_lockstep or bvh_equals or golden" > gpurun_out/r02v_tests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/r02v_tests.log
S=synthetic4k:8
tools/r02_exp.sh r02v "$S:" "$S:wf_descend_min=16" "$S:wf_descend_min=20" "$S:wf_descend_min=28" "$S:wf_descend_min=12" "mesh_1080p:128:bvh_wide=1"
