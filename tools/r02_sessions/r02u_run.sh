#!/bin/bash
# round 2, step u: eight-wide BVH as the default for large sets: GPU suite (release + check build), numbers, ncu capture of k_wf_trace8
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --durations=5 > gpurun_out/r02u_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r02u_tests.log
PTB_LIBRARY=$PWD/path_tracer_rust_b200/libptb_check.so timeout 1500 python -m pytest tests -m gpu -q -k "not fullsize_synthetic_rays and not statistics" > gpurun_out/r02u_tests_check.log 2>&1; echo "check-build tests rc=$?"; tail -2 gpurun_out/r02u_tests_check.log
tools/r02_exp.sh r02u "synthetic4k:8:" "synthetic4k:8:bvh_wide=0" "mesh_1080p:128:" "synthetic4k:32:"
timeout 600 python tools/profile_render.py synthetic 1920 1080 4 2 > gpurun_out/r02u_plain.log 2>&1; tail -1 gpurun_out/r02u_plain.log
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'k_wf_trace8' --launch-skip 3 -c 1 -f \
   -o gpurun_out/prof_wf_r02u_syn python tools/profile_render.py synthetic 1920 1080 4 1 > gpurun_out/r02u_ncu_full.log 2>&1; echo "full rc=$?"
