#!/bin/bash
# round 2, step n: persistent shade kernel that keeps continuations which cannot reach the BVH; parity, then A/B
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x -k "not fullsize_synthetic_rays and not statistics" > gpurun_out/r02n_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r02n_tests.log
M=mesh_1080p:128; S=synthetic4k:8
tools/r02_exp.sh r02n_lb3 "$M:" "$S:" "cornell_default:100:integrator=2" "cornell4k:16:integrator=2"
PTB_LIBRARY=$PWD/path_tracer_rust_b200/libptb_alt2.so tools/r02_exp.sh r02n_lb2 "$M:" "$S:" "cornell_default:100:integrator=2"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r02n_launches_mesh.csv \
   python tools/profile_render.py mesh 1920 1080 16 2 > gpurun_out/r02n_ncu_mesh.log 2>&1
