#!/bin/bash
# round 2, step g: sample-parallel megakernel for small frames + shade occupancy A/B; full GPU suite on release and check builds
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --durations=5 > gpurun_out/r02g_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r02g_tests.log
PTB_LIBRARY=$PWD/path_tracer_rust_b200/libptb_check.so timeout 1500 python -m pytest tests -m gpu -q -k "not fullsize_synthetic and not statistics" > gpurun_out/r02g_tests_check.log 2>&1; echo "check-build tests rc=$?"; tail -2 gpurun_out/r02g_tests_check.log
tools/r02_exp.sh r02g "cornell_default:100:" "cornell_default:100:integrator=2" "cornell_default:100:integrator=1" "single_sphere_1080p:256:" "three_spheres_1080p:256:" "three_spheres_1080p:256:integrator=3" \
   "mesh_1080p:128:" "synthetic4k:8:" "cornell4k:64:"
PTB_LIBRARY=$PWD/path_tracer_rust_b200/libptb_alt4.so tools/r02_exp.sh r02g_lb4 "mesh_1080p:128:" "synthetic4k:8:"
