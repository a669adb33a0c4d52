#!/bin/bash
# round 2, step i: binned-SAH topology for small primitive sets (mesh.json), parity first
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -k "bvh or fuzz or golden or lockstep or wavefront_equals or smoke" > gpurun_out/r02i_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r02i_tests.log
M=mesh_1080p:128
tools/r02_exp.sh r02i "$M:" "$M:bvh_sah_max_prims=0" "$M:bvh_leaf_max=4" "$M:bvh_leaf_max=1" "$M:bvh_leaf_max=3" "$M:integrator=1" "$M:bvh_top_levels=0"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r02i_launches_mesh.csv \
   python tools/profile_render.py mesh 1920 1080 16 2 > gpurun_out/r02i_ncu_mesh.log 2>&1
