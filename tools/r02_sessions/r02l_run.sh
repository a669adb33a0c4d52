#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/profile_render.py three-spheres 1920 1080 256 2 > gpurun_out/r02l_plain.log 2>&1; tail -1 gpurun_out/r02l_plain.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_render' -c 1 -f \
   -o gpurun_out/prof_k_render_three_spheres_r02l python tools/profile_render.py three-spheres 1920 1080 256 1 > gpurun_out/r02l_ncu.log 2>&1; echo "rc=$?"
