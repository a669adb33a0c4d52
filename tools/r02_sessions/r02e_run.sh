#!/bin/bash
# round 2, step e: ncu evidence for the rebuilt wavefront kernels on the synthetic scene (same frame as profiles/r01g: 1080p x 4 spp, bounce 3)
mkdir -p gpurun_out
timeout 600 python tools/profile_render.py synthetic 1920 1080 4 2 > gpurun_out/r02e_plain.log 2>&1; echo "plain rc=$?"; tail -2 gpurun_out/r02e_plain.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02e_launches_syn.csv \
   python tools/profile_render.py synthetic 1920 1080 4 1 > gpurun_out/r02e_ncu_list.log 2>&1; echo "list rc=$?"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'k_wf_(trace|shade)' --launch-skip 6 -c 2 -f \
   -o gpurun_out/prof_wf_r02e_syn python tools/profile_render.py synthetic 1920 1080 4 1 > gpurun_out/r02e_ncu_full.log 2>&1; echo "full rc=$?"
S=synthetic4k:8
tools/r02_exp.sh r02e "$S:" "$S:wf_descend_min=16" "$S:wf_descend_min=20" "$S:wf_trace_threads=1024" "$S:wf_trace_threads=256,bvh_top_levels=4" "$S:wf_refill=12,wf_descend_min=16" "mesh_1080p:128:" "mesh_1080p:128:wf_descend_min=16"
