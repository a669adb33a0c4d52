#!/bin/bash
# round 2, step y: eight-wide kernel: how many nodes in shared memory (the north star's "shared-memory-staged BVH top levels" on the new kernel)
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02y_smoke.log 2>&1; echo "smoke rc=$?"; grep -c "bit-exact" gpurun_out/r02y_smoke.log
S=synthetic4k:8
tools/r02_exp.sh r02y "$S:wf_top8_nodes=0" "$S:wf_top8_nodes=73" "$S:wf_top8_nodes=256" "$S:wf_top8_nodes=512" "$S:wf_top8_nodes=585,wf_trace_threads=1024" "$S:wf_top8_nodes=1024,wf_trace_threads=1024"
