#!/bin/bash
# round 2, step t: eight-wide BVH (one primitive per leaf child): thresholds / CTA size
mkdir -p gpurun_out
S=synthetic4k:8
tools/r02_exp.sh r02t "$S:bvh_wide=1" "$S:bvh_wide=1,wf_descend_min=8" "$S:bvh_wide=1,wf_descend_min=12" "$S:bvh_wide=1,wf_descend_min=24" "$S:bvh_wide=1,wf_trace_threads=1024" "$S:bvh_wide=1,wf_refill=4" "$S:bvh_wide=1,wf_refill=12" "$S:bvh_wide=1,wavefront_paths=67108864"
