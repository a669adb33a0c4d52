#!/bin/bash
# round 2, step h: streaming queue stores, L2 persisting window for the BVH nodes (experiment), if-if schedule (descend_min 32)
mkdir -p gpurun_out
S=synthetic4k:8; M=mesh_1080p:128
tools/r02_exp.sh r02h "$S:" "$S:l2_persist=1" "$S:wf_descend_min=32" "$S:wf_descend_min=24" "$S:wf_descend_min=12" "$M:" "$M:l2_persist=1" "$S:bvh_leaf_max=1" "$S:bvh_leaf_max=3"
