#!/bin/bash
# round 2, step s: eight-wide BVH with one / two primitives per leaf child
mkdir -p gpurun_out
S=synthetic4k:8; M=mesh_1080p:128
PTB_LIBRARY=$PWD/path_tracer_rust_b200/libptb_leaf1.so tools/r02_exp.sh r02s_leaf1 "$S:bvh_wide=1" "$M:bvh_wide=1"
PTB_LIBRARY=$PWD/path_tracer_rust_b200/libptb_leaf2.so tools/r02_exp.sh r02s_leaf2 "$S:bvh_wide=1" "$M:bvh_wide=1"
