#!/bin/bash
mkdir -p gpurun_out
tools/r02_exp.sh r02m "cornell4k:64:" "three_spheres_1080p:256:" "single_sphere_1080p:256:" "cornell_default:100:" "cornell4k:64:"
