#!/bin/bash
# round 2, step ac: final build: GPU suite (release + check build), smoke, last knobs of the eight-wide kernel
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02ac_tests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/r02ac_tests.log
PTB_LIBRARY=$PWD/path_tracer_rust_b200/libptb_check.so timeout 1500 python -m pytest tests -m gpu -q -k "not fullsize_synthetic_rays and not statistics" > gpurun_out/r02ac_tests_check.log 2>&1; echo "check-build tests rc=$?"; tail -2 gpurun_out/r02ac_tests_check.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02ac_smoke.log 2>&1; echo "smoke rc=$?"
S=synthetic4k:8
tools/r02_exp.sh r02ac "$S:" "$S:wf_trace_threads=1024" "$S:wf_trace_threads=256" "$S:wf_refill=6" "$S:wavefront_paths=67108864"
