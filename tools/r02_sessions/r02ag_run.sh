#!/bin/bash
# round 2, step ag: FINAL build: GPU suite (release + check build), smoke, driver-style bench with all extras
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02ag_tests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/r02ag_tests.log
PTB_LIBRARY=$PWD/path_tracer_rust_b200/libptb_check.so timeout 1500 python -m pytest tests -m gpu -q -k "not fullsize_synthetic_rays and not statistics" > gpurun_out/r02ag_tests_check.log 2>&1; echo "check-build tests rc=$?"; tail -2 gpurun_out/r02ag_tests_check.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02ag_smoke.log 2>&1; echo "smoke rc=$?"
( time timeout 1500 python bench.py --gpus 1 --steps 3 --warmup 2 > gpurun_out/r02ag_bench_n1.json 2> gpurun_out/r02ag_bench_n1.err ) 2> gpurun_out/r02ag_bench_n1.time; echo "bench rc=$?"; tail -3 gpurun_out/r02ag_bench_n1.time
