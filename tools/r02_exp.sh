#!/bin/bash
# usage: tools/r02_exp.sh TAG "workload:spp:opts" ...   (opts = comma-separated k=v backend options)
# one JSON line per experiment into gpurun_out/exp_<TAG>.jsonl
TAG=$1; shift
mkdir -p gpurun_out
: > gpurun_out/exp_$TAG.jsonl
for e in "$@"; do
  IFS=':' read -r wl spp opts <<< "$e"
  args=""
  IFS=',' read -ra kv <<< "$opts"
  for o in "${kv[@]}"; do [ -n "$o" ] && args="$args --opt $o"; done
  line=$(timeout 600 python bench.py --workload $wl --spp $spp --steps 2 --warmup 1 --no-cpu-baseline $args 2>>gpurun_out/exp_$TAG.err | tail -1)
  echo "$line" >> gpurun_out/exp_$TAG.jsonl
  echo "$line" | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); print('$e', round(d['value'],1), 'Mpaths/s', 'ms', round(d['ms_per_step'],1), d['roofline']['tests_per_segment'])
except Exception as ex: print('$e', 'FAILED', ex)"
done
