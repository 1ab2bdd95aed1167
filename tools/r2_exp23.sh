#!/bin/bash
# round 2, experiment 23: per-channel rows of 4096 elements with 16 vectors per lane (one warp per row, no named barrier)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
IWQ_SEG_VPL=16 timeout 600 python -m pytest tests -m gpu -x -q -k "per_channel or int_matches or fuzz or wrappers or full_size" > gpurun_out/r2_tests23.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2_tests23.log
: > gpurun_out/r2_bench23.txt
for rep in 1 2 3; do
  for v in 4 16 8; do
    r=$(IWQ_SEG_VPL=$v python tools/sweep_r2.py --model llama-2-7b --iters 20 --rounds 1 --configs int8pc,int4pc 2>/dev/null | python -c "
import sys,json
o=[]
for l in sys.stdin:
    d=json.loads(l)
    if 'knobs' in d: o.append(d['config']+'='+str(round(d['GBps'])))
print(' '.join(o))")
    echo "$rep 7b per-channel vpl=$v -> $r" >> gpurun_out/r2_bench23.txt
  done
done
tail -n 3 gpurun_out/r2_tests23.log; cat gpurun_out/r2_bench23.txt
