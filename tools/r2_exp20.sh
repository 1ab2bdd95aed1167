#!/bin/bash
# round 2, experiment 20: INT stage, 4 vs 16 (8) vectors per lane, fresh process per measurement, alternating (bench regime)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
: > gpurun_out/r2_bench20.txt
for rep in 1 2 3; do
  for v in 4 16 8; do
    r=$(IWQ_SEG_VPL=$v python bench.py --steps 20 --warmup 3 --skip-e2e --skip-cpu --skip-configs 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value'],1), d['clocks']['sm_mhz'])")
    echo "$rep headline vpl=$v -> $r" >> gpurun_out/r2_bench20.txt
  done
done
for rep in 1 2; do
  for v in 4 16 8; do
    r=$(IWQ_SEG_VPL=$v python tools/sweep_r2.py --model llama-2-13b --layers 20 --iters 20 --rounds 1 --configs int4g64,int4g128,int4g256 2>/dev/null | python -c "
import sys,json
o=[]
for l in sys.stdin:
    d=json.loads(l)
    if 'knobs' in d: o.append(d['config']+'='+str(round(d['GBps'])))
print(' '.join(o))")
    echo "$rep 13b-half vpl=$v -> $r" >> gpurun_out/r2_bench20.txt
  done
done
cat gpurun_out/r2_bench20.txt
