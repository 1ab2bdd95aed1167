#!/bin/bash
# round 2, experiment 17: vectors per lane measured the way the bench measures (fresh process, 3 warm-up + 20 timed steps), alternating
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
: > gpurun_out/r2_bench17.txt
for rep in 1 2 3; do
  for v in "IWQ_SEG_VPL=4" "IWQ_SEG_VPL=8" "IWQ_SEG_VPL=8 IWQ_SEG_STAGES=3" "IWQ_SEG_VPL=16"; do
    r=$(env $v python bench.py --steps 20 --warmup 3 --skip-e2e --skip-cpu --skip-configs 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value'],1), d['clocks']['sm_mhz'], d['clocks']['reasons'])")
    echo "$rep $v -> $r" >> gpurun_out/r2_bench17.txt
  done
done
cat gpurun_out/r2_bench17.txt
