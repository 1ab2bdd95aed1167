#!/bin/bash
# Knob sweep for tools/microbench.py (one process per setting; prints one JSON line each).
cd "$(dirname "$0")/.."
for v in 0 1 2; do
  IWQ_VARIANT=$v python tools/microbench.py --configs int4g128,int4g128sym,int4g32,fp4e2m1g128 "$@"
done
python tools/microbench.py --configs int8pc,int4pc,int4pt,int4g64,int4g256,fp8e4m3g128,fp4e2m1g128asym "$@"
python tools/microbench.py --dtype bf16 --configs int4g128,int8pc,fp4e2m1g128 "$@"
for st in 3 4 5; do
  IWQ_SEG_STAGES=$st python tools/microbench.py --configs int4g128 "$@"
done
IWQ_SEG_STAGES=4 IWQ_SEG_CTAS=2 python tools/microbench.py --configs int4g128 "$@"
IWQ_SEG_STAGES=8 IWQ_SEG_CTAS=1 python tools/microbench.py --configs int4g128 "$@"
