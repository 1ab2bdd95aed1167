#!/bin/bash
# Knob sweep for tools/microbench.py (one process per setting; prints one JSON line each).
cd "$(dirname "$0")/.."
for v in 0 1; do
  IWQ_VARIANT=$v python tools/microbench.py --configs int4g128,int4g128sym,fp4e2m1g128 "$@"
done
for st in 3 4 8; do
  IWQ_SEG_STAGES=$st python tools/microbench.py --configs int4g128 "$@"
done
for c in 1 2; do
  IWQ_SEG_CTAS=$c IWQ_SEG_STAGES=6 python tools/microbench.py --configs int4g128 "$@"
done
IWQ_SEG_STAGES=3 IWQ_SEG_CTAS=4 python tools/microbench.py --configs int4g128 "$@"
