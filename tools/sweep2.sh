#!/bin/bash
cd "$(dirname "$0")/.."
for st in 2 3 4; do for c in 0 3; do
  IWQ_SEG_STAGES=$st IWQ_SEG_CTAS=$c python tools/microbench.py --configs int4g128,fp4e2m1g128 --layers 3 "$@"
done; done
