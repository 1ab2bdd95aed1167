#!/bin/bash
cd "$(dirname "$0")/.."
for st in 2 3 4 5; do for c in 0 2; do
  IWQ_SEG_STAGES=$st IWQ_SEG_CTAS=$c python tools/microbench.py --configs int4g128,int4g128sym,fp4e2m1g128 --layers 3 "$@"
done; done
python tools/microbench.py --configs int8pc,int4pt,int4g32,int4g64,int4g256 --layers 3
python tools/microbench.py --dtype bf16 --configs int4g128,int8pc,fp4e2m1g128 --layers 3
