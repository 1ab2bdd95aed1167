#!/bin/bash
cd "$(dirname "$0")/.."
python tools/microbench.py --configs int4g128,fp4e2m1g128,fp4e2m1g128asym,fp8e4m3g128 --layers 3
IWQ_SEG_STAGES=4 python tools/microbench.py --configs fp4e2m1g128,fp8e4m3g128 --layers 3
IWQ_SEG_STAGES=2 python tools/microbench.py --configs fp4e2m1g128,fp8e4m3g128 --layers 3
IWQ_SEG_STAGES=5 IWQ_SEG_CTAS=2 python tools/microbench.py --configs fp4e2m1g128,fp8e4m3g128 --layers 3
python tools/microbench.py --dtype bf16 --configs int4g128,fp4e2m1g128,fp8e4m3g128 --layers 3
