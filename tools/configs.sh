#!/bin/bash
# BASELINE.json configs 2-4 on one GPU (kernel throughput on the named model shapes; parity is covered by tests/)
cd "$(dirname "$0")/.."
python tools/microbench.py --model llama-2-7b --layers 0 --iters 5 --configs int8pc,int4pc,int4pc_sym
python tools/microbench.py --model llama-2-13b --layers 0 --iters 5 --configs int4g32,int4g64,int4g128,int4g256,fp4e2m1g128,fp4e1m2g128,fp6e3m2g128,fp6e2m3g128,fp8e4m3g128
python tools/microbench.py --model opt-6.7b --layers 0 --iters 5 --configs int4pt,int4pt_sym
python tools/microbench.py --model llama-2-13b --layers 0 --iters 5 --dtype bf16 --configs int4g128,int8pc,fp4e2m1g128,fp8e4m3g128
