#!/bin/bash
cd "$(dirname "$0")/.."
python tools/microbench.py --model llama-2-7b --layers 8 --configs int8pc,int4pc_sym
IWQ_ROW_VARIANT=1 python tools/microbench.py --model llama-2-7b --layers 8 --configs int8pc,int4pc_sym
python tools/microbench.py --model llama-2-70b --layers 2 --configs int8pc,int4pc_sym
IWQ_ROW_VARIANT=1 python tools/microbench.py --model llama-2-70b --layers 2 --configs int8pc
python tools/microbench.py --model opt-125m --layers 0 --configs int8pc,int4g128
python tools/microbench.py --model llama-2-7b --layers 8 --dtype bf16 --configs int8pc,int4pc
for w in 2 8; do IWQ_ROW_WARPS=$w python tools/microbench.py --model llama-2-7b --layers 8 --configs int8pc; done
