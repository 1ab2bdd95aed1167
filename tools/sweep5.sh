#!/bin/bash
cd "$(dirname "$0")/.."
export IWQ_ROW_VARIANT=1
for w in 2 4 8; do for kb in 24 36 64; do
  IWQ_ROW_WARPS=$w IWQ_ROW_SMEM_KB=$kb python tools/microbench.py --model llama-2-7b --layers 8 --configs int8pc
done; done
for w in 8 16; do for kb in 72 110; do
  IWQ_ROW_WARPS=$w IWQ_ROW_SMEM_KB=$kb python tools/microbench.py --model llama-2-70b --layers 2 --configs int8pc
done; done
