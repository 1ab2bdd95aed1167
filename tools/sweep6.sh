#!/bin/bash
cd "$(dirname "$0")/.."
for rv in 1 0; do
for w in 1 2 3 4; do for kb in 28 36 48; do
  IWQ_ROW_VARIANT=$rv IWQ_ROW_WARPS=$w IWQ_ROW_SMEM_KB=$kb python tools/microbench.py --model llama-2-7b --layers 8 --configs int8pc
done; done
for w in 2 4 6; do for kb in 36 56 72; do
  IWQ_ROW_VARIANT=$rv IWQ_ROW_WARPS=$w IWQ_ROW_SMEM_KB=$kb python tools/microbench.py --model llama-2-70b --layers 2 --configs int8pc
done; done
done
