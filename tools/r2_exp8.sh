#!/bin/bash
# round 2, experiment 8: the seg4 headline kernel against what tools/tmabench.cu found (fewer bytes in flight, contiguous tile runs
# per CTA, evict_last on the bulk loads / stores)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
K='seg_map=0;seg_ctas=2;seg_stages=2;seg_map=1;seg_hint=2;seg_hint=18;seg_map=1,seg_hint=2;seg_map=1,seg_hint=18;seg_map=1,seg_hint=2,seg_ctas=2;seg_map=1,seg_hint=18,seg_ctas=2;seg_map=1,seg_hint=2,seg_stages=2;seg_map=1,seg_hint=2,seg_stages=4,seg_ctas=2;seg_map=1,seg_hint=18,seg_stages=4,seg_ctas=2;seg_map=1,seg_hint=2,seg_stages=4;seg_map=1,seg_hint=2,seg_stages=5,seg_ctas=2'
python tools/sweep_r2.py --model llama-2-70b --layers 10 --configs int4g128 --iters 8 --rounds 3 --knobs "$K" > gpurun_out/r2_sweep_seg8.jsonl 2> gpurun_out/r2_sweep8.err
cut -c1-200 gpurun_out/r2_sweep_seg8.jsonl; tail -5 gpurun_out/r2_sweep8.err
