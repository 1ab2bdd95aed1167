#!/bin/bash
# round 2, experiment 10: seg4 headline kernel, tile size (consumer warps per CTA) x ring depth around the 32 KB x 4 stages x 1 CTA point
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
P=iron_weight_only_quant_b200
run() { IWQ_LIB=$PWD/$P/libiwq_b200_$1.so python tools/sweep_r2.py --model llama-2-70b --layers 10 --configs int4g128 --iters 8 --rounds 2 --knobs "$2" | sed "s/^{/{\"lib\": \"$1\", /"; }
(
python tools/sweep_r2.py --model llama-2-70b --layers 10 --configs int4g128 --iters 8 --rounds 2 --knobs "seg_hint=0;seg_hint=2"
run cw16 'seg_stages=4;seg_stages=4,seg_hint=2;seg_stages=4,seg_map=1;seg_stages=4,seg_hint=2,seg_map=1;seg_stages=4,seg_hint=18;seg_stages=4,seg_hint=1;seg_stages=6;seg_stages=4'
run cw12 'seg_stages=3;seg_stages=4;seg_stages=4,seg_ctas=1;seg_stages=5,seg_ctas=1;seg_stages=6,seg_ctas=1;seg_stages=3,seg_hint=2'
run cw20 'seg_stages=3;seg_stages=4;seg_stages=5;seg_stages=4,seg_hint=2'
run cw24 'seg_stages=3;seg_stages=4;seg_stages=4,seg_hint=2;seg_stages=2'
) > gpurun_out/r2_sweep_seg10.jsonl 2> gpurun_out/r2_sweep10.err
cut -c1-175 gpurun_out/r2_sweep_seg10.jsonl; tail -5 gpurun_out/r2_sweep10.err
