#!/bin/bash
# round 2, experiment 9: seg4 headline kernel with 8 KB tiles (4 consumer warps, 5-6 CTAs/SM) and 32 KB tiles (16 consumer warps)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
P=iron_weight_only_quant_b200
run() { IWQ_LIB=$PWD/$P/libiwq_b200_$1.so python tools/sweep_r2.py --model llama-2-70b --layers 10 --configs int4g128 --iters 8 --rounds 2 --knobs "$2" | sed "s/^{/{\"lib\": \"$1\", /"; }
(
python tools/sweep_r2.py --model llama-2-70b --layers 10 --configs int4g128 --iters 8 --rounds 2 --knobs "seg_hint=0;seg_hint=2"
run cw4 'seg_stages=3;seg_stages=2;seg_stages=4;seg_stages=3,seg_ctas=5;seg_stages=3,seg_ctas=4;seg_stages=2,seg_ctas=5;seg_stages=3,seg_hint=2;seg_stages=3,seg_hint=2,seg_map=1;seg_stages=4,seg_ctas=4;seg_stages=4,seg_ctas=5'
run cw4b 'seg_stages=3;seg_stages=2;seg_stages=4;seg_stages=3,seg_ctas=4;seg_stages=3,seg_hint=2;seg_stages=4,seg_hint=2'
run cw16 'seg_stages=3;seg_stages=2;seg_stages=4;seg_stages=3,seg_ctas=1;seg_stages=2,seg_ctas=1;seg_stages=4,seg_ctas=1;seg_stages=3,seg_ctas=1,seg_hint=2;seg_stages=3,seg_ctas=1,seg_hint=2,seg_map=1;seg_stages=5,seg_ctas=1'
) > gpurun_out/r2_sweep_seg9.jsonl 2> gpurun_out/r2_sweep9.err
cut -c1-175 gpurun_out/r2_sweep_seg9.jsonl; tail -5 gpurun_out/r2_sweep9.err
