#!/bin/bash
# round 2, experiment 1: gate tests + per-tensor slab ordering / L2 hints / headline hints / read-only ceiling
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem --format=csv > gpurun_out/gpu.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests1.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2_tests1.log
tools/membench 8192 8192 24 l2 > gpurun_out/r2_membench_l2.log 2>&1
python tools/sweep_r2.py --model opt-6.7b --configs int4pt --iters 5 \
  --knobs 'pt_slab_mb=0,pt_reverse=0;pt_slab_mb=0,pt_reverse=1;pt_slab_mb=40;pt_slab_mb=40,pt_hint=1;pt_slab_mb=40,pt_hint=5;pt_slab_mb=70;pt_slab_mb=140;pt_slab_mb=40,seg_ctas=2,seg_stages=4;pt_slab_mb=0,pt_hint=1;pt_slab_mb=0,pt_hint=5' > gpurun_out/r2_sweep_pt.jsonl 2> gpurun_out/r2_sweep_pt.err
python tools/sweep_r2.py --model opt-6.7b --shards 8 --configs int4pt --iters 10 \
  --knobs 'pt_slab_mb=0,pt_reverse=0;pt_slab_mb=0,pt_reverse=1;pt_slab_mb=40;pt_slab_mb=24;pt_slab_mb=40,pt_hint=5' > gpurun_out/r2_sweep_pt_shard8.jsonl 2>> gpurun_out/r2_sweep_pt.err
python tools/sweep_r2.py --model llama-2-70b --layers 10 --configs int4g128 --iters 10 \
  --knobs 'seg_hint=0;seg_hint=1;seg_hint=4;seg_hint=5;seg_hint=2;seg_hint=0' > gpurun_out/r2_sweep_hint.jsonl 2>> gpurun_out/r2_sweep_pt.err
tail -3 gpurun_out/r2_tests1.log; cat gpurun_out/r2_sweep_pt.jsonl gpurun_out/r2_sweep_pt_shard8.jsonl gpurun_out/r2_sweep_hint.jsonl; tail -5 gpurun_out/r2_membench_l2.log; tail -5 gpurun_out/r2_sweep_pt.err
