#!/bin/bash
# round 2, experiment 7 (after the container was re-created): full gate + bench both arms + the sweeps whose results were lost
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem --format=csv > gpurun_out/gpu.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests7.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2_tests7.log
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2_bench7_ref.json 2> gpurun_out/r2_bench7.err
python bench.py --steps 20 --warmup 3 > gpurun_out/r2_bench7.json 2>> gpurun_out/r2_bench7.err; echo "bench rc $?" >> gpurun_out/r2_bench7.err
K='col_mode=0;col_mode=2;col_mode=2,col_stages=1;col_mode=2,col_ctas=2;col_mode=2,col_stages=1,col_ctas=5'
python tools/sweep_r2.py --fresh --model llama-2-70b --layers 3 --iters 4 --rounds 2 --configs int4g128_d1,int4g128sym_d1,int4g64_d1,int4g256_d1,fp4e2m1g128_d1,fp4e2m1g128_d1_approx --knobs "$K" > gpurun_out/r2_sweep_col7.jsonl 2> gpurun_out/r2_sweep7.err
K='fp_mode=0;fp_mode=1,fp_rep=3;fp_mode=1,fp_rep=4;fp_mode=2;fp_mode=2,fp_stages=4;fp_mode=2,fp_stages=2'
python tools/sweep_r2.py --fresh --model llama-2-13b --layers 8 --iters 4 --rounds 2 --configs fp4e2m1g128,fp8e4m3g128,fp6e3m2g128,fp6e2m3g128,fp4e1m2g128,int4g128 --knobs "$K" > gpurun_out/r2_sweep_fp7.jsonl 2>> gpurun_out/r2_sweep7.err
python tools/sweep_r2.py --fresh --model llama-2-13b --layers 8 --iters 4 --rounds 2 --dtype bf16 --configs fp4e2m1g128,fp8e4m3g128 --knobs "$K" > gpurun_out/r2_sweep_fp7_bf16.jsonl 2>> gpurun_out/r2_sweep7.err
python tools/sweep_r2.py --model opt-6.7b --configs int4pt --iters 5 \
  --knobs 'pt_slab_mb=0,pt_reverse=0;pt_slab_mb=0,pt_reverse=1;pt_slab_mb=40;pt_slab_mb=40,pt_hint=1;pt_slab_mb=40,pt_hint=5;pt_slab_mb=70;pt_slab_mb=140;pt_slab_mb=24' > gpurun_out/r2_sweep_pt7.jsonl 2>> gpurun_out/r2_sweep7.err
python tools/sweep_r2.py --model llama-2-70b --layers 10 --configs int4g128 --iters 10 \
  --knobs 'seg_hint=0;seg_hint=1;seg_hint=4;seg_hint=5;seg_hint=2;seg_stages=4,seg_ctas=2;seg_hint=0' > gpurun_out/r2_sweep_hint7.jsonl 2>> gpurun_out/r2_sweep7.err
tail -15 gpurun_out/r2_tests7.log; tail -5 gpurun_out/r2_bench7.err; cut -c1-170 gpurun_out/r2_sweep_col7.jsonl gpurun_out/r2_sweep_fp7.jsonl gpurun_out/r2_sweep_fp7_bf16.jsonl gpurun_out/r2_sweep_pt7.jsonl gpurun_out/r2_sweep_hint7.jsonl; tail -5 gpurun_out/r2_sweep7.err
