#!/bin/bash
# round 2, experiment 5: minifloat lookup variants on FRESH data every pass + gate tests for the new pieces
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests5.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2_tests5.log
K='fp_mode=0;fp_mode=1,fp_rep=3;fp_mode=1,fp_rep=4;fp_mode=2;fp_mode=1,fp_rep=3,fp_stages=4'
python tools/sweep_r2.py --fresh --model llama-2-13b --layers 8 --iters 4 --rounds 2 --configs fp4e2m1g128,fp8e4m3g128,fp6e3m2g128,fp6e2m3g128,fp4e1m2g128,int4g128 --knobs "$K" > gpurun_out/r2_sweep_fp5.jsonl 2> gpurun_out/r2_sweep_fp5.err
python tools/sweep_r2.py --fresh --model llama-2-13b --layers 8 --iters 4 --rounds 2 --dtype bf16 --configs fp4e2m1g128,fp8e4m3g128 --knobs "$K" > gpurun_out/r2_sweep_fp5_bf16.jsonl 2>> gpurun_out/r2_sweep_fp5.err
python tools/sweep_r2.py --model llama-2-13b --layers 8 --iters 4 --rounds 2 --configs fp4e2m1g128,fp8e4m3g128 --knobs "fp_mode=0;fp_mode=1,fp_rep=3" > gpurun_out/r2_sweep_fp5_stale.jsonl 2>> gpurun_out/r2_sweep_fp5.err
tail -15 gpurun_out/r2_tests5.log; cut -c1-150 gpurun_out/r2_sweep_fp5.jsonl gpurun_out/r2_sweep_fp5_bf16.jsonl gpurun_out/r2_sweep_fp5_stale.jsonl; tail -5 gpurun_out/r2_sweep_fp5.err
