#!/bin/bash
# round 2, experiment 2: replicated minifloat step table (bank-conflict-free lookups)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "fp or fuzz or variants or smoke or wrapper or model" > gpurun_out/r2_tests2.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2_tests2.log
K='fp_rep=0;fp_rep=4;fp_rep=3;fp_rep=4,fp_stages=4;fp_rep=4,fp_stages=2;fp_rep=4,fp_rep_kb=64;fp_rep=4,fp_rep_kb=64,fp_stages=4;fp_rep=4,fp_rep_kb=64,fp_stages=2'
python tools/sweep_r2.py --model llama-2-13b --layers 10 --iters 5 --rounds 3 --configs fp4e2m1g128,fp8e4m3g128,fp6e3m2g128,fp6e2m3g128,fp4e1m2g128 --knobs "$K" > gpurun_out/r2_sweep_fp.jsonl 2> gpurun_out/r2_sweep_fp.err
python tools/sweep_r2.py --model llama-2-13b --layers 10 --iters 5 --rounds 3 --dtype bf16 --configs fp4e2m1g128,fp8e4m3g128 --knobs "$K" > gpurun_out/r2_sweep_fp_bf16.jsonl 2>> gpurun_out/r2_sweep_fp.err
tail -3 gpurun_out/r2_tests2.log; cut -c1-200 gpurun_out/r2_sweep_fp.jsonl gpurun_out/r2_sweep_fp_bf16.jsonl; tail -5 gpurun_out/r2_sweep_fp.err
