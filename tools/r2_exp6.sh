#!/bin/bash
# round 2, experiment 6: quant_dim = 1 through 2-D tensor maps (col_tma_kernel) vs the cp.async kernel; FP8 stage variants
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "dim1 or quant_dim or approx or fuzz or variants or bfp" > gpurun_out/r2_tests6.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2_tests6.log
K='col_mode=0;col_mode=2;col_mode=2,col_stages=1;col_mode=2,col_ctas=2;col_mode=2,col_stages=1,col_ctas=5'
python tools/sweep_r2.py --fresh --model llama-2-70b --layers 3 --iters 4 --rounds 2 --configs int4g128_d1,int4g128sym_d1,int4g64_d1,int4g256_d1,fp4e2m1g128_d1,fp4e2m1g128_d1_approx --knobs "$K" > gpurun_out/r2_sweep_col6.jsonl 2> gpurun_out/r2_sweep_col6.err
python tools/sweep_r2.py --fresh --model llama-2-13b --layers 8 --iters 4 --rounds 2 --configs fp8e4m3g128 --knobs "fp_mode=2;fp_mode=2,fp_stages=2;fp_mode=2,fp_stages=4;fp_mode=2,fp_stages=5" > gpurun_out/r2_sweep_fp6.jsonl 2>> gpurun_out/r2_sweep_col6.err
tail -15 gpurun_out/r2_tests6.log; cut -c1-170 gpurun_out/r2_sweep_col6.jsonl gpurun_out/r2_sweep_fp6.jsonl; tail -5 gpurun_out/r2_sweep_col6.err
