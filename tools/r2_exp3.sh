#!/bin/bash
# round 2, experiment 3: direct minifloat table vs replicated step table vs step table; bench.py with configs
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "fp or fuzz or variants or smoke or wrapper or model" > gpurun_out/r2_tests3.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2_tests3.log
K='fp_mode=0;fp_mode=1,fp_rep=3;fp_mode=2;fp_mode=2,fp_stages=4;fp_mode=2,fp_rep_kb=27;fp_mode=2,seg_ctas=2,fp_stages=4'
python tools/sweep_r2.py --model llama-2-13b --layers 10 --iters 5 --rounds 3 --configs fp4e2m1g128,fp8e4m3g128,fp6e3m2g128,fp6e2m3g128,fp4e1m2g128,fp4e2m1g128asym,int4g128 --knobs "$K" > gpurun_out/r2_sweep_fp3.jsonl 2> gpurun_out/r2_sweep_fp3.err
python tools/sweep_r2.py --model llama-2-13b --layers 10 --iters 5 --rounds 3 --dtype bf16 --configs fp4e2m1g128,fp8e4m3g128 --knobs "$K" > gpurun_out/r2_sweep_fp3_bf16.jsonl 2>> gpurun_out/r2_sweep_fp3.err
python bench.py --steps 20 --warmup 3 > gpurun_out/r2_bench3.json 2> gpurun_out/r2_bench3.err; echo "bench rc $?" >> gpurun_out/r2_bench3.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2_bench3_ref.json 2>> gpurun_out/r2_bench3.err
tail -3 gpurun_out/r2_tests3.log; cut -c1-150 gpurun_out/r2_sweep_fp3.jsonl gpurun_out/r2_sweep_fp3_bf16.jsonl; tail -5 gpurun_out/r2_sweep_fp3.err; tail -20 gpurun_out/r2_bench3.err; cut -c1-1500 gpurun_out/r2_bench3.json; cut -c1-800 gpurun_out/r2_bench3_ref.json
