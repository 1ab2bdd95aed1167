#!/bin/bash
# round 2, experiment 18: col_wide_kernel with the direct minifloat table and 128-column tiles (the reference's own scripts run
# --quant_dim 1 --approximate with FP4 / FP6 / FP8 at group 64 and 128)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "dim1 or quant_dim or fuzz or variants or approx or bfp" > gpurun_out/r2_tests18.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2_tests18.log
K='fp_mode=2;fp_mode=0'
python tools/sweep_r2.py --fresh --model llama-2-70b --layers 3 --iters 4 --rounds 2 --configs fp4e2m1g128_d1,fp4e2m1g128_d1_approx,fp8e4m3g128_d1,fp6e3m2g128_d1,fp4e2m1g64_d1,fp8e4m3g64_d1,fp6e3m2g64_d1,int4g128_d1 --knobs "$K" > gpurun_out/r2_sweep_col18.jsonl 2> gpurun_out/r2_sweep18.err
python tools/sweep_r2.py --fresh --model llama-2-70b --layers 3 --iters 4 --rounds 2 --dtype bf16 --configs fp8e4m3g128_d1,fp6e3m2g64_d1 --knobs "$K" > gpurun_out/r2_sweep_col18_bf16.jsonl 2>> gpurun_out/r2_sweep18.err
tail -n 4 gpurun_out/r2_tests18.log
python - <<'PY'
import json
for f in ['col18','col18_bf16']:
    for l in open(f'gpurun_out/r2_sweep_{f}.jsonl'):
        d=json.loads(l)
        if 'knobs' in d: print(f, d['config'], d['knobs'], round(d.get('GBps',0)), d.get('GBps_all'), d.get('error',''))
PY
tail -n 5 gpurun_out/r2_sweep18.err
