#!/bin/bash
# round 2, experiment 15: col_wide_kernel (CTA-wide ring of 128 x 256 tensor-map tiles) vs col_tma_kernel
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "dim1 or quant_dim or fuzz or variants or approx or bfp" > gpurun_out/r2_tests15.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2_tests15.log
K='col_mode=3'
python tools/sweep_r2.py --fresh --model llama-2-70b --layers 3 --iters 4 --rounds 2 --configs int4g128_d1,int4g128sym_d1,int4g64_d1,int4g32_d1,fp4e2m1g128_d1,fp4e2m1g64_d1,fp8e4m3g128_d1,fp8e4m3g64_d1,fp6e3m2g64_d1,bfp4g128_d1 --knobs "$K" > gpurun_out/r2_sweep_col15.jsonl 2> gpurun_out/r2_sweep15.err
tail -6 gpurun_out/r2_tests15.log
python - <<'PY'
import json
for l in open('gpurun_out/r2_sweep_col15.jsonl'):
    d=json.loads(l)
    if 'knobs' in d: print(d['config'], d['knobs'], round(d.get('GBps',0)), d.get('GBps_all'), d.get('error',''))
PY
tail -5 gpurun_out/r2_sweep15.err
