#!/bin/bash
# round 2, experiment 14: col_tma_kernel box width / stages / CTAs / L2 hint
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "dim1 or quant_dim" > gpurun_out/r2_tests14.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2_tests14.log
K='col_box=0;col_box=0,col_ctas=2;col_box=0,col_hint=2;col_box=0,col_ctas=2,col_hint=2;col_box=64;col_box=64,col_ctas=2;col_box=64,col_stages=2;col_box=64,col_hint=2;col_box=128;col_box=128,col_ctas=1;col_box=128,col_stages=2;col_box=128,col_hint=2;col_box=64,col_ctas=3,col_hint=2'
python tools/sweep_r2.py --fresh --model llama-2-70b --layers 3 --iters 4 --rounds 2 --configs int4g128_d1,int4g128sym_d1,int4g64_d1,fp4e2m1g128_d1 --knobs "$K" > gpurun_out/r2_sweep_col14.jsonl 2> gpurun_out/r2_sweep14.err
tail -4 gpurun_out/r2_tests14.log
python - <<'PY'
import json
for l in open('gpurun_out/r2_sweep_col14.jsonl'):
    d=json.loads(l)
    if 'knobs' in d: print(d['config'], d['knobs'], round(d.get('GBps',0)), d.get('GBps_all'), d.get('error',''))
PY
tail -5 gpurun_out/r2_sweep14.err
