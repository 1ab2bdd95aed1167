#!/bin/bash
# round 2, experiment 19: col_double_kernel with 64-column tiles (second sweep from L2) vs 256-column tiles
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "double or fuzz or approx" > gpurun_out/r2_tests19.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2_tests19.log
K='col_dbl_tile=256'
python tools/sweep_r2.py --fresh --model llama-2-70b --layers 3 --iters 4 --rounds 2 --configs fp4e2m1g128_d1_double,fp8e4m3g64_d1_double --knobs "$K" > gpurun_out/r2_sweep_col19.jsonl 2> gpurun_out/r2_sweep19.err
tail -n 4 gpurun_out/r2_tests19.log
python - <<'PY'
import json
for l in open('gpurun_out/r2_sweep_col19.jsonl'):
    d=json.loads(l)
    if 'knobs' in d: print(d['config'], d['knobs'], round(d.get('GBps',0)), d.get('GBps_all'), d.get('error',''))
PY
tail -n 5 gpurun_out/r2_sweep19.err
