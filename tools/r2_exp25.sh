#!/bin/bash
# round 2, experiment 25: per-channel on Llama-2-70B shapes (8192- and 28672-wide rows): row kernel vs seg4 with 8 warps per row
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
K='wide_seg4=1;wide_seg4=2;wide_seg4=1,row_warps=8;wide_seg4=1,row_warps=8,row_ctas=3;wide_seg4=1'
python tools/sweep_r2.py --model llama-2-70b --layers 4 --configs int4pc,int8pc --iters 10 --rounds 2 --knobs "$K" > gpurun_out/r2_sweep_pc25.jsonl 2> gpurun_out/r2_sweep25.err
python tools/sweep_r2.py --shape 8192,8192,24 --configs int4pc --iters 10 --rounds 2 --knobs "$K;wide_seg4=1,row_warps=4,row_ctas=4;wide_seg4=1,row_warps=2,row_ctas=6" >> gpurun_out/r2_sweep_pc25.jsonl 2>> gpurun_out/r2_sweep25.err
python - <<'PY'
import json
for l in open('gpurun_out/r2_sweep_pc25.jsonl'):
    d=json.loads(l)
    if 'knobs' in d: print(d['config'], d['knobs'], round(d.get('GBps',0)), d.get('GBps_all'), d.get('error',''))
    elif 'model' in d: print(d)
PY
tail -n 3 gpurun_out/r2_sweep25.err
