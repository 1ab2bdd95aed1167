#!/bin/bash
# round 2, experiment 24: row_tma_kernel (per-channel rows that are not 2048 / 4096 elements) -- consumer warps x CTAs per SM x ring budget
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
K='row_warps=0;row_warps=4,row_ctas=3;row_warps=8,row_ctas=2;row_warps=8,row_ctas=2,row_smem_kb=100;row_warps=16,row_ctas=1,row_smem_kb=200;row_warps=16,row_ctas=1,row_smem_kb=110;row_warps=12,row_ctas=1,row_smem_kb=200;row_warps=8,row_ctas=1,row_smem_kb=200;row_warps=8,row_ctas=3;row_warps=6,row_ctas=2,row_smem_kb=100'
(
python tools/sweep_r2.py --shape 4096,11008,64 --configs int4pc --iters 10 --rounds 2 --knobs "$K" | sed 's/^{/{"shape": "4096x11008", /'
python tools/sweep_r2.py --shape 5120,13824,40 --configs int4pc --iters 10 --rounds 2 --knobs "$K" | sed 's/^{/{"shape": "5120x13824", /'
python tools/sweep_r2.py --shape 8192,28672,12 --configs int4pc --iters 10 --rounds 2 --knobs "$K" | sed 's/^{/{"shape": "8192x28672", /'
) > gpurun_out/r2_sweep_row24.jsonl 2> gpurun_out/r2_sweep24.err
python - <<'PY'
import json
for l in open('gpurun_out/r2_sweep_row24.jsonl'):
    d=json.loads(l)
    if 'knobs' in d: print(d['shape'], d['knobs'], round(d.get('GBps',0)), d.get('GBps_all'), d.get('error',''))
PY
tail -n 3 gpurun_out/r2_sweep24.err
