#!/bin/bash
# round 2, experiment 22: the round-1 fused per-tensor kernel on the new ring shape (48 KB tiles, one CTA per SM)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "fused" > gpurun_out/r2_tests22.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2_tests22.log
K='fused_batch_mb=16,fused_lookahead=2;fused_batch_mb=32,fused_lookahead=1;fused_batch_mb=32,fused_lookahead=2;fused_batch_mb=48,fused_lookahead=1;fused_batch_mb=8,fused_lookahead=3;fused_batch_mb=24,fused_lookahead=2'
python tools/sweep_r2.py --model opt-6.7b --configs int4pt_fused,int4pt --iters 5 --rounds 2 --knobs "$K" > gpurun_out/r2_sweep_pt22.jsonl 2> gpurun_out/r2_sweep22.err
python tools/sweep_r2.py --model opt-6.7b --shards 8 --configs int4pt_fused,int4pt --iters 10 --rounds 2 --knobs "$K" > gpurun_out/r2_sweep_pt22_s8.jsonl 2>> gpurun_out/r2_sweep22.err
tail -n 3 gpurun_out/r2_tests22.log
python - <<'PY'
import json
for f in ['pt22','pt22_s8']:
    for l in open(f'gpurun_out/r2_sweep_{f}.jsonl'):
        d=json.loads(l)
        if 'knobs' in d: print(f, d['config'], d['knobs'], round(d.get('GBps',0)), d.get('GBps_all'), d.get('error',''))
PY
tail -n 3 gpurun_out/r2_sweep22.err
