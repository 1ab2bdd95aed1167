#!/bin/bash
# round 2, experiment 12: 48 KB tiles / 24 consumer warps as the default -- full gate, bench, minifloat + per-tensor + per-channel sweeps
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests12.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2_tests12.log
python bench.py --steps 20 --warmup 3 > gpurun_out/r2_bench12.json 2> gpurun_out/r2_bench12.err; echo "bench rc $?" >> gpurun_out/r2_bench12.err
K='fp_mode=2;fp_mode=2,fp_stages=3;fp_mode=0;fp_mode=1,fp_rep=3'
python tools/sweep_r2.py --fresh --model llama-2-13b --layers 8 --iters 4 --rounds 3 --configs fp4e2m1g128,fp8e4m3g128,fp6e3m2g128,fp6e2m3g128,fp4e1m2g128 --knobs "$K" > gpurun_out/r2_sweep_fp12.jsonl 2> gpurun_out/r2_sweep12.err
python tools/sweep_r2.py --fresh --model llama-2-13b --layers 8 --iters 4 --rounds 3 --dtype bf16 --configs fp4e2m1g128,fp8e4m3g128,int4g128 --knobs "$K" > gpurun_out/r2_sweep_fp12_bf16.jsonl 2>> gpurun_out/r2_sweep12.err
python tools/sweep_r2.py --model opt-6.7b --configs int4pt --iters 5 --rounds 3 --knobs 'pt_hint=0;pt_hint=2;pt_hint=1;pt_reverse=0;pt_hint=2,seg_stages=3' > gpurun_out/r2_sweep_pt12.jsonl 2>> gpurun_out/r2_sweep12.err
python tools/sweep_r2.py --model llama-2-7b --configs int8pc,int4pc --iters 5 --rounds 3 --knobs 'seg_hint=2;seg_hint=0;wide_seg4=0' > gpurun_out/r2_sweep_pc12.jsonl 2>> gpurun_out/r2_sweep12.err
tail -8 gpurun_out/r2_tests12.log; tail -3 gpurun_out/r2_bench12.err; cut -c1-1200 gpurun_out/r2_bench12.json
python - <<'PY'
import json,glob
for f in ['fp12','fp12_bf16','pt12','pc12']:
    for l in open(f'gpurun_out/r2_sweep_{f}.jsonl'):
        d=json.loads(l)
        if 'knobs' in d: print(f, d['config'], d['knobs'], round(d.get('GBps',0)), d.get('GBps_all'), d.get('error',''))
PY
tail -5 gpurun_out/r2_sweep12.err
