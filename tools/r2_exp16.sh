#!/bin/bash
# round 2, experiment 16: seg4 with 8 / 16 vectors per lane (fewer lanes repeat the segment parameters), one bulk store per tile
# (the seg_store knob of this script -- one bulk store per tile -- was measured 1-2 % slower and removed again; those knob sets now report "unknown knob")
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for v in 8 16; do
  IWQ_SEG_VPL=$v timeout 900 python -m pytest tests -m gpu -x -q -k "int_matches or fp_chain or fp_codec or fuzz or oracle_all or per_tensor or bfp or odd_group or packed" > gpurun_out/r2_tests16_vpl$v.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2_tests16_vpl$v.log
done
IWQ_SEG_STORE=1 timeout 900 python -m pytest tests -m gpu -x -q -k "int_matches or fuzz or oracle_all or per_tensor" > gpurun_out/r2_tests16_store.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2_tests16_store.log
K='seg_vpl=4;seg_vpl=8;seg_vpl=16;seg_vpl=4,seg_store=1;seg_vpl=8,seg_store=1;seg_vpl=16,seg_store=1;seg_vpl=8,seg_stages=3;seg_vpl=16,seg_stages=3;seg_vpl=4'
python tools/sweep_r2.py --model llama-2-70b --layers 10 --configs int4g128 --iters 10 --rounds 4 --knobs "$K" > gpurun_out/r2_sweep_seg16.jsonl 2> gpurun_out/r2_sweep16.err
K='seg_vpl=4;seg_vpl=8;seg_vpl=16'
python tools/sweep_r2.py --fresh --model llama-2-13b --layers 8 --iters 4 --rounds 2 --configs fp4e2m1g128,fp8e4m3g128,fp6e3m2g128,int4g32,int4g64,int4g256,bfp4g128 --knobs "$K" > gpurun_out/r2_sweep_fp16.jsonl 2>> gpurun_out/r2_sweep16.err
python tools/sweep_r2.py --model opt-6.7b --configs int4pt --iters 5 --rounds 2 --knobs "$K" > gpurun_out/r2_sweep_pt16.jsonl 2>> gpurun_out/r2_sweep16.err
tail -3 gpurun_out/r2_tests16_vpl8.log gpurun_out/r2_tests16_vpl16.log gpurun_out/r2_tests16_store.log
python - <<'PY'
import json
for f in ['seg16','fp16','pt16']:
    for l in open(f'gpurun_out/r2_sweep_{f}.jsonl'):
        d=json.loads(l)
        if 'knobs' in d: print(f, d['config'], d['knobs'], round(d.get('GBps',0)), d.get('GBps_all'), d.get('error',''))
PY
tail -5 gpurun_out/r2_sweep16.err
