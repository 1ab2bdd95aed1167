#!/bin/bash
# round 2, experiment 26: FP8 E4M3 on fp16 (38 KB direct table): 40 KB tiles x 4 stages vs 48 KB tiles x 3 stages.  The IWQ_FP_TILE40 knob and the
# 40 KB-tile instantiation this script drove were measured slower and removed again (profiles/r02_sweep_fp26_tile40.jsonl); kept for the record.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "fp_chain or fp_codec or fp_rne or fuzz or variants or approx or smoke or hf_arch" > gpurun_out/r2_tests26.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2_tests26.log
K='fp_tile40=1;fp_tile40=0'
python tools/sweep_r2.py --fresh --model llama-2-13b --layers 8 --iters 4 --rounds 3 --configs fp8e4m3g128,fp6e3m2g128,fp4e2m1g128 --knobs "$K" > gpurun_out/r2_sweep_fp26.jsonl 2> gpurun_out/r2_sweep26.err
tail -n 3 gpurun_out/r2_tests26.log
python - <<'PY'
import json
for l in open('gpurun_out/r2_sweep_fp26.jsonl'):
    d=json.loads(l)
    if 'knobs' in d: print(d['config'], d['knobs'], round(d.get('GBps',0)), d.get('GBps_all'), d.get('error',''))
PY
tail -n 3 gpurun_out/r2_sweep26.err
