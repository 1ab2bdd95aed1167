#!/bin/bash
# round 2, experiment 11: repeatability of the tile-size result (6 interleaved rounds per knob set)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
P=iron_weight_only_quant_b200
run() { IWQ_LIB=$PWD/$P/libiwq_b200_$1.so python tools/sweep_r2.py --model llama-2-70b --layers 10 --configs int4g128 --iters 10 --rounds 6 --knobs "$2" | sed "s/^{/{\"lib\": \"$1\", /"; }
(
python tools/sweep_r2.py --model llama-2-70b --layers 10 --configs int4g128 --iters 10 --rounds 6 --knobs "seg_hint=0;seg_hint=2"
run cw16 'seg_stages=4;seg_stages=4,seg_hint=2;seg_stages=5;seg_stages=3'
run cw20 'seg_stages=4;seg_stages=4,seg_hint=2'
run cw24 'seg_stages=4;seg_stages=4,seg_hint=2'
python tools/sweep_r2.py --model llama-2-70b --layers 10 --configs int4g128 --iters 10 --rounds 6 --knobs "seg_hint=0;seg_hint=2"
) > gpurun_out/r2_sweep_seg11.jsonl 2> gpurun_out/r2_sweep11.err
python - <<'PY'
import json
for l in open('gpurun_out/r2_sweep_seg11.jsonl'):
    d=json.loads(l)
    if 'knobs' in d: print(d.get('lib','cw8'), d['knobs'], round(d['GBps']), d['GBps_all'])
PY
tail -5 gpurun_out/r2_sweep11.err
