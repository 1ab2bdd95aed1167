#!/bin/bash
# round 2, experiment 21: per-channel on 7B shapes, the tree of commit f50f163 (first 48 KB-tile version) against HEAD, same box
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
(
for rep in 1 2; do
  (cd _old && python tools/sweep_r2.py --model llama-2-7b --configs int8pc,int4pc --iters 10 --rounds 2) | sed "s/^{/{\"tree\": \"old\", /"
  python tools/sweep_r2.py --model llama-2-7b --configs int8pc,int4pc --iters 10 --rounds 2 | sed "s/^{/{\"tree\": \"new\", /"
done
) > gpurun_out/r2_sweep_pc21.jsonl 2> gpurun_out/r2_sweep21.err
python - <<'PY'
import json
for l in open('gpurun_out/r2_sweep_pc21.jsonl'):
    d=json.loads(l)
    if 'knobs' in d: print(d['tree'], d['config'], round(d.get('GBps',0)), d.get('GBps_all'))
PY
tail -n 3 gpurun_out/r2_sweep21.err
