#!/bin/bash
# round 2: full gate + both bench arms (what the driver runs at round end)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests_final.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2_tests_final.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; echo "smoke rc $?" >> gpurun_out/r2_smoke.log
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2_bench_ref_final.json 2> gpurun_out/r2_bench_final.err
python bench.py > gpurun_out/r2_bench_final.json 2>> gpurun_out/r2_bench_final.err; echo "bench rc $?" >> gpurun_out/r2_bench_final.err
tail -n 6 gpurun_out/r2_tests_final.log; tail -n 2 gpurun_out/r2_smoke.log; tail -n 3 gpurun_out/r2_bench_final.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_final.json').read().strip().splitlines()[-1])
c=d.pop('configs')
print(round(d['value'],1), d['ms_per_step'], d['roofline']['frac'], 'e2e', round(d['e2e']['value'],1), d['clocks'], 'cpu', d['cpu_baseline']['value'], d['cpu_baseline']['kind'])
for k,v in c.items():
    if isinstance(v,dict):
        for kk,vv in v.items():
            if isinstance(vv,dict): print(' ',k,kk,round(vv.get('GBps',0)),round(vv.get('frac',0),3),vv.get('parity_ok'))
        if 'GBps' in v: print(' ',k, round(v['GBps'],1), v.get('frac'), v.get('parity_ok'), v.get('ms'))
r=json.loads(open('gpurun_out/r2_bench_ref_final.json').read().strip().splitlines()[-1])
print('ref', r['value'], r['cpu_baseline']['kind'], r['configs']['config1']['GBps'])
PY
