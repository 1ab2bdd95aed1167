#!/bin/bash
# round 2, experiment 13: per-tensor range pass through the bulk-copy ring (minmax_tma_kernel) vs 128-bit global loads
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "per_tensor or int_matches or wrappers or fuzz or sharded or odd" > gpurun_out/r2_tests13.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2_tests13.log
for v in "IWQ_MM_TMA=0" "IWQ_MM_TMA=1" "IWQ_MM_TMA=1 IWQ_MM_HINT=2" "IWQ_MM_TMA=1 IWQ_MM_STAGES=3" "IWQ_MM_TMA=1 IWQ_MM_HINT=2 IWQ_PT_HINT=2"; do
  echo "== $v world 1"; env $v python tools/config4_bench.py 2>>gpurun_out/r2_c4_13.err | cut -c100-330
  echo "== $v fake world 8"; env $v IWQ_FAKE_WORLD=8 python tools/config4_bench.py 2>>gpurun_out/r2_c4_13.err | cut -c100-330
done > gpurun_out/r2_c4_13.txt
tail -5 gpurun_out/r2_tests13.log; cat gpurun_out/r2_c4_13.txt; tail -3 gpurun_out/r2_c4_13.err
