#!/bin/bash
# Round-2 evidence (B200_PROFILING.md recipe): plain run first (must exit 0), the ncu launch list of the bench command, one
# full capture of the bench kernel, and full captures of the other kernel families (one launch each, on fresh data).
# Numbers printed under ncu are never bench values.  Summaries (tools/ncu_summary.py) are written next to the reports.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
B="python bench.py --steps 3 --warmup 3 --skip-e2e --skip-cpu --skip-configs"
$B > gpurun_out/r2_bench_plain.log 2>&1 || { tail -5 gpurun_out/r2_bench_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:seg4|row_tma|col_|minmax|range_init|seg_" -c 400 --csv --log-file gpurun_out/r2_launches_bench.csv $B > gpurun_out/r2_ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:seg4 --launch-skip 4 -c 1 -f -o gpurun_out/r2_prof_bench $B > gpurun_out/r2_ncu_full_bench.log 2>&1
python tools/ncu_summary.py gpurun_out/r2_prof_bench.ncu-rep > gpurun_out/r2_ncu_bench_kernel.txt 2>&1
ncu -i gpurun_out/r2_prof_bench.ncu-rep --page raw --csv > gpurun_out/r2_ncu_bench_kernel.csv 2>/dev/null
cap() {  # name kernel-regex skip microbench-args...
  local name=$1 rx=$2 skip=$3; shift 3
  python tools/microbench.py "$@" > gpurun_out/r2_mb_$name.log 2>&1 || { echo "plain run of $name failed"; tail -3 gpurun_out/r2_mb_$name.log; return; }
  ncu --set full --clock-control none --import-source on -k regex:$rx --launch-skip $skip -c 1 -f -o gpurun_out/r2_prof_$name python tools/microbench.py "$@" > gpurun_out/r2_ncu_$name.log 2>&1
  python tools/ncu_summary.py gpurun_out/r2_prof_$name.ncu-rep > gpurun_out/r2_ncu_$name.txt 2>&1
  rm -f gpurun_out/r2_prof_$name.ncu-rep
}
cap fp4e2m1_f16 seg4 0 --fresh --model llama-2-13b --layers 2 --configs fp4e2m1g128
cap fp8e4m3_f16 seg4 0 --fresh --model llama-2-13b --layers 2 --configs fp8e4m3g128
cap fp8e4m3_bf16 seg4 0 --fresh --model llama-2-13b --layers 2 --dtype bf16 --configs fp8e4m3g128
cap col_int4g128 col_wide 0 --fresh --layers 1 --configs int4g128_d1
cap col_fp4g128 col_wide 0 --fresh --layers 1 --configs fp4e2m1g128_d1
cap col_double col_double 0 --fresh --layers 1 --configs fp4e2m1g128_d1_double
cap minmax minmax 3 --model opt-6.7b --layers 4 --iters 1 --configs int4pt
cap pt_apply seg4 3 --model opt-6.7b --layers 4 --iters 1 --configs int4pt
cap row_tma row_tma 3 --model llama-2-7b --layers 4 --iters 1 --configs int4pc
cap pc_seg4 seg4 3 --model llama-2-7b --layers 4 --iters 1 --configs int4pc
cap w8pc_bf16 'seg4|row_tma' 3 --model llama-2-7b --layers 4 --iters 1 --dtype bf16 --configs int8pc
ls -la gpurun_out | grep r2_ncu | head -30; cat gpurun_out/r2_ncu_bench_kernel.txt; head -12 gpurun_out/r2_launches_bench.csv | cut -c1-200
