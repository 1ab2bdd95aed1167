#!/bin/bash
# Round-end evidence for the bench kernel: plain run first (must exit 0), then the ncu launch list and one
# full capture of the same command (B200_PROFILING.md recipe).  Numbers printed under ncu are never bench values.
set -x
mkdir -p gpurun_out
python bench.py --steps 3 --warmup 3 --skip-e2e --skip-cpu > gpurun_out/bench_plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 3 --warmup 3 --skip-e2e --skip-cpu > gpurun_out/ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:seg4 --launch-skip 4 -c 1 -f -o gpurun_out/prof_bench \
  python bench.py --steps 3 --warmup 3 --skip-e2e --skip-cpu > gpurun_out/ncu_full_bench.log 2>&1
tail -2 gpurun_out/bench_plain.log | cut -c1-300
