#!/bin/bash
# ncu full captures of the minifloat and column kernels (one launch each, after warm-up launches)
set -x
mkdir -p gpurun_out
for c in fp4e2m1g128 fp8e4m3g128 int4g128_d1 int4pc; do
  ncu --set full --clock-control none --import-source on --launch-skip 3 -c 1 -k regex:'seg4|col_kernel|row' -f -o gpurun_out/prof_$c \
    python tools/microbench.py --layers 1 --iters 1 --configs $c > gpurun_out/ncu_$c.log 2>&1
done
python tools/microbench.py --layers 3 --iters 10 --configs int4g128,fp4e2m1g128,fp8e4m3g128,fp4e2m1g128asym,int4pc,int8pc,int4pt,bfp4g128,int4g128_d1 > gpurun_out/mb_now.log 2>&1
cat gpurun_out/mb_now.log
