#!/bin/bash
# round 2, experiment 4: ncu of the minifloat lookup variants (why is the direct table slow for fp16 and fast for bf16?)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() {  # name dtype config env...
  local name=$1 dt=$2 cfgn=$3; shift 3
  env "$@" ncu --set full --clock-control none --import-source on -k regex:seg4 --launch-skip 3 -c 1 -f -o gpurun_out/r2_prof_$name \
    python tools/microbench.py --layers 1 --iters 2 --dtype $dt --configs $cfgn > gpurun_out/r2_ncu_$name.log 2>&1
  python tools/ncu_summary.py gpurun_out/r2_prof_$name.ncu-rep > gpurun_out/r2_ncu_$name.txt 2>&1
}
run fp4_f16_direct f16 fp4e2m1g128 IWQ_FP_MODE=2
run fp4_f16_step f16 fp4e2m1g128 IWQ_FP_MODE=0
run fp8_bf16_direct bf16 fp8e4m3g128 IWQ_FP_MODE=2
run fp8_f16_repl f16 fp8e4m3g128 IWQ_FP_MODE=1 IWQ_FP_REP=3
run fp4_bf16_direct bf16 fp4e2m1g128 IWQ_FP_MODE=2
cat gpurun_out/r2_ncu_fp4_f16_direct.txt gpurun_out/r2_ncu_fp4_f16_step.txt gpurun_out/r2_ncu_fp8_bf16_direct.txt gpurun_out/r2_ncu_fp8_f16_repl.txt gpurun_out/r2_ncu_fp4_bf16_direct.txt
