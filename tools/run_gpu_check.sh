#!/bin/bash
# GPU box: parity tests, then the kernel micro-benchmark on the configurations given as $1 (comma list)
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
python tools/microbench.py --layers 3 --iters 10 --configs ${1:-int4g128,fp4e2m1g128,fp8e4m3g128} > gpurun_out/mb_now.log 2>&1
cat gpurun_out/mb_now.log
