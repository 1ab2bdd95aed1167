#!/bin/bash
# quant_dim = 1 kernel sweep: two-sweep kernel (TILE=-1), register-resident (MODE=1), warp-private cp.async ring (MODE=0)
for knobs in "IWQ_COL_STAGES=1" "IWQ_COL_STAGES=1 IWQ_COL_CTAS=4" "IWQ_COL_STAGES=1 IWQ_COL_CTAS=3" "IWQ_COL_STAGES=2" "IWQ_COL_STAGES=2 IWQ_COL_CTAS=2" "IWQ_COL_STAGES=3"; do
  env $knobs python tools/microbench.py --layers 3 --iters 10 --configs int4g128_d1,int4g128sym_d1,int4g64_d1,int4g256_d1,fp4e2m1g128_d1
done
