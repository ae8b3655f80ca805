#!/bin/bash
# ncu --set full of the weight-gradient (TN) pair kernels (gpurun -- bash tools/profile_tn.sh <tag>)
tag=${1:-r02}
G="python tools/gemm_bench.py tf32x3 3"
$G > gpurun_out/${tag}_tn_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gemm_x3ts2 -s 20 -c 4 -o gpurun_out/${tag}_prof_gemm_tn $G > gpurun_out/${tag}_ncu_tn.log 2>&1
tail -3 gpurun_out/${tag}_ncu_tn.log
