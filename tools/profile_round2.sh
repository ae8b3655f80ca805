#!/bin/bash
# Round-2 profile (gpurun -- bash tools/profile_round2.sh <tag>): plain run first (must exit 0), then the ncu launch
# list of ONE eager training step, then `ncu --set full` of the hot kernels.  Everything lands in gpurun_out/.
tag=${1:-r02}
out=gpurun_out
mkdir -p $out
S="python tools/one_step.py"
$S > $out/${tag}_one_step_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $out/${tag}_launches.csv $S > $out/${tag}_ncu_launches.log 2>&1
G="python tools/gemm_bench.py tf32x3 3"
$G > $out/${tag}_gemm_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gemm_x3ntw -s 4 -c 2 -o $out/${tag}_prof_gemm_ntw $G > $out/${tag}_ncu_gemm.log 2>&1
ls -la $out | tail -8
