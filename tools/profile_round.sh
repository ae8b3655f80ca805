#!/bin/bash
# Round profile: run on the GPU box through gpurun (tools/profile_round.sh <tag>); everything lands in gpurun_out/.
#   1. plain bench (must exit 0) then the ncu launch list of the same command (device time of every launch);
#   2. one `ncu --set full` capture per hot kernel (after the same command ran clean without ncu).
# Read the reports here with tools/summarize_launches.py / tools/ncu_traffic.py and commit the summaries under profiles/.
tag=${1:-r01}
out=gpurun_out
mkdir -p $out
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$B > $out/${tag}_bench_plain.json 2> $out/${tag}_bench_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 700 -c 420 --csv --log-file $out/${tag}_launches.csv $B > $out/${tag}_ncu_launches.log 2>&1
python scratch/gemm_one.py tf32x3 nt > $out/${tag}_g_nt.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gemm_x3ts2 -s 6 -c 1 -o $out/${tag}_prof_gemm_nt python scratch/gemm_one.py tf32x3 nt > $out/${tag}_ncu_gnt.log 2>&1
python scratch/gemm_one.py tf32x3 tn > $out/${tag}_g_tn.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gemm_x3ts2 -s 6 -c 1 -o $out/${tag}_prof_gemm_tn python scratch/gemm_one.py tf32x3 tn > $out/${tag}_ncu_gtn.log 2>&1
ITERS=20 python scratch/seg_only.py > $out/${tag}_seg.log 2>&1 &&
ITERS=20 ncu --set full --clock-control none --import-source on -k regex:segmax -s 30 -c 2 -o $out/${tag}_prof_seg python scratch/seg_only.py > $out/${tag}_ncu_seg.log 2>&1
ls -la $out | tail -20
