#!/bin/bash
# ncu --set full of the seg-max forward / backward kernels on the headline shapes (gpurun -- bash tools/profile_seg.sh <tag>)
tag=${1:-r02}
ITERS=5 python scratch/seg_only.py > gpurun_out/${tag}_seg_plain.log 2>&1 &&
ITERS=5 ncu --set full --clock-control none --import-source on -k regex:segmax -s 6 -c 6 -o gpurun_out/${tag}_prof_seg python scratch/seg_only.py > gpurun_out/${tag}_ncu_seg.log 2>&1
tail -3 gpurun_out/${tag}_ncu_seg.log
