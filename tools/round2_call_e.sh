#!/bin/bash
# gpurun -- bash tools/round2_call_e.sh : `ncu --set full` of the kernels round 2 had no capture of — the three fused GAT
# edge kernels (wide layers, H x F = 4 x 256) and the weight-gradient (TN) CTA-pair GEMM — each after a plain run.
out=gpurun_out; mkdir -p $out
N="ncu --set full --clock-control none --import-source on"
timeout 40 python tools/one_step_gat.py 2 > $out/r02f_gat_plain.log 2>&1; rc=$?; echo "gat plain rc=$rc"
if [ $rc -eq 0 ]; then
  timeout 75 $N --profile-from-start off --kernel-name-base demangled -k 'regex:gat_(fwd|bwd_dst|bwd_src)_kernel<2' -c 6 \
    -o $out/r02f_prof_gat python tools/one_step_gat.py 2 > $out/r02f_ncu_gat.log 2>&1; echo "gat ncu rc=$?"
  if [ ! -f $out/r02f_prof_gat.ncu-rep ]; then
    timeout 75 $N --profile-from-start off -k 'regex:gat_(fwd|bwd_dst|bwd_src)_kernel' -s 2 -c 7 \
      -o $out/r02f_prof_gat python tools/one_step_gat.py 2 > $out/r02f_ncu_gat2.log 2>&1; echo "gat ncu (base names) rc=$?"
  fi
fi
timeout 60 $N -k regex:gemm_x3ts2 -s 3 -c 2 -o $out/r02f_prof_gemm_tn python tools/one_step.py 1 > $out/r02f_ncu_tn.log 2>&1; echo "tn ncu rc=$?"
ls -la $out/*.ncu-rep; tail -2 $out/r02f_ncu_gat.log $out/r02f_ncu_tn.log
