#!/bin/bash
for pf in 0 4 8 16; do for d in 0 16; do
  echo -n "pf=$pf dbg=$d: "; GTS_X3_PF=$pf GTS_X3_DBG=$d ITERS=30 timeout 60 python scratch/gemm_only.py tf32x3 2>&1 | tr '\n' ' '; echo
done; done
