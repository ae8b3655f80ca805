#!/bin/bash
# time the CTA-pair 3xTF32 kernel with parts of the pipeline disabled (results are garbage; timing only)
for d in 0 16 1 2 32 4 6 7 23; do
  echo -n "dbg=$d: "; GTS_X3_DBG=$d ITERS=30 timeout 60 python scratch/gemm_only.py tf32x3 2>&1 | tr '\n' ' '; echo
done
