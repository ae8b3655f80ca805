#!/bin/bash
# compute-sanitizer memcheck over the smallest SAGE / GAT / CTA-pair GEMM edge-shape cases.
# One tool per gpurun call (B200_PROFILING.md).  Usage: gpurun -- bash tools/sanitize_small.sh [tag]
TAG=${1:-r02}
SEL='integer_valued or gat_stack or pair_kernel_edges or weighted_ce or adamw or (test_gemm_nt and 333) or (csr_build and small) or (segmax_fwd and 20)'
mkdir -p gpurun_out
python -m pytest tests/test_gpu_models.py tests/test_gpu_gemm_loss.py tests/test_gpu_graph_segmax.py -q -m gpu -x -k "$SEL" > gpurun_out/${TAG}_memcheck_plain.log 2>&1 || { echo "plain run failed"; tail -20 gpurun_out/${TAG}_memcheck_plain.log; exit 1; }
tail -2 gpurun_out/${TAG}_memcheck_plain.log
timeout 1500 compute-sanitizer --tool memcheck --log-file gpurun_out/${TAG}_memcheck.log \
  python -m pytest tests/test_gpu_models.py tests/test_gpu_gemm_loss.py tests/test_gpu_graph_segmax.py -q -m gpu -x -k "$SEL" > gpurun_out/${TAG}_memcheck_pytest.log 2>&1
echo "sanitizer exit $?"
tail -3 gpurun_out/${TAG}_memcheck_pytest.log
grep -c "Invalid\|Error:" gpurun_out/${TAG}_memcheck.log; tail -5 gpurun_out/${TAG}_memcheck.log
