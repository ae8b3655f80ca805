"""Micro-benchmark of the hot GEMM shapes of the 7x256 SAGE-pool step (run on the GPU box):
NT K=512 two-source (+bias+ReLU), NT K=256 (+bias+ReLU), NT K=256 masked, NT K=512 masked, TN, TN2 — CUDA events,
3 rotating 92 MB operands (larger than L2 together with the outputs), plus the max relative error of the first
4096 rows against an fp64 product.  Usage: python tools/gemm_bench.py [mode] [iters]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gnn_tumor_seg_b200 import ops

dev = torch.device("cuda:0")
mode = sys.argv[1] if len(sys.argv) > 1 else "tf32x3"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 50
N, D = 90000, 256
torch.manual_seed(0)
A = [torch.randn(N, D, device=dev) for _ in range(3)]
W1 = torch.randn(D, D, device=dev) / 16
W2 = torch.randn(D, D, device=dev) / 16
b = torch.randn(D, device=dev)


def err(out, ref):
    return ((out[:4096].double() - ref).abs().max() / ref.abs().max()).item()


cases = {
    "nt_k512_relu": lambda i: ops.gemm_nt(A[i % 3], W1, A[(i + 1) % 3], W2, bias=b, act=ops.ACT_RELU, mode=mode),
    "nt_k256_relu": lambda i: ops.gemm_nt(A[i % 3], W1, bias=b, act=ops.ACT_RELU, mode=mode),
    "nt_k256_mask": lambda i: ops.gemm_nt(A[i % 3], W1, act=ops.ACT_MASK_POS, aux=A[(i + 2) % 3], mode=mode),
    "nt_k512_mask": lambda i: ops.gemm_nt(A[i % 3], W1, A[(i + 1) % 3], W2, act=ops.ACT_MASK_POS, aux=A[(i + 2) % 3], mode=mode),
    "tn_256": lambda i: ops.gemm_tn_colsum(A[i % 3], A[(i + 1) % 3], mode=mode),
    "tn2_256": lambda i: ops.gemm_tn2_colsum(A[i % 3], A[(i + 1) % 3], A[(i + 2) % 3], mode=mode),
}
a0, a1, a2 = (A[k][:4096].double() for k in range(3))
refs = {
    "nt_k512_relu": torch.relu(a0 @ W1.double().T + a1 @ W2.double().T + b.double()),
    "nt_k256_relu": torch.relu(a0 @ W1.double().T + b.double()),
    "nt_k256_mask": (a0 @ W1.double().T) * (a2 > 0),
    "nt_k512_mask": (a0 @ W1.double().T + a1 @ W2.double().T) * (a2 > 0),
}
for name, f in cases.items():
    try:
        out = f(0)
    except Exception as e:      # an op missing in this build must not hide the other numbers
        print(mode, name, "FAILED", repr(e)[:200])
        continue
    torch.cuda.synchronize()
    e = err(out, refs[name]) if name in refs else float("nan")
    for i in range(10):
        f(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        f(i)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / iters * 1e3
    print(f"{mode} {name:14s} {us:8.1f} us   max-rel-err(first 4096 rows vs fp64) {e:.2e}", flush=True)
