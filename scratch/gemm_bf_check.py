"""Accuracy and time of the NT GEMM in the current GTS_X3_BF16 setting against an fp64 product (K = 256 and the
two-source K = 512 form, bias + ReLU epilogue off for the error so that cancellation is visible)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gnn_tumor_seg_b200 import ops
dev = torch.device("cuda:0")
torch.manual_seed(0)
N, D = 90000, 256
A = [torch.randn(N, D, device=dev) for _ in range(3)]
W1 = torch.randn(D, D, device=dev) / 16; W2 = torch.randn(D, D, device=dev) / 16; b = torch.randn(D, device=dev)
tag = "x3_bf16=" + os.environ.get("GTS_X3_BF16", "default")
for name, fn, ref in (
        ("k256", lambda i: ops.gemm_nt(A[i % 3], W1, mode="tf32x3"), lambda: A[0][:4096].double() @ W1.double().t()),
        ("k512", lambda i: ops.gemm_nt(A[i % 3], W1, A[(i + 1) % 3], W2, bias=b, act=1, mode="tf32x3"),
         lambda: torch.relu(A[0][:4096].double() @ W1.double().t() + A[1][:4096].double() @ W2.double().t() + b.double()))):
    out = fn(0)
    r = ref()
    err = (out[:4096].double() - r).abs().max().item() / r.abs().max().item()
    rms = ((out[:4096].double() - r).pow(2).mean().sqrt() / r.pow(2).mean().sqrt()).item()
    for i in range(10):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(60):
        fn(i)
    e1.record(); torch.cuda.synchronize()
    print("%s %s: max rel err %.3e  rms rel err %.3e  %.1f us" % (tag, name, err, rms, e0.elapsed_time(e1) / 60 * 1e3), flush=True)
