import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gnn_tumor_seg_b200 import ops
dev = torch.device("cuda:0")
mode = sys.argv[1] if len(sys.argv) > 1 else "tf32x3"
N, D = 90000, 256
A = [torch.randn(N, D, device=dev) for _ in range(3)]
W1 = torch.randn(D, D, device=dev); W2 = torch.randn(D, D, device=dev); b = torch.randn(D, device=dev)
iters = int(os.environ.get("ITERS", "100"))
def nt(i): return ops.gemm_nt(A[i % 3], W1, A[(i + 1) % 3], W2, bias=b, act=1, mode=mode)
def nt1(i): return ops.gemm_nt(A[i % 3], W1, bias=b, act=1, mode=mode)
def tn(i): return ops.gemm_tn(A[i % 3], A[(i + 1) % 3], mode=mode)
for name, f in (("nt_k512", nt), ("nt_k256", nt1), ("tn_256x256", tn)):
    for i in range(iters): f(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters): f(i)
    e1.record(); torch.cuda.synchronize()
    print(mode, name, "ms", e0.elapsed_time(e1) / iters)
