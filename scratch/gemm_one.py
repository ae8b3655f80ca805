import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gnn_tumor_seg_b200 import ops
dev = torch.device("cuda:0")
mode = sys.argv[1] if len(sys.argv) > 1 else "tf32x3"
which = sys.argv[2] if len(sys.argv) > 2 else "nt"
N, D = 90000, 256
A = [torch.randn(N, D, device=dev) for _ in range(3)]
W1 = torch.randn(D, D, device=dev); W2 = torch.randn(D, D, device=dev); b = torch.randn(D, device=dev)
for i in range(12):
    if which == "nt":
        ops.gemm_nt(A[i % 3], W1, A[(i + 1) % 3], W2, bias=b, act=1, mode=mode)
    else:
        ops.gemm_tn_colsum(A[i % 3], A[(i + 1) % 3], mode=mode)
torch.cuda.synchronize()
print("ok")
