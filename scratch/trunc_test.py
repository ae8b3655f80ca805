import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gnn_tumor_seg_b200 import ops
dev = torch.device("cuda:0")
torch.manual_seed(0)
M, N, K = 512, 256, 256
A = torch.randn(M, K); B = torch.randn(N, K)
def trunc(x):
    return (x.view(torch.int32) & ~0x1FFF).view(torch.float32)
def rne(x):   # round to nearest even at 13 dropped bits
    i = x.view(torch.int32).to(torch.int64) & 0xFFFFFFFF
    lsb = (i >> 13) & 1
    i = (i + 0xFFF + lsb) & ~0x1FFF
    return (i & 0xFFFFFFFF).to(torch.int64).apply_(lambda v: v - (1 << 32) if v >= (1 << 31) else v).to(torch.int32).view(torch.float32)
def rna(x):   # round to nearest, ties away (cvt.rna.tf32)
    i = x.view(torch.int32).to(torch.int64) & 0xFFFFFFFF
    i = (i + 0x1000) & ~0x1FFF
    return (i & 0xFFFFFFFF).to(torch.int64).apply_(lambda v: v - (1 << 32) if v >= (1 << 31) else v).to(torch.int32).view(torch.float32)
out = ops.gemm_nt(A.to(dev), B.to(dev), mode="tf32").cpu().double()
for name, f in (("trunc", trunc), ("rne", rne), ("rna", rna)):
    ref = f(A).double() @ f(B).double().T
    print(name, "max rel err", ((out - ref).abs().max() / ref.abs().max()).item())
print("full", ((out - A.double() @ B.double().T).abs().max() / (A.double() @ B.double().T).abs().max()).item())
