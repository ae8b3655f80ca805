import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gnn_tumor_seg_b200 import ops
dev = torch.device("cuda:0")
def rel(a, b): return (a.double().cpu() - b).abs().max().item() / b.abs().max().item()
for (K, Mo, No) in [(100000, 1024, 1024), (100000, 512, 512), (100000, 256, 1024), (100000, 1024, 256), (20000, 1024, 1024)]:
    g = torch.Generator().manual_seed(1)
    A = torch.randn(K, Mo, generator=g); B1 = torch.randn(K, No, generator=g); B2 = torch.randn(K, No, generator=g)
    r1 = A.double().T @ B1.double(); r2 = A.double().T @ B2.double()
    c1, c2, cs = ops.gemm_tn2_colsum(A.to(dev), B1.to(dev), B2.to(dev), mode="tf32x3")
    d1, ds = ops.gemm_tn_colsum(A.to(dev), B1.to(dev), mode="tf32x3")
    e = (c1.double().cpu() - r1).abs()
    print(K, Mo, No, "dual", rel(c1, r1), rel(c2, r2), "single", rel(d1, r1), "colsum", rel(cs, A.double().sum(0)),
          "argmax err at", divmod(int(e.argmax()), No))
