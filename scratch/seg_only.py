import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gnn_tumor_seg_b200 import graph as G, ops, synth
dev = torch.device("cuda:0")
graphs = [synth.make_graph(s) for s in range(6)]
bg = G.batch([G.from_edge_list(g.src, g.dst, g.n_nodes) for g in graphs]).to(dev)
N = bg.number_of_nodes()
Ps = [torch.relu(torch.randn(N, 256, device=dev)) for _ in range(3)]
indptr, indices = bg.csr
iters = int(os.environ.get("ITERS", "300"))
for i in range(iters):
    n, a = ops.segmax_fwd(Ps[i % 3], indptr, indices)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(iters): ops.segmax_fwd(Ps[i % 3], indptr, indices)
e1.record(); torch.cuda.synchronize()
print(os.environ.get("GTS_SEGMAX_GENERIC", "wide"), "segmax_fwd ms", e0.elapsed_time(e1) / iters)
dN = torch.randn(N, 256, device=dev)
for i in range(min(50, iters)): ops.segmax_bwd(dN, a, N)
torch.cuda.synchronize()
e0.record()
for i in range(min(100, iters)): ops.segmax_bwd(dN, a, N)
e1.record(); torch.cuda.synchronize()
print("segmax_bwd ms", e0.elapsed_time(e1) / min(100, iters))
