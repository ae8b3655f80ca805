"""ONE eager GAT training step (BASELINE configs[2]: layer_sizes [256]*4, heads [4,4,4,4], residuals [F,F,T,F], B = 6 x
15k-node RAGs) after a warm-up step — the launch list for `ncu --metrics gpu__time_duration.sum`."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("GTS_SYNTH_CACHE", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), ".synth_cache"))
import numpy as np
import torch
from gnn_tumor_seg_b200 import graph as G, networks, ops, synth
from gnn_tumor_seg_b200.trainer import FusedAdamW

dev = torch.device("cuda:0")
graphs = [synth.make_graph(s) for s in range(6)]
bg = G.batch([G.from_edge_list(g.src, g.dst, g.n_nodes) for g in graphs]).to(dev)
x = torch.as_tensor(np.concatenate([g.features for g in graphs])).to(dev)
y = torch.as_tensor(np.concatenate([g.labels for g in graphs])).to(dev)
torch.manual_seed(0)
net = networks.GAT(20, [256] * 4, 4, [4, 4, 4, 4], [False, False, True, False]).to(dev)
w = torch.tensor([0.1, 1.0, 2.0, 2.0], device=dev)
opt = FusedAdamW(net.parameters(), lr=1e-4, weight_decay=1e-4)
for it in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2):
    if it == 1:
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStart()
    loss = ops.weighted_cross_entropy(net(bg, x), y, w)
    opt.zero_grad()
    loss.backward()
    opt.step()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("loss", float(loss))
