"""Three warm-up steps, then ONE eager training step of the headline config (7x256, B = 6 x 15k-node RAGs) between
two cudaProfilerStart/Stop marks — the launch list of a step for `ncu --profile-from-start off` or a plain kernel list."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("GTS_SYNTH_CACHE", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), ".synth_cache"))
import numpy as np
import torch
from gnn_tumor_seg_b200 import graph as G, networks, ops, synth
from gnn_tumor_seg_b200.trainer import SageTrainer

dev = torch.device("cuda:0")
graphs = [synth.make_graph(s) for s in range(6)]
bg = G.batch([G.from_edge_list(g.src, g.dst, g.n_nodes) for g in graphs]).to(dev)
x = torch.as_tensor(np.concatenate([g.features for g in graphs])).to(dev)
y = torch.as_tensor(np.concatenate([g.labels for g in graphs])).to(dev)
torch.manual_seed(0)
net = networks.GraphSage(20, [256] * 7, 4, "pool", 0).to(dev)
tr = SageTrainer(net, torch.tensor([0.1, 1.0, 2.0, 2.0], device=dev), lr=1e-4, weight_decay=1e-4)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1
for _ in range(n):
    loss = tr.step(bg, x, y)
torch.cuda.synchronize()
print("loss", float(loss))
