import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from gnn_tumor_seg_b200 import graph as G, networks, ops, synth
from oracle import gat_ref
dev = torch.device("cuda:0")
g = synth.make_graph(1)
bg = G.from_edge_list(g.src, g.dst, g.n_nodes)
feats, labels = torch.as_tensor(g.features), torch.as_tensor(g.labels)
w = torch.tensor([0.1, 1., 2., 2.])
cfg = (20, [256] * 4, 4, [4, 4, 4, 4], [False, False, True, False])
torch.manual_seed(0)
net = networks.GAT(*cfg)
ref64 = gat_ref.GATRef(*cfg).double()
ref64.load_state_dict({k: v.double() for k, v in net.state_dict().items()})
s, d = bg.edges()
F.cross_entropy(ref64((s, d), feats.double()), labels, weight=w.double()).backward()
net.to(dev)
rel = lambda a, b: float((a.double() - b.double()).abs().max() / b.double().abs().max())
for mode in ("fp32", "tf32x3"):
    ops.set_gemm_mode(mode)
    net.zero_grad()
    loss = ops.weighted_cross_entropy(net(bg.to(dev), feats.to(dev)), labels.to(dev), w.to(dev))
    loss.backward()
    print(mode, " ".join("%s=%.1e" % (n.replace("layers.", "L"), rel(p.grad.cpu(), q.grad)) for (n, p), (_, q) in zip(net.named_parameters(), ref64.named_parameters())))
