import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.nn.functional as F
from gnn_tumor_seg_b200 import graph as G, networks, ops, synth
from oracle import sage_ref
dev = torch.device("cuda:0")
g = synth.make_graph(0)
torch.manual_seed(0)
net = networks.GraphSage(20, [256] * 7, 4, "pool", 0)
ref = sage_ref.GraphSageRef(20, [256] * 7, 4).double()
ref.load_state_dict({k: v.double() for k, v in net.state_dict().items()})
net.to(dev)
bg = G.from_edge_list(g.src, g.dst, g.n_nodes).to(dev)
feats = torch.as_tensor(g.features); labels = torch.as_tensor(g.labels); w = torch.tensor([0.1, 1, 2, 2.])
indptr, indices = (t.cpu().numpy() for t in bg.csr)
rl = ref((indptr, indices), feats.double())
rloss = F.cross_entropy(rl, labels, weight=w.double()); rloss.backward()
rg = torch.cat([p.grad.reshape(-1) for p in ref.parameters()])
print("ref logits absmax", rl.abs().max().item(), "loss", rloss.item())
for mode in ("fp32", "tf32x3", "tf32"):
    ops.set_gemm_mode(mode)
    net.zero_grad()
    logits = net(bg, feats.to(dev))
    loss = ops.weighted_cross_entropy(logits, labels.to(dev), w.to(dev)); loss.backward()
    gg = torch.cat([p.grad.reshape(-1) for p in net.parameters()]).cpu().double()
    l = logits.detach().cpu().double()
    print(f"{mode:7s} logits max-rel {((l-rl).abs().max()/rl.abs().max()).item():.3e}  norm-rel {((l-rl).norm()/rl.norm()).item():.3e} "
          f"class agree {(l.argmax(1)==rl.argmax(1)).double().mean().item():.6f}  loss rel {abs(loss.item()-rloss.item())/rloss.item():.3e} "
          f"grad norm-rel {((gg-rg).norm()/rg.norm()).item():.3e}")
