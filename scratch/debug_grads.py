import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.nn.functional as F
from gnn_tumor_seg_b200 import graph as G, networks, ops, synth
from oracle import graph_ref, sage_ref
dev = torch.device("cuda:0")
mode = sys.argv[1] if len(sys.argv) > 1 else "fp32"
ops.set_gemm_mode(mode)
gs = [synth.make_small_graph(s, n_nodes=400 + 17 * i, isolated=3) for i, s in enumerate([1, 2, 3])]
bg = G.batch([G.from_edge_list(g.src, g.dst, g.n_nodes) for g in gs])
feats = torch.as_tensor(np.concatenate([g.features for g in gs])); labels = torch.as_tensor(np.concatenate([g.labels for g in gs]))
s, d = bg.edges(); csr = graph_ref.csr_by_dst_ref(s.numpy(), d.numpy(), bg.number_of_nodes())[:2]
torch.manual_seed(0)
net = networks.GraphSage(20, [256, 256, 64], 4, "pool", 0); ref = sage_ref.GraphSageRef(20, [256, 256, 64], 4)
ref.load_state_dict(net.state_dict()); net.to(dev)
w = torch.tensor([0.1, 1., 2., 2.])
x = feats.to(dev).requires_grad_(True)
logits = net(bg.to(dev), x); loss = ops.weighted_cross_entropy(logits, labels.to(dev), w.to(dev)); loss.backward()
xr = feats.clone().requires_grad_(True); rl = ref(csr, xr); rloss = F.cross_entropy(rl, labels, weight=w); rloss.backward()
print("loss", loss.item(), rloss.item(), "logit max", rl.abs().max().item())
def rn(a, b): return ((a.double()-b.double()).norm()/b.double().norm()).item()
def rm(a, b): return ((a.double()-b.double()).abs().max()/b.double().abs().max()).item()
print("logits", rn(logits.detach().cpu(), rl.detach()), rm(logits.detach().cpu(), rl.detach()))
for (n, p), (_, q) in zip(net.named_parameters(), ref.named_parameters()):
    print(f"{n:28s} norm-rel {rn(p.grad.cpu(), q.grad):.3e} max-rel {rm(p.grad.cpu(), q.grad):.3e}  |ref| {q.grad.norm().item():.3e}")
print("x.grad", rn(x.grad.cpu(), xr.grad), rm(x.grad.cpu(), xr.grad))
# same thing but the oracle in float64 as a third opinion
ref64 = sage_ref.GraphSageRef(20, [256, 256, 64], 4).double(); ref64.load_state_dict({k: v.double() for k, v in ref.state_dict().items()})
rl64 = ref64(csr, feats.double()); l64 = F.cross_entropy(rl64, labels, weight=w.double()); l64.backward()
for (n, p), (_, q), (_, r) in zip(net.named_parameters(), ref.named_parameters(), ref64.named_parameters()):
    print(f"{n:28s} gpu-vs-f64 {rn(p.grad.cpu(), r.grad):.3e}   cpu32-vs-f64 {rn(q.grad, r.grad):.3e}")
