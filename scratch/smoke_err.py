"""Per-parameter gradient errors of the smoke() configuration (two small graphs, [256, 256] hidden) against the oracle."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gnn_tumor_seg_b200 import graph as G, networks, ops, synth
from oracle import graph_ref, sage_ref
dev = torch.device("cuda:0")
for seed in range(4):
    torch.manual_seed(seed)
    graphs = [synth.make_small_graph(s + 10 * seed, n_nodes=300 + 50 * s, avg_deg=8) for s in range(2)]
    bg = G.batch([G.from_edge_list(g.src, g.dst, g.n_nodes) for g in graphs])
    feats = torch.as_tensor(np.concatenate([g.features for g in graphs]))
    labels = torch.as_tensor(np.concatenate([g.labels for g in graphs]))
    w = torch.tensor([0.1, 1.0, 2.0, 2.0])
    net = networks.GraphSage(20, [256, 256], 4, "pool", 0)
    ref = sage_ref.GraphSageRef(20, [256, 256], 4)
    ref.load_state_dict(net.state_dict())
    net.to(dev)
    s, d = bg.edges()
    indptr, indices, _ = graph_ref.csr_by_dst_ref(s.numpy(), d.numpy(), bg.number_of_nodes())
    ref.zero_grad()
    rl = ref((indptr, indices), feats)
    torch.nn.functional.cross_entropy(rl, labels, weight=w).backward()
    for mode in ("fp32", "tf32x3"):
        ops.set_gemm_mode(mode)
        net.zero_grad()
        logits = net(bg.to(dev), feats.to(dev))
        ops.weighted_cross_entropy(logits, labels.to(dev), w.to(dev)).backward()
        le = (logits.detach().cpu() - rl.detach()).abs().max().item() / rl.detach().abs().max().item()
        worst_max, worst_norm = 0.0, 0.0
        for (n, p), (_, q) in zip(net.named_parameters(), ref.named_parameters()):
            g = p.grad.cpu()
            worst_max = max(worst_max, (g - q.grad).abs().max().item() / max(q.grad.abs().max().item(), 1e-12))
            worst_norm = max(worst_norm, (g - q.grad).norm().item() / max(q.grad.norm().item(), 1e-12))
        print("seed %d %-6s bf=%s logits %.2e  grad worst max-rel %.2e  worst norm-rel %.2e" %
              (seed, mode, os.environ.get("GTS_X3_BF16", "1"), le, worst_max, worst_norm), flush=True)
