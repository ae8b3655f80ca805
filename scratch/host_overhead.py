import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gnn_tumor_seg_b200 import graph as G, networks, ops, synth, dp
dev = torch.device("cuda:0")
ops.set_gemm_mode(sys.argv[1] if len(sys.argv) > 1 else "tf32")
graphs = [synth.make_graph(s) for s in range(6)]
bg = G.batch([G.from_edge_list(g.src, g.dst, g.n_nodes) for g in graphs]).to(dev)
feats = torch.as_tensor(np.concatenate([g.features for g in graphs])).to(dev)
labels = torch.as_tensor(np.concatenate([g.labels for g in graphs])).to(dev)
torch.manual_seed(0)
net = networks.GraphSage(20, [256]*7, 4, "pool", 0).to(dev)
tr = dp.DataParallelTrainer(net, torch.tensor([0.1,1,2,2.], device=dev))
for _ in range(3): tr.forward_backward(bg, feats, labels)
torch.cuda.synchronize()
for rep in range(2):
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): tr.forward_backward(bg, feats, labels)
    e1.record()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"host enqueue {1e3*(t1-t0)/10:.3f} ms/step, total wall {1e3*(t2-t0)/10:.3f}, gpu events {e0.elapsed_time(e1)/10:.3f}")
# CUDA graph of one step
g = torch.cuda.CUDAGraph()
s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    for _ in range(2): tr.forward_backward(bg, feats, labels)
torch.cuda.current_stream().wait_stream(s)
try:
    with torch.cuda.graph(g):
        loss = tr.forward_backward(bg, feats, labels)
    torch.cuda.synchronize()
    for _ in range(3): g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): g.replay()
    e1.record(); torch.cuda.synchronize()
    print(f"cuda graph replay {e0.elapsed_time(e1)/10:.3f} ms/step  loss {float(loss):.4f}")
except Exception as ex:
    print("graph capture failed:", repr(ex)[:300])
