"""Full-size parity at the north-star tolerances (BASELINE configs[1] and [2]):

* GraphSAGE-pool 7x256, batch of 6 x 15k-node RAGs (90 000 nodes, 1.34 M edges), training step: logits within 1e-4
  (max-norm relative) of the CPU oracle, and every parameter gradient within 1e-4 (max-norm relative) of the oracle
  evaluated WITH THE GPU RUN'S DECISIONS (arg-max routing, ReLU masks) imposed — the stack is piecewise linear and a
  decision within rounding of a tie re-routes a whole contribution, so gradients are compared on the same routing and the
  number of decisions on which the two runs differ is counted and bounded separately.
* GAT 4-head x256 ([256]*4, heads [4,4,4,4], residuals [F,F,T,F]) on one 15k-node RAG: logits and gradients within 1e-4.
"""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from gnn_tumor_seg_b200 import graph as G, networks, ops, synth
from gnn_tumor_seg_b200.trainer import SageTrainer
from oracle import compare, gat_ref, graph_ref, sage_ref

pytestmark = pytest.mark.gpu
W = [0.1, 1.0, 2.0, 2.0]
TOL = 1e-4                    # north star: fp32 logits and gradients within 1e-4 relative (fp32 / 3xTF32 accumulate)
# Decisions allowed to differ between the two fp32 evaluations (inputs within rounding of a tie).  Measured on B200,
# tf32x3 against the fp32 CPU oracle: 2348 of 163 080 000 arg-max decisions (1.4e-5), 213 pool-ReLU and 145 of
# 161 280 000 output-ReLU decisions (~1e-6); the bounds leave ~3x headroom.
MAX_ARGMAX_FLIP_FRACTION = 5e-5
MAX_RELU_FLIP_FRACTION = 5e-6


def _rel_max(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


def _gpu_run_with_decisions(net, dg, x, y, w):
    return compare.gpu_sage_step_with_decisions(net, ops, dg, x, y, w)


@pytest.mark.parametrize("mode", ["tf32x3"])
def test_sage_7x256_batch6_training_step_full_size(cuda_dev, mode):
    ops.set_gemm_mode(mode)
    try:
        graphs = [synth.make_graph(s) for s in range(6)]
        bg = G.batch([G.from_edge_list(g.src, g.dst, g.n_nodes) for g in graphs])
        feats = torch.as_tensor(np.concatenate([g.features for g in graphs]))
        labels = torch.as_tensor(np.concatenate([g.labels for g in graphs]))
        assert bg.number_of_nodes() == 90000
        torch.manual_seed(0)
        net = networks.GraphSage(20, [256] * 7, 4, "pool", 0)
        ref = sage_ref.GraphSageRef(20, [256] * 7, 4)
        ref.load_state_dict(net.state_dict())
        net.to(cuda_dev)
        w = torch.tensor(W, device=cuda_dev)
        dg, x, y = bg.to(cuda_dev), feats.to(cuda_dev), labels.to(cuda_dev)
        logits, loss, grads, dec_gpu = _gpu_run_with_decisions(net, dg, x, y, w)

        indptr, indices = (t.cpu().numpy() for t in dg.csr)
        rl, dec_ref = sage_ref.graphsage_decisions(ref, (indptr, indices), feats)
        assert _rel_max(logits, rl) <= TOL
        assert (logits.argmax(1) == rl.argmax(1)).float().mean().item() >= 0.9999
        flips = sage_ref.count_decision_flips(dec_gpu, dec_ref)
        print("decision flips GPU vs oracle:", flips)
        n_dec = flips["argmax_decisions"]
        assert flips["argmax_flips"] <= MAX_ARGMAX_FLIP_FRACTION * n_dec, flips
        assert flips["pool_relu_flips"] <= MAX_RELU_FLIP_FRACTION * n_dec, flips
        assert flips["out_relu_flips"] <= MAX_RELU_FLIP_FRACTION * max(flips["out_relu_decisions"], 1), flips

        # oracle forward + backward on the GPU run's routing
        fl = sage_ref.graphsage_forward_forced(ref, feats, dec_gpu)
        rloss = F.cross_entropy(fl, labels, weight=torch.tensor(W))
        ref.zero_grad()
        rloss.backward()
        assert abs(rloss.item() - loss) <= TOL * abs(rloss.item())
        assert _rel_max(logits, fl.detach()) <= TOL
        worst = 0.0
        for n, q in ref.named_parameters():
            e = _rel_max(grads[n], q.grad)
            worst = max(worst, e)
            assert e <= TOL, (n, e)
        print("worst per-parameter gradient error (max-norm relative):", worst)

        # the product paths (whole-stack autograd Function, one-call trainer step) give the per-layer path's gradients
        for p in net.parameters():
            p.grad = None
        ops.weighted_cross_entropy(net(dg, x), y, w).backward()
        for n, p in net.named_parameters():
            assert _rel_max(p.grad.cpu(), grads[n]) <= 1e-5, n
        tr = SageTrainer(net, w, lr=1e-4, weight_decay=1e-4)
        loss_t = tr.forward_backward(dg, x, y)
        assert abs(float(loss_t) - loss) <= 1e-5 * abs(loss)
        for n, p in net.named_parameters():
            assert _rel_max(p.grad.cpu(), grads[n]) <= 1e-5, n
    finally:
        ops.set_gemm_mode("tf32x3")


@pytest.mark.parametrize("mode", ["fp32", "tf32x3"])
def test_gat_4head_x256_training_step_full_size(cuda_dev, mode):
    """BASELINE configs[2] on one 15k-node RAG (the CPU oracle materialises E x 256 temporaries per head).
    Yardstick = the oracle in fp64.  fp32 mode (SIMT FFMA GEMMs): every gradient within 1e-4 of it (or as close as the
    fp32 CPU oracle gets — the attention-vector gradients are sums over 15 000 x 4 x 256 products through the softmax
    backward's alpha * (dalpha - sum alpha dalpha) cancellation).  tf32x3 mode (tensor cores, the default): weights and
    biases within 1e-4; the attention vectors amplify the K = 1024 GEMMs' 2e-5 error through that cancellation and are
    held to the north star's TF32 bar, 1e-3 (measured: 4.4e-4 worst) — stated per mode, as BASELINE.json asks."""
    ops.set_gemm_mode(mode)
    try:
        g = synth.make_graph(1)
        bg = G.from_edge_list(g.src, g.dst, g.n_nodes)
        feats, labels = torch.as_tensor(g.features), torch.as_tensor(g.labels)
        cfg = (20, [256] * 4, 4, [4, 4, 4, 4], [False, False, True, False])
        torch.manual_seed(0)
        net = networks.GAT(*cfg)
        ref = gat_ref.GATRef(*cfg)
        ref.load_state_dict(net.state_dict())
        ref64 = gat_ref.GATRef(*cfg).double()
        ref64.load_state_dict({k: v.double() for k, v in net.state_dict().items()})
        net.to(cuda_dev)
        w = torch.tensor(W)
        logits = net(bg.to(cuda_dev), feats.to(cuda_dev))
        loss = ops.weighted_cross_entropy(logits, labels.to(cuda_dev), w.to(cuda_dev))
        loss.backward()
        s, d = bg.edges()
        rl = ref((s, d), feats)
        rloss = F.cross_entropy(rl, labels, weight=w)
        rloss.backward()
        rl64 = ref64((s, d), feats.double())
        F.cross_entropy(rl64, labels, weight=w.double()).backward()
        assert _rel_max(logits.detach().cpu(), rl.detach()) <= TOL
        assert _rel_max(logits.detach().cpu(), rl64.detach()) <= TOL
        assert abs(loss.item() - rloss.item()) <= TOL * abs(rloss.item())
        worst = worst_cpu = 0.0
        for (n, p), (_, q), (_, q64) in zip(net.named_parameters(), ref.named_parameters(), ref64.named_parameters()):
            e_gpu = _rel_max(p.grad.cpu(), q64.grad)
            e_cpu = _rel_max(q.grad, q64.grad)
            worst, worst_cpu = max(worst, e_gpu), max(worst_cpu, e_cpu)
            tol = 1e-3 if (mode == "tf32x3" and "attn" in n) else TOL
            assert e_gpu <= max(tol, 2.0 * e_cpu), (mode, n, e_gpu, e_cpu)
        print("GAT [%s] worst per-parameter gradient error vs fp64: GPU %.2e, fp32 CPU oracle %.2e" % (mode, worst, worst_cpu))
    finally:
        ops.set_gemm_mode("tf32x3")
