"""On-hardware data-parallel correctness: 2 ranks under NCCL, whole graphs per rank,
trainer.SageTrainer's bucketed all-reduce overlapped with the backward.  After the step the gradient arenas are
BITWISE equal across ranks and agree with the single-device step on the union batch to 1e-5 (norm-wise; the
all-reduce and the union batch sum the same per-graph contributions in a different order).
Needs >= 2 GPUs (gpurun --gpus 2); skipped otherwise."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
W = [0.1, 1.0, 2.0, 2.0]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    import torch.distributed as dist
    from gnn_tumor_seg_b200 import graph as G, networks, ops, synth
    from gnn_tumor_seg_b200.trainer import SageTrainer
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", device_id=dev)
    ops.set_deterministic_backward(True)
    graphs = [synth.make_small_graph(70 + s, n_nodes=500 + 31 * s, avg_deg=8) for s in range(4)]

    def batch_of(ids):
        sel = [graphs[i] for i in ids]
        bg = G.batch([G.from_edge_list(g.src, g.dst, g.n_nodes) for g in sel])
        return (bg.to(dev), torch.as_tensor(np.concatenate([g.features for g in sel])).to(dev),
                torch.as_tensor(np.concatenate([g.labels for g in sel])).to(dev))

    w = torch.tensor(W, device=dev)
    torch.manual_seed(0)
    net = networks.GraphSage(20, [256, 256, 64], 4, "pool", 0).to(dev)
    tr = SageTrainer(net, w, lr=1e-3, weight_decay=1e-4, n_buckets=2)
    assert tr.world_size == world and len(tr.buckets) == 2
    loss = tr.forward_backward(*batch_of(list(range(rank, 4, world))))        # rank r owns graphs r, r+R, ...
    torch.cuda.synchronize()
    n = tr.arena.total
    flat = tr.arena.grads[:n + 2].clone()
    gathered = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    bitwise = all(torch.equal(gathered[0], g) for g in gathered[1:])
    # global gradient = reduced sums / global denominator
    g_dp = (flat[:n] / flat[n + 1]).cpu()
    # the same step on ONE device over the union batch (local trainer on an identical network)
    torch.manual_seed(0)
    net1 = networks.GraphSage(20, [256, 256, 64], 4, "pool", 0).to(dev)
    tr1 = SageTrainer(net1, w, lr=1e-3, weight_decay=1e-4, data_parallel=False)
    loss1 = tr1.forward_backward(*batch_of([0, 1, 2, 3]))
    torch.cuda.synchronize()
    g_1 = tr1.arena.grads[:n].cpu()
    rel = float((g_dp - g_1).norm() / g_1.norm())
    # then one optimiser step each: parameters agree as well
    tr.optimizer.step(grad_denom=tr.denominator)
    tr1.optimizer.step()
    torch.cuda.synchronize()
    prel = float((tr.arena.params - tr1.arena.params).abs().max())
    # the same data-parallel step replayed from CUDA-graph segments (trainer.GraphedStep) vs eager: 3 steps each
    from gnn_tumor_seg_b200.trainer import GraphedStep
    sel = [graphs[i] for i in range(rank, 4, world)]
    hg = G.batch([G.from_edge_list(g.src, g.dst, g.n_nodes) for g in sel], pin=True)
    hx = torch.as_tensor(np.concatenate([g.features for g in sel])).pin_memory()
    hy = torch.as_tensor(np.concatenate([g.labels for g in sel])).pin_memory()
    nets = []
    for _ in range(2):
        torch.manual_seed(5)
        nets.append(networks.GraphSage(20, [256, 256, 64], 4, "pool", 0).to(dev))
    tr_e = SageTrainer(nets[0], w, lr=1e-3, weight_decay=1e-4)
    tr_g = SageTrainer(nets[1], w, lr=1e-3, weight_decay=1e-4)
    gs = GraphedStep(tr_g, hg, hx, hy)                      # one eager warm-up step + capture of the segments
    assert gs.segments is not None and len(gs.segments) == len(tr_g.buckets) + 1
    le = lg = None
    for it in range(3):
        le = tr_e.step(hg.to(dev), hx.to(dev), hy.to(dev))
        if it > 0:
            lg = gs(hg, hx, hy)
    torch.cuda.synchronize()
    gdiff = float((tr_e.arena.params - tr_g.arena.params).abs().mean())
    torch.save({"bitwise": bitwise, "rel": rel, "loss": float(loss), "loss1": float(loss1), "prel": prel,
                "graph_param_diff": gdiff, "loss_eager": float(le), "loss_graph": float(lg)},
               os.path.join(out_dir, f"r{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs 2 GPUs (gpurun --gpus 2)")
def test_two_rank_nccl_step_equals_single_device_union_batch(tmp_path):
    import torch.multiprocessing as mp
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    for r in range(2):
        d = torch.load(os.path.join(str(tmp_path), f"r{r}.pt"))
        assert d["bitwise"], "gradient arenas differ across ranks after the all-reduce"
        assert d["rel"] <= 1e-5, d
        assert abs(d["loss"] - d["loss1"]) <= 1e-5 * abs(d["loss1"]), d
        assert d["prel"] <= 2e-4, d          # |lr| = 1e-3: a first AdamW step moves every element by ~lr
        assert d["graph_param_diff"] < 1e-6 and abs(d["loss_eager"] - d["loss_graph"]) <= 1e-5 * abs(d["loss_eager"]), d
