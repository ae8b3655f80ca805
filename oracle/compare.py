"""Tie-aware comparison helpers (TEST INFRASTRUCTURE, see oracle/__init__.py): run the CUDA path with its decisions
recorded, so that the oracle can be evaluated on the same routing (oracle/sage_ref.py: graphsage_forward_forced).
Imported by tests/ and __graft_entry__.smoke() only."""
import torch


def gpu_sage_step_with_decisions(net, ops, dg, x, y, w):
    """Forward + weighted CE + backward of a GraphSage('pool') net through the per-layer CUDA path, recording every
    decision: arg-max and (neigh > 0) from each segmax_fwd call, (out > 0) of each hidden layer.
    Returns (logits cpu, loss float, {param name: grad cpu}, decisions)."""
    rec = {"seg": [], "out": []}
    orig = ops.segmax_fwd

    def seg(P, indptr, indices, want_argmax=True):
        neigh, arg = orig(P, indptr, indices, want_argmax=want_argmax)
        rec["seg"].append((arg.cpu().long(), (neigh > 0).cpu()))
        return neigh, arg

    hooks = [l.register_forward_hook(lambda m, i, o: rec["out"].append((o > 0).cpu())) for l in net.layers]
    stack = ops.use_stack_path()
    ops.set_stack_path(False)
    ops.segmax_fwd = seg
    try:
        for p in net.parameters():
            p.grad = None
        logits = net(dg, x)
        loss = ops.weighted_cross_entropy(logits, y, w)
        loss.backward()
    finally:
        ops.segmax_fwd = orig
        ops.set_stack_path(stack)
        for h in hooks:
            h.remove()
    L = len(net.layers)
    decisions = [{"arg": rec["seg"][l][0], "neigh_pos": rec["seg"][l][1],
                  "out_pos": rec["out"][l] if l + 1 < L else None} for l in range(L)]
    grads = {n: p.grad.detach().cpu().clone() for n, p in net.named_parameters()}
    return logits.detach().cpu(), float(loss.detach()), grads, decisions
