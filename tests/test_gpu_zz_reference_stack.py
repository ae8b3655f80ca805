"""GPU parity against the committed fixture of the reference-built stacks (tests/golden/reference_stack.npz, see
tests/test_reference_stack.py): the product's GraphSage / GAT load the reference-built checkpoints and reproduce the
logits within the north star's 1e-4 (fp32-accurate modes).  Named to run after every other GPU test file."""
import os

import numpy as np
import pytest
import torch

from gnn_tumor_seg_b200 import graph as G, networks, ops

pytestmark = pytest.mark.gpu
GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_stack.npz"))


def _state(prefix):
    return {k[len(prefix):]: torch.as_tensor(GOLD[k]) for k in GOLD.files if k.startswith(prefix)}


def _rel(a, b):
    return (a.double() - b.double()).abs().max().item() / max(b.double().abs().max().item(), 1e-30)


@pytest.mark.parametrize("mode", ["fp32", "tf32x3"])
@pytest.mark.parametrize("model", ["sage", "gat"])
def test_product_reproduces_reference_built_stack(cuda_dev, model, mode):
    ops.set_gemm_mode(mode)
    try:
        if model == "sage":
            net = networks.GraphSage(20, [64, 32, 128], 4, "pool", 0.0)
        else:
            net = networks.GAT(20, [16, 8, 32], 4, [2, 4, 2], [True, True, True])
        net.load_state_dict(_state(model + "_sd/"))
        net.to(cuda_dev)
        bg = G.from_edge_list(GOLD["src"], GOLD["dst"], int(GOLD["n_nodes"])).to(cuda_dev)
        x = torch.as_tensor(GOLD["x"]).to(cuda_dev)
        want = torch.as_tensor(GOLD[model + "_logits"])
        got_train = net(bg, x).detach().cpu()                      # autograd (training) forward
        net.eval()
        with torch.no_grad():
            got_eval = net(bg, x).cpu()                            # inference forward
        assert _rel(got_train, want) < 1e-4, _rel(got_train, want)
        assert _rel(got_eval, want) < 1e-4, _rel(got_eval, want)
    finally:
        ops.set_gemm_mode("tf32x3")
