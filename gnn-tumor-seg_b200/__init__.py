"""gnn-tumor-seg_b200 — B200-native message-passing hot path of GNN-Tumor-Seg.

Host-side mirror of the reference's module API (model/networks.py,
model/gnn_model.py, data_processing/data_loader.py, data_processing/graph_io.py)
over hand-written sm_100a CUDA kernels reached through the C-ABI library
``libgts.so`` (include/gts.h).  Import as ``gnn_tumor_seg_b200``.

Sub-modules are imported lazily so that CPU-only tooling (the synthetic graph
generator, the data-parallel host logic) works without loading the CUDA
library; anything that computes goes through ``_lib`` and fails loudly if the
library or a GPU is missing — there is no CPU fallback.
"""
__version__ = "0.1.0"

_LAZY = {
    "GraphSage": "networks", "GAT": "networks", "init_graph_net": "networks",
    "SAGEConv": "networks", "GATConv": "networks",
    "GNN": "gnn_model",
    "BatchedGraph": "graph", "minibatch_graphs": "graph", "from_networkx": "graph",
    "batch": "graph", "from_edge_list": "graph",
    "project_nodes_to_img": "project", "project_labels_to_brats": "project",
    "project_logits_to_img": "project", "determine_tumor_crop": "project",
    "ImageGraphDataset": "data_loader", "PredLogitDataset": "data_loader",
    "load_graph_json": "graph_io", "parse_node_link_json": "graph_io",
}


def __getattr__(name):
    if name in _LAZY:
        import importlib
        mod = importlib.import_module(f"{__name__}.{_LAZY[name]}")
        return getattr(mod, name)
    raise AttributeError(f"module {__name__!r} has no attribute {name!r}")
