"""ctypes binding of libgts.so (C-ABI declared in include/gts.h).

The library is built in-tree by ``__graft_entry__.build()`` (or
``make -C gnn-tumor-seg_b200/csrc``).  There is deliberately no fallback: if the
shared object is missing or a CUDA device is absent, every compute entry point
raises — the product never computes on the CPU.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GTS_LIB_PATH") or os.path.join(_HERE, "libgts.so")     # override: A/B builds of the library

GTS_OK = 0
ACT_NONE, ACT_RELU, ACT_MASK_POS, ACT_MASK_POS_SCATTER, ACT_MASK_BITS, ACT_MASK_BITS_SCATTER = 0, 1, 2, 3, 4, 5
GEMM_FP32, GEMM_TF32, GEMM_TF32X3 = 0, 1, 2
GEMM_MODES = {"fp32": GEMM_FP32, "tf32": GEMM_TF32, "tf32x3": GEMM_TF32X3}

c_f32p = C.c_void_p
c_i32p = C.c_void_p
c_i64p = C.c_void_p
c_i16p = C.c_void_p
c_stream = C.c_void_p


class GemmNtArgs(C.Structure):
    """struct gts_gemm_nt_args (include/gts.h)."""
    _fields_ = [
        ("A1", C.c_void_p), ("lda1", C.c_int64), ("K1", C.c_int32),
        ("A2", C.c_void_p), ("lda2", C.c_int64), ("K2", C.c_int32),
        ("B1", C.c_void_p), ("ldb1", C.c_int64),
        ("B2", C.c_void_p), ("ldb2", C.c_int64),
        ("bias", C.c_void_p),
        ("aux", C.c_void_p), ("ldaux", C.c_int64),
        ("C", C.c_void_p), ("ldc", C.c_int64),
        ("M", C.c_int32), ("N", C.c_int32),
        ("act", C.c_int32), ("mode", C.c_int32),
        ("bias2", C.c_void_p),
        ("scatter_idx", C.c_void_p), ("ld_idx", C.c_int64),
        ("scatter_out", C.c_void_p), ("ld_out", C.c_int64),
        ("relu_bits_out", C.c_void_p), ("ld_bits_out", C.c_int64),
        ("aux_bits", C.c_void_p), ("ld_aux_bits", C.c_int64),
        ("zero_fill", C.c_void_p), ("zero_fill_bytes", C.c_size_t),
    ]


class SageLayer(C.Structure):
    """struct gts_sage_layer (include/gts.h)."""
    _fields_ = [("din", C.c_int32), ("dout", C.c_int32), ("relu", C.c_int32), ("reserved", C.c_int32),
                ("Wp", C.c_void_p), ("bp", C.c_void_p), ("Ws", C.c_void_p), ("Wn", C.c_void_p), ("b", C.c_void_p),
                ("b2", C.c_void_p)]


class SageLayerGrads(C.Structure):
    """struct gts_sage_layer_grads (include/gts.h)."""
    _fields_ = [("dWp", C.c_void_p), ("dbp", C.c_void_p), ("dWs", C.c_void_p), ("dWn", C.c_void_p), ("db", C.c_void_p),
                ("db2", C.c_void_p)]


class SageStepArgs(C.Structure):
    """struct gts_sage_step_args (include/gts.h)."""
    _fields_ = [("layers", C.POINTER(SageLayer)), ("grads", C.POINTER(SageLayerGrads)), ("n_layers", C.c_int32),
                ("n_nodes", C.c_int32),
                ("indptr", C.c_void_p), ("indices", C.c_void_p), ("csc_indptr", C.c_void_p), ("csc_indices", C.c_void_p),
                ("feats", C.c_void_p), ("ldf", C.c_int64),
                ("labels", C.c_void_p), ("class_w", C.c_void_p),
                ("sums", C.c_void_p),
                ("logits", C.c_void_p), ("ldl", C.c_int64),
                ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t),
                ("mode", C.c_int32), ("normalize", C.c_int32), ("bwd_layer_lo", C.c_int32), ("reserved", C.c_int32)]


MAX_PEERS = 16
PEER_HANDLE_BYTES = 64


class PeerComm(C.Structure):
    """struct gts_peer_comm (include/gts.h)."""
    _fields_ = [("rank", C.c_int32), ("world", C.c_int32), ("n", C.c_int64), ("base", C.c_void_p * MAX_PEERS)]


# name -> (restype, argtypes); mirrors include/gts.h one to one
_SIGNATURES = {
    "gts_version": (C.c_int, []),
    "gts_last_error": (C.c_char_p, []),
    "gts_device_info": (C.c_int, [C.POINTER(C.c_int32)] * 3),
    "gts_batch_edges": (C.c_int, [c_i32p, c_i32p, C.c_int64, c_i64p, c_i32p, C.c_int32, c_i32p, c_i32p, c_stream]),
    "gts_csr_build_workspace_bytes": (C.c_size_t, [C.c_int64, C.c_int32]),
    "gts_csr_build": (C.c_int, [c_i32p, c_i32p, C.c_int64, C.c_int32, c_i32p, c_i32p, c_i32p,
                                C.c_void_p, C.c_size_t, c_stream]),
    "gts_edge_perm_compose": (C.c_int, [c_i32p, c_i32p, C.c_int64, c_i32p, c_i32p, c_stream]),
    "gts_gemm_nt": (C.c_int, [C.POINTER(GemmNtArgs), c_stream]),
    "gts_gemm_nt_bits_supported": (C.c_int, [C.c_int32, C.c_int32, C.c_int32]),
    "gts_gemm_tn_workspace_bytes": (C.c_size_t, [C.c_int32, C.c_int32, C.c_int64, C.c_int32]),
    "gts_gemm_tn": (C.c_int, [c_f32p, C.c_int64, c_f32p, C.c_int64, c_f32p, C.c_int64, C.c_int32, C.c_int32,
                              C.c_int64, C.c_int32, C.c_void_p, C.c_size_t, c_stream]),
    "gts_gemm_tn_colsum_workspace_bytes": (C.c_size_t, [C.c_int32, C.c_int32, C.c_int64, C.c_int32]),
    "gts_gemm_tn_colsum": (C.c_int, [c_f32p, C.c_int64, c_f32p, C.c_int64, c_f32p, C.c_int64, C.c_int32, C.c_int32,
                                     C.c_int64, C.c_int32, c_f32p, C.c_void_p, C.c_size_t, c_stream]),
    "gts_gemm_tn2_colsum_workspace_bytes": (C.c_size_t, [C.c_int32, C.c_int32, C.c_int64, C.c_int32]),
    "gts_gemm_tn2_colsum": (C.c_int, [c_f32p, C.c_int64, c_f32p, C.c_int64, c_f32p, C.c_int64, c_f32p, c_f32p, C.c_int64,
                                      C.c_int32, C.c_int32, C.c_int64, C.c_int32, c_f32p, C.c_void_p, C.c_size_t, c_stream]),
    "gts_colsum_workspace_bytes": (C.c_size_t, [C.c_int64, C.c_int32]),
    "gts_colsum": (C.c_int, [c_f32p, C.c_int64, C.c_int64, C.c_int32, c_f32p, C.c_void_p, C.c_size_t, c_stream]),
    "gts_transpose": (C.c_int, [c_f32p, C.c_int64, C.c_int32, C.c_int32, c_f32p, C.c_int64, c_stream]),
    "gts_segmax_fwd": (C.c_int, [c_f32p, C.c_int64, c_i32p, c_i32p, C.c_int32, C.c_int32, c_f32p, C.c_int64,
                                 c_i32p, C.c_int64, c_stream]),
    "gts_segmax_fwd_bits_supported": (C.c_int, [C.c_int32, C.c_int32, C.c_int64]),
    "gts_segmax_fwd_bits": (C.c_int, [c_f32p, C.c_int64, c_i32p, c_i32p, C.c_int32, C.c_int32, c_f32p, C.c_int64,
                                      c_i32p, C.c_int64, C.c_void_p, C.c_int64, c_stream]),
    "gts_segmax_bwd": (C.c_int, [c_f32p, C.c_int64, c_i32p, C.c_int64, C.c_int32, C.c_int32, c_f32p, C.c_int64,
                                 C.c_int32, c_stream]),
    "gts_segmax_bwd_add": (C.c_int, [c_f32p, C.c_int64, c_i32p, C.c_int64, C.c_int32, C.c_int32, c_f32p, C.c_int64,
                                     c_stream]),
    "gts_segmax_bwd_det": (C.c_int, [c_f32p, C.c_int64, c_i32p, C.c_int64, c_i32p, c_i32p, C.c_int32, C.c_int32,
                                     c_f32p, C.c_int64, c_stream]),
    "gts_segsum_fwd": (C.c_int, [c_f32p, C.c_int64, c_i32p, c_i32p, C.c_int32, C.c_int32, C.c_int32, c_f32p,
                                 C.c_int64, c_stream]),
    "gts_segsum_bwd": (C.c_int, [c_f32p, C.c_int64, c_i32p, c_i32p, c_i32p, C.c_int32, C.c_int32, C.c_int32,
                                 c_f32p, C.c_int64, c_stream]),
    "gts_mask_pos": (C.c_int, [c_f32p, c_f32p, C.c_int64, c_f32p, c_stream]),
    "gts_sage_workspace_bytes": (C.c_size_t, [C.POINTER(SageLayer), C.c_int32, C.c_int32, C.c_int32, C.c_int32]),
    "gts_sage_forward": (C.c_int, [C.POINTER(SageLayer), C.c_int32, c_i32p, c_i32p, C.c_int32, c_f32p, C.c_int64,
                                   c_f32p, C.c_int64, C.c_void_p, C.c_size_t, C.c_int32, C.c_int32, c_stream]),
    "gts_sage_backward": (C.c_int, [C.POINTER(SageLayer), C.POINTER(SageLayerGrads), C.c_int32, c_i32p, c_i32p,
                                    C.c_int32, c_f32p, C.c_int64, c_f32p, C.c_int64, c_f32p, C.c_int64,
                                    C.c_void_p, C.c_size_t, C.c_int32, c_stream]),
    "gts_sage_backward_range": (C.c_int, [C.POINTER(SageLayer), C.POINTER(SageLayerGrads), C.c_int32, C.c_int32, C.c_int32,
                                          c_i32p, c_i32p, C.c_int32, c_f32p, C.c_int64, c_f32p, C.c_int64, c_f32p,
                                          C.c_int64, C.c_void_p, C.c_size_t, C.c_int32, c_stream]),
    "gts_sage_step": (C.c_int, [C.POINTER(SageStepArgs), c_stream]),
    "gts_sage_step_backward_rest": (C.c_int, [C.POINTER(SageStepArgs), C.c_int32, C.c_int32, c_stream]),
    "gts_sage_profile": (C.c_int, [C.c_int32]),
    "gts_sage_profile_read": (C.c_int, [C.POINTER(C.c_float), C.POINTER(C.c_int32), C.c_int32]),
    "gts_ce_weighted": (C.c_int, [c_f32p, C.c_int64, c_i64p, c_f32p, C.c_int32, C.c_int32, c_f32p, c_f32p,
                                  C.c_int64, c_stream]),
    "gts_scale_by_inv": (C.c_int, [c_f32p, C.c_int64, C.c_float, c_f32p, c_stream]),
    "gts_argmax_rows": (C.c_int, [c_f32p, C.c_int64, C.c_int32, C.c_int32, c_i32p, c_stream]),
    "gts_project_nodes": (C.c_int, [c_i16p, C.c_int64, c_i64p, C.c_int32, c_i64p, c_i32p, c_stream]),
    "gts_project_labels": (C.c_int, [c_i16p, C.c_int32, C.c_int32, C.c_int32, c_i32p, c_i32p, c_i32p, c_i32p,
                                     C.c_int32, c_i16p, C.c_int32, c_i16p, C.c_int32, C.c_int32, C.c_int32,
                                     c_i32p, c_stream]),
    "gts_tumor_plane_occupancy": (C.c_int, [c_i16p, C.c_int32, C.c_int32, C.c_int32, c_i32p, C.c_int32, c_i32p, c_i32p,
                                            c_i32p, c_i32p, c_stream]),
    "gts_project_logits": (C.c_int, [c_i16p, C.c_int64, c_f32p, C.c_int64, C.c_int32, C.c_int32, c_f32p, c_f32p,
                                     c_i32p, c_stream]),
    "gts_gat_scores": (C.c_int, [c_f32p, C.c_int64, c_f32p, c_f32p, C.c_int32, C.c_int32, C.c_int32, c_f32p,
                                 c_f32p, c_stream]),
    "gts_gat_fwd": (C.c_int, [c_f32p, C.c_int64, c_f32p, c_f32p, c_i32p, c_i32p, C.c_int32, C.c_int32, C.c_int32,
                              C.c_float, c_f32p, C.c_int64, c_f32p, C.c_int32, c_f32p, C.c_int64, c_f32p, c_f32p,
                              c_i32p, c_stream]),
    "gts_gat_act_bwd": (C.c_int, [c_f32p, c_f32p, C.c_int64, C.c_int32, c_f32p, c_stream]),
    "gts_gat_bwd_dst": (C.c_int, [c_f32p, C.c_int64, c_f32p, c_f32p, c_f32p, c_f32p, c_i32p, c_i32p, c_f32p,
                                  C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_float, c_f32p, c_f32p, c_stream]),
    "gts_gat_bwd_src": (C.c_int, [c_f32p, c_f32p, c_f32p, c_f32p, c_i32p, c_i32p, c_i32p, c_f32p, C.c_int64,
                                  c_f32p, c_f32p, c_f32p, c_f32p, C.c_int32, C.c_int32, C.c_int32, C.c_float,
                                  c_f32p, C.c_int64, c_f32p, c_stream]),
    "gts_gat_attn_grad_workspace_bytes": (C.c_size_t, [C.c_int32, C.c_int32, C.c_int32]),
    "gts_gat_attn_grad": (C.c_int, [c_f32p, C.c_int64, c_f32p, C.c_int32, C.c_int32, C.c_int32, c_f32p,
                                    C.c_void_p, C.c_size_t, c_stream]),
    "gts_gat_attn_grad2": (C.c_int, [c_f32p, C.c_int64, c_f32p, c_f32p, C.c_int32, C.c_int32, C.c_int32, c_f32p, c_f32p,
                                     C.c_void_p, C.c_size_t, c_stream]),
    "gts_adamw_step": (C.c_int, [c_f32p, c_f32p, c_f32p, c_f32p, C.c_int64, C.c_float, C.c_float, C.c_float,
                                 C.c_float, C.c_float, C.c_int32, C.c_float, c_f32p, c_stream]),
    "gts_adamw_step_dev": (C.c_int, [c_f32p, c_f32p, c_f32p, c_f32p, C.c_int64, c_f32p, C.c_float, c_f32p, c_stream]),
    "gts_peer_buffer_bytes": (C.c_size_t, [C.c_int64]),
    "gts_peer_alloc": (C.c_int, [C.c_size_t, C.POINTER(C.c_void_p), C.c_char_p]),
    "gts_peer_free": (C.c_int, [C.c_void_p]),
    "gts_peer_open": (C.c_int, [C.c_char_p, C.POINTER(C.c_void_p)]),
    "gts_peer_close": (C.c_int, [C.c_void_p]),
    "gts_peer_publish": (C.c_int, [C.POINTER(PeerComm), c_f32p, c_stream]),
    "gts_peer_allreduce_adamw": (C.c_int, [C.POINTER(PeerComm), c_f32p, C.c_int64, c_f32p, c_f32p, c_f32p, c_f32p,
                                           C.c_int64, C.c_int32, c_stream]),
    "gts_peer_status": (C.c_int, [C.POINTER(PeerComm), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)

_lib = None


class GtsError(RuntimeError):
    """A libgts.so call returned a non-zero status."""


def lib_available() -> bool:
    return os.path.exists(LIB_PATH)


def load():
    """Load libgts.so (once) and bind every symbol of include/gts.h."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise GtsError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            f"or `make -C {os.path.join(_HERE, 'csrc')}`. There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (restype, argtypes) in _SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError here = header/library mismatch
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


def check(rc: int, what: str = "") -> None:
    if rc != GTS_OK:
        msg = load().gts_last_error()
        raise GtsError(f"{what or 'libgts'} failed (status {rc}): {msg.decode() if msg else ''}")


def require_cuda(*tensors) -> None:
    """Fail loudly if a tensor is not on a CUDA device (no CPU path exists)."""
    import torch
    cur = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise GtsError("gnn_tumor_seg_b200 computes only on a CUDA device (B200, sm_100a); "
                           f"got a tensor on {t.device}. Move inputs with .to('cuda').")
        if cur is None:
            cur = torch.cuda.current_device()
        # launches go to the CURRENT device's current stream: a tensor of another GPU would fault there
        if t.device.index is not None and t.device.index != cur:
            raise GtsError(f"tensor on {t.device} but the current CUDA device is cuda:{cur}: wrap the call in "
                           f"torch.cuda.device({t.device.index}) (one process per GPU is the supported layout)")


def stream_ptr() -> int:
    import torch
    return torch.cuda.current_stream().cuda_stream


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()
