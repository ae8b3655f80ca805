"""The C-ABI library loads and exports every symbol include/gts.h declares
(no compute calls: runs without a GPU)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    txt = open(os.path.join(ROOT, "include", "gts.h")).read()
    return sorted(set(re.findall(r"GTS_API\s+[\w\s\*]+?\b(gts_\w+)\s*\(", txt)))


def test_header_declares_expected_families():
    syms = _declared_symbols()
    assert len(syms) >= 30
    for s in ("gts_csr_build", "gts_gemm_nt", "gts_gemm_tn", "gts_segmax_fwd", "gts_segmax_bwd", "gts_gat_fwd",
              "gts_project_labels", "gts_ce_weighted", "gts_last_error"):
        assert s in syms


def test_library_exports_every_declared_symbol(built_lib):
    for s in _declared_symbols():
        assert hasattr(built_lib, s), f"{s} declared in gts.h but not exported by libgts.so"


def test_binding_table_matches_header(built_lib):
    from gnn_tumor_seg_b200 import _lib
    assert sorted(_lib.EXPORTED_SYMBOLS) == _declared_symbols()


def test_version_and_error_string(built_lib):
    assert built_lib.gts_version() >= 100
    assert isinstance(built_lib.gts_last_error(), bytes)


def test_argument_validation_without_gpu(built_lib):
    """Argument checks run before any CUDA call, so they are testable on CPU."""
    from gnn_tumor_seg_b200 import _lib
    rc = built_lib.gts_csr_build(None, None, -1, 4, None, None, None, None, 0, None)
    assert rc == 1 and b"negative" in built_lib.gts_last_error()
    rc = built_lib.gts_gemm_nt(None, None)
    assert rc == 1
    assert built_lib.gts_csr_build_workspace_bytes(1000, 100) > 0
    assert built_lib.gts_gemm_tn_workspace_bytes(256, 512, 90000, 0) >= 256 * 512 * 4
    with pytest.raises(_lib.GtsError):
        _lib.check(rc, "gts_gemm_nt")


def test_gemm_args_struct_layout():
    from gnn_tumor_seg_b200._lib import GemmNtArgs
    # matches the C struct under the SysV x86-64 ABI (natural alignment)
    assert ctypes.sizeof(GemmNtArgs) == 224
    assert GemmNtArgs.K1.offset == 16 and GemmNtArgs.A2.offset == 24 and GemmNtArgs.mode.offset == 132 and GemmNtArgs.bias2.offset == 136
    assert GemmNtArgs.scatter_idx.offset == 144 and GemmNtArgs.ld_out.offset == 168
    assert GemmNtArgs.relu_bits_out.offset == 176 and GemmNtArgs.ld_aux_bits.offset == 200
    assert GemmNtArgs.zero_fill.offset == 208 and GemmNtArgs.zero_fill_bytes.offset == 216


def test_peer_exchange_abi(built_lib):
    """struct gts_peer_comm layout, buffer sizing and argument checks of the peer-memory exchange (no CUDA call)."""
    from gnn_tumor_seg_b200 import _lib
    lib = _lib.load()
    assert ctypes.sizeof(_lib.PeerComm) == 16 + 8 * _lib.MAX_PEERS and _lib.PeerComm.base.offset == 16
    assert lib.gts_peer_buffer_bytes(1000) == 256 + 2 * 1000 * 4
    assert lib.gts_peer_buffer_bytes(1001) == 256 + 2 * 1004 * 4          # staging buffers are whole float4s
    assert lib.gts_peer_publish(None, None, None) == 1 and b"null comm" in lib.gts_last_error()
    c = _lib.PeerComm()
    c.rank, c.world, c.n = 0, 40, 1024
    assert lib.gts_peer_publish(ctypes.byref(c), None, None) == 1 and b"at most" in lib.gts_last_error()
    c.world = 2
    assert lib.gts_peer_allreduce_adamw(ctypes.byref(c), None, 0, None, None, None, None, -1, 0, None) == 1
    assert b"base[0] is null" in lib.gts_last_error()


def test_integration_doc_stub_matches_the_binding():
    """INTEGRATION.md's Level-2 ctypes stub is what a reference maintainer would paste: its struct must be the library's
    struct (same fields, types and order as _lib.GemmNtArgs, which test_gemm_args_struct_layout pins to the header) and
    its positional constructor calls must fill every field."""
    import ast
    from gnn_tumor_seg_b200 import _lib
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    blocks = re.findall(r"```python\n(.*?)```", doc, re.S)
    stub = next(b for b in blocks if "class GemmNtArgs" in b)
    tree = ast.parse(stub)
    cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "GemmNtArgs")
    fields_node = next(n for n in cls.body if isinstance(n, ast.Assign) and n.targets[0].id == "_fields_")
    doc_fields = [(e.elts[0].value, getattr(ctypes, e.elts[1].attr)) for e in fields_node.value.elts]
    lib_fields = list(_lib.GemmNtArgs._fields_)         # (ctypes aliases: c_int64 is c_long here)
    assert doc_fields == lib_fields
    calls = [n for n in ast.walk(tree) if isinstance(n, ast.Call) and getattr(n.func, "id", "") == "GemmNtArgs"]
    assert len(calls) == 2 and all(len(c.args) == len(lib_fields) for c in calls)
    # the one argtypes list the stub spells out
    a = next(n for n in ast.walk(tree) if isinstance(n, ast.Assign) and isinstance(n.targets[0], ast.Attribute)
             and n.targets[0].attr == "argtypes")
    assert len(a.value.elts) == len(_lib._SIGNATURES["gts_segmax_fwd"][1])
    # every entry point of the header is named in the document's table
    for s in _declared_symbols():
        assert s in doc or (s.endswith("_workspace_bytes") and "gts_*_workspace_bytes" in doc), s
