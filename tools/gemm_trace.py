"""In-kernel clock64 trace of gemm_x3ntw_kernel (instrumented build: make -C gnn-tumor-seg_b200/csrc trace;
run with GTS_LIB_PATH=gnn-tumor-seg_b200/libgts_trace.so python tools/gemm_trace.py [K2]).
Per CTA and work item (first 8): MMA issuer waits (accumulator L / R, ready barriers), producer empty-waits,
A-split waits, epilogue wait / drain times — all in SM clocks."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from gnn_tumor_seg_b200 import _lib, ops

dev = torch.device("cuda:0")
K2 = int(sys.argv[1]) if len(sys.argv) > 1 else 256
masked = len(sys.argv) > 2 and sys.argv[2] == "mask"
dbg_flags = int(sys.argv[3]) if len(sys.argv) > 3 else 0
N, D = 90000, 256
torch.manual_seed(0)
A = [torch.randn(N, D, device=dev) for _ in range(3)]
W1 = torch.randn(D, D, device=dev) / 16
W2 = torch.randn(D, D, device=dev) / 16
b = torch.randn(D, device=dev)
lib = _lib.load()
raw = ctypes.CDLL(_lib.LIB_PATH)
ITEMS, SLOTS, CTAS = 8, 16, 148
trace = torch.zeros(CTAS * ITEMS * SLOTS, dtype=torch.int64, device=dev)


def run(i):
    kw = dict(act=ops.ACT_MASK_POS, aux=A[(i + 2) % 3]) if masked else dict(bias=b, act=ops.ACT_RELU)
    if K2:
        return ops.gemm_nt(A[i % 3], W1, A[(i + 1) % 3], W2, mode="tf32x3", **kw)
    return ops.gemm_nt(A[i % 3], W1, mode="tf32x3", **kw)


if dbg_flags:
    raw.gts_debug_set_flags(ctypes.c_int(dbg_flags))
for i in range(5):
    run(i)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(20):
    run(i)
e1.record()
torch.cuda.synchronize()
print(f"K2={K2} masked={masked} flags={dbg_flags}: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us per launch")
torch.cuda.synchronize()
raw.gts_debug_set_trace(ctypes.c_void_p(trace.data_ptr()))
run(0)
torch.cuda.synchronize()
raw.gts_debug_set_trace(ctypes.c_void_p(0))
t = trace.cpu().numpy().reshape(CTAS, ITEMS, SLOTS)
names = ["mma:item_start", "mma:accL_ok", "mma:before_accR", "mma:accR_ok", "mma:item_end", "mma:ready_wait_sum",
         "tma:empty_wait_sum", "tma:item_end", "epi:L_start", "epi:L_full", "epi:L_done", "epi:R_start", "epi:R_full",
         "epi:R_done", "asplit:full_wait_sum", "asplit:afree_wait_sum"]
for cta in (0, 1, 74, 75):
    print(f"--- CTA {cta} (rank {cta & 1})")
    base = t[cta, 0, 0] if t[cta, 0, 0] else t[cta, 0, 8]
    for it in range(6):
        r = t[cta, it]
        rel = lambda x: int(x - base) if x else -1
        if cta & 1 == 0:
            print(f" item {it}: mma start {rel(r[0])} accL_wait {r[1]-r[0]} accR_wait {r[3]-r[2]} kloop {r[4]-r[3]} "
                  f"end {rel(r[4])} ready_wait_sum {r[5]}")
        print(f"         tma empty_wait_sum {r[6]} tma_end {rel(r[7])} | epi L: wait {r[9]-r[8]} drain {r[10]-r[9]} "
              f"R: wait {r[12]-r[11]} drain {r[13]-r[12]} done {rel(r[13])} | asplit full_wait {r[14]} afree_wait {r[15]}")
# whole-kernel marks (record 7): entry -> set-up done -> first MMA item -> last epilogue -> all roles done
k = t[0::2, 7]
first = lead_first = t[0::2, 0, 0]
n_items = [max(i for i in range(7) if t[c, i, 13] or t[c, i, 10]) for c in range(0, CTAS, 2)]
last_epi = np.array([max(t[c, i, 13], t[c, i, 10]) for c, i in zip(range(0, CTAS, 2), n_items)])
last_mma = np.array([t[c, i, 4] for c, i in zip(range(0, CTAS, 2), n_items)])
print(f"kernel marks (median over leaders, clk): entry->setup {np.median(k[:,1]-k[:,0]):.0f}  setup->first item "
      f"{np.median(first-k[:,1]):.0f}  first item->last MMA issued {np.median(last_mma-first):.0f}  last MMA issued->last epilogue done "
      f"{np.median(last_epi-last_mma):.0f}  last epilogue->exit {np.median(k[:,2]-last_epi):.0f}  total {np.median(k[:,2]-k[:,0]):.0f}; "
      f"items per cluster: {sorted(set(int(x)+1 for x in n_items))}")
# aggregate over leader CTAs
lead = t[0::2]
for it in range(5):
    r = lead[:, it]
    print(f"item {it} (median over 74 leaders): accL_wait {np.median(r[:,1]-r[:,0]):.0f} accR_wait {np.median(r[:,3]-r[:,2]):.0f} "
          f"kloop {np.median(r[:,4]-r[:,3]):.0f} ready_wait {np.median(r[:,5]):.0f} tma_empty_wait {np.median(r[:,6]):.0f} "
          f"epiL drain {np.median(r[:,10]-r[:,9]):.0f} epiR drain {np.median(r[:,13]-r[:,12]):.0f} "
          f"asplit full_wait {np.median(r[:,14]):.0f} afree_wait {np.median(r[:,15]):.0f}")
