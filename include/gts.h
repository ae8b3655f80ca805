/*
 * gts.h — C-ABI of libgts.so, the B200 (sm_100a) message-passing hot path of
 * GNN-Tumor-Seg (reference: rsinghlab/GNN-Tumor-Seg, /root/reference).
 *
 * The reference has no FFI of its own: its hot path runs inside DGL
 * (SAGEConv / GATConv / dgl.batch) and numpy.  Each entry point below names
 * the reference call site (file:line under /root/reference) whose arithmetic
 * it replaces; INTEGRATION.md shows the ctypes binding a maintainer of the
 * reference would add.
 *
 * Conventions (every function):
 *   - returns 0 (GTS_OK) or a gts_status error code; gts_last_error() returns
 *     a thread-local message for the last failure on the calling thread;
 *   - takes raw DEVICE pointers with explicit sizes / leading dimensions (in
 *     elements) and a cudaStream_t passed as void*; work is enqueued on that
 *     stream, nothing synchronises the host;
 *   - never allocates or frees caller memory; scratch comes from a
 *     caller-provided workspace sized by the matching *_workspace_bytes();
 *   - matrices are row-major fp32; ids are int32 unless stated.
 */
#ifndef GTS_H_
#define GTS_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* gts_stream_t; /* cudaStream_t */

#if defined(__GNUC__)
#define GTS_API __attribute__((visibility("default")))
#else
#define GTS_API
#endif

enum gts_status {
  GTS_OK = 0,
  GTS_ERR_INVALID = 1,     /* bad argument (null pointer, negative size, misalignment) */
  GTS_ERR_CUDA = 2,        /* a CUDA runtime/driver call failed */
  GTS_ERR_WORKSPACE = 3,   /* workspace too small */
  GTS_ERR_UNSUPPORTED = 4  /* shape/mode not supported by this build */
};

/* GEMM epilogues */
enum gts_act {
  GTS_ACT_NONE = 0,
  GTS_ACT_RELU = 1,     /* C = max(acc + bias, 0) */
  GTS_ACT_MASK_POS = 2, /* C = (aux > 0) ? acc + bias : 0   (ReLU backward mask) */
  GTS_ACT_MASK_BITS = 4,        /* as GTS_ACT_MASK_POS with the mask handed over as a BIT matrix (aux_bits, 1 bit per element
                                 * instead of a 4-byte float: the ReLU mask of a 92 MB activation is 2.9 MB).  Word (m, n / 32),
                                 * bit b  <->  column 32 * (n / 32) + 4 * (b & 7) + (b >> 3); written by gts_gemm_nt
                                 * (relu_bits_out) and gts_segmax_fwd_bits.  CTA-pair tensor-core path only
                                 * (gts_gemm_nt_bits_supported). */
  GTS_ACT_MASK_BITS_SCATTER = 5, /* GTS_ACT_MASK_POS_SCATTER with the mask as a bit matrix (aux_bits), in the epilogue of the 256-wide
                                 * CTA-pair kernel: the masked tile never goes to memory, every non-zero element is added (fp32 RED)
                                 * to scatter_out[scatter_idx[m,n], n] while the next item's MMAs run.  C may be NULL. */
  GTS_ACT_MASK_POS_SCATTER = 3  /* v = (aux > 0) ? acc : 0 is NOT stored to C: it is routed through saved arg-max indices,
                                 * scatter_out[scatter_idx[m,n], n] += v (fp32 RED; rows with idx < 0 and v == 0 skipped).
                                 * The backward of the neighbour max fused into the GEMM that produces its input:
                                 * dP[argU[v,k],k] += ((dZ Wn) * (neigh > 0))[v,k]  (SURVEY.md Appendix A.1).
                                 * scatter_out must be zero-filled by the caller; C may be NULL. */
};

/* GEMM arithmetic */
enum gts_gemm_mode {
  GTS_GEMM_FP32 = 0,    /* SIMT FFMA, fp32 multiply + fp32 accumulate */
  GTS_GEMM_TF32 = 1,    /* tcgen05.mma kind::tf32, fp32 accumulate in TMEM */
  GTS_GEMM_TF32X3 = 2   /* 3xTF32 split (hi*hi + hi*lo + lo*hi) on tcgen05: fp32-accurate */
};

GTS_API int gts_version(void);
GTS_API const char* gts_last_error(void);
/* SM count / compute capability of the current device (for grid sizing & gating). */
GTS_API int gts_device_info(int32_t* sm_count, int32_t* cc_major, int32_t* cc_minor);

/* ------------------------------------------------------------------------
 * K6 — batched graph -> device CSR.
 * Replaces dgl.batch + dgl.from_networkx + DGL's lazy COO->CSR
 * (data_processing/data_loader.py:72,168; model/gnn_model.py:38).
 * ------------------------------------------------------------------------ */

/* Block-diagonal union: global ids = local ids + node_off[g(e)], where graph
 * g(e) is the one whose edge range [edge_off[g], edge_off[g+1]) holds e.
 * edge_off (int64, n_graphs+1) and node_off (int32, n_graphs+1) are device arrays. */
GTS_API int gts_batch_edges(const int32_t* src_local, const int32_t* dst_local, int64_t n_edges,
                    const int64_t* edge_off, const int32_t* node_off, int32_t n_graphs,
                    int32_t* src_global, int32_t* dst_global, gts_stream_t stream);

GTS_API size_t gts_csr_build_workspace_bytes(int64_t n_edges, int32_t n_nodes);

/* Canonical CSR keyed by `row` (pass dst for the in-edge CSR, src for the
 * out-edge CSC): indptr[n_nodes+1], indices[e] = col of the e-th entry, rows
 * ordered by edge id (== stable sort of edge ids by row), eid[e] = original
 * edge id (nullable).  Bit-exact with np.argsort(row, kind='stable'). */
GTS_API int gts_csr_build(const int32_t* row, const int32_t* col, int64_t n_edges, int32_t n_nodes,
                  int32_t* indptr, int32_t* indices, int32_t* eid,
                  void* workspace, size_t workspace_bytes, gts_stream_t stream);

/* csc2csr[q] = position in the CSR of the edge stored at CSC slot q.
 * scratch: n_edges int32. */
GTS_API int gts_edge_perm_compose(const int32_t* eid_csr, const int32_t* eid_csc, int64_t n_edges,
                          int32_t* scratch, int32_t* csc2csr, gts_stream_t stream);

/* ------------------------------------------------------------------------
 * K1 / K3 / K4 — dense contractions.  Replace nn.Linear inside DGL SAGEConv /
 * GATConv (invoked at model/networks.py:35,63,65) and their autograd.
 * ------------------------------------------------------------------------ */

/* C[M,N] = act( A1[M,K1] * B1[N,K1]^T + A2[M,K2] * B2[N,K2]^T + bias[N] ).
 * B* use the nn.Linear weight layout [out,in].  A2/B2 may be NULL (K2 = 0):
 * the two-source form is the concatenated fc_self || fc_neigh contraction
 * without materialising the concatenation.  aux/ldaux: mask source for
 * GTS_ACT_MASK_POS (same shape as C). */
typedef struct gts_gemm_nt_args {
  const float* A1; int64_t lda1; int32_t K1;
  const float* A2; int64_t lda2; int32_t K2;
  const float* B1; int64_t ldb1;
  const float* B2; int64_t ldb2;
  const float* bias;
  const float* aux; int64_t ldaux;
  float* C; int64_t ldc;
  int32_t M; int32_t N;
  int32_t act;   /* gts_act */
  int32_t mode;  /* gts_gemm_mode */
  const float* bias2;   /* optional second bias vector, added to bias (DGL<=0.7 keeps fc_self.bias and fc_neigh.bias) */
  const int32_t* scatter_idx; int64_t ld_idx;   /* GTS_ACT_MASK_POS_SCATTER: arg-max indices [M,N] */
  float* scatter_out; int64_t ld_out;           /* GTS_ACT_MASK_POS_SCATTER: destination [rows,N], zero-filled */
  uint32_t* relu_bits_out; int64_t ld_bits_out; /* optional with GTS_ACT_RELU: bit matrix [M, ld_bits_out words] of (C > 0) */
  const uint32_t* aux_bits; int64_t ld_aux_bits; /* GTS_ACT_MASK_BITS: the mask, [M, ld_aux_bits words] */
  /* Optional side job: clear zero_fill[0, zero_fill_bytes) (16-byte aligned, a multiple of 16 bytes, must not overlap
   * any operand or output of this call).  The 256-wide CTA-pair kernel does it with its two spare warps while the tiles
   * run (the whole-stack backward clears the next layer's dP this way instead of a separate memset pass); every other
   * kernel choice issues a plain memset on the stream first.  Either way the buffer is zero when the call's work is done. */
  void* zero_fill; size_t zero_fill_bytes;
} gts_gemm_nt_args;

GTS_API int gts_gemm_nt(const gts_gemm_nt_args* args, gts_stream_t stream);
/* 1 when gts_gemm_nt honours relu_bits_out / GTS_ACT_MASK_BITS for this shape and mode (the 256-wide CTA-pair kernel:
 * 3xTF32, M >= 256, N a multiple of 256); callers fall back to the float mask otherwise. */
GTS_API int gts_gemm_nt_bits_supported(int32_t M, int32_t N, int32_t mode);

/* Weight gradient: C[Mo,No] = A[K,Mo]^T * B[K,No]  (K = number of nodes).
 * Split-K over the grid with a deterministic second-pass reduction. */
GTS_API size_t gts_gemm_tn_workspace_bytes(int32_t Mo, int32_t No, int64_t K, int32_t mode);
GTS_API int gts_gemm_tn(const float* A, int64_t lda, const float* B, int64_t ldb,
                float* C, int64_t ldc, int32_t Mo, int32_t No, int64_t K, int32_t mode,
                void* workspace, size_t workspace_bytes, gts_stream_t stream);

/* Weight gradient plus the bias gradient that shares its A operand:
 * C = A^T B and colsum_out[m] = sum_k A[k,m].  In GTS_GEMM_TF32X3 the column sums are a by-product of the
 * kernel that stages A through tensor memory (no extra pass over A); otherwise gts_gemm_tn + gts_colsum.
 * Replaces the autograd of nn.Linear (weight.grad and bias.grad) inside DGL SAGEConv
 * (reference model/networks.py:35). */
GTS_API size_t gts_gemm_tn_colsum_workspace_bytes(int32_t Mo, int32_t No, int64_t K, int32_t mode);
GTS_API int gts_gemm_tn_colsum(const float* A, int64_t lda, const float* B, int64_t ldb,
                       float* C, int64_t ldc, int32_t Mo, int32_t No, int64_t K, int32_t mode,
                       float* colsum_out, void* workspace, size_t workspace_bytes, gts_stream_t stream);

/* Two weight gradients that share their A operand, plus its column sums, in ONE pass over A:
 * C1 = A^T B1, C2 = A^T B2, colsum_out[m] = sum_k A[k,m]  (dWs = dZ^T h, dWn = dZ^T neigh, db = sum dZ of one
 * SAGEConv layer).  3xTF32 CTA-pair kernel: the split A tile staged in tensor memory feeds both products;
 * other modes / shapes: gts_gemm_tn_colsum + gts_gemm_tn behind the same entry point. */
GTS_API size_t gts_gemm_tn2_colsum_workspace_bytes(int32_t Mo, int32_t No, int64_t K, int32_t mode);
GTS_API int gts_gemm_tn2_colsum(const float* A, int64_t lda, const float* B1, int64_t ldb1, const float* B2, int64_t ldb2,
                        float* C1, float* C2, int64_t ldc, int32_t Mo, int32_t No, int64_t K, int32_t mode,
                        float* colsum_out, void* workspace, size_t workspace_bytes, gts_stream_t stream);

/* out[c] = sum_r A[r,c]  (bias gradients).  Deterministic two-pass. */
GTS_API size_t gts_colsum_workspace_bytes(int64_t rows, int32_t cols);
GTS_API int gts_colsum(const float* A, int64_t lda, int64_t rows, int32_t cols, float* out,
               void* workspace, size_t workspace_bytes, gts_stream_t stream);

/* out[c, r] = in[r, c]  (small weight transposes for the data-gradient GEMMs). */
GTS_API int gts_transpose(const float* in, int64_t ldin, int32_t rows, int32_t cols,
                  float* out, int64_t ldout, gts_stream_t stream);

/* ------------------------------------------------------------------------
 * K2 / K2b — neighbour max-aggregation.  Replace DGL
 * update_all(copy_src, max) and its scatter-add backward inside
 * SAGEConv('pool') (invoked at model/networks.py:35).
 * ------------------------------------------------------------------------ */

/* neigh[v,k] = max_{u in row v} P[u,k]; argmax[v,k] = that u, FIRST maximum in
 * CSR order wins; rows with no entries give 0 / -1.  argmax may be NULL
 * (inference). */
GTS_API int gts_segmax_fwd(const float* P, int64_t ldp, const int32_t* indptr, const int32_t* indices,
                   int32_t n_nodes, int32_t D, float* neigh, int64_t ldn,
                   int32_t* argmax, int64_t ldarg, gts_stream_t stream);

/* gts_segmax_fwd that also writes pos_bits = the bit matrix of (neigh > 0) in the layout of GTS_ACT_MASK_BITS
 * ([n_nodes, ld_bits words]): the ReLU mask of fc_pool at the selected entries, consumed by the backward's dNeigh GEMM.
 * D = 256 only (gts_segmax_fwd_bits_supported); returns GTS_ERR_UNSUPPORTED otherwise. */
GTS_API int gts_segmax_fwd_bits_supported(int32_t n_nodes, int32_t D, int64_t ldp);
GTS_API int gts_segmax_fwd_bits(const float* P, int64_t ldp, const int32_t* indptr, const int32_t* indices,
                        int32_t n_nodes, int32_t D, float* neigh, int64_t ldn,
                        int32_t* argmax, int64_t ldarg, uint32_t* pos_bits, int64_t ld_bits, gts_stream_t stream);

/* dP[argmax[v,k], k] += dNeigh[v,k]; dP ([n_src_rows, D], leading dim lddp) is
 * zero-filled by the call.  fp32 atomics (order not deterministic). */
GTS_API int gts_segmax_bwd(const float* dNeigh, int64_t ldd, const int32_t* argmax, int64_t ldarg,
                   int32_t n_nodes, int32_t D, float* dP, int64_t lddp, int32_t n_src_rows,
                   gts_stream_t stream);

/* The scatter half of gts_segmax_bwd alone: dP[arg[v,k], k] += dNeigh[v,k] into a buffer the caller initialised
 * (zero for the plain backward; gts_sage_backward has the previous layer's GEMM clear it as a side job). */
GTS_API int gts_segmax_bwd_add(const float* dNeigh, int64_t ldd, const int32_t* argmax, int64_t ldarg,
                       int32_t n_nodes, int32_t D, float* dP, int64_t lddp, gts_stream_t stream);

/* Deterministic form: transposed gather over the out-edge CSC.
 * dP[u,k] = sum over out-edges (u->v) with argmax[v,k]==u of dNeigh[v,k].
 * Needs a SIMPLE graph: a duplicated (u,v) edge is counted once per copy here and once by gts_segmax_bwd
 * (BatchedGraph.has_duplicate_edges() checks; graphs from the reference's networkx Graph objects are simple). */
GTS_API int gts_segmax_bwd_det(const float* dNeigh, int64_t ldd, const int32_t* argmax, int64_t ldarg,
                       const int32_t* csc_indptr, const int32_t* csc_indices,
                       int32_t n_nodes, int32_t D, float* dP, int64_t lddp, gts_stream_t stream);

/* Sum-type aggregators of SAGEConv('mean'|'gcn') (model/networks.py:72-75):
 * mode 0: sum; 1: mean over in-edges (0 if none); 2: gcn = (sum + self[v]) / (deg+1). */
GTS_API int gts_segsum_fwd(const float* P, int64_t ldp, const int32_t* indptr, const int32_t* indices,
                   int32_t n_nodes, int32_t D, int32_t mode, float* out, int64_t ldo,
                   gts_stream_t stream);

/* Backward of gts_segsum_fwd over the out-edge CSC:
 * dP[u,k] = sum_{(u->v)} s(v) * dOut[v,k]  (+ s(u)*dOut[u,k] for gcn), with
 * s(v) = 1 | 1/deg_in(v) | 1/(deg_in(v)+1) for mode 0 | 1 | 2; deg_in from in_indptr. */
GTS_API int gts_segsum_bwd(const float* dOut, int64_t ldd, const int32_t* csc_indptr, const int32_t* csc_indices,
                   const int32_t* in_indptr, int32_t n_nodes, int32_t D, int32_t mode,
                   float* dP, int64_t lddp, gts_stream_t stream);

/* out[i] = ref[i] > 0 ? grad[i] : 0   (stand-alone ReLU backward, used when the
 * mask could not be fused into a GEMM epilogue). */
GTS_API int gts_mask_pos(const float* grad, const float* ref, int64_t n, float* out, gts_stream_t stream);

/* ------------------------------------------------------------------------
 * Whole-stack entry points: GraphSage.forward over all SAGEConv('pool') layers
 * (model/networks.py:32-36) and its backward, as ONE host call each, so the
 * per-layer kernels are enqueued back to back without returning to Python.
 * Layer l: P = relu(h Wp^T + bp); neigh = segmax(P); out = act(h Ws^T + neigh Wn^T + b),
 * act = ReLU when relu != 0.  All pointers are device pointers; weights use
 * the nn.Linear layout [out,in]; the output bias is b + b2 (either may be NULL).
 * ------------------------------------------------------------------------ */
typedef struct gts_sage_layer {
  int32_t din; int32_t dout; int32_t relu; int32_t reserved;
  const float* Wp; const float* bp; const float* Ws; const float* Wn;
  const float* b;    /* output bias (fc_self.bias), nullable */
  const float* b2;   /* second output bias (fc_neigh.bias), nullable; effective bias = b + b2 */
} gts_sage_layer;

typedef struct gts_sage_layer_grads {
  float* dWp; float* dbp; float* dWs; float* dWn; float* db;
  float* db2;   /* nullable: second copy of db (gradient of fc_neigh.bias when the module keeps both DGL<=0.7 biases) */
} gts_sage_layer_grads;

/* Workspace for n_nodes nodes.  training != 0 keeps every layer's neigh /
 * arg-max / output for the backward; the same buffer must then be handed,
 * untouched, to gts_sage_backward. */
GTS_API size_t gts_sage_workspace_bytes(const gts_sage_layer* layers, int32_t n_layers, int32_t n_nodes,
                                int32_t training, int32_t mode);

GTS_API int gts_sage_forward(const gts_sage_layer* layers, int32_t n_layers,
                     const int32_t* indptr, const int32_t* indices, int32_t n_nodes,
                     const float* feats, int64_t ldf, float* logits, int64_t ldl,
                     void* workspace, size_t workspace_bytes, int32_t training, int32_t mode,
                     gts_stream_t stream);

/* dlogits [n_nodes, dout_last] (ldd).  csc_indptr/csc_indices non-NULL selects
 * the deterministic arg-max backward.  dfeats (nullable): gradient w.r.t. the
 * input features.  dlogits is not modified. */
GTS_API int gts_sage_backward(const gts_sage_layer* layers, const gts_sage_layer_grads* grads, int32_t n_layers,
                      const int32_t* csc_indptr, const int32_t* csc_indices, int32_t n_nodes,
                      const float* feats, int64_t ldf, const float* dlogits, int64_t ldd,
                      float* dfeats, int64_t lddf,
                      void* workspace, size_t workspace_bytes, int32_t mode, gts_stream_t stream);

/* Layers layer_hi-1 .. layer_lo of gts_sage_backward.  The gradient entering layer l is kept at a fixed place of the
 * workspace, so consecutive calls [L,k) then [k,0) equal one gts_sage_backward: the data-parallel trainer
 * all-reduces the gradients of the finished layers on a side stream in between. */
GTS_API int gts_sage_backward_range(const gts_sage_layer* layers, const gts_sage_layer_grads* grads, int32_t n_layers,
                            int32_t layer_hi, int32_t layer_lo,
                            const int32_t* csc_indptr, const int32_t* csc_indices, int32_t n_nodes,
                            const float* feats, int64_t ldf, const float* dlogits, int64_t ldd,
                            float* dfeats, int64_t lddf,
                            void* workspace, size_t workspace_bytes, int32_t mode, gts_stream_t stream);

/* One training step of GNN.run_epoch's loop body minus the optimiser (model/gnn_model.py:41-45) as ONE host call:
 * logits = GraphSage.forward; sums = [sum w*nll, sum w] (zeroed by the call; loss = sums[0]/sums[1]);
 * gradients of every layer >= bwd_layer_lo into grads (normalize != 0: of the weighted MEAN, the reference's loss;
 * normalize == 0: of the un-normalised sum, for data-parallel training where the denominator is global).
 * All scratch incl. d(loss)/d(logits) lives in the caller's workspace (gts_sage_workspace_bytes, training = 1);
 * nothing is allocated, nothing synchronises: the call can be captured into a CUDA graph. */
typedef struct gts_sage_step_args {
  const gts_sage_layer* layers; const gts_sage_layer_grads* grads; int32_t n_layers;
  int32_t n_nodes;
  const int32_t* indptr; const int32_t* indices;           /* in-edge CSR */
  const int32_t* csc_indptr; const int32_t* csc_indices;   /* nullable: deterministic arg-max backward */
  const float* feats; int64_t ldf;
  const int64_t* labels; const float* class_w;
  float* sums;                  /* device float[2] */
  float* logits; int64_t ldl;   /* out: [n_nodes, dout_last] */
  void* workspace; size_t workspace_bytes;
  int32_t mode;                 /* gts_gemm_mode */
  int32_t normalize;
  int32_t bwd_layer_lo;         /* 0: whole backward; k > 0: stop after layer k (continue with gts_sage_step_backward_rest) */
  int32_t reserved;
} gts_sage_step_args;
GTS_API int gts_sage_step(const gts_sage_step_args* args, gts_stream_t stream);
/* Backward layers layer_hi-1 .. layer_lo of a step started with bwd_layer_lo = layer_hi. */
GTS_API int gts_sage_step_backward_rest(const gts_sage_step_args* args, int32_t layer_hi, int32_t layer_lo, gts_stream_t stream);

/* Per-class device timing of the whole-stack entry points above (bench.py's step breakdown of the product path):
 * gts_sage_profile(1) brackets every launch group of the following gts_sage_forward / _backward / _step calls with CUDA
 * events; gts_sage_profile_read sums them per class — 0 gemm_nt, 1 segmax_fwd, 2 segmax_bwd, 3 gemm_tn2_colsum,
 * 4 gemm_tn_colsum, 5 transpose, 6 ce_weighted (7 slots) — after synchronising on the events; gts_sage_profile(0) stops
 * and discards.  A measurement aid: not thread-safe, never enable it around a CUDA-graph capture. */
GTS_API int gts_sage_profile(int32_t enable);
GTS_API int gts_sage_profile_read(float* ms_by_kind, int32_t* calls_by_kind, int32_t n_kinds);

/* ------------------------------------------------------------------------
 * K8 — weighted-mean cross entropy (model/gnn_model.py:30,42).
 * ------------------------------------------------------------------------ */

/* sums[0] += sum_i w[y_i]*nll_i ; sums[1] += sum_i w[y_i]  (caller zeroes sums);
 * dlogits[i,c] = w[y_i]*(softmax(z_i)[c] - [c==y_i])   (NOT yet divided by
 * sums[1]; nullable).  labels are int64 as torch.LongTensor; label -100 (torch's
 * ignore_index) contributes nothing; any other label outside [0, n_classes) — where
 * torch raises — poisons sums[0] with NaN, so the loss cannot silently look fine. */
GTS_API int gts_ce_weighted(const float* logits, int64_t ld, const int64_t* labels, const float* class_w,
                    int32_t n_nodes, int32_t n_classes, float* sums, float* dlogits, int64_t ldd,
                    gts_stream_t stream);

/* x[i] *= alpha / (*denom)   (denominator read on the device: no host sync). */
GTS_API int gts_scale_by_inv(float* x, int64_t n, float alpha, const float* denom, gts_stream_t stream);

/* ------------------------------------------------------------------------
 * K7 — node -> voxel reprojection.
 * ------------------------------------------------------------------------ */

/* cls[i] = index of the first maximum of logits[i,:]  (torch.max(logits,1),
 * scripts/generate_gnn_predictions.py:66; model/gnn_model.py:65). */
GTS_API int gts_argmax_rows(const float* logits, int64_t ld, int32_t n_nodes, int32_t n_classes,
                    int32_t* cls, gts_stream_t stream);

/* project_nodes_to_img (data_processing/graph_io.py:21-24):
 * out[i] = svs[i] == -1 ? 0 : node_labels[svs[i]].  err_flag (device int32,
 * caller zeroes) is set to 1 if an id is outside [-1, n_nodes). */
GTS_API int gts_project_nodes(const int16_t* svs, int64_t n_vox, const int64_t* node_labels,
                      int32_t n_nodes, int64_t* out, int32_t* err_flag, gts_stream_t stream);

/* save_voxel_preds minus the NIfTI write
 * (scripts/generate_gnn_predictions.py:64-73): project int32 node classes
 * through the cropped int16 map, paste into the full volume
 * (uncrop_to_brats_size, data_processing/image_processing.py:21-25) and
 * relabel through lut (swap_labels_to_brats, scripts/preprocess_dataset.py:159-169).
 * inv_x/inv_y/inv_z (lengths VX,VY,VZ): crop coordinate of each full-volume
 * plane or -1.  Every voxel of vol[VX,VY,VZ] is written exactly once.
 * err_flag set to 1 on an id outside [-1,n_nodes) or a class outside [0,n_lut). */
GTS_API int gts_project_labels(const int16_t* svs, int32_t X, int32_t Y, int32_t Z,
                       const int32_t* inv_x, const int32_t* inv_y, const int32_t* inv_z,
                       const int32_t* node_cls, int32_t n_nodes,
                       const int16_t* lut, int32_t n_lut,
                       int16_t* vol, int32_t VX, int32_t VY, int32_t VZ,
                       int32_t* err_flag, gts_stream_t stream);

/* determine_tumor_crop (data_processing/image_processing.py:8-17; used by
 * scripts/generate_joint_predictions.py:67 and data_loader.py PredLogitDataset.get_crop:148-150):
 * occ_x[x] = 1 iff plane x of the cropped map holds a voxel whose supervoxel class is non-zero
 * (same for y, z; background -1 counts as healthy).  The reference's crop = planes whose occupancy,
 * dilated by one plane on each side (scipy binary_dilation, 3-D cross), is set; all planes when
 * nothing is predicted tumorous.  err_flag set to 1 on an id outside [-1,n_nodes). */
GTS_API int gts_tumor_plane_occupancy(const int16_t* svs, int32_t X, int32_t Y, int32_t Z,
                              const int32_t* node_cls, int32_t n_nodes,
                              int32_t* occ_x, int32_t* occ_y, int32_t* occ_z,
                              int32_t* err_flag, gts_stream_t stream);

/* save_voxel_logits (scripts/generate_gnn_predictions.py:55-62;
 * scripts/generate_joint_predictions.py:64-66): out[i,:] = svs[i]==-1 ?
 * bg_row : node_logits[svs[i],:]. */
GTS_API int gts_project_logits(const int16_t* svs, int64_t n_vox, const float* node_logits, int64_t ld,
                       int32_t n_nodes, int32_t n_classes, const float* bg_row, float* out,
                       int32_t* err_flag, gts_stream_t stream);

/* ------------------------------------------------------------------------
 * K5 — GATConv edge-score + edge-softmax + weighted aggregation, fused
 * (DGL GATConv.forward invoked at model/networks.py:63,65).
 * Z is [n_nodes, H*F] (ldz), el/er/rowmax/rowsum are [n_nodes, H].
 * ------------------------------------------------------------------------ */

/* el[u,h] = <Z[u,h,:], attn_l[h,:]>, er likewise. */
GTS_API int gts_gat_scores(const float* Z, int64_t ldz, const float* attn_l, const float* attn_r,
                   int32_t n_nodes, int32_t H, int32_t F, float* el, float* er, gts_stream_t stream);

/* out[v,h,:] = act( sum_{e=(u->v)} softmax_e(leaky_relu(el[u,h]+er[v,h])) * Z[u,h,:]
 *                   + res[v,h,:] + bias[h,:] ),  act: 0 none, 1 ELU.
 * rowmax/rowsum save the softmax statistics for the backward.  A row with no
 * in-edges sets *err_flag = 1 (DGL raises DGLError). */
GTS_API int gts_gat_fwd(const float* Z, int64_t ldz, const float* el, const float* er,
                const int32_t* indptr, const int32_t* indices,
                int32_t n_nodes, int32_t H, int32_t F, float slope,
                const float* res, int64_t ldres, const float* bias, int32_t act,
                float* out, int64_t ldo, float* rowmax, float* rowsum,
                int32_t* err_flag, gts_stream_t stream);

/* dpre = dout * ELU'(out) (act==1) or dout (act==0). */
GTS_API int gts_gat_act_bwd(const float* dout, const float* out, int64_t n, int32_t act, float* dpre,
                    gts_stream_t stream);

/* Backward pass A (by destination): per edge dt[e,h] (CSR order), der[v,h]. */
GTS_API int gts_gat_bwd_dst(const float* Z, int64_t ldz, const float* el, const float* er,
                    const float* rowmax, const float* rowsum,
                    const int32_t* indptr, const int32_t* indices,
                    const float* dO, int64_t lddo,
                    int32_t n_nodes, int32_t H, int32_t F, float slope,
                    float* dt_edge, float* der, gts_stream_t stream);

/* Backward pass B (by source, over the out-edge CSC): dZ[u,h,:] =
 * sum_{e=(u->v)} alpha_e*dO[v,h,:] + del[u,h]*attn_l[h,:] + der[u,h]*attn_r[h,:];
 * del[u,h] = sum_{out(u)} dt_e is also written. */
GTS_API int gts_gat_bwd_src(const float* el, const float* er, const float* rowmax, const float* rowsum,
                    const int32_t* csc_indptr, const int32_t* csc_indices, const int32_t* csc2csr,
                    const float* dO, int64_t lddo, const float* dt_edge, const float* der,
                    const float* attn_l, const float* attn_r,
                    int32_t n_nodes, int32_t H, int32_t F, float slope,
                    float* dZ, int64_t lddz, float* del, gts_stream_t stream);

/* dattn[h,f] = sum_u coef[u,h] * Z[u,h,f]   (gradients of attn_l / attn_r). */
GTS_API size_t gts_gat_attn_grad_workspace_bytes(int32_t n_nodes, int32_t H, int32_t F);
GTS_API int gts_gat_attn_grad(const float* Z, int64_t ldz, const float* coef, int32_t n_nodes,
                      int32_t H, int32_t F, float* dattn, void* workspace, size_t workspace_bytes,
                      gts_stream_t stream);

/* Both attention-vector gradients in one pass over Z (dattn_l from coef_l = del, dattn_r from coef_r = der);
 * workspace: 2 x gts_gat_attn_grad_workspace_bytes. */
GTS_API int gts_gat_attn_grad2(const float* Z, int64_t ldz, const float* coef_l, const float* coef_r, int32_t n_nodes,
                       int32_t H, int32_t F, float* dattn_l, float* dattn_r, void* workspace, size_t workspace_bytes,
                       gts_stream_t stream);

/* ------------------------------------------------------------------------
 * Optimiser step on a flat parameter arena (model/gnn_model.py:28,46):
 * AdamW exactly as torch.optim.AdamW (decoupled weight decay, bias
 * correction), grads pre-scaled by grad_scale / (*grad_denom) when
 * grad_denom != NULL (data-parallel loss denominator).
 * ------------------------------------------------------------------------ */
GTS_API int gts_adamw_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n,
                   float lr, float beta1, float beta2, float eps, float weight_decay, int32_t step,
                   float grad_scale, const float* grad_denom, gts_stream_t stream);

/* Same update with the hyper-parameters read on the DEVICE, for steps replayed from a CUDA graph:
 * hyper = device float[8] {lr, beta1, beta2, eps, weight_decay, step, 0, 0}; hyper[5] (the step count,
 * exact in fp32 up to 2^24) is incremented by the call before use; ExponentialLR = the host rewriting hyper[0]
 * between replays (model/gnn_model.py:29,47). */
GTS_API int gts_adamw_step_dev(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n,
                       float* hyper, float grad_scale, const float* grad_denom, gts_stream_t stream);

/* ------------------------------------------------------------------------
 * Data-parallel gradient exchange over NVLink peer memory, fused with the
 * optimiser (the reference trains on one device,
 * model/gnn_model.py:23,41-47 — whole graphs per rank is the natural shard).
 *
 * One process per GPU.  Every rank owns ONE exchange buffer in its own HBM
 *   [ 256-byte header | staging buffer 0: n floats | staging buffer 1: n floats ]
 * that the other ranks of the box map through CUDA IPC (NVLink / NVSwitch
 * peer access).  A step is two launches, both capturable in a CUDA graph:
 *   gts_peer_publish          copies the rank's un-normalised gradient arena
 *                             (+ the two loss sums behind it) into staging
 *                             buffer (epoch & 1) and then raises the rank's
 *                             flag in EVERY peer's header (release, system scope);
 *   gts_peer_allreduce_adamw  waits until every rank's flag has reached the
 *                             epoch (acquire, bounded spin), reads all ranks'
 *                             staging buffers over NVLink, sums them in rank
 *                             order (identical bits on every rank), writes the
 *                             sums back to the gradient arena and — apply != 0 —
 *                             runs the AdamW update of gts_adamw_step_dev on
 *                             sum / (summed loss denominator) in the same pass.
 * The epoch counter lives in the header (device side), so replays of a
 * captured graph and eager calls interleave freely; the two staging buffers
 * alternate by epoch parity, which makes one flag exchange per step enough
 * (a rank can run at most one step ahead of the slowest one).  A peer that
 * never arrives sets the header's error word after ~20 s instead of hanging
 * the GPU (gts_peer_status).
 *
 * gts_peer_alloc / gts_peer_free / gts_peer_open / gts_peer_close are the only
 * entry points of this library that allocate or map device memory (an
 * IPC-exportable buffer cannot be carved from a caller's pool).
 * ------------------------------------------------------------------------ */
#define GTS_MAX_PEERS 16
#define GTS_PEER_HANDLE_BYTES 64

typedef struct gts_peer_comm {
  int32_t rank, world;            /* world <= GTS_MAX_PEERS */
  int64_t n;                      /* floats per staging buffer (multiple of 4) */
  void* base[GTS_MAX_PEERS];      /* base[r]: rank r's exchange buffer as mapped in THIS process
                                   * (gts_peer_alloc for r == rank, gts_peer_open of r's handle otherwise) */
} gts_peer_comm;

GTS_API size_t gts_peer_buffer_bytes(int64_t n);
/* cudaMalloc + zero-fill + cudaIpcGetMemHandle on the current device; handle: GTS_PEER_HANDLE_BYTES bytes to send to
 * the other ranks (any byte transport: the host side uses torch.distributed.all_gather_object). */
GTS_API int gts_peer_alloc(size_t bytes, void** dptr, unsigned char* handle);
GTS_API int gts_peer_free(void* dptr);
/* cudaIpcOpenMemHandle with lazy peer access from the current device; fails (GTS_ERR_CUDA) where the two devices have
 * no peer path — the caller then keeps the NCCL all-reduce. */
GTS_API int gts_peer_open(const unsigned char* handle, void** dptr);
GTS_API int gts_peer_close(void* dptr);
/* src: n floats (16-byte aligned). */
GTS_API int gts_peer_publish(const gts_peer_comm* comm, const float* src, gts_stream_t stream);
/* grads: n floats, receives the sums.  apply != 0: AdamW on the first n_params elements (param / exp_avg /
 * exp_avg_sq / hyper as in gts_adamw_step_dev, hyper[5] incremented by the call) with the summed gradient divided by
 * the summed element denom_index of the exchanged vector (denom_index < 0: no division). */
GTS_API int gts_peer_allreduce_adamw(const gts_peer_comm* comm, float* grads, int64_t n_params, float* param,
                             float* exp_avg, float* exp_avg_sq, float* hyper, int64_t denom_index, int32_t apply,
                             gts_stream_t stream);
/* Synchronous read of the local header: completed epochs and the error word (0 = ok, 1 = a peer's flag timed out). */
GTS_API int gts_peer_status(const gts_peer_comm* comm, uint32_t* epoch, uint32_t* error);

#ifdef __cplusplus
}
#endif
#endif /* GTS_H_ */
