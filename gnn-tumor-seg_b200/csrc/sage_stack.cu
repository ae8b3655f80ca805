// sage_stack.cu — GraphSage.forward / backward over the whole SAGEConv('pool')
// stack (reference model/networks.py:20-36) as one host call each.  Pure
// orchestration of the K1/K2/K3 kernels: every launch goes to the caller's
// stream, all scratch is carved from the caller's workspace.
#include "common.cuh"
#include <cstdlib>
#include <vector>

namespace gts {

struct StackPlan {
  static constexpr int kMaxLayers = 64;
  int L = 0;
  int64_t N = 0;
  size_t neigh[kMaxLayers], arg[kMaxLayers], out[kMaxLayers];
  size_t P = 0, g0 = 0, g1 = 0, dP = 0, dP2 = 0, gemm_ws = 0, colsum_ws = 0, dlogits = 0;
  size_t wnT[kMaxLayers], wsT[kMaxLayers], wpT[kMaxLayers];      // transposed weights of every layer (one batched launch)
  // ReLU masks as bit matrices (1 bit per element): neigh_bits[l] = (neigh_l > 0) written by the seg-max forward,
  // out_bits[l] = (out_l > 0) written by the concat GEMM's epilogue; 0 = not available for that layer (float mask then)
  size_t neigh_bits[kMaxLayers], out_bits[kMaxLayers];
  size_t gemm_ws_bytes = 0, colsum_ws_bytes = 0;
  // deferred split-K reductions (defer_reduce()): every weight-gradient GEMM of the backward keeps its partial products
  // in its own slice until the ONE reduction launch at the end of the layer range
  size_t tn2_ws[kMaxLayers], tn_ws[kMaxLayers], tn2_ws_bytes[kMaxLayers], tn_ws_bytes[kMaxLayers];
  bool defer = false;
  size_t total = 0;
};

// One split-K reduction launch per backward range instead of one to three per weight-gradient GEMM (30 launches of
// ~7 us in the 7x256 step); GTS_DEFER_REDUCE=0 keeps the immediate launches and the single shared GEMM workspace.
static bool defer_reduce() {
  static const bool on = !(getenv("GTS_DEFER_REDUCE") && atoi(getenv("GTS_DEFER_REDUCE")) == 0);
  return on;
}

struct DeferGuard {      // installs the batch on this thread for the lifetime of a backward range
  explicit DeferGuard(ReduceBatch* b) { splitk_defer_set(b); }
  ~DeferGuard() { splitk_defer_set(nullptr); }
};

// The dh GEMM of layer l+1 clears layer l's dP with its spare warps (gts_gemm_nt_args.zero_fill) instead of a memset
// pass in front of every scatter; GTS_ZERO_IN_GEMM=0 keeps the memsets (A/B runs) and the single dP buffer.
static bool zero_in_gemm() {
  static const bool on = !(getenv("GTS_ZERO_IN_GEMM") && atoi(getenv("GTS_ZERO_IN_GEMM")) == 0);
  return on;
}

static bool make_plan(const gts_sage_layer* layers, int L, int64_t N, bool training, int mode, StackPlan& pl) {
  if (L < 1 || L > StackPlan::kMaxLayers) return false;
  pl.L = L; pl.N = N;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += align_up(bytes ? bytes : 16, 256); return o; };
  int max_din = 0, max_dim = 0;
  for (int l = 0; l < L; ++l) {
    if (layers[l].din < 1 || layers[l].dout < 1) return false;
    if (l > 0 && layers[l].din != layers[l - 1].dout) return false;
    max_din = layers[l].din > max_din ? layers[l].din : max_din;
    max_dim = layers[l].din > max_dim ? layers[l].din : max_dim;
    max_dim = layers[l].dout > max_dim ? layers[l].dout : max_dim;
  }
  const size_t n = (size_t)N;
  if (training) {
    static const bool bits_on = !(getenv("GTS_MASK_BITS") && atoi(getenv("GTS_MASK_BITS")) == 0);
    for (int l = 0; l < L; ++l) {
      pl.neigh[l] = take(n * layers[l].din * 4);
      pl.arg[l] = take(n * layers[l].din * 4);
      pl.out[l] = (l + 1 < L) ? take(n * layers[l].dout * 4) : 0;
      // bit masks where both the producer and the consumer (256-wide CTA-pair GEMM) support them
      const bool nb = bits_on && gts_segmax_fwd_bits_supported((int32_t)N, layers[l].din, layers[l].din) &&
                      gts_gemm_nt_bits_supported((int32_t)N, layers[l].din, mode);
      pl.neigh_bits[l] = nb ? take(n * (layers[l].din / 32) * 4) : 0;
      const bool ob = bits_on && l + 1 < L && layers[l].relu && gts_gemm_nt_bits_supported((int32_t)N, layers[l].dout, mode);
      pl.out_bits[l] = ob ? take(n * (layers[l].dout / 32) * 4) : 0;
    }
    pl.P = take(n * max_din * 4);        // forward: pooled features; backward: dNeigh'
    pl.g0 = take(n * max_dim * 4);       // backward: dZ / dh ping-pong
    pl.g1 = take(n * max_dim * 4);
    pl.dP = take(n * max_din * 4);
    // second dP: layer l scatters into dP[l & 1] while layer l+1's dh GEMM (the last reader of the other one) clears it
    pl.dP2 = zero_in_gemm() ? take(n * max_din * 4) : pl.dP;
    for (int l = 0; l < L; ++l) {
      pl.wnT[l] = take((size_t)layers[l].din * layers[l].dout * 4);
      pl.wsT[l] = take((size_t)layers[l].din * layers[l].dout * 4);
      pl.wpT[l] = take((size_t)layers[l].din * layers[l].din * 4);
    }
    size_t gw = 256, cw = 256;
    for (int l = 0; l < L; ++l) {
      size_t a = gts_gemm_tn2_colsum_workspace_bytes(layers[l].dout, layers[l].din, N, mode);
      size_t b = gts_gemm_tn_colsum_workspace_bytes(layers[l].din, layers[l].din, N, mode);
      gw = a > gw ? a : gw; gw = b > gw ? b : gw;
      size_t c = gts_colsum_workspace_bytes(N, layers[l].dout), d = gts_colsum_workspace_bytes(N, layers[l].din);
      cw = c > cw ? c : cw; cw = d > cw ? d : cw;
    }
    pl.gemm_ws_bytes = gw; pl.colsum_ws_bytes = cw;
    pl.gemm_ws = take(gw);
    pl.colsum_ws = take(cw);
    pl.defer = defer_reduce();
    for (int l = 0; l < L; ++l) {
      if (pl.defer) {
        pl.tn2_ws_bytes[l] = gts_gemm_tn2_colsum_workspace_bytes(layers[l].dout, layers[l].din, N, mode);
        pl.tn_ws_bytes[l] = gts_gemm_tn_colsum_workspace_bytes(layers[l].din, layers[l].din, N, mode);
        pl.tn2_ws[l] = take(pl.tn2_ws_bytes[l]);
        pl.tn_ws[l] = take(pl.tn_ws_bytes[l]);
      } else {
        pl.tn2_ws[l] = pl.tn_ws[l] = pl.gemm_ws;
        pl.tn2_ws_bytes[l] = pl.tn_ws_bytes[l] = gw;
      }
    }
    pl.dlogits = take(n * layers[L - 1].dout * 4);     // gts_sage_step: gradient of the loss w.r.t. the logits
  } else {
    for (int l = 0; l < L; ++l) pl.neigh_bits[l] = pl.out_bits[l] = 0;
    pl.P = take(n * max_din * 4);
    pl.neigh[0] = take(n * max_din * 4);      // single reused neigh buffer
    pl.g0 = take(n * max_dim * 4);            // activations ping-pong
    pl.g1 = take(n * max_dim * 4);
  }
  pl.total = off;
  return true;
}

// Optional per-class device timing of the whole-stack calls (bench.py's step breakdown of the PRODUCT path): when
// enabled, every launch group below is bracketed by two CUDA events on the caller's stream.  Off by default; never
// enable it around a CUDA-graph capture.
enum ProfKind { kProfGemmNt = 0, kProfSegFwd, kProfSegBwd, kProfTn2, kProfTn, kProfTranspose, kProfCe, kProfKinds };
struct ProfState {
  bool on = false;
  struct Rec { int kind; cudaEvent_t e0, e1; };
  std::vector<Rec> recs;
  std::vector<cudaEvent_t> pool;
  cudaEvent_t get() {
    if (!pool.empty()) { cudaEvent_t e = pool.back(); pool.pop_back(); return e; }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
  }
};
static ProfState g_prof;
struct Prof {
  cudaStream_t st; int idx = -1;
  Prof(int kind, gts_stream_t stream) : st(as_stream(stream)) {
    if (!g_prof.on) return;
    ProfState::Rec r{kind, g_prof.get(), g_prof.get()};
    cudaEventRecord(r.e0, st);
    g_prof.recs.push_back(r);
    idx = (int)g_prof.recs.size() - 1;
  }
  ~Prof() { if (idx >= 0) cudaEventRecord(g_prof.recs[idx].e1, st); }
};

static inline float* at(void* ws, size_t off) { return reinterpret_cast<float*>(reinterpret_cast<char*>(ws) + off); }

#define GTS_TRY(expr)            \
  do {                           \
    int _rc = (expr);            \
    if (_rc != GTS_OK) return _rc; \
  } while (0)

static int nt(const float* A1, int64_t lda1, int K1, const float* B1, int64_t ldb1, const float* A2, int64_t lda2, int K2,
              const float* B2, int64_t ldb2, const float* bias, int act, const float* aux, int64_t ldaux, float* C,
              int64_t ldc, int M, int N, int mode, gts_stream_t st, const float* bias2 = nullptr,
              uint32_t* relu_bits_out = nullptr, const uint32_t* aux_bits = nullptr,
              const int32_t* scatter_idx = nullptr, float* scatter_out = nullptr, void* zero_fill = nullptr,
              size_t zero_fill_bytes = 0) {
  gts_gemm_nt_args a{};
  a.zero_fill = zero_fill; a.zero_fill_bytes = zero_fill_bytes;
  a.scatter_idx = scatter_idx; a.ld_idx = N; a.scatter_out = scatter_out; a.ld_out = N;
  a.relu_bits_out = relu_bits_out; a.ld_bits_out = N / 32; a.aux_bits = aux_bits; a.ld_aux_bits = N / 32;
  a.bias2 = bias2;
  a.A1 = A1; a.lda1 = lda1; a.K1 = K1; a.A2 = A2; a.lda2 = lda2; a.K2 = K2;
  a.B1 = B1; a.ldb1 = ldb1; a.B2 = B2; a.ldb2 = ldb2; a.bias = bias; a.aux = aux; a.ldaux = ldaux;
  a.C = C; a.ldc = ldc; a.M = M; a.N = N; a.act = act; a.mode = mode;
  Prof prof(kProfGemmNt, st);
  return gts_gemm_nt(&a, st);
}

}  // namespace gts

using namespace gts;

extern "C" {

size_t gts_sage_workspace_bytes(const gts_sage_layer* layers, int32_t n_layers, int32_t n_nodes,
                                int32_t training, int32_t mode) {
  StackPlan pl;
  if (!layers || n_nodes < 0 || !make_plan(layers, n_layers, n_nodes, training != 0, mode, pl)) return 0;
  return pl.total;
}

int gts_sage_forward(const gts_sage_layer* layers, int32_t n_layers,
                     const int32_t* indptr, const int32_t* indices, int32_t n_nodes,
                     const float* feats, int64_t ldf, float* logits, int64_t ldl,
                     void* workspace, size_t workspace_bytes, int32_t training, int32_t mode,
                     gts_stream_t stream) {
  GTS_CHECK_ARG(layers && n_layers >= 1, "gts_sage_forward: no layers");
  GTS_CHECK_ARG(n_nodes >= 0, "gts_sage_forward: negative n_nodes");
  if (n_nodes == 0) return GTS_OK;
  GTS_CHECK_ARG(indptr && feats && logits && workspace, "gts_sage_forward: null pointer");
  StackPlan pl;
  GTS_CHECK_ARG(make_plan(layers, n_layers, n_nodes, training != 0, mode, pl), "gts_sage_forward: inconsistent layer dims");
  if (workspace_bytes < pl.total) {
    set_error("gts_sage_forward: workspace %zu < required %zu", workspace_bytes, pl.total);
    return GTS_ERR_WORKSPACE;
  }
  const int N = n_nodes;
  const float* h = feats;
  int64_t ldh = ldf;
  for (int l = 0; l < n_layers; ++l) {
    const gts_sage_layer& ly = layers[l];
    GTS_CHECK_ARG(ly.Wp && ly.bp && ly.Ws && ly.Wn, "gts_sage_forward: layer %d has a null weight", l);
    float* P = at(workspace, pl.P);
    GTS_TRY(nt(h, ldh, ly.din, ly.Wp, ly.din, nullptr, 0, 0, nullptr, 0, ly.bp, GTS_ACT_RELU, nullptr, 0, P, ly.din, N,
               ly.din, mode, stream));
    float* neigh = at(workspace, training ? pl.neigh[l] : pl.neigh[0]);
    int32_t* arg = training ? reinterpret_cast<int32_t*>(at(workspace, pl.arg[l])) : nullptr;
    {
      Prof prof(kProfSegFwd, stream);
      if (training && pl.neigh_bits[l])
        GTS_TRY(gts_segmax_fwd_bits(P, ly.din, indptr, indices, N, ly.din, neigh, ly.din, arg, ly.din,
                                    reinterpret_cast<uint32_t*>(at(workspace, pl.neigh_bits[l])), ly.din / 32, stream));
      else
        GTS_TRY(gts_segmax_fwd(P, ly.din, indptr, indices, N, ly.din, neigh, ly.din, arg, ly.din, stream));
    }
    const bool last = (l + 1 == n_layers);
    float* out;
    int64_t ldo;
    if (last) { out = logits; ldo = ldl; }
    else if (training) { out = at(workspace, pl.out[l]); ldo = ly.dout; }
    else { out = at(workspace, (l & 1) ? pl.g1 : pl.g0); ldo = ly.dout; }
    GTS_TRY(nt(h, ldh, ly.din, ly.Ws, ly.din, neigh, ly.din, ly.din, ly.Wn, ly.din, ly.b,
               ly.relu ? GTS_ACT_RELU : GTS_ACT_NONE, nullptr, 0, out, ldo, N, ly.dout, mode, stream, ly.b2,
               (training && pl.out_bits[l]) ? reinterpret_cast<uint32_t*>(at(workspace, pl.out_bits[l])) : nullptr));
    h = out; ldh = ldo;
  }
  return GTS_OK;
}

// Layers layer_hi-1 .. layer_lo of the backward pass.  The gradient that enters layer l (dZ) lives at a fixed place:
// dlogits for the top layer, otherwise the ping-pong buffer (L-2-l) & 1 of the workspace — so consecutive range calls
// continue each other (the data-parallel trainer all-reduces the finished layers' gradients in between).
static int backward_range(const gts_sage_layer* layers, const gts_sage_layer_grads* grads, int32_t n_layers,
                          int32_t layer_hi, int32_t layer_lo,
                          const int32_t* csc_indptr, const int32_t* csc_indices, int32_t n_nodes,
                          const float* feats, int64_t ldf, const float* dlogits, int64_t ldd,
                          float* dfeats, int64_t lddf,
                          void* workspace, size_t workspace_bytes, int32_t mode, gts_stream_t stream) {
  GTS_CHECK_ARG(layers && grads && n_layers >= 1, "gts_sage_backward: no layers");
  GTS_CHECK_ARG(n_nodes >= 0, "gts_sage_backward: negative n_nodes");
  GTS_CHECK_ARG(0 <= layer_lo && layer_lo <= layer_hi && layer_hi <= n_layers, "gts_sage_backward: bad layer range [%d,%d)", layer_lo, layer_hi);
  // the gradient handed in is taken as the gradient of the top layer's OUTPUT: a ReLU there would need its mask,
  // which this entry point does not see (the reference's top layer has no activation, model/networks.py:30)
  GTS_CHECK_ARG(!layers[n_layers - 1].relu, "gts_sage_backward: a ReLU on the last layer is not supported (apply its mask to dlogits and clear the flag)");
  StackPlan pl;
  GTS_CHECK_ARG(make_plan(layers, n_layers, n_nodes, true, mode, pl), "gts_sage_backward: inconsistent layer dims");
  if (n_nodes > 0 && workspace_bytes < pl.total) {
    set_error("gts_sage_backward: workspace %zu < required %zu", workspace_bytes, pl.total);
    return GTS_ERR_WORKSPACE;
  }
  GTS_CHECK_ARG(n_nodes == 0 || (feats && dlogits && workspace), "gts_sage_backward: null pointer");
  const int N = n_nodes;
  // every weight transpose of the range (data-gradient GEMMs consume K-major B operands) in launches of <= 96 jobs
  {
    TransposeBatch tb;
    auto push = [&](const float* in, int rows, int cols, size_t off) -> int {
      if (tb.n == TransposeBatch::kMaxJobs) { int rc = launch_transpose_batch(tb, as_stream(stream)); if (rc != GTS_OK) return rc; tb.n = 0; }
      tb.job[tb.n++] = TransposeJob{in, at(workspace, off), cols, rows, rows, cols};
      return GTS_OK;
    };
    for (int l = layer_lo; l < layer_hi && N > 0; ++l) {
      const gts_sage_layer& ly = layers[l];
      GTS_CHECK_ARG(ly.Wp && ly.Ws && ly.Wn, "gts_sage_backward: layer %d has a null weight", l);
      GTS_TRY(push(ly.Wn, ly.dout, ly.din, pl.wnT[l]));
      if (l > 0 || dfeats) {
        GTS_TRY(push(ly.Ws, ly.dout, ly.din, pl.wsT[l]));
        GTS_TRY(push(ly.Wp, ly.din, ly.din, pl.wpT[l]));
      }
    }
    Prof prof(kProfTranspose, stream);
    if (tb.n > 0) GTS_TRY(launch_transpose_batch(tb, as_stream(stream)));
  }
  ReduceBatch reduce_batch;
  DeferGuard defer_guard(pl.defer ? &reduce_batch : nullptr);
  struct BiasCopy { float* dst; const float* src; size_t bytes; };
  BiasCopy bias_copy[StackPlan::kMaxLayers];
  int n_bias_copy = 0;
  bool dp_cleared = false;       // this layer's dP was zeroed by the previous (upper) layer's dh GEMM
  for (int l = layer_hi - 1; l >= layer_lo; --l) {
    const gts_sage_layer& ly = layers[l];
    const gts_sage_layer_grads& g = grads[l];
    GTS_CHECK_ARG(g.dWp && g.dbp && g.dWs && g.dWn && g.db, "gts_sage_backward: layer %d has a null gradient pointer", l);
    const float* dZ = dlogits;
    int64_t ldz = ldd;
    if (l + 1 < n_layers) { dZ = at(workspace, ((n_layers - 2 - l) & 1) ? pl.g1 : pl.g0); ldz = ly.dout; }
    const float* h = (l == 0) ? feats : at(workspace, pl.out[l - 1]);
    const int64_t ldh = (l == 0) ? ldf : ly.din;
    const float* neigh = at(workspace, pl.neigh[l]);
    const int32_t* arg = reinterpret_cast<const int32_t*>(at(workspace, pl.arg[l]));
    // (dZ already carries this layer's ReLU mask: the consumer's epilogue applied (out > 0))
    // dWs = dZ^T h, dWn = dZ^T neigh and db = column sums of dZ: one pass over dZ in the 3xTF32 mode
    {
      Prof prof(kProfTn2, stream);
      GTS_TRY(gts_gemm_tn2_colsum(dZ, ldz, h, ldh, neigh, ly.din, g.dWs, g.dWn, ly.din, ly.dout, ly.din, N, mode, g.db,
                                  at(workspace, pl.tn2_ws[l]), pl.tn2_ws_bytes[l], stream));
    }
    if (g.db2 && N > 0) {    // fc_neigh.bias enters the output as fc_self.bias does: same gradient, its own arena slot
      if (pl.defer)          // (db exists only after the deferred reduction: copied behind it)
        bias_copy[n_bias_copy++] = BiasCopy{g.db2, g.db, sizeof(float) * (size_t)ly.dout};
      else
        GTS_CUDA(cudaMemcpyAsync(g.db2, g.db, sizeof(float) * (size_t)ly.dout, cudaMemcpyDeviceToDevice, as_stream(stream)));
    }
    // dNeigh' = (dZ Wn) * (neigh > 0)
    const float* WnT = at(workspace, pl.wnT[l]);
    // Measured (profiles/r01_gemm_x3_pipeline.md): routing the masked tile through the arg-max from inside the GEMM
    // epilogue (GTS_ACT_MASK_POS_SCATTER) is correct but slower — 320 us against 60 (GEMM) + 118 (zero-fill + scatter):
    // four epilogue warps per SM cannot keep as many REDs in flight as a full-occupancy scatter kernel.
    float* dNeigh = at(workspace, pl.P);
    float* dP = at(workspace, (l & 1) ? pl.dP2 : pl.dP);
    const bool det = csc_indptr && csc_indices;       // the gather form writes every element: nothing to clear
    // dP[arg[v,k],k] += ((dZ Wn) * (neigh > 0))[v,k] straight out of the GEMM's epilogue (GTS_ACT_MASK_BITS_SCATTER): the
    // epilogue warps keep the tile in registers, the accumulator is already back with the MMA issuer, and the REDs go
    // out while the next item's MMAs run — no dNeigh tensor, no separate scatter pass.  Correct (tests) and MEASURED
    // SLOWER again (round 1: 320 us with the accumulator held; round 2: 230 us against 64 + 84): 512 scattered 32-lane
    // REDs per half tile from four warps cost the SM's LSU ~70 clk each (37 000 clk per half tile, 3x the main loop),
    // where the stand-alone kernel spreads them over 64 resident warps.  Off unless GTS_FUSED_SCATTER=1.
    static const bool fused_on = getenv("GTS_FUSED_SCATTER") && atoi(getenv("GTS_FUSED_SCATTER")) == 1;
    const bool fused_scatter = fused_on && pl.neigh_bits[l] && !det;
    if (!det && !dp_cleared && N > 0) {
      Prof prof(kProfSegBwd, stream);
      GTS_CUDA(cudaMemsetAsync(dP, 0, sizeof(float) * (size_t)N * ly.din, as_stream(stream)));
    }
    if (fused_scatter) {
      GTS_TRY(nt(dZ, ldz, ly.dout, WnT, ly.dout, nullptr, 0, 0, nullptr, 0, nullptr, GTS_ACT_MASK_BITS_SCATTER, nullptr, 0,
                 nullptr, ly.din, N, ly.din, mode, stream, nullptr, nullptr,
                 reinterpret_cast<const uint32_t*>(at(workspace, pl.neigh_bits[l])), arg, dP));
    } else if (pl.neigh_bits[l])      // the mask (neigh > 0) as bits: 1/32 of the float operand's traffic in the epilogue
      GTS_TRY(nt(dZ, ldz, ly.dout, WnT, ly.dout, nullptr, 0, 0, nullptr, 0, nullptr, GTS_ACT_MASK_BITS, nullptr, 0, dNeigh,
                 ly.din, N, ly.din, mode, stream, nullptr, nullptr,
                 reinterpret_cast<const uint32_t*>(at(workspace, pl.neigh_bits[l]))));
    else
      GTS_TRY(nt(dZ, ldz, ly.dout, WnT, ly.dout, nullptr, 0, 0, nullptr, 0, nullptr, GTS_ACT_MASK_POS, neigh, ly.din, dNeigh,
                 ly.din, N, ly.din, mode, stream));
    if (!fused_scatter) {
      Prof prof(kProfSegBwd, stream);
      if (det)
        GTS_TRY(gts_segmax_bwd_det(dNeigh, ly.din, arg, ly.din, csc_indptr, csc_indices, N, ly.din, dP, ly.din, stream));
      else
        GTS_TRY(gts_segmax_bwd_add(dNeigh, ly.din, arg, ly.din, N, ly.din, dP, ly.din, stream));
    }
    dp_cleared = false;
    {
      Prof prof(kProfTn, stream);
      GTS_TRY(gts_gemm_tn_colsum(dP, ly.din, h, ldh, g.dWp, ly.din, ly.din, ly.din, N, mode, g.dbp,
                                 at(workspace, pl.tn_ws[l]), pl.tn_ws_bytes[l], stream));
    }
    if (l > 0 || dfeats) {
      const float* WsT = at(workspace, pl.wsT[l]);
      const float* WpT = at(workspace, pl.wpT[l]);
      float* dh;
      int64_t lddh;
      if (l == 0) { dh = dfeats; lddh = lddf; }
      else { dh = at(workspace, ((n_layers - 1 - l) & 1) ? pl.g1 : pl.g0); lddh = ly.din; }
      // dh = (dZ Ws + dP' Wp) * (h > 0): h is the ReLU output of layer l-1 (no mask for the input features)
      const bool mask = l > 0 && layers[l - 1].relu != 0;
      // side job: clear the NEXT layer's dP (the other buffer; this GEMM reads dP[l & 1]) — see zero_in_gemm()
      void* zf = nullptr;
      size_t zf_bytes = 0;
      if (l > layer_lo && !det && pl.dP2 != pl.dP && N > 0 && ((size_t)N * layers[l - 1].din) % 4 == 0) {
        zf = at(workspace, ((l - 1) & 1) ? pl.dP2 : pl.dP);
        zf_bytes = sizeof(float) * (size_t)N * layers[l - 1].din;
        dp_cleared = true;
      }
      if (mask && pl.out_bits[l - 1])
        GTS_TRY(nt(dZ, ldz, ly.dout, WsT, ly.dout, dP, ly.din, ly.din, WpT, ly.din, nullptr, GTS_ACT_MASK_BITS, nullptr, 0,
                   dh, lddh, N, ly.din, mode, stream, nullptr, nullptr,
                   reinterpret_cast<const uint32_t*>(at(workspace, pl.out_bits[l - 1])), nullptr, nullptr, zf, zf_bytes));
      else
        GTS_TRY(nt(dZ, ldz, ly.dout, WsT, ly.dout, dP, ly.din, ly.din, WpT, ly.din, nullptr,
                   mask ? GTS_ACT_MASK_POS : GTS_ACT_NONE, mask ? h : nullptr, ldh, dh, lddh, N, ly.din, mode, stream,
                   nullptr, nullptr, nullptr, nullptr, nullptr, zf, zf_bytes));
    }
  }
  if (pl.defer) {              // every queued split-K reduction of the range in one launch, then the bias copies
    Prof prof(kProfTn, stream);
    GTS_TRY(launch_splitk_reduce_batch(reduce_batch, as_stream(stream)));
    for (int i = 0; i < n_bias_copy; ++i)
      GTS_CUDA(cudaMemcpyAsync(bias_copy[i].dst, bias_copy[i].src, bias_copy[i].bytes, cudaMemcpyDeviceToDevice, as_stream(stream)));
  }
  return GTS_OK;
}

int gts_sage_backward(const gts_sage_layer* layers, const gts_sage_layer_grads* grads, int32_t n_layers,
                      const int32_t* csc_indptr, const int32_t* csc_indices, int32_t n_nodes,
                      const float* feats, int64_t ldf, const float* dlogits, int64_t ldd,
                      float* dfeats, int64_t lddf,
                      void* workspace, size_t workspace_bytes, int32_t mode, gts_stream_t stream) {
  return backward_range(layers, grads, n_layers, n_layers, 0, csc_indptr, csc_indices, n_nodes, feats, ldf, dlogits, ldd,
                        dfeats, lddf, workspace, workspace_bytes, mode, stream);
}

int gts_sage_backward_range(const gts_sage_layer* layers, const gts_sage_layer_grads* grads, int32_t n_layers,
                            int32_t layer_hi, int32_t layer_lo,
                            const int32_t* csc_indptr, const int32_t* csc_indices, int32_t n_nodes,
                            const float* feats, int64_t ldf, const float* dlogits, int64_t ldd,
                            float* dfeats, int64_t lddf,
                            void* workspace, size_t workspace_bytes, int32_t mode, gts_stream_t stream) {
  return backward_range(layers, grads, n_layers, layer_hi, layer_lo, csc_indptr, csc_indices, n_nodes, feats, ldf, dlogits,
                        ldd, dfeats, lddf, workspace, workspace_bytes, mode, stream);
}

int gts_sage_step(const gts_sage_step_args* a, gts_stream_t stream) {
  GTS_CHECK_ARG(a && a->layers && a->grads && a->n_layers >= 1, "gts_sage_step: no layers");
  GTS_CHECK_ARG(a->n_nodes >= 1, "gts_sage_step: empty batch");
  GTS_CHECK_ARG(a->labels && a->class_w && a->sums && a->logits && a->workspace, "gts_sage_step: null pointer");
  GTS_CHECK_ARG(0 <= a->bwd_layer_lo && a->bwd_layer_lo <= a->n_layers, "gts_sage_step: bad bwd_layer_lo");
  StackPlan pl;
  GTS_CHECK_ARG(make_plan(a->layers, a->n_layers, a->n_nodes, true, a->mode, pl), "gts_sage_step: inconsistent layer dims");
  if (a->workspace_bytes < pl.total) {
    set_error("gts_sage_step: workspace %zu < required %zu", a->workspace_bytes, pl.total);
    return GTS_ERR_WORKSPACE;
  }
  const int C = a->layers[a->n_layers - 1].dout;
  float* dlogits = at(a->workspace, pl.dlogits);
  GTS_TRY(gts_sage_forward(a->layers, a->n_layers, a->indptr, a->indices, a->n_nodes, a->feats, a->ldf, a->logits, a->ldl,
                           a->workspace, a->workspace_bytes, 1, a->mode, stream));
  {
    Prof prof(kProfCe, stream);
    GTS_CUDA(cudaMemsetAsync(a->sums, 0, 2 * sizeof(float), as_stream(stream)));
    GTS_TRY(gts_ce_weighted(a->logits, a->ldl, a->labels, a->class_w, a->n_nodes, C, a->sums, dlogits, C, stream));
    if (a->normalize)      // gradients of the weighted MEAN (model/gnn_model.py:30,42): d/dz of sum(w nll) / sum(w)
      GTS_TRY(gts_scale_by_inv(dlogits, (int64_t)a->n_nodes * C, 1.0f, a->sums + 1, stream));
  }
  return backward_range(a->layers, a->grads, a->n_layers, a->n_layers, a->bwd_layer_lo, a->csc_indptr, a->csc_indices,
                        a->n_nodes, a->feats, a->ldf, dlogits, C, nullptr, 0, a->workspace, a->workspace_bytes, a->mode, stream);
}

int gts_sage_profile(int32_t enable) {
  for (auto& r : g_prof.recs) { g_prof.pool.push_back(r.e0); g_prof.pool.push_back(r.e1); }
  g_prof.recs.clear();
  g_prof.on = enable != 0;
  return GTS_OK;
}

int gts_sage_profile_read(float* ms_by_kind, int32_t* calls_by_kind, int32_t n_kinds) {
  GTS_CHECK_ARG(ms_by_kind && calls_by_kind && n_kinds >= kProfKinds, "gts_sage_profile_read: need %d slots", (int)kProfKinds);
  for (int k = 0; k < n_kinds; ++k) { ms_by_kind[k] = 0.f; calls_by_kind[k] = 0; }
  for (auto& r : g_prof.recs) {
    GTS_CUDA(cudaEventSynchronize(r.e1));
    float ms = 0.f;
    GTS_CUDA(cudaEventElapsedTime(&ms, r.e0, r.e1));
    ms_by_kind[r.kind] += ms;
    calls_by_kind[r.kind] += 1;
  }
  return GTS_OK;
}

int gts_sage_step_backward_rest(const gts_sage_step_args* a, int32_t layer_hi, int32_t layer_lo, gts_stream_t stream) {
  GTS_CHECK_ARG(a && a->layers && a->grads && a->n_layers >= 1, "gts_sage_step_backward_rest: no layers");
  StackPlan pl;
  GTS_CHECK_ARG(make_plan(a->layers, a->n_layers, a->n_nodes, true, a->mode, pl), "gts_sage_step_backward_rest: inconsistent layer dims");
  const int C = a->layers[a->n_layers - 1].dout;
  return backward_range(a->layers, a->grads, a->n_layers, layer_hi, layer_lo, a->csc_indptr, a->csc_indices, a->n_nodes,
                        a->feats, a->ldf, at(a->workspace, pl.dlogits), C, nullptr, 0, a->workspace, a->workspace_bytes,
                        a->mode, stream);
}

}  // extern "C"
