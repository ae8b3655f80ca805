// graph.cu — K6: batched edge lists -> canonical device CSR/CSC (int32).
//
// Replaces dgl.batch + dgl.from_networkx + DGL's lazy COO->CSR conversion
// (reference data_processing/data_loader.py:72,168; model/gnn_model.py:38).
// Result is bit-exact with the stable sort of edge ids by row
// (oracle/graph_ref.py:csr_by_dst_ref).
//
// Pipeline (all on the caller's stream, no host sync):
//   count (atomic histogram) -> exclusive scan (two-level) -> fill (atomic slot
//   claim, arbitrary order inside a row) -> per-row rank sort by edge id, which
//   restores the deterministic edge-id order.  HBM-bound integer work:
//   algorithmic bytes 4*(2E read + E write + 2(N+1)).
#include "common.cuh"

namespace gts {

constexpr int kScanThreads = 1024;
constexpr int kScanItems = 4;
constexpr int kScanTile = kScanThreads * kScanItems;

__global__ void batch_edges_kernel(const int32_t* __restrict__ src_l, const int32_t* __restrict__ dst_l,
                                   int64_t E, const int64_t* __restrict__ edge_off,
                                   const int32_t* __restrict__ node_off, int B,
                                   int32_t* __restrict__ src_g, int32_t* __restrict__ dst_g) {
  extern __shared__ unsigned char smem_raw[];
  int64_t* s_eoff = reinterpret_cast<int64_t*>(smem_raw);
  int32_t* s_noff = reinterpret_cast<int32_t*>(s_eoff + (B + 1));
  for (int i = threadIdx.x; i <= B; i += blockDim.x) {
    s_eoff[i] = edge_off[i];
    s_noff[i] = node_off[i];
  }
  __syncthreads();
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < E;
       e += (int64_t)gridDim.x * blockDim.x) {
    int lo = 0, hi = B;   // largest g with edge_off[g] <= e
    while (hi - lo > 1) {
      int mid = (lo + hi) >> 1;
      if (s_eoff[mid] <= e) lo = mid; else hi = mid;
    }
    int32_t off = s_noff[lo];
    src_g[e] = src_l[e] + off;
    dst_g[e] = dst_l[e] + off;
  }
}

__global__ void count_rows_kernel(const int32_t* __restrict__ row, int64_t E, int32_t N,
                                  int32_t* __restrict__ cnt) {
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < E;
       e += (int64_t)gridDim.x * blockDim.x) {
    int32_t r = row[e];
    if (r >= 0 && r < N) atomicAdd(&cnt[r], 1);
  }
}

// Level 1: per-tile exclusive scan of cnt -> indptr[0..N), tile totals -> tile_sum.
__global__ void scan_tiles_kernel(const int32_t* __restrict__ cnt, int32_t N,
                                  int32_t* __restrict__ out, int32_t* __restrict__ tile_sum) {
  __shared__ int32_t s_warp[32];
  const int tile = blockIdx.x;
  const int base = tile * kScanTile + threadIdx.x * kScanItems;
  int32_t v[kScanItems];
  int32_t t = 0;
#pragma unroll
  for (int i = 0; i < kScanItems; ++i) {
    v[i] = (base + i < N) ? cnt[base + i] : 0;
    t += v[i];
  }
  // inclusive warp scan of thread totals
  int32_t x = t;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int32_t y = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= o) x += y;
  }
  if (lane == 31) s_warp[warp] = x;
  __syncthreads();
  if (warp == 0) {
    int32_t w = s_warp[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int32_t y = __shfl_up_sync(0xffffffffu, w, o);
      if (lane >= o) w += y;
    }
    s_warp[lane] = w;   // inclusive over warps
  }
  __syncthreads();
  int32_t excl = x - t + (warp > 0 ? s_warp[warp - 1] : 0);
#pragma unroll
  for (int i = 0; i < kScanItems; ++i) {
    if (base + i < N) out[base + i] = excl;
    excl += v[i];
  }
  if (threadIdx.x == kScanThreads - 1) tile_sum[tile] = s_warp[31];
}

// Level 2: single block, sequential chunks: exclusive scan of tile sums in place,
// total -> *total_out.
__global__ void scan_tile_sums_kernel(int32_t* __restrict__ tile_sum, int32_t n_tiles,
                                      int32_t* __restrict__ total_out) {
  __shared__ int32_t s_warp[32];
  __shared__ int32_t s_carry;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int base = 0; base < n_tiles; base += blockDim.x) {
    int i = base + threadIdx.x;
    int32_t t = (i < n_tiles) ? tile_sum[i] : 0;
    int32_t x = t;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int32_t y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    if (lane == 31) s_warp[warp] = x;
    __syncthreads();
    if (warp == 0) {
      int32_t w = s_warp[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        int32_t y = __shfl_up_sync(0xffffffffu, w, o);
        if (lane >= o) w += y;
      }
      s_warp[lane] = w;
    }
    __syncthreads();
    int32_t carry = s_carry;
    int32_t excl = carry + x - t + (warp > 0 ? s_warp[warp - 1] : 0);
    if (i < n_tiles) tile_sum[i] = excl;
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) s_carry = carry + s_warp[31];
    __syncthreads();
  }
  if (threadIdx.x == 0) *total_out = s_carry;
}

// Level 3: add tile offsets; write indptr[N] = total.
__global__ void scan_add_kernel(int32_t* __restrict__ indptr, int32_t N,
                                const int32_t* __restrict__ tile_off, const int32_t* __restrict__ total) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N) indptr[i] += tile_off[i / kScanTile];
  if (i == 0) indptr[N] = *total;
}

__global__ void fill_rows_kernel(const int32_t* __restrict__ row, const int32_t* __restrict__ col,
                                 int64_t E, int32_t N, const int32_t* __restrict__ indptr,
                                 int32_t* __restrict__ cursor, int32_t* __restrict__ tmp_col,
                                 int32_t* __restrict__ tmp_eid) {
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < E;
       e += (int64_t)gridDim.x * blockDim.x) {
    int32_t r = row[e];
    if (r < 0 || r >= N) continue;
    int32_t p = indptr[r] + atomicAdd(&cursor[r], 1);
    tmp_col[p] = col[e];
    tmp_eid[p] = (int32_t)e;
  }
}

// One warp per row: rank each entry by edge id and write it to its final slot.
__global__ void sort_rows_kernel(const int32_t* __restrict__ indptr, int32_t N,
                                 const int32_t* __restrict__ tmp_col, const int32_t* __restrict__ tmp_eid,
                                 int32_t* __restrict__ indices, int32_t* __restrict__ eid_out) {
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  for (int64_t r = blockIdx.x * (int64_t)warps_per_block + (threadIdx.x >> 5); r < N;
       r += (int64_t)gridDim.x * warps_per_block) {
    const int32_t beg = indptr[r], end = indptr[r + 1];
    const int32_t deg = end - beg;
    if (deg <= 32) {
      int32_t my_e = (lane < deg) ? tmp_eid[beg + lane] : 0x7fffffff;
      int32_t my_c = (lane < deg) ? tmp_col[beg + lane] : 0;
      int rank = 0;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        int32_t oe = __shfl_sync(0xffffffffu, my_e, j);
        rank += (oe < my_e) ? 1 : 0;
      }
      if (lane < deg) {
        indices[beg + rank] = my_c;
        if (eid_out) eid_out[beg + rank] = my_e;
      }
    } else {
      for (int32_t i = lane; i < deg; i += 32) {
        const int32_t my_e = tmp_eid[beg + i];
        int rank = 0;
        for (int32_t j = 0; j < deg; ++j) rank += (tmp_eid[beg + j] < my_e) ? 1 : 0;
        indices[beg + rank] = tmp_col[beg + i];
        if (eid_out) eid_out[beg + rank] = my_e;
      }
    }
  }
}

__global__ void invert_perm_kernel(const int32_t* __restrict__ eid_csr, int64_t E, int32_t* __restrict__ pos_of_eid) {
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < E; p += (int64_t)gridDim.x * blockDim.x)
    pos_of_eid[eid_csr[p]] = (int32_t)p;
}
__global__ void compose_perm_kernel(const int32_t* __restrict__ eid_csc, const int32_t* __restrict__ pos_of_eid,
                                    int64_t E, int32_t* __restrict__ csc2csr) {
  for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < E; q += (int64_t)gridDim.x * blockDim.x)
    csc2csr[q] = pos_of_eid[eid_csc[q]];
}

static inline int grid_for(int64_t n, int threads, int waves = 8) {
  int64_t blocks = ceil_div<int64_t>(n, threads);
  int64_t cap = (int64_t)sm_count() * waves;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

struct CsrWs {
  int32_t* cursor;   // N (count, then cursor)
  int32_t* tile_sum; // n_tiles
  int32_t* total;    // 1
  int32_t* tmp_col;  // E
  int32_t* tmp_eid;  // E
  size_t bytes;
};

static CsrWs carve_csr_ws(void* ws, int64_t E, int32_t N) {
  CsrWs w;
  size_t off = 0;
  auto take = [&](size_t n_int) {
    int32_t* p = ws ? reinterpret_cast<int32_t*>(reinterpret_cast<char*>(ws) + off) : nullptr;
    off += align_up(n_int * sizeof(int32_t), 256);
    return p;
  };
  const int64_t n_tiles = ceil_div<int64_t>((int64_t)N > 0 ? N : 1, kScanTile);
  w.cursor = take((size_t)(N > 0 ? N : 1));
  w.tile_sum = take((size_t)n_tiles);
  w.total = take(1);
  w.tmp_col = take((size_t)(E > 0 ? E : 1));
  w.tmp_eid = take((size_t)(E > 0 ? E : 1));
  w.bytes = off;
  return w;
}

}  // namespace gts

using namespace gts;

extern "C" {

int gts_batch_edges(const int32_t* src_local, const int32_t* dst_local, int64_t n_edges,
                    const int64_t* edge_off, const int32_t* node_off, int32_t n_graphs,
                    int32_t* src_global, int32_t* dst_global, gts_stream_t stream) {
  GTS_CHECK_ARG(n_edges >= 0 && n_graphs >= 1, "gts_batch_edges: n_edges=%lld n_graphs=%d", (long long)n_edges, n_graphs);
  if (n_edges == 0) return GTS_OK;
  GTS_CHECK_ARG(src_local && dst_local && edge_off && node_off && src_global && dst_global, "gts_batch_edges: null pointer");
  GTS_CHECK_ARG(n_graphs <= 2048, "gts_batch_edges: at most 2048 graphs per batch (got %d)", n_graphs);
  const int threads = 256;
  size_t smem = (size_t)(n_graphs + 1) * (sizeof(int64_t) + sizeof(int32_t));
  batch_edges_kernel<<<grid_for(n_edges, threads), threads, smem, as_stream(stream)>>>(
      src_local, dst_local, n_edges, edge_off, node_off, n_graphs, src_global, dst_global);
  GTS_LAUNCH_CHECK();
  return GTS_OK;
}

size_t gts_csr_build_workspace_bytes(int64_t n_edges, int32_t n_nodes) {
  if (n_edges < 0 || n_nodes < 0) return 0;
  return carve_csr_ws(nullptr, n_edges, n_nodes).bytes;
}

int gts_csr_build(const int32_t* row, const int32_t* col, int64_t n_edges, int32_t n_nodes,
                  int32_t* indptr, int32_t* indices, int32_t* eid,
                  void* workspace, size_t workspace_bytes, gts_stream_t stream) {
  GTS_CHECK_ARG(n_edges >= 0 && n_nodes >= 0, "gts_csr_build: negative size");
  GTS_CHECK_ARG(n_edges < (int64_t)0x7fffffff, "gts_csr_build: n_edges must fit int32 offsets");
  GTS_CHECK_ARG(indptr != nullptr, "gts_csr_build: indptr is null");
  cudaStream_t st = as_stream(stream);
  if (n_nodes == 0 || n_edges == 0) {
    GTS_CUDA(cudaMemsetAsync(indptr, 0, sizeof(int32_t) * ((size_t)n_nodes + 1), st));
    return GTS_OK;
  }
  GTS_CHECK_ARG(row && col && indices && workspace, "gts_csr_build: null pointer");
  CsrWs w = carve_csr_ws(workspace, n_edges, n_nodes);
  if (workspace_bytes < w.bytes) {
    set_error("gts_csr_build: workspace %zu < required %zu", workspace_bytes, w.bytes);
    return GTS_ERR_WORKSPACE;
  }
  const int threads = 256;
  const int n_tiles = (int)ceil_div<int64_t>(n_nodes, kScanTile);
  GTS_CUDA(cudaMemsetAsync(w.cursor, 0, sizeof(int32_t) * (size_t)n_nodes, st));
  count_rows_kernel<<<grid_for(n_edges, threads), threads, 0, st>>>(row, n_edges, n_nodes, w.cursor);
  GTS_LAUNCH_CHECK();
  scan_tiles_kernel<<<n_tiles, kScanThreads, 0, st>>>(w.cursor, n_nodes, indptr, w.tile_sum);
  GTS_LAUNCH_CHECK();
  scan_tile_sums_kernel<<<1, 1024, 0, st>>>(w.tile_sum, n_tiles, w.total);
  GTS_LAUNCH_CHECK();
  scan_add_kernel<<<(int)ceil_div<int64_t>(n_nodes, 256), 256, 0, st>>>(indptr, n_nodes, w.tile_sum, w.total);
  GTS_LAUNCH_CHECK();
  GTS_CUDA(cudaMemsetAsync(w.cursor, 0, sizeof(int32_t) * (size_t)n_nodes, st));
  fill_rows_kernel<<<grid_for(n_edges, threads), threads, 0, st>>>(row, col, n_edges, n_nodes, indptr,
                                                                   w.cursor, w.tmp_col, w.tmp_eid);
  GTS_LAUNCH_CHECK();
  sort_rows_kernel<<<grid_for((int64_t)n_nodes * 32, threads), threads, 0, st>>>(indptr, n_nodes, w.tmp_col,
                                                                               w.tmp_eid, indices, eid);
  GTS_LAUNCH_CHECK();
  return GTS_OK;
}

int gts_edge_perm_compose(const int32_t* eid_csr, const int32_t* eid_csc, int64_t n_edges,
                          int32_t* scratch, int32_t* csc2csr, gts_stream_t stream) {
  GTS_CHECK_ARG(n_edges >= 0, "gts_edge_perm_compose: negative size");
  if (n_edges == 0) return GTS_OK;
  GTS_CHECK_ARG(eid_csr && eid_csc && scratch && csc2csr, "gts_edge_perm_compose: null pointer");
  const int threads = 256;
  invert_perm_kernel<<<grid_for(n_edges, threads), threads, 0, as_stream(stream)>>>(eid_csr, n_edges, scratch);
  GTS_LAUNCH_CHECK();
  compose_perm_kernel<<<grid_for(n_edges, threads), threads, 0, as_stream(stream)>>>(eid_csc, scratch, n_edges, csc2csr);
  GTS_LAUNCH_CHECK();
  return GTS_OK;
}

}  // extern "C"
