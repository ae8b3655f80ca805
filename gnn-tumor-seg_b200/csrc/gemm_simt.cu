// gemm_simt.cu — fp32 SIMT (FFMA) contractions: the exact-fp32 arithmetic mode
// (GTS_GEMM_FP32) of K1/K3/K4 and their backward, and the shape fallback for
// the tcgen05 path (gemm_tcgen05.cu).  Also column sums and small transposes.
//
// Replaces nn.Linear inside DGL SAGEConv/GATConv and its autograd (invoked at
// reference model/networks.py:35,63,65).
#include "common.cuh"

namespace gts {

int gemm_nt_tcgen05(const gts_gemm_nt_args* a, cudaStream_t st);   // gemm_tcgen05.cu
int gemm_tn_tcgen05(const float* A, int64_t lda, const float* B, int64_t ldb, float* C, int64_t ldc,
                    int32_t Mo, int32_t No, int64_t K, int32_t mode, float* colsum_out, void* ws, size_t ws_bytes,
                    cudaStream_t st);
bool gemm_tn_tcgen05_fuses_colsum(int32_t mode);
bool gemm_tn2_tcgen05_supported(const float* A, int64_t lda, const float* B1, int64_t ldb1, const float* B2, int64_t ldb2,
                                int32_t Mo, int32_t No, int64_t K, int32_t mode);
size_t gemm_tn2_tcgen05_ws(int32_t Mo, int32_t No, int64_t K);
int gemm_tn2_tcgen05(const float* A, int64_t lda, const float* B1, int64_t ldb1, const float* B2, int64_t ldb2,
                     float* C1, float* C2, int64_t ldc, int32_t Mo, int32_t No, int64_t K, float* colsum_out,
                     void* ws, size_t ws_bytes, cudaStream_t st);
size_t gemm_tn_tcgen05_ws(int32_t Mo, int32_t No, int64_t K, int32_t mode);
bool gemm_nt_tcgen05_supported(const gts_gemm_nt_args* a);
bool gemm_nt_bits_supported(int32_t M, int32_t N, int32_t mode);
bool gemm_tn_tcgen05_supported(const float* A, int64_t lda, const float* B, int64_t ldb, int32_t Mo, int32_t No, int64_t K);

constexpr int kSimtThreads = 256;
constexpr int kBK = 16;

// One operand source of the K loop: `rows` x K, element (r,k) at
// base[r*ld + k] (K-major) or base[k*ld + r] (MN-major).
struct Operand {
  const float* base;
  int64_t ld;
};

template <int BR, bool KMAJOR>
__device__ __forceinline__ void load_tile(float (*dst)[BR + 4], const Operand& op, int64_t r0, int64_t rows,
                                          int64_t k0, int64_t k_end) {
  // dst[k][r] for k in [0,kBK), r in [0,BR)
  constexpr int ELEMS = BR * kBK;
  static_assert(ELEMS % kSimtThreads == 0, "tile must divide evenly");
#pragma unroll
  for (int i = 0; i < ELEMS / kSimtThreads; ++i) {
    const int e = i * kSimtThreads + threadIdx.x;
    int r, k;
    if (KMAJOR) { k = e % kBK; r = e / kBK; }       // consecutive threads walk along K (contiguous)
    else        { r = e % BR;  k = e / BR;  }       // consecutive threads walk along rows (contiguous)
    const int64_t gr = r0 + r, gk = k0 + k;
    float v = 0.f;
    if (gr < rows && gk < k_end)
      v = KMAJOR ? __ldg(op.base + gr * op.ld + gk) : __ldg(op.base + gk * op.ld + gr);
    dst[k][r] = v;
  }
}

// C tile = sum over sources s of A_s * B_s^T over K_s.  A_KMAJOR/B_KMAJOR select
// the operand layouts: NT GEMM (forward / data-gradient) has both K-major; the
// weight-gradient (TN) form has both MN-major and a single source.
// SPLITK: blockIdx.z selects a K range and the tile is written raw to a
// partial buffer C + z*split_stride.
template <int BM, int BN, int TM, int TN, bool A_KMAJOR, bool B_KMAJOR, bool SPLITK>
__global__ void __launch_bounds__(kSimtThreads)
gemm_simt_kernel(Operand A1, Operand B1, int64_t K1, Operand A2, Operand B2, int64_t K2,
                 int64_t M, int64_t N, float* __restrict__ C, int64_t ldc,
                 const float* __restrict__ bias, const float* __restrict__ bias2, const float* __restrict__ aux, int64_t ldaux, int act,
                 int64_t k_per_split, int64_t split_stride,
                 const int32_t* __restrict__ sidx = nullptr, int64_t ld_sidx = 0, float* __restrict__ sout = nullptr,
                 int64_t ld_sout = 0) {
  static_assert((BM / TM) * (BN / TN) == kSimtThreads, "thread tiling must cover the block tile");
  __shared__ float As[kBK][BM + 4];
  __shared__ float Bs[kBK][BN + 4];
  const int tx = threadIdx.x % (BN / TN);
  const int ty = threadIdx.x / (BN / TN);
  const int64_t m0 = (int64_t)blockIdx.y * BM;
  const int64_t n0 = (int64_t)blockIdx.x * BN;

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

#pragma unroll 1
  for (int s = 0; s < 2; ++s) {
    const Operand& A = s == 0 ? A1 : A2;
    const Operand& B = s == 0 ? B1 : B2;
    const int64_t K = s == 0 ? K1 : K2;
    if (K <= 0 || A.base == nullptr) continue;
    int64_t kb = 0, ke = K;
    if (SPLITK) {
      kb = (int64_t)blockIdx.z * k_per_split;
      ke = kb + k_per_split < K ? kb + k_per_split : K;
    }
#pragma unroll 1
    for (int64_t k0 = kb; k0 < ke; k0 += kBK) {
      load_tile<BM, A_KMAJOR>(As, A, m0, M, k0, ke);
      load_tile<BN, B_KMAJOR>(Bs, B, n0, N, k0, ke);
      __syncthreads();
#pragma unroll
      for (int k = 0; k < kBK; ++k) {
        float a[TM], b[TN];
#pragma unroll
        for (int i = 0; i < TM; ++i) a[i] = As[k][ty + i * (BM / TM)];
#pragma unroll
        for (int j = 0; j < TN; ++j) b[j] = Bs[k][tx + j * (BN / TN)];
#pragma unroll
        for (int i = 0; i < TM; ++i)
#pragma unroll
          for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
      }
      __syncthreads();
    }
  }

  float* Cout = C + (SPLITK ? (int64_t)blockIdx.z * split_stride : 0);
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int64_t m = m0 + ty + i * (BM / TM);
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int64_t n = n0 + tx + j * (BN / TN);
      if (n >= N) continue;
      float v = acc[i][j];
      if (!SPLITK) {
        if (bias) v += bias[n];
        if (bias2) v += bias2[n];
        if (act == GTS_ACT_RELU) v = fmaxf(v, 0.f);
        else if (act == GTS_ACT_MASK_POS || act == GTS_ACT_MASK_POS_SCATTER) v = (aux[m * ldaux + n] > 0.f) ? v : 0.f;
        if (act == GTS_ACT_MASK_POS_SCATTER) {
          const int32_t u = sidx[m * ld_sidx + n];
          if (u >= 0 && v != 0.f) atomicAdd(sout + (int64_t)u * ld_sout + n, v);
          continue;
        }
      }
      Cout[m * ldc + n] = v;
    }
  }
}

// C[i] = sum_z partial[z][i]  (deterministic order)
__global__ void splitk_reduce_kernel(const float* __restrict__ partial, int64_t split_stride, int splits,
                                     int64_t rows, int64_t cols, float* __restrict__ C, int64_t ldc) {
  const int64_t total = rows * cols;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    float s = 0.f;
    for (int z = 0; z < splits; ++z) s += partial[(int64_t)z * split_stride + i];
    const int64_t r = i / cols, c = i - r * cols;
    C[r * ldc + c] = s;
  }
}

// float4 variant: rows*cols % 4 == 0, contiguous C (ldc == cols), 16-byte aligned.
__global__ void splitk_reduce_vec_kernel(const float4* __restrict__ partial, int64_t split_stride4, int splits,
                                         int64_t total4, float4* __restrict__ C) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    int z = 0;
    for (; z + 8 <= splits; z += 8) {
      float4 v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = ldg_nc_na(partial + (int64_t)(z + j) * split_stride4 + i);
#pragma unroll
      for (int j = 0; j < 8; ++j) { s.x += v[j].x; s.y += v[j].y; s.z += v[j].z; s.w += v[j].w; }
    }
    for (; z < splits; ++z) {
      const float4 v = ldg_nc_na(partial + (int64_t)z * split_stride4 + i);
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
    C[i] = s;
  }
}

// Wide variant: 128 threads = 32 consecutive float4 outputs x 4 split groups (warp g sums the splits z = g, g+4, ...,
// all loads in flight), then warp 0 adds the four partial sums in a fixed order -> deterministic, 4x the parallelism
// of one thread per output (a 256x256 weight gradient has only 16 384 float4 outputs).  Outputs [0, main4) go to C,
// outputs [main4, total4) to `tail` (the fused column sums of the weight-gradient GEMM).
__device__ __forceinline__ void splitk_reduce_wide_body(const float4* __restrict__ partial, int64_t split_stride4, int splits,
                                                        int64_t main4, int64_t total4, float4* __restrict__ C,
                                                        float4* __restrict__ tail, int64_t block) {
  __shared__ float4 red[3][32];
  const int o = threadIdx.x & 31, g = threadIdx.x >> 5;
  const int64_t i = block * 32 + o;
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  if (i < total4) {
    int z = g;
    for (; z + 12 < splits; z += 16) {
      float4 v[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] = ldg_nc_na(partial + (int64_t)(z + 4 * j) * split_stride4 + i);
#pragma unroll
      for (int j = 0; j < 4; ++j) { s.x += v[j].x; s.y += v[j].y; s.z += v[j].z; s.w += v[j].w; }
    }
    for (; z < splits; z += 4) {
      const float4 v = ldg_nc_na(partial + (int64_t)z * split_stride4 + i);
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
  }
  if (g > 0) red[g - 1][o] = s;
  __syncthreads();
  if (g == 0 && i < total4) {
#pragma unroll
    for (int j = 0; j < 3; ++j) { const float4 v = red[j][o]; s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w; }
    if (i < main4) C[i] = s;
    else tail[i - main4] = s;
  }
}

__global__ void __launch_bounds__(128)
splitk_reduce_wide_kernel(const float4* __restrict__ partial, int64_t split_stride4, int splits, int64_t main4,
                          int64_t total4, float4* __restrict__ C, float4* __restrict__ tail) {
  splitk_reduce_wide_body(partial, split_stride4, splits, main4, total4, C, tail, blockIdx.x);
}

// Every queued reduction of a backward range in one launch: block -> (job, block of that job), same body.
__global__ void __launch_bounds__(128) splitk_reduce_batch_kernel(const ReduceBatch b) {
  int j = 0;
  while (j + 1 < b.n && (int)blockIdx.x >= b.job[j + 1].blk0) ++j;
  const ReduceJob& q = b.job[j];
  splitk_reduce_wide_body(q.partial, q.stride4, q.splits, q.main4, q.total4, q.C, q.tail, (int64_t)blockIdx.x - q.blk0);
}

// Column sums, wide fast path (cols % 4 == 0, cols <= 1024, 16-byte aligned rows):
// block = 256 threads = (cols/4 column groups) x (256/(cols/4) row lanes); float4 loads.
__global__ void __launch_bounds__(256)
colsum_partial_vec_kernel(const float* __restrict__ A, int64_t lda, int64_t rows, int cols4, int64_t rows_per_block,
                          float* __restrict__ partial) {
  extern __shared__ float4 s_red[];                // [row_lanes][cols4]
  const int row_lanes = 256 / cols4;
  const int cg = threadIdx.x % cols4, rl = threadIdx.x / cols4;
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_block;
  const int64_t r1 = r0 + rows_per_block < rows ? r0 + rows_per_block : rows;
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  if (rl < row_lanes) {
    int64_t r = r0 + rl;
    for (; r + 3 * row_lanes < r1; r += 4 * row_lanes) {
      float4 v[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] = ldg_nc_na(reinterpret_cast<const float4*>(A + (r + j * row_lanes) * lda) + cg);
#pragma unroll
      for (int j = 0; j < 4; ++j) { s.x += v[j].x; s.y += v[j].y; s.z += v[j].z; s.w += v[j].w; }
    }
    for (; r < r1; r += row_lanes) {
      const float4 v = ldg_nc_na(reinterpret_cast<const float4*>(A + r * lda) + cg);
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
    s_red[rl * cols4 + cg] = s;
  }
  __syncthreads();
  if (rl == 0) {
    for (int j = 1; j < row_lanes; ++j) {
      const float4 v = s_red[j * cols4 + cg];
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
    reinterpret_cast<float4*>(partial + (int64_t)blockIdx.x * cols4 * 4)[cg] = s;
  }
}

// final: out[c] = sum_b partial[b][c]; one block of 256 threads per 32 columns (8 block-lanes x 32 columns)
__global__ void __launch_bounds__(256)
colsum_final_wide_kernel(const float* __restrict__ partial, int n_blocks, int32_t cols, float* __restrict__ out) {
  __shared__ float red[8][33];
  const int cx = threadIdx.x & 31, by = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cx;
  float s = 0.f;
  if (c < cols)
    for (int b = by; b < n_blocks; b += 8) s += partial[(int64_t)b * cols + c];
  red[by][cx] = s;
  __syncthreads();
  if (by == 0 && c < cols) {
    float t = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) t += red[j][cx];
    out[c] = t;
  }
}

// Column sums, pass 1: block b sums rows [b*rows_per_block, ...) of every column.
__global__ void colsum_partial_kernel(const float* __restrict__ A, int64_t lda, int64_t rows, int32_t cols,
                                      int64_t rows_per_block, float* __restrict__ partial) {
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_block;
  const int64_t r1 = r0 + rows_per_block < rows ? r0 + rows_per_block : rows;
  // threads: x over columns (coalesced), y over row lanes
  __shared__ float red[8][33];
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;   // 32 x 8
  for (int c0 = 0; c0 < cols; c0 += 32) {
    const int c = c0 + cx;
    float s = 0.f;
    if (c < cols)
      for (int64_t r = r0 + ry; r < r1; r += 8) s += A[r * lda + c];
    red[ry][cx] = s;
    __syncthreads();
    if (ry == 0 && c < cols) {
      float t = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) t += red[j][cx];
      partial[(int64_t)blockIdx.x * cols + c] = t;
    }
    __syncthreads();
  }
}

__global__ void colsum_final_kernel(const float* __restrict__ partial, int n_blocks, int32_t cols, float* __restrict__ out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cols) return;
  float s = 0.f;
  for (int b = 0; b < n_blocks; ++b) s += partial[(int64_t)b * cols + c];
  out[c] = s;
}

__global__ void transpose_kernel(const float* __restrict__ in, int64_t ldin, int32_t rows, int32_t cols,
                                 float* __restrict__ out, int64_t ldout) {
  __shared__ float t[32][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  for (int j = threadIdx.y; j < 32; j += 8) {
    const int r = blockIdx.y * 32 + j;
    if (r < rows && c < cols) t[j][threadIdx.x] = in[(int64_t)r * ldin + c];
  }
  __syncthreads();
  const int r2 = blockIdx.y * 32 + threadIdx.x;   // column of out
  for (int j = threadIdx.y; j < 32; j += 8) {
    const int c2 = blockIdx.x * 32 + j;           // row of out
    if (c2 < cols && r2 < rows) out[(int64_t)c2 * ldout + r2] = t[threadIdx.x][j];
  }
}

// partial[z] = [rows*cols product | tail_len extra floats]; C contiguous; everything a multiple of 4 floats and
// 16-byte aligned (the caller checks with splitk_reduce_fused_ok)
bool splitk_reduce_fused_ok(int64_t rows, int64_t cols, int64_t ldc, int64_t tail_len, int64_t split_stride,
                            const void* partial, const void* C, const void* tail) {
  auto al = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15u) == 0; };
  return ldc == cols && (rows * cols) % 4 == 0 && tail_len % 4 == 0 && split_stride % 4 == 0 && al(partial) && al(C) &&
         (tail_len == 0 || al(tail));
}
static thread_local ReduceBatch* g_defer = nullptr;
void splitk_defer_set(ReduceBatch* batch) { g_defer = batch; }
ReduceBatch* splitk_defer_target() { return g_defer; }

int launch_splitk_reduce_batch(ReduceBatch& b, cudaStream_t st) {
  if (b.n > 0) {
    splitk_reduce_batch_kernel<<<b.blocks, 128, 0, st>>>(b);
    b.n = 0; b.blocks = 0;
    GTS_LAUNCH_CHECK();
  }
  return GTS_OK;
}

void launch_splitk_reduce_fused(const float* partial, int64_t split_stride, int splits, int64_t rows, int64_t cols,
                                float* C, float* tail, int64_t tail_len, cudaStream_t st) {
  const int64_t main4 = rows * cols / 4, total4 = main4 + tail_len / 4;
  const int blocks = (int)ceil_div<int64_t>(total4, 32);
  if (ReduceBatch* q = g_defer) {                       // queued: the owner of the batch launches it
    if (q->n == ReduceBatch::kMaxJobs) launch_splitk_reduce_batch(*q, st);
    ReduceJob& j = q->job[q->n++];
    j.partial = reinterpret_cast<const float4*>(partial); j.C = reinterpret_cast<float4*>(C);
    j.tail = reinterpret_cast<float4*>(tail);
    j.stride4 = split_stride / 4; j.main4 = main4; j.total4 = total4; j.splits = splits; j.blk0 = q->blocks;
    q->blocks += blocks;
    return;
  }
  splitk_reduce_wide_kernel<<<blocks, 128, 0, st>>>(reinterpret_cast<const float4*>(partial), split_stride / 4, splits, main4,
                                                    total4, reinterpret_cast<float4*>(C), reinterpret_cast<float4*>(tail));
}

// All weight transposes of one backward pass in ONE launch: blockIdx.z picks the matrix.
__global__ void transpose_batch_kernel(const TransposeBatch b) {
  const TransposeJob j = b.job[blockIdx.z];
  if ((int)blockIdx.x * 32 >= j.cols || (int)blockIdx.y * 32 >= j.rows) return;
  __shared__ float t[32][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int r = blockIdx.y * 32 + i;
    if (r < j.rows && c < j.cols) t[i][threadIdx.x] = j.in[(int64_t)r * j.ldin + c];
  }
  __syncthreads();
  const int r2 = blockIdx.y * 32 + threadIdx.x;
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int c2 = blockIdx.x * 32 + i;
    if (c2 < j.cols && r2 < j.rows) j.out[(int64_t)c2 * j.ldout + r2] = t[threadIdx.x][i];
  }
}

int launch_transpose_batch(const TransposeBatch& b, cudaStream_t st) {
  if (b.n <= 0) return GTS_OK;
  int max_r = 0, max_c = 0;
  for (int i = 0; i < b.n; ++i) { max_r = b.job[i].rows > max_r ? b.job[i].rows : max_r; max_c = b.job[i].cols > max_c ? b.job[i].cols : max_c; }
  if (max_r == 0 || max_c == 0) return GTS_OK;
  dim3 grid((max_c + 31) / 32, (max_r + 31) / 32, b.n), block(32, 8);
  transpose_batch_kernel<<<grid, block, 0, st>>>(b);
  GTS_LAUNCH_CHECK();
  return GTS_OK;
}

void launch_splitk_reduce(const float* partial, int64_t split_stride, int splits, int64_t rows, int64_t cols,
                          float* C, int64_t ldc, cudaStream_t st) {
  const int64_t total = rows * cols;
  if (splits >= 8 && splitk_reduce_fused_ok(rows, cols, ldc, 0, split_stride, partial, C, nullptr)) {
    launch_splitk_reduce_fused(partial, split_stride, splits, rows, cols, C, nullptr, 0, st);
    return;
  }
  const bool vec = ldc == cols && total % 4 == 0 && split_stride % 4 == 0 &&
                   (reinterpret_cast<uintptr_t>(partial) & 15u) == 0 && (reinterpret_cast<uintptr_t>(C) & 15u) == 0;
  if (vec) {
    int blocks = (int)ceil_div<int64_t>(total / 4, 128);
    if (blocks < 1) blocks = 1;
    splitk_reduce_vec_kernel<<<blocks, 128, 0, st>>>(reinterpret_cast<const float4*>(partial), split_stride / 4, splits,
                                                     total / 4, reinterpret_cast<float4*>(C));
    return;
  }
  int blocks = (int)ceil_div<int64_t>(total, 256);
  if (blocks < 1) blocks = 1;
  splitk_reduce_kernel<<<blocks, 256, 0, st>>>(partial, split_stride, splits, rows, cols, C, ldc);
}

static int colsum_blocks(int64_t rows) {
  int64_t b = ceil_div<int64_t>(rows, 128);
  const int64_t cap = (int64_t)sm_count() * 8;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

static int tn_splits(int tiles, int64_t K) {
  int64_t want = ceil_div<int64_t>((int64_t)sm_count() * 2, tiles);
  int64_t max_by_k = K / 256;
  if (max_by_k < 1) max_by_k = 1;
  if (want > max_by_k) want = max_by_k;
  if (want < 1) want = 1;
  return (int)want;
}

static size_t gemm_tn_simt_ws(int32_t Mo, int32_t No, int64_t K) {
  const int bm = Mo <= 16 ? 16 : 128;
  const int tiles = (int)(ceil_div<int64_t>(Mo, bm) * ceil_div<int64_t>(No, 128));
  const int splits = tn_splits(tiles, K);
  return align_up((size_t)splits * (size_t)Mo * (size_t)No * sizeof(float), 256);
}

static int gemm_nt_simt(const gts_gemm_nt_args* a, cudaStream_t st) {
  Operand A1{a->A1, a->lda1}, B1{a->B1, a->ldb1}, A2{a->A2, a->lda2}, B2{a->B2, a->ldb2};
  const int64_t K2 = (a->A2 && a->B2) ? a->K2 : 0;
  if (a->N <= 16) {
    dim3 grid((unsigned)ceil_div<int64_t>(a->N, 16), (unsigned)ceil_div<int64_t>(a->M, 128));
    gemm_simt_kernel<128, 16, 8, 1, true, true, false><<<grid, kSimtThreads, 0, st>>>(
        A1, B1, a->K1, A2, B2, K2, a->M, a->N, a->C, a->ldc, a->bias, a->bias2, a->aux, a->ldaux, a->act, 0, 0,
        a->scatter_idx, a->ld_idx, a->scatter_out, a->ld_out);
  } else {
    dim3 grid((unsigned)ceil_div<int64_t>(a->N, 128), (unsigned)ceil_div<int64_t>(a->M, 128));
    gemm_simt_kernel<128, 128, 8, 8, true, true, false><<<grid, kSimtThreads, 0, st>>>(
        A1, B1, a->K1, A2, B2, K2, a->M, a->N, a->C, a->ldc, a->bias, a->bias2, a->aux, a->ldaux, a->act, 0, 0,
        a->scatter_idx, a->ld_idx, a->scatter_out, a->ld_out);
  }
  GTS_LAUNCH_CHECK();
  return GTS_OK;
}

static int gemm_tn_simt(const float* A, int64_t lda, const float* B, int64_t ldb, float* C, int64_t ldc,
                        int32_t Mo, int32_t No, int64_t K, void* ws, size_t ws_bytes, cudaStream_t st) {
  const size_t need = gemm_tn_simt_ws(Mo, No, K);
  if (ws_bytes < need || ws == nullptr) {
    set_error("gts_gemm_tn: workspace %zu < required %zu", ws_bytes, need);
    return GTS_ERR_WORKSPACE;
  }
  Operand Ao{A, lda}, Bo{B, ldb}, none{nullptr, 0};
  const bool small_m = Mo <= 16;
  const int bm = small_m ? 16 : 128;
  const int tiles_m = (int)ceil_div<int64_t>(Mo, bm), tiles_n = (int)ceil_div<int64_t>(No, 128);
  const int splits = tn_splits(tiles_m * tiles_n, K);
  int64_t kps = ceil_div<int64_t>(K, splits);
  kps = ceil_div<int64_t>(kps, kBK) * kBK;
  float* partial = reinterpret_cast<float*>(ws);
  const int64_t stride = (int64_t)Mo * No;
  dim3 grid(tiles_n, tiles_m, splits);
  if (small_m)
    gemm_simt_kernel<16, 128, 1, 8, false, false, true><<<grid, kSimtThreads, 0, st>>>(
        Ao, Bo, K, none, none, 0, Mo, No, partial, No, nullptr, nullptr, nullptr, 0, 0, kps, stride);
  else
    gemm_simt_kernel<128, 128, 8, 8, false, false, true><<<grid, kSimtThreads, 0, st>>>(
        Ao, Bo, K, none, none, 0, Mo, No, partial, No, nullptr, nullptr, nullptr, 0, 0, kps, stride);
  GTS_LAUNCH_CHECK();
  const int64_t total = stride;
  int blocks = (int)ceil_div<int64_t>(total, 256);
  splitk_reduce_kernel<<<blocks, 256, 0, st>>>(partial, stride, splits, Mo, No, C, ldc);
  GTS_LAUNCH_CHECK();
  return GTS_OK;
}

}  // namespace gts

using namespace gts;

extern "C" {

int gts_gemm_nt(const gts_gemm_nt_args* a, gts_stream_t stream) {
  GTS_CHECK_ARG(a != nullptr, "gts_gemm_nt: args is null");
  GTS_CHECK_ARG(a->M >= 0 && a->N >= 0 && a->K1 >= 0 && a->K2 >= 0, "gts_gemm_nt: negative size");
  if (a->M == 0 || a->N == 0) return GTS_OK;
  GTS_CHECK_ARG(a->C != nullptr || a->act == GTS_ACT_MASK_POS_SCATTER || a->act == GTS_ACT_MASK_BITS_SCATTER, "gts_gemm_nt: C is null");
  GTS_CHECK_ARG(a->act != GTS_ACT_MASK_POS_SCATTER || (a->scatter_idx && a->scatter_out && a->aux),
                "gts_gemm_nt: GTS_ACT_MASK_POS_SCATTER needs aux, scatter_idx and scatter_out");
  GTS_CHECK_ARG(a->K1 == 0 || (a->A1 && a->B1), "gts_gemm_nt: A1/B1 null with K1 > 0");
  GTS_CHECK_ARG(a->act >= GTS_ACT_NONE && a->act <= GTS_ACT_MASK_BITS_SCATTER, "gts_gemm_nt: unknown act %d", a->act);
  GTS_CHECK_ARG(a->act != GTS_ACT_MASK_BITS_SCATTER || (a->scatter_idx && a->scatter_out && a->aux_bits),
                "gts_gemm_nt: GTS_ACT_MASK_BITS_SCATTER needs aux_bits, scatter_idx and scatter_out");
  GTS_CHECK_ARG(a->act != GTS_ACT_MASK_POS || a->aux != nullptr, "gts_gemm_nt: GTS_ACT_MASK_POS needs aux");
  const bool bits = a->act == GTS_ACT_MASK_BITS || a->act == GTS_ACT_MASK_BITS_SCATTER || a->relu_bits_out != nullptr;
  if (bits && !(gemm_nt_bits_supported(a->M, a->N, a->mode) && gemm_nt_tcgen05_supported(a))) {
    set_error("gts_gemm_nt: bit-matrix masks (GTS_ACT_MASK_BITS / relu_bits_out) are not supported for M=%d N=%d mode=%d "
              "(see gts_gemm_nt_bits_supported)", a->M, a->N, a->mode);
    return GTS_ERR_UNSUPPORTED;
  }
  cudaStream_t st = as_stream(stream);
  if (a->zero_fill && a->zero_fill_bytes) {
    GTS_CHECK_ARG((reinterpret_cast<uintptr_t>(a->zero_fill) & 15u) == 0 && a->zero_fill_bytes % 16 == 0, "gts_gemm_nt: zero_fill must be 16-byte aligned and sized");
    // the 256-wide kernel clears it with its spare warps (same condition as its dispatch); anything else: memset first
    const bool wide = a->mode == GTS_GEMM_TF32X3 && gemm_nt_bits_supported(a->M, a->N, a->mode) && gemm_nt_tcgen05_supported(a);
    if (!wide) GTS_CUDA(cudaMemsetAsync(a->zero_fill, 0, a->zero_fill_bytes, st));
  }
  if (a->mode == GTS_GEMM_FP32) return gemm_nt_simt(a, st);
  if (a->mode == GTS_GEMM_TF32 || a->mode == GTS_GEMM_TF32X3) {
    if (gemm_nt_tcgen05_supported(a)) return gemm_nt_tcgen05(a, st);
    return gemm_nt_simt(a, st);   // shape outside the tensor-core tiling: exact fp32 SIMT kernel
  }
  set_error("gts_gemm_nt: unknown mode %d", a->mode);
  return GTS_ERR_INVALID;
}

int gts_gemm_nt_bits_supported(int32_t M, int32_t N, int32_t mode) { return gemm_nt_bits_supported(M, N, mode) ? 1 : 0; }

size_t gts_gemm_tn_workspace_bytes(int32_t Mo, int32_t No, int64_t K, int32_t mode) {
  if (Mo <= 0 || No <= 0 || K <= 0) return 256;
  size_t s = gemm_tn_simt_ws(Mo, No, K);
  if (mode != GTS_GEMM_FP32) {
    size_t t = gemm_tn_tcgen05_ws(Mo, No, K, mode);
    if (t > s) s = t;
  }
  return s;
}

int gts_gemm_tn(const float* A, int64_t lda, const float* B, int64_t ldb,
                float* C, int64_t ldc, int32_t Mo, int32_t No, int64_t K, int32_t mode,
                void* workspace, size_t workspace_bytes, gts_stream_t stream) {
  GTS_CHECK_ARG(Mo >= 0 && No >= 0 && K >= 0, "gts_gemm_tn: negative size");
  if (Mo == 0 || No == 0) return GTS_OK;
  GTS_CHECK_ARG(C != nullptr, "gts_gemm_tn: C is null");
  cudaStream_t st = as_stream(stream);
  if (K == 0) {
    GTS_CUDA(cudaMemset2DAsync(C, sizeof(float) * ldc, 0, sizeof(float) * No, Mo, st));
    return GTS_OK;
  }
  GTS_CHECK_ARG(A && B, "gts_gemm_tn: null operand");
  GTS_CHECK_ARG(mode >= GTS_GEMM_FP32 && mode <= GTS_GEMM_TF32X3, "gts_gemm_tn: unknown mode %d", mode);
  if (mode != GTS_GEMM_FP32 && gemm_tn_tcgen05_supported(A, lda, B, ldb, Mo, No, K))
    return gemm_tn_tcgen05(A, lda, B, ldb, C, ldc, Mo, No, K, mode, nullptr, workspace, workspace_bytes, st);
  return gemm_tn_simt(A, lda, B, ldb, C, ldc, Mo, No, K, workspace, workspace_bytes, st);
}

size_t gts_gemm_tn_colsum_workspace_bytes(int32_t Mo, int32_t No, int64_t K, int32_t mode) {
  return gts_gemm_tn_workspace_bytes(Mo, No, K, mode) + gts_colsum_workspace_bytes(K, Mo);
}

int gts_gemm_tn_colsum(const float* A, int64_t lda, const float* B, int64_t ldb,
                       float* C, int64_t ldc, int32_t Mo, int32_t No, int64_t K, int32_t mode,
                       float* colsum_out, void* workspace, size_t workspace_bytes, gts_stream_t stream) {
  GTS_CHECK_ARG(Mo >= 0 && No >= 0 && K >= 0, "gts_gemm_tn_colsum: negative size");
  GTS_CHECK_ARG(colsum_out != nullptr || Mo == 0, "gts_gemm_tn_colsum: colsum_out is null");
  GTS_CHECK_ARG(mode >= GTS_GEMM_FP32 && mode <= GTS_GEMM_TF32X3, "gts_gemm_tn_colsum: unknown mode %d", mode);
  const size_t tn_bytes = gts_gemm_tn_workspace_bytes(Mo, No, K, mode);
  const size_t need = tn_bytes + gts_colsum_workspace_bytes(K, Mo);
  if (workspace_bytes < need || (need > 0 && workspace == nullptr)) {
    set_error("gts_gemm_tn_colsum: workspace %zu < required %zu", workspace_bytes, need);
    return GTS_ERR_WORKSPACE;
  }
  if (Mo > 0 && No > 0 && K > 0 && C && A && B && gemm_tn_tcgen05_fuses_colsum(mode) &&
      gemm_tn_tcgen05_supported(A, lda, B, ldb, Mo, No, K))
    return gemm_tn_tcgen05(A, lda, B, ldb, C, ldc, Mo, No, K, mode, colsum_out, workspace, tn_bytes, as_stream(stream));
  int rc = gts_gemm_tn(A, lda, B, ldb, C, ldc, Mo, No, K, mode, workspace, tn_bytes, stream);
  if (rc != GTS_OK) return rc;
  return gts_colsum(A, lda, K, Mo, colsum_out, reinterpret_cast<char*>(workspace) + tn_bytes, workspace_bytes - tn_bytes, stream);
}

size_t gts_gemm_tn2_colsum_workspace_bytes(int32_t Mo, int32_t No, int64_t K, int32_t mode) {
  size_t a = gts_gemm_tn_colsum_workspace_bytes(Mo, No, K, mode);
  size_t b = mode == GTS_GEMM_TF32X3 ? gemm_tn2_tcgen05_ws(Mo, No, K) : 0;
  return a > b ? a : b;
}

int gts_gemm_tn2_colsum(const float* A, int64_t lda, const float* B1, int64_t ldb1, const float* B2, int64_t ldb2,
                        float* C1, float* C2, int64_t ldc, int32_t Mo, int32_t No, int64_t K, int32_t mode,
                        float* colsum_out, void* workspace, size_t workspace_bytes, gts_stream_t stream) {
  GTS_CHECK_ARG(Mo >= 0 && No >= 0 && K >= 0, "gts_gemm_tn2_colsum: negative size");
  GTS_CHECK_ARG(mode >= GTS_GEMM_FP32 && mode <= GTS_GEMM_TF32X3, "gts_gemm_tn2_colsum: unknown mode %d", mode);
  if (Mo > 0 && No > 0 && K > 0 && A && B1 && B2 && C1 && C2 && colsum_out &&
      gemm_tn2_tcgen05_supported(A, lda, B1, ldb1, B2, ldb2, Mo, No, K, mode))
    return gemm_tn2_tcgen05(A, lda, B1, ldb1, B2, ldb2, C1, C2, ldc, Mo, No, K, colsum_out, workspace, workspace_bytes,
                            as_stream(stream));
  // two products through ONE workspace: the first one's split-K reduction must run before the second product
  // overwrites the partial sums, so neither is deferred (narrow first / last layer shapes only)
  ReduceBatch* const deferred = splitk_defer_target();
  splitk_defer_set(nullptr);
  int rc = gts_gemm_tn_colsum(A, lda, B1, ldb1, C1, ldc, Mo, No, K, mode, colsum_out, workspace, workspace_bytes, stream);
  if (rc == GTS_OK) rc = gts_gemm_tn(A, lda, B2, ldb2, C2, ldc, Mo, No, K, mode, workspace, workspace_bytes, stream);
  splitk_defer_set(deferred);
  return rc;
}

size_t gts_colsum_workspace_bytes(int64_t rows, int32_t cols) {
  if (rows <= 0 || cols <= 0) return 256;
  return align_up((size_t)colsum_blocks(rows) * (size_t)cols * sizeof(float), 256);
}

int gts_colsum(const float* A, int64_t lda, int64_t rows, int32_t cols, float* out,
               void* workspace, size_t workspace_bytes, gts_stream_t stream) {
  GTS_CHECK_ARG(rows >= 0 && cols >= 0, "gts_colsum: negative size");
  if (cols == 0) return GTS_OK;
  GTS_CHECK_ARG(out != nullptr, "gts_colsum: out is null");
  cudaStream_t st = as_stream(stream);
  if (rows == 0) {
    GTS_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * cols, st));
    return GTS_OK;
  }
  GTS_CHECK_ARG(A != nullptr && workspace != nullptr, "gts_colsum: null pointer");
  const size_t need = gts_colsum_workspace_bytes(rows, cols);
  if (workspace_bytes < need) {
    set_error("gts_colsum: workspace %zu < required %zu", workspace_bytes, need);
    return GTS_ERR_WORKSPACE;
  }
  const int nb = colsum_blocks(rows);
  const int64_t rpb = ceil_div<int64_t>(rows, nb);
  float* partial = reinterpret_cast<float*>(workspace);
  const bool vec = cols % 4 == 0 && cols >= 16 && cols <= 1024 && lda % 4 == 0 && (reinterpret_cast<uintptr_t>(A) & 15u) == 0 &&
                   256 % (cols / 4) == 0;
  if (vec) {
    const int cols4 = cols / 4;
    const size_t smem = (size_t)(256 / cols4) * cols4 * sizeof(float4);
    colsum_partial_vec_kernel<<<nb, 256, smem, st>>>(A, lda, rows, cols4, rpb, partial);
    GTS_LAUNCH_CHECK();
    colsum_final_wide_kernel<<<(cols + 31) / 32, 256, 0, st>>>(partial, nb, cols, out);
  } else {
    colsum_partial_kernel<<<nb, 256, 0, st>>>(A, lda, rows, cols, rpb, partial);
    GTS_LAUNCH_CHECK();
    colsum_final_wide_kernel<<<(cols + 31) / 32, 256, 0, st>>>(partial, nb, cols, out);
  }
  GTS_LAUNCH_CHECK();
  return GTS_OK;
}

int gts_transpose(const float* in, int64_t ldin, int32_t rows, int32_t cols,
                  float* out, int64_t ldout, gts_stream_t stream) {
  GTS_CHECK_ARG(rows >= 0 && cols >= 0, "gts_transpose: negative size");
  if (rows == 0 || cols == 0) return GTS_OK;
  GTS_CHECK_ARG(in && out, "gts_transpose: null pointer");
  dim3 grid((cols + 31) / 32, (rows + 31) / 32), block(32, 8);
  transpose_kernel<<<grid, block, 0, as_stream(stream)>>>(in, ldin, rows, cols, out, ldout);
  GTS_LAUNCH_CHECK();
  return GTS_OK;
}

}  // extern "C"
