// gemm_tcgen05.cu — K1/K3/K4 dense contractions on the 5th-gen tensor cores.
//
//   tcgen05.mma.cta_group::1.kind::tf32 (M=128, N<=256, K=8 per instruction),
//   operands staged in shared memory by TMA (cp.async.bulk.tensor, 128B
//   swizzle), fp32 accumulators in TMEM (two 256-column buffers so the MMA of
//   tile i+1 overlaps the epilogue of tile i), epilogue read back with
//   tcgen05.ld and fused bias / ReLU / ReLU-mask.  Persistent CTAs, one per SM.
//
// Replaces nn.Linear inside DGL SAGEConv / GATConv and its autograd (invoked at
// reference model/networks.py:35,63,65):
//   NT form  C[M,N]  = act(A1 B1^T + A2 B2^T + bias)  — forward and data-gradient
//                      GEMMs; both operands K-major; the two-source K loop is the
//                      concatenated fc_self || fc_neigh contraction;
//   TN form  C[Mo,No] = A^T B over K = nodes          — weight gradients; both
//                      operands MN-major (TMA boxes of 32 nodes x 32 columns),
//                      split-K across the grid + deterministic reduction.
//
// Warp roles (192 threads): warp 0 = TMA producer (one elected lane), warp 1 =
// TMEM allocator + MMA issuer (one elected lane), warps 2..5 = epilogue (each
// owns the TMEM lane quarter warp_id % 4).
//
// Arithmetic modes: GTS_GEMM_TF32 — operands rounded to TF32 by the TMA load
// (CU_TENSOR_MAP_DATA_TYPE_TFLOAT32), one MMA pass; GTS_GEMM_TF32X3 — see the
// split kernel variant below (hi/lo split in shared memory, three MMA passes,
// fp32-accurate).
#include "common.cuh"
#include <cuda.h>
#include <cuda_bf16.h>
#include <cstdlib>

namespace gts {

void launch_splitk_reduce(const float* partial, int64_t split_stride, int splits, int64_t rows, int64_t cols,
                          float* C, int64_t ldc, cudaStream_t st);   // gemm_simt.cu
bool splitk_reduce_fused_ok(int64_t rows, int64_t cols, int64_t ldc, int64_t tail_len, int64_t split_stride,
                            const void* partial, const void* C, const void* tail);
void launch_splitk_reduce_fused(const float* partial, int64_t split_stride, int splits, int64_t rows, int64_t cols,
                                float* C, float* tail, int64_t tail_len, cudaStream_t st);

namespace tc {

constexpr int BM = 128;            // UMMA M
constexpr int BK = 32;             // fp32 elements per k-block = one 128-byte swizzle span
constexpr int UMMA_K = 8;          // tf32: 32 bytes per instruction along K
constexpr int MAX_BN = 256;
constexpr int A_STAGE_BYTES = BM * BK * 4;        // 16 KB
constexpr int B_STAGE_BYTES = MAX_BN * BK * 4;    // 32 KB
constexpr int OPER_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;   // one set of operands (A | B)
constexpr int EPI_LD = 36;                        // padded row (floats) of the per-warp staging tile
constexpr int NUM_THREADS = 320;                  // 10 warps, see roles below
constexpr int TMEM_COLS = 512;

// X3 = false: 3 stages of (A | B), 8 epilogue warps.
// X3 = true : 2 stages of (A | B | A_lo | B_lo), 4 epilogue warps + 4 split warps.
template <bool X3>
struct Cfg {
  static constexpr int STAGES = X3 ? 2 : 3;
  static constexpr int STAGE_BYTES = X3 ? 2 * OPER_BYTES : OPER_BYTES;
  static constexpr int EPI_WARPS = X3 ? 4 : 8;
  static constexpr int SPLIT_WARPS = X3 ? 4 : 0;
  static constexpr int stages_bytes = STAGES * STAGE_BYTES;
  static constexpr int epi_off = stages_bytes;
  static constexpr int epi_bytes = EPI_WARPS * 32 * EPI_LD * 4;
  static constexpr int bar_off = epi_off + epi_bytes;
  // full[STAGES], empty[STAGES], split[STAGES], tmem_full[2], tmem_empty[2], tmem_ptr
  static constexpr int total = bar_off + (3 * STAGES + 4) * 8 + 16;
  static constexpr int dyn_bytes = total + 1024;   // slack for the manual 1024-byte alignment
  static_assert(dyn_bytes <= 232448, "exceeds the 227 KB shared-memory limit per CTA");
  static_assert(2 + EPI_WARPS + SPLIT_WARPS == NUM_THREADS / 32, "warp roles must cover the CTA");
};

// ---------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done = 0;
  uint32_t spins = 0;
  while (true) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(addr), "r"(parity) : "memory");
    if (done) break;
    if (++spins > (1u << 22)) __trap();    // watchdog: a protocol bug must fault, not hang the GPU
  }
}
// non-blocking probe: the MMA issuer's barriers are almost always complete already, and every clock it spends
// not issuing is exposed once the (shallow) tensor-core issue queue has drained
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return done != 0;
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* tmap, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(tmap) : "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// Whole-warp forms (one elected lane issues).  The MMA warp runs its loop with all 32 lanes and warp-uniform
// operands: operands that the compiler can prove uniform stay in uniform registers, whereas operands of a
// single-lane branch cost an ELECT / R2UR.BROADCAST loop of ~11 instructions around every UTCHMMA — and the
// issuing thread is effectively synchronous with the tensor pipe, so that overhead is not hidden (measured:
// 98 -> 66 clk per 64-clk MMA).
__device__ __forceinline__ void tcgen05_commit_elect(uint32_t bar_addr) {
  asm volatile(
      "{\n\t.reg .pred pe;\n\t"
      "elect.sync _|pe, 0xffffffff;\n\t"
      "@pe tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}"
      ::"r"(bar_addr) : "memory");
}
__device__ __forceinline__ void tcgen05_mma_tf32_elect(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                                       uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, pe;\n\t"
      "elect.sync _|pe, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tcgen05_mma_tf32_ts_elect(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, pe;\n\t"
      "elect.sync _|pe, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Explicit shared-state-space accesses.  Through generic pointers (uint8_t* smem + offset) the compiler emitted generic
// LD.E / ST.E for every tile access of the split and epilogue warps (63-79 per kernel in round 1's SASS, against 5-9
// LDS / STS): generic accesses take the slower address path and queue behind outstanding global loads.
__device__ __forceinline__ void sts_128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void sts_128f(uint32_t addr, const float4& v) {
  asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ float4 lds_128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void red_add_f32(float* p, float v) {
  asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}
__device__ __forceinline__ float lds_32(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
  return v;
}

// Shared-memory matrix descriptor (SM100 UMMA), version 1.
//   K-major  (layout 2 = SWIZZLE_128B, 16-byte swizzle granule): LBO unused (1),
//            SBO = 1024 B between 8-row groups;
//   MN-major (layout 1 = SWIZZLE_128B_BASE32B — the only MN-major layout the
//            tensor core accepts for 32-bit (tf32) operands; pairs with the TMA
//            swizzle CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B): rows of 32 contiguous
//            MN elements, 4 k-rows per 512-byte swizzle atom; LBO = bytes between
//            32-element MN chunks, SBO = 512 B between groups of 4 k-rows.
constexpr uint32_t kLayoutSw128 = 2, kLayoutSw128Base32 = 1;
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                                   uint32_t layout_type) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;     // descriptor version (Blackwell)
  d |= (uint64_t)layout_type << 61;
  return d;
}
// Instruction descriptor: D = f32, A = B = tf32, dense, no negate.
__host__ __device__ constexpr uint32_t make_idesc(int m, int n, bool mn_major) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((mn_major ? 1u : 0u) << 15) | ((mn_major ? 1u : 0u) << 16) |
         ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

struct Params {
  // common
  int tiles_m, tiles_n;
  int BN;                 // UMMA N of this launch (multiple of 16, <= 256)
  int M, N;               // output extents (NT: rows/cols of C; TN: Mo/No)
  float* C; int64_t ldc;
  // NT
  int kb1, kb2;           // k-blocks of source 1 / 2
  const float* bias; const float* bias2; const float* aux; int64_t ldaux; int act;
  // TN (split-K)
  int splits; int kb_per_split; int kb_total; int64_t split_stride;
  int64_t c2_off;          // TN two-B form: offset (floats) of the second product inside a split record
  float* colsum_partial;   // TN, A-in-TMEM kernel: [splits][M] column sums of the A operand (or null)
  uint32_t* bits_out; int64_t ld_bits_out;            // wide NT kernel: bit matrix of (C > 0) beside a ReLU output (or null)
  const uint32_t* aux_bits; int64_t ld_aux_bits;      // wide NT kernel, GTS_ACT_MASK_BITS: the mask as a bit matrix
  const int32_t* scatter_idx; int64_t ld_idx;         // wide NT kernel, GTS_ACT_MASK_BITS_SCATTER: arg-max rows [M, N]
  float* scatter_out; int64_t ld_out;                 //   and the zero-filled destination the masked tile is added into
  uint4* zero_fill; unsigned long long zero_n16;      // wide NT kernel: side job of the two spare warps (16-byte units)
  long long* trace;        // GTS_TRACE builds only: per-CTA clock64 records of gemm_x3ntw_kernel (tools/gemm_trace.py)
  int dbg;                 // GTS_TRACE builds only: ablation switches
};

// In-kernel trace (compile with -DGTS_TRACE; the shipped library has none of it): record r of CTA b at
// trace[(b * kTraceItems + item) * kTraceSlots + r].
#ifdef GTS_TRACE
constexpr int kTraceItems = 8, kTraceSlots = 16;
#define GTS_TR(item, r, val) do { if (p.trace && (item) < kTraceItems) p.trace[((long long)blockIdx.x * kTraceItems + (item)) * kTraceSlots + (r)] = (val); } while (0)
#define GTS_TR_ON 1
#else
#define GTS_TR(item, r, val) do { (void)(item); } while (0)
#define GTS_TR_ON 0
#endif

// ---------------------------------------------------------------------------
// The kernel.  TN = false: NT form;  TN = true: weight-gradient form.
// X3 = true: 3xTF32.  The tensor core TRUNCATES fp32 operands to TF32 (measured
// on B200: a FLOAT32 tensor map reproduces the truncated-operand product to
// 1e-6), so the landed fp32 tile itself serves as the "hi" part; four split
// warps write lo = x - trunc(x) next to it and the MMA warp issues
// lo*hi + hi*lo + hi*hi into the same fp32 TMEM accumulator.
// ---------------------------------------------------------------------------
// lo = x - trunc_tf32(x) is exact in fp32; the MMA will truncate lo itself to TF32, so bias it by
// half a TF32 ulp first (round-to-nearest instead of toward zero).  Measured effect on the 8-layer
// logits: 3.4e-5 -> 3.1e-5 relative, i.e. the residual error of this mode is dominated by the
// tensor core's fp32 accumulation, not by the operand split.
__device__ __forceinline__ float tf32_lo(float x) {
  const float lo = x - __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
  return __uint_as_float(__float_as_uint(lo) + 0x1000u);
}

// Epilogue of one accumulator tile for one warp: the warp owns TMEM lanes [q*32, q*32+32) (rows m0..m0+31 of C)
// and the 32-column chunks c_first, c_first + c_step, ...  TMEM -> registers -> padded smem tile -> coalesced
// 128-bit global stores with the fused bias / ReLU / ReLU-mask.
#ifdef GTS_TRACE
#define GTS_DBG(bit) (p.dbg & (1 << (bit)))     // timing-only ablation switches of the instrumented build (garbage results)
#else
#define GTS_DBG(bit) 0
#endif

template <bool TN>
__device__ __forceinline__ void epilogue_tile(const Params& p, uint32_t t_base, int m0, int n0, float* Cout, float* stg,
                                              int lane, int c_first, int c_step, bool masked, uint64_t* full_bar,
                                              uint32_t full_phase) {
  const int cc = (lane & 7) * 4;
  const int rsub = lane >> 3;
  const uint32_t stg_u = smem_u32(stg);
  // ReLU-mask source: fetched one chunk ahead so its latency hides behind the TMEM load
  float4 aux_cur[8], aux_nxt[8];
  auto load_aux = [&](int c0, float4 (&dst)[8]) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int64_t row = (int64_t)m0 + 4 * j + rsub;
      const int col = n0 + c0 + cc;
      dst[j] = (row < p.M && col < p.N && c0 + cc < p.BN)
                   ? ldg_nc_na(reinterpret_cast<const float4*>(p.aux + row * p.ldaux + col))
                   : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
    if (masked && c_first < p.BN) load_aux(c_first, aux_cur);

  mbar_wait(full_bar, full_phase);
  tcgen05_fence_after();
  for (int c0 = c_first; c0 < p.BN; c0 += c_step) {
    uint32_t v[32];
    if (!GTS_DBG(2)) tmem_ld_32x32b_x32(t_base + c0, v);
    if (masked && c0 + c_step < p.BN) load_aux(c0 + c_step, aux_nxt);
    if (!GTS_DBG(2)) tmem_ld_wait();
    // lane = row: park the 32 columns of this row in the padded staging tile
#pragma unroll
    for (int j = 0; j < 8; ++j)
      sts_128(stg_u + (uint32_t)(lane * EPI_LD + 4 * j) * 4, v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
    __syncwarp();
    // coalesced write-out: 8 lanes cover one 128-byte row segment, 4 rows per instruction
    const int col = n0 + c0 + cc;
    const bool col_ok = col < p.N && c0 + cc < p.BN;
    float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
    if (!TN && p.bias && col_ok && !GTS_DBG(1)) b = *reinterpret_cast<const float4*>(p.bias + col);
    if (!TN && p.bias2 && col_ok && !GTS_DBG(1)) {
      const float4 b2 = *reinterpret_cast<const float4*>(p.bias2 + col);
      b.x += b2.x; b.y += b2.y; b.z += b2.z; b.w += b2.w;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int rr = 4 * j + rsub;
      const int64_t row = (int64_t)m0 + rr;
      if (row < p.M && col_ok) {
        float4 x = lds_128(stg_u + (uint32_t)(rr * EPI_LD + cc) * 4);
        if (!TN) {
          x.x += b.x; x.y += b.y; x.z += b.z; x.w += b.w;
          if (p.act == GTS_ACT_RELU) {
            x.x = fmaxf(x.x, 0.f); x.y = fmaxf(x.y, 0.f); x.z = fmaxf(x.z, 0.f); x.w = fmaxf(x.w, 0.f);
          } else if (masked) {
            const float4 a = aux_cur[j];
            x.x = a.x > 0.f ? x.x : 0.f; x.y = a.y > 0.f ? x.y : 0.f;
            x.z = a.z > 0.f ? x.z : 0.f; x.w = a.w > 0.f ? x.w : 0.f;
          }
        }
        if (!GTS_DBG(0)) *reinterpret_cast<float4*>(Cout + row * p.ldc + col) = x;
      }
    }
    __syncwarp();
    if (masked) {
#pragma unroll
      for (int j = 0; j < 8; ++j) aux_cur[j] = aux_nxt[j];
    }
  }
}

template <bool TN, bool X3>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA1, const __grid_constant__ CUtensorMap tmA2,
               const __grid_constant__ CUtensorMap tmB1, const __grid_constant__ CUtensorMap tmB2, const Params p) {
  using L = Cfg<X3>;
  constexpr int STAGES = L::STAGES;
  constexpr int STAGE_BYTES = L::STAGE_BYTES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::bar_off);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* split_bar = full_bar + 2 * STAGES;
  uint64_t* tmem_full = full_bar + 3 * STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  float* epi_stage = reinterpret_cast<float*>(smem + L::epi_off);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);    // provably warp-uniform
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
      mbar_init(&split_bar[s], L::SPLIT_WARPS * 32);
    }
    for (int a = 0; a < 2; ++a) { mbar_init(&tmem_full[a], 1); mbar_init(&tmem_empty[a], L::EPI_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    prefetch_tmap(&tmA1); prefetch_tmap(&tmB1);
    if (!TN && p.kb2 > 0) { prefetch_tmap(&tmA2); prefetch_tmap(&tmB2); }
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  const int n_work = TN ? p.tiles_m * p.tiles_n * p.splits : p.tiles_m * p.tiles_n;
  const uint32_t stage_tx_bytes = (uint32_t)(BM + p.BN) * BK * 4;

  if (warp == 0) {
    // ======================= TMA producer =======================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int w = blockIdx.x; w < n_work; w += gridDim.x) {
        int tile = w, split = 0;
        if (TN) { split = w % p.splits; tile = w / p.splits; }
        const int m0 = (tile / p.tiles_n) * BM;
        const int n0 = (tile % p.tiles_n) * p.BN;
        int kb_beg = 0, kb_end = p.kb1 + p.kb2;
        if (TN) { kb_beg = split * p.kb_per_split; kb_end = min(p.kb_total, kb_beg + p.kb_per_split); }
        for (int kb = kb_beg; kb < kb_end; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_arrive_expect_tx(&full_bar[stage], stage_tx_bytes);
          uint8_t* sa = smem + stage * STAGE_BYTES;
          uint8_t* sb = sa + A_STAGE_BYTES;
          if (!TN) {
            const bool second = kb >= p.kb1;
            const int k0 = (second ? kb - p.kb1 : kb) * BK;
            tma_load_2d(sa, second ? &tmA2 : &tmA1, &full_bar[stage], k0, m0);
            tma_load_2d(sb, second ? &tmB2 : &tmB1, &full_bar[stage], k0, n0);
          } else {
            const int k0 = kb * BK;    // node rows
#pragma unroll
            for (int c = 0; c < BM / 32; ++c) tma_load_2d(sa + c * 4096, &tmA1, &full_bar[stage], m0 + 32 * c, k0);
            for (int c = 0; c < p.BN / 32; ++c) tma_load_2d(sb + c * 4096, &tmB1, &full_bar[stage], n0 + 32 * c, k0);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ======================= MMA issuer (whole warp, one elected lane issues) =======================
    {
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
      const uint32_t idesc = make_idesc(BM, p.BN, TN);
      const uint32_t smem0 = smem_u32(smem);
      int stage = 0; uint32_t phase = 0;
      int it = 0;
      for (int w = blockIdx.x; w < n_work; w += gridDim.x, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        int kb_beg = 0, kb_end = p.kb1 + p.kb2;
        if (TN) { const int split = w % p.splits; kb_beg = split * p.kb_per_split; kb_end = min(p.kb_total, kb_beg + p.kb_per_split); }
        if (!mbar_test(&tmem_empty[acc], acc_phase ^ 1)) mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tcgen05_fence_after();
        const uint32_t d_tmem = tmem_u + (uint32_t)acc * MAX_BN;
        for (int kb = kb_beg; kb < kb_end; ++kb) {
          uint64_t* rb = X3 ? &split_bar[stage] : &full_bar[stage];
          if (!mbar_test(rb, phase)) mbar_wait(rb, phase);
          tcgen05_fence_after();
          const uint32_t sa = smem0 + stage * STAGE_BYTES;
          const uint32_t sb = sa + A_STAGE_BYTES;
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            // K-major: advance 32 bytes inside the 128-byte swizzle span;
            // MN-major: 8 node rows = two 512-byte swizzle atoms per 32-column chunk
            const uint32_t koff = TN ? k * 1024 : k * UMMA_K * 4;
            const uint32_t lbo = TN ? 4096 : 16, sbo = TN ? 512 : 1024;
            const uint32_t lay = TN ? kLayoutSw128Base32 : kLayoutSw128;
            const uint64_t da = make_smem_desc(sa + koff, lbo, sbo, lay);
            const uint64_t db = make_smem_desc(sb + koff, lbo, sbo, lay);
            const uint32_t first = (kb > kb_beg || k > 0) ? 1u : 0u;
            if (X3) {
              const uint64_t da_lo = make_smem_desc(sa + OPER_BYTES + koff, lbo, sbo, lay);
              const uint64_t db_lo = make_smem_desc(sb + OPER_BYTES + koff, lbo, sbo, lay);
              tcgen05_mma_tf32_elect(d_tmem, da_lo, db, idesc, first);   // lo * hi
              tcgen05_mma_tf32_elect(d_tmem, da, db_lo, idesc, 1u);      // hi * lo
              tcgen05_mma_tf32_elect(d_tmem, da, db, idesc, 1u);         // hi * hi
            } else {
              tcgen05_mma_tf32_elect(d_tmem, da, db, idesc, first);
            }
          }
          tcgen05_commit_elect(smem_u32(&empty_bar[stage]));       // smem slot reusable once these MMAs retire
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        tcgen05_commit_elect(smem_u32(&tmem_full[acc]));           // accumulator complete -> epilogue
      }
    }
  } else if (warp < 2 + L::EPI_WARPS) {
    // ======================= epilogue warps =======================
    const int ew = warp - 2;
    const int q = warp & 3;                        // TMEM lane quarter this warp may access
    const int n_col_groups = L::EPI_WARPS / 4;     // warps sharing a lane quarter split the column chunks
    const int col_group = ew >> 2;
    float* stg = epi_stage + ew * 32 * EPI_LD;
    const bool masked = !TN && p.act == GTS_ACT_MASK_POS;
    int it = 0;
    for (int w = blockIdx.x; w < n_work; w += gridDim.x, ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      int tile = w, split = 0;
      if (TN) { split = w % p.splits; tile = w / p.splits; }
      const int m0 = (tile / p.tiles_n) * BM + q * 32;
      const int n0 = (tile % p.tiles_n) * p.BN;
      float* Cout = p.C + (TN ? (int64_t)split * p.split_stride : 0);
      const uint32_t t_base = tmem_base + (uint32_t)acc * MAX_BN + ((uint32_t)(q * 32) << 16);

      epilogue_tile<TN>(p, t_base, m0, n0, Cout, stg, lane, col_group * 32, n_col_groups * 32, masked,
                        &tmem_full[acc], acc_phase);
      tcgen05_fence_before();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
    }
  } else if (X3) {
    // ======================= split warps (3xTF32) =======================
    const int t = threadIdx.x - (2 + L::EPI_WARPS) * 32;          // 0..127
    const int n_vec = (A_STAGE_BYTES + p.BN * BK * 4) / 16;       // float4 count of (A | used part of B)
    int stage = 0; uint32_t phase = 0;
    for (int w = blockIdx.x; w < n_work; w += gridDim.x) {
      int kb_beg = 0, kb_end = p.kb1 + p.kb2;
      if (TN) { const int split = w % p.splits; kb_beg = split * p.kb_per_split; kb_end = min(p.kb_total, kb_beg + p.kb_per_split); }
      for (int kb = kb_beg; kb < kb_end; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        float4* hi = reinterpret_cast<float4*>(smem + stage * STAGE_BYTES);
        float4* lo = reinterpret_cast<float4*>(smem + stage * STAGE_BYTES + OPER_BYTES);
#pragma unroll 4
        for (int i = t; i < n_vec; i += L::SPLIT_WARPS * 32) {
          const float4 x = lds_128(smem_u32(hi) + (uint32_t)i * 16);
          sts_128f(smem_u32(lo) + (uint32_t)i * 16, make_float4(tf32_lo(x.x), tf32_lo(x.y), tf32_lo(x.z), tf32_lo(x.w)));
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the MMA (async proxy)
        mbar_arrive(&split_bar[stage]);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}


// ---------------------------------------------------------------------------
// 3xTF32 with the A operand in TENSOR MEMORY (tcgen05.mma ... [d], [a_tmem], b_desc).
//
// Why: in the all-shared-memory 3xTF32 scheme above every k-block moves 288 KB
// through shared memory (TMA 48 + split 96 + three MMA passes reading A and B,
// 144) against 1536 tensor-pipe cycles -> the 128 B/clk shared-memory port caps
// the tensor pipe at ~68 % (ncu: ~50 % achieved).  Here four warps read the
// landed A tile ONCE, split it in registers and park hi | lo in TMEM
// (tcgen05.st, lane = A row, column = k), so the MMA only streams B from shared
// memory: 224 KB per k-block.
//
// Warp roles (576 threads): 0 = TMA producer, 1 = MMA issuer + TMEM allocator,
// 2..5 = A split -> TMEM (warp % 4 = TMEM lane quarter), 6..9 = B split in shared
// memory, 10..17 = epilogue (two warps per lane quarter).
// TMEM: columns [0,256) one fp32 accumulator, [256 + 64 s, +64) A hi | lo of stage s.
// TN form: A = [nodes, Mo] row-major (MN-major box), so the split thread of feature
// m walks its column of the landed box — and sums it on the way: the bias
// gradient (column sum of the A operand) is a by-product of the weight-gradient GEMM.
// ---------------------------------------------------------------------------
// BNC = widest N tile (UMMA N) of the variant:
//   256: 2 stages of (A 16 | B 32 | B lo 32 KB), ONE 256-column accumulator (epilogue exposed), 8 epilogue warps;
//   128: 4 stages of (A 16 | B 16 | B lo 16 KB), TWO 128-column accumulators (epilogue overlaps the next
//        tile), 4 epilogue warps.  A 256-wide output is two N tiles, i.e. A is fetched twice (the second time
//        from L2) — but twice the bytes are in flight per SM, and with only ~48 KB per stage the 2-stage form
//        is bound by the L2 -> SM round trip (measured: 3000+ clk per k-block against 1536 of tensor work).
template <int BNC>
struct TsCfg {
  static constexpr int STAGES = BNC == 256 ? 2 : 4;
  static constexpr int ACC_BUFS = BNC == 256 ? 1 : 2;
  static constexpr int EPI_WARPS = BNC == 256 ? 8 : 4;
  static constexpr int B_BYTES = BNC * BK * 4;
  static constexpr int STAGE_BYTES = A_STAGE_BYTES + 2 * B_BYTES;     // A | B hi | B lo
  static constexpr int THREADS = (10 + EPI_WARPS) * 32;
  static constexpr int epi_off = STAGES * STAGE_BYTES;
  static constexpr int epi_bytes = EPI_WARPS * 32 * EPI_LD * 4;
  static constexpr int bar_off = epi_off + epi_bytes;
  // full[S], empty[S], a_ready[S], b_ready[S], tmem_full[ACC_BUFS], tmem_empty[ACC_BUFS], tmem_ptr
  static constexpr int total = bar_off + (4 * STAGES + 2 * ACC_BUFS) * 8 + 16;
  static constexpr int dyn_bytes = total + 1024;
  static_assert(dyn_bytes <= 232448, "exceeds the 227 KB shared-memory limit per CTA");
  static constexpr uint32_t A_COL0 = 256;       // first TMEM column of the A ring (64 columns per stage: hi | lo)
  static_assert(A_COL0 + 64 * STAGES <= TMEM_COLS && ACC_BUFS * BNC <= (int)A_COL0, "TMEM column budget");
};

__device__ __forceinline__ void tmem_st_32x32b_x32(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,"
      "%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
        "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]),
        "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]),
        "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// D = f32, A = B = tf32; A always K-major (it lives in TMEM), B K-major (NT) or MN-major (TN).
__host__ __device__ constexpr uint32_t make_idesc_ts(int m, int n, bool b_mn_major) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((b_mn_major ? 1u : 0u) << 16) | ((uint32_t)(n >> 3) << 17) |
         ((uint32_t)(m >> 4) << 24);
}

template <bool TN, int BNC>
__global__ void __launch_bounds__(TsCfg<BNC>::THREADS, 1)
gemm_x3ts_kernel(const __grid_constant__ CUtensorMap tmA1, const __grid_constant__ CUtensorMap tmA2,
                 const __grid_constant__ CUtensorMap tmB1, const __grid_constant__ CUtensorMap tmB2, const Params p) {
  using L = TsCfg<BNC>;
  constexpr int STAGES = L::STAGES;
  constexpr int STAGE_BYTES = L::STAGE_BYTES;
  constexpr int ACC_BUFS = L::ACC_BUFS;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::bar_off);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* a_ready = full_bar + 2 * STAGES;
  uint64_t* b_ready = full_bar + 3 * STAGES;
  uint64_t* tmem_full = full_bar + 4 * STAGES;
  uint64_t* tmem_empty = tmem_full + ACC_BUFS;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + ACC_BUFS);
  float* epi_stage = reinterpret_cast<float*>(smem + L::epi_off);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);    // provably warp-uniform
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
      mbar_init(&a_ready[s], 128);
      mbar_init(&b_ready[s], 128);
    }
    for (int a = 0; a < ACC_BUFS; ++a) { mbar_init(&tmem_full[a], 1); mbar_init(&tmem_empty[a], L::EPI_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    prefetch_tmap(&tmA1); prefetch_tmap(&tmB1);
    if (!TN && p.kb2 > 0) { prefetch_tmap(&tmA2); prefetch_tmap(&tmB2); }
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  const int n_work = TN ? p.tiles_m * p.tiles_n * p.splits : p.tiles_m * p.tiles_n;
  const uint32_t stage_tx_bytes = (uint32_t)(BM + p.BN) * BK * 4;

  // k-block range of work item w
  auto k_range = [&](int w, int& kb_beg, int& kb_end) {
    kb_beg = 0; kb_end = p.kb1 + p.kb2;
    if (TN) { const int split = w % p.splits; kb_beg = split * p.kb_per_split; kb_end = min(p.kb_total, kb_beg + p.kb_per_split); }
  };

  if (warp == 0) {
    // ======================= TMA producer =======================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int w = blockIdx.x; w < n_work; w += gridDim.x) {
        const int tile = TN ? w / p.splits : w;
        const int m0 = (tile / p.tiles_n) * BM;
        const int n0 = (tile % p.tiles_n) * p.BN;
        int kb_beg, kb_end;
        k_range(w, kb_beg, kb_end);
        for (int kb = kb_beg; kb < kb_end; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_arrive_expect_tx(&full_bar[stage], stage_tx_bytes);
          uint8_t* sa = smem + stage * STAGE_BYTES;
          uint8_t* sb = sa + A_STAGE_BYTES;
          if (!TN) {
            const bool second = kb >= p.kb1;
            const int k0 = (second ? kb - p.kb1 : kb) * BK;
            tma_load_2d(sa, second ? &tmA2 : &tmA1, &full_bar[stage], k0, m0);
            tma_load_2d(sb, second ? &tmB2 : &tmB1, &full_bar[stage], k0, n0);
          } else {
            const int k0 = kb * BK;    // node rows
#pragma unroll
            for (int c = 0; c < BM / 32; ++c) tma_load_2d(sa + c * 4096, &tmA1, &full_bar[stage], m0 + 32 * c, k0);
            for (int c = 0; c < p.BN / 32; ++c) tma_load_2d(sb + c * 4096, &tmB1, &full_bar[stage], n0 + 32 * c, k0);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ======================= MMA issuer (whole warp, one elected lane issues) =======================
    {
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
      const uint32_t idesc = make_idesc_ts(BM, p.BN, TN);
      const uint32_t smem0 = smem_u32(smem);
      int stage = 0; uint32_t phase = 0;
      int it = 0;
      for (int w = blockIdx.x; w < n_work; w += gridDim.x, ++it) {
        int kb_beg, kb_end;
        k_range(w, kb_beg, kb_end);
        const int acc = it % ACC_BUFS;
        const uint32_t acc_phase = (uint32_t)(it / ACC_BUFS) & 1;
        if (!mbar_test(&tmem_empty[acc], acc_phase ^ 1)) mbar_wait(&tmem_empty[acc], acc_phase ^ 1);   // the epilogue has drained this accumulator
        tcgen05_fence_after();
        const uint32_t d_tmem = tmem_u + (uint32_t)acc * BNC;
        for (int kb = kb_beg; kb < kb_end; ++kb) {
          if (!mbar_test(&a_ready[stage], phase)) mbar_wait(&a_ready[stage], phase);
          if (!mbar_test(&b_ready[stage], phase)) mbar_wait(&b_ready[stage], phase);
          tcgen05_fence_after();
          const uint32_t sb = smem0 + stage * STAGE_BYTES + A_STAGE_BYTES;
          const uint32_t a_hi = tmem_u + L::A_COL0 + (uint32_t)stage * 64;
          const uint32_t a_lo = a_hi + 32;
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            const uint32_t koff = TN ? k * 1024 : k * UMMA_K * 4;
            const uint32_t lbo = TN ? 4096 : 16, sbo = TN ? 512 : 1024;
            const uint32_t lay = TN ? kLayoutSw128Base32 : kLayoutSw128;
            const uint64_t db = make_smem_desc(sb + koff, lbo, sbo, lay);
            const uint64_t db_lo = make_smem_desc(sb + L::B_BYTES + koff, lbo, sbo, lay);
            const uint32_t first = (kb > kb_beg || k > 0) ? 1u : 0u;
            tcgen05_mma_tf32_ts_elect(d_tmem, a_lo + k * UMMA_K, db, idesc, first);   // lo * hi
            tcgen05_mma_tf32_ts_elect(d_tmem, a_hi + k * UMMA_K, db_lo, idesc, 1u);   // hi * lo
            tcgen05_mma_tf32_ts_elect(d_tmem, a_hi + k * UMMA_K, db, idesc, 1u);      // hi * hi
          }
          tcgen05_commit_elect(smem_u32(&empty_bar[stage]));       // smem slot and TMEM A slot reusable once these MMAs retire
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        tcgen05_commit_elect(smem_u32(&tmem_full[acc]));           // accumulator complete -> epilogue
      }
    }
  } else if (warp < 6) {
    // ======================= A split -> TMEM =======================
    const int q = warp & 3;                        // TMEM lane quarter of this warp
    const int r = q * 32 + lane;                   // row of the 128-row A tile (NT) / feature column (TN)
    int stage = 0; uint32_t phase = 0;
    for (int w = blockIdx.x; w < n_work; w += gridDim.x) {
      int kb_beg, kb_end;
      k_range(w, kb_beg, kb_end);
      float csum = 0.f;
      for (int kb = kb_beg; kb < kb_end; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        const uint8_t* sa = smem + stage * STAGE_BYTES;
        uint32_t hi[32], lo[32];
        if (!TN) {
          // K-major tile, 128B swizzle: row r at r * 128, 16-byte chunk c stored at chunk c ^ (r & 7)
          const uint32_t row = smem_u32(sa) + (uint32_t)r * 128;
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const float4 x = lds_128(row + ((uint32_t)(c ^ (r & 7)) << 4));
            hi[4 * c + 0] = __float_as_uint(x.x); hi[4 * c + 1] = __float_as_uint(x.y);
            hi[4 * c + 2] = __float_as_uint(x.z); hi[4 * c + 3] = __float_as_uint(x.w);
          }
        } else {
          // MN-major boxes [32 node rows][32 features], SWIZZLE_128B_ATOM_32B: box q holds this warp's 32
          // features; node row k at k * 128, 32-byte unit u stored at unit u ^ (k & 3)
          const uint32_t box = smem_u32(sa) + (uint32_t)(q * 4096 + (lane & 7) * 4);
          const int u = lane >> 3;
#pragma unroll
          for (int k = 0; k < 32; ++k) {
            const float x = lds_32(box + (uint32_t)(k * 128 + ((u ^ (k & 3)) << 5)));
            hi[k] = __float_as_uint(x);
            csum += x;
          }
        }
#pragma unroll
        for (int i = 0; i < 32; ++i) lo[i] = __float_as_uint(tf32_lo(__uint_as_float(hi[i])));
        const uint32_t t_a = tmem_base + ((uint32_t)(q * 32) << 16) + L::A_COL0 + (uint32_t)stage * 64;
        tmem_st_32x32b_x32(t_a, hi);
        tmem_st_32x32b_x32(t_a + 32, lo);
        tmem_st_wait();
        tcgen05_fence_before();
        mbar_arrive(&a_ready[stage]);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      if (TN && p.colsum_partial) {
        const int split = w % p.splits, tile = w / p.splits;
        const int m = (tile / p.tiles_n) * BM + r;
        if (tile % p.tiles_n == 0 && m < p.M) p.colsum_partial[(int64_t)split * p.split_stride + m] = csum;
      }
    }
  } else if (warp < 10) {
    // ======================= B split in shared memory =======================
    const int t = threadIdx.x - 6 * 32;            // 0..127
    const int n_vec = p.BN * BK * 4 / 16;          // float4 count of the used part of B
    int stage = 0; uint32_t phase = 0;
    for (int w = blockIdx.x; w < n_work; w += gridDim.x) {
      int kb_beg, kb_end;
      k_range(w, kb_beg, kb_end);
      for (int kb = kb_beg; kb < kb_end; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        const float4* hi = reinterpret_cast<const float4*>(smem + stage * STAGE_BYTES + A_STAGE_BYTES);
        float4* lo = reinterpret_cast<float4*>(smem + stage * STAGE_BYTES + A_STAGE_BYTES + L::B_BYTES);
#pragma unroll 4
        for (int i = t; i < n_vec; i += 128) {
          const float4 x = lds_128(smem_u32(hi) + (uint32_t)i * 16);
          sts_128f(smem_u32(lo) + (uint32_t)i * 16, make_float4(tf32_lo(x.x), tf32_lo(x.y), tf32_lo(x.z), tf32_lo(x.w)));
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the MMA (async proxy)
        mbar_arrive(&b_ready[stage]);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else {
    // ======================= epilogue warps =======================
    const int ew = warp - 10;
    const int q = warp & 3;
    const int col_group = ew >> 2;
    float* stg = epi_stage + ew * 32 * EPI_LD;
    const bool masked = !TN && p.act == GTS_ACT_MASK_POS;
    int it = 0;
    for (int w = blockIdx.x; w < n_work; w += gridDim.x, ++it) {
      int tile = w, split = 0;
      if (TN) { split = w % p.splits; tile = w / p.splits; }
      const int m0 = (tile / p.tiles_n) * BM + q * 32;
      const int n0 = (tile % p.tiles_n) * p.BN;
      float* Cout = p.C + (TN ? (int64_t)split * p.split_stride : 0);
      const int acc = it % ACC_BUFS;
      const uint32_t t_base = tmem_base + (uint32_t)acc * BNC + ((uint32_t)(q * 32) << 16);
      epilogue_tile<TN>(p, t_base, m0, n0, Cout, stg, lane, col_group * 32, (L::EPI_WARPS / 4) * 32, masked,
                        &tmem_full[acc], (uint32_t)(it / ACC_BUFS) & 1);
      tcgen05_fence_before();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

// ---------------------------------------------------------------------------
// CTA-pair form of the A-in-TMEM 3xTF32 kernel (tcgen05.mma.cta_group::2, M = 256, N <= 128 per tile).
//
// What the measurements of the one-CTA forms said (profiles/r01_gemm_x3_pipeline.md):
//  * the tensor pipe does 128 clk per 128x256x8 tf32 MMA (1.07 PFLOP/s chip-wide) — the kernels ran at half of it;
//  * per-hop clock64 traces: one ring slot takes ~5500 clk from "seen empty" to "seen empty again" (TMA ~1100,
//    split + hand-over ~2000, MMA 1536, commit ~300) — more than the 2-3 slots that fit beside the lo copies can
//    hide, and most of the hand-over was a cluster-scope release fence on the arrive, not work;
//  * an epilogue that is NOT overlapped costs 26 us per 92 MB output: all SMs store at once and the burst runs at
//    the ~3.5 TB/s HBM write rate with nothing else in flight.
// Hence: a CTA pair shares B (each CTA lands, splits and feeds half of the N rows: TMA bytes, split work and
// operand reads per flop halve), N tiles of 128 so that TWO accumulators (2 x 128 TMEM columns) fit beside a
// 4-slot A ring (4 x 64 columns) and the epilogue of tile i overlaps the main loop of tile i+1, and 32 KB
// stages so that SIX of them are in flight.  A 256-wide output is two N tiles (A comes from L2 the second time).
//
// Both CTAs run the same warp roles on their own 128 rows of A / C; the MMA is issued by the leader (cluster
// rank 0) only, its a_ready / b_ready / tmem_empty barriers collect (plain, CTA-scope-release) remote arrivals
// from both CTAs, and tcgen05.commit multicasts "stage free" / "A slot free" / "accumulator full" to both.
// ---------------------------------------------------------------------------
// DUAL (weight gradients only): one A operand against TWO B operands per ring slot (dWs = dZ^T h and dWn = dZ^T neigh
// share dZ): the split A tile in TMEM feeds 24 instead of 12 MMAs, into two accumulators.
template <bool DUAL, int BF = 0>
struct Ts2CfgT {
  static constexpr int NB = DUAL ? 2 : 1;                                // B operands per slot
  static constexpr int BN_MAX = 128;                                   // UMMA N of a pair tile
  static constexpr int KSUB = 1;                                       // k-blocks (of 32) per ring slot: one barrier round per 24 MMAs
  static constexpr int STAGES = DUAL ? 4 : 6;                          // shared-memory ring
  static constexpr int SLOT_COLS = BF == 2 ? 32 : 64;                  // BF == 2: only the bf16 parts of A live in TMEM
  static constexpr int A_SLOTS = BF == 2 ? 8 : 4;                      // TMEM ring of split A tiles
  static constexpr int ACC_BUFS = 2;
  static constexpr int EPI_WARPS = 4;
  static constexpr int B_HALF_BYTES = (BN_MAX / 2) * BK * 4;           // 8 KB: this CTA's half of the N rows
  static constexpr int SUB_BYTES = A_STAGE_BYTES + NB * 2 * B_HALF_BYTES; // one k-block: A | (B half hi | B half lo) x NB = 32 / 48 KB
  static constexpr int STAGE_BYTES = KSUB * SUB_BYTES;
  static constexpr int THREADS = (10 + EPI_WARPS) * 32;
  static constexpr int epi_off = STAGES * STAGE_BYTES;
  static constexpr int epi_bytes = EPI_WARPS * 32 * EPI_LD * 4;
  static constexpr int bar_off = epi_off + epi_bytes;
  // full[S], empty[S], ready[A], a_free[A], tmem_full[2], tmem_empty[2], tmem_ptr
  static constexpr int total = bar_off + (2 * STAGES + 2 * A_SLOTS + 2 * ACC_BUFS) * 8 + 16;
  static constexpr int dyn_bytes = total + 1024;
  static_assert(dyn_bytes <= 232448, "exceeds the 227 KB shared-memory limit per CTA");
  static constexpr uint32_t A_COL0 = ACC_BUFS * BN_MAX;                // 256
  static_assert(A_COL0 + SLOT_COLS * KSUB * A_SLOTS <= TMEM_COLS, "TMEM column budget");
};
using Ts2Cfg = Ts2CfgT<false>;

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `local` (a shared::cta address) inside CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(uint32_t local, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local), "r"(rank));
  return r;
}
// remote arrive with the default (release, CTA scope) semantics: a cluster-scope release here costs ~1400 clk
// per hand-over (measured); what crosses the CTAs is ordered by the tcgen05 / proxy fences, not by this arrive
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// whole-warp call, one elected lane commits
__device__ __forceinline__ void tcgen05_commit_mc2_elect(uint32_t bar_addr) {
  asm volatile(
      "{\n\t.reg .pred pe;\n\t"
      "elect.sync _|pe, 0xffffffff;\n\t"
      "@pe tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n\t}"
      ::"r"(bar_addr), "h"((uint16_t)3) : "memory");
}


// The twelve MMAs of one 32-wide k-block of the 3xTF32 scheme (4 k-steps x {lo*hi, hi*lo, hi*hi}) as ONE asm
// block: the issuing thread is effectively synchronous with the tensor pipe (measured: every scalar
// instruction between two tcgen05.mma adds to the issue time), so all descriptors and TMEM addresses are
// formed before the first MMA and the twelve instructions go out back to back.  Called by the WHOLE MMA warp with
// warp-uniform operands (one elected lane issues): operands that the compiler can prove uniform live in uniform
// registers; per-thread operands cost an ELECT / R2UR.BROADCAST loop of ~11 instructions around every UTCHMMA.
// b_lo32 / l_lo32: low words of the B hi / B lo descriptors of k-step 0, k16: descriptor step per k-step (16-byte units).
__device__ __forceinline__ void mma_x3_block_ts2(uint32_t d_tmem, uint32_t a_hi, uint32_t b_lo32, uint32_t l_lo32,
                                                 uint32_t desc_hi32, uint32_t k16, uint32_t idesc, uint32_t first) {
  asm volatile(
      "{\n\t"
      ".reg .pred pf, pt, pe;\n\t"
      ".reg .b32 x1, x2, x3, y1, y2, y3, ah1, ah2, ah3, al0, al1, al2, al3;\n\t"
      ".reg .b64 b0, b1, b2, b3, l0, l1, l2, l3;\n\t"
      "elect.sync _|pe, 0xffffffff;\n\t"
      "setp.ne.b32 pf, %7, 0;\n\t"
      "setp.eq.b32 pt, %7, %7;\n\t"
      "add.u32 x1, %2, %5;\n\t add.u32 x2, x1, %5;\n\t add.u32 x3, x2, %5;\n\t"
      "add.u32 y1, %3, %5;\n\t add.u32 y2, y1, %5;\n\t add.u32 y3, y2, %5;\n\t"
      "mov.b64 b0, {%2, %4};\n\t mov.b64 b1, {x1, %4};\n\t mov.b64 b2, {x2, %4};\n\t mov.b64 b3, {x3, %4};\n\t"
      "mov.b64 l0, {%3, %4};\n\t mov.b64 l1, {y1, %4};\n\t mov.b64 l2, {y2, %4};\n\t mov.b64 l3, {y3, %4};\n\t"
      "add.u32 ah1, %1, 8;\n\t add.u32 ah2, %1, 16;\n\t add.u32 ah3, %1, 24;\n\t"
      "add.u32 al0, %1, 32;\n\t add.u32 al1, %1, 40;\n\t add.u32 al2, %1, 48;\n\t add.u32 al3, %1, 56;\n\t"
      "@pe tcgen05.mma.cta_group::2.kind::tf32 [%0], [al0], b0, %6, pf;\n\t"
      "@pe tcgen05.mma.cta_group::2.kind::tf32 [%0], [%1], l0, %6, pt;\n\t"
      "@pe tcgen05.mma.cta_group::2.kind::tf32 [%0], [%1], b0, %6, pt;\n\t"
      "@pe tcgen05.mma.cta_group::2.kind::tf32 [%0], [al1], b1, %6, pt;\n\t"
      "@pe tcgen05.mma.cta_group::2.kind::tf32 [%0], [ah1], l1, %6, pt;\n\t"
      "@pe tcgen05.mma.cta_group::2.kind::tf32 [%0], [ah1], b1, %6, pt;\n\t"
      "@pe tcgen05.mma.cta_group::2.kind::tf32 [%0], [al2], b2, %6, pt;\n\t"
      "@pe tcgen05.mma.cta_group::2.kind::tf32 [%0], [ah2], l2, %6, pt;\n\t"
      "@pe tcgen05.mma.cta_group::2.kind::tf32 [%0], [ah2], b2, %6, pt;\n\t"
      "@pe tcgen05.mma.cta_group::2.kind::tf32 [%0], [al3], b3, %6, pt;\n\t"
      "@pe tcgen05.mma.cta_group::2.kind::tf32 [%0], [ah3], l3, %6, pt;\n\t"
      "@pe tcgen05.mma.cta_group::2.kind::tf32 [%0], [ah3], b3, %6, pt;\n\t"
      "}"
      ::"r"(d_tmem), "r"(a_hi), "r"(b_lo32), "r"(l_lo32), "r"(desc_hi32), "r"(k16), "r"(idesc), "r"(first) : "memory");
}

__device__ __forceinline__ void tmem_st_32x32b_x16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
        "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
// D = f32, A = B = bf16 (kind::f16), A in TMEM (K-major), B K-major
__host__ __device__ constexpr uint32_t make_idesc_ts_bf16(int m, int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo_elem, float hi_elem) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(lo_elem, hi_elem);     // .x (low half) = even k, .y = odd k
  return *reinterpret_cast<const uint32_t*>(&h);
}

// "BF" form of the k-block (NT only): hi*hi stays a TF32 product of the raw fp32 operands (4 MMAs, K = 8 each), the
// two cross terms A_lo*B and A*B_lo run as bf16 MMAs (kind::f16, K = 16 each: 2 + 2 MMAs) at twice the tensor rate:
// 8 MMAs per k-block instead of 12.  The factors of the cross terms only need ~8 bits (they multiply a 2^-11 term:
// error ~2^-20 relative, below the fp32 accumulation error of the tensor core that bounds this mode).
// TMEM slot (64 columns): [0,32) A fp32 | [32,48) bf16x2(A_lo) | [48,64) bf16x2(A);  shared memory: B fp32 tile, then
// one 128-byte-row tile [bf16(B) k 0..31 | bf16(B_lo) k 0..31] with the same 128B swizzle.
__device__ __forceinline__ void mma_x3bf_block_ts2(uint32_t d_tmem, uint32_t a_hi, uint32_t b_lo32, uint32_t l_lo32,
                                                   uint32_t desc_hi32, uint32_t idesc, uint32_t idesc_bf, uint32_t first) {
  asm volatile(
      "{\n\t"
      ".reg .pred pf, pt, pe;\n\t"
      ".reg .b32 x1, x2, x3, y1, y2, y3, ah1, ah2, ah3, al0, al1, ab0, ab1;\n\t"
      ".reg .b64 b0, b1, b2, b3, h0, h1, l0, l1;\n\t"
      "elect.sync _|pe, 0xffffffff;\n\t"
      "setp.ne.b32 pf, %7, 0;\n\t"
      "setp.eq.b32 pt, %7, %7;\n\t"
      "add.u32 x1, %2, 2;\n\t add.u32 x2, %2, 4;\n\t add.u32 x3, %2, 6;\n\t"
      "add.u32 y1, %3, 2;\n\t add.u32 y2, %3, 4;\n\t add.u32 y3, %3, 6;\n\t"
      "mov.b64 b0, {%2, %4};\n\t mov.b64 b1, {x1, %4};\n\t mov.b64 b2, {x2, %4};\n\t mov.b64 b3, {x3, %4};\n\t"
      "mov.b64 h0, {%3, %4};\n\t mov.b64 h1, {y1, %4};\n\t mov.b64 l0, {y2, %4};\n\t mov.b64 l1, {y3, %4};\n\t"
      "add.u32 ah1, %1, 8;\n\t add.u32 ah2, %1, 16;\n\t add.u32 ah3, %1, 24;\n\t"
      "add.u32 al0, %1, 32;\n\t add.u32 al1, %1, 40;\n\t add.u32 ab0, %1, 48;\n\t add.u32 ab1, %1, 56;\n\t"
      "@pe tcgen05.mma.cta_group::2.kind::tf32 [%0], [%1], b0, %5, pf;\n\t"
      "@pe tcgen05.mma.cta_group::2.kind::tf32 [%0], [ah1], b1, %5, pt;\n\t"
      "@pe tcgen05.mma.cta_group::2.kind::tf32 [%0], [ah2], b2, %5, pt;\n\t"
      "@pe tcgen05.mma.cta_group::2.kind::tf32 [%0], [ah3], b3, %5, pt;\n\t"
      "@pe tcgen05.mma.cta_group::2.kind::f16 [%0], [al0], h0, %6, pt;\n\t"
      "@pe tcgen05.mma.cta_group::2.kind::f16 [%0], [al1], h1, %6, pt;\n\t"
      "@pe tcgen05.mma.cta_group::2.kind::f16 [%0], [ab0], l0, %6, pt;\n\t"
      "@pe tcgen05.mma.cta_group::2.kind::f16 [%0], [ab1], l1, %6, pt;\n\t"
      "}"
      ::"r"(d_tmem), "r"(a_hi), "r"(b_lo32), "r"(l_lo32), "r"(desc_hi32), "r"(idesc), "r"(idesc_bf), "r"(first) : "memory");
}

// TN (weight-gradient) form of the bf16-cross-term k-block.  hi*hi: four TF32 MMAs on the MN-major fp32 B tile as in
// mma_x3_block_ts2 (k-step = k16 descriptor units); the cross terms: 2 + 2 bf16 MMAs (K = 16) whose B operand is an
// MN-major bf16 tile in the standard 128-byte swizzle (one 128-byte row of this CTA's 64 n-values per k, 8 k-rows per
// 1024-byte swizzle atom, SBO = 1024): h_lo32 / l_lo32 = low descriptor words of the bf16(B) / bf16(B_lo) tiles at
// k-step 0, the second K = 16 step is two atoms (2048 B = 128 units) further.  A parts in TMEM as in the NT form.
__device__ __forceinline__ void mma_x3bf_block_ts2_tn(uint32_t d_tmem, uint32_t a_hi, uint32_t b_lo32, uint32_t k16,
                                                      uint32_t desc_hi32, uint32_t h_lo32, uint32_t l_lo32,
                                                      uint32_t desc_bf_hi32, uint32_t idesc, uint32_t idesc_bf, uint32_t first) {
  asm volatile(
      "{\n\t"
      ".reg .pred pf, pt, pe;\n\t"
      ".reg .b32 x1, x2, x3, y1, z1, ah1, ah2, ah3, al0, al1, ab0, ab1;\n\t"
      ".reg .b64 b0, b1, b2, b3, h0, h1, l0, l1;\n\t"
      "elect.sync _|pe, 0xffffffff;\n\t"
      "setp.ne.b32 pf, %10, 0;\n\t"
      "setp.eq.b32 pt, %10, %10;\n\t"
      "add.u32 x1, %2, %3;\n\t add.u32 x2, x1, %3;\n\t add.u32 x3, x2, %3;\n\t"
      "add.u32 y1, %5, 128;\n\t add.u32 z1, %6, 128;\n\t"
      "mov.b64 b0, {%2, %4};\n\t mov.b64 b1, {x1, %4};\n\t mov.b64 b2, {x2, %4};\n\t mov.b64 b3, {x3, %4};\n\t"
      "mov.b64 h0, {%5, %7};\n\t mov.b64 h1, {y1, %7};\n\t mov.b64 l0, {%6, %7};\n\t mov.b64 l1, {z1, %7};\n\t"
      "add.u32 ah1, %1, 8;\n\t add.u32 ah2, %1, 16;\n\t add.u32 ah3, %1, 24;\n\t"
      "add.u32 al0, %1, 32;\n\t add.u32 al1, %1, 40;\n\t add.u32 ab0, %1, 48;\n\t add.u32 ab1, %1, 56;\n\t"
      "@pe tcgen05.mma.cta_group::2.kind::tf32 [%0], [%1], b0, %8, pf;\n\t"
      "@pe tcgen05.mma.cta_group::2.kind::tf32 [%0], [ah1], b1, %8, pt;\n\t"
      "@pe tcgen05.mma.cta_group::2.kind::tf32 [%0], [ah2], b2, %8, pt;\n\t"
      "@pe tcgen05.mma.cta_group::2.kind::tf32 [%0], [ah3], b3, %8, pt;\n\t"
      "@pe tcgen05.mma.cta_group::2.kind::f16 [%0], [al0], h0, %9, pt;\n\t"
      "@pe tcgen05.mma.cta_group::2.kind::f16 [%0], [al1], h1, %9, pt;\n\t"
      "@pe tcgen05.mma.cta_group::2.kind::f16 [%0], [ab0], l0, %9, pt;\n\t"
      "@pe tcgen05.mma.cta_group::2.kind::f16 [%0], [ab1], l1, %9, pt;\n\t"
      "}"
      ::"r"(d_tmem), "r"(a_hi), "r"(b_lo32), "r"(k16), "r"(desc_hi32), "r"(h_lo32), "r"(l_lo32), "r"(desc_bf_hi32),
        "r"(idesc), "r"(idesc_bf), "r"(first) : "memory");
}

// BF == 2: as above, but the hi*hi TF32 product reads A straight from the TMA-landed shared-memory tile (SS form:
// the raw fp32 tile IS the hi operand, the MMA truncates it), so a TMEM slot holds only the two bf16 parts
// ([0,16) bf16x2(A_lo) | [16,32) bf16x2(A)) and the ring is 8 slots deep instead of 4.
__device__ __forceinline__ void mma_x3bf2_block_ts2(uint32_t d_tmem, uint32_t a_tm, uint32_t a_lo32, uint32_t b_lo32,
                                                    uint32_t l_lo32, uint32_t desc_hi32, uint32_t idesc, uint32_t idesc_bf,
                                                    uint32_t first) {
  asm volatile(
      "{\n\t"
      ".reg .pred pf, pt, pe;\n\t"
      ".reg .b32 x1, x2, x3, y1, y2, y3, z1, z2, z3, al1, ab0, ab1;\n\t"
      ".reg .b64 a0, a1, a2, a3, b0, b1, b2, b3, h0, h1, l0, l1;\n\t"
      "elect.sync _|pe, 0xffffffff;\n\t"
      "setp.ne.b32 pf, %8, 0;\n\t"
      "setp.eq.b32 pt, %8, %8;\n\t"
      "add.u32 x1, %3, 2;\n\t add.u32 x2, %3, 4;\n\t add.u32 x3, %3, 6;\n\t"
      "add.u32 y1, %4, 2;\n\t add.u32 y2, %4, 4;\n\t add.u32 y3, %4, 6;\n\t"
      "add.u32 z1, %2, 2;\n\t add.u32 z2, %2, 4;\n\t add.u32 z3, %2, 6;\n\t"
      "mov.b64 a0, {%2, %5};\n\t mov.b64 a1, {z1, %5};\n\t mov.b64 a2, {z2, %5};\n\t mov.b64 a3, {z3, %5};\n\t"
      "mov.b64 b0, {%3, %5};\n\t mov.b64 b1, {x1, %5};\n\t mov.b64 b2, {x2, %5};\n\t mov.b64 b3, {x3, %5};\n\t"
      "mov.b64 h0, {%4, %5};\n\t mov.b64 h1, {y1, %5};\n\t mov.b64 l0, {y2, %5};\n\t mov.b64 l1, {y3, %5};\n\t"
      "add.u32 al1, %1, 8;\n\t add.u32 ab0, %1, 16;\n\t add.u32 ab1, %1, 24;\n\t"
      "@pe tcgen05.mma.cta_group::2.kind::tf32 [%0], a0, b0, %6, pf;\n\t"
      "@pe tcgen05.mma.cta_group::2.kind::tf32 [%0], a1, b1, %6, pt;\n\t"
      "@pe tcgen05.mma.cta_group::2.kind::tf32 [%0], a2, b2, %6, pt;\n\t"
      "@pe tcgen05.mma.cta_group::2.kind::tf32 [%0], a3, b3, %6, pt;\n\t"
      "@pe tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], h0, %7, pt;\n\t"
      "@pe tcgen05.mma.cta_group::2.kind::f16 [%0], [al1], h1, %7, pt;\n\t"
      "@pe tcgen05.mma.cta_group::2.kind::f16 [%0], [ab0], l0, %7, pt;\n\t"
      "@pe tcgen05.mma.cta_group::2.kind::f16 [%0], [ab1], l1, %7, pt;\n\t"
      "}"
      ::"r"(d_tmem), "r"(a_tm), "r"(a_lo32), "r"(b_lo32), "r"(l_lo32), "r"(desc_hi32), "r"(idesc), "r"(idesc_bf), "r"(first)
      : "memory");
}

template <bool TN, bool DUAL, int BF = 0>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(Ts2Cfg::THREADS, 1)
gemm_x3ts2_kernel(const __grid_constant__ CUtensorMap tmA1, const __grid_constant__ CUtensorMap tmA2,
                  const __grid_constant__ CUtensorMap tmB1, const __grid_constant__ CUtensorMap tmB2, const Params p) {
  static_assert(!DUAL || TN, "the two-B form exists for the weight-gradient GEMM only");
  static_assert(!(TN && BF == 2), "TN form: bf16 cross terms with the A operand in tensor memory (BF == 1) only");
  using L = Ts2CfgT<DUAL, BF>;
  constexpr int NB = L::NB;
  constexpr int STAGES = L::STAGES;
  constexpr int A_SLOTS = L::A_SLOTS;
  constexpr int ACC_BUFS = L::ACC_BUFS;
  constexpr int STAGE_BYTES = L::STAGE_BYTES;
  constexpr int KSUB = L::KSUB;
  constexpr int SUB_BYTES = L::SUB_BYTES;
  extern __shared__ uint8_t smem_raw[];
  // the same offset in both CTAs (the dynamic window starts at the same shared::cta address in every CTA of a kernel)
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::bar_off);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* ready = full_bar + 2 * STAGES;           // A split in TMEM and B lo in shared memory, both CTAs (leader's copy is used)
  uint64_t* a_free = ready + A_SLOTS;
  uint64_t* tmem_full = a_free + A_SLOTS;
  uint64_t* tmem_empty = tmem_full + ACC_BUFS;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + ACC_BUFS);
  float* epi_stage = reinterpret_cast<float*>(smem + L::epi_off);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);    // provably warp-uniform
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();           // 0 = leader (issues the MMAs), 1 = peer
  const int cluster_id = blockIdx.x >> 1;
  const int n_clusters = gridDim.x >> 1;
  const int half_bn = p.BN >> 1;                     // N rows of B this CTA lands and feeds

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);                   // multicast commit of the leader
    }
    for (int s = 0; s < A_SLOTS; ++s) {
      mbar_init(&ready[s], 16);                      // (4 A-split + 4 B-split warps) x 2 CTAs
      mbar_init(&a_free[s], 1);                      // multicast commit of the leader
    }
    for (int a = 0; a < ACC_BUFS; ++a) {
      mbar_init(&tmem_full[a], 1);
      mbar_init(&tmem_empty[a], 2 * L::EPI_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    prefetch_tmap(&tmA1); prefetch_tmap(&tmB1);
    if (!TN && p.kb2 > 0) { prefetch_tmap(&tmA2); prefetch_tmap(&tmB2); }
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  cluster_sync_all();                                // barriers initialised and TMEM allocated in BOTH CTAs
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  // work item w -> (pair tile of 256 rows, N tile, split);  tiles_m counts 256-row pair tiles
  const int n_work = TN ? p.tiles_m * p.tiles_n * p.splits : p.tiles_m * p.tiles_n;
  const uint32_t stage_tx_bytes = (uint32_t)(BM + NB * half_bn) * BK * 4;

  auto k_range = [&](int w, int& kb_beg, int& kb_end) {
    kb_beg = 0; kb_end = p.kb1 + p.kb2;
    if (TN) { const int split = w % p.splits; kb_beg = split * p.kb_per_split; kb_end = min(p.kb_total, kb_beg + p.kb_per_split); }
  };

  if (warp == 0) {
    // ======================= TMA producer (each CTA: its 128 rows of A, its half of B) =======================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int w = cluster_id; w < n_work; w += n_clusters) {
        const int tile = TN ? w / p.splits : w;
        const int m0 = (tile / p.tiles_n) * (2 * BM) + (int)rank * BM;
        const int n0 = (tile % p.tiles_n) * p.BN + (int)rank * half_bn;
        int kb_beg, kb_end;
        k_range(w, kb_beg, kb_end);
        for (int kb = kb_beg; kb < kb_end; kb += KSUB) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          const int nsub = min(KSUB, kb_end - kb);
          mbar_arrive_expect_tx(&full_bar[stage], stage_tx_bytes * (uint32_t)nsub);
          for (int j = 0; j < nsub; ++j) {
            uint8_t* sa = smem + stage * STAGE_BYTES + j * SUB_BYTES;
            uint8_t* sb = sa + A_STAGE_BYTES;
            const int kj = kb + j;
            if (!TN) {
              const bool second = kj >= p.kb1;
              const int k0 = (second ? kj - p.kb1 : kj) * BK;
              tma_load_2d(sa, second ? &tmA2 : &tmA1, &full_bar[stage], k0, m0);
              tma_load_2d(sb, second ? &tmB2 : &tmB1, &full_bar[stage], k0, n0);
            } else {
              const int k0 = kj * BK;    // node rows
#pragma unroll
              for (int c = 0; c < BM / 32; ++c) tma_load_2d(sa + c * 4096, &tmA1, &full_bar[stage], m0 + 32 * c, k0);
              for (int b = 0; b < NB; ++b)
                for (int c = 0; c < half_bn / 32; ++c)
                  tma_load_2d(sb + b * 2 * L::B_HALF_BYTES + c * 4096, b ? &tmB2 : &tmB1, &full_bar[stage], n0 + 32 * c, k0);
            }
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ======================= MMA issuer (leader CTA only; the whole warp runs the loop, one elected lane issues) =======================
    if (rank == 0) {
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);          // warp-uniform copy
      const uint32_t idesc = make_idesc_ts(2 * BM, p.BN, TN);
      const uint32_t idesc_bf = make_idesc_ts_bf16(2 * BM, p.BN) | (TN ? (1u << 16) : 0u);      // TN: B MN-major
      // bf16 B tiles of the TN form: MN-major, standard 128-byte swizzle, SBO = 1024 B between 8-k-row atoms
      const uint32_t desc_bf_hi = (uint32_t)(make_smem_desc(0, 2048, 1024, kLayoutSw128) >> 32);
      // descriptor of a B tile at shared-memory address 0; the address field (bits 0..13, 16-byte units) is added per use
      const uint64_t desc0 = make_smem_desc(0, TN ? 4096 : 16, TN ? 512 : 1024, TN ? kLayoutSw128Base32 : kLayoutSw128);
      const uint32_t desc0_lo = (uint32_t)desc0, desc0_hi = (uint32_t)(desc0 >> 32);
      const uint32_t koff16 = (TN ? 1024 : UMMA_K * 4) >> 4;
      const uint32_t sb0 = (smem_u32(smem) + A_STAGE_BYTES) >> 4;
      const uint32_t empty0 = smem_u32(&empty_bar[0]), afree0 = smem_u32(&a_free[0]), tfull0 = smem_u32(&tmem_full[0]);
      int stage = 0;
      int slot = 0; uint32_t sphase = 0;
      int it = 0;
      for (int w = cluster_id; w < n_work; w += n_clusters, ++it) {
        int kb_beg, kb_end;
        k_range(w, kb_beg, kb_end);
        const int acc = DUAL ? 0 : it % ACC_BUFS;              // DUAL: both accumulators belong to the same work item
        const uint32_t acc_phase = DUAL ? (uint32_t)it & 1 : (uint32_t)(it / ACC_BUFS) & 1;
        if (!mbar_test(&tmem_empty[acc], acc_phase ^ 1)) mbar_wait(&tmem_empty[acc], acc_phase ^ 1);   // drained by both CTAs' epilogues
        tcgen05_fence_after();
        const uint32_t d_tmem = tmem_u + (uint32_t)acc * L::BN_MAX;
        for (int kb = kb_beg; kb < kb_end; kb += KSUB) {
          if (!mbar_test(&ready[slot], sphase)) mbar_wait(&ready[slot], sphase);
          tcgen05_fence_after();
          const int nsub = min(KSUB, kb_end - kb);
          for (int j = 0; j < nsub; ++j) {
            const uint32_t b_lo32 = desc0_lo + sb0 + (uint32_t)(stage * STAGE_BYTES + j * SUB_BYTES) / 16;
            const uint32_t a_hi = tmem_u + L::A_COL0 + (uint32_t)(slot * KSUB + j) * L::SLOT_COLS;
#pragma unroll
            for (int b = 0; b < NB; ++b) {
              const uint32_t bb = b_lo32 + (uint32_t)b * (2 * L::B_HALF_BYTES >> 4);
              if (BF == 2)     // A tile of this stage: same descriptor fields as a K-major B tile (128 rows instead of 64)
                mma_x3bf2_block_ts2(d_tmem, a_hi, bb - (A_STAGE_BYTES >> 4), bb, bb + (L::B_HALF_BYTES >> 4), desc0_hi,
                                    make_idesc(2 * BM, p.BN, false), idesc_bf, (kb > kb_beg || j > 0) ? 1u : 0u);
              else if (BF == 1 && TN)      // bf16 tiles: [bf16(B) 4 KB | bf16(B_lo) 4 KB] behind the fp32 half tile
                mma_x3bf_block_ts2_tn(d_tmem + (uint32_t)b * L::BN_MAX, a_hi, bb, koff16, desc0_hi, bb + (L::B_HALF_BYTES >> 4),
                                      bb + (L::B_HALF_BYTES >> 4) + (4096 >> 4), desc_bf_hi, idesc, idesc_bf,
                                      (kb > kb_beg || j > 0) ? 1u : 0u);
              else if (BF == 1)
                mma_x3bf_block_ts2(d_tmem, a_hi, bb, bb + (L::B_HALF_BYTES >> 4), desc0_hi, idesc, idesc_bf,
                                   (kb > kb_beg || j > 0) ? 1u : 0u);
              else
              mma_x3_block_ts2(d_tmem + (uint32_t)b * L::BN_MAX, a_hi, bb, bb + (L::B_HALF_BYTES >> 4), desc0_hi, koff16, idesc,
                               (kb > kb_beg || j > 0) ? 1u : 0u);
            }
          }
          tcgen05_commit_mc2_elect(empty0 + (uint32_t)stage * 8);    // both CTAs: shared-memory stage reusable
          tcgen05_commit_mc2_elect(afree0 + (uint32_t)slot * 8);     // both CTAs: TMEM A slot (and its ready barrier) reusable
          if (++stage == STAGES) stage = 0;
          if (++slot == A_SLOTS) { slot = 0; sphase ^= 1; }
        }
        tcgen05_commit_mc2_elect(tfull0 + (uint32_t)acc * 8);        // both CTAs: accumulator complete -> epilogue
      }
    }
  } else if (warp < 6) {
    // ======================= A split -> TMEM (own 128 rows) =======================
    const int q = warp & 3;
    const int r = q * 32 + lane;
    int stage = 0; uint32_t phase = 0;
    int slot = 0; uint32_t sphase = 0;
    const uint32_t ready_leader = mapa_u32(smem_u32(&ready[0]), 0);
    for (int w = cluster_id; w < n_work; w += n_clusters) {
      int kb_beg, kb_end;
      k_range(w, kb_beg, kb_end);
      float csum = 0.f;
      for (int kb = kb_beg; kb < kb_end; kb += KSUB) {
        mbar_wait(&full_bar[stage], phase);
        const int nsub = min(KSUB, kb_end - kb);
        for (int j = 0; j < nsub; ++j) {
          const uint8_t* sa = smem + stage * STAGE_BYTES + j * SUB_BYTES;
          uint32_t hi[32], lo[32];
          if (!TN) {
            // K-major tile, 128B swizzle: row r at r * 128, 16-byte chunk c stored at chunk c ^ (r & 7)
            const uint32_t row = smem_u32(sa) + (uint32_t)r * 128;
#pragma unroll
            for (int c = 0; c < 8; ++c) {
              const float4 x = lds_128(row + ((uint32_t)(c ^ (r & 7)) << 4));
              hi[4 * c + 0] = __float_as_uint(x.x); hi[4 * c + 1] = __float_as_uint(x.y);
              hi[4 * c + 2] = __float_as_uint(x.z); hi[4 * c + 3] = __float_as_uint(x.w);
            }
          } else {
            // MN-major boxes [32 node rows][32 features], SWIZZLE_128B_ATOM_32B: box q holds this warp's 32
            // features; node row k at k * 128, 32-byte unit u stored at unit u ^ (k & 3)
            const uint32_t box = smem_u32(sa) + (uint32_t)(q * 4096 + (lane & 7) * 4);
            const int u = lane >> 3;
#pragma unroll
            for (int k = 0; k < 32; ++k) {
              const float x = lds_32(box + (uint32_t)(k * 128 + ((u ^ (k & 3)) << 5)));
              hi[k] = __float_as_uint(x);
              csum += x;
            }
          }
          uint32_t lob[16], hib[16];
          if (BF) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const float x0 = __uint_as_float(hi[2 * i]), x1 = __uint_as_float(hi[2 * i + 1]);
              lob[i] = pack_bf16x2(x0 - __uint_as_float(hi[2 * i] & 0xFFFFE000u), x1 - __uint_as_float(hi[2 * i + 1] & 0xFFFFE000u));
              hib[i] = pack_bf16x2(x0, x1);
            }
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) lo[i] = __float_as_uint(tf32_lo(__uint_as_float(hi[i])));
          }
          if (j == 0) {
            mbar_wait(&a_free[slot], sphase ^ 1);      // the MMAs that read this TMEM slot have retired
            tcgen05_fence_after();
          }
          const uint32_t t_a = tmem_base + ((uint32_t)(q * 32) << 16) + L::A_COL0 + (uint32_t)(slot * KSUB + j) * L::SLOT_COLS;
          if (BF != 2) tmem_st_32x32b_x32(t_a, hi);
          if (BF == 2) {
            tmem_st_32x32b_x16(t_a, lob);
            tmem_st_32x32b_x16(t_a + 16, hib);
          } else if (BF == 1) {
            tmem_st_32x32b_x16(t_a + 32, lob);
            tmem_st_32x32b_x16(t_a + 48, hib);
          } else {
            tmem_st_32x32b_x32(t_a + 32, lo);
          }
        }
        tmem_st_wait();
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) { if (rank == 0) mbar_arrive(&ready[slot]); else mbar_arrive_remote(ready_leader + (uint32_t)slot * 8); }
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
        if (++slot == A_SLOTS) { slot = 0; sphase ^= 1; }
      }
      if (TN && p.colsum_partial) {
        const int split = w % p.splits, tile = w / p.splits;
        const int m = (tile / p.tiles_n) * (2 * BM) + (int)rank * BM + r;
        if (tile % p.tiles_n == 0 && m < p.M) p.colsum_partial[(int64_t)split * p.split_stride + m] = csum;
      }
    }
  } else if (warp < 10) {
    // ======================= B split in shared memory (own half of the N rows) =======================
    const int t = threadIdx.x - 6 * 32;
    const int n_vec = half_bn * BK * 4 / 16;
    int stage = 0; uint32_t phase = 0;
    int slot = 0; uint32_t sphase = 0;
    const uint32_t ready_leader = mapa_u32(smem_u32(&ready[0]), 0);
    for (int w = cluster_id; w < n_work; w += n_clusters) {
      int kb_beg, kb_end;
      k_range(w, kb_beg, kb_end);
      for (int kb = kb_beg; kb < kb_end; kb += KSUB) {
        mbar_wait(&full_bar[stage], phase);
        const int nsub = min(KSUB, kb_end - kb);
        {
          for (int j = 0; j < nsub; ++j) {
            for (int b = 0; b < NB; ++b) {
              uint8_t* bt = smem + stage * STAGE_BYTES + j * SUB_BYTES + A_STAGE_BYTES + b * 2 * L::B_HALF_BYTES;
              const float4* hi = reinterpret_cast<const float4*>(bt);
              float4* lo = reinterpret_cast<float4*>(bt + L::B_HALF_BYTES);
              if (BF && TN) {
                // MN-major fp32 boxes [32 k][32 n] (SWIZZLE_128B_ATOM_32B: 32-byte unit u of row k stored at u ^ (k & 3);
                // box c = n / 32 at + 4096 c) -> bf16 tiles [32 k-rows x 128 B] in the standard 128-byte swizzle
                // (16-byte chunk j of row k stored at j ^ (k & 7)): a thread converts one 32-byte unit (8 n-values of
                // one k) into one chunk of bf16(B) and one of bf16(B_lo).  Lane -> (c, u) = lane & 7, k = 4 rows per
                // warp: a quarter-warp reads 8 distinct 16-byte slots (c = 1 lanes take the upper half first) and
                // writes the 8 chunks of one row.
                const int cu = lane & 7, c = cu >> 2, u = cu & 3;
                const uint32_t src0 = smem_u32(hi) + (uint32_t)c * 4096, dst_hi = smem_u32(lo), dst_lo = dst_hi + 4096;
#pragma unroll
                for (int it2 = 0; it2 < 2; ++it2) {
                  const int k = it2 * 16 + (t >> 5) * 4 + (lane >> 3);
                  const uint32_t src = src0 + (uint32_t)(k * 128 + ((u ^ (k & 3)) << 5));
                  const float4 f0 = lds_128(src + (c ? 16u : 0u)), f1 = lds_128(src + (c ? 0u : 16u));
                  const float4 x = c ? f1 : f0, y = c ? f0 : f1;
                  const float v[8] = {x.x, x.y, x.z, x.w, y.x, y.y, y.z, y.w};
                  uint32_t hb[4], lb[4];
#pragma unroll
                  for (int e = 0; e < 4; ++e) {
                    const float a0 = v[2 * e], a1 = v[2 * e + 1];
                    hb[e] = pack_bf16x2(a0, a1);
                    lb[e] = pack_bf16x2(a0 - __uint_as_float(__float_as_uint(a0) & 0xFFFFE000u),
                                        a1 - __uint_as_float(__float_as_uint(a1) & 0xFFFFE000u));
                  }
                  const uint32_t off = (uint32_t)(k * 128 + ((cu ^ (k & 7)) << 4));
                  sts_128(dst_hi + off, hb[0], hb[1], hb[2], hb[3]);
                  sts_128(dst_lo + off, lb[0], lb[1], lb[2], lb[3]);
                }
              } else if (BF) {
                // row n (128 B, 16-byte chunk c stored at c ^ (n & 7)): fp32 chunks 2d, 2d+1 (k = 8d .. 8d+7) ->
                // bf16(B) into chunk d and bf16(B_lo) into chunk 4 + d of the same row of the second tile
                for (int i = t; i < half_bn * 4; i += 128) {
                  const int n = i >> 2, d = i & 3, sw = n & 7;
                  const float4 x = lds_128(smem_u32(hi) + (uint32_t)(n * 8 + ((2 * d) ^ sw)) * 16);
                  const float4 y = lds_128(smem_u32(hi) + (uint32_t)(n * 8 + ((2 * d + 1) ^ sw)) * 16);
                  const float v[8] = {x.x, x.y, x.z, x.w, y.x, y.y, y.z, y.w};
                  uint32_t hb[4], lb[4];
#pragma unroll
                  for (int e = 0; e < 4; ++e) {
                    const float a0 = v[2 * e], a1 = v[2 * e + 1];
                    hb[e] = pack_bf16x2(a0, a1);
                    lb[e] = pack_bf16x2(a0 - __uint_as_float(__float_as_uint(a0) & 0xFFFFE000u),
                                        a1 - __uint_as_float(__float_as_uint(a1) & 0xFFFFE000u));
                  }
                  const uint32_t orow = smem_u32(lo) + (uint32_t)n * 128;
                  sts_128(orow + (uint32_t)(d ^ sw) * 16, hb[0], hb[1], hb[2], hb[3]);
                  sts_128(orow + (uint32_t)((4 + d) ^ sw) * 16, lb[0], lb[1], lb[2], lb[3]);
                }
              } else {
#pragma unroll 4
              for (int i = t; i < n_vec; i += 128) {
                const float4 x = lds_128(smem_u32(hi) + (uint32_t)i * 16);
                sts_128f(smem_u32(lo) + (uint32_t)i * 16, make_float4(tf32_lo(x.x), tf32_lo(x.y), tf32_lo(x.z), tf32_lo(x.w)));
              }
              }
            }
          }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the MMA (async proxy)
        // the ready barrier is per A slot: do not arrive for round i before the phase of round i - A_SLOTS is over
        mbar_wait(&a_free[slot], sphase ^ 1);
        __syncwarp();
        if (lane == 0) { if (rank == 0) mbar_arrive(&ready[slot]); else mbar_arrive_remote(ready_leader + (uint32_t)slot * 8); }
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
        if (++slot == A_SLOTS) { slot = 0; sphase ^= 1; }
      }
    }
  } else {
    // ======================= epilogue warps (own 128 rows of C) =======================
    const int ew = warp - 10;
    const int q = warp & 3;
    float* stg = epi_stage + ew * 32 * EPI_LD;
    const bool masked = !TN && p.act == GTS_ACT_MASK_POS;
    const uint32_t tmem_empty_leader = mapa_u32(smem_u32(&tmem_empty[0]), 0);
    int it = 0;
    for (int w = cluster_id; w < n_work; w += n_clusters, ++it) {
      int tile = w, split = 0;
      if (TN) { split = w % p.splits; tile = w / p.splits; }
      const int m0 = (tile / p.tiles_n) * (2 * BM) + (int)rank * BM + q * 32;
      const int n0 = (tile % p.tiles_n) * p.BN;
      float* Cout = p.C + (TN ? (int64_t)split * p.split_stride : 0);
      const int acc = DUAL ? 0 : it % ACC_BUFS;
      const uint32_t full_phase = DUAL ? (uint32_t)it & 1 : (uint32_t)(it / ACC_BUFS) & 1;
      const uint32_t t_base = tmem_base + (uint32_t)acc * L::BN_MAX + ((uint32_t)(q * 32) << 16);
      epilogue_tile<TN>(p, t_base, m0, n0, Cout, stg, lane, 0, 32, masked, &tmem_full[acc], full_phase);
      if (DUAL)      // second product of the pair (same A): its own accumulator and its own block of the split record
        epilogue_tile<TN>(p, t_base + L::BN_MAX, m0, n0, Cout + p.c2_off, stg, lane, 0, 32, false, &tmem_full[acc], full_phase);
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) { if (rank == 0) mbar_arrive(&tmem_empty[acc]); else mbar_arrive_remote(tmem_empty_leader + (uint32_t)acc * 8); }
    }
  }

  tcgen05_fence_before();
  cluster_sync_all();                                // no CTA may exit (or free TMEM) while its peer still uses it
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

template <int N>
__device__ __forceinline__ void setmaxnreg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N>
__device__ __forceinline__ void setmaxnreg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }

// Epilogue of one 128-column half tile for one warp of the wide kernel, FULL tiles only (all 32 rows inside M; N is
// a multiple of 256, so every column is valid).
//
// What the in-kernel trace showed (profiles/r02_gemm_ntw.md): draining a half tile chunk by chunk (TMEM -> registers
// -> padded smem -> coalesced stores) takes ~5000-7000 clk, and not because of the instruction count: all 148 SMs
// drain at the same moment and 9.5 MB of stores go out at the ~3.5 TB/s the memory system accepts.  With the
// accumulator held until its last chunk was stored, that burst sat on the MMA issuer's critical path (the R half of
// the next item re-uses the accumulator).  Here the warp pulls its whole 32 x 128 slice into REGISTERS (128 per
// thread; the epilogue warpgroup raises its register budget with setmaxnreg), hands the accumulator back at once,
// and then stores at whatever rate the memory system takes while the next item's MMAs run.
// Row pointers are formed once, chunk offsets are immediates, bias is fetched before the accumulator wait, the
// ReLU-mask operand is fetched one chunk ahead row by row (no second register set).
// ACT: GTS_ACT_NONE / GTS_ACT_RELU (bias, bias2 optional) / GTS_ACT_MASK_POS (aux with ldaux == ldc, no bias).
template <int ACT>
__device__ __forceinline__ void epilogue_half_regs(const Params& p, uint32_t t_base, int m0, int n0, float* stg, int lane,
                                                   uint64_t* full_bar, uint32_t full_phase, uint64_t* empty_local,
                                                   uint32_t empty_remote, bool leader) {
  const int cc = (lane & 7) * 4;
  const int rsub = lane >> 3;
  // rows of this 32-row group inside M (32 everywhere but in the last tile); row 4j + rsub is valid iff j < jmax
  // (the float-mask form holds 64 more registers for its double-buffered operand: it takes full row groups only, the
  // caller sends partial ones through the generic path — with the predication it spilled 110 registers)
  const int nrows = min(32, p.M - m0);
  const int jmax = ACT == GTS_ACT_MASK_POS ? 8 : (nrows > rsub ? (nrows - rsub + 3) >> 2 : 0);
  // explicit shared-state-space addresses: through a generic float* the compiler emitted generic LD / ST for the staging
  // tile, which queue behind the outstanding global loads of the mask operand (seen: 14 000+ clk per masked half tile)
  const uint32_t stg_w = smem_u32(stg) + (uint32_t)(lane * EPI_LD) * 4;                 // this lane's row (write side)
  const uint32_t stg_r = smem_u32(stg) + (uint32_t)(rsub * EPI_LD + cc) * 4;            // row rsub, columns cc.. (read side)
  float* c_row = p.C + (int64_t)(m0 + rsub) * p.ldc + n0 + cc;       // row m0 + rsub; row 4j + rsub is j * step4 further
  const int step4 = 4 * (int)p.ldc;                                    // 32-bit element offsets: one IMAD.WIDE per access
  const float* a_row = ACT == GTS_ACT_MASK_POS ? p.aux + (int64_t)(m0 + rsub) * p.ldc + n0 + cc : nullptr;   // ldaux == ldc
  float4 b[4];
  if (ACT == GTS_ACT_NONE || ACT == GTS_ACT_RELU) {
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      b[c] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (p.bias) b[c] = *reinterpret_cast<const float4*>(p.bias + n0 + 32 * c + cc);
      if (p.bias2) {
        const float4 b2 = *reinterpret_cast<const float4*>(p.bias2 + n0 + 32 * c + cc);
        b[c].x += b2.x; b[c].y += b2.y; b[c].z += b2.z; b[c].w += b2.w;
      }
    }
  }
  // mask operand (float form): two register sets, the loads of chunk c+1 all issued before chunk c is stored.  (Refilling
  // aux[j] right after its use — one register set — serialised on the scoreboard: every LDS of the store loop then
  // waited for the previous row's global load, 21 000 clk per half tile.)
  float4 aux[2][8];
  if (ACT == GTS_ACT_MASK_POS) {
#pragma unroll
    for (int j = 0; j < 8; ++j)
      aux[0][j] = j < jmax ? ldg_nc_na(reinterpret_cast<const float4*>(a_row + j * step4)) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  // mask operand (bit form): the four words (one per 32-column chunk of this half tile) of each of this lane's 8 rows,
  // fetched before the accumulator is even complete
  uint4 mw[8];
  if (ACT == GTS_ACT_MASK_BITS) {
    const uint32_t* brow = p.aux_bits + (int64_t)(m0 + rsub) * p.ld_aux_bits + (n0 >> 5);
    const int64_t bstep4 = 4 * p.ld_aux_bits;
#pragma unroll
    for (int j = 0; j < 8; ++j)
      mw[j] = j < jmax ? *reinterpret_cast<const uint4*>(brow + j * bstep4) : make_uint4(0u, 0u, 0u, 0u);
  }
  // scatter form: the arg-max rows and the mask words of this lane's elements, fetched per chunk (one register set: with
  // 128 accumulator registers live, a second set spilled)
  int4 sidx[8];
  uint32_t sw[8];
  const int32_t* i_row = ACT == GTS_ACT_MASK_BITS_SCATTER ? p.scatter_idx + (int64_t)(m0 + rsub) * p.ld_idx + n0 + cc : nullptr;
  const int istep4 = 4 * (int)p.ld_idx;
  const uint32_t* sb_row = ACT == GTS_ACT_MASK_BITS_SCATTER ? p.aux_bits + (int64_t)(m0 + rsub) * p.ld_aux_bits + (n0 >> 5) : nullptr;
  const int sbstep4 = 4 * (int)p.ld_aux_bits;
  float* s_col = ACT == GTS_ACT_MASK_BITS_SCATTER ? p.scatter_out + n0 + cc : nullptr;      // column of this lane in the destination
  const uint32_t s_ld_bytes = (uint32_t)(p.ld_out * 4);
  mbar_wait(full_bar, full_phase);
  tcgen05_fence_after();
  uint32_t v[4][32];
#pragma unroll
  for (int c = 0; c < 4; ++c) tmem_ld_32x32b_x32(t_base + 32 * c, v[c]);
  tmem_ld_wait();
  tcgen05_fence_before();
  __syncwarp();
  if (lane == 0) { if (leader) mbar_arrive(empty_local); else mbar_arrive_remote(empty_remote); }   // accumulator free again
  uint32_t* bits_row = (ACT == GTS_ACT_RELU && p.bits_out) ? p.bits_out + (int64_t)(m0 + rsub) * p.ld_bits_out + (n0 >> 5) : nullptr;
  const int64_t bits_step4 = 4 * p.ld_bits_out;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    if (ACT == GTS_ACT_MASK_POS && c < 3) {
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (j < jmax) aux[(c + 1) & 1][j] = ldg_nc_na(reinterpret_cast<const float4*>(a_row + j * step4 + 32 * (c + 1)));
    }
    if (ACT == GTS_ACT_MASK_BITS_SCATTER) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        sidx[j] = j < jmax ? __ldg(reinterpret_cast<const int4*>(i_row + j * istep4 + 32 * c)) : make_int4(-1, -1, -1, -1);
        sw[j] = j < jmax ? __ldg(sb_row + j * sbstep4 + c) : 0u;
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j)
      sts_128(stg_w + 16 * j, v[c][4 * j], v[c][4 * j + 1], v[c][4 * j + 2], v[c][4 * j + 3]);
    __syncwarp();
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float4 x = lds_128(stg_r + j * (4 * EPI_LD * 4));
      if (ACT == GTS_ACT_MASK_POS) {
        const float4 a = aux[c & 1][j];
        x.x = a.x > 0.f ? x.x : 0.f; x.y = a.y > 0.f ? x.y : 0.f;
        x.z = a.z > 0.f ? x.z : 0.f; x.w = a.w > 0.f ? x.w : 0.f;
      } else if (ACT == GTS_ACT_MASK_BITS || ACT == GTS_ACT_MASK_BITS_SCATTER) {
        // bit 8 * comp + (lane & 7) of the chunk's word <-> column 4 * (lane & 7) + comp
        const uint32_t w = (ACT == GTS_ACT_MASK_BITS_SCATTER ? sw[j]
                                                             : (c == 0 ? mw[j].x : c == 1 ? mw[j].y : c == 2 ? mw[j].z : mw[j].w)) >> (lane & 7);
        x.x = (w & 0x1u) ? x.x : 0.f; x.y = (w & 0x100u) ? x.y : 0.f;
        x.z = (w & 0x10000u) ? x.z : 0.f; x.w = (w & 0x1000000u) ? x.w : 0.f;
        if (ACT == GTS_ACT_MASK_BITS_SCATTER) {
          // dP[arg[v,k], k] += x: the backward of the neighbour max, straight out of the GEMM that produces its input
          // (zeros — half of the tile after the mask — and rows without an arg-max are skipped)
          const int4 a = sidx[j];
          auto dst = [&](int32_t u, int comp) {
            uint64_t addr;
            asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(addr) : "r"((uint32_t)u), "r"(s_ld_bytes), "l"(s_col + 32 * c + comp));
            return reinterpret_cast<float*>(addr);
          };
          if (x.x != 0.f && a.x >= 0) red_add_f32(dst(a.x, 0), x.x);
          if (x.y != 0.f && a.y >= 0) red_add_f32(dst(a.y, 1), x.y);
          if (x.z != 0.f && a.z >= 0) red_add_f32(dst(a.z, 2), x.z);
          if (x.w != 0.f && a.w >= 0) red_add_f32(dst(a.w, 3), x.w);
        }
      } else {
        x.x += b[c].x; x.y += b[c].y; x.z += b[c].z; x.w += b[c].w;
        if (ACT == GTS_ACT_RELU) {
          x.x = fmaxf(x.x, 0.f); x.y = fmaxf(x.y, 0.f); x.z = fmaxf(x.z, 0.f); x.w = fmaxf(x.w, 0.f);
        }
      }
      if (ACT == GTS_ACT_RELU && bits_row) {
        // the ReLU mask of this output as bits, for the backward's data-gradient GEMM: lanes 8 * rsub .. + 7 hold the 32
        // columns of row 4j + rsub; byte rsub of each component's ballot, component-major, is that row's word
        const uint32_t b0 = __ballot_sync(0xffffffffu, x.x > 0.f), b1 = __ballot_sync(0xffffffffu, x.y > 0.f);
        const uint32_t b2 = __ballot_sync(0xffffffffu, x.z > 0.f), b3 = __ballot_sync(0xffffffffu, x.w > 0.f);
        const uint32_t sel = (uint32_t)rsub | ((4u + (uint32_t)rsub) << 4);
        const uint32_t w01 = __byte_perm(b0, b1, sel), w23 = __byte_perm(b2, b3, sel);
        if ((lane & 7) == 0 && j < jmax) bits_row[j * bits_step4 + c] = __byte_perm(w01, w23, 0x5410);
      }
      if (ACT != GTS_ACT_MASK_BITS_SCATTER && j < jmax && !GTS_DBG(0)) *reinterpret_cast<float4*>(c_row + j * step4 + 32 * c) = x;
    }
    __syncwarp();
  }
}

// ---------------------------------------------------------------------------
// NT "wide" form of the CTA-pair kernel (round 2): ONE pass over a 256-row A tile feeds BOTH 128-column halves of a
// 256-wide output tile.
//
// Why (profiles/r01_gemm_x3_pipeline.md): the N = 128 pair kernel above fetches every A tile
// twice per 256-wide output (once per N tile; 24 KB of TMA traffic per CTA for 512 clk of tensor work) and moves
// 88 KB through shared memory per 512 clk; its time is the L2 -> SM operand delivery (TMA + barrier skeleton = 3/4 of
// the kernel).  Here a k-block lands A (16 KB) once next to BOTH B halves (2 x 8 KB per CTA) and the split A tile in
// TMEM feeds 16 MMAs instead of 8: 32 KB of TMA traffic and 144 KB of shared-memory traffic per 1024 clk of tensor
// work, i.e. 1.5x / 1.25x less per flop.
//
// TMEM (512 columns): THREE 128-column accumulators, rotated over the half tiles (half tile h -> accumulator h % 3),
// so the epilogue of item i (L half first, then R) overlaps the main loop of item i+1: the L half of item i+1 goes
// to the spare accumulator, its R half re-uses the accumulator item i's L half drains first.  The A ring lives in
// the remaining 128 columns as FOUR slots of HALF a k-block (16 k-values x 128 rows: 16 columns fp32 A | 8 columns
// bf16x2(A_lo) | 8 columns bf16x2(A)): a slot feeds 2 x (2 TF32 + 2 bf16) MMAs = 512 clk, and three slots of tensor
// work cover the hand-over (multicast commit -> a_free -> tcgen05.st -> remote arrive) of the fourth.
// Shared memory: 4 stages of (A 16 | B_L fp32 8 | B_L bf16 8 | B_R fp32 8 | B_R bf16 8 KB) = 192 KB.
// Warp roles (512 threads, warpgroup-aligned for setmaxnreg): 0 TMA, 1 MMA issuer (leader CTA), 2..3 idle,
// 4..7 A split -> TMEM, 8..11 B split, 12..15 epilogue.
// Requires N % 256 == 0 (each work item = 256 rows x 256 columns); other shapes stay on the N = 128 pair kernel.
// ---------------------------------------------------------------------------
struct NtwCfg {
  static constexpr int STAGES = 4;
  static constexpr int A_SLOTS = 4;                                    // TMEM ring of split half-k-block A tiles
  static constexpr int SLOT_COLS = 32;
  static constexpr int ACC_BUFS = 3;
  static constexpr int BN_HALF = 128;                                  // UMMA N of one half tile
  static constexpr int EPI_WARPS = 4;
  static constexpr int B_HALF_BYTES = (BN_HALF / 2) * BK * 4;          // 8 KB: this CTA's 64 rows of one half tile
  static constexpr int STAGE_BYTES = A_STAGE_BYTES + 2 * 2 * B_HALF_BYTES;   // 48 KB
  // Roles are WARPGROUP-aligned so that setmaxnreg can move registers to where they are needed:
  // warps 0..3: TMA producer (0), MMA issuer (1), two idle warps;  4..7: A split;  8..11: B split;  12..15: epilogue.
  static constexpr int THREADS = 16 * 32;
  static constexpr int REGS_CTRL = 80, REGS_ASPLIT = 104, REGS_BSPLIT = 72, REGS_EPI = 240;   // sum x 128 <= 64 K
  static_assert(128 * (REGS_CTRL + REGS_ASPLIT + REGS_BSPLIT + REGS_EPI) <= 65536, "register file budget");
  static constexpr int epi_off = STAGES * STAGE_BYTES;
  static constexpr int epi_bytes = EPI_WARPS * 32 * EPI_LD * 4;
  static constexpr int bar_off = epi_off + epi_bytes;
  // full[S], empty[S], ready[A], a_free[A], tmem_full[3], tmem_empty[3], tmem_ptr
  static constexpr int total = bar_off + (2 * STAGES + 2 * A_SLOTS + 2 * ACC_BUFS) * 8 + 16;
  static constexpr int dyn_bytes = total + 1024;
  static_assert(dyn_bytes <= 232448, "exceeds the 227 KB shared-memory limit per CTA");
  static constexpr uint32_t A_COL0 = ACC_BUFS * BN_HALF;               // 384
  static_assert(A_COL0 + SLOT_COLS * A_SLOTS <= TMEM_COLS, "TMEM column budget");
};

__device__ __forceinline__ void tmem_st_32x32b_x8(uint32_t taddr, const uint32_t (&v)[8]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
      : "memory");
}

// The four MMAs of HALF a k-block (16 k-values) into one 128-column half tile: hi*hi as two TF32 MMAs (K = 8) on the
// raw fp32 operands, A_lo*B and A*B_lo as one bf16 MMA (K = 16) each.  a_tm: TMEM slot ([0,16) fp32 | [16,24)
// bf16x2(A_lo) | [24,32) bf16x2(A)); b_lo32 / h_lo32: low descriptor words of the fp32 B tile at this half's first
// k-step and of the bf16 tile [bf16(B) | bf16(B_lo)] at this half's 32-byte chunk.  Whole-warp call, one elected
// lane issues (see mma_x3_block_ts2).
__device__ __forceinline__ void mma_x3bf_half_ts2(uint32_t d_tmem, uint32_t a_tm, uint32_t b_lo32, uint32_t h_lo32,
                                                  uint32_t desc_hi32, uint32_t idesc, uint32_t idesc_bf, uint32_t first) {
  asm volatile(
      "{\n\t"
      ".reg .pred pf, pt, pe;\n\t"
      ".reg .b32 x1, y1, a1, a2, a3;\n\t"
      ".reg .b64 b0, b1, h0, l0;\n\t"
      "elect.sync _|pe, 0xffffffff;\n\t"
      "setp.ne.b32 pf, %7, 0;\n\t"
      "setp.eq.b32 pt, %7, %7;\n\t"
      "add.u32 x1, %2, 2;\n\t"
      "add.u32 y1, %3, 4;\n\t"
      "mov.b64 b0, {%2, %4};\n\t mov.b64 b1, {x1, %4};\n\t"
      "mov.b64 h0, {%3, %4};\n\t mov.b64 l0, {y1, %4};\n\t"
      "add.u32 a1, %1, 8;\n\t add.u32 a2, %1, 16;\n\t add.u32 a3, %1, 24;\n\t"
      "@pe tcgen05.mma.cta_group::2.kind::tf32 [%0], [%1], b0, %5, pf;\n\t"
      "@pe tcgen05.mma.cta_group::2.kind::tf32 [%0], [a1], b1, %5, pt;\n\t"
      "@pe tcgen05.mma.cta_group::2.kind::f16 [%0], [a2], h0, %6, pt;\n\t"
      "@pe tcgen05.mma.cta_group::2.kind::f16 [%0], [a3], l0, %6, pt;\n\t"
      "}"
      ::"r"(d_tmem), "r"(a_tm), "r"(b_lo32), "r"(h_lo32), "r"(desc_hi32), "r"(idesc), "r"(idesc_bf), "r"(first) : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NtwCfg::THREADS, 1)
gemm_x3ntw_kernel(const __grid_constant__ CUtensorMap tmA1, const __grid_constant__ CUtensorMap tmA2,
                  const __grid_constant__ CUtensorMap tmB1, const __grid_constant__ CUtensorMap tmB2, const Params p) {
  using L = NtwCfg;
  constexpr int STAGES = L::STAGES;
  constexpr int A_SLOTS = L::A_SLOTS;
  constexpr int ACC_BUFS = L::ACC_BUFS;
  constexpr int STAGE_BYTES = L::STAGE_BYTES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::bar_off);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* ready = full_bar + 2 * STAGES;           // leader's copy collects both CTAs' arrivals
  uint64_t* a_free = ready + A_SLOTS;
  uint64_t* tmem_full = a_free + A_SLOTS;
  uint64_t* tmem_empty = tmem_full + ACC_BUFS;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + ACC_BUFS);
  float* epi_stage = reinterpret_cast<float*>(smem + L::epi_off);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);    // provably warp-uniform
  const int lane = threadIdx.x & 31;
  if (threadIdx.x == 0) GTS_TR(7, 0, clock64());     // trace: kernel entry (record 7 = whole-kernel marks)
  const uint32_t rank = cluster_ctarank();           // 0 = leader (issues the MMAs), 1 = peer
  const int cluster_id = blockIdx.x >> 1;
  const int n_clusters = gridDim.x >> 1;
  const uint32_t smem_base = smem_u32(smem);         // operand tiles are read / written with explicit ld.shared / st.shared

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);                   // multicast commit of the leader
    }
    for (int s = 0; s < A_SLOTS; ++s) {
      // even slots = first half of a k-block: 4 A-split + 4 B-split warps per CTA; odd slots: the A-split warps only
      mbar_init(&ready[s], (s & 1) ? 8 : 16);
      mbar_init(&a_free[s], 1);                      // multicast commit of the leader
    }
    for (int a = 0; a < ACC_BUFS; ++a) {
      mbar_init(&tmem_full[a], 1);
      mbar_init(&tmem_empty[a], 2 * L::EPI_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    prefetch_tmap(&tmA1); prefetch_tmap(&tmB1);
    if (p.kb2 > 0) { prefetch_tmap(&tmA2); prefetch_tmap(&tmB2); }
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  cluster_sync_all();                                // barriers initialised and TMEM allocated in BOTH CTAs
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  if (threadIdx.x == 0) GTS_TR(7, 1, clock64());     // trace: set-up done

  // work item w -> (pair tile of 256 rows, 256-column tile); tiles_m counts 256-row pair tiles, tiles_n 256-column tiles
  const int n_work = p.tiles_m * p.tiles_n;
  const int n_kb = p.kb1 + p.kb2;
  constexpr uint32_t stage_tx_bytes = (uint32_t)A_STAGE_BYTES + 2u * L::B_HALF_BYTES;

  // Registers follow the roles (the epilogue warpgroup holds a 32 x 128 accumulator slice per warp).  Each setmaxnreg
  // sits INSIDE its warpgroup's branch: ptxas sizes the register allocation of the code a setmaxnreg dominates, and a
  // join after the four instructions would fall back to the launch bound (seen: 1184 bytes of spills).
  if (warp < 4) {
  setmaxnreg_dec<L::REGS_CTRL>();
  if (warp == 0) {
    // ======================= TMA producer (each CTA: its 128 rows of A, its 64 rows of each B half) =======================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int w = cluster_id; w < n_work; w += n_clusters) {
        const int m0 = (w / p.tiles_n) * (2 * BM) + (int)rank * BM;
        const int n0 = (w % p.tiles_n) * (2 * L::BN_HALF) + (int)rank * (L::BN_HALF / 2);
        [[maybe_unused]] long long tr_empty = 0;
        [[maybe_unused]] const int tr_item = (w - cluster_id) / n_clusters;
        for (int kb = 0; kb < n_kb; ++kb) {
          if (GTS_TR_ON) { const long long t0 = clock64(); mbar_wait(&empty_bar[stage], phase ^ 1); tr_empty += clock64() - t0; }
          else mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_arrive_expect_tx(&full_bar[stage], stage_tx_bytes);
          uint8_t* sa = smem + stage * STAGE_BYTES;
          uint8_t* sb = sa + A_STAGE_BYTES;
          const bool second = kb >= p.kb1;
          const int k0 = (second ? kb - p.kb1 : kb) * BK;
          tma_load_2d(sa, second ? &tmA2 : &tmA1, &full_bar[stage], k0, m0);
          tma_load_2d(sb, second ? &tmB2 : &tmB1, &full_bar[stage], k0, n0);
          tma_load_2d(sb + 2 * L::B_HALF_BYTES, second ? &tmB2 : &tmB1, &full_bar[stage], k0, n0 + L::BN_HALF);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        GTS_TR(tr_item, 6, tr_empty); GTS_TR(tr_item, 7, clock64());
      }
    }
  } else if (warp == 1) {
    // ======================= MMA issuer (leader CTA only; the whole warp runs the loop, one elected lane issues) =======================
    if (rank == 0) {
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);          // warp-uniform copy
      const uint32_t idesc = make_idesc_ts(2 * BM, L::BN_HALF, false);
      const uint32_t idesc_bf = make_idesc_ts_bf16(2 * BM, L::BN_HALF);
      const uint64_t desc0 = make_smem_desc(0, 16, 1024, kLayoutSw128);
      const uint32_t desc0_lo = (uint32_t)desc0, desc0_hi = (uint32_t)(desc0 >> 32);
      const uint32_t sb0 = (smem_u32(smem) + A_STAGE_BYTES) >> 4;
      const uint32_t empty0 = smem_u32(&empty_bar[0]), afree0 = smem_u32(&a_free[0]), tfull0 = smem_u32(&tmem_full[0]);
      int stage = 0;
      uint32_t slot_ctr = 0;                         // half k-blocks issued so far: slot = ctr & 3, phase = (ctr >> 2) & 1
      uint32_t h = 0;                                // half tiles issued so far: accumulator = h % 3, use count = h / 3
      for (int w = cluster_id; w < n_work; w += n_clusters, h += 2) {
        const uint32_t accL = h % 3u, accR = (h + 1u) % 3u;
        const uint32_t phL = ((h / 3u) & 1u) ^ 1u, phR = (((h + 1u) / 3u) & 1u) ^ 1u;
        [[maybe_unused]] const int tr_item = (int)(h >> 1);
        [[maybe_unused]] long long tr_ready = 0;
        if (lane == 0) GTS_TR(tr_item, 0, clock64());
        if (!mbar_test(&tmem_empty[accL], phL)) mbar_wait(&tmem_empty[accL], phL);   // drained by both CTAs' epilogues
        tcgen05_fence_after();
        if (lane == 0) GTS_TR(tr_item, 1, clock64());
        const uint32_t dL = tmem_u + accL * L::BN_HALF, dR = tmem_u + accR * L::BN_HALF;
        // The R half re-uses the accumulator that the previous item's L half drains FIRST; that drain is still in
        // flight when this item starts.  Head of the item: the L-half MMAs of the first (up to) four half-k-block slots
        // go out back to back (1024 clk of tensor work that needs only the spare accumulator and the A ring's
        // depth), then the wait for the R accumulator, then the R-half MMAs of the same slots (which free them).
        const int head_kb = n_kb < 2 ? n_kb : (GTS_DBG(4) ? 0 : 2);
        const int stage_head = stage;
        const uint32_t ctr_head = slot_ctr;
        for (int kb = 0; kb < head_kb; ++kb) {
          const uint32_t bL = desc0_lo + sb0 + (uint32_t)(stage * STAGE_BYTES) / 16;
#pragma unroll
          for (int hs = 0; hs < 2; ++hs, ++slot_ctr) {
            const uint32_t slot = slot_ctr & (A_SLOTS - 1), sphase = (slot_ctr >> 2) & 1u;
            if (GTS_TR_ON) {
              const long long t0 = clock64();
              if (!mbar_test(&ready[slot], sphase)) mbar_wait(&ready[slot], sphase);
              tr_ready += clock64() - t0;
            } else if (!mbar_test(&ready[slot], sphase)) mbar_wait(&ready[slot], sphase);
            tcgen05_fence_after();
            mma_x3bf_half_ts2(dL, tmem_u + L::A_COL0 + slot * L::SLOT_COLS, bL + 4u * hs,
                              bL + (L::B_HALF_BYTES >> 4) + 2u * hs, desc0_hi, idesc, idesc_bf, (kb > 0 || hs > 0) ? 1u : 0u);
          }
          if (++stage == STAGES) stage = 0;
        }
        if (lane == 0) GTS_TR(tr_item, 2, clock64());
        if (!mbar_test(&tmem_empty[accR], phR)) mbar_wait(&tmem_empty[accR], phR);
        tcgen05_fence_after();
        if (lane == 0) GTS_TR(tr_item, 3, clock64());
        stage = stage_head;
        slot_ctr = ctr_head;
        for (int kb = 0; kb < head_kb; ++kb) {
          const uint32_t bR = desc0_lo + sb0 + (uint32_t)(stage * STAGE_BYTES) / 16 + (2 * L::B_HALF_BYTES >> 4);
#pragma unroll
          for (int hs = 0; hs < 2; ++hs, ++slot_ctr) {
            const uint32_t slot = slot_ctr & (A_SLOTS - 1);
            mma_x3bf_half_ts2(dR, tmem_u + L::A_COL0 + slot * L::SLOT_COLS, bR + 4u * hs,
                              bR + (L::B_HALF_BYTES >> 4) + 2u * hs, desc0_hi, idesc, idesc_bf, (kb > 0 || hs > 0) ? 1u : 0u);
            tcgen05_commit_mc2_elect(afree0 + slot * 8);             // both CTAs: TMEM A slot (and its ready barrier) reusable
          }
          tcgen05_commit_mc2_elect(empty0 + (uint32_t)stage * 8);    // both CTAs: shared-memory stage reusable
          if (++stage == STAGES) stage = 0;
        }
        for (int kb = head_kb; kb < n_kb; ++kb) {
          const uint32_t bL = desc0_lo + sb0 + (uint32_t)(stage * STAGE_BYTES) / 16;   // fp32 tile of the L half
          const uint32_t bR = bL + (2 * L::B_HALF_BYTES >> 4);
#pragma unroll
          for (int hs = 0; hs < 2; ++hs, ++slot_ctr) {
            const uint32_t slot = slot_ctr & (A_SLOTS - 1), sphase = (slot_ctr >> 2) & 1u;
            if (GTS_TR_ON) {
              const long long t0 = clock64();
              if (!mbar_test(&ready[slot], sphase)) mbar_wait(&ready[slot], sphase);
              tr_ready += clock64() - t0;
            } else if (!mbar_test(&ready[slot], sphase)) mbar_wait(&ready[slot], sphase);
            tcgen05_fence_after();
            const uint32_t a_tm = tmem_u + L::A_COL0 + slot * L::SLOT_COLS;
            const uint32_t first = (kb > 0 || hs > 0) ? 1u : 0u;
            mma_x3bf_half_ts2(dL, a_tm, bL + 4u * hs, bL + (L::B_HALF_BYTES >> 4) + 2u * hs, desc0_hi, idesc, idesc_bf, first);
            mma_x3bf_half_ts2(dR, a_tm, bR + 4u * hs, bR + (L::B_HALF_BYTES >> 4) + 2u * hs, desc0_hi, idesc, idesc_bf, first);
            tcgen05_commit_mc2_elect(afree0 + slot * 8);             // both CTAs: TMEM A slot (and its ready barrier) reusable
          }
          tcgen05_commit_mc2_elect(empty0 + (uint32_t)stage * 8);    // both CTAs: shared-memory stage reusable
          if (++stage == STAGES) stage = 0;
        }
        tcgen05_commit_mc2_elect(tfull0 + accL * 8);                 // both CTAs: accumulators complete -> epilogue
        tcgen05_commit_mc2_elect(tfull0 + accR * 8);
        if (lane == 0) { GTS_TR(tr_item, 4, clock64()); GTS_TR(tr_item, 5, tr_ready); }
      }
    }
  } else if (p.zero_fill != nullptr) {
    // ======================= warps 2, 3: side job — clear a buffer the NEXT kernels need zeroed =======================
    // (the whole-stack backward hands the next layer's dP here: 92 MB of stores ride under a tensor-bound main loop
    // instead of a 16 us memset pass between two kernels; plain coalesced 16-byte stores, 512 B per warp instruction)
    const unsigned long long stride = (unsigned long long)gridDim.x * 64ull;
    unsigned long long i = (unsigned long long)blockIdx.x * 64ull + (unsigned long long)((warp - 2) * 32 + lane);
    const uint4 z = make_uint4(0u, 0u, 0u, 0u);
    for (; i + 3ull * stride < p.zero_n16; i += 4ull * stride) {
      p.zero_fill[i] = z; p.zero_fill[i + stride] = z; p.zero_fill[i + 2ull * stride] = z; p.zero_fill[i + 3ull * stride] = z;
    }
    for (; i < p.zero_n16; i += stride) p.zero_fill[i] = z;
  }
  } else if (warp < 8) {
    // ======================= A split -> TMEM (own 128 rows), two half-k-block slots per stage =======================
    setmaxnreg_dec<L::REGS_ASPLIT>();
    const int q = warp & 3;
    const int r = q * 32 + lane;
    int stage = 0; uint32_t phase = 0;
    uint32_t slot_ctr = 0;
    const uint32_t ready_leader = mapa_u32(smem_u32(&ready[0]), 0);
    const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + L::A_COL0;
    for (int w = cluster_id; w < n_work; w += n_clusters) {
      [[maybe_unused]] long long tr_full = 0, tr_afree = 0;
      for (int kb = 0; kb < n_kb; ++kb) {
        if (GTS_TR_ON) { const long long t0 = clock64(); mbar_wait(&full_bar[stage], phase); tr_full += clock64() - t0; }
        else mbar_wait(&full_bar[stage], phase);
        // K-major tile, 128B swizzle: row r at r * 128, 16-byte chunk c stored at chunk c ^ (r & 7)
        const uint32_t row = smem_base + (uint32_t)(stage * STAGE_BYTES + r * 128);      // shared-state-space address
        uint32_t hi[32];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const float4 x = lds_128(row + ((uint32_t)(c ^ (r & 7)) << 4));
          hi[4 * c + 0] = __float_as_uint(x.x); hi[4 * c + 1] = __float_as_uint(x.y);
          hi[4 * c + 2] = __float_as_uint(x.z); hi[4 * c + 3] = __float_as_uint(x.w);
        }
#pragma unroll
        for (int hs = 0; hs < 2; ++hs, ++slot_ctr) {
          uint32_t f32[16], lob[8], hib[8];
#pragma unroll
          for (int i = 0; i < 16; ++i) f32[i] = hi[16 * hs + i];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const uint32_t u0 = hi[16 * hs + 2 * i], u1 = hi[16 * hs + 2 * i + 1];
            const float x0 = __uint_as_float(u0), x1 = __uint_as_float(u1);
            lob[i] = pack_bf16x2(x0 - __uint_as_float(u0 & 0xFFFFE000u), x1 - __uint_as_float(u1 & 0xFFFFE000u));
            hib[i] = pack_bf16x2(x0, x1);
          }
          const uint32_t slot = slot_ctr & (A_SLOTS - 1), sphase = (slot_ctr >> 2) & 1u;
          if (GTS_TR_ON) { const long long t0 = clock64(); mbar_wait(&a_free[slot], sphase ^ 1); tr_afree += clock64() - t0; }
          else mbar_wait(&a_free[slot], sphase ^ 1);      // the MMAs that read this TMEM slot have retired
          tcgen05_fence_after();
          const uint32_t t_a = t_row + slot * L::SLOT_COLS;
          tmem_st_32x32b_x16(t_a, f32);
          tmem_st_32x32b_x8(t_a + 16, lob);
          tmem_st_32x32b_x8(t_a + 24, hib);
          tmem_st_wait();
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) { if (rank == 0) mbar_arrive(&ready[slot]); else mbar_arrive_remote(ready_leader + slot * 8); }
        }
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      if (GTS_TR_ON && warp == 4 && lane == 0) {
        const int tr_item = (w - cluster_id) / n_clusters;
        GTS_TR(tr_item, 14, tr_full); GTS_TR(tr_item, 15, tr_afree);
      }
    }
  } else if (warp < 12) {
    // ======================= B split in shared memory (own 64 rows of both half tiles) =======================
    setmaxnreg_dec<L::REGS_BSPLIT>();
    const int t = threadIdx.x - 8 * 32;
    int stage = 0; uint32_t phase = 0;
    uint32_t slot_ctr = 0;
    const uint32_t ready_leader = mapa_u32(smem_u32(&ready[0]), 0);
    for (int w = cluster_id; w < n_work; w += n_clusters) {
      for (int kb = 0; kb < n_kb; ++kb, slot_ctr += 2) {
        mbar_wait(&full_bar[stage], phase);
#pragma unroll
        for (int b = 0; b < 2; ++b) {
          const uint32_t hi = smem_base + (uint32_t)(stage * STAGE_BYTES + A_STAGE_BYTES + b * 2 * L::B_HALF_BYTES);
          const uint32_t lo = hi + L::B_HALF_BYTES;
          // row n (128 B, 16-byte chunk c stored at c ^ (n & 7)): fp32 chunks 2d, 2d+1 (k = 8d .. 8d+7) ->
          // bf16(B) into chunk d and bf16(B_lo) into chunk 4 + d of the same row of the bf16 tile
#pragma unroll
          for (int i = t; i < (L::BN_HALF / 2) * 4; i += 128) {
            // lane -> (row n, 32-byte k-group d).  The two rows of a quarter-warp (8 lanes = one shared-memory
            // wavefront of a 128-bit access) are n and n ^ 5: their swizzle terms differ in bit 0, so the loads
            // (chunks {0,2,4,6} ^ sw and {1,3,5,7} ^ sw) do not collide, and in bit 2, so the stores (chunks {0..3} ^ sw
            // and {4..7} ^ sw) do not either.  (Rows n, n + 1 per quarter-warp: every store was a 2-way bank conflict,
            // ncu: 2.1 M of 8.8 M shared-memory wavefronts of the kernel.)
            const int d = i & 3, r8 = (i >> 3) & 3, n = (i >> 5) * 8 + (r8 ^ (((i >> 2) & 1) * 5)), sw = n & 7;
            const float4 x = lds_128(hi + (uint32_t)(n * 8 + ((2 * d) ^ sw)) * 16), y = lds_128(hi + (uint32_t)(n * 8 + ((2 * d + 1) ^ sw)) * 16);
            const float v[8] = {x.x, x.y, x.z, x.w, y.x, y.y, y.z, y.w};
            uint32_t hb[4], lb[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float a0 = v[2 * e], a1 = v[2 * e + 1];
              hb[e] = pack_bf16x2(a0, a1);
              lb[e] = pack_bf16x2(a0 - __uint_as_float(__float_as_uint(a0) & 0xFFFFE000u),
                                  a1 - __uint_as_float(__float_as_uint(a1) & 0xFFFFE000u));
            }
            const uint32_t orow = lo + (uint32_t)n * 128;
            sts_128(orow + (uint32_t)(d ^ sw) * 16, hb[0], hb[1], hb[2], hb[3]);
            sts_128(orow + (uint32_t)((4 + d) ^ sw) * 16, lb[0], lb[1], lb[2], lb[3]);
          }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the MMA (async proxy)
        // the ready barrier is per A slot: do not arrive for round i before the phase of round i - A_SLOTS is over
        const uint32_t slot = slot_ctr & (A_SLOTS - 1), sphase = (slot_ctr >> 2) & 1u;
        mbar_wait(&a_free[slot], sphase ^ 1);
        __syncwarp();
        if (lane == 0) { if (rank == 0) mbar_arrive(&ready[slot]); else mbar_arrive_remote(ready_leader + slot * 8); }
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else {
    // ======================= epilogue warps (own 128 rows of C; L half, then R half) =======================
    setmaxnreg_inc<L::REGS_EPI>();
    const int ew = warp - 12;
    const int q = warp & 3;
    float* stg = epi_stage + ew * 32 * EPI_LD;
    const bool masked = p.act == GTS_ACT_MASK_POS;
    const uint32_t tmem_empty_leader = mapa_u32(smem_u32(&tmem_empty[0]), 0);
    // fast path: full 32-row groups, the mask operand laid out like C
    const bool fast_ok = !GTS_DBG(3) && (p.act != GTS_ACT_MASK_POS || p.ldaux == p.ldc);
    uint32_t h = 0;
    for (int w = cluster_id; w < n_work; w += n_clusters) {
      const int m0 = (w / p.tiles_n) * (2 * BM) + (int)rank * BM + q * 32;
#pragma unroll 1
      for (int half = 0; half < 2; ++half, ++h) {
        const uint32_t acc = h % 3u, full_phase = (h / 3u) & 1u;
        const int n0 = (w % p.tiles_n) * (2 * L::BN_HALF) + half * L::BN_HALF;
        const uint32_t t_base = tmem_base + acc * L::BN_HALF + ((uint32_t)(q * 32) << 16);
        if (GTS_TR_ON && ew == 0) {
          if (lane == 0) GTS_TR((int)(h >> 1), 8 + 3 * half, clock64());
          mbar_wait(&tmem_full[acc], full_phase);
          if (lane == 0) GTS_TR((int)(h >> 1), 9 + 3 * half, clock64());
        }
        if (fast_ok && m0 < p.M && (p.act != GTS_ACT_MASK_POS || m0 + 32 <= p.M)) {
          // whole slice into registers, accumulator released inside, stores afterwards (partial row groups predicated)
#define GTS_EPI(ACT_) epilogue_half_regs<ACT_>(p, t_base, m0, n0, stg, lane, &tmem_full[acc], full_phase, &tmem_empty[acc], \
                                               tmem_empty_leader + acc * 8, rank == 0)
          if (p.act == GTS_ACT_MASK_POS) GTS_EPI(GTS_ACT_MASK_POS);
          else if (p.act == GTS_ACT_MASK_BITS) GTS_EPI(GTS_ACT_MASK_BITS);
          else if (p.act == GTS_ACT_MASK_BITS_SCATTER) GTS_EPI(GTS_ACT_MASK_BITS_SCATTER);
          else if (p.act == GTS_ACT_RELU) GTS_EPI(GTS_ACT_RELU);
          else GTS_EPI(GTS_ACT_NONE);
#undef GTS_EPI
        } else if (fast_ok && m0 >= p.M) {
          // row group entirely beyond M (last tile): nothing to store, only the accumulator hand-back
          mbar_wait(&tmem_full[acc], full_phase);
          tcgen05_fence_after();
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) { if (rank == 0) mbar_arrive(&tmem_empty[acc]); else mbar_arrive_remote(tmem_empty_leader + acc * 8); }
        } else {
          // a float mask operand with its own leading dimension: generic chunk-wise path
          epilogue_tile<false>(p, t_base, m0, n0, p.C, stg, lane, 0, 32, masked, &tmem_full[acc], full_phase);
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) { if (rank == 0) mbar_arrive(&tmem_empty[acc]); else mbar_arrive_remote(tmem_empty_leader + acc * 8); }
        }
        if (GTS_TR_ON && ew == 0 && lane == 0) GTS_TR((int)(h >> 1), 10 + 3 * half, clock64());
      }
    }
  }

  tcgen05_fence_before();
  cluster_sync_all();                                // no CTA may exit (or free TMEM) while its peer still uses it
  if (threadIdx.x == 0) GTS_TR(7, 2, clock64());     // trace: all roles done
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

// 2-D fp32 tensor [rows, cols] with row stride ld (elements); box = box_cols x box_rows.
static bool encode_2d(CUtensorMap* tm, const float* base, int64_t rows, int64_t cols, int64_t ld, int box_cols, int box_rows,
                      bool round_tf32, bool swizzle_atom_32b = false) {
  EncodeTiledFn enc = get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled entry point unavailable"); return false; }
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(float)};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  static const bool no_round = getenv("GTS_TMA_NO_ROUND") != nullptr;
  if (no_round) round_tf32 = false;
  CUresult r = enc(tm, round_tf32 ? CU_TENSOR_MAP_DATA_TYPE_TFLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                   const_cast<float*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   swizzle_atom_32b ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r); return false; }
  return true;
}

static inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
static inline bool operand_ok(const float* p, int64_t ld) { return p && al16(p) && ld % 4 == 0 && ld > 0; }

// cudaFuncAttributeMaxDynamicSharedMemorySize is a PER-DEVICE attribute: remember it per device ordinal (a process
// that drives several GPUs, or switches the current device, must set it on each).
constexpr int kMaxDevices = 64;
template <typename Kernel>
static int ensure_dyn_smem(Kernel kernel, int bytes, bool (&done)[kMaxDevices]) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) dev = -1;
  if (dev < 0 || dev >= kMaxDevices || !done[dev]) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e != cudaSuccess) { set_error("cudaFuncSetAttribute(smem=%d) failed: %s", bytes, cudaGetErrorString(e)); return GTS_ERR_CUDA; }
    if (dev >= 0 && dev < kMaxDevices) done[dev] = true;
  }
  return GTS_OK;
}

template <bool TN, bool X3>
static int configure_kernel() {
  static bool done[kMaxDevices] = {};
  return ensure_dyn_smem(gemm_tc_kernel<TN, X3>, Cfg<X3>::dyn_bytes, done);
}

template <bool TN, bool X3>
static int launch(const CUtensorMap& a1, const CUtensorMap& a2, const CUtensorMap& b1, const CUtensorMap& b2,
                  const Params& p, int n_work, cudaStream_t st) {
  int rc = configure_kernel<TN, X3>();
  if (rc != GTS_OK) return rc;
  const int grid = n_work < sm_count() ? n_work : sm_count();
  gemm_tc_kernel<TN, X3><<<grid, NUM_THREADS, Cfg<X3>::dyn_bytes, st>>>(a1, a2, b1, b2, p);
  GTS_LAUNCH_CHECK();
  return GTS_OK;
}

template <bool TN, int BNC>
static int launch_ts(const CUtensorMap& a1, const CUtensorMap& a2, const CUtensorMap& b1, const CUtensorMap& b2,
                     const Params& p, int n_work, cudaStream_t st) {
  using L = TsCfg<BNC>;
  static bool done[kMaxDevices] = {};
  if (int rc = ensure_dyn_smem(gemm_x3ts_kernel<TN, BNC>, L::dyn_bytes, done)) return rc;
  const int grid = n_work < sm_count() ? n_work : sm_count();
  gemm_x3ts_kernel<TN, BNC><<<grid, L::THREADS, L::dyn_bytes, st>>>(a1, a2, b1, b2, p);
  GTS_LAUNCH_CHECK();
  return GTS_OK;
}

template <bool TN, bool DUAL = false, int BF = 0>
static int launch_ts2(const CUtensorMap& a1, const CUtensorMap& a2, const CUtensorMap& b1, const CUtensorMap& b2,
                      const Params& p, int n_work, cudaStream_t st) {
  using L = Ts2CfgT<DUAL, BF>;
  static bool done[kMaxDevices] = {};
  if (int rc = ensure_dyn_smem(gemm_x3ts2_kernel<TN, DUAL, BF>, L::dyn_bytes, done)) return rc;
  const int pairs = sm_count() / 2;
  const int grid = 2 * (n_work < pairs ? n_work : pairs);          // one CTA pair (cluster of 2) per TPC
  gemm_x3ts2_kernel<TN, DUAL, BF><<<grid, L::THREADS, L::dyn_bytes, st>>>(a1, a2, b1, b2, p);
  GTS_LAUNCH_CHECK();
  return GTS_OK;
}

static long long* g_ntw_trace = nullptr;      // GTS_TRACE builds: set through gts_debug_set_trace
#ifdef GTS_TRACE
extern "C" GTS_API void gts_debug_set_trace(void* device_ptr) { g_ntw_trace = reinterpret_cast<long long*>(device_ptr); }
static int g_dbg_flags_host = 0;
extern "C" GTS_API void gts_debug_set_flags(int flags) { g_dbg_flags_host = flags; }
#endif

static int launch_ntw(const CUtensorMap& a1, const CUtensorMap& a2, const CUtensorMap& b1, const CUtensorMap& b2,
                      const Params& p, int n_work, cudaStream_t st) {
  using L = NtwCfg;
  static bool done[kMaxDevices] = {};
  if (int rc = ensure_dyn_smem(gemm_x3ntw_kernel, L::dyn_bytes, done)) return rc;
  const int pairs = sm_count() / 2;
  const int grid = 2 * (n_work < pairs ? n_work : pairs);          // one CTA pair (cluster of 2) per TPC
  Params q = p;
  q.trace = g_ntw_trace;
#ifdef GTS_TRACE
  q.dbg = g_dbg_flags_host;
#endif
  gemm_x3ntw_kernel<<<grid, L::THREADS, L::dyn_bytes, st>>>(a1, a2, b1, b2, q);
  GTS_LAUNCH_CHECK();
  return GTS_OK;
}

// CTA-pair (cta_group::2) form for the wide shapes; GTS_X3_CTAS=1 keeps every shape on the one-CTA kernel
static bool ts_pair_enabled() {
  static const bool off = getenv("GTS_X3_CTAS") && atoi(getenv("GTS_X3_CTAS")) == 1;
  return !off;
}
static bool nt_pair_shape(int M, int N) { return ts_pair_enabled() && M >= 256 && N >= 128 && N % 64 == 0; }
static bool tn_pair_shape(int Mo, int No) { return ts_pair_enabled() && Mo > 128 && No >= 128 && No % 64 == 0; }

// bf16 cross terms in the weight-gradient (TN) pair kernels: 8 instead of 12 MMA-times per k-block (these kernels run
// at the tensor pipe's sustained rate, so the executed work is what counts).  Needs the full 128-wide N tile (64 n per
// CTA = one 128-byte bf16 row per k).  GTS_X3_TN_BF16=0 keeps the all-TF32 form.
static bool tn_bf_cross(int bn) {
  static const bool on = !(getenv("GTS_X3_TN_BF16") && atoi(getenv("GTS_X3_TN_BF16")) == 0);
  return on && bn == 128;
}

// widest N tile of the A-in-TMEM kernel: 128 (default: 4 stages, overlapped epilogue) or 256 (GTS_X3_BN=256)
static int ts_bn_cap() {
  static const int cap = (getenv("GTS_X3_BN") && atoi(getenv("GTS_X3_BN")) == 256) ? 256 : 128;
  return cap;
}

// A/B switch for profiling: GTS_X3_SMEM=1 keeps the all-shared-memory 3xTF32 kernel.
static bool x3_in_tmem() {
  static const bool legacy = getenv("GTS_X3_SMEM") != nullptr;
  return !legacy;
}

static int pick_bn(int n, int granule, int cap = MAX_BN) {
  int bn = n >= cap ? cap : ((n + granule - 1) / granule) * granule;
  return bn;
}

}  // namespace tc

// the shapes / modes the 256-wide kernel takes (mirrors the dispatch in gemm_nt_tcgen05)
bool gemm_nt_bits_supported(int32_t M, int32_t N, int32_t mode) {
  static const bool ntw_on = !(getenv("GTS_X3_NTW") && atoi(getenv("GTS_X3_NTW")) == 0);
  static const int bf_cross = getenv("GTS_X3_BF16") ? atoi(getenv("GTS_X3_BF16")) : 1;
  return mode == GTS_GEMM_TF32X3 && tc::x3_in_tmem() && tc::nt_pair_shape(M, N) && ntw_on && bf_cross == 1 && N % 256 == 0 &&
         tc::get_encode() != nullptr;
}

bool gemm_nt_tcgen05_supported(const gts_gemm_nt_args* a) {
  if (a->M < 1 || a->N < 4 || a->N % 4 != 0) return false;
  if (a->K1 < 1 || !tc::operand_ok(a->A1, a->lda1) || !tc::operand_ok(a->B1, a->ldb1)) return false;
  const bool two = a->A2 && a->B2 && a->K2 > 0;
  if (two && (!tc::operand_ok(a->A2, a->lda2) || !tc::operand_ok(a->B2, a->ldb2))) return false;
  if (a->act == GTS_ACT_MASK_POS_SCATTER) return false;     // scatter epilogue: SIMT kernel only (measured slower when fused into the tcgen05 epilogue)
  if (a->act != GTS_ACT_MASK_BITS_SCATTER && (!tc::al16(a->C) || a->ldc % 4 != 0)) return false;
  if (a->bias && !tc::al16(a->bias)) return false;
  if (a->bias2 && !tc::al16(a->bias2)) return false;
  if (a->act == GTS_ACT_MASK_POS && (!tc::al16(a->aux) || a->ldaux % 4 != 0)) return false;
  if ((a->act == GTS_ACT_MASK_BITS || a->act == GTS_ACT_MASK_BITS_SCATTER) &&
      (!a->aux_bits || !tc::al16(a->aux_bits) || a->ld_aux_bits % 4 != 0)) return false;
  if (a->act == GTS_ACT_MASK_BITS_SCATTER &&
      (!a->scatter_idx || !a->scatter_out || !tc::al16(a->scatter_idx) || a->ld_idx % 4 != 0 || a->ld_out * 4 >= ((int64_t)1 << 32)))
    return false;
  if (a->relu_bits_out && (a->act != GTS_ACT_RELU || a->ld_bits_out < a->N / 32)) return false;
  return tc::get_encode() != nullptr;
}

int gemm_nt_tcgen05(const gts_gemm_nt_args* a, cudaStream_t st) {
  using namespace tc;
  const bool x3 = a->mode == GTS_GEMM_TF32X3;
  const bool rnd = !x3;      // 1xTF32: TMA rounds to nearest; 3xTF32: raw fp32 (the MMA truncates, lo = x - trunc(x))
  const bool two = a->A2 && a->B2 && a->K2 > 0;
  const bool in_tmem = x3 && x3_in_tmem();
  const bool pair = in_tmem && nt_pair_shape(a->M, a->N);
  Params p{};
  p.BN = pair ? pick_bn(a->N, 64, Ts2Cfg::BN_MAX) : pick_bn(a->N, 16, in_tmem ? ts_bn_cap() : MAX_BN);
  p.tiles_m = pair ? (a->M + 2 * BM - 1) / (2 * BM) : (a->M + BM - 1) / BM;
  p.tiles_n = (a->N + p.BN - 1) / p.BN;
  // 256-wide output tiles: one pass over A feeds both 128-column halves (gemm_x3ntw_kernel; GTS_X3_NTW=0 keeps the
  // N = 128 pair kernel for A/B runs)
  static const bool ntw_on = !(getenv("GTS_X3_NTW") && atoi(getenv("GTS_X3_NTW")) == 0);
  static const int bf_cross = getenv("GTS_X3_BF16") ? atoi(getenv("GTS_X3_BF16")) : 1;    // bf16 cross terms: default on
  const bool wide = pair && ntw_on && bf_cross == 1 && a->N % 256 == 0;
  if (wide) p.tiles_n = a->N / 256;          // p.BN stays 128: the UMMA N of one half tile
  p.M = a->M; p.N = a->N; p.C = a->C; p.ldc = a->ldc;
  p.kb1 = (a->K1 + BK - 1) / BK;
  p.kb2 = two ? (a->K2 + BK - 1) / BK : 0;
  p.bias = a->bias; p.bias2 = a->bias2; p.aux = a->aux; p.ldaux = a->ldaux; p.act = a->act;
  p.bits_out = a->relu_bits_out; p.ld_bits_out = a->ld_bits_out; p.aux_bits = a->aux_bits; p.ld_aux_bits = a->ld_aux_bits;
  p.scatter_idx = a->scatter_idx; p.ld_idx = a->ld_idx; p.scatter_out = a->scatter_out; p.ld_out = a->ld_out;
  // side job (gts_gemm_nt_args.zero_fill): only the 256-wide kernel has spare warps; gts_gemm_nt issued a memset otherwise
  if (wide && a->zero_fill) { p.zero_fill = reinterpret_cast<uint4*>(a->zero_fill); p.zero_n16 = a->zero_fill_bytes >> 4; }
  p.splits = 1;
  CUtensorMap tA1, tA2, tB1, tB2;
  if (!encode_2d(&tA1, a->A1, a->M, a->K1, a->lda1, BK, BM, rnd)) return GTS_ERR_CUDA;
  const int b_rows = pair ? p.BN / 2 : p.BN;        // a CTA pair lands half of the N rows per CTA
  if (!encode_2d(&tB1, a->B1, a->N, a->K1, a->ldb1, BK, b_rows, rnd)) return GTS_ERR_CUDA;
  if (two) {
    if (!encode_2d(&tA2, a->A2, a->M, a->K2, a->lda2, BK, BM, rnd)) return GTS_ERR_CUDA;
    if (!encode_2d(&tB2, a->B2, a->N, a->K2, a->ldb2, BK, b_rows, rnd)) return GTS_ERR_CUDA;
  } else {
    tA2 = tA1; tB2 = tB1;
  }
  const int n_work = p.tiles_m * p.tiles_n;
  // cross terms of the 3xTF32 scheme as bf16 MMAs (8 instead of 12 MMAs per k-block; GTS_X3_BF16=0: all-TF32 form).
  // Measured: same error against fp64 (2.7e-6 max on K=256 products, logits 2.6e-5 on the 8-layer stack), K=256
  // 59.8 -> 55.9 us, K=512 100.6 -> 98.3 us, training step 5.19 -> 5.10 ms.
  if (wide) return launch_ntw(tA1, tA2, tB1, tB2, p, n_work, st);
  if (a->relu_bits_out || a->act == GTS_ACT_MASK_BITS || a->act == GTS_ACT_MASK_BITS_SCATTER) {
    set_error("gts_gemm_nt: bit-matrix masks need the 256-wide CTA-pair kernel (gts_gemm_nt_bits_supported)");
    return GTS_ERR_UNSUPPORTED;
  }
  if (pair && bf_cross == 2) return launch_ts2<false, false, 2>(tA1, tA2, tB1, tB2, p, n_work, st);
  if (pair && bf_cross == 1) return launch_ts2<false, false, 1>(tA1, tA2, tB1, tB2, p, n_work, st);
  if (pair) return launch_ts2<false>(tA1, tA2, tB1, tB2, p, n_work, st);
  if (in_tmem)
    return ts_bn_cap() == 256 ? launch_ts<false, 256>(tA1, tA2, tB1, tB2, p, n_work, st)
                              : launch_ts<false, 128>(tA1, tA2, tB1, tB2, p, n_work, st);
  return x3 ? launch<false, true>(tA1, tA2, tB1, tB2, p, n_work, st)
            : launch<false, false>(tA1, tA2, tB1, tB2, p, n_work, st);
}

bool gemm_tn_tcgen05_supported(const float* A, int64_t lda, const float* B, int64_t ldb, int32_t Mo, int32_t No, int64_t K) {
  if (Mo < 4 || Mo % 4 != 0 || No < 4 || No % 4 != 0 || K < 1 || K > 0x7fffffff) return false;
  if (!tc::operand_ok(A, lda) || !tc::operand_ok(B, ldb)) return false;
  return tc::get_encode() != nullptr;
}

constexpr int kMaxKbPerSplit = 128;   // k-blocks of 32 rows
static void tn_plan(int32_t Mo, int32_t No, int64_t K, int32_t mode, tc::Params& p) {
  using namespace tc;
  const bool in_tmem = mode == GTS_GEMM_TF32X3 && x3_in_tmem();
  const bool pair = in_tmem && tn_pair_shape(Mo, No);
  p.BN = pair ? pick_bn(No, 64, Ts2Cfg::BN_MAX) : pick_bn(No, 32, in_tmem ? ts_bn_cap() : MAX_BN);
  p.tiles_m = pair ? (Mo + 2 * BM - 1) / (2 * BM) : (Mo + BM - 1) / BM;
  p.tiles_n = (No + p.BN - 1) / p.BN;
  p.kb_total = (int)((K + BK - 1) / BK);
  const int tiles = p.tiles_m * p.tiles_n;
  int splits = (pair ? sm_count() / 2 : sm_count()) / tiles;
  if (splits < 1) splits = 1;
  if (splits > p.kb_total) splits = p.kb_total;
  p.kb_per_split = (p.kb_total + splits - 1) / splits;
  // accuracy bound: the tensor core's fp32 accumulation error grows with the length of one accumulation chain
  // (measured: 3.7e-4 for 45 000 node rows per split, 1e-6 for 2 400); cap a split at 4096 rows and let the
  // deterministic fp32 reduction add the partial products instead.
  if (p.kb_per_split > kMaxKbPerSplit) p.kb_per_split = kMaxKbPerSplit;
  p.splits = (p.kb_total + p.kb_per_split - 1) / p.kb_per_split;
  p.split_stride = (int64_t)Mo * No;
}

// workspace: per split one record [Mo*No partial product | Mo (rounded up to 4) partial column sums of A]
static inline int64_t tn_record(int32_t Mo, int32_t No) { return (int64_t)Mo * No + (((int64_t)Mo + 3) / 4) * 4; }

// A 256-wide product as the two-B form over its two 128-column halves (B1 = B[:, 0:128], B2 = B[:, 128:256]): the split A
// tile in tensor memory feeds both halves, so A is read once instead of once per N tile, and every cluster owns one
// K split of the whole product (74 splits instead of 2 x 37).  GTS_X3_TN_HALVES=0 keeps one work item per N tile.
static bool tn_split_halves(int32_t Mo, int32_t No, int32_t mode) {
  static const bool on = !(getenv("GTS_X3_TN_HALVES") && atoi(getenv("GTS_X3_TN_HALVES")) == 0);
  return on && mode == GTS_GEMM_TF32X3 && tc::x3_in_tmem() && No == 256 && tc::tn_pair_shape(Mo, 128) && tc::tn_bf_cross(128);
}

size_t gemm_tn_tcgen05_ws(int32_t Mo, int32_t No, int64_t K, int32_t mode) {
  if (Mo < 1 || No < 1 || K < 1) return 0;
  tc::Params p{};
  tn_plan(Mo, tn_split_halves(Mo, No, mode) ? 128 : No, K, mode, p);
  return align_up((size_t)p.splits * (size_t)tn_record(Mo, No) * sizeof(float), 256);
}

// true when gemm_tn_tcgen05 can produce the column sums of A as a by-product (3xTF32, A staged through TMEM)
bool gemm_tn_tcgen05_fuses_colsum(int32_t mode) { return mode == GTS_GEMM_TF32X3 && tc::x3_in_tmem(); }

int gemm_tn_tcgen05(const float* A, int64_t lda, const float* B, int64_t ldb, float* C, int64_t ldc,
                    int32_t Mo, int32_t No, int64_t K, int32_t mode, float* colsum_out, void* ws, size_t ws_bytes,
                    cudaStream_t st) {
  using namespace tc;
  const bool x3 = mode == GTS_GEMM_TF32X3;
  const bool halves = tn_split_halves(Mo, No, mode);
  Params p{};
  tn_plan(Mo, halves ? 128 : No, K, mode, p);
  const int64_t record = tn_record(Mo, No);
  const size_t need = align_up((size_t)p.splits * (size_t)record * sizeof(float), 256);
  if (!ws || ws_bytes < need) { set_error("gts_gemm_tn: workspace %zu < required %zu", ws_bytes, need); return GTS_ERR_WORKSPACE; }
  p.M = Mo; p.N = halves ? 128 : No;
  p.C = reinterpret_cast<float*>(ws); p.ldc = No;
  p.split_stride = record;
  if (halves) {
    // record = [Mo x 256 product | column sums of A]: the second half tile lands 128 columns to the right
    p.c2_off = 128;
    p.colsum_partial = colsum_out ? p.C + (int64_t)Mo * No : nullptr;
    CUtensorMap tA, tB1, tB2;
    if (!encode_2d(&tA, A, K, Mo, lda, 32, BK, false, true)) return GTS_ERR_CUDA;
    if (!encode_2d(&tB1, B, K, 128, ldb, 32, BK, false, true)) return GTS_ERR_CUDA;
    if (!encode_2d(&tB2, B + 128, K, 128, ldb, 32, BK, false, true)) return GTS_ERR_CUDA;
    const int n_work_h = p.tiles_m * p.tiles_n * p.splits;
    int rc_h = launch_ts2<true, true, 1>(tA, tA, tB1, tB2, p, n_work_h, st);
    if (rc_h != GTS_OK) return rc_h;
    if (colsum_out && Mo % 4 == 0 && p.splits >= 8 && splitk_reduce_fused_ok(Mo, No, ldc, Mo, record, p.C, C, colsum_out)) {
      launch_splitk_reduce_fused(p.C, record, p.splits, Mo, No, C, colsum_out, Mo, st);
      GTS_LAUNCH_CHECK();
      return GTS_OK;
    }
    launch_splitk_reduce(p.C, record, p.splits, Mo, No, C, ldc, st);
    GTS_LAUNCH_CHECK();
    if (colsum_out) {
      launch_splitk_reduce(p.C + (int64_t)Mo * No, record, p.splits, 1, Mo, colsum_out, Mo, st);
      GTS_LAUNCH_CHECK();
    }
    return GTS_OK;
  }
  const bool in_tmem = x3 && x3_in_tmem();
  if (colsum_out && !in_tmem) { set_error("gts_gemm_tn: fused column sums need the 3xTF32 A-in-TMEM kernel"); return GTS_ERR_INVALID; }
  float* cs_partial = p.C + (int64_t)Mo * No;
  p.colsum_partial = colsum_out ? cs_partial : nullptr;
  CUtensorMap tA, tB;
  if (!encode_2d(&tA, A, K, Mo, lda, 32, BK, !x3, true)) return GTS_ERR_CUDA;
  if (!encode_2d(&tB, B, K, No, ldb, 32, BK, !x3, true)) return GTS_ERR_CUDA;
  const int n_work = p.tiles_m * p.tiles_n * p.splits;
  const bool pair = in_tmem && tn_pair_shape(Mo, No);
  int rc = pair ? (tn_bf_cross(p.BN) ? launch_ts2<true, false, 1>(tA, tA, tB, tB, p, n_work, st)
                                      : launch_ts2<true>(tA, tA, tB, tB, p, n_work, st))
         : in_tmem ? (ts_bn_cap() == 256 ? launch_ts<true, 256>(tA, tA, tB, tB, p, n_work, st)
                                         : launch_ts<true, 128>(tA, tA, tB, tB, p, n_work, st))
                   : (x3 ? launch<true, true>(tA, tA, tB, tB, p, n_work, st) : launch<true, false>(tA, tA, tB, tB, p, n_work, st));
  if (rc != GTS_OK) return rc;
  // one deterministic reduction over the splits for the product and (when fused) the column sums
  if (colsum_out && Mo % 4 == 0 && p.splits >= 8 &&
      splitk_reduce_fused_ok(Mo, No, ldc, Mo, record, p.C, C, colsum_out)) {
    launch_splitk_reduce_fused(p.C, record, p.splits, Mo, No, C, colsum_out, Mo, st);
    GTS_LAUNCH_CHECK();
    return GTS_OK;
  }
  launch_splitk_reduce(p.C, record, p.splits, Mo, No, C, ldc, st);
  GTS_LAUNCH_CHECK();
  if (colsum_out) {
    launch_splitk_reduce(cs_partial, record, p.splits, 1, Mo, colsum_out, Mo, st);
    GTS_LAUNCH_CHECK();
  }
  return GTS_OK;
}

// [C1 | C2] = A^T [B1 | B2] (+ column sums of A) in one pass over A; CTA-pair 3xTF32 kernel only.
bool gemm_tn2_tcgen05_supported(const float* A, int64_t lda, const float* B1, int64_t ldb1, const float* B2, int64_t ldb2,
                                int32_t Mo, int32_t No, int64_t K, int32_t mode) {
  return mode == GTS_GEMM_TF32X3 && tc::x3_in_tmem() && tc::tn_pair_shape(Mo, No) && Mo % 4 == 0 &&
         gemm_tn_tcgen05_supported(A, lda, B1, ldb1, Mo, No, K) && tc::operand_ok(B2, ldb2);
}
size_t gemm_tn2_tcgen05_ws(int32_t Mo, int32_t No, int64_t K) {
  if (Mo < 1 || No < 1 || K < 1) return 0;
  tc::Params p{};
  tn_plan(Mo, No, K, GTS_GEMM_TF32X3, p);
  return align_up((size_t)p.splits * (size_t)(tn_record(Mo, No) + (int64_t)Mo * No) * sizeof(float), 256);
}
int gemm_tn2_tcgen05(const float* A, int64_t lda, const float* B1, int64_t ldb1, const float* B2, int64_t ldb2,
                     float* C1, float* C2, int64_t ldc, int32_t Mo, int32_t No, int64_t K, float* colsum_out,
                     void* ws, size_t ws_bytes, cudaStream_t st) {
  using namespace tc;
  Params p{};
  tn_plan(Mo, No, K, GTS_GEMM_TF32X3, p);
  const int64_t prod = (int64_t)Mo * No;
  const int64_t record = tn_record(Mo, No) + prod;               // [C1 | C2 | column sums]
  const size_t need = align_up((size_t)p.splits * (size_t)record * sizeof(float), 256);
  if (!ws || ws_bytes < need) { set_error("gts_gemm_tn2: workspace %zu < required %zu", ws_bytes, need); return GTS_ERR_WORKSPACE; }
  p.M = Mo; p.N = No;
  p.C = reinterpret_cast<float*>(ws); p.ldc = No;
  p.split_stride = record;
  p.c2_off = prod;
  float* cs_partial = p.C + 2 * prod;
  p.colsum_partial = colsum_out ? cs_partial : nullptr;
  CUtensorMap tA, tB1, tB2;
  if (!encode_2d(&tA, A, K, Mo, lda, 32, BK, false, true)) return GTS_ERR_CUDA;
  if (!encode_2d(&tB1, B1, K, No, ldb1, 32, BK, false, true)) return GTS_ERR_CUDA;
  if (!encode_2d(&tB2, B2, K, No, ldb2, 32, BK, false, true)) return GTS_ERR_CUDA;
  const int n_work = p.tiles_m * p.tiles_n * p.splits;
  int rc = tn_bf_cross(p.BN) ? launch_ts2<true, true, 1>(tA, tA, tB1, tB2, p, n_work, st)
                             : launch_ts2<true, true>(tA, tA, tB1, tB2, p, n_work, st);
  if (rc != GTS_OK) return rc;
  const bool contiguous = C2 == C1 + prod && ldc == No;
  if (contiguous && colsum_out && p.splits >= 8 && splitk_reduce_fused_ok(2 * Mo, No, ldc, Mo, record, p.C, C1, colsum_out)) {
    launch_splitk_reduce_fused(p.C, record, p.splits, 2 * Mo, No, C1, colsum_out, Mo, st);   // one reduction for all three
    GTS_LAUNCH_CHECK();
    return GTS_OK;
  }
  launch_splitk_reduce(p.C, record, p.splits, Mo, No, C1, ldc, st);
  launch_splitk_reduce(p.C + prod, record, p.splits, Mo, No, C2, ldc, st);
  if (colsum_out) launch_splitk_reduce(cs_partial, record, p.splits, 1, Mo, colsum_out, Mo, st);
  GTS_LAUNCH_CHECK();
  return GTS_OK;
}

}  // namespace gts
